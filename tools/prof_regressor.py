"""Small driver for ncu: one regressor forward+backward on `--images` images (10 crops each), bf16 tcgen05 path."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import oracle as O
from regressor_guided_image_editing_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=16)
ap.add_argument("--iters", type=int, default=1)
args = ap.parse_args()
sd = O.make_regressor_state_dict()
B = args.images
reg = ops.Regressor(sd, max_crops=B * 10, precision="bf16")
img = torch.rand(B, 3, 480, 480, device="cuda")
offs = torch.randint(0, 33, (B, 10, 2), dtype=torch.int32, device="cuda")
dl = torch.randn(B * 10, 4, device="cuda") * 1e-3
dimg = torch.empty_like(img)
for _ in range(args.iters):
    logits = reg.forward(img, offs)
    reg.backward(dl, dimg)
torch.cuda.synchronize()
print("ok", logits.float().abs().mean().item(), dimg.abs().mean().item())

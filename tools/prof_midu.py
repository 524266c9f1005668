"""Timing of the MiDU guidance head (configs[3] shape: batch 32 mid-block features) -- forward + d/d(feature), CUDA events,
median of --reps.  SD head: [32,1280,8,8]; SDXL head: [32,1280,32,32].  FLOPs: 2*M*N*K of the convolutions (fwd + dgrad)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from regressor_guided_image_editing_b200.guidance_classifier.MiduClassifier import NativeMiduHead, _MiduHeadFn

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--reps", type=int, default=30)
a = ap.parse_args()
for sdxl, hw, chans in ((False, 8, [(1280, 256, 8), (256, 128, 4)]), (True, 32, [(1280, 512, 32), (512, 256, 16), (256, 128, 8), (128, 64, 4)])):
    m = O._midu_module(sdxl, 2).cuda()
    flops = 2 * sum(2 * a.batch * h * h * co * 9 * ci for ci, co, h in chans)      # fwd + dgrad
    for prec in ("bf16", "fp32"):
        head = NativeMiduHead(m, prec, sdxl)
        f = torch.randn(a.batch, 1280, hw, hw, device="cuda", requires_grad=True)
        ts = []
        for _ in range(a.reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            p = _MiduHeadFn.apply(f, head)
            g, = torch.autograd.grad(p.square().sum(), f)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = sorted(ts[3:])
        t = ts[len(ts) // 2]
        print(f"{'SDXL' if sdxl else 'SD  '} head {prec}: batch {a.batch}, fwd+bwd {t:.3f} ms, {flops / t / 1e9:.1f} TFLOP/s (convs only)")

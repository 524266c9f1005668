"""Diagnostic (not a test): how well conditioned is d(objective)/d(style) of the latent loop's GENERATOR terms?
Compares the gradient of the content-reconstruction term and of a smooth stand-in for the classifier term through the
random-init MUNIT mirror in fp32 vs fp64 on the CPU, and -- when a GPU is present -- CPU fp32 vs GPU fp32 under the
TF32 switches PyTorch offers.  Usage: python tools/diag_latent_conditioning.py [batch]"""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from regressor_guided_image_editing_b200.external.imaginaire.generators.munit import Generator

torch.manual_seed(0)
gen = Generator()
B, h = (int(sys.argv[1]) if len(sys.argv) > 1 else 2), 256
image = torch.stack([2.0 * O.synthetic_image(300 + i, h, h) - 1.0 for i in range(B)])


def grads(g, img, dt, dev="cpu"):
    g = copy.deepcopy(g).to(dev).to(dt)
    img = img.to(dev).to(dt)
    with torch.no_grad():
        content, style = g.autoencoder_a.encode(img)
    x = style.clone().requires_grad_(True)
    im = torch.clamp(g.autoencoder_a.decode(content, x), -1, 1)
    rec = torch.nn.functional.l1_loss(g.autoencoder_a.encode(im)[0], content)
    gr, = torch.autograd.grad(rec, x, retain_graph=True)
    torch.manual_seed(5)
    w = torch.randn(im.shape[1:]).to(dev).to(dt)
    gc, = torch.autograd.grad(0.2 * ((im * w).mean((1, 2, 3)) ** 2).mean(), x)
    return gr.flatten().double().cpu(), gc.flatten().double().cpu(), rec.item()


def report(tag, a, b):
    (ra, ca, la), (rb, cb, lb) = a, b
    print(f"{tag}: recon loss {la:.9f} / {lb:.9f};  recon grad |max| {rb.abs().max().item():.2e}, max diff {(ra - rb).abs().max().item():.2e} "
          f"= {((ra - rb).abs().max() / rb.abs().max()).item():.2e} of max;  smooth-term grad diff {((ca - cb).abs().max() / cb.abs().max()).item():.2e} of max")


t = time.time()
c32, c64 = grads(gen, image, torch.float32), grads(gen, image, torch.float64)
print(f"CPU passes: {time.time() - t:.1f} s (batch {B})")
report("CPU fp32 vs CPU fp64", c32, c64)
if torch.cuda.is_available():
    print("cudnn.allow_tf32 =", torch.backends.cudnn.allow_tf32, " cuda.matmul.allow_tf32 =", torch.backends.cuda.matmul.allow_tf32)
    report("GPU fp32 (torch defaults) vs CPU fp64", grads(gen, image, torch.float32, "cuda"), c64)
    torch.backends.cudnn.allow_tf32 = False
    report("GPU fp32 (cudnn.allow_tf32=False) vs CPU fp64", grads(gen, image, torch.float32, "cuda"), c64)
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.backends.fp32_precision = "ieee"
        torch.backends.cudnn.conv.fp32_precision = "ieee"
        print("fp32_precision API present: set to ieee")
    except Exception as e:
        print("fp32_precision API:", repr(e))
    report("GPU fp32 (all TF32 off) vs CPU fp64", grads(gen, image, torch.float32, "cuda"), c64)
    torch.backends.cudnn.deterministic = True
    report("GPU fp32 (all TF32 off, deterministic) vs CPU fp64", grads(gen, image, torch.float32, "cuda"), c64)
    report("GPU fp64 vs CPU fp64", grads(gen, image, torch.float64, "cuda"), c64)

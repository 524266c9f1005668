"""GPU: antialiased resize (+transpose), loss head, Adam/best-x and guidance update kernels against torch / the oracle."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(256, 256, 480, 480), (512, 512, 480, 480), (300, 400, 480, 640), (600, 480, 600, 480),
                                   (97, 131, 480, 648)])
def test_resize_aa_fwd_bwd(shape):
    from regressor_guided_image_editing_b200 import ops
    ih, iw, oh, ow = shape
    x = torch.rand(2, 3, ih, iw, generator=torch.Generator().manual_seed(1))
    xc = x.clone().requires_grad_(True)
    ref = F.interpolate(xc, size=(oh, ow), mode="bilinear", align_corners=False, antialias=True)
    gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(2))
    gref, = torch.autograd.grad((ref * gout).sum(), xc)
    rs = ops.Resize(ih, iw, oh, ow)
    out = rs.fwd(x.to(DEV))
    gin = rs.bwd(gout.to(DEV).contiguous())
    assert (out.cpu() - ref.detach()).abs().max().item() <= 2e-6
    assert (gin.cpu() - gref).abs().max().item() <= 2e-5


def test_fused_resize_is_bit_identical_to_the_two_pass_kernels(tmp_path):
    """The single-kernel resize keeps the tap order and the fp32 rounding of the intermediate, so it must reproduce the two
    separable kernels (RGIE_RESIZE_FUSED=0, read once per process -> child process) bit for bit, forward and transpose."""
    import os, subprocess, sys
    from regressor_guided_image_editing_b200 import ops
    code = (
        "import sys, torch\n"
        "sys.path.insert(0, %r)\n"
        "from regressor_guided_image_editing_b200 import ops\n"
        "x = torch.rand(3, 3, 512, 512, generator=torch.Generator().manual_seed(11)).cuda()\n"
        "g = torch.randn(3, 3, 480, 480, generator=torch.Generator().manual_seed(12)).cuda()\n"
        "rs = ops.Resize(512, 512, 480, 480)\n"
        "torch.save({'out': rs.fwd(x).cpu(), 'gin': rs.bwd(g).cpu()}, %r)\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), str(tmp_path / "two_pass.pt"))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, RGIE_RESIZE_FUSED="0"), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    want = torch.load(tmp_path / "two_pass.pt")
    x = torch.rand(3, 3, 512, 512, generator=torch.Generator().manual_seed(11)).to(DEV)
    g = torch.randn(3, 3, 480, 480, generator=torch.Generator().manual_seed(12)).to(DEV)
    rs = ops.Resize(512, 512, 480, 480)
    assert torch.equal(rs.fwd(x).cpu(), want["out"])
    assert torch.equal(rs.bwd(g).cpu(), want["gin"])


def test_resize_output_size_rule():
    from regressor_guided_image_editing_b200 import ops
    for h, w in [(256, 256), (512, 512), (300, 400), (400, 300), (1024, 683)]:
        assert ops.resize_output_size(h, w, 480) == O.resize_output_size(h, w, 480)


def test_va_head_matches_oracle():
    from regressor_guided_image_editing_b200 import ops
    B, reps = 5, 10
    logits = torch.randn(B * reps, 4, generator=torch.Generator().manual_seed(3))
    target = torch.rand(B, 2, generator=torch.Generator().manual_seed(4))
    lg = logits.clone().requires_grad_(True)
    pred = torch.sigmoid(lg.view(B, reps, 4).mean(1))
    losses = torch.stack([0.15 * O.va_loss(pred[b:b + 1, :2], target[b:b + 1], 1.0) for b in range(B)])
    g, = torch.autograd.grad(losses.sum(), lg)
    preds_d = torch.empty(B, 4, device=DEV); loss_d = torch.empty(B, device=DEV); dl = torch.empty(B * reps, 4, device=DEV)
    ops.va_head(logits.to(DEV), B, reps, True, target.to(DEV), 0.5, 0.0, 3, 0.15, preds_d, loss_d, dl)
    assert (preds_d.cpu() - pred.detach()).abs().max().item() <= 1e-6
    assert (loss_d.cpu() - losses.detach()).abs().max().item() <= 1e-7
    assert (dl.cpu() - g).abs().max().item() <= 1e-8 + 1e-5 * g.abs().max().item()
    # untargeted defaults (ValenceArousalLoss.py:82-109): valence -> 0.5, arousal -> 0 when minimised
    ops.va_head(logits.to(DEV), B, reps, True, None, 0.5, 0.0, 3, 1.0, preds_d, loss_d, dl)
    ref = torch.stack([O.va_loss(pred[b:b + 1, :2].detach(), None, 1.0, True) for b in range(B)])
    assert (loss_d.cpu() - ref).abs().max().item() <= 1e-6


def test_adam_and_best_tracking_match_torch():
    from regressor_guided_image_editing_b200 import ops
    B, n, steps = 3, 41, 12
    gen = torch.Generator().manual_seed(5)
    x0 = torch.randn(B, n, generator=gen)
    grads = torch.randn(steps, B, n, generator=gen) * torch.logspace(-6, 0, n)
    grads[:, :, 7] = 0.0                                   # an element whose gradient is identically zero never moves
    losses = torch.rand(steps, B, generator=gen)
    # torch reference: one Adam per problem, lr rewritten each step as in optimize_image.py:69-75
    xr = [x0[b].clone().requires_grad_(True) for b in range(B)]
    opts = [torch.optim.Adam([xr[b]], betas=(0.9, 0.999), lr=0.05) for b in range(B)]
    best_loss = [math.inf] * B
    best_x = [x0[b].clone() for b in range(B)]
    for s in range(steps):
        lr = O.lr_schedule(s, steps, 0.05)
        for b in range(B):
            for gpar in opts[b].param_groups:
                gpar["lr"] = lr
            if losses[s, b].item() < best_loss[b]:
                best_loss[b] = losses[s, b].item(); best_x[b] = xr[b].detach().clone()
            xr[b].grad = grads[s, b].clone()
            opts[b].step()
    xd = x0.to(DEV).clone(); m = torch.zeros_like(xd); v = torch.zeros_like(xd)
    bl = torch.full((B,), float("inf"), device=DEV); bx = xd.clone(); bs = torch.zeros(B, dtype=torch.int32, device=DEV)
    for s in range(steps):
        ops.adam_step(xd, grads[s].to(DEV).contiguous(), m, v, O.lr_schedule(s, steps, 0.05), s + 1,
                      loss=losses[s].to(DEV).contiguous(), best_loss=bl, best_x=bx, best_step=bs, step=s)
    torch.cuda.synchronize()
    xref = torch.stack([t.detach() for t in xr])
    assert (xd.cpu() - xref).abs().max().item() <= 1e-6
    assert (bx.cpu() - torch.stack(best_x)).abs().max().item() <= 1e-6
    assert torch.allclose(bl.cpu(), torch.tensor(best_loss))
    assert (xd.cpu()[:, 7] == x0[:, 7]).all()


def test_guidance_update_matches_reference_lines():
    from regressor_guided_image_editing_b200 import ops
    gen = torch.Generator().manual_seed(6)
    lat = torch.randn(1, 4, 64, 64, generator=gen); g = torch.randn(1, 4, 64, 64, generator=gen) * 1e-3
    ref = O.guidance_update(lat, g, 0.2, True)
    out = ops.guidance_update(lat.to(DEV).clone(), g.to(DEV).contiguous(), 0.2, True)
    assert (out.cpu() - ref).abs().max().item() <= 1e-6
    # batched: one independent problem per latent (reference semantics at its batch size of 1)
    lat = torch.randn(32, 4, 64, 64, generator=gen); g = torch.randn(32, 4, 64, 64, generator=gen)
    ref = torch.cat([O.guidance_update(lat[i:i + 1], g[i:i + 1], 0.2, True) for i in range(32)])
    out = ops.guidance_update(lat.to(DEV).clone(), g.to(DEV).contiguous(), 0.2, True, per_problem=4 * 64 * 64)
    assert (out.cpu() - ref).abs().max().item() <= 1e-6

"""GPU: native resnet50 regressor (forward + input-gradient backward) against the CPU oracle.

fp32 mode (CUDA-core GEMMs) is the parity mode: activations, logits and d(image) must match the oracle to fp32
round-off.  bf16 mode (tcgen05/TMEM/TMA GEMMs) is checked against the bf16 CUDA-core cross-check (same storage
precision, different accumulation order) and, loosely, against the fp32 oracle.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

TAPS = ["stem", "pool", "layer1.0.c1", "layer1.0.c2", "layer1.0", "layer1.2", "layer2.0.c1", "layer2.0.c2", "layer2.0",
        "layer2.3", "layer3.0", "layer3.5", "layer4.0.c2", "layer4.0", "layer4.2"]


@pytest.fixture(scope="module")
def sd():
    return O.make_regressor_state_dict()


def _inputs(reps, hr=480, wr=480, seed=5):
    img = O.synthetic_image(seed, hr, wr)[None]
    g = torch.Generator().manual_seed(99)
    offs = torch.stack([torch.randint(0, hr - 448 + 1, (1, reps), generator=g),
                        torch.randint(0, wr - 448 + 1, (1, reps), generator=g)], -1).int()
    return img, offs


def _oracle(sd, img, offs, dlogits):
    img = img.clone().requires_grad_(True)
    crops = O.replicate_and_crop(img, offs, 448, True)
    taps = {}
    logits = O.resnet50_forward(crops, sd, taps)
    g, = torch.autograd.grad((logits * dlogits).sum(), img)
    return logits.detach(), {k: v.detach() for k, v in taps.items()}, g


def _native(sd, img, offs, dlogits, precision, want_taps=True):
    from regressor_guided_image_editing_b200 import ops
    reps = offs.shape[1]
    reg = ops.Regressor(sd, max_crops=reps, precision=precision)
    img_d, offs_d = img.to(DEV).contiguous(), offs.to(DEV).contiguous()
    logits = reg.forward(img_d, offs_d, normalize=True)
    taps = {}
    if want_taps:
        shapes = {"stem": (64, 224), "pool": (64, 112)}
        for name in TAPS:
            if name in shapes:
                c, h = shapes[name]
            else:
                s = int(name[5])
                cm = 64 << (s - 1)
                h = 224 >> s
                if name.endswith(".c1"):
                    c = cm
                    h = h * 2 if (name[7] == "0" and s > 1) else h
                elif name.endswith(".c2"):
                    c = cm
                else:
                    c = 4 * cm
            taps[name] = reg.tap(name, (reps, c, h, h)).cpu()
    dimg = torch.empty_like(img_d)
    reg.backward(dlogits.to(DEV).contiguous(), dimg)
    torch.cuda.synchronize()
    return logits.cpu(), taps, dimg.cpu()


def test_fp32_mode_matches_oracle(sd):
    reps = 2
    img, offs = _inputs(reps)
    dlogits = torch.randn(reps, 4, generator=torch.Generator().manual_seed(3))
    lo, to, go = _oracle(sd, img, offs, dlogits)
    ln, tn, gn = _native(sd, img, offs, dlogits, "fp32")
    for name in TAPS:
        ref = to[name]
        err = (tn[name] - ref).abs().max().item()
        scale = ref.abs().max().item()
        print(f"{name:14s} max-abs err {err:.3e} (max |ref| {scale:.3f})")
        assert err <= 2e-4 * max(scale, 1.0), f"{name}: {err}"
    lerr = (ln - lo).abs().max().item()
    print("logits", ln.flatten()[:4], lo.flatten()[:4], lerr)
    assert lerr <= 2e-4 * max(lo.abs().max().item(), 1.0)
    gerr = (gn - go).abs().max().item() / go.abs().max().item()
    grel = (gn - go).abs().mean().item() / go.abs().mean().item()
    print(f"d(image): max-rel {gerr:.3e} mean-rel {grel:.3e}")
    # isolated pixels sit on ReLU / max-pool kinks whose side is decided by fp32 round-off: bound the max loosely,
    # the mean tightly
    assert gerr <= 5e-2 and grel <= 4e-3


def test_bf16_tcgen05_matches_bf16_simt_and_oracle(sd):
    reps = 2
    img, offs = _inputs(reps, seed=6)
    dlogits = torch.randn(reps, 4, generator=torch.Generator().manual_seed(4))
    lo, to, go = _oracle(sd, img, offs, dlogits)
    ls, ts, gs = _native(sd, img, offs, dlogits, "bf16_simt")
    lt, tt, gt = _native(sd, img, offs, dlogits, "bf16")
    for name in TAPS:
        ref = to[name]
        scale = max(ref.abs().max().item(), 1.0)
        e_ts = (tt[name] - ts[name]).abs().max().item() / scale
        e_to = (tt[name] - ref).abs().max().item() / scale
        print(f"{name:14s} tcgen05 vs simt-bf16 {e_ts:.3e}   tcgen05 vs fp32 oracle {e_to:.3e}")
        assert e_ts <= 3e-2, f"{name}: tcgen05 vs CUDA-core bf16 {e_ts}"
        assert e_to <= 6e-2, f"{name}: tcgen05 vs oracle {e_to}"
    print("logits tcgen05", lt.flatten()[:4], "simt", ls.flatten()[:4], "oracle", lo.flatten()[:4])
    assert (lt - lo).abs().max().item() <= 5e-2 * max(lo.abs().max().item(), 1.0)
    assert (lt - ls).abs().max().item() <= 2e-2 * max(lo.abs().max().item(), 1.0)
    cos = F.cosine_similarity(gt.flatten(), go.flatten(), dim=0).item()
    cos_s = F.cosine_similarity(gt.flatten(), gs.flatten(), dim=0).item()
    print(f"d(image) cosine: tcgen05 vs oracle {cos:.5f}, tcgen05 vs simt-bf16 {cos_s:.5f}; "
          f"norm ratio {gt.norm().item() / go.norm().item():.4f}")
    assert cos >= 0.95 and cos_s >= 0.97      # per-pixel gradient through ~50 bf16 layers; parameter gradients average this out
    assert abs(gt.norm().item() / go.norm().item() - 1.0) <= 0.05


def test_regressor_golden(sd, golden_dir):
    """Reference-generated golden (ValenceArousalLoss forward + backward on CPU) through the native fp32 path."""
    from regressor_guided_image_editing_b200 import ops
    gold = torch.load(os.path.join(golden_dir, "regressor.pt"))
    for k, v in gold["weights_probe"].items():
        assert torch.equal(sd[k].flatten()[:4], v), "seeded state_dict differs from the one the golden was made with"
    for tag in ("up256", "down512", "same480", "rect300x400"):
        g = gold[tag]
        img = O.synthetic_image(g["image_index"], g["h"], g["w"])[None].to(DEV).contiguous()
        oh, ow = ops.resize_output_size(g["h"], g["w"], 480)
        rs = ops.Resize(g["h"], g["w"], oh, ow)
        x = rs.fwd(img)
        reg = ops.Regressor(sd, max_crops=10, precision="fp32")
        offs = g["offsets"].to(DEV).contiguous()
        logits = reg.forward(x, offs, normalize=True)
        preds = torch.empty(1, 4, device=DEV)
        loss = torch.empty(1, device=DEV)
        dlogits = torch.empty_like(logits)
        ops.va_head(logits, 1, 10, True, g["target"].to(DEV), 0.5, 0.0, 3, g["weight_clf"], preds, loss, dlogits)
        dx = torch.empty_like(x)
        reg.backward(dlogits, dx)
        dimg = rs.bwd(dx)
        torch.cuda.synchronize()
        perr = (preds.cpu()[:, :2] - g["pred"]).abs().max().item()
        lerr = abs(loss.item() - g["loss"].item())
        print(tag, "pred", preds.cpu()[0, :2].tolist(), g["pred"][0].tolist(), "loss", loss.item(), g["loss"].item())
        assert perr <= 2e-5 and lerr <= 1e-6
        gd = dimg.cpu()[0, :, ::8, ::8]
        gerr = (gd - g["grad_ds"]).abs().max().item() / g["grad_max"].item()
        serr = abs(dimg.abs().sum().item() - g["grad_abs_sum"].item()) / g["grad_abs_sum"].item()
        merr = (gd - g["grad_ds"]).abs().mean().item() / g["grad_ds"].abs().mean().item()
        print(tag, f"grad: max-rel {gerr:.3e}, mean-rel {merr:.3e}, abs-sum rel {serr:.3e}")
        # isolated pixels sit on ReLU / max-pool kinks whose side is decided by fp32 round-off (accumulation order):
        # the max over the sampled pixels is bounded loosely, the mean and the |.|-sum tightly
        # (mean-rel measured: 0.4e-3 .. 1.4e-3 with the CUDA-core fp32 GEMM, up to 2.2e-3 with the tensor-core bf16x3 GEMM)
        assert gerr <= 5e-2 and merr <= 4e-3 and serr <= 5e-3
        del reg


def test_back_to_back_fusion_opt_in_matches():
    """RGIE_GEMM_B2B (gemm_b2b_kernel: layer1 conv3 + skip -> the next block's conv1 in ONE launch, and the same for the
    input-gradient pairs; the second GEMM reads the first one's output tile from the shared-memory boxes of its TMA store)
    must give the same numbers as the two separate launches.  The switch is read once per process, so the check runs the
    bf16 parity tests in a child process with the fusion forced ON and forced OFF."""
    import subprocess, sys
    for flag in ("1", "0"):
        env = dict(os.environ, RGIE_GEMM_B2B=flag)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", __file__, "-k",
                            "test_bf16_tcgen05_matches_bf16_simt_and_oracle or test_regressor_golden"],
                           env=env, capture_output=True, text=True, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, f"RGIE_GEMM_B2B={flag}: " + r.stdout[-2000:] + r.stderr[-2000:]
        assert "2 passed" in r.stdout


_STEM_CHILD = r"""
import sys, torch
sys.path.insert(0, {root!r})
from oracle import oracle as O
from regressor_guided_image_editing_b200 import ops
sd = O.make_regressor_state_dict()
torch.manual_seed(5)
img = torch.stack([O.synthetic_image(20 + i, 480, 480) for i in range(2)]).cuda()
offs = O.draw_crop_offsets(1, 2, 480, 480)[0].int().cuda()
reg = ops.Regressor(sd, max_crops=6, crop_size=448, precision="bf16")
logits = reg.forward(img, offs[:, :3].contiguous(), normalize=True)
pool = reg.tap("pool", (6, 64, 112, 112))
dimg = torch.empty_like(img)
reg.backward(torch.randn(6, 4, generator=torch.Generator().manual_seed(3)).cuda(), dimg)
torch.cuda.synchronize()
torch.save(dict(logits=logits.cpu(), pool=pool.cpu(), dimg=dimg.cpu()), sys.argv[1])
"""


def test_fused_conv1_maxpool_is_bit_identical(tmp_path):
    """gemm_conv1_pool_kernel (conv1 + 3x3/2 max-pool in one launch; the 224 x 224 x 64 stem activation never reaches HBM)
    against conv1 -> maxpool_fwd_kernel (RGIE_STEM_POOL=0): pooled tensor, logits and the image gradient (which goes through
    the argmax bytes) must be bit-identical -- same bf16 rounding of the conv output, same first-maximum scan order.
    The switch is read once per process: two child processes."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for flag in ("1", "0"):
        path = str(tmp_path / f"stem_{flag}.pt")
        r = subprocess.run([sys.executable, "-c", _STEM_CHILD.format(root=root), path], env=dict(os.environ, RGIE_STEM_POOL=flag),
                           capture_output=True, text=True, timeout=300, cwd=root)
        assert r.returncode == 0, f"RGIE_STEM_POOL={flag}: " + r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(torch.load(path))
    a, b = outs
    assert a["pool"].abs().max().item() > 0
    assert torch.equal(a["pool"], b["pool"]), (a["pool"] - b["pool"]).abs().max().item()
    assert torch.equal(a["logits"], b["logits"])
    assert torch.equal(a["dimg"], b["dimg"])


def test_conv3_hshare_on_and_off_pass_the_bf16_parity_tests():
    """RGIE_CONV3_HSHARE (conv3_hshare_kernel: layer1's 3x3 convs and their input gradients with the three horizontal taps as
    the N dimension, horizontal sum by warp shuffles in the epilogue) against the CTA-pair patch kernel: both settings must
    pass the bf16 parity tests (tcgen05 vs the CUDA-core bf16 GEMM and vs the fp32 oracle, reference-generated golden).
    The switch is read once per process: child processes."""
    import subprocess, sys
    for flag in ("1", "0"):
        env = dict(os.environ, RGIE_CONV3_HSHARE=flag)
        r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", __file__, "-k",
                            "test_bf16_tcgen05_matches_bf16_simt_and_oracle or test_regressor_golden"],
                           env=env, capture_output=True, text=True, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0, f"RGIE_CONV3_HSHARE={flag}: " + r.stdout[-2000:] + r.stderr[-2000:]
        assert "2 passed" in r.stdout

"""GPU: the filter kernels (8 default filters + gamma / brightness / b&w / hue / white balance; forward + backward)
against the CPU oracle and the reference-generated goldens."""
import os

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _mirror():
    from regressor_guided_image_editing_b200.baselines.image_transformations import image_transformations as IT
    return IT


def _oracle_single(name, im, p):
    return torch.clamp(O.apply_one(name, im, p), 0.0, 1.0)


def _param_tensor(name, val):
    if name == "tone":
        return torch.tensor(val, dtype=torch.float32).view(1, 1, 8, 1)
    if name == "color":
        return torch.tensor(val, dtype=torch.float32).view(1, 3, 8, 1)
    if name == "scale":
        return torch.tensor(val, dtype=torch.float32).view(1, 4)
    if name == "affine":
        return torch.tensor(val, dtype=torch.float32).view(1, 2, 3)
    if name == "bw":
        return torch.tensor(val, dtype=torch.float32).view(1)      # the reference indexes bw_param[:, None, None, None]
    return torch.tensor(val, dtype=torch.float32)


g7 = torch.Generator().manual_seed(7)
SINGLE_CASES = [
    ("exposure", 0.0), ("exposure", 0.45), ("exposure", -0.8),
    ("saturation", 1.0), ("saturation", 0.0), ("saturation", 1.7), ("saturation", 0.35),
    ("tone", [1.0] * 8), ("tone", (1.0 + 0.3 * torch.randn(8, generator=g7)).tolist()),
    ("color", [1.0] * 24), ("color", (1.0 + 0.3 * torch.randn(24, generator=g7)).tolist()),
    ("contrast", 1.0), ("contrast", 0.5), ("contrast", 1.8), ("contrast", 0.0),
    ("sharp", 0.0), ("sharp", 0.5), ("sharp", 1.0), ("sharp", 2.5),
    ("blur", 1e-4), ("blur", 0.5), ("blur", 2.0), ("blur", 6.0),
    # generic (non-"nice") values: with e.g. sx=1.5, cx=30 a third of the sample points land exactly on pixel centres,
    # where bilinear sampling has a kink and the side the reference takes is decided by float round-off of
    # torch.linspace (platform dependent); identity is checked separately (forward + d(image) only)
    ("scale", [1.0, 1.0, 0.0, 0.0]), ("scale", [1.2371, 1.1113, 9.37, 14.21]), ("scale", [1.5311, 1.0173, 30.19, 5.23]),
    ("scale", [1.0537, 1.3071, 0.0, 40.43]),
    # the remaining pointwise filters of apply_params.  gamma < 1 is left out on purpose: the test
    # image holds exact zeros, where d(x^gamma)/dx is +inf in the reference too
    ("gamma", 1.0), ("gamma", 2.2), ("gamma", 1.3), ("gamma", 0.0),
    ("bright", 0.0), ("bright", 0.3), ("bright", 1.0),
    ("bw", 0.0), ("bw", 0.4), ("bw", 1.0),
    ("hue", 0.0), ("hue", 0.7), ("hue", -2.1), ("hue", 3.0),
    ("wb", 0.0), ("wb", 0.5), ("wb", 1.0),
    # general affine warp with border padding (generic values; identity is a kink for d(param) like scale's)
    ("affine", [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]), ("affine", [1.0371, 0.0513, 1.37, -0.0431, 0.9713, -2.21]),
    ("affine", [0.8713, -0.1211, 3.19, 0.0917, 1.1307, 1.43]),
]


@pytest.mark.parametrize("hw", [(36, 44), (33, 57), (128, 96)])
@pytest.mark.parametrize("name,val", SINGLE_CASES, ids=[f"{n}-{i}" for i, (n, _) in enumerate(SINGLE_CASES)])
def test_single_filter_vs_oracle(name, val, hw):
    IT = _mirror()
    h, w = hw
    im = O.synthetic_image(6, h, w)[None]
    # make clamps / ties bite: push some pixels to the ends of the range
    im = torch.clamp(im * 1.25 - 0.1, 0.0, 1.0)
    p = _param_tensor(name, val)
    gout = torch.randn(im.shape, generator=torch.Generator().manual_seed(13))

    im_c, p_c = im.clone().requires_grad_(True), p.clone().requires_grad_(True)
    out_c = _oracle_single(name, im_c, p_c)
    gi_c, gp_c = torch.autograd.grad((out_c * gout).sum(), [im_c, p_c], allow_unused=True)
    gp_c = torch.zeros_like(p) if gp_c is None else gp_c

    im_g, p_g = im.to(DEV).requires_grad_(True), p.to(DEV).requires_grad_(True)
    out_g = IT._DISPATCH[name](im_g, p_g)
    gi_g, gp_g = torch.autograd.grad((out_g * gout.to(DEV)).sum(), [im_g, p_g], allow_unused=True)
    gp_g = torch.zeros_like(p_g) if gp_g is None else gp_g

    warp = name in ("scale", "affine")
    fwd_tol = 2e-4 if warp else 2e-6                # bilinear sampling positions differ at the 1e-5 px level
    err = (out_g.cpu() - out_c).abs().max().item()
    assert err <= fwd_tol, f"{name} forward max-abs {err}"
    # image gradient: robust relative-L1 criterion (isolated pixels sit exactly on clamp / tie boundaries)
    gden = gi_c.abs().mean().item() + 1e-12
    gerr = (gi_g.cpu() - gi_c).abs().mean().item() / gden
    gtol = 2e-3 if warp else 1e-4
    assert gerr <= gtol, f"{name} d(image) relative L1 {gerr}"
    pden = gp_c.abs().max().item() + 1e-6 * gout.numel() ** 0.5
    perr = (gp_g.cpu() - gp_c).abs().max().item() / pden
    ptol = 5e-3 if warp else 5e-4
    if (name == "scale" and list(val) == [1.0, 1.0, 0.0, 0.0]) or (name == "affine" and list(val) == [1.0, 0.0, 0.0, 0.0, 1.0, 0.0]):
        return            # d(param) at exact identity is a kink (see SINGLE_CASES comment)
    assert perr <= ptol, f"{name} d(param) {gp_g.cpu().flatten()[:4]} vs {gp_c.flatten()[:4]} rel {perr}"


def test_identity_known_answer():
    """run_img_trans.py SAME preset / init_params start values: exposure, tone, color, contrast, scale leave the image
    unchanged; blur at 1e-4 is a delta; sharp at 0 returns the 3x3-smoothed image."""
    IT = _mirror()
    im = O.synthetic_image(2, 64, 80)[None].to(DEV)
    x0 = O.init_x0()
    p = O.get_params_from_vector(x0, O.DEFAULT_FILTERS, 64)
    for name in ("exposure", "tone", "color", "contrast", "blur"):
        out = IT._DISPATCH[name](im, p[name].to(DEV) if isinstance(p[name], torch.Tensor) else p[name])
        assert (out - im).abs().max().item() <= 1e-6, name
    out = IT.apply_scale(im, p["scale"].to(DEV))
    assert (out - im).abs().max().item() <= 1e-4
    out = IT.apply_saturation(im, p["saturation"].to(DEV))
    assert (out - im).abs().max().item() <= 2e-6


def test_chain_vs_golden(golden_dir):
    """apply_params on the reference-generated goldens (tests/golden/filters.pt, made by oracle/gen_golden.py)."""
    IT = _mirror()
    gold = torch.load(os.path.join(golden_dir, "filters.pt"))
    checked = 0
    for key, g in gold.items():
        if key == "singles" or any(t not in IT._DISPATCH for t in g["trans"]):
            continue
        im = O.synthetic_image(g["image_index"], g["h"], g["w"])[None].to(DEV).requires_grad_(True)
        x = g["x"].to(DEV).requires_grad_(True)
        params = O.get_params_from_vector(x, g["trans"], g["h"])
        outs = IT.apply_params(im, params)
        ref = g["stages"][-1]
        err = (outs[-1].detach().cpu() - ref).abs().max().item()
        assert err <= 5e-4, f"{key}: forward max-abs {err}"
        gout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(g["gout_seed"])).to(DEV)
        gx, gim = torch.autograd.grad((outs[-1] * gout).sum(), [x, im], allow_unused=True)
        gx_c, gx_r = gx.cpu().clone(), g["grad_x"].clone()
        if key.startswith(("identity", "branches")) and "scale" in g["trans"]:
            lay = O.param_layout(g["trans"])["scale"]
            gx_c[lay[0]:lay[0] + 4] = 0; gx_r[lay[0]:lay[0] + 4] = 0      # scale at exact identity: kink, side = round-off
        den = gx_r.abs().max().item() + 1e-3
        rel = (gx_c - gx_r).abs().max().item() / den
        assert rel <= 2e-2, f"{key}: d(x) rel {rel}\n{gx.cpu()}\n{g['grad_x']}"
        gden = g["grad_im"].abs().mean().item() + 1e-12
        grel = (gim.cpu() - g["grad_im"]).abs().mean().item() / gden
        assert grel <= 2e-2, f"{key}: d(image) rel-L1 {grel}"
        checked += 1
    assert checked >= 8


def _raw_fwd(kind, x, p, off, stride):
    """rgie_filter_fwd on the parameter block that starts `off` floats into each row of the [B, stride] table."""
    import ctypes as C
    from regressor_guided_image_editing_b200 import _lib, ops
    B, _, H, W = x.shape
    out, ws = torch.empty_like(x), ops.filter_workspace(B, H, W, x.device)
    _lib.check(_lib.load().rgie_filter_fwd(kind, _lib.ptr(x), _lib.ptr(out), C.c_void_p(p.data_ptr() + 4 * off), stride, B, H, W,
                                           _lib.ptr(ws), _lib.stream_ptr(x.device)))
    return out


def _raw_bwd(kind, x, gout, p, gp, off, stride):
    import ctypes as C
    from regressor_guided_image_editing_b200 import _lib, ops
    B, _, H, W = x.shape
    gin, ws = torch.empty_like(x), ops.filter_workspace(B, H, W, x.device)
    _lib.check(_lib.load().rgie_filter_bwd(kind, _lib.ptr(x), _lib.ptr(gout), _lib.ptr(gin), C.c_void_p(p.data_ptr() + 4 * off),
                                           stride, C.c_void_p(gp.data_ptr() + 4 * off), stride, B, H, W, _lib.ptr(ws),
                                           _lib.stream_ptr(x.device)))
    return gin


@pytest.mark.parametrize("hw", [(36, 44), (33, 57), (64, 64)])
def test_fused_prefix_equals_four_passes(hw):
    """rgie_filter_prefix_fwd/bwd (exposure -> saturation -> tone -> colour as one pixel pass, the head of the default
    list) against the four separate kernels: bit-identical image, parameter gradients equal up to summation order."""
    from regressor_guided_image_editing_b200 import _lib, ops
    h, w = hw
    B = 3
    g = torch.Generator().manual_seed(21)
    im = torch.stack([torch.clamp(O.synthetic_image(10 + i, h, w) * 1.3 - 0.1, 0, 1) for i in range(B)]).to(DEV)
    p = torch.cat([0.4 * torch.randn(B, 1, generator=g), 1.0 + 0.5 * torch.rand(B, 1, generator=g),
                   1.0 + 0.3 * torch.randn(B, 32, generator=g), torch.zeros(B, 7)], dim=1).to(DEV).contiguous()   # stride 41
    NP = p.shape[1]
    gout = torch.randn(im.shape, generator=g).to(DEV)
    fused = ops.filter_prefix_fwd(im, p, NP)
    names, offs, cnt = ["exposure", "saturation", "tone", "color"], [0, 1, 2, 10], [1, 1, 8, 24]
    stages = [im]
    for n, o in zip(names, offs):
        stages.append(_raw_fwd(_lib.FILTER_KINDS[n], stages[-1], p, o, NP))
    assert torch.equal(fused, stages[-1])
    gp_seq = torch.zeros(B, NP, device=DEV)
    gcur = gout
    for k in reversed(range(4)):
        gcur = _raw_bwd(_lib.FILTER_KINDS[names[k]], stages[k], gcur, p, gp_seq, offs[k], NP)
    gp_fused = torch.zeros(B, NP, device=DEV)
    ops.filter_prefix_bwd(im, gout, p, NP, gp_fused, NP)
    torch.cuda.synchronize()
    assert torch.equal(gp_fused[:, 34:], torch.zeros(B, NP - 34, device=DEV))
    for o, c in zip(offs, cnt):
        a, b = gp_fused[:, o:o + c], gp_seq[:, o:o + c]
        assert (a - b).abs().max().item() <= 2e-5 * (b.abs().max().item() + 1e-3), (o, a, b)


def test_single_filters_vs_reference_generated_goldens(golden_dir):
    """Each filter of the default list alone against the per-filter fixtures the REFERENCE generated (filters.pt `singles`,
    oracle/gen_golden.py: the reference's apply_params on one filter, its autograd for d/d(param) and d/d(image))."""
    IT = _mirror()
    s = torch.load(os.path.join(golden_dir, "filters.pt"))["singles"]
    im0 = O.synthetic_image(s["image_index"], s["h"], s["w"])[None]
    gout = torch.randn(im0.shape, generator=torch.Generator().manual_seed(s["gout_seed"])).to(DEV)
    assert len(s["cases"]) >= 18
    for key, c in s["cases"].items():
        im = im0.to(DEV).requires_grad_(True)
        p = torch.tensor(float(c["value"]), device=DEV, requires_grad=True)
        out = IT._DISPATCH[c["name"]](im, p)
        gi, gp = torch.autograd.grad((out * gout).sum(), [im, p], allow_unused=True)
        gp = torch.zeros(()) if gp is None else gp.cpu()
        err = (out.detach().cpu() - c["out"]).abs().max().item()
        assert err <= 2e-6, f"{key}: forward max-abs {err}"
        gerr = (gi.cpu() - c["grad_im"]).abs().mean().item() / (c["grad_im"].abs().mean().item() + 1e-12)
        assert gerr <= 1e-4, f"{key}: d(image) relative L1 {gerr}"
        perr = abs(gp.item() - c["grad_p"].item()) / (abs(c["grad_p"].item()) + 1e-6 * gout.numel() ** 0.5)
        assert perr <= 5e-4, f"{key}: d(param) {gp.item()} vs {c['grad_p'].item()}"


def test_extra_filters_vs_reference_generated_goldens(golden_dir):
    """filters_extra.pt (oracle/gen_golden.py --only filters_extra): tone / colour curves, scale, gamma, brightness, b&w (1-d
    parameter), hue, white balance and the general affine warp, each alone through the REFERENCE's apply_params with its
    autograd gradients, against the native kernels."""
    IT = _mirror()
    s = torch.load(os.path.join(golden_dir, "filters_extra.pt"))
    im0 = torch.clamp(O.synthetic_image(s["image_index"], s["h"], s["w"])[None] * 1.25 - 0.1, 0.0, 1.0)
    gout = torch.randn(im0.shape, generator=torch.Generator().manual_seed(s["gout_seed"])).to(DEV)
    assert len(s["cases"]) == 21
    for key, c in s["cases"].items():
        im = im0.to(DEV).requires_grad_(True)
        p = c["param"].to(DEV).requires_grad_(True)
        out = IT._DISPATCH[c["name"]](im, p)
        gi, gp = torch.autograd.grad((out * gout).sum(), [im, p], allow_unused=True)
        gp = torch.zeros_like(c["param"]) if gp is None else gp.cpu()
        warp = c["name"] in ("scale", "affine")
        err = (out.detach().cpu() - c["out"]).abs().max().item()
        assert err <= (2e-4 if warp else 2e-6), f"{key}: forward max-abs {err}"
        gerr = (gi.cpu() - c["grad_im"]).abs().mean().item() / (c["grad_im"].abs().mean().item() + 1e-12)
        assert gerr <= (2e-3 if warp else 1e-4), f"{key}: d(image) relative L1 {gerr}"
        perr = (gp - c["grad_p"]).abs().max().item() / (c["grad_p"].abs().max().item() + 1e-6 * gout.numel() ** 0.5)
        assert perr <= (5e-3 if warp else 5e-4), f"{key}: d(param) {gp.flatten()[:4]} vs {c['grad_p'].flatten()[:4]} rel {perr}"

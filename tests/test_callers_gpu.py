"""GPU: `output_transform` (optimize_image_param.py:295-312) -- the caller right after the optimisation loop: evaluation of
the edit at the working size (two native regressor predictions + statistics) and the full-resolution re-render of the same
parameters through the native filter kernels, against the oracle's apply_params on the same file."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_output_transform_rerenders_at_full_resolution(tmp_path, capsys):
    import PIL.Image
    from regressor_guided_image_editing_b200 import optimize_image_param as oip
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss
    (tmp_path / "images" / "set").mkdir(parents=True)
    rng = np.random.default_rng(4)
    # smooth synthetic photo, non-square, grayscale-free: low-frequency field + noise
    yy, xx = np.mgrid[0:700, 0:900]
    base = np.stack([0.5 + 0.4 * np.sin(xx / 90.0 + c) * np.cos(yy / 70.0 - c) for c in range(3)], -1)
    arr = np.clip(base + 0.05 * rng.standard_normal(base.shape), 0, 1)
    PIL.Image.fromarray((arr * 255).astype(np.uint8)).save(tmp_path / "images" / "set" / "a.png")
    oip.configure_output(output_size=640, data_dir=str(tmp_path))
    oip.STATS.clear()

    sd = O.make_regressor_state_dict()
    clf = ValenceArousalLoss(sd, torch.device(DEV), 1, is_minimized=True, requires_grad=False, precision="fp32")
    work = O.synthetic_image(3, 96, 96)[None].to(DEV)
    params_trans, x0 = oip.init_params(oip.DEFAULT_TRANS)
    x = x0 + 0.08 * torch.randn(41, generator=torch.Generator().manual_seed(5))
    x[36], x[35] = 1.1, 0.6                                  # blur sigma, sharpen amount away from their flat start values
    obj_params = {"params": params_trans, "clf": clf}
    eval_params = {"emotion_type_labels": ['Valence', 'Arousal']}
    torch.manual_seed(77)
    full_in, full_out = oip.output_transform(work, x.to(DEV), obj_params, eval_params, 0.1, ["set/a.png"])
    assert full_in.shape == (1, 3, 640, 640) and full_out.shape == full_in.shape and full_out.is_cuda

    # same parameters, same file, oracle filters on CPU (scale centres unpacked against the WORKING size, as the reference does)
    ref = O.apply_params(full_in.cpu(), O.get_params_from_vector(x, O.DEFAULT_FILTERS, 96))[-1]
    err = (full_out.cpu() - ref).abs()
    assert err.max().item() <= 5e-4 and err.mean().item() <= 2e-5, (err.max().item(), err.mean().item())

    # evaluation side: one entry per statistic, adapted prediction / delta consistent with the regressor on the edit
    st = oip.STATS[0.1]
    assert all(len(st[k]) == 1 for k in ("valence", "arousal", "delta_valence", "delta_arousal", "rec_error"))
    edited = O.apply_params(work.cpu(), O.get_params_from_vector(x, O.DEFAULT_FILTERS, 96))[-1]
    assert abs(st["rec_error"][0] - (edited - work.cpu()).abs().mean().item()) <= 1e-5
    torch.manual_seed(77)
    offs = O.draw_crop_offsets(2, 1, 480, 480)
    p_before = O.regressor_predict(work.cpu(), sd, offs[0])
    p_after = O.regressor_predict(edited, sd, offs[1])
    assert abs(st["valence"][0] - p_after[0, 0].item()) <= 2e-3
    assert abs(st["delta_arousal"][0] - (p_after - p_before)[0, 1].item()) <= 2e-3
    assert "reconstruction error" in capsys.readouterr().out


def test_output_transform_needs_configuration():
    from regressor_guided_image_editing_b200 import optimize_image_param as oip, _lib
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss
    oip.OUTPUT_TRANSFORM = None
    clf = ValenceArousalLoss(O.make_regressor_state_dict(), torch.device(DEV), 1, precision="fp32")
    params_trans, x0 = oip.init_params(oip.DEFAULT_TRANS)
    with pytest.raises(_lib.RgieError):
        oip.output_transform(O.synthetic_image(3, 64, 64)[None].to(DEV), x0.to(DEV), {"params": params_trans, "clf": clf},
                             {"emotion_type_labels": ['Valence', 'Arousal']}, 0.1, ["x.png"])


def test_parametric_objective_with_clip_reconstruction_term():
    """weight_recon > 0 (the script's default objective, optimize_image_param.py:249-257): loss and d(loss)/d(x) of
    objective_function_parametric with the native filters + native regressor + the CLIP term around a seeded stand-in tower,
    against the CPU oracle's regressor term plus the same term evaluated on the oracle's edited image (compute_clip_loss is
    pinned bit for bit to the reference's on CPU: tests/test_callers_cpu.py::test_clip_loss_matches_reference)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_callers_cpu import _FakeClipTower
    from oracle import oracle as O
    from regressor_guided_image_editing_b200 import optimize_image_param as oip
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss
    sd = O.make_regressor_state_dict()
    h = w = 96
    image = O.smooth_image(3, h, w)[None]
    x = O.init_x0().clone()
    x[0], x[1], x[34], x[35], x[36] = 0.2, 1.2, 1.1, 0.3, 0.8                      # exposure, saturation, contrast, sharp, blur
    x[2:34] += 0.1 * torch.randn(32, generator=torch.Generator().manual_seed(41))   # tone + colour curves
    x[37:41] = torch.tensor([1.2371, 1.1113, 9.37, 14.21])                          # scale at generic values (no kink)
    target = torch.tensor([[0.7, 0.5]])
    w_clf, w_rec = 0.15, 0.5
    torch.manual_seed(2011)
    offs = O.draw_crop_offsets(1, 1, 480, 480)

    prev, tf32 = oi.CLIP_MODEL, torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the stand-in tower's convolution would otherwise run in TF32 on the GPU
    try:
        oi.CLIP_MODEL = _FakeClipTower()
        xc = x.clone().requires_grad_(True)
        loss_reg, _, edited = O.objective_parametric(xc, image, sd, offs[0], target, w_clf)
        clip_c = oi.compute_clip_loss(image, edited)
        loss_c = loss_reg + w_rec * clip_c
        g_c, = torch.autograd.grad(loss_c, xc)

        oi.CLIP_MODEL = _FakeClipTower().to(DEV)
        clf = ValenceArousalLoss(sd, torch.device(DEV), 1, is_minimized=True, requires_grad=True, precision="fp32")
        _, params = oip.initialize_parametric(image.to(DEV), {"clf": clf, "dis": None, "weight_clf": w_clf, "weight_dis": 0.0,
                                                              "weight_recon": w_rec, "target": target.to(DEV)})
        xg = x.to(DEV).requires_grad_(True)
        torch.manual_seed(2011)
        loss_g = oip.objective_function_parametric(xg, **params)
        g_g, = torch.autograd.grad(loss_g, xg)
    finally:
        oi.CLIP_MODEL, torch.backends.cudnn.allow_tf32 = prev, tf32
    print("loss", loss_g.item(), loss_c.item(), "clip term", clip_c.item())
    assert clip_c.item() > 1e-4                                                     # the term is live
    assert abs(loss_g.item() - loss_c.item()) <= 1e-4
    g_g = g_g.cpu()
    for name, (o, n) in dict(exposure=(0, 1), saturation=(1, 1), tone=(2, 8), color=(10, 24), contrast=(34, 1), sharp=(35, 1),
                             blur=(36, 1), scale=(37, 4)).items():
        ref = g_c[o:o + n]
        rel = (g_g[o:o + n] - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        assert rel <= (2e-2 if name == "scale" else 5e-3), (name, rel, g_g[o:o + n], ref)

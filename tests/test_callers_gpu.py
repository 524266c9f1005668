"""GPU: `output_transform` (optimize_image_param.py:295-312) -- the caller right after the optimisation loop: evaluation of
the edit at the working size (two native regressor predictions + statistics) and the full-resolution re-render of the same
parameters through the native filter kernels, against the oracle's apply_params on the same file."""
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

"""GPU: the full per-image optimisation loop (ParametricEditEngine and the drop-in `optimization`) against the
reference-generated golden (tests/golden/loop_c1.pt = BASELINE.json configs[0]: one synthetic 256x256 image, random-init
regressor, 50 steps toward target valence, reference run on CPU) and against the CPU oracle.

Stated tolerances (SURVEY.md 7 "precision vs parity"):
  fp32 mode (CUDA-core GEMMs, fp32 everywhere): all 50 per-step losses |d| <= 5e-5 and predictions |d| <= 2e-3 against the
      reference's own run; the filter chain at the reference's best_x reproduces its edited image to max-abs <= 1e-3.
  bf16 mode (tcgen05 GEMMs): per-step predictions |d| <= 1e-2, losses |d| <= 2e-4.
  Raw parameter vectors are compared tightly for the first steps only: the scale filter's gradient at EXACT identity (its
  start value) is a kink whose side the reference picks by float round-off of torch.linspace (tests/test_filters_gpu.py),
  and Adam turns noise-level gradients into full-size steps, so late parameters are bounded loosely and their
  image-space effect (mean |pixel| difference) is what is asserted.
"""
import os

import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def sd():
    return O.make_regressor_state_dict()


@pytest.fixture(scope="module")
def gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "loop_c1.pt"))


def _run_engine(sd, gold, precision, steps=None, use_graph=True):
    from regressor_guided_image_editing_b200 import engine
    steps = gold["num_steps"] if steps is None else steps
    img = O.synthetic_image(gold["image_index"], gold["h"], gold["w"])[None].to(DEV)
    eng = engine.ParametricEditEngine(sd, batch=1, height=gold["h"], width=gold["w"], num_steps=gold["num_steps"],
                                      precision=precision, use_graph=use_graph)
    eng.load_problem(img, gold["offsets"].to(DEV), alpha=gold["alpha"], learning_rate=gold["learning_rate"],
                     weight_clf=gold["weight_clf"])
    eng.advance(steps)
    torch.cuda.synchronize()
    return eng, {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in eng.results().items()}


def _edit_with(sd_unused, x, h, w, image_index):
    """Edited image for a given raw parameter vector through the native filter chain."""
    from regressor_guided_image_editing_b200 import optimize_image_param as oip
    from regressor_guided_image_editing_b200.baselines.image_transformations.image_transformations import apply_params
    img = O.synthetic_image(image_index, h, w)[None].to(DEV)
    ptmpl, _ = oip.init_params(oip.DEFAULT_TRANS)
    with torch.no_grad():
        return apply_params(img, oip.get_params_from_vector(x.to(DEV), 1, ptmpl, h))[-1].cpu()


def _check_against_golden(out, gold, tol_loss, tol_pred, tol_x, tag):
    dl = (out["losses"][:, 0] - gold["losses"]).abs()
    dp = (out["preds"][:, 0, :2] - gold["preds"]).abs().max(1).values
    dx = (out["xs"][:, 0] - gold["xs"]).abs().max(1).values          # x BEFORE each step, all 50 steps
    print(f"{tag} per-step |dloss|:", [f"{v:.1e}" for v in dl.tolist()])
    print(f"{tag} per-step |dpred|:", [f"{v:.1e}" for v in dp.tolist()])
    print(f"{tag} per-step |dx|   :", [f"{v:.1e}" for v in dx.tolist()])
    dlast = (out["xs"][-1, 0] - gold["xs"][-1]).abs()
    top = torch.topk(dlast, 6)
    print(f"{tag} largest parameter drifts at the last step (index: |d|, mine, golden):",
          [(int(i), round(float(v), 4), round(float(out["xs"][-1, 0, i]), 4), round(float(gold["xs"][-1, i]), 4))
           for v, i in zip(top.values, top.indices)])
    # image-space effect of the parameter drift: edit with my last x and with the golden's last x (same native filters)
    ed_a = _edit_with(None, out["xs"][-1, 0], gold["h"], gold["w"], gold["image_index"])
    ed_b = _edit_with(None, gold["xs"][-1], gold["h"], gold["w"], gold["image_index"])
    print(f"{tag} edited(x_mine[-1]) vs edited(x_golden[-1]): max-abs {(ed_a - ed_b).abs().max().item():.3e} "
          f"mean-abs {(ed_a - ed_b).abs().mean().item():.3e}")
    assert dl.max().item() <= tol_loss, f"{tag}: per-step loss drifted ({dl.max().item():.3e})"
    assert dp.max().item() <= tol_pred, f"{tag}: per-step prediction drifted ({dp.max().item():.3e})"
    # Raw parameters: tight while no kink / noise-level gradient has been amplified by Adam's normalisation (first steps),
    # then only bounded (the scale parameters random-walk from the identity kink; module docstring).  What the drift means
    # in image space is asserted through the mean absolute pixel difference.
    assert dx[:3].max().item() <= 1e-4, f"{tag}: parameters differ before any kink could matter"
    assert dx.max().item() <= tol_x, f"{tag}: parameter trajectory drifted ({dx.max().item():.3e})"
    assert (ed_a - ed_b).abs().mean().item() <= 5e-3, f"{tag}: image-space drift"
    # best-x is an argmin over per-step losses that differ by less than the loss tolerance late in the run: accept any
    # step whose golden loss is within tolerance of the golden minimum
    ok_steps = (gold["losses"] <= gold["losses"].min() + 2 * tol_loss).nonzero().flatten().tolist()
    assert out["best_step"].item() in ok_steps, (out["best_step"].item(), ok_steps)
    assert abs(out["best_loss"].item() - gold["losses"].min().item()) <= 2 * tol_loss


def test_loop_c1_fp32_matches_reference_golden(sd, gold):
    """BASELINE.json configs[0] against the reference's own CPU run (fp32 parity mode)."""
    eng, out = _run_engine(sd, gold, "fp32")
    assert (out["target"] - gold["target"]).abs().max().item() <= 1e-5
    dl = (out["losses"][:, 0] - gold["losses"]).abs()
    dp = (out["preds"][:, 0, :2] - gold["preds"]).abs().max(1).values
    assert dl[:3].max().item() <= 2e-6 and dp[:3].max().item() <= 2e-5      # before any kink can matter
    _check_against_golden(out, gold, tol_loss=5e-5, tol_pred=2e-3, tol_x=0.5, tag="fp32")
    # edited image: same parameters -> same pixels (max-abs <= 1e-3, north_star fp32 tolerance)
    ed = _edit_with(sd, gold["best_x"], gold["h"], gold["w"], gold["image_index"])
    err = (ed - gold["edited"]).abs()
    print("edited image at the golden best_x: max-abs", err.max().item(), "mean-abs", err.mean().item())
    assert err.max().item() <= 1e-3
    # and the image the engine itself ends with, at ITS best_x, is what its own filters give
    ed2 = _edit_with(sd, out["best_x"][0], gold["h"], gold["w"], gold["image_index"])
    assert (ed2 - out["edited"]).abs().max().item() <= 1e-6


def test_loop_c1_bf16_tracks_reference_golden(sd, gold):
    eng, out = _run_engine(sd, gold, "bf16")
    """Throughput mode (tcgen05 bf16 regressor): stated tolerance |dpred| <= 1e-2, |dloss| <= 2e-4 per step."""
    _check_against_golden(out, gold, tol_loss=2e-4, tol_pred=1e-2, tol_x=1.5, tag="bf16")


def test_graph_replay_equals_eager(sd, gold):
    _, a = _run_engine(sd, gold, "bf16", steps=6, use_graph=True)
    _, b = _run_engine(sd, gold, "bf16", steps=6, use_graph=False)
    assert torch.equal(a["losses"][:6], b["losses"][:6])
    assert torch.equal(a["x_last"], b["x_last"])


def test_batch_is_independent_problems(sd):
    """Image b of a batch gives bit-identical results to running it alone (sharding across GPUs relies on this)."""
    from regressor_guided_image_editing_b200 import engine
    steps, h, w = 4, 128, 160
    imgs = torch.stack([O.synthetic_image(i, h, w) for i in range(3)]).to(DEV)
    offs = []
    for i in range(3):
        torch.manual_seed(2000 + i)
        oh, ow = O.resize_output_size(h, w, 480)
        offs.append(O.draw_crop_offsets(1 + steps, 1, oh, ow))
    offs = torch.cat(offs, 1).to(DEV)
    eng3 = engine.ParametricEditEngine(sd, batch=3, height=h, width=w, num_steps=steps, precision="bf16")
    r3 = eng3.run(imgs, offs)
    eng3m = engine.ParametricEditEngine(sd, batch=3, height=h, width=w, num_steps=steps, precision="bf16", micro_batch=1)
    r3m = eng3m.run(imgs, offs)
    eng1 = engine.ParametricEditEngine(sd, batch=1, height=h, width=w, num_steps=steps, precision="bf16")
    for i in range(3):
        r1 = eng1.run(imgs[i:i + 1].contiguous(), offs[:, i:i + 1].contiguous())
        assert torch.equal(r1["losses"][:, 0], r3["losses"][:, i]), f"image {i}: batch != alone"
        assert torch.equal(r1["best_x"][0], r3["best_x"][i])
        assert torch.equal(r1["edited"][0], r3["edited"][i])
        assert torch.equal(r1["losses"][:, 0], r3m["losses"][:, i]), f"image {i}: micro-batched != alone"


def test_dropin_optimization_fused_and_generic_paths(sd):
    """The reference call surface: ValenceArousalLoss + initialize_parametric + objective_function_parametric +
    optimization(), fused path vs generic autograd path vs the CPU oracle, same seeds."""
    from regressor_guided_image_editing_b200 import optimize_image_param as oip
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss
    steps, h, w = 3, 96, 96
    image = O.synthetic_image(7, h, w)[None]
    torch.manual_seed(2007)
    offs = O.draw_crop_offsets(1 + steps, 1, 480, 480)
    ref = O.optimize_parametric(image, sd, offs, alpha=0.1, learning_rate=0.05, num_steps=steps, weight_clf=0.15)

    def run(force_generic):
        clf = ValenceArousalLoss(sd, torch.device(DEV), 1, is_minimized=True, requires_grad=True, precision="fp32")
        params = {"clf": clf, "dis": None, "weight_clf": 0.15, "weight_dis": 0.0, "weight_recon": 0.0, "alpha": 0.1}
        torch.manual_seed(2007)
        x0, params = oip.initialize_parametric(image.to(DEV), params)
        params["target"] = oi.get_condition_from_alpha(params["alpha"], params["clf"], image.to(DEV))
        del params["alpha"]
        if force_generic:
            obj = lambda x, **kw: oip.objective_function_parametric(x, **kw)     # a different callable -> generic path
        else:
            obj = oip.objective_function_parametric
        best = oi.optimization(x0, params, obj, learning_rate=0.05, num_steps=steps)
        return best.cpu(), params["target"].cpu()

    best_f, tgt_f = run(False)
    best_g, tgt_g = run(True)
    assert (tgt_f - ref["target"]).abs().max().item() <= 1e-5
    assert (tgt_g - ref["target"]).abs().max().item() <= 1e-5
    # indices 37..40 are the scale filter, whose gradient at exact identity is a round-off-decided kink: excluded
    assert (best_f[:37] - ref["best_x"][:37]).abs().max().item() <= 1e-4, (best_f, ref["best_x"])
    assert (best_g[:37] - ref["best_x"][:37]).abs().max().item() <= 1e-4, (best_g, ref["best_x"])
    assert (best_f[:37] - best_g[:37]).abs().max().item() <= 1e-5

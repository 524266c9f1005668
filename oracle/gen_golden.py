"""ORACLE / TEST INFRASTRUCTURE ONLY.

Generates tests/golden/*.pt by running the UNMODIFIED reference Python (imported from /root/reference/src through
oracle/ref_harness.py) on CPU.  Run in the build container only:

    python -m oracle.gen_golden [--only filters|filters_extra|emonet|regressor|loop|loopk|loops|loop512|midu|munit]

The reference ships no tests/fixtures, so these vectors are the pin for the standalone oracle
(oracle/oracle.py) and, through it, for the CUDA path.  Inputs are regenerated from seeds at test time; only outputs
and small inputs are stored.
"""
from __future__ import annotations

import argparse
import os
import tempfile
import time

import torch

from . import oracle as O
from . import ref_harness

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_clf(r, sd):
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "va_pred_all")
    torch.save(sd, path)
    return r.ValenceArousalLoss(path, torch.device("cpu"), 1, is_minimized=True, requires_grad=True)


def filter_cases():
    """(name, trans list, x vector) cases exercising every branch of the default 8 filters + the 6 'next' ones."""
    cases = []
    d = O.DEFAULT_FILTERS
    lay = O.param_layout(d)
    x_id = O.init_x0(d)
    cases.append(("identity", d, x_id.clone()))
    g = torch.Generator().manual_seed(7)
    x = x_id + 0.15 * torch.randn(41, generator=g)
    x[lay['blur'][0]] = 1.3
    x[lay['sharp'][0]] = 0.7
    x[lay['scale'][0]:lay['scale'][0] + 4] = torch.tensor([1.2371, 1.1113, 9.37, 14.21])
    cases.append(("perturbed", d, x.clone()))
    x = x_id + 0.3 * torch.randn(41, generator=g)
    x[lay['blur'][0]] = 3.0
    x[lay['sharp'][0]] = 1.8          # clamped-blend branch
    x[lay['contrast'][0]] = 1.6
    x[lay['exposure'][0]] = 0.8       # saturating exposure -> many ties at 1.0 in rgb_to_hsv
    x[lay['scale'][0]:lay['scale'][0] + 4] = torch.tensor([1.5311, 1.0173, 30.19, 5.23])
    cases.append(("strong", d, x.clone()))
    x = x_id.clone()
    x[lay['contrast'][0]] = -0.2      # :291 host branch -> python float 0.0
    x[lay['sharp'][0]] = 1.0          # blend returns input2
    x[lay['saturation'][0]] = 0.0
    x[lay['scale'][0]:lay['scale'][0] + 4] = torch.tensor([0.7, 0.9, -3.0, 500.0])   # clamps active
    cases.append(("branches", d, x.clone()))
    # single filters, non default ('next' row): bw needs a 1-d param in the reference -> skipped there
    for name, val in (("gamma", 1.4), ("bright", 0.2), ("hue", 0.6), ("wb", 0.5)):
        cases.append((f"single_{name}", [name, "contrast"], torch.tensor([val, 1.0])))
    return cases


def gen_filters(r):
    out = {}
    for h, w in ((40, 48), (33, 57)):
        im = O.synthetic_image(5, h, w)[None]
        for name, trans, x in filter_cases():
            p_t, _ = r.optimize_image_param.init_params(trans)
            xv = x.clone().requires_grad_(True)
            imv = im.clone().requires_grad_(True)
            px = r.optimize_image_param.get_params_from_vector(xv, 1, p_t, im.size(2))
            outs = r.image_transformations.apply_params(imv, px)
            gen = torch.Generator().manual_seed(11)
            gout = torch.randn(outs[-1].shape, generator=gen)
            gx, gim = torch.autograd.grad((outs[-1] * gout).sum(), [xv, imv], allow_unused=True)
            out[f"{name}_{h}x{w}"] = dict(trans=list(trans), x=x.clone(), h=h, w=w, image_index=5,
                                         stages=[o.detach().clone() for o in (outs if name in ('perturbed', 'strong') else outs[-1:])],
                                         gout_seed=11,
                                         grad_x=gx.detach().clone() if gx is not None else torch.zeros_like(x),
                                         grad_im=gim.detach().clone())
    # per-filter chains (each filter alone, with its own gradient) for per-kernel parity
    im = O.synthetic_image(6, 36, 44)[None]
    singles = {}
    for name, vals in (("exposure", [0.0, 0.4, -0.7]), ("saturation", [1.0, 0.0, 1.7, 0.4]),
                       ("contrast", [1.0, 0.5, 1.8]), ("sharp", [0.0, 0.5, 1.0, 2.5]),
                       ("blur", [1e-4, 0.5, 2.0, 6.0])):
        for v in vals:
            trans = [name] if name == "contrast" else [name, "contrast"]
            x = torch.tensor([v] if name == "contrast" else [v, 1.0])
            p_t, _ = r.optimize_image_param.init_params(trans)
            xv = x.clone().requires_grad_(True)
            imv = im.clone().requires_grad_(True)
            px = r.optimize_image_param.get_params_from_vector(xv, 1, p_t, im.size(2))
            o1 = r.image_transformations.apply_params(imv, {name: px[name]})[-1]
            gen = torch.Generator().manual_seed(13)
            gout = torch.randn(o1.shape, generator=gen)
            gx, gim = torch.autograd.grad((o1 * gout).sum(), [xv, imv], allow_unused=True)
            singles[f"{name}_{v}"] = dict(name=name, value=v, out=o1.detach().clone(),
                                          grad_p=(gx[0].clone() if gx is not None else torch.tensor(0.0)),
                                          grad_im=gim.detach().clone())
    out["singles"] = dict(image_index=6, h=36, w=44, gout_seed=13, cases=singles)
    torch.save(out, os.path.join(GOLDEN_DIR, "filters.pt"))
    print("filters.pt:", len(out), "entries")


def extra_filter_cases():
    """(key, filter name, parameter tensor in the shape the reference's apply_* expects) -- the filters that filters.pt's
    `singles` leave out: tone / colour curves, scale, and the six 'next' filters incl. bw (1-d parameter) and affine."""
    g = torch.Generator().manual_seed(17)
    cases = []
    for v in (1.0, 2.2, 1.3):
        cases.append((f"gamma_{v}", "gamma", torch.tensor(v)))
    for v in (0.0, 0.3, 1.0):
        cases.append((f"bright_{v}", "bright", torch.tensor(v)))
    for v in (0.0, 0.4, 1.0):
        cases.append((f"bw_{v}", "bw", torch.tensor([v])))                       # img_trans_torch_diff.py:67-70 indexes [:, None, None, None]
    for v in (0.7, -2.1, 3.0):
        cases.append((f"hue_{v}", "hue", torch.tensor(v)))
    for v in (0.0, 0.5, 1.0):
        cases.append((f"wb_{v}", "wb", torch.tensor(v)))
    cases.append(("tone_random", "tone", (1.0 + 0.3 * torch.randn(8, generator=g)).view(1, 1, 8, 1)))
    cases.append(("color_random", "color", (1.0 + 0.3 * torch.randn(24, generator=g)).view(1, 3, 8, 1)))
    cases.append(("scale_a", "scale", torch.tensor([[1.2371, 1.1113, 9.37, 14.21]])))
    cases.append(("scale_b", "scale", torch.tensor([[1.5311, 1.0173, 30.19, 5.23]])))
    cases.append(("affine_a", "affine", torch.tensor([[[1.0371, 0.0513, 1.37], [-0.0431, 0.9713, -2.21]]])))
    cases.append(("affine_b", "affine", torch.tensor([[[0.8713, -0.1211, 3.19], [0.0917, 1.1307, 1.43]]])))
    return cases


def gen_filters_extra(r):
    """tests/golden/filters_extra.pt: one filter at a time through the reference's own apply_params (filter + clamp), with the
    reference's autograd for d/d(param) and d/d(image)."""
    h, w, image_index, gout_seed = 36, 44, 6, 19
    im = torch.clamp(O.synthetic_image(image_index, h, w)[None] * 1.25 - 0.1, 0.0, 1.0)     # clamps / channel ties bite
    gout = torch.randn(im.shape, generator=torch.Generator().manual_seed(gout_seed))
    cases = {}
    for key, name, p in extra_filter_cases():
        pv = p.clone().requires_grad_(True)
        imv = im.clone().requires_grad_(True)
        out = r.image_transformations.apply_params(imv, {name: pv})[-1]
        gp, gim = torch.autograd.grad((out * gout).sum(), [pv, imv], allow_unused=True)
        cases[key] = dict(name=name, param=p.clone(), out=out.detach().clone(),
                          grad_p=(gp.detach().clone() if gp is not None else torch.zeros_like(p)), grad_im=gim.detach().clone())
    torch.save(dict(image_index=image_index, h=h, w=w, gout_seed=gout_seed, image_map="clamp(1.25 * synthetic - 0.1, 0, 1)",
                    cases=cases), os.path.join(GOLDEN_DIR, "filters_extra.pt"))
    print("filters_extra.pt:", len(cases), "cases")


def gen_emonet(r):
    """tests/golden/emonet.pt: the reference's own ValenceArousalLoss on an "EmoNet" checkpoint path (EmoNet.py:33-130: resnet50
    with a 1-output head, Resize(256) + deterministic ten-crop of 224, denorm + ImageNet normalisation, fake arousal column),
    forward + autograd w.r.t. the image on CPU, for a square and a non-square image."""
    sd = O.make_regressor_state_dict(num_classes=1)
    ck = {"state_dict": {("module.model." + k).replace("module.model.fc.", "module.model.last_linear."): v for k, v in sd.items()}}
    path = os.path.join(tempfile.mkdtemp(), "EmoNet_valence_test.pth.tar")
    torch.save(ck, path)
    clf = r.ValenceArousalLoss(path, torch.device("cpu"), 1, is_minimized=True, requires_grad=True)
    target = torch.tensor([[0.3, 0.0]])
    out = dict(image_index=9, image_map="clamp(1.2 * synthetic - 0.1, 0, 1)", target=target, cases={})
    for h, w in ((256, 256), (300, 340)):
        img = torch.clamp(O.synthetic_image(9, h, w)[None] * 1.2 - 0.1, 0.0, 1.0).requires_grad_(True)
        loss = clf(img, target=target)
        g, = torch.autograd.grad(loss, img)
        out["cases"][f"{h}x{w}"] = dict(h=h, w=w, pred=clf.fake_loss_metric.detach().clone(), loss=loss.detach().clone(),
                                        grad_ds=g[0, :, ::4, ::4].clone(), grad_abs_sum=g.abs().sum())
    torch.save(out, os.path.join(GOLDEN_DIR, "emonet.pt"))
    print("emonet.pt written")


def gen_regressor(r):
    sd = O.make_regressor_state_dict()
    clf = _ref_clf(r, sd)
    out = {"weights_probe": {k: sd[k].flatten()[:4].clone() for k in
                             ("conv1.weight", "layer3.2.conv2.weight", "layer4.2.bn3.bias", "fc.weight")}}
    for tag, (h, w) in {"up256": (256, 256), "down512": (512, 512), "same480": (480, 480),
                        "rect300x400": (300, 400)}.items():
        img = O.synthetic_image(3, h, w)[None].requires_grad_(True)
        torch.manual_seed(4242)
        loss = 0.15 * clf(img, target=torch.tensor([[0.9, 0.4]]))
        g, = torch.autograd.grad(loss, img)
        torch.manual_seed(4242)
        oh, ow = O.resize_output_size(h, w, 480)
        offs = O.draw_crop_offsets(1, 1, oh, ow)
        out[tag] = dict(h=h, w=w, image_index=3, offsets=offs[0], pred=clf.fake_loss_metric.detach().clone(),
                        loss=loss.detach().clone(), grad_abs_sum=g.abs().sum(), grad_max=g.abs().max(),
                        grad_ds=g[0, :, ::8, ::8].clone(), grad_patch=g[0, :, :32, :32].clone(),
                        target=torch.tensor([[0.9, 0.4]]), weight_clf=0.15)
        print(tag, out[tag]["pred"], float(loss), float(g.abs().max()))
    torch.save(out, os.path.join(GOLDEN_DIR, "regressor.pt"))


# Start points off the identity presets.  The 50-step run keeps blur at its preset: with a real sigma the optimiser walks
# sigma down through 0 within ~20 steps and the REFERENCE itself turns NaN there (kornia's gaussian with sigma <= 0), and at
# the preset 1e-4 d/d(sigma) underflows to exactly 0 in every implementation (a dead parameter, not a kink).  The short
# 512x512 run does start from a real sigma so that the separable blur and its sigma-gradient are inside the loop test.
# Scale values must be GENERIC: with sx = 1.05 = 21/20 and cx = 3 the sample position (x - cx) / sx + cx is an exact
# integer for every 21st column (a whole family of pixels sits ON the bilinear kink, and the reference's own d/d(scale)
# changes by 4-10 % between an fp32 and an fp64 evaluation of the same expression -- measured, tools/scale_kink_check.py).
KINK_FREE_X0 = dict(sharp=[0.3], scale=[1.0537, 1.0311, 3.3, 5.7])
KINK_FREE_X0_BLUR = dict(sharp=[0.3], blur=[0.8], scale=[1.0537, 1.0311, 3.3, 5.7])


def gen_loop(r, h=256, w=256, num_steps=50, tag="c1", image_index=0, x0_override=None, smooth=False):
    """BASELINE.json configs[0]: one synthetic 256x256 image, random-init regressor, 50 steps, CPU.
    x0_override ({filter: values}) moves the start point off the identity presets: the scale filter's bilinear kink,
    the sharp == 0 and the blur sigma -> 0 branches all sit exactly AT the reference's start values, so a start point
    next to them gives a trajectory on which every parameter gradient is smooth (tag "c1k").
    Also stored: d(loss)/d(x) of every step (one extra autograd.grad per step on the reference's own graph)."""
    sd = O.make_regressor_state_dict()
    clf = _ref_clf(r, sd)
    image = (O.smooth_image if smooth else O.synthetic_image)(image_index, h, w)[None]
    torch.manual_seed(2000 + image_index)
    obj = {"clf": clf, "dis": None, "weight_clf": 0.15, "weight_dis": 0.0, "weight_recon": 0.0, "alpha": 0.1}
    x0, obj = r.optimize_image_param.initialize_parametric(image, obj)
    if x0_override:
        lay = O.param_layout(O.DEFAULT_FILTERS)
        x0 = x0.clone()
        for name, vals in x0_override.items():
            x0[lay[name][0]:lay[name][0] + len(vals)] = torch.tensor(vals)
    obj["target"] = r.optimize_image.get_condition_from_alpha(obj["alpha"], obj["clf"], image)
    del obj["alpha"]
    losses, preds, xs, grads = [], [], [], []
    orig = r.optimize_image_param.objective_function_parametric

    def wrapped(x, **kw):
        xs.append(x.detach().clone())
        l = orig(x, **kw)
        grads.append(torch.autograd.grad(l, x, retain_graph=True)[0].detach().clone())
        losses.append(float(l))
        preds.append(clf.fake_loss_metric.detach().clone()[0])
        return l

    t0 = time.perf_counter()
    best_x = r.optimize_image.optimization(x0, obj, wrapped, learning_rate=0.05, num_steps=num_steps)
    dt = time.perf_counter() - t0
    with torch.no_grad():
        px = r.optimize_image_param.get_params_from_vector(best_x, 1, obj["params"], image.size(2))
        edited = r.image_transformations.apply_params(image, px)[-1]
    torch.manual_seed(2000 + image_index)
    offs = O.draw_crop_offsets(1 + num_steps, 1, 480, 480)
    out = dict(h=h, w=w, num_steps=num_steps, image_index=image_index, smooth=smooth, alpha=0.1, learning_rate=0.05, weight_clf=0.15,
               target=obj["target"].clone(), losses=torch.tensor(losses), preds=torch.stack(preds),
               xs=torch.stack(xs), grads=torch.stack(grads), x0=x0.detach().clone(), best_x=best_x.clone(),
               edited=(edited if h * w <= 256 * 256 else edited[..., ::4, ::4]).clone(), offsets=offs,
               ref_seconds=dt, ref_threads=torch.get_num_threads())
    torch.save(out, os.path.join(GOLDEN_DIR, f"loop_{tag}.pt"))
    print(f"loop_{tag}.pt: {num_steps} steps in {dt:.1f}s; loss {losses[0]:.6f} -> {min(losses):.6f}")


def gen_midu(r):
    import importlib
    ref_harness.install()
    Midu = importlib.import_module("guidance_classifier.MiduClassifier").MiduClassifier
    scores = importlib.import_module("guidance_classifier.guidance_scores")
    out = {}
    for is_sdxl, hw in ((False, 8), (True, 32)):
        torch.manual_seed(0)
        m = Midu._create_midu_classifier("cpu", 2, is_sdxl)
        g = torch.Generator().manual_seed(3000)
        feat = torch.randn(2, 1280, hw, hw, generator=g).requires_grad_(True)
        pred = m(feat)
        loss = scores.valence_arousal_score(pred, "cpu", True, None)
        gf, = torch.autograd.grad(loss, feat)
        out["sdxl" if is_sdxl else "sd"] = dict(seed=0, feat_seed=3000, hw=hw, pred=pred.detach().clone(),
                                                  loss=loss.detach().clone(), grad_abs_sum=gf.abs().sum(),
                                                  grad_slice=gf[:, ::64, ::2, ::2].clone())
    torch.save(out, os.path.join(GOLDEN_DIR, "midu.pt"))
    print("midu.pt written")


MUNIT_SMALL = dict(num_filters=8, num_filters_mlp=32, num_res_blocks=2)     # width / depth overrides of imagenet2imagenet.yaml


def gen_munit(r, h=64, w=64, batch=2, num_steps=3):
    """BASELINE.json configs[2] in miniature, run by the reference itself: its MUNIT `Generator` (imagenet2imagenet.yaml
    with the MUNIT_SMALL overrides so that the state_dict fits a fixture; the full-width network is compared directly
    in tests/test_munit_cpu.py when /root/reference is mounted), `initialize_imaginaire`, `objective_function_imaginaire`
    (weight_clf 0.2, weight_recon 1.0, optimize_image_imaginaire.py:32-37) and `optimization` on a batch of images in
    [-1, 1].  Stored: the generator's state_dict BEFORE any forward (spectral-norm u / v included: every training-mode
    forward advances them), content / style codes, per-step losses, the style codes visited, their gradients, best_x."""
    import importlib
    Config = importlib.import_module("external.imaginaire.config").Config
    RefGen = importlib.import_module("external.imaginaire.generators.munit").Generator
    oii = importlib.import_module("optimize_image_imaginaire")
    cfg = Config(os.path.join(ref_harness.REFERENCE_SRC, "external/imaginaire/imagenet2imagenet.yaml"))
    for k, v in MUNIT_SMALL.items():
        setattr(cfg.gen, k, v)
    torch.manual_seed(0)
    gen = RefGen(cfg.gen, cfg.data)
    state0 = {k: v.detach().clone() for k, v in gen.autoencoder_a.state_dict().items()}      # the loop only uses domain A
    sd = O.make_regressor_state_dict()
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "va_pred_all")
    torch.save(sd, path)
    clf = r.ValenceArousalLoss(path, torch.device("cpu"), 1, is_input_range_0_1=False, is_minimized=True, requires_grad=True)
    image = torch.stack([2.0 * O.synthetic_image(40 + i, h, w) - 1.0 for i in range(batch)])
    torch.manual_seed(2040)
    obj = {"gen": gen, "clf": clf, "dis": None, "gan_loss": None, "weight_clf": 0.2, "weight_dis": 0.0, "weight_recon": 1.0}
    x0, obj = oii.initialize_imaginaire(image, obj)
    obj["target"] = r.optimize_image.get_condition_from_alpha(0.1, clf, image)
    losses, xs, grads, imgs = [], [], [], []
    orig = oii.objective_function_imaginaire

    def wrapped(x, **kw):
        xs.append(x.detach().clone())
        l = orig(x, **kw)
        grads.append(torch.autograd.grad(l, x, retain_graph=True)[0].detach().clone())
        losses.append(float(l))
        return l

    best_x = r.optimize_image.optimization(x0, obj, wrapped, learning_rate=0.05, num_steps=num_steps)
    with torch.no_grad():
        edited = torch.clamp(gen.autoencoder_a.decode(obj["content"].detach(), best_x), -1, 1)
    torch.manual_seed(2040)
    offs = O.draw_crop_offsets(1 + num_steps, batch, 480, 480)
    out = dict(h=h, w=w, batch=batch, num_steps=num_steps, overrides=dict(MUNIT_SMALL), state0=state0, image_index0=40,
               content=obj["content"].clone(), x0=x0.clone(), target=obj["target"].clone(), losses=torch.tensor(losses),
               xs=torch.stack(xs), grads=torch.stack(grads), best_x=best_x.clone(), edited=edited.clone(), offsets=offs,
               crop_seed=2040, weight_clf=0.2, weight_recon=1.0, learning_rate=0.05)
    torch.save(out, os.path.join(GOLDEN_DIR, "munit_small.pt"))
    print("munit_small.pt written: losses", losses)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    r = ref_harness.ref()
    todo = [args.only] if args.only else ["filters", "midu", "regressor", "loop"]
    if "filters" in todo:
        gen_filters(r)
    if "filters_extra" in todo:
        gen_filters_extra(r)
    if "emonet" in todo:
        gen_emonet(r)
    if "midu" in todo:
        gen_midu(r)
    if "regressor" in todo:
        gen_regressor(r)
    if "loop" in todo:
        gen_loop(r)
    if "loopk" in todo:
        gen_loop(r, tag="c1k", x0_override=KINK_FREE_X0)
    if "loops" in todo:       # band-limited image, the reference's own start point: the end-to-end edited-image parity case
        gen_loop(r, tag="c1s", smooth=True)
    if "munit" in todo:
        gen_munit(r)
    if "loop512" in todo:
        gen_loop(r, 512, 512, 3, "c2_3steps", image_index=1, x0_override=KINK_FREE_X0_BLUR)


if __name__ == "__main__":
    main()

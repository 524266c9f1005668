"""CPU: the callers right after the optimisation loop -- `compare_emotions`, the statistics helpers --
against the reference's own functions (imported from /root/reference when mounted) and on fixed inputs otherwise."""
import pytest
import torch

from oracle import ref_harness
from regressor_guided_image_editing_b200.baselines import utils as U
from regressor_guided_image_editing_b200.baselines.run_img_trans import compare_emotions

needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted (GPU box)")


class _FakeLoss:
    """predict_loss_metric stand-in: two deterministic statistics of the image (no regressor needed on CPU)."""

    def predict_loss_metric(self, imgs):
        return torch.stack((imgs.mean((1, 2, 3)), imgs[:, 0].amax((1, 2))), dim=1)


def _images():
    g = torch.Generator().manual_seed(12)
    a = torch.rand(3, 3, 16, 20, generator=g)
    return a, (a * 0.9 + 0.03).clamp(0, 1)


def test_interweave_and_stats_fixed():
    a = torch.arange(6.).view(3, 2)
    out = U.interweave_batch_tensors(a, -a)
    assert out.shape == (6, 2) and torch.equal(out[0::2], a) and torch.equal(out[1::2], -a)
    stats = {}
    U.check_init_stats_adapt(stats, 0.1)
    U.check_init_stats_adapt(stats, 0.1)
    assert list(stats[0.1]) == ["valence", "arousal", "delta_valence", "delta_arousal", "rec_error"]
    img, adapted = _images()
    compare_emotions(_FakeLoss(), img, adapted, None, stats[0.1])
    p0, p1 = _FakeLoss().predict_loss_metric(img), _FakeLoss().predict_loss_metric(adapted)
    assert stats[0.1]["valence"] == [p1[0, 0].item()] and stats[0.1]["arousal"] == [p1[0, 1].item()]
    assert stats[0.1]["delta_valence"] == [(p1 - p0)[0, 0].item()]
    assert abs(stats[0.1]["rec_error"][0] - (adapted - img).abs().mean().item()) < 1e-7


@needs_ref
def test_helpers_equal_reference(capsys):
    import importlib
    ref_harness.install()
    RU = importlib.import_module("baselines.utils")
    RT = importlib.import_module("baselines.run_img_trans")
    img, adapted = _images()
    assert torch.equal(U.interweave_batch_tensors(img, adapted), RU.interweave_batch_tensors(img, adapted))
    s_ref, s_new = {}, {}
    for alpha in (0.1, -0.1):
        RU.check_init_stats_adapt(s_ref, alpha)
        U.check_init_stats_adapt(s_new, alpha)
        for k in range(2):
            RT.compare_emotions(_FakeLoss(), img + 0.01 * k, adapted, ['Valence', 'Arousal'], s_ref[alpha])
            compare_emotions(_FakeLoss(), img + 0.01 * k, adapted, ['Valence', 'Arousal'], s_new[alpha])
    assert s_ref.keys() == s_new.keys()
    for alpha in s_ref:
        assert s_ref[alpha].keys() == s_new[alpha].keys()
        for key in s_ref[alpha]:
            assert s_ref[alpha][key] == pytest.approx(s_new[alpha][key], abs=1e-7), key
    capsys.readouterr()
    RU.print_stats(s_ref)
    want = capsys.readouterr().out
    U.print_stats(s_new)
    assert capsys.readouterr().out == want


@needs_ref
@pytest.mark.parametrize("is_minimized", [True, False])
def test_loss_expressions_equal_reference(is_minimized):
    """The [B, 2]-sized loss expressions of the ValenceArousalLoss mirror against the reference class (both built without
    their regressor: the model is irrelevant to these methods), values and gradients bit for bit; same RNG consumption."""
    import importlib
    ref_harness.install()
    RefLoss = importlib.import_module("baselines.losses.ValenceArousalLoss").ValenceArousalLoss
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss, _MODES

    def bare(cls, mode):
        obj = cls.__new__(cls)
        torch.nn.Module.__init__(obj)
        obj.device, obj.weight, obj.is_minimized = torch.device("cpu"), 0.15, is_minimized
        cols, err = _MODES[mode]
        obj.output_ixs, obj.get_error = list(cols), getattr(obj, err)
        return obj

    g = torch.Generator().manual_seed(3)
    pred = torch.rand(5, 2, generator=g)
    target = torch.rand(5, 2, generator=g)
    for mode in ("va", "valence", "arousal"):
        a, b = bare(RefLoss, mode), bare(ValenceArousalLoss, mode)
        for tgt in (None, target):
            pa = pred.clone().requires_grad_(True)
            pb = pred.clone().requires_grad_(True)
            sel = (lambda t: t) if mode == "va" else (lambda t: t[:, a.output_ixs[0]])
            ea = a.get_error(sel(pa), None if tgt is None else sel(tgt))
            eb = b.get_error(sel(pb), None if tgt is None else sel(tgt))
            assert torch.equal(ea, eb), (mode, tgt is None)
            la, lb = torch.mean(a.weight * ea), torch.mean(b.weight * eb)
            ga, = torch.autograd.grad(la, pa)
            gb, = torch.autograd.grad(lb, pb)
            assert torch.equal(la, lb) and torch.equal(ga, gb)
        torch.manual_seed(11)
        ra = a.get_random_condition_tensor(7)
        torch.manual_seed(11)
        rb = b.get_random_condition_tensor(7)
        assert torch.equal(ra, rb)


@needs_ref
def test_parameter_packing_equals_reference():
    """init_params / initialize_parametric / get_params_from_vector of the mirror against the reference's own functions:
    same dictionaries (keys, order, types, shapes, values), same flat start vector, same gradients through the clamps, and
    the same behaviour on a second call (when the scalar entries have become 0-d tensors)."""
    r = ref_harness.ref().optimize_image_param
    from regressor_guided_image_editing_b200 import optimize_image_param as m
    names = ['exposure', 'saturation', 'tone', 'color', 'contrast', 'sharp', 'blur', 'scale', 'gamma', 'wb', 'bright', 'bw',
             'hue', 'affine', 'not_a_filter']
    for lst in (m.DEFAULT_TRANS, names, list(reversed(names))):
        pr, xr = r.init_params(lst)
        pm, xm = m.init_params(lst)
        assert list(pr) == list(pm) and torch.equal(xr, xm) and xr.dtype == xm.dtype
        for k in pr:
            assert type(pr[k]) is type(pm[k]), k
            assert (pr[k] == pm[k]) if isinstance(pr[k], float) else torch.equal(pr[k], pm[k]), k
        g = torch.Generator().manual_seed(len(lst))
        x = xr + 0.3 * torch.randn(xr.shape, generator=g)
        if "scale" in pr:
            x[-1 if lst[-1] == "scale" else 0] += 0.0
        for call in range(2):
            xa = x.clone().requires_grad_(True)
            xb = x.clone().requires_grad_(True)
            qa = r.get_params_from_vector(xa, 1, pr, 480)
            qb = m.get_params_from_vector(xb, 1, pm, 480)
            assert list(qa) == list(qb)
            tot_a = sum((v.float().sum() if torch.is_tensor(v) else torch.tensor(v)) for v in qa.values())
            tot_b = sum((v.float().sum() if torch.is_tensor(v) else torch.tensor(v)) for v in qb.values())
            for k in qa:
                if torch.is_tensor(qa[k]):
                    assert torch.is_tensor(qb[k]) and qa[k].shape == qb[k].shape and torch.equal(qa[k], qb[k]), (k, call)
                else:
                    assert qa[k] == qb[k], (k, call)
            ga, = torch.autograd.grad(tot_a, xa)
            gb, = torch.autograd.grad(tot_b, xb)
            assert torch.equal(ga, gb), call
    img = torch.rand(1, 3, 8, 8)
    xa, oa = r.initialize_parametric(img, {"clf": None})
    xb, ob = m.initialize_parametric(img, {"clf": None})
    assert torch.equal(xa, xb) and list(oa) == list(ob) and oa["image"] is img and ob["image"] is img


@needs_ref
def test_score_and_crop_mean_equal_reference():
    """guidance_scores.valence_arousal_score (value and gradient) and MeanReplicatedCrops against the reference's own."""
    import importlib
    ref_harness.install()
    R = importlib.import_module("guidance_classifier.guidance_scores")
    RM = importlib.import_module("baselines.models.utilities.MeanReplicatedCrops").MeanReplicatedCrops
    from regressor_guided_image_editing_b200.guidance_classifier.guidance_scores import valence_arousal_score
    from regressor_guided_image_editing_b200.baselines.models.utilities.MeanReplicatedCrops import MeanReplicatedCrops
    g = torch.Generator().manual_seed(1)
    for B in (1, 4):
        p = torch.rand(B, 2, generator=g)
        for minimised in (True, False):
            for ref_value in (None, torch.rand(B, 2, generator=g)):
                a, b = p.clone().requires_grad_(True), p.clone().requires_grad_(True)
                sa = R.valence_arousal_score(a, "cpu", minimised, ref_value)
                sb = valence_arousal_score(b, "cpu", minimised, ref_value)
                ga, = torch.autograd.grad(sa, a)
                gb, = torch.autograd.grad(sb, b)
                assert torch.equal(sa, sb) and torch.equal(ga, gb)
    x = torch.rand(30, 4, generator=g)
    assert torch.equal(MeanReplicatedCrops(10)(x), RM(10)(x))


# ---------------------------------------------------------------------------------------------------------------
# CLIP reconstruction term: everything AROUND the third-party image tower
# ---------------------------------------------------------------------------------------------------------------
class _FakeClipTower(torch.nn.Module):
    """Seeded stand-in with CLIP's call surface (`encode_image` on a [B,3,224,224] batch -> [B,D] features)."""

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(31)
        self.w1 = torch.nn.Parameter(0.2 * torch.randn(8, 3, 16, 16, generator=g))
        self.w2 = torch.nn.Parameter(0.2 * torch.randn(32, 8 * 14 * 14, generator=g))

    def encode_image(self, x):
        assert x.shape[1:] == (3, 224, 224)
        return torch.tanh(torch.nn.functional.conv2d(x, self.w1, stride=16)).flatten(1) @ self.w2.t()


def _clip_pair(lo):
    g = torch.Generator().manual_seed(32)
    a = torch.rand(2, 3, 96, 130, generator=g)
    b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    return (a, b) if lo >= 0 else (2 * a - 1, 2 * b - 1)


def test_clip_loss_needs_a_tower_and_says_so(monkeypatch):
    import sys
    from regressor_guided_image_editing_b200 import _lib
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    monkeypatch.setattr(oi, "CLIP_MODEL", None)
    monkeypatch.setitem(sys.modules, "clip", None)            # `import clip` raises ImportError
    a, b = _clip_pair(0)
    with pytest.raises(_lib.RgieError, match="clip"):
        oi.compute_clip_loss(a, b)


@pytest.mark.parametrize("lo", [0, -1])
def test_clip_loss_properties(monkeypatch, lo):
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    monkeypatch.setattr(oi, "CLIP_MODEL", _FakeClipTower())
    a, b = _clip_pair(lo)
    assert oi.compute_clip_loss(a, a).abs().item() <= 1e-6          # identical images: cosine 1
    l = oi.compute_clip_loss(a, b)
    assert l.dim() == 0 and 0 < l.item() < 2
    # only the FIRST pair counts (the reference indexes [0] after the sum over features)
    b2 = b.clone(); b2[1] = 0.5
    assert oi.compute_clip_loss(a, b2).item() == l.item()


@needs_ref
@pytest.mark.parametrize("lo", [0, -1])
def test_clip_loss_matches_reference(monkeypatch, lo):
    """compute_clip_loss of the reference (baselines/optimize_image.py:152-183: torchvision Resize((224, 224)), Normalize(.5, .5)
    when image1 >= 0, two encode_image calls, normalised features, 1 - cosine of the first pair) with the same tower."""
    import importlib
    ref_harness.install()
    ref_oi = importlib.import_module("baselines.optimize_image")
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    tower = _FakeClipTower()
    monkeypatch.setattr(ref_oi, "CLIP_MODEL", tower)
    monkeypatch.setattr(oi, "CLIP_MODEL", tower)
    a, b = _clip_pair(lo)
    b_r, b_m = b.clone().requires_grad_(True), b.clone().requires_grad_(True)
    l_r, l_m = ref_oi.compute_clip_loss(a, b_r), oi.compute_clip_loss(a, b_m)
    assert torch.equal(l_r.detach(), l_m.detach())
    g_r, = torch.autograd.grad(l_r, b_r)
    g_m, = torch.autograd.grad(l_m, b_m)
    assert torch.equal(g_r, g_m)

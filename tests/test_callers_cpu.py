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

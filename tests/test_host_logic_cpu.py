"""CPU: the host-side logic of the native path that involves no kernel -- BN folding of the regressor weights, the crop-offset
replay of torch's global generator, the learning-rate ramp and Adam's host scalars, the resize output-size rule -- each
against the library code the reference itself calls (torchvision, torch.optim.Adam), so these run on the GPU box as well."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import oracle as O
from regressor_guided_image_editing_b200 import engine, ops
from regressor_guided_image_editing_b200.baselines.models.utilities.ReplicateAndCrop import draw_crop_offsets


# ---------------------------------------------------------------------------------------------------------------
def _folded_forward(arrs, x):
    """resnet50 with conv + bias only (what librgie.so computes), from the arrays in rgie_regressor_create's order."""
    it = iter(torch.from_numpy(a) for a in arrs)

    def conv(x, stride=1, pad=0):
        w, b = next(it), next(it)
        return F.conv2d(x, w, b, stride=stride, padding=pad)

    x = F.max_pool2d(F.relu(conv(x, 2, 3)), 3, 2, 1)
    for li, nb in enumerate(ops.RESNET50_LAYERS, start=1):
        for bi in range(nb):
            stride = 2 if (bi == 0 and li > 1) else 1
            h = F.relu(conv(x))
            h = F.relu(conv(h, stride, 1))                  # torchvision's v1.5 bottleneck: the stride sits in the 3x3 conv
            h = conv(h)
            skip = conv(x, stride) if bi == 0 else x
            x = F.relu(h + skip)
    w, b = next(it), next(it)
    assert next(it, None) is None
    return F.adaptive_avg_pool2d(x, 1).flatten(1) @ w.t() + b


def test_bn_folding_reproduces_the_eval_mode_network():
    """EmotionPredictionModel.py:24-32,46: torchvision resnet50 in eval mode, fc -> 4.  ops.fold_resnet50 folds every
    BatchNorm (running statistics, eps 1e-5) into its conv in float64; the folded conv + bias network must be the same function."""
    from torchvision import models
    sd = O.make_regressor_state_dict()
    net = models.resnet50()
    net.fc = torch.nn.Linear(2048, 4)
    net.load_state_dict(sd)
    net.eval()
    x = torch.randn(2, 3, 96, 80, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        want = net(x)
        arrs = ops.fold_resnet50(sd)
        got = _folded_forward(arrs, x)
    assert len(arrs) == 2 * 53 + 2 and all(a.dtype.name == "float32" and a.flags["C_CONTIGUOUS"] for a in arrs)
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


# ---------------------------------------------------------------------------------------------------------------
def test_crop_offsets_replay_torchvision_random_crop_draws():
    """ReplicateAndCrop.py:30-45 crops with torchvision's RandomCrop: per crop two `torch.randint(...).item()` draws on the
    global generator (RandomCrop.get_params).  Same seed -> same (top, left), in the same order, and the same generator state after."""
    from torchvision import transforms
    img = torch.zeros(3, 480, 500)
    torch.manual_seed(77)
    want = [transforms.RandomCrop.get_params(img, (448, 448))[:2] for _ in range(2 * 10)]
    state_want = torch.get_rng_state()
    torch.manual_seed(77)
    got = draw_crop_offsets(2, 480, 500, 448, 10)
    assert torch.equal(torch.get_rng_state(), state_want)
    assert got.dtype == torch.int32 and got.shape == (2, 10, 2)
    assert got.view(-1, 2).tolist() == [list(t) for t in want]
    assert int(got[..., 0].max()) <= 32 and int(got[..., 1].max()) <= 52 and int(got.min()) >= 0


def test_crop_offsets_edge_cases():
    torch.manual_seed(3)
    before = torch.get_rng_state()
    assert torch.equal(draw_crop_offsets(3, 448, 448), torch.zeros(3, 10, 2, dtype=torch.int32))
    assert torch.equal(torch.get_rng_state(), before)          # no draw at all when the image already has the crop size
    with pytest.raises(ValueError, match="larger than input image size"):
        draw_crop_offsets(1, 447, 480)
    g = torch.Generator().manual_seed(9)
    a = draw_crop_offsets(1, 480, 480, generator=g)
    assert torch.equal(torch.get_rng_state(), before)          # an explicit generator leaves the global one alone
    assert torch.equal(a, draw_crop_offsets(1, 480, 480, generator=torch.Generator().manual_seed(9)))


def test_crop_offsets_equal_the_oracles():
    torch.manual_seed(2000)
    a = O.draw_crop_offsets(3, 2, 480, 480)                    # [calls, batch, 10, 2]
    torch.manual_seed(2000)
    b = torch.stack([draw_crop_offsets(2, 480, 480) for _ in range(3)])
    assert torch.equal(a.to(torch.int32), b)


# ---------------------------------------------------------------------------------------------------------------
def _reference_ramp(step, num_steps, learning_rate, down=0.25, up=0.05):
    import numpy as np                                          # optimize_image.py:69-75, written out
    t = step / num_steps
    r = min(1.0, (1.0 - t) / down)
    r = 0.5 - 0.5 * np.cos(r * np.pi)
    r = r * min(1.0, t / up)
    return learning_rate * r


@pytest.mark.parametrize("num_steps", [50, 100, 7])
def test_lr_ramp(num_steps):
    lrs = [engine.lr_schedule(s, num_steps, 0.05) for s in range(num_steps)]
    assert lrs == [float(_reference_ramp(s, num_steps, 0.05)) for s in range(num_steps)]
    assert lrs[0] == 0.0                                        # the first step never moves x
    assert max(lrs) <= 0.05 + 1e-12
    if num_steps >= 50:
        up_end, down_start = math.ceil(0.05 * num_steps), int(0.75 * num_steps)
        assert all(abs(v - 0.05) <= 1e-12 for v in lrs[up_end:down_start + 1])
        assert all(a >= b for a, b in zip(lrs[down_start:], lrs[down_start + 1:]))


def test_adam_host_scalars_reproduce_torch_adam():
    """The device kernel receives (lr / (1 - b1^k), sqrt(1 - b2^k)) computed on the host in float64, as torch's
    _single_tensor_adam does; stepping with them on CPU must equal torch.optim.Adam with the per-step lr of the ramp."""
    g = torch.Generator().manual_seed(1)
    x_ref = torch.randn(41, generator=g).requires_grad_(True)
    x, m, v = x_ref.detach().clone(), torch.zeros(41), torch.zeros(41)
    opt = torch.optim.Adam([x_ref], betas=(0.9, 0.999), lr=0.05)
    for k in range(1, 13):
        lr = engine.lr_schedule(k - 1, 12, 0.05)
        grad = torch.randn(41, generator=g) * (0.0 if k == 4 else 1.0)
        grad[7] = 0.0                                           # an element whose gradient is always 0 never moves
        for gr in opt.param_groups:
            gr["lr"] = lr
        x_ref.grad = grad.clone()
        opt.step()
        step_size, bc2_sqrt = ops.adam_scalars(lr, k)
        m.lerp_(grad, 1 - 0.9)
        v.mul_(0.999).addcmul_(grad, grad, value=1 - 0.999)
        x.addcdiv_(m, (v.sqrt() / bc2_sqrt).add_(1e-8), value=-step_size)
        assert torch.equal(x, x_ref.detach()), k
    assert x[7].item() == x_ref.detach()[7].item()


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hw", [(256, 256), (512, 512), (300, 400), (400, 300), (480, 480), (481, 777), (1024, 683), (97, 1001)])
def test_resize_output_size_is_torchvisions(hw):
    from torchvision.transforms import functional as tvF
    h, w = hw
    out = tvF.resize(torch.zeros(1, 3, h, w), 480, antialias=True)
    assert ops.resize_output_size(h, w, 480) == tuple(out.shape[-2:])
    assert O.resize_output_size(h, w, 480) == tuple(out.shape[-2:])
    assert O.resize_aa(torch.zeros(1, 3, h, w), 480).shape == out.shape

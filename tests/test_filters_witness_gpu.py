"""GPU: the native filter kernels against OTHER people's implementations of the same published algorithms (torchvision 0.26,
scipy), not against this repository's reading of kornia (`oracle/kornia_shim.py`).  Companion of
tests/test_shim_witness_cpu.py; the witnesses run on the CPU (torchvision's convolutions on a GPU would use TF32).
Reference call sites: /root/reference/src/baselines/image_transformations/image_transformations.py:98,109,122,173,185,195,221.
"""
import math

import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

tvF = pytest.importorskip("torchvision.transforms.functional")
tvFT = pytest.importorskip("torchvision.transforms._functional_tensor")
ndimage = pytest.importorskip("scipy.ndimage")


def _mirror():
    from regressor_guided_image_editing_b200.baselines.image_transformations import image_transformations as IT
    return IT


def _image(seed, b=2, h=72, w=88):
    im = torch.stack([O.synthetic_image(seed + i, h, w) for i in range(b)])
    return torch.clamp(im * 1.25 - 0.1, 0.0, 1.0)          # clamps and channel ties bite


def _native(fn, im, p):
    return fn(im.cuda(), torch.as_tensor(p, dtype=torch.float32).cuda()).cpu()


@pytest.mark.parametrize("factor", [0.0, 0.5, 1.0, 1.7])
def test_sharpen_kernel_equals_torchvision(factor):
    im = _image(20)
    err = (_native(_mirror().apply_sharpening, im, factor) - tvF.adjust_sharpness(im, factor)).abs().max().item()
    assert err <= 5e-6, err


@pytest.mark.parametrize("sigma", [0.7, 1.0, 2.0, 5.0])
def test_blur_kernel_equals_torchvision(sigma):
    im = _image(21)
    err = (_native(_mirror().apply_gaussian_blur, im, sigma) - tvF.gaussian_blur(im, [25, 25], [sigma, sigma])).abs().max().item()
    assert err <= 5e-6, err


@pytest.mark.parametrize("factor", [0.5, 1.0, 1.8])
def test_contrast_kernel_equals_torchvision_up_to_gray_weight(factor):
    im = _image(22)
    mine = _native(_mirror().apply_contrast, im, factor)
    for b in range(im.shape[0]):
        err = (mine[b] - tvF.adjust_contrast(im[b], factor)).abs().max().item()
        assert err <= 1.1e-4 * abs(1 - factor) + 3e-6, err          # torchvision's gray weight is 0.2989, kornia's 0.299


@pytest.mark.parametrize("factor", [0.0, 0.35, 1.0, 1.7])
def test_saturation_kernel_is_the_hsv_scaling_of_torchvisions_converters(factor):
    im = _image(23)
    h, s, v = tvFT._rgb2hsv(im).unbind(1)
    theirs = tvFT._hsv2rgb(torch.stack((h, torch.clamp(s * factor, 0, 1), v), 1))
    err = (_native(_mirror().apply_saturation, im, factor) - theirs).abs().max().item()
    assert err <= 1e-5, err


@pytest.mark.parametrize("factor", [0.7, -2.1, 3.0])
def test_hue_kernel_equals_torchvision(factor):
    im = _image(24)
    err = (_native(_mirror().apply_hue, im, factor) - tvF.adjust_hue(im, factor / (2 * math.pi))).abs().max().item()
    assert err <= 2e-5, err


@pytest.mark.parametrize("gamma", [1.0, 1.3, 2.2])
def test_gamma_kernel_equals_torchvision(gamma):
    im = _image(25)
    err = (_native(_mirror().apply_gamma, im, gamma) - tvF.adjust_gamma(im, gamma)).abs().max().item()
    assert err <= 5e-6, err


@pytest.mark.parametrize("sx,sy,cx,cy", [(1.2371, 1.1113, 9.37, 14.21), (1.5311, 1.0173, 30.19, 5.23), (2.3, 1.9, 40.0, 33.0)])
def test_scale_kernel_equals_scipy_inverse_mapping(sx, sy, cx, cy):
    """kornia's matrix [[sx, 0, (1 - sx) cx], [0, sy, (1 - sx) cy]] (its convention), then the inverse-mapped bilinear warp with
    zeros outside done by scipy in float64."""
    im = _image(26)
    mine = _mirror().apply_scale(im.cuda(), torch.tensor([[sx, sy, cx, cy]] * im.shape[0]).cuda()).cpu()
    tx, ty = (1 - sx) * cx, (1 - sx) * cy
    theirs = np.empty(im.shape, dtype=np.float64)
    src = im.double().numpy()
    for b in range(im.shape[0]):
        for c in range(3):
            theirs[b, c] = ndimage.affine_transform(src[b, c], np.array([1 / sy, 1 / sx]), offset=np.array([-ty / sy, -tx / sx]),
                                                    order=1, mode="grid-constant", cval=0.0)
    err = np.abs(mine.double().numpy() - theirs).max()
    assert err <= 2e-4, err                                  # fp32 sample positions: 1e-5 px times slopes of up to 1 per px

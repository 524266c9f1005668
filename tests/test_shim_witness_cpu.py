"""CPU: independent third-party witnesses for oracle/kornia_shim.py.

kornia 0.8.2 (the reference's pin, /root/reference/uv.lock:588-590) exists neither in the reference tree nor in this image, so
the shim cannot be diffed against kornia itself ("parity unpinned" for its nine functions).  What CAN be
checked offline is that the shim's restatement of each PUBLISHED algorithm agrees with OTHER implementations of the same
algorithm that are installed here and were written by other people:

* torchvision 0.26 (`transforms.functional` and its private HSV helpers): PIL's Sharpness enhancer (kornia documents its
  `sharpness` as that algorithm), the sampled-Gaussian separable blur with reflect padding, the HSV hexcone round trip behind
  saturation / hue, the gray-mean contrast blend, gain * x ** gamma;
* Python's `colorsys` for the HSV definitions on single pixels;
* scipy.ndimage.map_coordinates (order 1, `grid-constant`) for the inverse-mapped bilinear warp with zero padding that
  `warp_affine` reaches through normalise -> invert -> affine_grid -> grid_sample(align_corners=True).

What these witnesses do NOT cover is stated per test: conventions that are kornia's own (hue in radians, additive
brightness, (sigma_y, sigma_x) order, `get_rotation_matrix2d` using (1 - sx) for BOTH translations) stay a reading of kornia.
The call sites are /root/reference/src/baselines/image_transformations/image_transformations.py:98,109,122,143,173,185,195,205,221.
"""
import colorsys
import math

import numpy as np
import pytest
import torch

from oracle import kornia_shim as K

tvF = pytest.importorskip("torchvision.transforms.functional")
tvFT = pytest.importorskip("torchvision.transforms._functional_tensor")
ndimage = pytest.importorskip("scipy.ndimage")


def _image(seed, b=2, h=37, w=45, smooth=False):
    g = torch.Generator().manual_seed(seed)
    im = torch.rand(b, 3, h, w, generator=g)
    if smooth:      # some saturated / flat regions: the clamp branches and the delta == 0 branch of the hexcone
        im = torch.clamp(1.4 * im - 0.2, 0, 1)
        im[:, :, :4, :5] = 0.5
    return im


@pytest.mark.parametrize("factor", [0.0, 0.3, 1.0, 1.7, 3.0])
def test_sharpness_equals_torchvision_pil_algorithm(factor):
    """F6.  torchvision: blend(img, smoothed-with-[[1,1,1],[1,5,1],[1,1,1]]/13-interior-only, factor), clamped."""
    im = _image(1, smooth=True)
    mine = K.sharpness(im, torch.tensor(factor))
    theirs = tvF.adjust_sharpness(im, factor)
    assert (mine - theirs).abs().max().item() <= 1e-6


@pytest.mark.parametrize("sigma", [0.6, 1.0, 2.0, 4.5])
def test_gaussian_blur_equals_torchvision(sigma):
    """F7.  25 taps of exp(-x^2 / 2 sigma^2) / sum, reflect padding, depthwise cross-correlation."""
    im = _image(2)
    mine = K.gaussian_blur2d(im, (25, 25), torch.tensor([[sigma, sigma]]))
    theirs = tvF.gaussian_blur(im, [25, 25], [sigma, sigma])
    assert (mine - theirs).abs().max().item() <= 2e-6


def test_gaussian_blur_anisotropic_axis_order():
    """kornia: kernel_size = (ky, kx), sigma = (sigma_y, sigma_x); torchvision: [kx, ky], [sigma_x, sigma_y].  Only the ORDER
    is kornia's convention; the reference always passes the same sigma twice (image_transformations.py:120-121)."""
    im = _image(3)
    mine = K.gaussian_blur2d(im, (25, 25), torch.tensor([[0.7, 2.5]]))           # sigma_y = 0.7, sigma_x = 2.5
    theirs = tvF.gaussian_blur(im, [25, 25], [2.5, 0.7])
    assert (mine - theirs).abs().max().item() <= 2e-6


def test_gaussian_blur_per_image_sigma():
    im = _image(4)
    mine = K.gaussian_blur2d(im, (25, 25), torch.tensor([[0.8, 0.8], [3.0, 3.0]]))
    for b, s in enumerate((0.8, 3.0)):
        theirs = tvF.gaussian_blur(im[b:b + 1], [25, 25], [s, s])
        assert (mine[b:b + 1] - theirs).abs().max().item() <= 2e-6


def test_hsv_round_trip_against_torchvision_and_colorsys():
    """F2 / hue.  Same hexcone model; kornia's h is an angle in [0, 2 pi), torchvision's and colorsys' a fraction of a turn."""
    im = _image(5, smooth=True)
    hsv = K.rgb_to_hsv(im)
    tv = tvFT._rgb2hsv(im)
    dh = (hsv[:, 0] / (2 * math.pi) - tv[:, 0]).abs()
    dh = torch.minimum(dh, 1 - dh)                                  # hue is periodic
    chroma = im.max(1).values - im.min(1).values
    assert dh[chroma > 1e-3].max().item() <= 1e-5                   # hue is ill-defined (and unused: s = 0) on gray pixels
    # kornia: s = delta / (max + 1e-8) (its eps, a kornia convention); torchvision: delta / max -> relative 1e-8 / max
    assert ((hsv[:, 1] - tv[:, 1]).abs() <= 1e-6 + 2e-8 / tv[:, 2].clamp_min(1e-8)).all()
    assert torch.equal(hsv[:, 2], tv[:, 2])
    assert (K.hsv_to_rgb(hsv) - im).abs().max().item() <= 2e-6
    assert (K.hsv_to_rgb(hsv) - tvFT._hsv2rgb(tv)).abs().max().item() <= 2e-6
    px = im[0, :, ::9, ::11].reshape(3, -1).T.tolist()
    mine = hsv[0, :, ::9, ::11].reshape(3, -1).T.tolist()
    for (r, g, b), (h, s, v) in zip(px, mine):
        ch, cs, cv = colorsys.rgb_to_hsv(r, g, b)
        if max(r, g, b) - min(r, g, b) > 1e-3:
            d = abs(h / (2 * math.pi) - ch)
            assert min(d, 1 - d) <= 1e-5
        assert abs(s - cs) <= 1e-6 + 2e-8 / max(cv, 1e-8) and abs(v - cv) <= 1e-7


@pytest.mark.parametrize("factor", [0.0, 0.4, 1.0, 1.6, 5.0])
def test_saturation_is_an_hsv_scaling(factor):
    """F2.  kornia.enhance.adjust_saturation scales S in HSV (torchvision's own adjust_saturation is a gray blend, a different
    algorithm, so the witness is the same scaling done with torchvision's converters)."""
    im = _image(6, smooth=True)
    mine = K.adjust_saturation(im, torch.tensor(factor))
    h, s, v = tvFT._rgb2hsv(im).unbind(1)
    theirs = tvFT._hsv2rgb(torch.stack((h, torch.clamp(s * factor, 0, 1), v), 1))
    assert (mine - theirs).abs().max().item() <= 3e-6


@pytest.mark.parametrize("factor", [-2.0, -0.5, 0.0, 0.9, 3.0])
def test_hue_shift_equals_torchvision(factor):
    """hue (8f rank 1).  kornia: radians in [-pi, pi]; torchvision: turns in [-0.5, 0.5]."""
    im = _image(7, smooth=True)
    mine = K.adjust_hue(im, torch.tensor(factor))
    theirs = tvF.adjust_hue(im, factor / (2 * math.pi))
    # a pixel whose shifted hue lands within round-off of a sextant boundary may take either branch: both give the same colour
    assert (mine - theirs).abs().max().item() <= 5e-6


@pytest.mark.parametrize("factor", [0.0, 0.5, 1.0, 1.8])
def test_contrast_equals_torchvision_up_to_the_gray_weights(factor):
    """F5.  blend(img, mean of the gray image, factor); torchvision's gray uses 0.2989 where kornia documents 0.299."""
    im = _image(8, smooth=True)
    mine = K.adjust_contrast_with_mean_subtraction(im, torch.tensor(factor))
    for b in range(im.shape[0]):        # the shim's mean is per image (kornia: mean((-2, -1), keepdim)); torchvision's too when called per image
        theirs = tvF.adjust_contrast(im[b], factor)
        assert (mine[b] - theirs).abs().max().item() <= 1.1e-4 * abs(1 - factor) + 1e-6
    # with the same weights the blend itself is exact
    mu = (0.299 * im[:, 0] + 0.587 * im[:, 1] + 0.114 * im[:, 2]).mean((-2, -1))[:, None, None, None]
    assert (mine - torch.clamp(factor * im + (1 - factor) * mu, 0, 1)).abs().max().item() <= 1e-6


@pytest.mark.parametrize("gamma,gain", [(0.5, 1.0), (1.0, 1.0), (2.2, 1.0), (1.5, 1.3)])
def test_gamma_equals_torchvision(gamma, gain):
    im = _image(9)
    mine = K.adjust_gamma(im, torch.tensor(gamma), gain)
    theirs = tvF.adjust_gamma(im, gamma, gain)
    assert (mine - theirs).abs().max().item() <= 1e-6


def _scipy_warp(im, minv, shift):
    """out[y, x] = bilinear(in, minv @ (y, x) + shift) with zero padding that takes part in the interpolation (float64)."""
    out = np.empty_like(im)
    for b in range(im.shape[0]):
        for c in range(im.shape[1]):
            out[b, c] = ndimage.affine_transform(im[b, c], minv, offset=shift, order=1, mode="grid-constant", cval=0.0)
    return out


@pytest.mark.parametrize("sx,sy,cx,cy", [(1.0, 1.0, 0.0, 0.0), (1.3, 1.3, 10.0, 7.0), (1.7, 1.2, 22.0, 18.0), (2.5, 1.05, 3.0, 30.0)])
def test_scale_warp_equals_scipy_inverse_mapping(sx, sy, cx, cy):
    """F8.  With kornia's matrix M = [[sx, 0, (1 - sx) cx], [0, sy, (1 - sx) cy]] (its get_rotation_matrix2d puts alpha = M00 in
    both translations: kornia's convention, NOT witnessed here), `scale` must be the pixel-space inverse mapping
    src = M^-1 (dst) sampled bilinearly with zeros outside -- the normalise / invert / affine_grid / grid_sample chain
    restated by an independent resampler."""
    im = _image(10, h=40, w=48).double()
    mine = K.scale(im, torch.tensor([[sx, sy]], dtype=torch.float64), torch.tensor([[cx, cy]], dtype=torch.float64))
    tx, ty = (1 - sx) * cx, (1 - sx) * cy
    # scipy indexes (row, col) = (y, x): src_y = (y - ty) / sy, src_x = (x - tx) / sx
    theirs = _scipy_warp(im.numpy(), np.array([1 / sy, 1 / sx]), np.array([-ty / sy, -tx / sx]))
    assert np.abs(mine.numpy() - theirs).max() <= 1e-9


def test_general_affine_warp_equals_scipy_inverse_mapping():
    """affine (8f rank 1): kornia.geometry.transform.affine(im, M) = inverse-mapped bilinear warp, zeros outside."""
    im = _image(11, h=40, w=48).double()
    a = math.radians(17.0)
    M = torch.tensor([[[1.2 * math.cos(a), 1.2 * math.sin(a), 3.5], [-0.9 * math.sin(a), 0.9 * math.cos(a), -2.25]]], dtype=torch.float64)
    mine = K.affine(im, M)
    A = np.array([[M[0, 0, 0], M[0, 0, 1]], [M[0, 1, 0], M[0, 1, 1]]], dtype=np.float64)     # acts on (x, y)
    t = np.array([M[0, 0, 2], M[0, 1, 2]], dtype=np.float64)
    Ainv = np.linalg.inv(A)
    P = np.array([[0.0, 1.0], [1.0, 0.0]])                                                  # (y, x) <-> (x, y)
    theirs = _scipy_warp(im.numpy(), P @ Ainv @ P, -(P @ Ainv @ t))
    assert np.abs(mine.numpy() - theirs).max() <= 1e-9


def test_scale_fp32_path_stays_within_round_off_of_the_float64_witness():
    """The loop runs the warp in fp32 (linalg.inv, affine_grid and grid_sample in fp32): bound against the fp64 witness."""
    im = _image(12, h=64, w=64)
    sx, sy, cx, cy = 1.37, 1.11, 20.0, 41.0
    mine = K.scale(im, torch.tensor([[sx, sy]]), torch.tensor([[cx, cy]]))
    tx, ty = (1 - sx) * cx, (1 - sx) * cy
    theirs = _scipy_warp(im.double().numpy(), np.array([1 / sy, 1 / sx]), np.array([-ty / sy, -tx / sx]))
    assert np.abs(mine.double().numpy() - theirs).max() <= 5e-5

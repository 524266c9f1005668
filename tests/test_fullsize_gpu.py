"""GPU, at BASELINE.json's full configs[1] size (64 images of 512x512, micro-batches of 32 images = 320 crops of 448x448):
the oracle would need hours on these shapes, so parity is checked through properties that do not depend on size --
identity presets, linearity of every backward pass in its incoming gradient, row independence of the regressor (a
permutation of the batch permutes the outputs bit for bit, which is what sharding across GPUs relies on), exact scaling of
the input gradient by a power of two, and the whole engine run in one batch of 64 vs two shards of 32."""
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
B, H = 64, 512


@pytest.fixture(scope="module")
def images():
    g = torch.Generator().manual_seed(123)
    low = torch.rand(B, 3, H // 8, H // 8, generator=g)
    img = torch.nn.functional.interpolate(low, size=(H, H), mode="bilinear", align_corners=False)
    return (0.05 + 0.9 * img).clamp(0, 1).contiguous().to(DEV)


def _mirror():
    from regressor_guided_image_editing_b200.baselines.image_transformations import image_transformations as IT
    return IT


def test_identity_presets_leave_full_size_batch_unchanged(images):
    IT = _mirror()
    p = O.get_params_from_vector(O.init_x0(), O.DEFAULT_FILTERS, H)
    for name in ("exposure", "tone", "color", "contrast", "blur", "saturation", "scale"):
        v = p[name].to(DEV) if isinstance(p[name], torch.Tensor) else p[name]
        out = IT._DISPATCH[name](images, v)
        tol = 1e-4 if name == "scale" else 2e-6
        assert (out - images).abs().max().item() <= tol, name


@pytest.mark.parametrize("name,n,val", [("exposure", 1, [0.3]), ("saturation", 1, [1.3]), ("tone", 8, None), ("color", 24, None),
                                        ("contrast", 1, [1.2]), ("sharp", 1, [0.5]), ("blur", 1, [1.5]),
                                        ("scale", 4, [1.1, 1.2, 200.0, 300.0])])
def test_filter_backward_is_linear_in_gout_at_full_size(images, name, n, val):
    from regressor_guided_image_editing_b200 import ops, _lib
    kind = _lib.FILTER_KINDS[name]
    g = torch.Generator().manual_seed(5)
    if val is None:
        val = (1.0 + 0.1 * torch.randn(n, generator=g)).tolist()
    p = torch.tensor(val, device=DEV).repeat(B, 1).contiguous()
    g1 = torch.randn(images.shape, generator=g).to(DEV)
    g2 = torch.randn(images.shape, generator=g).to(DEV)
    ws = ops.filter_workspace(B, H, H, DEV)

    def bwd(gout):
        gp = torch.empty(B, n, device=DEV)
        gin = ops.filter_bwd(kind, images, gout.contiguous(), p, n, gp, n, ws=ws).clone()
        return gin, gp
    a_in, a_p = bwd(g1)
    b_in, b_p = bwd(g2)
    c_in, c_p = bwd(g1 + 2.0 * g2)
    scale_in = (a_in.abs() + 2 * b_in.abs()).max().item() + 1e-12
    assert (c_in - (a_in + 2 * b_in)).abs().max().item() <= 1e-5 * scale_in
    scale_p = (a_p.abs() + 2 * b_p.abs()).max().item() + 1e-12
    assert (c_p - (a_p + 2 * b_p)).abs().max().item() <= 2e-4 * scale_p      # sums of 786k products: fp32 reduction order


def test_regressor_rows_are_independent_and_backward_scales_exactly():
    """320 crops per launch (the bench micro-batch): permuting the images permutes logits and image gradients bit for bit;
    doubling the incoming gradient doubles the image gradient exactly (every op of the backward pass is linear, and a factor
    of two is exact in bf16 and fp32)."""
    from regressor_guided_image_editing_b200 import ops
    nb = 32
    sd = O.make_regressor_state_dict()
    reg = ops.Regressor(sd, max_crops=nb * 10, precision="bf16")
    g = torch.Generator().manual_seed(9)
    img = torch.rand(nb, 3, 480, 480, generator=g).to(DEV)
    offs = torch.randint(0, 33, (nb, 10, 2), generator=g, dtype=torch.int32).to(DEV)
    dl = (torch.randn(nb * 10, 4, generator=g) * 1e-2).to(DEV)
    logits = reg.forward(img, offs).clone()
    dimg = reg.backward(dl, torch.empty_like(img)).clone()
    assert torch.isfinite(logits).all() and torch.isfinite(dimg).all() and dimg.abs().max().item() > 0

    perm = torch.randperm(nb, generator=g).to(DEV)
    logits_p = reg.forward(img[perm].contiguous(), offs[perm].contiguous()).clone()
    dl_p = dl.view(nb, 10, 4)[perm].reshape(nb * 10, 4).contiguous()
    dimg_p = reg.backward(dl_p, torch.empty_like(img)).clone()
    assert torch.equal(logits_p.view(nb, 10, -1), logits.view(nb, 10, -1)[perm])
    assert torch.equal(dimg_p, dimg[perm])

    reg.forward(img, offs)
    dimg2 = reg.backward((2.0 * dl).contiguous(), torch.empty_like(img))
    assert torch.equal(dimg2, 2.0 * dimg)


def test_engine_batch_of_64_equals_two_shards_of_32():
    """configs[1] shape, 3 optimisation steps: one engine with 64 problems (two micro-batches) and two engines with 32
    problems each (what two GPUs would run) give bit-identical losses, parameters and edited images."""
    from regressor_guided_image_editing_b200 import engine
    steps = 3
    sd = O.make_regressor_state_dict()
    imgs = torch.stack([O.synthetic_image(i, H, H) for i in range(B)]).to(DEV)
    g = torch.Generator().manual_seed(2000)
    offs = torch.randint(0, 33, (1 + steps, B, 10, 2), generator=g, dtype=torch.int32).to(DEV)
    full = engine.ParametricEditEngine(sd, batch=B, height=H, width=H, num_steps=steps, precision="bf16", micro_batch=32)
    r = full.run(imgs, offs)
    del full
    half = engine.ParametricEditEngine(sd, batch=32, height=H, width=H, num_steps=steps, precision="bf16", micro_batch=32)
    for s in range(2):
        sl = slice(32 * s, 32 * (s + 1))
        rs = half.run(imgs[sl].contiguous(), offs[:, sl].contiguous())
        assert torch.equal(rs["losses"], r["losses"][:, sl]), f"shard {s}: losses"
        assert torch.equal(rs["best_x"], r["best_x"][sl]), f"shard {s}: best_x"
        assert torch.equal(rs["edited"], r["edited"][sl]), f"shard {s}: edited images"
    assert torch.isfinite(r["losses"]).all()
    assert (r["edited"] - imgs).abs().max().item() > 0          # the optimisation moved the images


def test_crops_of_a_320_crop_handle_equal_the_same_image_alone():
    """Crops are independent: image b of a 320-crop handle must come out bit-identical (logits and image gradient) to the same
    image run alone on a 10-crop handle.  This crosses every tiling that packs several crops into one launch -- the bands of
    the fused conv1 + max-pool kernel, the 8-line tiles of conv3_hshare_kernel and the flat 128-row tiles, which all straddle
    image boundaries in the big handle and not in the small one.  Runs in a child process (its own 45 GB workspace)."""
    import os, subprocess, sys
    code = r"""
import torch
from oracle import oracle as O
from regressor_guided_image_editing_b200 import ops
DEV = 'cuda'
nb = 32
sd = O.make_regressor_state_dict()
g = torch.Generator().manual_seed(11)
img = torch.rand(nb, 3, 480, 480, generator=g).to(DEV)
offs = torch.randint(0, 33, (nb, 10, 2), generator=g, dtype=torch.int32).to(DEV)
dl = (torch.randn(nb * 10, 4, generator=g) * 1e-2).to(DEV)
big = ops.Regressor(sd, max_crops=nb * 10, precision='bf16')
logits = big.forward(img, offs).clone()
dimg = big.backward(dl, torch.empty_like(img)).clone()
del big
one = ops.Regressor(sd, max_crops=10, precision='bf16')
for b in (0, 7, 8, 31):
    lg = one.forward(img[b:b + 1].contiguous(), offs[b:b + 1].contiguous()).clone()
    dg = one.backward(dl[10 * b:10 * b + 10].contiguous(), torch.empty_like(img[b:b + 1]))
    assert torch.equal(lg, logits[10 * b:10 * b + 10]), f'image {b}: logits'
    assert torch.equal(dg[0], dimg[b]), f'image {b}: image gradient'
print('CROPS_INDEPENDENT_OK')
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=root), cwd=root,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CROPS_INDEPENDENT_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]

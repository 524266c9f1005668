"""CPU: the MUNIT generator mirror (regressor_guided_image_editing_b200/external/imaginaire/generators/munit.py) -- the
differentiable producer upstream of the native regressor in BASELINE.json configs[2].

  * against the reference's own module from the same state_dict (when /root/reference is mounted): full-width network of
    imagenet2imagenet.yaml, encode / decode / d(decode)/d(style) bit for bit over two consecutive training-mode forwards
    (each advances the spectral-norm power iteration);
  * against the reference-generated golden of the latent loop in miniature (tests/golden/munit_small.pt,
    oracle/gen_golden.py::gen_munit): state_dict loads by name, content / style codes and the decoded best image equal
    the reference's, and the ORACLE's latent loop over this mirror reproduces the reference's per-step losses, visited
    style codes, gradients and best_x.
"""
import os

import pytest
import torch

from oracle import oracle as O
from oracle import ref_harness
from regressor_guided_image_editing_b200.external.imaginaire.generators import munit as M

needs_ref = pytest.mark.skipif(not ref_harness.available(), reason="/root/reference not mounted (GPU box)")


def test_default_configuration_shapes():
    torch.manual_seed(0)
    ae = M.Generator().autoencoder_a
    names = set(ae.state_dict())
    for k in ("style_encoder.model.6.weight", "content_encoder.model.3.layers.norm.weight",
              "content_encoder.model.7.conv_block_1.layers.conv.weight_orig", "decoder.decoder.0.conv_block_0.layers.norm.fc.layers.conv.weight",
              "decoder.decoder.9.layers.conv.weight_u", "decoder.decoder.10.layers.conv.weight", "mlp.model.1.layers.conv.bias"):
        assert k in names, k
    img = 2 * O.synthetic_image(3, 32, 32)[None] - 1
    with torch.no_grad():
        content, style = ae.encode(img)
        out = ae.decode(content, style)
    assert content.shape == (1, 256, 4, 4) and style.shape == (1, 8, 1, 1) and out.shape == img.shape


@needs_ref
def test_full_width_network_equals_reference_module():
    import importlib
    ref_harness.install()
    Config = importlib.import_module("external.imaginaire.config").Config
    RefGen = importlib.import_module("external.imaginaire.generators.munit").Generator
    cfg = Config(os.path.join(ref_harness.REFERENCE_SRC, "external/imaginaire/imagenet2imagenet.yaml"))
    torch.manual_seed(0)
    ref = RefGen(cfg.gen, cfg.data)
    mine = M.Generator(cfg.gen, cfg.data)
    assert set(ref.state_dict()) == set(mine.state_dict())
    mine.load_state_dict(ref.state_dict())
    default = M.Generator()                                     # the built-in defaults ARE the yaml's gen section
    assert {k: tuple(v.shape) for k, v in default.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    img = torch.stack([2 * O.synthetic_image(7 + i, 64, 64) - 1 for i in range(2)])
    for _ in range(2):
        c1, s1 = ref.autoencoder_a.encode(img)
        c2, s2 = mine.autoencoder_a.encode(img)
        assert torch.equal(c1, c2) and torch.equal(s1, s2)
        a = s1.detach().clone().requires_grad_(True)
        b = s1.detach().clone().requires_grad_(True)
        o1 = ref.autoencoder_a.decode(c1.detach(), a)
        o2 = mine.autoencoder_a.decode(c2.detach(), b)
        assert torch.equal(o1, o2)
        g1, = torch.autograd.grad(o1.clamp(-1, 1).square().mean(), a)
        g2, = torch.autograd.grad(o2.clamp(-1, 1).square().mean(), b)
        assert torch.equal(g1, g2)


def _small_generator(gold):
    gen = M.Generator(dict(M.DEFAULT_GEN_CFG, **gold["overrides"]))
    gen.autoencoder_a.load_state_dict(gold["state0"])
    return gen


def test_latent_loop_golden_from_reference(golden_dir):
    gold = torch.load(os.path.join(golden_dir, "munit_small.pt"))
    gen = _small_generator(gold)
    B, h, w, steps = gold["batch"], gold["h"], gold["w"], gold["num_steps"]
    image = torch.stack([2.0 * O.synthetic_image(gold["image_index0"] + i, h, w) - 1.0 for i in range(B)])
    sd = O.make_regressor_state_dict()
    offs = gold["offsets"]
    with torch.no_grad():
        content, style = gen.autoencoder_a.encode(image)
    assert torch.equal(content, gold["content"]) and torch.equal(style, gold["x0"])
    with torch.no_grad():
        pred0 = O.regressor_predict(image, sd, offs[0], normalize=False)[:, [0, 1]]
    target = O.get_condition_from_alpha(pred0, 0.1)
    assert (target - gold["target"]).abs().max().item() <= 1e-6
    xs, grads = [], []

    def objective(x, s):
        xs.append(x.detach().clone())
        loss = O.objective_imaginaire(x, gen, content, sd, offs[1 + s], target, gold["weight_clf"], gold["weight_recon"])[0]
        grads.append(torch.autograd.grad(loss, x, retain_graph=True)[0].detach().clone())
        return loss

    out = O.optimize_generic(style, objective, gold["learning_rate"], steps)
    assert (out["losses"] - gold["losses"]).abs().max().item() <= 2e-6, (out["losses"], gold["losses"])
    assert (torch.stack(xs) - gold["xs"]).abs().max().item() <= 1e-5
    gscale = gold["grads"].abs().max().item()
    assert (torch.stack(grads) - gold["grads"]).abs().max().item() <= 1e-4 * gscale
    assert (out["best_x"] - gold["best_x"]).abs().max().item() <= 1e-5
    with torch.no_grad():
        edited = torch.clamp(gen.autoencoder_a.decode(content, out["best_x"]), -1, 1)
    assert (edited - gold["edited"]).abs().max().item() <= 1e-4

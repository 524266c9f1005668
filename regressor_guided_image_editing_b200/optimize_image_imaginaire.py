"""Drop-in for the hot-path functions of src/optimize_image_imaginaire.py: `initialize_imaginaire` (:112-117) and
`objective_function_imaginaire` (:120-145) -- latent (style-code) optimisation through a MUNIT generator.

The generator (`gen.autoencoder_a.encode/decode`, external/imaginaire) stays a PyTorch module (its
kernels are a later row); the regressor term `clf(img, target)` is the native ValenceArousalLoss (autograd Function over
librgie.so) and the Adam update / best-x tracking of `baselines.optimize_image.optimization` are the fused native
kernels, so the loop's regressor forward + input-gradient backward -- ~90 % of its FLOPs -- run on the tcgen05 path.
Script glue (main, checkpoints, discriminator loading, file I/O) is out of scope.

The objective is the reference's sum of up to three terms, evaluated in its order (regressor, discriminator hinge, content
reconstruction) on the decoded image clamped to [-1, 1]; `tests/test_oracle_cpu.py::test_imaginaire_objective_matches_reference`
pins the oracle restatement of it to the reference bit for bit and `tests/test_imaginaire_gpu.py` pins this module to the oracle.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

STYLE_DIMS = 8            # MUNIT style code: [1, 8, 1, 1]


def initialize_imaginaire(image, obj_params):
    """Start vector = the style code of the input image; its content code and the image itself become objective arguments."""
    encoder = obj_params["gen"].autoencoder_a
    with torch.no_grad():
        content, style = encoder.encode(image)
    obj_params.update(content=content, orig_image=image)
    return style, obj_params


def _style_code(x_opt):
    """The optimiser hands the style code over flat; the decoder wants [1, 8, 1, 1] float32."""
    return x_opt.view(1, STYLE_DIMS, 1, 1).to(torch.float32) if x_opt.dim() == 1 else x_opt


def _hinge_term(dis, gan_loss, img):
    """MUNIT trains with a hinge loss: only a NEGATIVE generator-side GAN loss is penalised."""
    logits, _, _ = dis.discriminator_a(img)
    return torch.relu(-gan_loss(logits, True, dis_update=False))


class _GraphedGenerator:
    """The two generator passes of the objective -- decode + clamp, re-encode + L1 -- as CUDA graphs
    (torch.cuda.make_graphed_callables: one forward and one backward graph each), built once per (generator, shapes).

    The generator is ~70 small layers (reflect pad, convolution, instance norm, AdaIN, spectral-norm power iteration ...):
    run eagerly at batch 16 / 256x256 it is bound by the host's launch rate, not by the GPU (measured: 24.4 ms per objective
    evaluation + gradient against 17.5 ms for the native regressor).  Graph replay issues the same kernels on the same
    values -- results are those of the eager modules; the in-place spectral-norm state (weight_u / weight_v) advances once
    per replay exactly as once per eager forward, and the warm-up / capture passes that building the graphs needs are
    undone by restoring that state afterwards."""

    def __init__(self, autoencoder, content, style):
        self.key = (tuple(content.shape), tuple(style.shape), content.device)
        state = {k: v.detach().clone() for k, v in autoencoder.state_dict().items()}

        def decode_clamped(content_, style_):
            return torch.clamp(autoencoder.decode(content_, style_), min=-1, max=1)

        def recon_l1(img_, content_):
            return F.l1_loss(autoencoder.encode(img_)[0], content_)

        c = content.detach().clone()
        x = style.detach().clone().float().requires_grad_(True)
        with torch.no_grad():
            img = decode_clamped(c, x).detach().clone()
        self.decode_clamped, self.recon_l1 = torch.cuda.make_graphed_callables(
            (decode_clamped, recon_l1), ((c, x), (img.requires_grad_(True), c.clone())))
        autoencoder.load_state_dict(state)                 # warm-up + capture advanced the power iteration: rewind


def _graphed(gen, content, style):
    """Graphed passes for this generator and these shapes; None for CPU tensors or RGIE_LATENT_GRAPHS=0 (eager modules)."""
    import os
    if not content.is_cuda or os.environ.get("RGIE_LATENT_GRAPHS", "1") == "0":
        return None
    key = (tuple(content.shape), tuple(style.shape), content.device)
    cached = getattr(gen, "_rgie_graphed", None)
    if cached is None or cached.key != key:
        gen._rgie_graphed = None                           # drop the old graphs (and their private memory pool) first
        gen._rgie_graphed = _GraphedGenerator(gen.autoencoder_a, content, style)
    return gen._rgie_graphed


def objective_function_imaginaire(x_opt, gen, orig_image, content, clf, weight_clf, weight_dis, weight_recon, dis=None,
                                  target=None, gan_loss=None):
    autoencoder = gen.autoencoder_a
    fixed_content = content.detach()
    style = _style_code(x_opt)
    graphs = _graphed(gen, fixed_content, style) if (x_opt.requires_grad and torch.is_grad_enabled()) else None
    # the decoder overshoots the image range, so the decoded image is clamped (as imaginaire itself does)
    if graphs is not None:
        img = graphs.decode_clamped(fixed_content, style)
    else:
        img = torch.clamp(autoencoder.decode(fixed_content, style), min=-1, max=1)
    loss = weight_clf * clf(img, target=target)
    if dis is not None and weight_dis > 0:
        loss = loss + weight_dis * _hinge_term(dis, gan_loss, img)
    if weight_recon > 0:
        if graphs is not None:
            loss = loss + weight_recon * graphs.recon_l1(img, fixed_content)
        else:
            re_encoded, _ = autoencoder.encode(img)                  # L1 distance of the content codes, as in imaginaire
            loss = loss + weight_recon * F.l1_loss(re_encoded, fixed_content)
    return loss

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


def objective_function_imaginaire(x_opt, gen, orig_image, content, clf, weight_clf, weight_dis, weight_recon, dis=None,
                                  target=None, gan_loss=None):
    autoencoder = gen.autoencoder_a
    fixed_content = content.detach()
    # the decoder overshoots the image range, so the decoded image is clamped (as imaginaire itself does)
    img = torch.clamp(autoencoder.decode(fixed_content, _style_code(x_opt)), min=-1, max=1)
    loss = weight_clf * clf(img, target=target)
    if dis is not None and weight_dis > 0:
        loss = loss + weight_dis * _hinge_term(dis, gan_loss, img)
    if weight_recon > 0:
        re_encoded, _ = autoencoder.encode(img)                      # L1 distance of the content codes, as in imaginaire
        loss = loss + weight_recon * F.l1_loss(re_encoded, fixed_content)
    return loss

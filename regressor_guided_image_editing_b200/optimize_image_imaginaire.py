"""Drop-in for the hot-path functions of src/optimize_image_imaginaire.py: `initialize_imaginaire` (:112-117) and
`objective_function_imaginaire` (:120-145) -- latent (style-code) optimisation through a MUNIT generator.

The generator (`gen.autoencoder_a.encode/decode`, external/imaginaire) stays a PyTorch module (its
kernels are a later row); the regressor term `clf(img, target)` is the native ValenceArousalLoss (autograd Function over
librgie.so) and the Adam update / best-x tracking of `baselines.optimize_image.optimization` are the fused native
kernels, so the loop's regressor forward + input-gradient backward -- ~90 % of its FLOPs -- run on the tcgen05 path.
Script glue (main, checkpoints, discriminator loading, file I/O) is out of scope.
"""
from __future__ import annotations

import torch


def initialize_imaginaire(image, obj_params):                                             # :112-117
    with torch.no_grad():
        content, style = obj_params["gen"].autoencoder_a.encode(image)
    obj_params["content"] = content
    obj_params["orig_image"] = image
    return style, obj_params


def objective_function_imaginaire(x_opt, gen, orig_image, content, clf, weight_clf, weight_dis, weight_recon, dis=None,
                                  target=None, gan_loss=None):                            # :120-145
    if len(x_opt.shape) == 1:
        x_opt = x_opt.view(1, 8, 1, 1).to(torch.float32)

    content = content.detach()
    img = gen.autoencoder_a.decode(content, x_opt)
    # the decoder overshoots the image range; the reference clamps (as the imaginaire repo does)
    img = torch.clamp(img, min=-1, max=1)

    loss = weight_clf * clf(img, target=target)

    if dis is not None and weight_dis > 0:
        out_ba, _, _ = dis.discriminator_a(img)
        dis_loss = gan_loss(out_ba, True, dis_update=False)
        # hinge loss: penalise negative discriminator outputs, accept positive ones
        loss = loss + weight_dis * torch.relu(-dis_loss)

    if weight_recon > 0:
        # L1 reconstruction on the content code, as in imaginaire
        content_new, _ = gen.autoencoder_a.encode(img)
        loss = loss + weight_recon * torch.nn.functional.l1_loss(content_new, content)

    return loss

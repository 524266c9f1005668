"""Drop-in for the hot-path functions of src/optimize_image_param.py: `init_params` (:121-209),
`initialize_parametric` (:212-234), `objective_function_parametric` (:237-259), `get_params_from_vector` (:262-292),
and the caller right after the loop, `output_transform` (:295-312: evaluation + full-resolution re-render of the edit,
SURVEY.md 8f rank 4).  Script glue (main, dataset loop, JPEG save, CLIP / discriminator terms) is out of scope (section 8).
"""
from __future__ import annotations

import torch

from . import _lib
from .baselines.image_transformations.image_transformations import apply_params

STATS = {}                 # :21   per-adaptation evaluation statistics filled by output_transform
OUTPUT_TRANSFORM = None    # :22   set by configure_output(); the reference builds it in main() (:78-82)
DATA_DIR = None            # the reference takes it from paths.py


def configure_output(output_size=1024, data_dir=None):
    """The part of the reference's main() (:78-82) that output_transform depends on: Resize + CenterCrop + ToTensor."""
    global OUTPUT_TRANSFORM, DATA_DIR
    from torchvision import transforms
    OUTPUT_TRANSFORM = transforms.Compose([transforms.Resize(output_size), transforms.CenterCrop(output_size),
                                           transforms.ToTensor()])
    if data_dir is not None:
        DATA_DIR = data_dir


DEFAULT_TRANS = ['exposure', 'saturation', 'tone', 'color', 'contrast', 'sharp', 'blur', 'scale']


def init_params(trans_to_apply):                                                          # :121-209
    x0, params = [], {}
    for name in trans_to_apply:
        if name in ("gamma", "contrast", "saturation"):
            params[name] = 1.0
            x0.append(params[name])
        elif name in ("sharp", "wb", "bright", "exposure", "bw", "hue"):
            params[name] = 0.0
            x0.append(params[name])
        elif name == "blur":
            params["blur"] = 1e-4
            x0.append(params["blur"])
        elif name == "tone":
            params["tone"] = torch.ones(1, 8, 1)
            x0.extend(params["tone"].view(-1).tolist())
        elif name == "color":
            params["color"] = torch.ones(3, 8, 1)
            x0.extend(params["color"].view(-1).tolist())
        elif name == "affine":
            params["affine"] = torch.eye(2, 3)
            x0.extend(params["affine"].view(-1).tolist())
        elif name == "scale":
            params["scale"] = torch.ones(1, 4)
            params["scale"][0, 2:4] = 0.0
            x0.extend(params["scale"].view(-1).tolist())
    return params, torch.tensor(x0)


def initialize_parametric(image, obj_params):                                             # :212-234
    params_trans, x0 = init_params(DEFAULT_TRANS)
    x0 = x0.to(image.device)
    obj_params["image"] = image
    obj_params["params"] = params_trans
    return x0, obj_params


def get_params_from_vector(x, batch_size, params, input_size=480):                        # :262-292
    ix_start = 0
    for name in params.keys():
        len_param = 1 if isinstance(params[name], float) else len(params[name].view(-1))
        if len_param == 1:
            params[name] = x[ix_start]
        else:
            param_tensor = x[ix_start:ix_start + len_param]
            if "tone" == name:
                params["tone"] = param_tensor.view(batch_size, 1, 8, 1)
            elif "color" == name:
                params["color"] = param_tensor.view(batch_size, 3, 8, 1)
            elif "affine" == name:
                params["affine"] = param_tensor.view(2, 3)[None].repeat(batch_size, 1, 1)
            elif "scale" == name:
                if param_tensor.size(0) == 4:
                    clamped_scale = param_tensor[0:2].clamp(min=1.0, max=torch.inf)
                    clamped_center = param_tensor[2:].clamp(min=0.0, max=input_size)
                    param_tensor = torch.cat((clamped_scale, clamped_center))
                else:
                    param_tensor = param_tensor.clamp(min=1.0, max=5.0)
                params["scale"] = param_tensor.repeat(batch_size, 1)
        ix_start += len_param
    params["contrast"] = 0.0 if params["contrast"] < 0 else params["contrast"]
    return params


def objective_function_parametric(x_opt, image, params, clf, weight_clf, weight_dis, weight_recon, dis=None,
                                  target=None):                                          # :237-259
    params_x = get_params_from_vector(x_opt, 1, params, image.size(2))
    outputs = apply_params(image, params_x)
    loss = weight_clf * clf(outputs[-1], target=target)
    if dis is not None and weight_dis > 0:
        loss = loss - weight_dis * dis(image)
    if weight_recon > 0:
        raise _lib.RgieError("the CLIP reconstruction term (optimize_image.py:152-183) is out of scope here "
                             "(SURVEY.md 8f rank 3): run with weight_recon=0")
    return loss


def output_transform(image, x_opt, obj_params, eval_params, adaptation, image_path):      # :295-312
    """Evaluate the optimised parameters on the working-size image, then re-render the edit on the full-resolution file
    (every filter is resolution-independent; `scale` centres are rescaled by get_params_from_vector's input_size)."""
    from .baselines.run_img_trans import compare_emotions
    from .baselines.utils import check_init_stats_adapt
    params_x = get_params_from_vector(x_opt, 1, obj_params["params"], image.size(2))
    outputs = apply_params(image, params_x)
    if "scale" in params_x:
        print(f"optimized scale: {params_x['scale'].tolist()}")

    check_init_stats_adapt(STATS, adaptation)
    compare_emotions(obj_params["clf"], image, outputs[-1], eval_params["emotion_type_labels"], STATS[adaptation])

    if OUTPUT_TRANSFORM is None:
        raise _lib.RgieError("output_transform: call configure_output(output_size, data_dir) first (reference main() :78-82)")
    import PIL.Image
    path = f"{DATA_DIR}/images/{image_path[0]}"
    print(path)
    input_image = PIL.Image.open(path)
    if input_image.mode != "RGB":
        input_image = input_image.convert('RGB')
    input_image_tensor = OUTPUT_TRANSFORM(input_image).unsqueeze(0).to(image.device)
    output_image_tensor = apply_params(input_image_tensor, params_x)[-1]
    return input_image_tensor, output_image_tensor

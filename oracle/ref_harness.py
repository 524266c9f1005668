"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference Python from /root/reference/src (read-only mount, exists only in the build
container, not on the GPU box) so that (a) the standalone restatement in oracle/oracle.py can be validated against the
reference's own code and (b) golden vectors can be generated (oracle/gen_golden.py).

Third-party packages the reference imports at module import time but that are not installed here are satisfied by
inert stub modules (`clip`, `matplotlib`, `plot_utils`, `diffusers`, ...) and by the pure-torch kornia restatement in
oracle/kornia_shim.py (kornia 0.8.2 is un-vendored; see that file's header).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = os.environ.get("RGIE_REFERENCE_SRC", "/root/reference/src")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "baselines"))


class _Anything:
    """Attribute sink: any attribute access / call returns another sink (enough for `from x import Y`)."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _stub(name: str, attrs=()):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = []  # behave like a package

    def _getattr(attr, _name=name):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return type(attr, (_Anything,), {})

    m.__getattr__ = _getattr  # PEP 562
    for a in attrs:
        setattr(m, a, type(a, (_Anything,), {}))
    sys.modules[name] = m
    return m


_installed = False


def install():
    """Make `import baselines...`, `import optimize_image_param` resolve to the reference tree."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_SRC}")
    from . import kornia_shim
    kornia_shim.install()
    for name in ("clip", "plot_utils", "wandb", "pycocotools", "pycocotools.coco", "albumentations", "skimage",
                 "accelerate"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        mpl.pyplot = plt
        _stub("matplotlib.patches")
        _stub("matplotlib.gridspec")
    try:
        import diffusers  # noqa: F401
    except Exception:
        _stub("diffusers", ("StableDiffusionXLPipeline", "StableDiffusionPipeline", "DDIMScheduler",
                            "DDIMInverseScheduler", "AutoencoderKL"))
        _stub("diffusers.utils")
        _stub("diffusers.utils.torch_utils")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # The reference's local `datasets` package must win over HuggingFace `datasets` if that was imported earlier.
    for k in [k for k in sys.modules if k == "datasets" or k.startswith("datasets.")]:
        mod = sys.modules[k]
        f = getattr(mod, "__file__", "") or ""
        if REFERENCE_SRC not in f:
            del sys.modules[k]
    _installed = True


def ref():
    """Return a namespace with the reference's hot-path callables."""
    install()
    ns = types.SimpleNamespace()
    ns.optimize_image = importlib.import_module("baselines.optimize_image")
    ns.optimize_image_param = importlib.import_module("optimize_image_param")
    ns.image_transformations = importlib.import_module("baselines.image_transformations.image_transformations")
    ns.ittf = importlib.import_module("baselines.image_transformations.img_trans_torch_diff")
    ns.ValenceArousalLoss = importlib.import_module("baselines.losses.ValenceArousalLoss").ValenceArousalLoss
    ns.EmotionPredictionModel = importlib.import_module("baselines.models.EmotionPredictionModel")
    ns.guidance_scores = importlib.import_module("guidance_classifier.guidance_scores")
    return ns

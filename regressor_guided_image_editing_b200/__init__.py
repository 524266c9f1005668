"""B200-native (sm_100a) drop-in for the hot path of christophgebhardt/regressor-guided-image-editing.

The sub-packages mirror the reference's `src/` layout (`baselines.optimize_image`, `baselines.image_transformations`,
`baselines.losses.ValenceArousalLoss`, `baselines.models.EmotionPredictionModel`, `optimize_image_param`,
`guidance_classifier`) so the reference scripts can import them unchanged by putting this package directory on
`sys.path`.  All arithmetic runs in librgie.so (hand-written CUDA, C ABI in include/rgie.h);
there is no CPU / PyTorch fallback.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]

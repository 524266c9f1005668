"""Crop-offset replay for src/baselines/models/utilities/ReplicateAndCrop.py:30-45.

The reference draws each crop's offsets with torchvision RandomCrop.get_params, i.e. two `torch.randint(...).item()`
calls per crop on torch's GLOBAL CPU generator (none at all when the image already has the crop size).  The native
regressor takes the offsets as an explicit int32 table; this helper consumes exactly the same draws in the same order so
that a reference script seeded with torch.manual_seed sees identical crops.
"""
from __future__ import annotations

import torch


def draw_crop_offsets(batch: int, h: int, w: int, crop: int = 448, reps: int = 10, generator=None) -> torch.Tensor:
    """int32 [batch, reps, 2] of (top, left), CPU tensor."""
    out = torch.zeros(batch, reps, 2, dtype=torch.int32)
    if h < crop or w < crop:
        raise ValueError(f"Required crop size {(crop, crop)} is larger than input image size {(h, w)}")
    if h == crop and w == crop:
        return out
    for b in range(batch):
        for r in range(reps):
            out[b, r, 0] = torch.randint(0, h - crop + 1, size=(1,), generator=generator).item()
            out[b, r, 1] = torch.randint(0, w - crop + 1, size=(1,), generator=generator).item()
    return out

"""Drop-in for the evaluation helpers of src/baselines/utils.py that sit right after the optimisation loop
(SURVEY.md 8f rank 4): the per-adaptation statistics dictionary and the row interleave used by `compare_emotions`.
Plotting / time-stamp / dataset helpers of the reference file are script glue and stay out of scope."""
from __future__ import annotations

import numpy as np
import torch


def interweave_batch_tensors(batch1: torch.Tensor, batch2: torch.Tensor) -> torch.Tensor:      # utils.py:231-238
    """Rows of the two batches alternated: (N, ...) x 2 -> (2N, prod(...))."""
    n = batch1.size(0)
    out = torch.empty(2 * n, batch1[0].numel(), dtype=batch1.dtype, device=batch1.device)
    out[0::2] = batch1.reshape(n, -1)
    out[1::2] = batch2.reshape(n, -1)
    return out


def print_stats(stats: dict) -> None:                                                        # utils.py:274-281
    for label, data in stats.items():
        parts = [f"{stat}: mean {np.mean(v):.4f}, std {np.std(v):.4f}; " for stat, v in data.items() if len(v) > 0]
        print(f"{label}: " + "".join(parts))


def check_init_stats_adapt(stats_dict: dict, adaptation) -> None:                            # utils.py:284-288
    stats_dict.setdefault(adaptation, {k: [] for k in ("valence", "arousal", "delta_valence", "delta_arousal", "rec_error")})

"""Batch sharding of the per-image optimisation across the GPUs of one box.

Images are independent optimisation problems (the reference runs them one by one, baselines/optimize_image.py:18,28),
so the job shards with NO collective inside the loop: rank r of W owns a contiguous block of image indices, runs its
micro-batches on its own GPU, and one final gather moves the edited images / predictions / target-error statistics to
rank 0 over NCCL (NVLink 5 through NVSwitch on a B200 box).  Per-image seeds make every result independent of W.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch


def partition(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block [begin, end) of rank `rank`: blocks of ceil(n/W), the last ranks may be short or empty."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = -(-n_items // world)
    begin = min(rank * per, n_items)
    return begin, min(begin + per, n_items)


def micro_batches(begin: int, end: int, batch: int) -> List[Tuple[int, int]]:
    """Split a rank's block into engine batches; the last one may be short (the caller pads it by repeating an image)."""
    return [(s, min(s + batch, end)) for s in range(begin, end, batch)]


def gather_to_rank0(local: Dict[str, torch.Tensor], counts: Optional[List[int]] = None, group=None
                    ) -> Optional[Dict[str, torch.Tensor]]:
    """Final gather of per-image results (first dimension = images of this rank) to rank 0.
    Works on any torch.distributed backend: nccl on the GPU box, gloo in the CPU tests.  Ranks may own different counts."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        return {k: v.clone() for k, v in local.items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    any_t = next(iter(local.values()))
    n_local = torch.tensor([any_t.shape[0]], dtype=torch.int64, device=any_t.device)
    all_n = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(all_n, n_local, group=group)
    ns = [int(t.item()) for t in all_n]
    n_max = max(ns)
    out = {}
    for k, v in local.items():
        pad = torch.zeros((n_max,) + tuple(v.shape[1:]), dtype=v.dtype, device=v.device)
        pad[:v.shape[0]] = v
        # a true gather: only rank 0 holds the receive buffers (N x 3 x H x W edited images are the bulk of the job's output)
        bufs = [torch.zeros_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, gather_list=bufs, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        if rank == 0:
            out[k] = torch.cat([b[:n] for b, n in zip(bufs, ns)], 0)
    return out if rank == 0 else None


def target_error_stats(preds_va: torch.Tensor, target: torch.Tensor, pred0_va: torch.Tensor) -> Dict[str, float]:
    """The statistics the reference prints per adaptation (baselines/run_img_trans.py:361-386 `compare_emotions`):
    mean valence/arousal before and after, and the mean absolute distance to the target."""
    return {"valence_before": float(pred0_va[:, 0].mean()), "arousal_before": float(pred0_va[:, 1].mean()),
            "valence_after": float(preds_va[:, 0].mean()), "arousal_after": float(preds_va[:, 1].mean()),
            "target_abs_error": float((preds_va - target).abs().mean())}

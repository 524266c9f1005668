"""CPU: sharding host logic, incl. the final gather on a world_size-2 gloo group (the N>1 path of bench.py / SURVEY 8e)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from regressor_guided_image_editing_b200 import shard


def test_partition_covers_everything_without_overlap():
    for n in (0, 1, 7, 64, 4096, 4099):
        for w in (1, 2, 4, 8):
            blocks = [shard.partition(n, w, r) for r in range(w)]
            flat = [i for b, e in blocks for i in range(b, e)]
            assert flat == list(range(n))
            assert max(e - b for b, e in blocks) - min(e - b for b, e in blocks) <= -(-n // w)


def test_micro_batches():
    assert shard.micro_batches(10, 75, 32) == [(10, 42), (42, 74), (74, 75)]
    assert shard.micro_batches(5, 5, 8) == []


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard.partition(n_items, world, rank)
    # per-image "results" that depend only on the global image index, like the per-image seeds of the real job
    idx = torch.arange(b, e)
    local = {"best_loss": idx.float() * 0.5, "preds": torch.stack([idx.float(), -idx.float()], 1),
             "edited": idx.float()[:, None, None, None].expand(e - b, 3, 4, 4).contiguous()}
    out = shard.gather_to_rank0(local)
    if rank == 0:
        q.put({k: v.clone() for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_final_gather_world2_gloo():
    n_items, world, port = 7, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    idx = torch.arange(n_items).float()
    assert torch.equal(out["best_loss"], idx * 0.5)
    assert torch.equal(out["preds"], torch.stack([idx, -idx], 1))
    assert out["edited"].shape == (n_items, 3, 4, 4) and torch.equal(out["edited"][:, 0, 0, 0], idx)


def test_target_error_stats():
    s = shard.target_error_stats(torch.tensor([[0.8, 0.5]]), torch.tensor([[0.9, 0.6]]), torch.tensor([[0.7, 0.5]]))
    assert abs(s["target_abs_error"] - 0.1) < 1e-6 and abs(s["valence_before"] - 0.7) < 1e-6

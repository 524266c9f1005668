#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): edited images/sec for the parametric-filter edit,
batch 64 synthetic 512x512 images, 100 optimisation steps per image, random-init regressor (configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one optimisation step of the whole batch (8 filters fwd -> AA resize -> 10 random crops -> resnet50 fwd ->
VA loss -> resnet50 input-gradient bwd -> resize^T -> 8 filters bwd -> Adam + best-x) -- one pass of the hot path over one
batch.  An image is "edited" after 100 such steps, so value = (N * B * K / 100) / elapsed [images/s]; every step costs the
same, the default K=100 is exactly one full edit of the batch.
  value : device-resident inputs, CUDA-graph replay, CUDA events, max over ranks.
  e2e   : same job through the public engine API with HOST buffers: pinned images/offsets H2D, per-step D2H of the loss
          vector (the reference prints float(loss) every step), edited images + predictions D2H; all inside the timed region.
  roofline : the dominant kernel family (row-shifted tcgen05 GEMM, ~125 launches/step = regressor fwd + dgrad), timed
          live with cudaEvent pairs around every launch in a few extra eager steps; algorithmic FLOPs / time vs the
          MEASURED bf16 peak (MEASURED_PEAKS.json, sustained figure since the kernels run inside a long step).
  cpu_baseline : the oracle port of the reference loop (oracle/oracle.py, weight-gradient work included as the reference
          does it) timed on this box's host cores on a bounded sample (rank 0, N=1 only).
  regressor_fwd_bwd_ms : cudaEvents around resize -> crop pack -> resnet50 fwd -> VA head -> resnet50 dgrad -> crop gather ->
          resize^T of eager steps; roofline.achieved / frac are computed on THIS span (all of the regressor,
          not the GEMM subset); the GEMM-only figures sit under roofline.gemm_only.
  --images N : BASELINE.json configs[4] -- a strong-scaling sweep of N images sharded over the ranks (shard.partition), each
          rank running ceil(N/W)/B engine batches of 100 steps through ONE engine + CUDA graph, then ONE gather of uint8 edited
          images + predictions + per-step losses to rank 0 (NCCL) and their D2H; everything inside the timed region.
  --latent : BASELINE.json configs[2] -- latent (style-code) optimisation through the random-init MUNIT generator + the
          native regressor, batch 16 at 256x256, through the reference's own call surface (initialize_imaginaire,
          objective_function_imaginaire, optimization); the generator is PyTorch (library convolutions),
          the regressor forward + input gradient and the Adam / best-x update are librgie.so.  Single GPU, extra line.
  inputs : seeded synthetic images / regressor weights from bench_inputs.py; oracle/ is imported by the CPU
          leg (cpu_reference_leg) only.
--impl reference : times that same CPU implementation alone (the reference is pure Python/PyTorch; /root/reference does
          not exist on the GPU box, the oracle port is its restatement validated bit-exactly against it).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEPS_PER_IMAGE = 100
FLOP_PER_IMAGE_STEP = 653.9e9          # 10 crops x 65.39 GFLOP (fwd + input-gradient), SURVEY.md 8(d)


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def _lib_sha():
    import hashlib
    try:
        with open(os.path.join(ROOT, "regressor_guided_image_editing_b200", "librgie.so"), "rb") as f:
            return hashlib.sha256(f.read()).hexdigest()[:16]
    except Exception:
        return None


def _src_sha():
    """sha256 over the sources the GEMM family of the regressor is built from (the kernels, their PTX helpers, the launch
    plan, the Makefile with the compiler flags), in name order.  nvcc's output is not byte-reproducible (two clean builds of
    one tree give different librgie.so hashes), so an ncu capture under profiles/ is matched to the running build by this
    hash; tools/summarize_launches.py stamps the same value.  `roofline.traffic` is a statement about these kernels only, so
    the filter / resize / update / MiDU sources are not part of it."""
    import hashlib
    csrc = os.path.join(ROOT, "regressor_guided_image_editing_b200", "csrc")
    names = ["Makefile", "common.cuh", "gemm_sm100.cu", "gemm_sm100.cuh", "regressor.cu", "sm100_ptx.cuh"]
    h = hashlib.sha256()
    try:
        for fn in names:
            h.update(fn.encode() + b"\0")
            with open(os.path.join(csrc, fn), "rb") as f:
                h.update(f.read())
        return h.hexdigest()[:16]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop, self.index = [], set(), threading.Event(), index
        self.max_mhz = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                f = [x.strip() for x in out.split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_leg(steps: int, warmup: int, h: int, w: int, threads: int, budget_s: float = 1e9):
    """The reference's own CPU implementation of the path (oracle port; weight gradients computed like the reference's
    requires_grad=True, optimize_image_param.py:62-63).  One 'step' = one optimisation step of ONE image."""
    import torch
    from oracle import oracle as O
    torch.set_num_threads(threads)
    sd = O.make_regressor_state_dict()
    wkeys = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k]
    for k in wkeys:
        sd[k].requires_grad_(True)
    image = O.synthetic_image(0, h, w)[None]
    torch.manual_seed(2000)
    offs = O.draw_crop_offsets(1 + warmup + steps, 1, 480, 480)
    x = O.init_x0().clone().requires_grad_(True)
    m, v_ = torch.zeros(41), torch.zeros(41)
    with torch.no_grad():
        target = O.get_condition_from_alpha(O.regressor_predict(image, sd, offs[0])[:, [0, 1]], 0.1)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        loss, _, _ = O.objective_parametric(x, image, sd, offs[1 + s], target, 0.15)
        for k in wkeys:
            sd[k].grad = None
        x.grad = None
        loss.backward()
        with torch.no_grad():
            O.adam_step(x, x.grad, m, v_, s + 1, O.lr_schedule(s, STEPS_PER_IMAGE, 0.05))
        if s >= warmup:
            times.append(time.perf_counter() - t0)
            if sum(times) > budget_s:
                break
    sec_per_step = sum(times) / len(times)
    return 1.0 / (STEPS_PER_IMAGE * sec_per_step), sec_per_step, len(times)


def run_sweep(args, rank, world, dev, lib):
    """BASELINE.json configs[4]: --images N synthetic images, strong split over the ranks, final gather to rank 0."""
    import torch
    import torch.distributed as dist
    import bench_inputs as BI               # seeded synthetic inputs / weights
    from regressor_guided_image_editing_b200 import engine, shard

    N, B, H, S = args.images, args.batch, args.size, args.sweep_steps
    begin, end = shard.partition(N, world, rank)
    batches = shard.micro_batches(begin, end, B)
    sd = BI.make_regressor_state_dict()
    eng = engine.ParametricEditEngine(sd, batch=B, height=H, width=H, num_steps=S, precision=args.precision,
                                      micro_batch=args.micro_batch, device=dev)
    n_local = end - begin
    # host inputs of this rank (pinned), per-image seeds: results do not depend on the world size
    images_h = torch.empty(max(n_local, 1), 3, H, H, dtype=torch.float32).pin_memory()
    offs_h = torch.empty(1 + S, max(n_local, 1), 10, 2, dtype=torch.int32).pin_memory()
    for k, i in enumerate(range(begin, end)):
        images_h[k] = BI.synthetic_image(i, H, H)
        g = torch.Generator().manual_seed(2000 + i)
        offs_h[:, k] = torch.randint(0, eng.Hr - 448 + 1, (1 + S, 10, 2), generator=g, dtype=torch.int32)
    edited_d = torch.empty(max(n_local, 1), 3, H, H, dtype=torch.uint8, device=dev)
    preds_d = torch.empty(max(n_local, 1), 2, dtype=torch.float32, device=dev)
    pred0_d = torch.empty_like(preds_d); target_d = torch.empty_like(preds_d)
    losses_d = torch.empty(max(n_local, 1), S, dtype=torch.float32, device=dev)

    def one_batch(b0, b1):
        n = b1 - b0
        idx = list(range(b0 - begin, b1 - begin)) + [b0 - begin] * (B - n)          # a short last batch repeats an image
        img = images_h[idx] if n < B else images_h[b0 - begin:b1 - begin]
        off = offs_h[:, idx] if n < B else offs_h[:, b0 - begin:b1 - begin]
        eng.load_problem(img.to(dev, non_blocking=True), off.contiguous().to(dev, non_blocking=True))
        eng.advance(S)
        res = eng.results()
        sl = slice(b0 - begin, b1 - begin)
        edited_d[sl] = (res["edited"][:n] * 255.0).round_().clamp_(0, 255).to(torch.uint8)
        preds_d[sl] = res["preds"][-1, :n, :2]; pred0_d[sl] = res["pred0"][:n, :2]; target_d[sl] = res["target"][:n]
        losses_d[sl] = res["losses"][:, :n].t()

    # warm-up: one short run so that every kernel is instantiated and the step graph is captured (untimed)
    lps = 0
    if batches:
        eng.load_problem(images_h[:1].expand(B, 3, H, H).contiguous().to(dev), offs_h[:, :1].expand(1 + S, B, 10, 2).contiguous().to(dev))
        l0 = lib.rgie_launch_count()
        eng.advance(1)
        torch.cuda.synchronize(dev)
        lps = int(lib.rgie_launch_count() - l0)
        eng.advance(min(3, S) - 1)
    torch.cuda.synchronize(dev)
    # rank 0's pinned landing buffers and the gather's communicator exist before the clock starts (allocating 3 GB of
    # pinned memory and NCCL's lazy point-to-point set-up are not part of the job's steady state)
    host = None
    if rank == 0:
        host = {"edited": torch.empty(N, 3, H, H, dtype=torch.uint8).pin_memory(),
                "preds": torch.empty(N, 2).pin_memory(), "pred0": torch.empty(N, 2).pin_memory(),
                "target": torch.empty(N, 2).pin_memory(), "losses": torch.empty(N, S).pin_memory()}
    shard.gather_to_rank0({"warm": torch.zeros(1, 4, device=dev)})
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t0 = time.perf_counter()
    with ClockSampler(dev.index) as clk:
        e0.record()
        for b0, b1 in batches:
            one_batch(b0, b1)
        e1.record()
        local = {"edited": edited_d[:n_local], "preds": preds_d[:n_local], "pred0": pred0_d[:n_local],
                 "target": target_d[:n_local], "losses": losses_d[:n_local]}
        out = shard.gather_to_rank0(local)
        if rank == 0:
            for k, v in out.items():
                host[k].copy_(v, non_blocking=True)
        e2.record()
        torch.cuda.synchronize(dev)
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e2), e0.elapsed_time(e1)], device=dev)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)
        tmax = torch.stack(per_rank).max(0).values
    else:
        tmax = t
    if rank != 0:
        return None
    total_ms, compute_ms = float(tmax[0]), float(tmax[1])
    stats = shard.target_error_stats(host["preds"], host["target"], host["pred0"])
    gathered = sum(v.numel() * v.element_size() for v in host.values())
    steps_total = len(batches) * S
    line = {"metric": "edited images/sec (100 steps, 512^2)", "value": N * (S / STEPS_PER_IMAGE) / (total_ms / 1e3),
            "unit": "images/s", "n_gpus": world, "steps": steps_total, "warmup": min(3, S), "ms_per_step": compute_ms / max(steps_total, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else args.precision, "data": "synthetic",
            "config": {"workload": f"configs[4]: batch-sharded sweep of {N} synthetic {H}x{H} images, {S} steps/image, "
                                   f"8 default filters, random-init resnet50 VA regressor on 10 random 448 crops",
                       "images": N, "batch_per_engine": B, "micro_batch": args.micro_batch, "engine_batches_per_rank": len(batches),
                       "precision": args.precision, "parallelism": f"dp{world} (contiguous image blocks, no collective in the loop, "
                                                                     f"one final gather to rank 0)",
                       "cache": "working set (>10 GB of activations per step) exceeds the 126 MB L2; no flush needed"},
            "clocks": clk.summary(),
            "e2e": {"value": N * (S / STEPS_PER_IMAGE) / (total_ms / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": int((images_h.numel() * 4 + offs_h.numel() * 4) / max(steps_total, 1)),
                    "d2h_bytes_per_step": int(gathered / max(steps_total, 1))},
            "sweep": {"images": N, "total_s": total_ms / 1e3, "loop_s": compute_ms / 1e3, "gather_and_d2h_s": (total_ms - compute_ms) / 1e3,
                      "wall_s": wall, "gathered_bytes": gathered, "per_rank_total_ms": [float(x[0]) for x in per_rank],
                      "target_error_stats": stats,
                      "edited_uint8_checksum": int(host["edited"].to(torch.int64).sum().item()),
                      "mean_final_loss": float(host["losses"][:, -1].mean().item())},
            "gpu_launches": lps * steps_total, "launches_per_step": lps}
    return line


def run_latent(args, dev, lib):
    """BASELINE.json configs[2]: optimize_image_imaginaire.py's loop at batch 16, 256x256 (one GPU)."""
    import torch
    import bench_inputs as BI               # seeded synthetic inputs / regressor weights
    from regressor_guided_image_editing_b200 import optimize_image_imaginaire as oii
    from regressor_guided_image_editing_b200.baselines import optimize_image as oi
    from regressor_guided_image_editing_b200.baselines.losses.ValenceArousalLoss import ValenceArousalLoss
    from regressor_guided_image_editing_b200.external.imaginaire.generators.munit import Generator

    B, H, K, W = args.latent_batch, args.latent_size, max(args.steps, 1), max(args.warmup, 3)
    sd = BI.make_regressor_state_dict()
    torch.manual_seed(0)
    gen = Generator().to(dev)                                   # imagenet2imagenet.yaml, random init, training mode (:75-79)
    clf = ValenceArousalLoss(sd, dev, 1, is_minimized=True, is_input_range_0_1=False, requires_grad=True,
                             precision=args.precision)
    images_h = torch.stack([2.0 * BI.synthetic_image(300 + i, H, H) - 1.0 for i in range(B)]).pin_memory()
    edited_h = torch.empty(B, 3, H, H).pin_memory()
    style_h = torch.empty(B, 8, 1, 1).pin_memory()

    def job(steps, images_d=None):
        """One whole latent edit of the batch: encode -> target -> `steps` optimisation steps -> decode(best)."""
        img = images_h.to(dev, non_blocking=True) if images_d is None else images_d
        params = {"gen": gen, "clf": clf, "dis": None, "gan_loss": None, "weight_clf": 0.2, "weight_dis": 0.0,
                  "weight_recon": 1.0}
        x0, params = oii.initialize_imaginaire(img, params)
        params["target"] = oi.get_condition_from_alpha(0.1, clf, img)
        best = oi.optimization(x0, params, oii.objective_function_imaginaire, learning_rate=0.05, num_steps=steps)
        with torch.no_grad():
            edited = torch.clamp(gen.autoencoder_a.decode(params["content"].detach(), best), -1, 1)
        return best, edited, params

    torch.manual_seed(2300)
    l0 = lib.rgie_launch_count()
    _, _, params = job(W)                                       # warm-up: W optimisation steps, every kernel instantiated
    torch.cuda.synchronize(dev)
    launches_warm = lib.rgie_launch_count() - l0
    images_d = images_h.to(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    with ClockSampler(dev.index) as clk:
        torch.cuda.synchronize(dev)
        l0 = lib.rgie_launch_count()
        ev[0].record()
        job(K, images_d)                                        # value: inputs resident in HBM
        ev[1].record()
        torch.cuda.synchronize(dev)
        launches = lib.rgie_launch_count() - l0
        ev[2].record()
        best, edited, _ = job(K)                                # e2e: pinned host images in, edited images + style codes out
        edited_h.copy_(edited, non_blocking=True); style_h.copy_(best, non_blocking=True)
        ev[3].record()
        torch.cuda.synchronize(dev)
    ms, ms_e2e = ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])

    # split of one objective evaluation + gradient: generator (PyTorch) vs native regressor
    def timed(fn, n=5):
        fn(); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record(); torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / n

    content, x = params["content"].detach(), best.detach().clone().requires_grad_(True)
    ae = gen.autoencoder_a

    def gen_only():
        img = torch.clamp(ae.decode(content, x), -1, 1)
        loss = torch.nn.functional.l1_loss(ae.encode(img)[0], content) + img.mean()
        torch.autograd.grad(loss, x)

    img_fixed = torch.clamp(ae.decode(content, x), -1, 1).detach()

    def reg_only():
        im = img_fixed.clone().requires_grad_(True)
        torch.autograd.grad(clf(im, target=params["target"]), im)

    gen_ms, reg_ms = timed(gen_only), timed(reg_only)
    reg_flop = B * FLOP_PER_IMAGE_STEP
    peaks, which = _peaks()
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / (6.0 if args.precision == "fp32" else 1.0)
    line = {"metric": "edited images/sec (100 steps, 256^2, latent)", "value": B * (K / STEPS_PER_IMAGE) / (ms / 1e3), "unit": "images/s",
            "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else args.precision, "data": "synthetic",
            "config": {"workload": f"configs[2]: optimize_image_imaginaire.py latent optimisation through the random-init MUNIT "
                                   f"generator (imagenet2imagenet.yaml) + random-init resnet50 VA regressor on 10 random 448 crops, "
                                   f"batch {B} at {H}x{H}, weight_clf 0.2, weight_recon 1.0, lr 0.05, {STEPS_PER_IMAGE} steps/image",
                       "batch_per_gpu": B, "image": f"{H}x{H}", "precision": args.precision,
                       "generator": "PyTorch eager fp32 (cuDNN, TF32 convolutions as torch defaults), training mode: one spectral-norm "
                                    "power iteration per forward",
                       "regressor_and_update": "librgie.so (tcgen05 regressor forward + input gradient, fused Adam + best-x)",
                       "includes": "encode + target prediction + K optimisation steps + decode(best) per job",
                       "cache": "activations per step exceed the 126 MB L2; no flush needed"},
            "clocks": clk.summary(),
            "e2e": {"value": B * (K / STEPS_PER_IMAGE) / (ms_e2e / 1e3), "unit": "images/s",
                    "h2d_bytes_per_step": int(images_h.numel() * 4 / K), "d2h_bytes_per_step": int((edited_h.numel() + style_h.numel()) * 4 / K)},
            "gpu_launches": int(launches), "launches_per_step": int(launches // K),
            "breakdown_ms_per_step": {"objective_and_update": ms / K, "generator_decode_encode_fwd_bwd_torch": gen_ms,
                                      "regressor_fwd_bwd_native": reg_ms},
            "roofline": {"bound": "tensor", "achieved": reg_flop / (reg_ms / 1e3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                         "frac": reg_flop / (reg_ms / 1e3) / 1e12 / peak, "traffic": None,
                         "span": "native regressor fwd+bwd of one step (resize 256->480, 160 crops, resnet50 fwd + dgrad, resize^T), "
                                 "653.9 GFLOP per image", "peak_source": which},
            "cpu_baseline": None, "lib_sha256": _lib_sha(), "src_sha256": _src_sha()}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=int(os.environ.get("RGIE_MICRO_BATCH", "32")))
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--cpu-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-GEMM timing table (json) here")
    ap.add_argument("--images", type=int, default=0,
                    help="configs[4]: strong-scaling sweep of this many images over all ranks (0 = the weak-scaling headline)")
    ap.add_argument("--sweep-steps", type=int, default=STEPS_PER_IMAGE, help="optimisation steps per image in --images mode")
    ap.add_argument("--latent", action="store_true", help="configs[2]: MUNIT latent optimisation, batch 16 at 256x256 (one GPU)")
    ap.add_argument("--latent-batch", type=int, default=16)
    ap.add_argument("--latent-size", type=int, default=256)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W, B, H = max(args.steps, 1), max(args.warmup, 3), args.batch, args.size
    workload = (f"configs[1]: parametric-filter edit (8 default filters), batch {B} synthetic {H}x{H} images per GPU, "
                f"{STEPS_PER_IMAGE} steps/image, random-init resnet50 VA regressor on 10 random 448 crops")
    config = {"workload": workload, "batch_per_gpu": B, "micro_batch": args.micro_batch, "image": f"{H}x{H}",
              "steps_per_image": STEPS_PER_IMAGE, "precision": args.precision, "parallelism": f"dp{world} (batch-sharded, no collective in the loop)",
              "cache": "working set (>10 GB of activations per step) exceeds the 126 MB L2; no flush needed"}

    # ---------------------------------------------------------------------------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        budget = float(os.environ.get("RGIE_CPU_BUDGET_S", "120"))
        Wc = min(W, 1)
        val, sec, K = cpu_reference_leg(K, Wc, H, H, threads, budget)
        W = Wc
        # `config` stays byte-identical to the native arm's (same workload, same keys): how the CPU arm runs it goes
        # under its own key
        how = {"precision": "fp32 (torch CPU)", "micro_batch": 1,
               "parallelism": f"{threads} host threads, one image at a time (the reference's own loop)"}
        line = {"metric": "edited images/sec (100 steps, 512^2)", "value": val, "unit": "images/s", "n_gpus": args.gpus,
                "steps": K, "warmup": W, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config,
                "reference_run": how,
                "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port",
                                 "sample": f"1 image x {K} optimisation steps (of {STEPS_PER_IMAGE}; bounded to ~{budget:.0f} s "
                                           f"of CPU work) at {H}x{H}, per-step time extrapolated to 100 steps; all "
                                           f"{threads} host threads"},
                "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist
    import bench_inputs as BI               # seeded synthetic inputs / weights; oracle/ is touched by cpu_reference_leg only
    from regressor_guided_image_editing_b200 import _lib, engine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback of the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    if args.latent:
        if rank == 0:
            print(json.dumps(run_latent(args, dev, lib)))
        if world > 1:
            dist.destroy_process_group()
        return

    if args.images > 0:
        line = run_sweep(args, rank, world, dev, lib)
        if line is not None:
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    sd = BI.make_regressor_state_dict()
    n_steps_total = W + K
    eng = engine.ParametricEditEngine(sd, batch=B, height=H, width=H, num_steps=max(n_steps_total, 1),
                                      precision=args.precision, micro_batch=args.micro_batch, device=dev)
    g = torch.Generator().manual_seed(1234 + rank)
    images_h = torch.stack([BI.synthetic_image(rank * B + i, H, H) for i in range(B)]).pin_memory()
    offs_h = torch.randint(0, eng.Hr - 448 + 1, (1 + n_steps_total, B, 10, 2), generator=g, dtype=torch.int32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident run: W warm-up steps then K timed steps (graph replays)
    eng.load_problem(images_h.to(dev, non_blocking=True), offs_h.to(dev, non_blocking=True))
    l0 = lib.rgie_launch_count()
    eng.advance(1)                                          # eager step: every kernel instantiated, launches counted
    torch.cuda.synchronize(dev)
    launches_per_step = lib.rgie_launch_count() - l0
    eng.ensure_graph()
    eng.advance(W - 1)                                      # remaining warm-up steps are graph replays
    done = W
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    Kt = K
    with ClockSampler(local_rank) as clk:
        e0.record()
        eng.advance(Kt)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    clk_own = clk.summary()
    t = torch.tensor([ms, float(clk_own["sm_mhz"] or 0.0)], device=dev)
    per_rank = [t.clone() for _ in range(world)]
    if world > 1:
        dist.all_gather(per_rank, t)            # attribution of the max: every rank's own time and median SM clock
    ms_max = max(float(x[0]) for x in per_rank)
    value = world * B * (Kt / STEPS_PER_IMAGE) / (ms_max / 1e3)

    # ---- live roofline of the GEMM family: a few eager steps with cudaEvent pairs around every GEMM launch
    prof = None
    if rank == 0:
        import ctypes as C
        eng.counter.fill_(min(done, eng.steps - 1))
        lib.rgie_regressor_set_profiling(eng.reg._h, 1)
        nops = lib.rgie_regressor_num_ops(eng.reg._h)
        tot_ms, tot_fl, table = 0.0, 0.0, None
        reps_prof = int(os.environ.get("RGIE_PROF_REPS", "4"))
        per_rep = []
        for r in range(reps_prof):
            eng.counter.fill_(min(done, eng.steps - 1))
            eng._step()
            torch.cuda.synchronize(dev)
            ms_a = (C.c_float * nops)(); fl_a = (C.c_double * nops)(); by_a = (C.c_double * nops)()
            info = (C.c_int * (4 * nops))(); n = C.c_int(0)
            _lib.check(lib.rgie_regressor_get_profile(eng.reg._h, ms_a, fl_a, by_a, info, nops, C.byref(n)))
            # the profile holds the LAST micro-batch of the step; all micro-batches are identical in shape
            tot_ms += sum(ms_a) * (B // eng.mb)
            tot_fl += sum(fl_a) * (B // eng.mb)
            tot_by = sum(by_a) * (B // eng.mb)
            per_rep.append(list(ms_a))
        med = [sorted(x)[len(x) // 2] for x in zip(*per_rep)]
        mn = [min(x) for x in zip(*per_rep)]
        table = [{"dir": "fwd" if info[4 * i] == 0 else "bwd", "N": info[4 * i + 1], "K": info[4 * i + 2],
                  "m_tiles": info[4 * i + 3], "ms": med[i], "ms_min": mn[i], "tflops": fl_a[i] / max(med[i], 1e-9) / 1e9,
                  "algorithmic_GB": by_a[i] / 1e9, "algorithmic_GBs": by_a[i] / max(med[i], 1e-9) / 1e6}
                 for i in range(n.value)]
        lib.rgie_regressor_set_profiling(eng.reg._h, 0)
        # regressor fwd+bwd span: events around resize -> ... -> resize^T of eager steps, no per-op events
        span = []
        eng.span_events = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        for r in range(reps_prof + 1):
            eng.counter.fill_(min(done, eng.steps - 1))
            eng._step()
            torch.cuda.synchronize(dev)
            if r > 0:
                span.append(eng.span_events[0].elapsed_time(eng.span_events[1]))
        eng.span_events = None
        span.sort()
        regressor_span_ms = span[len(span) // 2]
        gemm_ms_per_step = tot_ms / reps_prof
        peaks, which = _peaks()
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_note = " (sustained bf16 cuBLAS)"
        if args.precision == "fp32":
            # fp32 parity mode = exact bf16x3 operand split: six bf16 tensor instructions per fp32 product block
            peak, peak_note = peak / 6.0, " (sustained bf16 cuBLAS / 6: the fp32 mode issues six bf16 MMAs per fp32 product block)"
        achieved = tot_fl / reps_prof / (gemm_ms_per_step / 1e3) / 1e12
        n_launch = nops * (B // eng.mb)
        # DRAM bytes per launch of the same kernel family from the committed ncu capture (profiles/*_step_B32.json:
        # dram__bytes_read.sum + dram__bytes_write.sum over the GEMM launches of one optimisation step, micro-batch 32)
        # ... only from a capture of THIS build: tools/summarize_launches.py stamps the sha256 of the librgie.so it profiled
        traffic, traffic_src = None, None
        try:
            import glob
            sha, ssha = _lib_sha(), _src_sha()
            for cand in sorted(glob.glob(os.path.join(ROOT, "profiles", "*_step_B32.json")), reverse=True):
                pj = json.load(open(cand))
                same = (sha is not None and pj.get("lib_sha256") == sha) or (ssha is not None and pj.get("src_sha256") == ssha)
                if eng.mb == 32 and same:
                    traffic, traffic_src = pj["gemm_dram_bytes_per_launch"], os.path.relpath(cand, ROOT)
                    break
            if traffic is None:
                traffic_src = "no ncu capture of this build (librgie.so sha256 %s, sources %s) under profiles/: not reported" % (sha, ssha)
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        span_achieved = FLOP_PER_IMAGE_STEP * B / (regressor_span_ms / 1e3) / 1e12
        prof = {"bound": "tensor", "achieved": span_achieved, "peak": peak, "unit": "TFLOP/s", "frac": span_achieved / peak,
                "span": "whole regressor fwd+bwd (resize -> pack -> resnet50 fwd -> head -> dgrad -> crop gather -> resize^T), "
                        "algorithmic 653.9 GFLOP per image per step over regressor_fwd_bwd_ms",
                "regressor_fwd_bwd_ms": regressor_span_ms,
                "frac_of_whole_step": FLOP_PER_IMAGE_STEP * B / (ms_max / Kt / 1e3) / 1e12 / peak,
                "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": which + peak_note,
                "kernel": "gemm_tc32_kernel" if args.precision == "fp32" else "gemm_sm100_kernel<BN,STAGES,EPI,NEW> family",
                "gemm_only": {"achieved": achieved, "frac": achieved / peak, "ms_per_step": gemm_ms_per_step},
                "launches_per_step": n_launch, "gemm_ms_per_step": gemm_ms_per_step,
                "share_of_step": gemm_ms_per_step / (ms_max / Kt),
                "algorithmic_flop_per_step": tot_fl / reps_prof,
                "algorithmic_flop_per_launch": tot_fl / reps_prof / n_launch,
                "algorithmic_bytes_per_launch": tot_by / n_launch,
                "avg_launch_ms": gemm_ms_per_step / n_launch,
                "hbm": {"achieved_GBs": tot_by / (gemm_ms_per_step / 1e3) / 1e9, "peak_GBs": hbm_peak,
                        "frac": tot_by / (gemm_ms_per_step / 1e3) / 1e9 / hbm_peak,
                        "note": "the family mixes tensor-bound (K >= 1024) and HBM-bound (1x1 expansions) launches"}}
        # per-class reading: the family mixes tensor-bound launches (K >= 1024: the 3x3 convs of layer2-4 and the deep 1x1s)
        # and HBM-bound ones (everything else); each class against its own measured peak
        cls = {"tensor_bound": [0, 0.0, 0.0, 0.0], "hbm_bound": [0, 0.0, 0.0, 0.0]}
        for i in range(n.value):
            c = cls["tensor_bound" if info[4 * i + 2] >= 1024 and info[4 * i + 1] >= 128 else "hbm_bound"]
            c[0] += 1; c[1] += med[i]; c[2] += fl_a[i]; c[3] += by_a[i]
        # composite roofline of the family as it is launched (one kernel per conv): every launch against the SLOWER of its
        # own tensor floor (flops / sustained peak) and HBM floor (algorithmic bytes / measured copy bandwidth)
        floor_ms = sum(max(fl_a[i] / (peak * 1e12), by_a[i] / (hbm_peak * 1e9)) * 1e3 for i in range(n.value))
        prof["composite"] = {"floor_ms_per_micro_batch": floor_ms, "measured_ms_per_micro_batch": sum(med[:n.value]),
                             "frac": floor_ms / max(sum(med[:n.value]), 1e-9),
                             "note": "sum over launches of max(tensor floor, HBM floor) / sum of measured launch times"}
        prof["classes"] = {
            k: {"launches_per_micro_batch": v[0], "ms_per_micro_batch": v[1],
                "TFLOPs": v[2] / max(v[1], 1e-9) / 1e9, "tensor_frac": v[2] / max(v[1], 1e-9) / 1e9 / peak,
                "GBs": v[3] / max(v[1], 1e-9) / 1e6, "hbm_frac": v[3] / max(v[1], 1e-9) / 1e6 / hbm_peak}
            for k, v in cls.items()}
        if args.profile_out:
            os.makedirs(os.path.dirname(os.path.abspath(args.profile_out)), exist_ok=True)
            json.dump({"per_gemm": table, "summary": prof, "ms_per_step": ms_max / Kt}, open(args.profile_out, "w"), indent=1)

    # ---- end-to-end run through the public API with host buffers
    barrier()
    loss_h = torch.empty(B, dtype=torch.float32).pin_memory()
    edited_h = torch.empty(B, 3, H, H, dtype=torch.float32).pin_memory()
    preds_h = torch.empty(B, eng.nc, dtype=torch.float32).pin_memory()
    Ke = min(K, eng.steps)
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record()
    eng.load_problem(images_h.to(dev, non_blocking=True), offs_h[:1 + eng.steps].to(dev, non_blocking=True))
    # every step's loss vector is read back (the reference prints float(loss) per step); the read of step s is awaited
    # while step s + 1 is already running (two pinned buffers), so the host round trip never idles the GPU
    loss_hh = [loss_h, torch.empty_like(loss_h).pin_memory()]
    evs = [torch.cuda.Event(), torch.cuda.Event()]
    loss_trace = []
    for s in range(Ke):
        eng.advance(1)
        loss_hh[s & 1].copy_(eng.loss, non_blocking=True)
        evs[s & 1].record()
        if s > 0:
            evs[(s - 1) & 1].synchronize()
            loss_trace.append(float(loss_hh[(s - 1) & 1][0]))
    evs[(Ke - 1) & 1].synchronize()
    loss_trace.append(float(loss_hh[(Ke - 1) & 1][0]))
    res = eng.results()
    edited_h.copy_(res["edited"], non_blocking=True)
    preds_h.copy_(eng.preds, non_blocking=True)
    ee1.record()
    torch.cuda.synchronize(dev)
    e2e_ms = torch.tensor([ee0.elapsed_time(ee1)], device=dev)
    if world > 1:
        # final gather of the target-error statistics (the only communication of the job)
        stats = torch.stack([res["best_loss"].mean(), (eng.preds[:, :2] - eng.target).abs().mean()])
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * (Ke / STEPS_PER_IMAGE) / (float(e2e_ms.item()) / 1e3)
    h2d = (images_h.numel() * 4 + offs_h[:1 + eng.steps].numel() * 4) / Ke
    d2h = (edited_h.numel() * 4 + preds_h.numel() * 4) / Ke + loss_h.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, sec, _ = cpu_reference_leg(args.cpu_steps, 1, H, H, threads)
        cpu_base = {"value": v, "unit": "images/s", "cores": threads, "kind": "port",
                    "sample": f"1 image x {args.cpu_steps} optimisation steps (+1 warm-up) at {H}x{H}; {sec:.2f} s/step "
                              f"extrapolated to {STEPS_PER_IMAGE} steps/image"}
    line = {"metric": "edited images/sec (100 steps, 512^2)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": Kt, "warmup": done, "ms_per_step": ms_max / Kt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32 (bf16x3 split on tcgen05)", "data": "synthetic",
            "config": config, "clocks": clk_own,
            "per_rank": {"ms_per_step": [float(x[0]) / Kt for x in per_rank], "sm_mhz": [float(x[1]) for x in per_rank]},
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches_per_step * Kt), "launches_per_step": int(launches_per_step),
            "roofline": prof, "cpu_baseline": cpu_base,
            "regressor_fwd_bwd_ms": None if prof is None else prof["regressor_fwd_bwd_ms"],
            "gemm_ms_per_step": None if prof is None else prof["gemm_ms_per_step"],
            "lib_sha256": _lib_sha(), "src_sha256": _src_sha(),
            "final_mean_best_loss": float(res["best_loss"].mean().item())}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

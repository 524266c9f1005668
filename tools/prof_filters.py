"""Per-filter timing (CUDA events, median of --reps) at the bench shape: achieved HBM GB/s against the algorithmic bytes of
SURVEY.md 8(d): forward 2N (read + write), backward 3N (read x, read g, write gin), N = 12 * H * W bytes per image.
Also times the AA resize (N + 0.879 N each way) and the fused Adam update."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from regressor_guided_image_editing_b200 import ops, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--out", default="")
a = ap.parse_args()
B, H = a.batch, a.size
dev = "cuda"
peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6536.7
x = (0.05 + 0.9 * torch.rand(B, 3, H, H, device=dev)).contiguous()
g = torch.randn_like(x)
N = x.numel() * 4
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > L2: evict between timed launches

def timeit(fn):
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

PARAMS = {"exposure": [0.3], "saturation": [1.2], "tone": [1.0 + 0.05 * i for i in range(8)], "color": [1.0 + 0.01 * i for i in range(24)],
          "contrast": [1.1], "sharp": [0.4], "blur": [1e-4], "scale": [1.05, 1.1, 3.0, 5.0],
          "gamma": [1.2], "bright": [0.1], "bw": [0.3], "hue": [0.4], "wb": [0.5]}
rows = []
for name, pv in PARAMS.items():
    kind = _lib.FILTER_KINDS[name]
    n = len(pv)
    p = torch.tensor(pv, device=dev).repeat(B, 1).contiguous()
    out = torch.empty_like(x); gin = torch.empty_like(x); gp = torch.empty(B, n, device=dev)
    ws = ops.filter_workspace(B, H, H, dev)
    tf = timeit(lambda: ops.filter_fwd(kind, x, p, n, out=out, ws=ws))
    tb = timeit(lambda: ops.filter_bwd(kind, x, g, p, n, gp, n, gin=gin, ws=ws))
    rows.append({"filter": name, "fwd_ms": tf, "fwd_GBs": 2 * N / tf / 1e6, "fwd_frac": 2 * N / tf / 1e6 / peak,
                 "bwd_ms": tb, "bwd_GBs": 3 * N / tb / 1e6, "bwd_frac": 3 * N / tb / 1e6 / peak})
if True:
    name = "blur(sigma=2)"
    kind = _lib.FILTER_KINDS["blur"]
    p = torch.full((B, 1), 2.0, device=dev)
    out = torch.empty_like(x); gin = torch.empty_like(x); gp = torch.empty(B, 1, device=dev)
    ws = ops.filter_workspace(B, H, H, dev)
    tf = timeit(lambda: ops.filter_fwd(kind, x, p, 1, out=out, ws=ws))
    tb = timeit(lambda: ops.filter_bwd(kind, x, g, p, 1, gp, 1, gin=gin, ws=ws))
    rows.append({"filter": name, "fwd_ms": tf, "fwd_GBs": 2 * N / tf / 1e6, "fwd_frac": 2 * N / tf / 1e6 / peak,
                 "bwd_ms": tb, "bwd_GBs": 3 * N / tb / 1e6, "bwd_frac": 3 * N / tb / 1e6 / peak})
if True:
    # the fused head of the default chain: exposure -> saturation -> tone -> colour in one pass (fwd 2N; bwd reads x and g: 2N)
    pv = PARAMS["exposure"] + PARAMS["saturation"] + PARAMS["tone"] + PARAMS["color"]
    p = torch.tensor(pv, device=dev).repeat(B, 1).contiguous()
    out = torch.empty_like(x); gp = torch.empty(B, 34, device=dev)
    ws = ops.filter_workspace(B, H, H, dev)
    tf = timeit(lambda: ops.filter_prefix_fwd(x, p, 34, out=out))
    tb = timeit(lambda: ops.filter_prefix_bwd(x, g, p, 34, gp, 34, ws=ws))
    rows.append({"filter": "prefix(exp,sat,tone,col)", "fwd_ms": tf, "fwd_GBs": 2 * N / tf / 1e6, "fwd_frac": 2 * N / tf / 1e6 / peak,
                 "bwd_ms": tb, "bwd_GBs": 2 * N / tb / 1e6, "bwd_frac": 2 * N / tb / 1e6 / peak})
rs = ops.Resize(H, H, 480, 480)
xo = torch.empty(B, 3, 480, 480, device=dev); go = torch.randn_like(xo); gi = torch.empty_like(x)
No = xo.numel() * 4
tf = timeit(lambda: rs.fwd(x, out=xo)); tb = timeit(lambda: rs.bwd(go, gin=gi))
rows.append({"filter": "resize 512->480 (aa)", "fwd_ms": tf, "fwd_GBs": (N + No) / tf / 1e6, "fwd_frac": (N + No) / tf / 1e6 / peak,
             "bwd_ms": tb, "bwd_GBs": (N + No) / tb / 1e6, "bwd_frac": (N + No) / tb / 1e6 / peak})
print(f"B={B} {H}x{H}: N = {N/1e6:.1f} MB per pass operand; HBM peak {peak:.0f} GB/s (measured copy)")
for r in rows:
    print(f"{r['filter']:22s} fwd {r['fwd_ms']:7.3f} ms {r['fwd_GBs']:7.0f} GB/s ({100*r['fwd_frac']:5.1f}%)   "
          f"bwd {r['bwd_ms']:7.3f} ms {r['bwd_GBs']:7.0f} GB/s ({100*r['bwd_frac']:5.1f}%)")
tot = sum(r["fwd_ms"] + r["bwd_ms"] for r in rows if r["filter"] in ("exposure", "saturation", "tone", "color", "contrast", "sharp", "blur", "scale", "resize 512->480 (aa)"))
print(f"default chain as 8 separate filters + resize, fwd+bwd: {tot:.3f} ms")
tot2 = sum(r["fwd_ms"] + r["bwd_ms"] for r in rows if r["filter"] in ("prefix(exp,sat,tone,col)", "contrast", "sharp", "blur", "scale", "resize 512->480 (aa)"))
print(f"default chain as the engine runs it (fused head + contrast, sharp, blur, scale + resize), fwd+bwd: {tot2:.3f} ms")
if a.out:
    json.dump(rows, open(a.out, "w"), indent=1)

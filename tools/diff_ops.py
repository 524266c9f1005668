"""Compare two tools/prof_ops.py --out tables: per-op ms and the ops that moved by more than --thr."""
import json, sys
a = json.load(open(sys.argv[1])); b = json.load(open(sys.argv[2])); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.03
ta = sum(x["ms_min"] for x in a); tb = sum(x["ms_min"] for x in b)
print(f"total {ta:.3f} -> {tb:.3f} ms")
for x, y in zip(a, b):
    r = y["ms_min"] / max(x["ms_min"], 1e-9) - 1
    if abs(r) > thr:
        print(f"op {x['i']:3d} {x['dir']} N={x['N']:5d} K={x['K']:5d} mt={x['m_tiles']:6d}  {x['ms_min']:.3f} -> {y['ms_min']:.3f} ms ({100*r:+.1f} %)  {y['tflops']:.0f} TF/s {y['alg_GBs']:.0f} GB/s")

"""Per-GEMM timing of one regressor forward+backward (cudaEvent pairs around every GEMM launch), min over --reps."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from regressor_guided_image_editing_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=32)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--only", type=str, default="")
ap.add_argument("--out", type=str, default="")
a = ap.parse_args()
lib = _lib.load()
sd = O.make_regressor_state_dict()
B = a.images
reg = ops.Regressor(sd, max_crops=B * 10, precision="bf16")
img = torch.rand(B, 3, 480, 480, device="cuda")
offs = torch.randint(0, 33, (B, 10, 2), dtype=torch.int32, device="cuda")
dl = torch.randn(B * 10, 4, device="cuda") * 1e-3
dimg = torch.empty_like(img)
lib.rgie_regressor_set_profiling(reg._h, 1)
nops = lib.rgie_regressor_num_ops(reg._h)
best = None
for r in range(a.reps + 1):
    reg.forward(img, offs); reg.backward(dl, dimg); torch.cuda.synchronize()
    ms = (C.c_float * nops)(); fl = (C.c_double * nops)(); by = (C.c_double * nops)(); info = (C.c_int * (4 * nops))(); n = C.c_int(0)
    _lib.check(lib.rgie_regressor_get_profile(reg._h, ms, fl, by, info, nops, C.byref(n)))
    if r == 0: continue
    best = list(ms) if best is None else [min(x, y) for x, y in zip(best, ms)]
only = [int(x) for x in a.only.split(",")] if a.only else range(nops)
tab = [{"i": i, "dir": "fwd" if info[4*i] == 0 else "bwd", "N": info[4*i+1], "K": info[4*i+2], "m_tiles": info[4*i+3], "ms_min": best[i], "tflops": fl[i]/max(best[i], 1e-9)/1e9, "alg_GBs": by[i]/max(best[i], 1e-9)/1e6} for i in range(nops)]
for t in tab:
    if t["i"] in only: print(t)
print("total_ms_min", sum(best), "TF/s", sum(fl)/sum(best)/1e9, "algorithmic GB", sum(by)/1e9, "GB/s", sum(by)/sum(best)/1e6)
if a.out: json.dump(tab, open(a.out, "w"), indent=1)

"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list:
takes the LAST complete optimisation step (the launches between two consecutive adam_sched_kernel launches) and prints
time / share / DRAM bytes per kernel family.  Usage: python tools/summarize_launches.py launches.csv [--json out.json]"""
import csv, json, re, sys
from collections import OrderedDict, defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
launch = OrderedDict()
for r in rows:
    e = launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    unit = r["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    e[r["Metric Name"]] = v * scale
ids = list(launch)
adam = [i for i in ids if "adam_sched_kernel" in launch[i]["name"]]
assert len(adam) >= 2, "need two adam_sched_kernel launches to delimit a step"
step = [i for i in ids if adam[-2] < i <= adam[-1]]

def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"rgie::(<unnamed>::)?", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n
TP, DT = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"
fam = defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])      # launches, us, DRAM bytes, us x tensor-pipe %, us x DRAM-throughput %
for i in step:
    e = launch[i]
    f = fam[short(e["name"])]
    t = e.get("gpu__time_duration.sum", 0.0)
    f[0] += 1
    f[1] += t
    f[2] += e.get("dram__bytes_read.sum", 0.0) + e.get("dram__bytes_write.sum", 0.0)
    f[3] += t * e.get(TP, 0.0)
    f[4] += t * e.get(DT, 0.0)
tot = sum(v[1] for v in fam.values())
is_gemm = lambda k: k.startswith("gemm_") or "hshare_kernel" in k      # the tcgen05 row-shifted GEMM family
gemm = sum(v[1] for k, v in fam.items() if is_gemm(k))
gemm_n = sum(v[0] for k, v in fam.items() if is_gemm(k))
gemm_b = sum(v[2] for k, v in fam.items() if is_gemm(k))
print(f"# one optimisation step: {len(step)} launches, {tot:.1f} us (cold-cache, serialised under ncu: compare SHARES)")
print(f"# GEMM family: {gemm_n} launches, {gemm:.1f} us, share {100 * gemm / tot:.1f}%, DRAM {gemm_b / 1e9:.2f} GB "
      f"({gemm_b / max(gemm_n, 1) / 1e6:.1f} MB per launch)")
has_pipe = any(TP in launch[i] for i in step)
if has_pipe:
    gt = sum(v[3] for k, v in fam.items() if is_gemm(k)) / max(gemm, 1e-9)
    gd = sum(v[4] for k, v in fam.items() if is_gemm(k)) / max(gemm, 1e-9)
    print(f"# GEMM family, time-weighted over its launches: tensor pipe active {gt:.1f} %, DRAM throughput {gd:.1f} % of peak")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    pipe = f"  tensor {v[3] / max(v[1], 1e-9):5.1f}%  dram {v[4] / max(v[1], 1e-9):5.1f}%" if has_pipe else ""
    print(f"{v[1]:10.1f} us {100 * v[1] / tot:5.1f}%  x{v[0]:3d}  {v[2] / 1e9:7.3f} GB{pipe}  {k}")
if "--json" in sys.argv:
    out = sys.argv[sys.argv.index("--json") + 1]
    import hashlib, os
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "regressor_guided_image_editing_b200", "librgie.so")
    sha = hashlib.sha256(open(so, "rb").read()).hexdigest()[:16] if os.path.exists(so) else None
    if "--sha" in sys.argv:          # the build that was profiled, when the in-tree library has been rebuilt since (bench.py prints lib_sha256)
        sha = sys.argv[sys.argv.index("--sha") + 1]
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench                     # the same source hash bench.py matches a capture by (nvcc output is not byte-reproducible)
    json.dump({"lib_sha256": sha, "src_sha256": bench._src_sha(), "step_launches": len(step), "step_us": tot, "gemm_launches": gemm_n, "gemm_us": gemm, "gemm_share": gemm / tot,
               "gemm_dram_bytes": gemm_b, "gemm_dram_bytes_per_launch": gemm_b / max(gemm_n, 1),
               "gemm_tensor_pipe_pct_time_weighted": (gt if has_pipe else None), "gemm_dram_throughput_pct_time_weighted": (gd if has_pipe else None),
               "families": {k: {"launches": v[0], "us": v[1], "dram_bytes": v[2]} for k, v in fam.items()}}, open(out, "w"), indent=1)

"""CPU: the driver-facing contract of bench.py that can be checked without a GPU -- the reference arm's JSON line
(`--impl reference`: the oracle port on the host cores, bounded sample), that its `config` is the native arm's own, that a
non-zero rank of the reference arm exits without work, that the native arm refuses to run without CUDA, and that the
source hash which ties an ncu capture under profiles/ to the running build is the one the newest capture carries."""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ, RGIE_CPU_BUDGET_S="5")
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=timeout)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "edited images/sec (100 steps, 512^2)" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 1.0 / (100 * d["ms_per_step"] / 1e3)) <= 1e-9 * d["value"] + 1e-12
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "1 image x 1" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the config is the native arm's, key for key (what the CPU arm does differently sits under its own key)
    cfg = d["config"]
    assert cfg["workload"].startswith("configs[1]: parametric-filter edit (8 default filters), batch 64 synthetic 512x512")
    assert cfg["precision"] == "bf16" and cfg["micro_batch"] == 32 and cfg["batch_per_gpu"] == 64
    assert cfg["parallelism"].startswith("dp1 ")
    assert d["reference_run"]["precision"].startswith("fp32")


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"},
             timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="CUDA present")
def test_native_arm_refuses_to_run_without_cuda():
    r = _run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_capture_to_build_matching_uses_the_source_hash():
    sys.path.insert(0, ROOT)
    import bench
    ssha = bench._src_sha()
    assert ssha is not None and len(ssha) == 16 and ssha == bench._src_sha()
    caps = [json.load(open(p)) for p in glob.glob(os.path.join(ROOT, "profiles", "*_step_B32.json"))]
    assert any("src_sha256" in c and c.get("gemm_dram_bytes_per_launch", 0) > 0 for c in caps)
    if not any(c.get("src_sha256") == ssha for c in caps):
        pytest.skip("the GEMM-family sources changed since the last ncu launch list under profiles/: bench.py reports "
                    "roofline.traffic = null until tools/run_final.sh + tools/summarize_launches.py are re-run")


def test_bench_inputs_are_the_oracles_inputs():
    """bench.py's native arm draws its synthetic images / regressor weights from bench_inputs.py (not from oracle/); they must be
    the very inputs the parity tests use."""
    sys.path.insert(0, ROOT)
    import bench_inputs as BI
    from oracle import oracle as O
    for idx, h, w in ((0, 64, 48), (63, 512, 512), (300, 256, 256)):
        assert torch.equal(BI.synthetic_image(idx, h, w), O.synthetic_image(idx, h, w))
    for nc in (4, 1):
        a, b = BI.make_regressor_state_dict(nc), O.make_regressor_state_dict(num_classes=nc)
        assert list(a.keys()) == list(b.keys())
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_bench_touches_the_oracle_only_in_the_cpu_leg():
    """Only `cpu_reference_leg` (the cpu_baseline leg and the --impl reference arm) may import oracle/."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    users = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef))]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                users.add(fn.name)
            if isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
                users.add(fn.name)
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.dump(n) for n in top)
    assert users == {"cpu_reference_leg"}, users


def test_roofline_numerator_is_an_independent_flop_count():
    """bench.FLOP_PER_IMAGE_STEP (653.9 GFLOP = 10 crops x 65.39: the numerator of roofline.achieved) against PyTorch's own
    FlopCounterMode over torchvision's resnet50 (fc -> 4) at one 448 x 448 crop: forward + input gradient, no weight gradients."""
    from torch.utils.flop_counter import FlopCounterMode
    from torchvision import models
    sys.path.insert(0, ROOT)
    import bench
    net = models.resnet50()
    net.fc = torch.nn.Linear(2048, 4)
    net.eval()
    for p in net.parameters():
        p.requires_grad_(False)
    x = torch.zeros(1, 3, 448, 448, requires_grad=True)
    with FlopCounterMode(display=False) as fc:
        net(x).sum().backward()
    per_crop = fc.get_total_flops()
    assert abs(10 * per_crop - bench.FLOP_PER_IMAGE_STEP) <= 1e-3 * bench.FLOP_PER_IMAGE_STEP, (per_crop, bench.FLOP_PER_IMAGE_STEP)
    assert bench.STEPS_PER_IMAGE == 100

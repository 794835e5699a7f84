"""The benchmark contract on the CPU side: `bench.py --impl reference` (the CPU arm the driver runs beside ours)
prints one JSON line with the agreed keys, and the GPU arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        "--workload", "cfg2", "--ref-rows", "40"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpix*disp/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "rows" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg2") and "reference_sample" in line["config"]
    assert line["warmup"] == 3                                   # the same warm-up rule as the GPU arm: max(W, 3)


def test_both_arms_share_metric_and_config_keys():
    """The driver pairs the arms by `metric` / `unit` / `config`: both are built by the same code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"metric": METRIC') == 2 and bench.METRIC == "Mpix*disp/s (ASW agg+WTA)"
    assert src.count('"config": make_config_dict(args, W, H, D,') == 2


def test_gpu_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return      # on a GPU box the arm runs (covered by the round-end bench); here only the refusal matters
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "CUDA" in (r.stderr + r.stdout)

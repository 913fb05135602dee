"""The CPU-runnable legs of bench.py, each in a child process as the driver runs them: the reference arm (`--impl reference`: the oracle port
on the host cores) and the torch-CPU figure of cpu_baseline. The GPU legs are exercised on the B200 (profiles/bench_r01_*.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    cp = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert cp.returncode == 0, cp.stderr[-2000:]
    lines = [l for l in cp.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, cp.stdout  # exactly ONE JSON line on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and d["metric"] == "alexnet_ng_conv_fwd_images_per_sec" and d["unit"] == "images/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "alexnet_ng_conv" in d["config"]["workload"] and d["vs_baseline"] is None


def test_torch_cpu_child_figure():
    d = _run(["--torch-cpu-child", "--net", "nin_imagenet", "--batch", "2"])
    assert d["value"] > 0 and d["forwards"] >= 2 and d["cores"] >= 1 and d["seconds"] > 0

"""bench.py's reference arm (the CPU oracle timed on the host cores) runs without a GPU: the JSON contract of the line the
driver parses, alone and under torch.distributed.run with two ranks (rank 0 prints, the other exits 0 without work)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


def _line(out):
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    return json.loads(lines[0])


def _check(d, n):
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["n_gpus"] == n and d["steps"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "views/s" and "512^2 views/sec" in d["metric"] and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("C2") and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "views/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(_line(r.stdout), 1)


def test_reference_arm_under_torchrun_two_ranks():
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29611", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    _check(_line(r.stdout), 2)

"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm) keeps its JSON contract: one line, the metric /
unit / config of the GPU arm, a cpu_baseline object describing the run and an e2e object without copies.  Runs on CPU; under
torchrun only rank 0 works and prints."""
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(line, gpus):
    d = json.loads(line)
    assert d["impl"] == "reference" and d["n_gpus"] == gpus and d["higher_is_better"] is True
    assert d["metric"].startswith("point-clouds/sec") and d["unit"] == "clouds/s" and d["value"] > 0
    assert d["warmup"] >= 3 and d["steps"] == 2 and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "configs[1]" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "clouds/s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "clouds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"], cwd=ROOT, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 1)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(_free_port()), "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    _check(lines[0], 2)

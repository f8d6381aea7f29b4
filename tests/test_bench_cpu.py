"""bench.py's reference arm (the CPU restatement of the reference algorithm on the host cores) follows the contract
without a GPU: one JSON line from rank 0, nothing from the other ranks, the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world):
    env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT="29555")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(world), "--steps", "2",
                           "--warmup", "1", "--cpu-sample", "200000"], env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_line():
    out = _run(0, 1)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle-steps/sec" and d["unit"] == "particle-steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64"
    assert d["value"] > 1e5 and abs(d["ms_per_step"] * 1e-3 * d["value"] - 200000) < 1.0
    assert d["config"]["workload"] == "vp_bump_on_tail_strang_selfconsistent"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    out = _run(1, 2)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ""

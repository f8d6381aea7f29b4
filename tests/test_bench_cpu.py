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
                           "--warmup", "1", "--particles", "200000", "--cpu-sample", "100000"], env=env, capture_output=True, text=True, timeout=300)


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
    # the arm honours --particles / --steps / --warmup (same config as our arm) and reports the serial rate as well
    assert d["config"]["particles_per_gpu"] == 200000 and "sample_particles" not in d["config"] and d["warmup"] == 1
    assert cb["single_thread_value"] > 1e5      # (at this tiny size the threaded rate may fall below the serial one on a busy host)
    for k in ("lb", "clb"):
        w = d["workloads"][k]
        assert w["value"] > 1e4 and w["cpu_baseline"]["kind"] == "port" and w["cpu_baseline"]["cores"] >= 1


def test_config_is_shared_by_both_arms():
    """the driver compares the two arms' `config`: both build it with the same function"""
    sys.path.insert(0, ROOT)
    import bench
    class A: particles, workload, load, field, gpus = 1e8, "vp", "bump_on_tail", "selfconsistent", 1
    cfg = bench.workload_config(A, 1)
    assert cfg["workload"] == "vp_bump_on_tail_strang_selfconsistent" and cfg["particles_per_gpu"] == 100000000
    assert cfg["n_basis"] == 16 and cfg["order"] == 4 and cfg["dt"] == 0.1 and "l2" in cfg


def test_reference_arm_other_ranks_stay_silent():
    out = _run(1, 2)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ""

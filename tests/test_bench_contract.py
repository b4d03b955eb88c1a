"""bench.py's output contract, checked on CPU through the reference arm (the only arm that runs without a GPU):
stdout carries exactly ONE JSON line with the keys the driver reads; the CPU arm describes its bounded sample."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, ZKM_BENCH_CPU_LOG_N="13")            # a 2^13-point sample keeps this test to seconds
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "ms" and j["higher_is_better"] is False
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in j, k
    assert "workload" in j["config"] and "2^24" in j["config"]["workload"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "2^13" in cb["sample"] and cb["value"] == j["value"]
    assert cb["runs"] == j["steps"] == 2 and j["warmup"] == 0 and cb["host_threads"] >= 1
    # the timed region the line claims (steps x ms_per_step of the sample actually run) fits the wall clock of the run
    assert sum(cb["run_ms"]) / 1e3 <= j["wall_s"]


def test_reference_arm_ignores_an_inherited_single_thread_openmp_setting():
    """torchrun exports OMP_NUM_THREADS=1 to its children; the CPU arm passes its thread count explicitly."""
    env = dict(os.environ, ZKM_BENCH_CPU_LOG_N="13", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["cpu_baseline"]["host_threads"] == len(os.sched_getaffinity(0))
    assert j["e2e"] == {"value": j["value"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["vs_baseline"] is None                                # BASELINE.md publishes no number for this metric


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", ZKM_BENCH_CPU_LOG_N="13")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""

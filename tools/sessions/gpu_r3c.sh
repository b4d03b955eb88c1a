#!/bin/bash
# GPU session r3c: the final library of the round once more: whole GPU suite, smoke, bench line (k_pair_bwd at 6 CTAs/SM
# for BN254 is the only kernel change since r2z)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --durations=5 > gpurun_out/pytest_gpu_r3c.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_r3c.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r3c.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r3c.log
timeout 900 python bench.py > gpurun_out/bench_r3c.json 2> gpurun_out/bench_r3c.err
echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_r3c.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "unreg", j["e2e"].get("unregistered_ms"), "frac", j["roofline"]["frac"], "traffic", j["roofline"]["traffic"])
print("stages", j["msm_stage_ms"]); print("ntt", j["ntt"]["ms"], j["ntt"]["int_pipe"]["frac"], j["ntt"]["e2e"]["value"])
print("proxy", j["groth16_proxy"]["proofs_per_s"], "bw6", j["groth16_proxy_bw6_761"]["proofs_per_s"], "marlin", j["marlin_proxy"]["ms_per_proof"], "pre", j["msm_precomputed_bases"]["ms"])
print("cpu", j["cpu_baseline"]["value"], "launches", j["gpu_launches"], "clocks", j["clocks"])
PY
tail -2 gpurun_out/bench_r3c.err

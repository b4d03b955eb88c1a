#!/bin/bash
# GPU session r3i: the library as committed at the end of round 2 (chunked scalar upload behind the histogram pass, x-array
# before the scalars): whole GPU suite, smoke, the default bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --maxfail=15 > gpurun_out/pytest_gpu_r3i.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r3i.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r3i.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r3i.log
timeout 400 python bench.py > gpurun_out/bench_r3i.json 2> gpurun_out/bench_r3i.err
echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_r3i.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", j["e2e"]["value"], "unreg", j["e2e"].get("unregistered_ms"), "frac", j["roofline"]["frac"])
print("stages", j["msm_stage_ms"]); print("ntt", j["ntt"]["ms"], j["ntt"]["e2e"]["value"])
print("proxy", j["groth16_proxy"]["proofs_per_s"], "bw6", j["groth16_proxy_bw6_761"]["proofs_per_s"], "marlin", j["marlin_proxy"]["ms_per_proof"], "pre", j["msm_precomputed_bases"]["ms"])
PY
tail -2 gpurun_out/bench_r3i.err

#!/bin/bash
# GPU session r2z: round-end evidence on the final library of round 2: the whole GPU suite, smoke, the bench line,
# the reference arm, the ncu launch list / DRAM traffic / --set full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --durations=8 > gpurun_out/pytest_gpu_r2z.log 2>&1
echo "pytest rc=$?"; tail -14 gpurun_out/pytest_gpu_r2z.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2z.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_r2z.log
timeout 900 python bench.py > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_r2z.json").read().strip().splitlines()[-1])
    for k in ("value", "e2e", "roofline", "cpu_baseline", "msm_stage_ms", "clocks", "gpu_launches", "msm_precomputed_bases"):
        print(k, json.dumps(j.get(k))[:700])
    print("ntt", json.dumps(j.get("ntt"))[:900])
    for k in ("groth16_proxy", "marlin_proxy", "groth16_proxy_bw6_761"):
        print(k, json.dumps(j.get(k))[:400])
except Exception as e:
    print("no bench json", e)
PY
tail -3 gpurun_out/bench_r2z.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2z.json 2> gpurun_out/bench_ref_r2z.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/bench_ref_r2z.json
bash tools/ncu_final.sh r2z

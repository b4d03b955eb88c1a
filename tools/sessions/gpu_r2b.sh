#!/bin/bash
# GPU session r2b: restructured library -- whole GPU suite (incl. multi-device on one GPU, KZG, verifier acceptance), bench.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --durations=12 -x > gpurun_out/pytest_all_r2b.log 2>&1
echo "pytest rc=$?"; tail -30 gpurun_out/pytest_all_r2b.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    j = json.loads(open("gpurun_out/bench_r2b.json").read().strip().splitlines()[-1])
    for k in ("value", "e2e", "roofline", "cpu_baseline", "msm_stage_ms", "clocks"):
        print(k, json.dumps(j.get(k))[:900])
    print("ntt", json.dumps(j.get("ntt"))[:1500])
except Exception as e:
    print("no bench json", e)
PY
tail -5 gpurun_out/bench_r2b.err

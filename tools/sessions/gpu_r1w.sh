#!/bin/bash
# GPU session r1w: host wait moved behind the accumulation launch: parity + proxies.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_all_r1w.log 2>&1
echo "all rc=$?"; tail -2 gpurun_out/pytest_all_r1w.log
timeout 300 python tools/groth16_proxy.py --log-n 16 --inflight 3 --proofs 90 > gpurun_out/proxy_r1w.json 2> gpurun_out/r1w.err
timeout 300 python tools/groth16_proxy.py --log-n 16 --inflight 1 --proofs 60 >> gpurun_out/proxy_r1w.json 2>> gpurun_out/r1w.err
timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 16 --max 16 --kind witness --precompute --reps 5 >> gpurun_out/proxy_r1w.json 2>> gpurun_out/r1w.err
timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 16 --max 18 --reps 5 >> gpurun_out/proxy_r1w.json 2>> gpurun_out/r1w.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_r1w.json"):
    r = json.loads(l)
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("op", "log_n", "ms", "proofs_per_s", "ms_per_proof", "proofs_in_flight", "precompute")})
PY
tail -3 gpurun_out/r1w.err

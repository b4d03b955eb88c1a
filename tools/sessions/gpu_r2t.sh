#!/bin/bash
# GPU session r2t: lane affinity (a job prefers an idle lane that last served the same registered bases): proxy stability
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py -m gpu -q --maxfail=5 -k "concurrent or proving or batch or multidev or fold" > gpurun_out/pytest_r2t.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2t.log
: > gpurun_out/proxy_r2t.jsonl
for rep in 1 2; do for k in 1 2 4 6 8; do timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 160 --inflight $k >> gpurun_out/proxy_r2t.jsonl 2>> gpurun_out/r2t.err; done; done
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 48 --inflight 4 --curve bw6_761 >> gpurun_out/proxy_r2t.jsonl 2>> gpurun_out/r2t.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_r2t.jsonl"):
    r = json.loads(l); print(r["curve"], r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
tail -3 gpurun_out/r2t.err

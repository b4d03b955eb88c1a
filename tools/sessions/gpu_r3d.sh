#!/bin/bash
# GPU session r3d: lane affinity keyed by (registered part, size class) on the host path as well: Marlin proxy stability
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_kzg.py tests/test_gpu_multidev.py -m gpu -q -x -k "concurrent or proving or batch or multidev or kzg" > gpurun_out/pytest_r3d.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r3d.log
: > gpurun_out/marlin_r3d.jsonl; : > gpurun_out/proxy_r3d.jsonl
for rep in 1 2 3; do timeout 300 python tools/marlin_proxy.py --log-h 16 --log-k 18 --proofs 6 >> gpurun_out/marlin_r3d.jsonl 2>> gpurun_out/r3d.err; done
timeout 300 python tools/marlin_proxy.py --log-h 16 --log-k 18 --proofs 6 --curve bw6_761 >> gpurun_out/marlin_r3d.jsonl 2>> gpurun_out/r3d.err
for k in 1 4 8; do timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 160 --inflight $k >> gpurun_out/proxy_r3d.jsonl 2>> gpurun_out/r3d.err; done
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 48 --inflight 4 --curve bw6_761 >> gpurun_out/proxy_r3d.jsonl 2>> gpurun_out/r3d.err
python - <<'PY'
import json
for l in open("gpurun_out/marlin_r3d.jsonl"):
    r = json.loads(l); print("marlin", r["curve"], round(r["ms_per_proof"], 2), r["kernel_launches_per_proof"])
for l in open("gpurun_out/proxy_r3d.jsonl"):
    r = json.loads(l); print("groth16", r["curve"], r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1))
PY
tail -2 gpurun_out/r3d.err

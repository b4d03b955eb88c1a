#!/bin/bash
# GPU session r2n: stream-ordered workspace growth (no device-wide stalls): parity incl. multi-device, proxy vs proofs in flight
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py tests/test_kzg.py -m gpu -q -x -k "not bw6" > gpurun_out/pytest_r2n.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2n.log
: > gpurun_out/proxy_conn_r2n.jsonl
for conn in 8 32; do for k in 1 2 3 4 6 8; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 120 --inflight $k >> gpurun_out/proxy_conn_r2n.jsonl 2>> gpurun_out/r2n.err
done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 64 --inflight 8 --serial >> gpurun_out/proxy_conn_r2n.jsonl 2>> gpurun_out/r2n.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 64 --inflight 16 --serial >> gpurun_out/proxy_conn_r2n.jsonl 2>> gpurun_out/r2n.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_conn_r2n.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], r["concurrent_msms"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2n.err; }
sw sweep_msm_bls12_381_g1_r2n.jsonl msm --curve bls12_381 --min 22 --max 26
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_msm_*_r2n.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][10:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
tail -3 gpurun_out/r2n.err

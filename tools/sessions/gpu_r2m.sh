#!/bin/bash
# GPU session r2m: three-class fold (parity), hardware work queues (CUDA_DEVICE_MAX_CONNECTIONS) vs proofs in flight,
# sweeps that regressed in r2l (BW6-761 2^16 fold, 2^24 sort).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "msm" > gpurun_out/pytest_r2m.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2m.log
: > gpurun_out/proxy_conn_r2m.jsonl
for conn in 8 32; do for k in 2 4 8; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 120 --inflight $k >> gpurun_out/proxy_conn_r2m.jsonl 2>> gpurun_out/r2m.err
done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 64 --inflight 4 --serial >> gpurun_out/proxy_conn_r2m.jsonl 2>> gpurun_out/r2m.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 64 --inflight 8 --serial >> gpurun_out/proxy_conn_r2m.jsonl 2>> gpurun_out/r2m.err
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 64 --inflight 16 --serial >> gpurun_out/proxy_conn_r2m.jsonl 2>> gpurun_out/r2m.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_conn_r2m.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], r["concurrent_msms"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2m.err; }
sw sweep_msm_bw6_761_g1_r2m.jsonl msm --curve bw6_761 --min 16 --max 18
sw sweep_msm_bls12_381_g1_r2m.jsonl msm --curve bls12_381 --min 24 --max 24
sw sweep_msm_bls12_381_g2_r2m.jsonl msm --curve bls12_381 --group 2 --min 16 --max 18
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_msm_*_r2m.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][10:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
nproc; tail -3 gpurun_out/r2m.err

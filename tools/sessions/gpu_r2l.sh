#!/bin/bash
# GPU session r2l: lane policy (same stream first, idle next): proxy vs proofs in flight, sweeps incl. 2^23 / 2^24.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py -m gpu -q -x -k "msm or concurrent or proving or multidev" > gpurun_out/pytest_r2l.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2l.log
: > gpurun_out/proxy_inflight_r2l.jsonl
for k in 1 2 3 4 6 8; do
  timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 120 --inflight $k >> gpurun_out/proxy_inflight_r2l.jsonl 2>> gpurun_out/r2l.err
done
timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 48 --inflight 4 --curve bw6_761 >> gpurun_out/proxy_inflight_r2l.jsonl 2>> gpurun_out/r2l.err
python - <<'PY'
import json
for l in open("gpurun_out/proxy_inflight_r2l.jsonl"):
    r = json.loads(l); print(r["curve"], r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2l.err; }
sw sweep_msm_bls12_381_g1_r2l.jsonl msm --curve bls12_381 --min 16 --max 24
sw sweep_msm_bls12_381_g1_witness_r2l.jsonl msm --curve bls12_381 --min 16 --max 24 --kind witness
sw sweep_msm_bls12_381_g1_pre_r2l.jsonl msm --curve bls12_381 --min 16 --max 18 --kind witness --precompute
sw sweep_msm_bw6_761_g1_r2l.jsonl msm --curve bw6_761 --min 16 --max 18
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_msm_*_r2l.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][10:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
tail -3 gpurun_out/r2l.err

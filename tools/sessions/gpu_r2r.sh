#!/bin/bash
# GPU session r2r: NULL-stream calls on one library stream per device, leader-key aggregation in the per-window scatter:
# whole GPU suite, witness / uniform 2^24, proxy.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --durations=5 > gpurun_out/pytest_gpu_r2r.log 2>&1
echo "pytest rc=$?"; tail -9 gpurun_out/pytest_gpu_r2r.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r2r.err; }
for f in sweep_msm_bls12_381_g1_r2r sweep_msm_bls12_381_g1_witness_r2r; do : > gpurun_out/$f.jsonl; done
sw sweep_msm_bls12_381_g1_r2r.jsonl msm --curve bls12_381 --min 23 --max 24
sw sweep_msm_bls12_381_g1_witness_r2r.jsonl msm --curve bls12_381 --min 22 --max 24 --kind witness
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2r.jsonl")):
    for l in open(f):
        r = json.loads(l)
        print(f.split("/")[-1][6:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
: > gpurun_out/proxy_r2r.jsonl
for k in 1 4 6 8; do timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 120 --inflight $k >> gpurun_out/proxy_r2r.jsonl 2>> gpurun_out/r2r.err; done
python - <<'PY'
import json
for l in open("gpurun_out/proxy_r2r.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
tail -3 gpurun_out/r2r.err

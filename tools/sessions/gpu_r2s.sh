#!/bin/bash
# GPU session r2s: leader-key aggregation in the histogram / one-launch scatter as well: msm parity, witness + uniform sweeps, proxy
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_groth16_proof.py -m gpu -q --maxfail=5 -k "msm or golden or proof or kzg" > gpurun_out/pytest_r2s.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_r2s.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r2s.err; }
for f in sweep_msm_bls12_381_g1_r2s sweep_msm_bls12_381_g1_witness_r2s sweep_msm_bls12_381_g1_pre_r2s; do : > gpurun_out/$f.jsonl; done
sw sweep_msm_bls12_381_g1_r2s.jsonl msm --curve bls12_381 --min 16 --max 24
sw sweep_msm_bls12_381_g1_witness_r2s.jsonl msm --curve bls12_381 --min 16 --max 24 --kind witness
sw sweep_msm_bls12_381_g1_pre_r2s.jsonl msm --curve bls12_381 --min 16 --max 18 --kind witness --precompute
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2s.jsonl")):
    for l in open(f):
        r = json.loads(l)
        print(f.split("/")[-1][6:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
: > gpurun_out/proxy_r2s.jsonl
for k in 4 6; do timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 120 --inflight $k >> gpurun_out/proxy_r2s.jsonl 2>> gpurun_out/r2s.err; done
python - <<'PY'
import json
for l in open("gpurun_out/proxy_r2s.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
tail -3 gpurun_out/r2s.err

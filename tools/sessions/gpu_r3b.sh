#!/bin/bash
# GPU session r3b: k_accum_affine at 4 CTAs per SM (G1 fields), k_pair_bwd at 6 for BN254: parity, sweeps, proxy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -q -x -k "msm or golden" > gpurun_out/pytest_r3b.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r3b.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r3b.err; }
for f in sweep_msm_bls12_381_g1_r3b sweep_msm_bn254_g1_r3b sweep_msm_bls12_381_g1_witness_r3b; do : > gpurun_out/$f.jsonl; done
sw sweep_msm_bls12_381_g1_r3b.jsonl msm --curve bls12_381 --min 16 --max 24
sw sweep_msm_bn254_g1_r3b.jsonl msm --curve bn254 --min 20 --max 24
sw sweep_msm_bls12_381_g1_witness_r3b.jsonl msm --curve bls12_381 --min 20 --max 24 --kind witness
python - <<'PY'
import json
for f in ["sweep_msm_bls12_381_g1_r3b", "sweep_msm_bn254_g1_r3b", "sweep_msm_bls12_381_g1_witness_r3b"]:
    for l in open("gpurun_out/%s.jsonl" % f):
        r = json.loads(l); print(f[10:-4], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
: > gpurun_out/proxy_r3b.jsonl
for k in 1 4 8; do timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 160 --inflight $k >> gpurun_out/proxy_r3b.jsonl 2>> gpurun_out/r3b.err; done
python - <<'PY'
import json
for l in open("gpurun_out/proxy_r3b.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1))
PY
tail -2 gpurun_out/r3b.err

#!/bin/bash
# GPU session r3a: k_pair_bwd at 5 (BLS12-381 G1) / 6 (BN254 G1) CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "affine or known_discrete" > gpurun_out/pytest_r3a.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r3a.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r3a.err; }
: > gpurun_out/sweep_msm_bls12_381_g1_r3a.jsonl; : > gpurun_out/sweep_msm_bn254_g1_r3a.jsonl
sw sweep_msm_bls12_381_g1_r3a.jsonl msm --curve bls12_381 --min 21 --max 24
sw sweep_msm_bn254_g1_r3a.jsonl msm --curve bn254 --min 22 --max 24
python - <<'PY'
import json
for f in ["sweep_msm_bls12_381_g1_r3a", "sweep_msm_bn254_g1_r3a"]:
    for l in open("gpurun_out/%s.jsonl" % f):
        r = json.loads(l); print(f[10:-4], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
PY
tail -2 gpurun_out/r3a.err

#!/bin/bash
# GPU session r1u: skew-aware window model: parity, then full automatic sweeps (compare with profiles/sweep_msm_*_r1.jsonl).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_all_r1u.log 2>&1
echo "all rc=$?"; tail -3 gpurun_out/pytest_all_r1u.log
timeout 900 python tools/sweep.py msm --curve bls12_381 --group 1 --min 14 --max 25 --reps 3 > gpurun_out/sweep_msm_bls12_381_g1_r1u.jsonl 2> gpurun_out/r1u.err
timeout 600 python tools/sweep.py msm --curve bn254 --group 1 --min 14 --max 24 --reps 3 > gpurun_out/sweep_msm_bn254_g1_r1u.jsonl 2>> gpurun_out/r1u.err
timeout 600 python tools/sweep.py msm --curve bw6_761 --group 1 --min 14 --max 22 --reps 3 > gpurun_out/sweep_msm_bw6_761_g1_r1u.jsonl 2>> gpurun_out/r1u.err
timeout 600 python tools/sweep.py msm --curve bls12_381 --group 2 --min 14 --max 20 --reps 3 > gpurun_out/sweep_msm_bls12_381_g2_r1u.jsonl 2>> gpurun_out/r1u.err
python - <<'PY'
import json
for f in ("bls12_381_g1", "bn254_g1", "bw6_761_g1", "bls12_381_g2"):
    print(f, [(r["log_n"], r["window_bits"], round(r["ms"], 2)) for r in map(json.loads, open("gpurun_out/sweep_msm_%s_r1u.jsonl" % f))])
PY
tail -3 gpurun_out/r1u.err

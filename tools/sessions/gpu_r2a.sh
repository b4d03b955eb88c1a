#!/bin/bash
# GPU session r2a: full-size parity tests (new), regression of the round-1 suite, sweeps with a `check` per row.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=20 --durations=15 > gpurun_out/pytest_large_r2a.log 2>&1
echo "large rc=$?"; tail -25 gpurun_out/pytest_large_r2a.log
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 --deselect tests/test_gpu_large.py > gpurun_out/pytest_rest_r2a.log 2>&1
echo "rest rc=$?"; tail -3 gpurun_out/pytest_rest_r2a.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 3 > gpurun_out/$out 2>> gpurun_out/r2a.err; }
sw sweep_msm_bls12_381_g1_r2a.jsonl msm --curve bls12_381 --group 1 --min 16 --max 26
sw sweep_msm_bn254_g1_r2a.jsonl msm --curve bn254 --group 1 --min 16 --max 26
sw sweep_ntt_bls12_381_r2a.jsonl ntt --curve bls12_381 --min 16 --max 26
sw sweep_ntt_bn254_r2a.jsonl ntt --curve bn254 --min 16 --max 26
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2a.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][6:-11], r["log_n"], round(r.get("ms", r.get("fft_ms")), 3), r.get("check"))
PY
tail -5 gpurun_out/r2a.err

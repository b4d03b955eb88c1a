#!/bin/bash
# GPU session r2d: fused reduction products for the p = 1 mod 2^32 fields (runtime INV) + unit-twiddle skip in the NTT:
# parity of everything that multiplies in Fr, then before/after timings.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_kzg.py tests/test_groth16_verify.py -m gpu -q --maxfail=10 -k "ntt or golden or witness or kzg or verifier or bw6" > gpurun_out/pytest_r2d.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/pytest_r2d.log
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -k "ntt or (msm and (24 or 22))" > gpurun_out/pytest_large_r2d.log 2>&1
echo "large rc=$?"; tail -3 gpurun_out/pytest_large_r2d.log
# hierarchical bucket reduction: every MSM test (all window sizes / curves / groups)
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py tests/test_groth16_proof.py -m gpu -q --maxfail=10 -k "msm or proof or shard or proving" > gpurun_out/pytest_msm_r2d.log 2>&1
echo "msm rc=$?"; tail -4 gpurun_out/pytest_msm_r2d.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 > gpurun_out/$out 2>> gpurun_out/r2d.err; }
sw sweep_ntt_bls12_381_r2d.jsonl ntt --curve bls12_381 --min 16 --max 26
sw sweep_ntt_bn254_r2d.jsonl ntt --curve bn254 --min 20 --max 24
sw sweep_ntt_bw6_761_r2d.jsonl ntt --curve bw6_761 --min 16 --max 24
sw sweep_msm_bls12_381_g1_r2d.jsonl msm --curve bls12_381 --group 1 --min 16 --max 24
sw sweep_msm_bn254_g1_r2d.jsonl msm --curve bn254 --group 1 --min 20 --max 24
sw sweep_msm_bw6_761_g1_r2d.jsonl msm --curve bw6_761 --group 1 --min 16 --max 20
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2d.jsonl")):
    for l in open(f):
        r = json.loads(l); print(f.split("/")[-1][6:-11], r["log_n"], round(r.get("ms", r.get("fft_ms")), 3), r.get("check"))
PY
tail -3 gpurun_out/r2d.err

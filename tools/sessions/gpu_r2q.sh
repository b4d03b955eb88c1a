#!/bin/bash
# GPU session r2q: Montgomery reduction without the m * p_0 / m * p_1 products (BLS12-381 Fr, BW6-761 Fr) and digit words
# for the per-window scatter: the whole GPU suite, then NTT and MSM sweeps.
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --durations=5 > gpurun_out/pytest_gpu_r2q.log 2>&1
echo "pytest rc=$?"; tail -9 gpurun_out/pytest_gpu_r2q.log
sw() { out=$1; shift; timeout 900 python tools/sweep.py "$@" --reps 5 >> gpurun_out/$out 2>> gpurun_out/r2q.err; }
for f in sweep_ntt_bls12_381_r2q sweep_ntt_bw6_761_r2q sweep_ntt_bn254_r2q sweep_msm_bls12_381_g1_r2q sweep_msm_bls12_381_g1_witness_r2q; do : > gpurun_out/$f.jsonl; done
sw sweep_ntt_bls12_381_r2q.jsonl ntt --curve bls12_381 --min 16 --max 26
sw sweep_ntt_bw6_761_r2q.jsonl ntt --curve bw6_761 --min 20 --max 24
sw sweep_ntt_bn254_r2q.jsonl ntt --curve bn254 --min 24 --max 24
sw sweep_msm_bls12_381_g1_r2q.jsonl msm --curve bls12_381 --min 22 --max 26
sw sweep_msm_bls12_381_g1_witness_r2q.jsonl msm --curve bls12_381 --min 24 --max 24 --kind witness
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/sweep_*_r2q.jsonl")):
    for l in open(f):
        r = json.loads(l)
        if "stage_ms" in r or "window_bits" in r:
            print(f.split("/")[-1][6:-10], r["log_n"], round(r["ms"], 3), r.get("window_bits"), {k: round(v, 2) for k, v in (r.get("stage_ms") or {}).items()}, r.get("check"))
        else:
            print(f.split("/")[-1][6:-10], r["log_n"], {k: round(v, 3) for k, v in r.items() if k.endswith("ms")}, r.get("check"))
PY
tail -3 gpurun_out/r2q.err

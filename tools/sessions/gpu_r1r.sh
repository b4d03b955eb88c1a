#!/bin/bash
# GPU session r1r: measured auto rule for the batched-affine levels: parity, then before/after on every group.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_all_r1r.log 2>&1
echo "all rc=$?"; tail -3 gpurun_out/pytest_all_r1r.log
: > gpurun_out/affine_rule_r1r.jsonl
run() { timeout 400 python tools/sweep.py msm "$@" --reps 3 >> gpurun_out/affine_rule_r1r.jsonl 2>> gpurun_out/r1r.err; }
run --curve bls12_381 --group 1 --min 20 --max 24
run --curve bw6_761 --group 1 --min 17 --max 21
run --curve bw6_761 --group 1 --min 17 --max 20 --opt msm_affine_levels=0
run --curve bn254 --group 1 --min 21 --max 24
run --curve bn254 --group 1 --min 21 --max 23 --opt msm_affine_levels=0
run --curve bls12_381 --group 2 --min 18 --max 20
run --curve bls12_381 --group 2 --min 18 --max 20 --opt msm_affine_levels=0
python - <<'PY'
import json
for l in open("gpurun_out/affine_rule_r1r.jsonl"):
    r = json.loads(l); print(r["curve"], r["group"], r["log_n"], r.get("opts", ""), r["window_bits"], round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/r1r.err

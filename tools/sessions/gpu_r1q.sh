#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/affine_levels_r1q.jsonl
for lg in 22 23; do
  for lv in 2 3 4 5 6; do
    timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min $lg --max $lg --reps 3 --opt msm_affine_levels=$lv >> gpurun_out/affine_levels_r1q.jsonl 2>> gpurun_out/r1q.err
  done
done
timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min 21 --max 21 --reps 3 --opt msm_affine_levels=1 >> gpurun_out/affine_levels_r1q.jsonl 2>> gpurun_out/r1q.err
for lv in 3 5; do timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min 24 --max 24 --reps 3 --opt msm_affine_levels=$lv >> gpurun_out/affine_levels_r1q.jsonl 2>> gpurun_out/r1q.err; done
python - <<'PY'
import json
for l in open("gpurun_out/affine_levels_r1q.jsonl"):
    r = json.loads(l); print(r["log_n"], r.get("opts"), r["window_bits"], round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/r1q.err

#!/bin/bash
# GPU session r1p: where should the batched-affine levels start?  2^18..2^21, forced level counts and window sizes.
mkdir -p gpurun_out
: > gpurun_out/affine_threshold_r1p.jsonl
for lg in 18 19 20 21; do
  for lv in 0 2 3 4 5; do
    timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min $lg --max $lg --reps 3 --opt msm_affine_levels=$lv >> gpurun_out/affine_threshold_r1p.jsonl 2>> gpurun_out/r1p.err
  done
done
for lg in 20 21; do
  for c in 17 18; do
    for lv in 0 4 5; do
      timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min $lg --max $lg --reps 3 --opt msm_affine_levels=$lv --opt msm_window_bits=$c >> gpurun_out/affine_threshold_r1p.jsonl 2>> gpurun_out/r1p.err
    done
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/affine_threshold_r1p.jsonl"):
    r = json.loads(l); print(r["log_n"], r.get("opts"), round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/r1p.err

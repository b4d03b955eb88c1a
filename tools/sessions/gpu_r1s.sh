#!/bin/bash
# GPU session r1s: window size around the automatic choice for 2^20..2^23 now that the affine rule moved.
mkdir -p gpurun_out
: > gpurun_out/window_tune_r1s.jsonl
for lg in 20 21 22 23; do
  for c in 14 15 16 17 18 19 20; do
    timeout 200 python tools/sweep.py msm --curve bls12_381 --group 1 --min $lg --max $lg --reps 3 --opt msm_window_bits=$c >> gpurun_out/window_tune_r1s.jsonl 2>> gpurun_out/r1s.err
  done
done
python - <<'PY'
import json
for l in open("gpurun_out/window_tune_r1s.jsonl"):
    r = json.loads(l); print(r["log_n"], r.get("opts", ""), round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/r1s.err

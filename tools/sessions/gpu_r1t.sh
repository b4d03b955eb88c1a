#!/bin/bash
# GPU session r1t: is c = 16 (full top window, 32-bit aligned digits) the sweet spot elsewhere too?
mkdir -p gpurun_out
: > gpurun_out/window16_r1t.jsonl
run() { timeout 400 python tools/sweep.py msm "$@" --reps 3 >> gpurun_out/window16_r1t.jsonl 2>> gpurun_out/r1t.err; }
for lg in 17 18 19; do for c in 12 16; do run --curve bls12_381 --group 1 --min $lg --max $lg --opt msm_window_bits=$c; done; done
for lg in 24 25; do for c in 16; do run --curve bls12_381 --group 1 --min $lg --max $lg --opt msm_window_bits=$c; done; done
for lg in 22 23 24; do run --curve bn254 --group 1 --min $lg --max $lg --opt msm_window_bits=16; done
run --curve bn254 --group 1 --min 24 --max 24 --opt msm_window_bits=20
for c in 14 18 19 20; do run --curve bw6_761 --group 1 --min 20 --max 20 --opt msm_window_bits=$c; done
for c in 16 19 20; do run --curve bw6_761 --group 1 --min 22 --max 22 --opt msm_window_bits=$c; done
python - <<'PY'
import json
for l in open("gpurun_out/window16_r1t.jsonl"):
    r = json.loads(l); print(r["curve"], r["log_n"], r.get("opts", ""), round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/r1t.err

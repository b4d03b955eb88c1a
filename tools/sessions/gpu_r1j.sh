#!/bin/bash
# GPU session r1j: single-site reduce loop + L2 prefetch of the level-0 gathers: parity, then 2^24 timings per prefetch distance.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/pytest_all_r1j.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/pytest_all_r1j.log
tail -3 gpurun_out/pytest_all_r1j.log
: > gpurun_out/prefetch_tune_r1j.jsonl
for cfg in "0 0" "2 1" "1 1" "3 1" "2 2" "4 2" "2 0" "0 1"; do
  set -- $cfg
  timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 24 --max 24 --reps 3 --opt msm_prefetch_fwd=$1 --opt msm_prefetch_bwd=$2 >> gpurun_out/prefetch_tune_r1j.jsonl 2>> gpurun_out/sweep_r1j.err
done
python - <<'PY'
import json
for l in open("gpurun_out/prefetch_tune_r1j.jsonl"):
    r = json.loads(l); print(r.get("opts"), round(r["ms"], 2), r["stage_ms"])
PY
tail -3 gpurun_out/sweep_r1j.err

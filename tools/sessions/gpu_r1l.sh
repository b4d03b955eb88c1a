#!/bin/bash
# GPU session r1l: regression check of the small-MSM / NTT / proxy numbers after the BW6-761 generalisation.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/box_r1l.txt; nproc >> gpurun_out/box_r1l.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/box_r1l.txt
cat gpurun_out/box_r1l.txt
timeout 300 python tools/sweep.py ntt --curve bls12_381 --min 22 --max 24 --reps 5 > gpurun_out/ntt_check_r1l.jsonl 2> gpurun_out/r1l.err
timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 16 --max 16 --kind witness --precompute --reps 5 > gpurun_out/msm16_check_r1l.jsonl 2>> gpurun_out/r1l.err
timeout 300 python tools/sweep.py msm --curve bls12_381 --group 2 --min 16 --max 16 --kind witness --precompute --reps 5 >> gpurun_out/msm16_check_r1l.jsonl 2>> gpurun_out/r1l.err
timeout 300 python tools/sweep.py msm --curve bls12_381 --group 1 --min 16 --max 16 --reps 5 >> gpurun_out/msm16_check_r1l.jsonl 2>> gpurun_out/r1l.err
timeout 600 python tools/groth16_proxy.py --log-n 16 --inflight 3 --proofs 90 > gpurun_out/proxy_check_r1l.json 2>> gpurun_out/r1l.err
timeout 600 python tools/groth16_proxy.py --log-n 16 --inflight 1 --proofs 60 >> gpurun_out/proxy_check_r1l.json 2>> gpurun_out/r1l.err
python - <<'PY'
import json
for f in ("ntt_check_r1l.jsonl", "msm16_check_r1l.jsonl", "proxy_check_r1l.json"):
    for l in open("gpurun_out/" + f):
        r = json.loads(l)
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items() if k in ("op", "log_n", "group", "fft_ms", "coset_fft_ms", "ms", "stage_ms", "proofs_per_s", "ms_per_proof", "proofs_in_flight", "precompute")})
PY
tail -3 gpurun_out/r1l.err

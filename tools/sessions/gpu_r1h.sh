#!/bin/bash
# GPU session r1h: BW6-761 parity first, then the whole GPU suite, then BW6-761 sweeps.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "bw6_761 or golden or host_cpp" > gpurun_out/pytest_bw6_r1h.log 2>&1
echo "bw6 rc=$?" | tee -a gpurun_out/pytest_bw6_r1h.log
tail -5 gpurun_out/pytest_bw6_r1h.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_all_r1h.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/pytest_all_r1h.log
tail -5 gpurun_out/pytest_all_r1h.log
timeout 600 python tools/sweep.py msm --curve bw6_761 --group 1 --min 14 --max 22 --reps 3 > gpurun_out/sweep_msm_bw6_761_g1_r1.jsonl 2> gpurun_out/sweep_bw6.err
timeout 300 python tools/sweep.py msm --curve bw6_761 --group 2 --min 16 --max 16 --kind witness --reps 3 > gpurun_out/sweep_msm_bw6_761_g2_r1.jsonl 2>> gpurun_out/sweep_bw6.err
timeout 600 python tools/sweep.py ntt --curve bw6_761 --min 14 --max 24 --reps 3 > gpurun_out/sweep_ntt_bw6_761_r1.jsonl 2>> gpurun_out/sweep_bw6.err
tail -3 gpurun_out/sweep_msm_bw6_761_g1_r1.jsonl; tail -2 gpurun_out/sweep_ntt_bw6_761_r1.jsonl; tail -5 gpurun_out/sweep_bw6.err

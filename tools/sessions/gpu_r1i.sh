#!/bin/bash
# GPU session r1i: full GPU suite (BW6-761 included), smoke, BW6-761 sweeps + Groth16 proxy, per-kernel times of a BW6 2^16 MSM.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/pytest_all_r1i.log 2>&1
echo "all rc=$?" | tee -a gpurun_out/pytest_all_r1i.log
tail -4 gpurun_out/pytest_all_r1i.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/sweep.py msm --curve bw6_761 --group 1 --min 14 --max 22 --reps 3 > gpurun_out/sweep_msm_bw6_761_g1_r1.jsonl 2> gpurun_out/sweep_bw6.err
timeout 300 python tools/sweep.py msm --curve bw6_761 --group 1 --min 15 --max 18 --kind witness --precompute --reps 3 > gpurun_out/sweep_msm_bw6_761_g1_witness_precomputed_r1.jsonl 2>> gpurun_out/sweep_bw6.err
timeout 300 python tools/sweep.py msm --curve bw6_761 --group 2 --min 16 --max 16 --kind witness --reps 3 > gpurun_out/sweep_msm_bw6_761_g2_r1.jsonl 2>> gpurun_out/sweep_bw6.err
timeout 600 python tools/sweep.py ntt --curve bw6_761 --min 14 --max 24 --reps 3 --cpu > gpurun_out/sweep_ntt_bw6_761_r1.jsonl 2>> gpurun_out/sweep_bw6.err
timeout 600 python tools/groth16_proxy.py --curve bw6_761 --log-n 16 > gpurun_out/proxy_bw6_761_r1.json 2> gpurun_out/proxy_bw6.err
tail -2 gpurun_out/proxy_bw6_761_r1.json; tail -3 gpurun_out/proxy_bw6.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bw6_2p16_r1.csv python tools/profile_target.py 16 2 > gpurun_out/ncu_bw6.log 2>&1
tail -3 gpurun_out/sweep_msm_bw6_761_g1_r1.jsonl; cat gpurun_out/sweep_msm_bw6_761_g1_witness_precomputed_r1.jsonl; tail -5 gpurun_out/sweep_bw6.err

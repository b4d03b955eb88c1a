#!/bin/bash
# GPU session r1n: Groth16 proof bytes vs the exact restatement, Marlin-shaped proxy on both bench curves.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_groth16_proof.py -q -m gpu > gpurun_out/pytest_proof_r1n.log 2>&1
echo "proof rc=$?"; tail -15 gpurun_out/pytest_proof_r1n.log
timeout 900 python tools/marlin_proxy.py --log-h 16 --log-k 18 --proofs 6 --cpu > gpurun_out/marlin_proxy_bls12_381_r1.json 2> gpurun_out/marlin.err
echo "rc=$?"; cat gpurun_out/marlin_proxy_bls12_381_r1.json; tail -3 gpurun_out/marlin.err
timeout 900 python tools/marlin_proxy.py --curve bw6_761 --log-h 16 --log-k 18 --proofs 4 --cpu > gpurun_out/marlin_proxy_bw6_761_r1.json 2>> gpurun_out/marlin.err
echo "rc=$?"; cat gpurun_out/marlin_proxy_bw6_761_r1.json; tail -3 gpurun_out/marlin.err

#!/bin/bash
# GPU session r2c: the rest of the GPU suite after the gather-alignment fix (test_gpu_large was green in r2b).
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 --durations=8 --deselect tests/test_gpu_large.py > gpurun_out/pytest_rest_r2c.log 2>&1
echo "pytest rc=$?"; tail -40 gpurun_out/pytest_rest_r2c.log

#!/bin/bash
# GPU session r3h: chunked scalar upload with the histogram pass behind it (host entry points, >= 2^22 scalars)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multidev.py tests/test_gpu_large.py -m gpu -q -x -k "known_discrete_log_large or non_canonical or multidev or (full_size and 24)" > gpurun_out/pytest_r3h.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_r3h.log
timeout 600 python bench.py --skip-cpu --skip-proxy --skip-ntt --skip-precompute > gpurun_out/bench_r3h.json 2> gpurun_out/bench_r3h.err
echo "bench rc=$?"; python - <<'PY'
import json
j = json.loads(open("gpurun_out/bench_r3h.json").read().strip().splitlines()[-1])
print("value", j["value"], "e2e", json.dumps(j["e2e"])[:700])
PY
tail -2 gpurun_out/bench_r3h.err

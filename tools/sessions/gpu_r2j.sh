mkdir -p gpurun_out
: > gpurun_out/proxy_conn_r2j.jsonl
for conn in 8 32; do for k in 1 2 3 4; do
  CUDA_DEVICE_MAX_CONNECTIONS=$conn timeout 300 python tools/groth16_proxy.py --log-n 16 --proofs 96 --inflight $k >> gpurun_out/proxy_conn_r2j.jsonl 2>> gpurun_out/r2j.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/proxy_conn_r2j.jsonl"):
    r = json.loads(l); print(r["proofs_in_flight"], round(r["ms_per_proof"], 3), round(r["proofs_per_s"], 1), r["kernel_launches_per_proof"])
PY
tail -3 gpurun_out/r2j.err

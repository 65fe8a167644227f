set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "exchange or slab" > gpurun_out/pytest_peer_r02u.log 2>&1; tail -5 gpurun_out/pytest_peer_r02u.log
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2956$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${n}gpu_r02u.json 2> gpurun_out/bench_${n}gpu_r02u.err; tail -c 300 gpurun_out/bench_${n}gpu_r02u.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${n}gpu_r02u.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], {k:v for k,v in d["parity"].items() if k not in ("per_k","checker")}); print(d["details"].get("e2e_stages_ms")); print(d["details"].get("value_stages_ms"))
PY
done

set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02r.log 2>&1; tail -4 gpurun_out/pytest_r02r.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${n}gpu_r02r.json 2> gpurun_out/bench_${n}gpu_r02r.err; tail -c 400 gpurun_out/bench_${n}gpu_r02r.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${n}gpu_r02r.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], {k:v for k,v in d["parity"].items() if k not in ("per_k","checker")}); print(d["details"].get("e2e_stages_ms")); print(d["details"].get("value_stages_ms")); print(d["details"]["parallelism"][-200:])
PY
done

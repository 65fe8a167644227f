set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02y.log 2>&1; tail -4 gpurun_out/pytest_r02y.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_1gpu_r02y.json 2> gpurun_out/bench_1gpu_r02y.err; tail -c 300 gpurun_out/bench_1gpu_r02y.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_1gpu_r02y.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["ms_per_step"], {k:v for k,v in d["parity"].items() if k not in ("per_k","checker")}, d["roofline"]["frac"], d["clocks"])
PY

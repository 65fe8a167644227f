set -x
for w in bunny_knn c1_torus; do
  timeout 600 python bench.py --workload $w > gpurun_out/bench_${w}_r02p.json 2> gpurun_out/bench_${w}_r02p.err; tail -c 400 gpurun_out/bench_${w}_r02p.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${w}_r02p.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"], d.get("parity"), d["cpu_baseline"]["value"], d["cpu_baseline"]["seconds"])
PY
done
timeout 300 python bench.py --impl reference --workload bunny_knn --steps 2 --warmup 1 > gpurun_out/bench_bunny_knn_ref_r02p.json 2>/dev/null; cut -c1-300 gpurun_out/bench_bunny_knn_ref_r02p.json
timeout 900 python bench.py > gpurun_out/bench_1gpu_r02p.json 2> gpurun_out/bench_1gpu_r02p.err; tail -c 300 gpurun_out/bench_1gpu_r02p.err; cut -c1-700 gpurun_out/bench_1gpu_r02p.json

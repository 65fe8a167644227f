# what a GPU-box call of this repo usually runs (gpurun -- 'bash scripts/_gpu_call.sh'); outputs under gpurun_out/
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -c 300 gpurun_out/bench_1gpu.err; cut -c1-300 gpurun_out/bench_1gpu.json

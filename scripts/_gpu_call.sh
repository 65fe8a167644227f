set -x
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu_r02g.txt
python scripts/qbench.py 2e7 20,32 6 > gpurun_out/qbench_r02g.log 2>&1; cat gpurun_out/qbench_r02g.log | cut -c1-200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02g.log 2>&1; tail -15 gpurun_out/pytest_r02g.log
PCT_B200_TRACE=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_1gpu_r02g.json 2> gpurun_out/bench_1gpu_r02g.err; tail -5 gpurun_out/bench_1gpu_r02g.err; cut -c1-1500 gpurun_out/bench_1gpu_r02g.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_2gpu_r02g.json 2> gpurun_out/bench_2gpu_r02g.err; tail -5 gpurun_out/bench_2gpu_r02g.err; cat gpurun_out/bench_2gpu_r02g.json

set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "exchange" > gpurun_out/pytest_peer_r03a.log 2>&1; tail -5 gpurun_out/pytest_peer_r03a.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 3 --warmup 3 --no-parity > gpurun_out/bench_2gpu_r03a.json 2> gpurun_out/bench_2gpu_r03a.err; tail -c 300 gpurun_out/bench_2gpu_r03a.err; cut -c1-200 gpurun_out/bench_2gpu_r03a.json

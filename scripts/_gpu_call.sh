set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "random_clouds" > gpurun_out/pytest_fuzz_r02x.log 2>&1; tail -25 gpurun_out/pytest_fuzz_r02x.log

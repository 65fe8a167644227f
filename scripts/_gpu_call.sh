set -x
timeout 900 python -m pytest tests -m gpu -x -q -k "random_clouds" > gpurun_out/pytest_fuzz_r02w.log 2>&1; tail -12 gpurun_out/pytest_fuzz_r02w.log

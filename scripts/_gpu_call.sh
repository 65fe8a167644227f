set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02j.log 2>&1; tail -8 gpurun_out/pytest_r02j.log

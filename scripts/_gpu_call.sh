set -x
L=gpurun_out/sweep_r02d.log
: > $L
run() { echo "== $*" >> $L; env "$@" python scripts/qbench.py 2e7 20,32 6 >> $L 2>&1; }
run PCT_STAGED_ROUNDS=1
run PCT_STAGED_ROUNDS=2
run PCT_STAGED_ROUNDS=1 PCT_LIST_ROWS=24
run PCT_STAGED_ROUNDS=1 PCT_LIST_ROWS=26
run PCT_STAGED_ROUNDS=1 PCT_LIST_ROWS=30
run PCT_STAGED_ROUNDS=1 PCT_LIST_ROWS=30 PCT_LIST_TARGET=38
run PCT_STAGED_ROUNDS=1 PCT_LIST_ROWS=40
run PCT_STAGED_ROUNDS=1 PCT_CUT_GAIN=3.0
run PCT_STAGED_ROUNDS=1 PCT_CUT_GAIN=3.6
grep -v "^+" $L | cut -c1-230
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02d.log 2>&1; tail -15 gpurun_out/pytest_r02d.log

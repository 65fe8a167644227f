set -x
timeout 900 python scripts/variant_bench.py run 2e7 20,32 8 --parity > gpurun_out/variants_r03b.log 2>&1; cat gpurun_out/variants_r03b.log

set -x
PCT_KNN_KERNEL=warp timeout 900 ncu --set full --import-source on --clock-control none -k regex:knn_warp -c 1 -o gpurun_out/prof_warp_r02k -f python scripts/qbench.py 1e7 20 1 > gpurun_out/ncu_warpfull_r02k.log 2>&1
tail -3 gpurun_out/ncu_warpfull_r02k.log

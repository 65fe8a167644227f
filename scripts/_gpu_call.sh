set -x
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/pcie_probe.py 128 > gpurun_out/pcie_probe_r02n.txt 2> gpurun_out/pcie_probe_r02n.err; cat gpurun_out/pcie_probe_r02n.txt; tail -3 gpurun_out/pcie_probe_r02n.err

"""Aggregate host<->device bandwidth of a box against the number of GPUs copying at the same time.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/pcie_probe.py [MB per rank]

Every rank page-locks its slice of one POSIX shared segment (the layout of distributed.curvature_knn_shared); for
each set of active ranks the members copy their slice H2D (then D2H) at the same time, the others wait at the barrier;
prints per-set aggregate GB/s (bytes of the set / the slowest member's CUDA-event time)."""
import os
import sys

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

from point_cloud_toolbox_b200 import distributed as pdist


def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rows = mb * (1 << 20) // 4
    n = rows * world
    (host,) = pdist.shared_arrays_placed(f"pct_probe_{os.environ.get('MASTER_PORT', '0')}", [(n,)], n)
    mine = host.tensor[rank * rows:(rank + 1) * rows]
    buf = torch.empty(rows, dtype=torch.float32, device=dev)
    sets = [[0], [world - 1], [0, world // 2], [0, 1], list(range(0, world, 2)), list(range(world // 2)), list(range(world))]
    for direction in ("h2d", "d2h"):
        for active in sets:
            best = None
            for rep in range(4):
                dist.barrier()
                torch.cuda.synchronize()
                ms = 0.0
                if rank in active:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    if direction == "h2d":
                        buf.copy_(mine, non_blocking=True)
                    else:
                        mine.copy_(buf, non_blocking=True)
                    e1.record()
                    torch.cuda.synchronize()
                    ms = e0.elapsed_time(e1)
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if rep and (best is None or t.item() < best):
                    best = t.item()
            if rank == 0:
                gb = len(active) * rows * 4 / 1e9
                print(f"{direction} ranks={active} {mb} MB each: slowest {best:.2f} ms, aggregate {gb / (best / 1e3):.1f} GB/s, per rank {gb / len(active) / (best / 1e3):.1f} GB/s", flush=True)
    host.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

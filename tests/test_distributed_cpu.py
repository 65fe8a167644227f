"""World-size-2 gloo tests of the multi-GPU plumbing (no GPU): shard bounds, gather, un-permute."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

from point_cloud_toolbox_b200 import distributed as pdist  # noqa: E402


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 101, 35947, 10 ** 8):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            sizes = []
            for r in range(world):
                b, e = pdist.shard_bounds(n, world, r)
                assert b == prev and e >= b
                prev = e
                sizes.append(e - b)
            assert prev == n and max(sizes) - min(sizes) <= 1
            assert max(sizes) <= pdist.padded_rows(n, world)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank knows the same "index": a permutation (sorted position -> original index)
        g = torch.Generator().manual_seed(5)
        perm = torch.randperm(n, generator=g).to(torch.int32)
        # the cloud exists on rank 0 only and is broadcast
        cloud = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3) if rank == 0 else None
        cloud = pdist.broadcast_cloud(cloud, n, None, 0, torch.device("cpu"))
        assert torch.equal(cloud, torch.arange(n * 3, dtype=torch.float32).reshape(n, 3))
        # "fused kernel" stand-in: row for sorted position s is a function of the ORIGINAL index perm[s]
        b, e = pdist.shard_bounds(n, world, rank)
        orig = perm[b:e].long()
        local = torch.stack((orig.float() * 2.0, cloud[orig, 1]), 1)      # slice-local layout
        gathered = pdist.gather_rows(local, n, None, 0)
        if rank == 0:
            full = pdist.unpermute(gathered, perm)
            want = torch.stack((torch.arange(n).float() * 2.0, cloud[:, 1]), 1)
            np.save(out_path, np.array([float(torch.equal(full, want)), float(gathered.shape[0])]))
        else:
            assert gathered is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 1001])
def test_gather_and_unpermute_world2(tmp_path, n):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    res = np.load(out)
    assert res[0] == 1.0 and res[1] == n

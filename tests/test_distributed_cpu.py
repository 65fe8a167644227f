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


def test_slabs_partition_the_cloud():
    """Cut planes, margins and ownership masks (pure tensor logic of the slab mode)."""
    g = torch.Generator().manual_seed(1)
    for n, world in ((5000, 3), (20000, 8), (300, 2), (50, 8)):
        x = torch.randn(n, generator=g) * 3.0
        x[: n // 10] = x[0]  # a pile of equal coordinates must not break the partition
        cuts = pdist.slab_cuts(x, world)
        assert len(cuts) == world + 1 and cuts[0] == float("-inf") and cuts[-1] == float("inf")
        assert all(a <= b for a, b in zip(cuts, cuts[1:]))
        owner = torch.zeros(n, dtype=torch.int64)
        margin = 0.25
        for r in range(world):
            b = pdist.slab_bounds(cuts, r, margin)
            assert b[0] <= b[2] <= b[3] <= b[1]
            sel, own = pdist.slab_select(x, b)
            own_pos = own.nonzero().squeeze(1)
            owner[sel[own_pos]] += 1
            assert bool((sel[1:] > sel[:-1]).all())  # ascending: ties keep the whole cloud's index order
            # the slab holds everything within the margin of what it owns
            if len(own_pos):
                lo, hi = x[sel[own_pos]].min() - margin, x[sel[own_pos]].max() + margin
                inside = (x >= lo + 1e-6) & (x <= hi - 1e-6)
                held = torch.zeros(n, dtype=torch.bool)
                held[sel] = True
                assert bool(held[inside].all())
        assert bool((owner == 1).all())  # every point is answered by exactly one rank
    sizes = [int(pdist.slab_select(x, pdist.slab_bounds(cuts, r, 0.0))[1].sum()) for r in range(world)]
    assert sum(sizes) == n


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
        # slab mode: ranks hold rows for disjoint, unordered sets of original indices
        x = cloud[:, 0]
        cuts = pdist.slab_cuts(x, world)
        sel, own = pdist.slab_select(x, pdist.slab_bounds(cuts, rank, 4.0))
        ids = sel[own]
        rows = torch.stack((ids.float() * 3.0, cloud[ids, 2]), 1)
        full = pdist.gather_scattered(ids, rows, n, None, 0)
        if rank == 0:
            want = torch.stack((torch.arange(n).float() * 3.0, cloud[:, 2]), 1)
            assert torch.equal(full, want)
        else:
            assert full is None
        # shared-host mode: rows go back to the ranks that own their original index ranges
        mine = pdist.exchange_by_owner(ids, rows, n, None)
        b, e = pdist.shard_bounds(n, world, rank)
        want = torch.stack((torch.arange(b, e).float() * 3.0, cloud[b:e, 2]), 1)
        assert torch.equal(mine, want)
        # ... and the host arrays are one segment mapped by every rank
        name = f"pct_test_{port}"
        if rank == 0:
            shared = pdist.SharedHostArray(name, (2, n), create=True)
            shared.array[:] = 0
        dist.barrier()
        if rank != 0:
            shared = pdist.SharedHostArray(name, (2, n), create=False)
        shared.tensor[0, b:e] = mine[:, 0]
        shared.tensor[1, b:e] = mine[:, 1]
        dist.barrier()
        if rank == 0:
            assert np.array_equal(shared.array[0], np.arange(n, dtype=np.float32) * 3.0)
            assert np.array_equal(shared.array[1], cloud[:, 2].numpy())
        dist.barrier()
        shared.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 1001])
def test_gather_and_unpermute_world2(tmp_path, n):
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    res = np.load(out)
    assert res[0] == 1.0 and res[1] == n


def _text_worker(rank, world, port, path, out):
    import torch.distributed as dist

    from point_cloud_toolbox_b200 import distributed as pdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shared = pdist.shared_cloud_from_text(path, f"pct_text_{port}")
        np.save(out + f".{rank}.npy", np.array(shared.array))
        dist.barrier()
        shared.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("cols", [3, 6])
def test_shared_cloud_from_text_world2(tmp_path, cols):
    """Every rank sees the file's cloud exactly as the reference's loader leaves it (golden loader case)."""
    from conftest import load_golden

    g = load_golden("loader_case")
    path = str(tmp_path / "cloud.txt")
    np.savetxt(path, g["table"][:, :cols], fmt="%.5f")
    out = str(tmp_path / "seen")
    mp.spawn(_text_worker, args=(2, _free_port(), path, out), nprocs=2, join=True)
    for rank in range(2):
        assert np.array_equal(np.load(out + f".{rank}.npy"), g["points"])

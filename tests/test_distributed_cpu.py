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


def test_ranks_are_spread_over_the_visible_gpus():
    assert [pdist.device_for_rank(r, 2, 8) for r in range(2)] == [0, 4]
    assert [pdist.device_for_rank(r, 4, 8) for r in range(4)] == [0, 2, 4, 6]
    assert [pdist.device_for_rank(r, 8, 8) for r in range(8)] == list(range(8))
    assert [pdist.device_for_rank(r, 2, 2) for r in range(2)] == [0, 1]       # exactly as many devices as ranks
    assert [pdist.device_for_rank(r, 3, 8) for r in range(3)] == [0, 2, 4]
    assert pdist.device_for_rank(0, 1, 8) == 0
    assert len({pdist.device_for_rank(r, 4, 5) for r in range(4)}) == 4        # never two ranks on one device


def test_gpu_local_cpus_parses_sysfs_or_gives_up():
    """No GPU here: the topology helpers must return None instead of raising."""
    assert pdist.gpu_local_cpus(0) is None or isinstance(pdist.gpu_local_cpus(0), set)
    assert pdist.bind_near_gpu(0) is None or isinstance(pdist.bind_near_gpu(0), set)


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
        # the host arrays of the end-to-end path are one segment mapped by every rank
        b, e = pdist.shard_bounds(n, world, rank)
        name = f"pct_test_{port}"
        if rank == 0:
            shared = pdist.SharedHostArray(name, (2, n), create=True)
            shared.array[:] = 0
        dist.barrier()
        if rank != 0:
            shared = pdist.SharedHostArray(name, (2, n), create=False)
        shared.tensor[0, b:e] = torch.arange(b, e).float() * 3.0
        shared.tensor[1, b:e] = cloud[b:e, 2]
        dist.barrier()
        if rank == 0:
            assert np.array_equal(shared.array[0], np.arange(n, dtype=np.float32) * 3.0)
            assert np.array_equal(shared.array[1], cloud[:, 2].numpy())
        dist.barrier()
        shared.close()
        # the placed form: every rank first-touches the part it will move, then all page-lock (a no-op without CUDA)
        a_in, a_out = pdist.shared_arrays_placed(f"pct_place_{port}", [(n, 3), (2, n)], n)
        assert a_in.array.shape == (n, 3) and a_out.array.shape == (2, n)
        a_in.array[b:e] = cloud[b:e].numpy()
        a_out.array[:, b:e] = float(rank + 1)
        dist.barrier()
        assert np.array_equal(a_in.array, cloud.numpy())
        for r in range(world):
            rb, re_ = pdist.shard_bounds(n, world, r)
            assert np.all(a_out.array[:, rb:re_] == float(r + 1))
        dist.barrier()
        a_in.close()
        a_out.close()
        # the peer-memory exchanges need NCCL and CUDA: on this backend every rank agrees to use the all-to-alls
        assert pdist.PeerResults.get(n, None, torch.device("cpu")) is None
        pdist.PeerResults.release()
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


# ---------------------------------------------------------------------------
# slab exchange (the cloud is never replicated): protocol, ordering and margins, with CPU stand-ins for the kernels
# ---------------------------------------------------------------------------
def _ref_bin(share, axis, bounds, id_base):
    """What pct_slab_bin_count / pct_slab_bin_fill compute, in torch (test stand-in, no GPU here)."""
    x = share[:, axis]
    recs, complete, owned, owned_local = [], [], [], []
    for c_lo, c_hi, own_lo, own_hi in bounds:
        idx = ((x >= c_lo) & (x <= c_hi)).nonzero().squeeze(1)
        ids = (idx + id_base).to(torch.int32).view(torch.float32)
        recs.append(torch.cat((share[idx, :3], ids[:, None]), 1))
        complete.append(int(idx.numel()))
        own = ((x >= own_lo) & (x < own_hi)).nonzero().squeeze(1)
        owned.append(int(own.numel()))
        owned_local.append(own.to(torch.int32))
    return torch.cat(recs), complete, owned, torch.cat(owned_local)


def _brute_rows(cloud, ids, queries, k):
    """Row per query: [sum of the k nearest other points' ORIGINAL indices, k-th distance]; ties by original index."""
    d = torch.cdist(cloud[queries].double(), cloud.double())
    order = torch.argsort(d * 1e6 + ids[None, :].double() * 1e-6, dim=1)[:, 1:k + 1]      # self first (distance 0)
    return torch.stack((ids[order].sum(1).float(), torch.gather(d, 1, order)[:, -1].float()), 1)


def _exchange_worker(rank, world, port, n, k, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        u, v = torch.rand(n, generator=g) * 6.283, torch.rand(n, generator=g) * 6.283
        cloud = torch.stack(((1 + 0.3 * torch.cos(v)) * torch.cos(u), (1 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)), 1)
        cloud[5] = torch.tensor([0.0, 0.0, 0.0])                                   # isolated, on the cut plane of two slabs: unresolved there
        b, e = pdist.shard_bounds(n, world, rank)
        share = cloud[b:e].contiguous()
        h = 0.08

        def answer(slab, plan, r, kk, n_own):
            c_lo, c_hi, own_lo, own_hi = plan.bounds[r]
            ids = slab[:, 3].contiguous().view(torch.int32).long()
            assert bool((ids[1:] > ids[:-1]).all())                                # ascending original index at the receiver
            xs = slab[:, plan.axis]
            assert bool(((xs >= c_lo) & (xs <= c_hi)).all())
            own = ((xs >= own_lo) & (xs < own_hi)).nonzero().squeeze(1)
            assert own.numel() == n_own
            rows = _brute_rows(slab[:, :3], ids, own, kk)
            rec = torch.zeros((n_own, 8))
            rec[:, 3:5] = rows
            # a query is resolved when its k-th neighbour is closer than the slab's knowledge ends
            reach = torch.minimum(xs[own] - c_lo, c_hi - xs[own])
            bad = rows[:, 1] > reach
            rec[bad, 3:5] = float("nan")
            rec[:, 7] = torch.where(bad, torch.tensor(16, dtype=torch.int32), torch.tensor(0, dtype=torch.int32)).view(torch.float32)
            return rec, None, int(bad.sum())

        def redo(whole, rec, own_ids, plan, kk):
            assert whole.shape[0] == n and torch.equal(whole, cloud)
            bad = rec[:, 7].contiguous().view(torch.int32) != 0
            rec[bad, 3:5] = _brute_rows(whole, torch.arange(n), own_ids[bad].long(), kk)
            return rec

        part = pdist.curvature_knn_exchange(share, b, n, k, bin_fn=_ref_bin, answer_fn=answer, redo_fn=redo,
                                            cell_fn=lambda sample, n_total, bbox, kk: h)
        want = _brute_rows(cloud, torch.arange(n), torch.arange(b, e), k)
        ok = torch.equal(part.rows, want)
        # every rank derived the same plan, slabs own every point exactly once
        owned_total = torch.tensor([int(part.own_ids.numel())])
        dist.all_reduce(owned_total)
        np.save(out_path + f".{rank}.npy", np.array([float(ok), float(owned_total.item()), float(part.unresolved), part.plan.h,
                                                    part.plan.axis] + list(part.plan.cuts[1:-1])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1500), (3, 2000)])
def test_slab_exchange_protocol(tmp_path, world, n):
    out = str(tmp_path / "ex")
    mp.spawn(_exchange_worker, args=(world, _free_port(), n, 12, out), nprocs=world, join=True)
    res = [np.load(out + f".{r}.npy") for r in range(world)]
    for r in res:
        assert r[0] == 1.0, "rows differ from the whole-cloud brute force"
        assert r[1] == n
        assert np.array_equal(r[3:], res[0][3:])           # same plan on every rank
    if world == 2:
        assert sum(r[2] for r in res) >= 1                  # the isolated point went through the whole-cloud redo

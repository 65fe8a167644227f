"""GPU experiment: stage times of one rank's slab work (run on one GPU; ranks are emulated)."""
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from point_cloud_toolbox_b200 import GridIndex, distributed as pdist, engine  # noqa: E402
from scripts.tune import torus  # noqa: E402


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    k = 20
    cloud = torus(n)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        ix = GridIndex(cloud, k_hint=k); ix.curvature_knn(k, want_coeffs=False); torch.cuda.synchronize(); ix.close()
    for rep in range(3):
        for rank in range(min(world, 2)):
            e = [ev()]
            h, lo, hi = engine.estimate_cell_size(cloud, k); e.append(ev())
            axis = max(range(3), key=lambda a: hi[a] - lo[a])
            x = cloud[:, axis]
            bounds = pdist.slab_bounds(pdist.slab_cuts(x, world), rank, pdist.SLAB_MARGIN_CELLS * h); e.append(ev())
            sel, local, row_map, n_own = engine.slab_select(cloud, axis, bounds); e.append(ev())
            e.append(ev())
            index = GridIndex(local, cell_hint=h, k_hint=k); e.append(ev())
            index.set_slab(axis, *bounds, row_map=row_map, mapped_rows=n_own)
            fit = index.curvature_knn(k, want_coeffs=False); e.append(ev())
            own = torch.ones_like(row_map, dtype=torch.bool); own[1:] = row_map[1:] != row_map[:-1]; own[0] = bool(row_map[0] == 0)
            ids = sel[own]; records = fit.records; e.append(ev())
            torch.cuda.synchronize()
            st = index.last_stats()
            names = ["cell", "cuts", "select", "gather", "build", "query", "mask"]
            print(f"rep{rep} rank{rank}/{world} n_local={index.n} owned={int(ids.numel())} " +
                  " ".join(f"{nm}={e[i].elapsed_time(e[i + 1]):.2f}" for i, nm in enumerate(names)) +
                  f" retries={st.level1_retries} unstaged={st.unstaged} exact={st.exact_path}", flush=True)
            index.close()


if __name__ == "__main__":
    main()

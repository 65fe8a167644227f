"""CPU oracle for the curvature hot path of masnottuh/point-cloud-toolbox.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as something the GPU path calls.

Parity status
-------------
* kNN -> PCA plane -> rotation -> quadric fit -> curvature: **pinned** against
  outputs of the unmodified reference (``/root/reference/pointCloudToolbox.py``
  imported with four stubbed plotting/mesh modules) on ``bunny.txt``,
  ``egg_carton.txt`` and the torus stand-in; the vectors and the script that
  made them live in ``tests/golden/`` (``oracle/make_golden.py``).
  The reference has no tests or golden vectors of its own.
* neighbour study (ref :732-800): **pinned** -- return values of the unmodified reference's
  ``explicit_quadratic_neighbor_study`` on ``bunny.txt`` for seeded samples and several tolerances
  (``tests/golden/neighbor_study.npz``).
* rows either side of the path (text / PLY I/O, mesh energies, PCA estimators; ``oracle/around_path.py``):
  **pinned** -- ``oracle/make_golden_io.py`` compiles the unmodified function definitions out of
  ``/root/reference/utils.py`` (the file itself needs open3d / pyvista to import) and stores their outputs in
  ``tests/golden/io_energy_pca.npz``; ``estimate_curvature`` is the exception (the reference's output is rounding
  noise, see its docstring in the product), **parity unpinned** for it.
* epsilon-ball query: **parity unpinned** -- the reference never implemented it
  (README.md:8 advertises it, pointCloudToolbox.py:101-102 only lists scipy's
  API).  The oracle composes ``scipy.spatial.cKDTree.query_ball_point`` with the
  reference's own per-neighbourhood functions.

The arithmetic the reference relies on lives in third-party packages that are
not vendored and not pinned by the reference (requirements.txt lists neither):
``scipy.spatial.cKDTree`` (here scipy 1.18.1) and ``numpy.cov / linalg.svd /
linalg.lstsq`` (here numpy 2.3.5).  Their published behaviour is restated in
``oracle/reference_path.py`` next to each call.
"""
from .reference_path import (  # noqa: F401
    load_points,
    squared_distance_key,
    knn_canonical,
    ball_canonical,
    best_fit_plane_and_rotate,
    fit_quadratic_surface,
    explicit_quadratic_curvatures,
    neighbourhood_pipeline,
    curvature_from_neighbors,
    curvature_from_neighbors_batched,
    curvature_from_csr,
    knn_curvature,
    neighbor_study,
)
from . import datasets, compare  # noqa: F401

"""B200-native curvature hot path of masnottuh/point-cloud-toolbox.

    from point_cloud_toolbox_b200 import PointCloud      # drop-in for pointCloudToolbox.PointCloud

kNN / epsilon-ball search on a Morton-sorted uniform grid, PCA tangent plane,
oriented rotation, least-squares quadric and Gaussian / mean / principal
curvature, all in hand-written CUDA for sm_100a behind a C ABI
(include/pct_b200.h, libpct_b200.so).  Importing this package loads that
library and fails if it is missing; nothing here computes on the CPU.
"""
from . import _lib  # noqa: F401  (loads libpct_b200.so; raises if it cannot)
from .engine import GridIndex, fit_from_neighbors, fit_from_csr, plane_rotate, quadric_fit, quadric_curvature  # noqa: F401
from .pointCloudToolbox import PointCloud  # noqa: F401

__all__ = ["PointCloud", "GridIndex", "fit_from_neighbors", "fit_from_csr", "plane_rotate", "quadric_fit",
           "quadric_curvature"]
__version__ = "0.1.0"

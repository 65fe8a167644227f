"""Build tests/host_harness/libharness.so (CPU build of the __host__ __device__ query / fit code)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libharness.so")


def build():
    src = os.path.join(HERE, "harness.cpp")
    inc = os.path.join(ROOT, "point_cloud_toolbox_b200", "csrc")
    deps = [src] + [os.path.join(inc, f) for f in ("pct_math.cuh", "pct_grid.cuh", "pct_dispatch.h", "pct_energy.cuh")]
    extra = os.environ.get("PCT_HARNESS_EXTRA", "").split()  # e.g. -DPCT_HIST_BINS=32: the logic tests against a knob setting
    lib = LIB if not extra else LIB.replace(".so", "_variant.so")
    if not extra and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps):
        return lib
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", *extra, "-I", inc, src, "-o", lib], check=True)
    return lib


if __name__ == "__main__":
    print(build())

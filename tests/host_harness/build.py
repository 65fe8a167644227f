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
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", inc, src, "-o", LIB], check=True)
    return LIB


if __name__ == "__main__":
    print(build())

"""Compare compile-time variants of the library on one GPU box, one index, one command (results of round 2: profiles/variants_r02a.txt).

Here (no GPU needed):   python scripts/variant_bench.py build red=-DPCT_HIST_RED=1 cull=-DPCT_CULL_PASS2=1 ...
    builds point_cloud_toolbox_b200/build/variants/<name>.so for every NAME=FLAGS (FLAGS: nvcc flags joined by ','),
    plus `base.so` with no extra flag; the .so files travel to the box with the repo snapshot.
On the box:             python scripts/variant_bench.py run [N] [k,k,..] [reps] [--parity]
    for every variant: the library is swapped in, scripts/qbench.py runs in a fresh process, and with --parity
    the kNN / fused-curvature parity tests run against it too; prints one line per variant and k; restores the
    library afterwards.
"""
import glob
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "point_cloud_toolbox_b200")
LIB = os.path.join(PKG, "libpct_b200.so")
VARIANTS = os.path.join(PKG, "build", "variants")


def build(specs):
    os.makedirs(VARIANTS, exist_ok=True)
    keep = os.path.join(VARIANTS, "_shipping.so")
    stamp = os.path.join(PKG, "build", "stamp")
    shutil.copy(LIB, keep)
    try:
        for name, flags in [("base", "")] + specs:
            env = dict(os.environ, PCT_NVCC_EXTRA=flags.replace(",", " "))
            subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from point_cloud_toolbox_b200 import build as b; "
                            "b.build(force=True)" % ROOT], env=env, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
            shutil.copy(LIB, os.path.join(VARIANTS, name + ".so"))
            print("built", name, flags)
    finally:
        shutil.copy(keep, LIB)
        os.remove(keep)
        # the object files now belong to the last variant: drop the stamp so that the next import / build() recompiles
        # the shipping configuration (scripts/sass_hash.py reads the objects)
        if os.path.exists(stamp):
            os.remove(stamp)


def run(argv):
    parity = "--parity" in argv
    argv = [a for a in argv if a != "--parity"]
    n = argv[0] if argv else "2e7"
    ks = argv[1] if len(argv) > 1 else "20"
    reps = argv[2] if len(argv) > 2 else "8"
    keep = LIB + ".keep"
    shutil.copy(LIB, keep)
    try:
        for so in sorted(glob.glob(os.path.join(VARIANTS, "*.so"))):
            name = os.path.basename(so)[:-3]
            shutil.copy(so, LIB)
            out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "qbench.py"), n, ks, reps], cwd=ROOT,
                                 capture_output=True, text=True)
            for line in out.stdout.splitlines():
                m = re.search(r"k=(\d+).*med=([0-9.]+)ms.*retries=(\d+) exact=(\d+) unstaged=(\d+)", line)
                if m:
                    print(f"{name:16s} k={m.group(1):>3s} med={m.group(2):>7s} ms  retries={m.group(3)} exact={m.group(4)} unstaged={m.group(5)}", flush=True)
            if out.returncode != 0:
                print(f"{name:16s} qbench failed: {out.stderr[-300:]}", flush=True)
            if parity:
                t = subprocess.run([sys.executable, "-m", "pytest", "tests", "-m", "gpu", "-x", "-q", "-k",
                                    "knn_lists or fused_curvature or exact_path or tiny or large_cloud"], cwd=ROOT,
                                   capture_output=True, text=True)
                print(f"{name:16s} parity: {t.stdout.strip().splitlines()[-1] if t.stdout.strip() else t.stderr[-200:]}", flush=True)
    finally:
        shutil.copy(keep, LIB)
        os.remove(keep)


if __name__ == "__main__":
    if len(sys.argv) >= 2 and sys.argv[1] == "build":
        build([tuple(a.split("=", 1)) for a in sys.argv[2:]])
    elif len(sys.argv) >= 2 and sys.argv[1] == "run":
        run(sys.argv[2:])
    else:
        print(__doc__)

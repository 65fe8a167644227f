"""Fingerprint of the device code of the current build: one SHA-256 per kernel of its SASS text (encodings stripped).

    python scripts/sass_hash.py > profiles/sass_hashes_rNN.json          # after a GPU-verified build
    python scripts/sass_hash.py profiles/sass_hashes_rNN.json            # later: which kernels changed?

A change that is meant to be compiled out (a knob that defaults to off) must leave every hash as it was; this is how
device code can be touched when no GPU time is left to re-verify it.  Kernel names are compared without their
template arguments' defaults (an added, defaulted template parameter changes the mangled name only).
"""
import glob
import hashlib
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hashes():
    out = {}
    for obj in sorted(glob.glob(os.path.join(ROOT, "point_cloud_toolbox_b200", "build", "*.o"))):
        text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
        name, body = None, []

        def flush():
            if name is not None:
                out[os.path.basename(obj) + ":" + name] = hashlib.sha256("".join(body).encode()).hexdigest()[:16]

        for line in text.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                flush()
                name, body = m.group(1), []
            elif name is not None and "/*" in line:
                line = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip()
                if line.strip():
                    body.append(line.strip() + "\n")
        flush()
    return out


def demangled_key(k):
    try:
        obj, name = k.split(":", 1)
        d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        return obj + ":" + re.sub(r"\(.*", "", d)
    except Exception:
        return k


if __name__ == "__main__":
    now = hashes()
    if len(sys.argv) < 2:
        print(json.dumps(now, indent=0, sort_keys=True))
        sys.exit(0)
    ref = json.load(open(sys.argv[1]))
    ref_by_hash = {}
    for k, v in ref.items():
        ref_by_hash.setdefault(v, []).append(k)
    changed = [k for k, v in now.items() if v not in ref_by_hash]
    gone = [k for k, v in ref.items() if v not in set(now.values())]
    print(f"{len(now)} kernels now, {len(ref)} in the reference; {len(changed)} without a match:")
    for k in changed:
        print("  changed / new:", demangled_key(k)[:150])
    for k in gone:
        print("  no longer present:", demangled_key(k)[:150])
    sys.exit(1 if changed else 0)

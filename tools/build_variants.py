#!/usr/bin/env python
"""Build tuning variants of the library next to the default one (kernel experiments).

    python tools/build_variants.py name1=-DFLAG_A,-DFLAG_B name2=-DFLAG_C ...

Each variant recompiles every translation unit with its flags into ethicaltrajectoryplanning_b200/build/variants/<name>/ and
links libetp_<name>.so there; select one at run time with ETP_B200_LIB=<path>.  Variants travel to the GPU box with the
snapshot (they are git-ignored like the default library)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ethicaltrajectoryplanning_b200 import _abi  # noqa: E402


def build(name, flags):
    out_dir = os.path.join(_abi.HERE, "build", "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    objs = []
    for src in _abi.SOURCES:
        obj = os.path.join(out_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = ["nvcc", *_abi.NVCC_FLAGS, *flags, "-c", "-o", obj, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise SystemExit(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    lib = os.path.join(out_dir, f"libetp_{name}.so")
    r = subprocess.run(["nvcc", *_abi.NVCC_FLAGS, "-shared", "-o", lib, *objs], capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stdout + r.stderr)
    return lib


def main():
    jobs = []
    for arg in sys.argv[1:]:
        name, _, flags = arg.partition("=")
        jobs.append((name, [f for f in flags.split(",") if f]))
    with ThreadPoolExecutor(max_workers=4) as ex:
        for lib in ex.map(lambda j: build(*j), jobs):
            print(lib)


if __name__ == "__main__":
    main()

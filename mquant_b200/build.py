"""Build libmquant_b200.so in-tree with nvcc for sm_100a (no JIT cache, the .so travels with the repo)."""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmquant_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmquant_b200.so")
    return nvcc


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, probes: bool = False, stamps: bool = False,
          defines: tuple = (), out: str | None = None) -> str:
    """probes=True defines MQ_PROBES: the bottleneck probes of tools/gemm_probe.py / had_probe.py
    (MQ_GEMM_DEBUG / MQ_HAD_DEBUG masks) are compiled in -- never for a shipped library."""
    # (defines / out: an experimental variant next to the product library, for same-box A/B runs with
    # MQUANT_B200_LIB=<out>; objects go to their own directory)
    lib = out or LIB
    obj_dir = OBJ_DIR if out is None else OBJ_DIR + "_" + os.path.basename(out).replace(".", "_")
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(os.path.dirname(HERE), "include", "mquant_b200.h")]
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    objs = []
    for s in srcs:
        o = os.path.join(obj_dir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or probes or stamps or _stale(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-DMQ_PROBES"] if probes else []) + (["-DMQ_STAMPS"] if stamps else []) + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    if jobs or force or _stale(lib, objs):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libmquant_b200.so failed")
    return lib


if __name__ == "__main__":
    _defs = tuple(a[2:] for a in sys.argv[1:] if a.startswith("-D"))
    _out = next((a[6:] for a in sys.argv[1:] if a.startswith("--out=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, probes="--probes" in sys.argv, stamps="--stamps" in sys.argv,
                defines=_defs, out=_out))

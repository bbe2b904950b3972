"""Build recipe for libmpassit_rg.so (CUDA engine + C ABI) -- nvcc, sm_100a only.

Objects are compiled in parallel and linked in-tree (mpassit_b200/libmpassit_rg.so)
so the library travels with the repository snapshot to the GPU box.  Geometry /
weight-generation units are compiled with -fmad=false: their index and mask
decisions must be reproducible bit-for-bit by a plain IEEE fp64 host restatement.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmpassit_rg.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr",
          "-Xptxas", "-v"]
# (source, extra flags, macros it provides)
UNITS = [
    ("capi.cu", [], []),
    ("mesh.cu", ["-fmad=false"], []),
    ("bvh.cu", ["-fmad=false"], []),
    ("locate.cu", ["-fmad=false"], []),
    ("conserve.cu", ["-fmad=false"], ["MPRG_HAVE_CONSERVE"]),
    ("stagger.cu", ["-fmad=false"], ["MPRG_HAVE_STAGGER", "MPRG_HAVE_NODE"]),
    ("compose.cu", ["-fmad=false"], []),
    ("apply.cu", [], []),
    ("wcache.cu", [], []),
    ("target_gen.cu", ["-fmad=false"], []),
    ("gather.cu", [], []),
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "mpassit_rg.h"))
    headers.append(os.path.abspath(__file__))
    units = [(s, fl, mac) for (s, fl, mac) in UNITS if os.path.exists(os.path.join(CSRC, s))]
    # the library travels to the GPU box without its object files (mpassit_b200/_build is not shipped):
    # a library newer than every source is up to date, whatever is in _build
    if not force and not _stale(LIB, [os.path.join(CSRC, s) for (s, _, _) in units] + headers):
        return LIB
    macros = [f"-D{m}" for (_, _, mac) in units for m in mac]
    nvcc = _nvcc()
    jobs = []
    for src, flags, _ in units:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            jobs.append((src, [nvcc, *ARCH, *COMMON, *flags, *macros, "-c", s, "-o", o]))

    def run(job):
        name, cmd = job
        p = subprocess.run(cmd, capture_output=True, text=True)
        return name, p.returncode, p.stdout + p.stderr

    logs = []
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for name, rc, out in ex.map(run, jobs):
            logs.append((name, out))
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError(f"nvcc failed on {name}")
    with open(os.path.join(OBJ, "ptxas.log"), "a") as fh:
        for name, out in logs:
            fh.write(f"==== {name}\n{out}\n")
            if verbose:
                print(f"==== {name}\n{out}")
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for (s, _, _) in units]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB


HOST_LIB = os.path.join(HERE, "libmpassit_host.so")
HOST_SRC = ["setup.cpp", "target_grid.cpp", "interp.cpp", "ncio.cpp", "run.cpp", "weights.cpp"]


def build_host(force: bool = False) -> str:
    """libmpassit_host.so: C++ mirror of the reference's Fortran host stages (no CUDA code;
    calls the engine through its C ABI, resolved from the same directory via rpath)."""
    srcs = [os.path.join(HERE, "host", s) for s in HOST_SRC]
    deps = srcs + [os.path.join(HERE, "host", "ncio.hpp"), os.path.join(HERE, "host", "par.hpp"), os.path.join(HERE, "..", "include", "mpassit_host.h"), os.path.join(HERE, "..", "include", "mpassit_rg.h"), LIB]
    if force or _stale(HOST_LIB, deps):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wall", "-o", HOST_LIB, *srcs,
               "-L" + HERE, "-lmpassit_rg", "-Wl,-rpath,$ORIGIN"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("host library build failed")
    return HOST_LIB


def build_all(force: bool = False, verbose: bool = False):
    return build(force, verbose), build_host(force)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))

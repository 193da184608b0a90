"""In-tree build of the native pieces (no JIT cache: the built files travel with the repo snapshot).

  emsar_b200/libemsar_cuda.so   CUDA kernels + C ABI (include/emsar_cuda.h), sm_100a only
  emsar_b200/libemsar_host.so   host-side C: rsh loader, alignment readers, writers (no CUDA dependency)
  emsar_b200/bin/emsar          the command-line program (host C linked against both)
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "emsar_b200")
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
OBJ = os.path.join(PKG, "csrc", "_obj")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built")
    return p


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
    return r.stdout + r.stderr


def build_cuda(force=False, verbose=False, extra=()):
    out = os.path.join(PKG, "libemsar_cuda.so")
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        if force or _newer(o, [s] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + list(extra) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
    logs = []
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            logs = list(ex.map(_run, jobs))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in srcs]
    if force or jobs or _newer(out, objs):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs)
    if verbose:
        print("\n".join(logs))
    return out


def build_host(force=False):
    """Host C library + CLI. Built only when the sources exist (they arrive after the CUDA path)."""
    srcs = sorted(glob.glob(os.path.join(HOST, "*.c")))
    outs = []
    lib_srcs = [s for s in srcs if not s.endswith("_main.c")]
    hdrs = glob.glob(os.path.join(HOST, "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    if lib_srcs:
        out = os.path.join(PKG, "libemsar_host.so")
        if force or _newer(out, lib_srcs + hdrs):
            _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", HOST,
                  "-o", out] + lib_srcs + ["-lz", "-lm", "-lpthread"])
        outs.append(out)
    bmain = os.path.join(HOST, "emsar_build_main.c")
    if os.path.exists(bmain):          # no link-time CUDA dependency: --device loads libemsar_cuda.so at run time
        os.makedirs(os.path.join(PKG, "bin"), exist_ok=True)
        out = os.path.join(PKG, "bin", "emsar-build")
        if force or _newer(out, srcs + hdrs):
            _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", out, bmain] + lib_srcs +
                 ["-lz", "-lm", "-lpthread", "-ldl"])
        outs.append(out)
    main = os.path.join(HOST, "emsar_main.c")
    if os.path.exists(main):
        os.makedirs(os.path.join(PKG, "bin"), exist_ok=True)
        out = os.path.join(PKG, "bin", "emsar")
        cuda_so = os.path.join(PKG, "libemsar_cuda.so")
        if force or _newer(out, srcs + hdrs + [cuda_so]):
            _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", HOST, "-o", out, main] + lib_srcs +
                 ["-L", PKG, "-lemsar_cuda", "-Wl,-rpath,$ORIGIN/..", "-lz", "-lm", "-lpthread", "-lstdc++"])
        outs.append(out)
    return outs


def build_tools(force=False):
    """Bench / test tooling: the fast synthetic-workload generator (no CUDA, not on the product path)."""
    src = os.path.join(PKG, "tools", "synth_gen.c")
    out = os.path.join(PKG, "libemsar_synth.so")
    if force or _newer(out, [src]):
        _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-fopenmp", "-fPIC", "-shared", "-o", out, src, "-lm"])
    return [out]


def build_all(force=False, verbose=False):
    outs = [build_cuda(force=force, verbose=verbose)]
    outs += build_host(force=force)
    outs += build_tools(force=force)
    return outs


if __name__ == "__main__":
    print("\n".join(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)))

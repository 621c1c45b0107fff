"""Builds libkvae_kalman.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

One translation unit per shape of kvae_configs.h (compiled in parallel), plus the C-ABI unit.
No torch headers are involved: the library's only dependency is the CUDA runtime.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development A/B builds: KVAE_NVCC_FLAGS="-DKV_TPB_LARGE=128" KVAE_LIB_OUT=/path/lib_variant.so (load it with KVAE_LIB=...)
_VARIANT = os.environ.get("KVAE_LIB_OUT")
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_variant" if _VARIANT else ""))
LIB = _VARIANT or os.path.join(HERE, "libkvae_kalman.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("KVAE_NVCC_FLAGS", "").split()


def shapes():
    env = os.environ.get("KVAE_SHAPES")   # development: "4,2,4,3;16,8,16,8" builds only those shapes
    if env:
        return [tuple(int(x) for x in part.split(",")) for part in env.split(";") if part]
    txt = open(os.path.join(CSRC, "kvae_configs.h")).read()
    body = txt[txt.index("#define KVAE_FOR_EACH_SHAPE"):]
    return [tuple(int(x) for x in m) for m in re.findall(r"X\((\d+),\s*(\d+),\s*(\d+),\s*(\d+)\)", body)]


def _sources():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))] + \
           [os.path.join(HERE, "..", "include", "kvae_kalman.h")]


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources())


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed ({' '.join(cmd)}):\n{p.stdout[-4000:]}")
    return p.stdout


def _compile(shape_list, obj_dir, lib_path, override_list, verbose=False):
    os.makedirs(obj_dir, exist_ok=True)
    jobs = []
    for (n, p, m, k) in shape_list:
        o = os.path.join(obj_dir, f"shape_{n}_{p}_{m}_{k}.o")
        cmd = [NVCC, *ARCH, *FLAGS, f"-DKV_N={n}", f"-DKV_P={p}", f"-DKV_M={m}", f"-DKV_K={k}",
               "-c", os.path.join(CSRC, "kvae_shape.cu"), "-o", o]
        jobs.append((cmd, o))
    o = os.path.join(obj_dir, "capi.o")
    extra = []
    if override_list:
        hdr = os.path.join(obj_dir, "shape_list_override.h")
        with open(hdr, "w") as f:
            f.write("#define KVAE_FOR_EACH_SHAPE(X) " + " ".join(f"X({n},{p},{m},{k})" for (n, p, m, k) in shape_list) + "\n")
        extra = ["--pre-include", hdr]
    jobs.append(([NVCC, *ARCH, *FLAGS, *extra, "-c", os.path.join(CSRC, "kvae_capi.cu"), "-o", o], o))
    for name in ("regime", "dp", "vae"):
        o = os.path.join(obj_dir, name + ".o")
        jobs.append(([NVCC, *ARCH, *FLAGS, "-c", os.path.join(CSRC, f"kvae_{name}.cu"), "-o", o], o))
    # incremental: an object is kept when it is newer than its own .cu, every header and the recorded command line
    headers = [s for s in _sources() if not s.endswith(".cu")]

    def fresh(job):
        cmd, obj = job
        src = [c for c in cmd if c.endswith(".cu")]
        cmdfile = obj + ".cmd"
        if not (os.path.exists(obj) and os.path.exists(cmdfile) and open(cmdfile).read() == " ".join(cmd)):
            return False
        t = os.path.getmtime(obj)
        return all(os.path.getmtime(d) <= t for d in headers + src)

    def compile_one(job):
        if fresh(job):
            return ""
        out = _run(job[0], job[1] + ".log")
        with open(job[1] + ".cmd", "w") as f:
            f.write(" ".join(job[0]))
        return out

    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        outs = list(ex.map(compile_one, jobs))
    if verbose:
        for out in outs:
            sys.stdout.write(out)
    tmp = lib_path + f".tmp{os.getpid()}"
    _run([NVCC, *ARCH, "-shared", "-o", tmp, *[j[1] for j in jobs]], os.path.join(obj_dir, "link.log"))
    os.replace(tmp, lib_path)   # atomic: a concurrent loader never sees a half-written library
    return lib_path


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    return _compile(shapes(), OBJ, LIB, bool(os.environ.get("KVAE_SHAPES")), verbose)


# ---------------------------------------------------------------------------------------------------------
# Shapes on demand.  KVAEConfig (kvae/utils/config.py:4-60) allows any (a_dim, z_dim, u_dim, num_modes); the default
# library instantiates the five tuples of kvae_configs.h.  For any other tuple capi.lib_for() calls build_shape_lib(),
# which compiles the SAME sources for that one tuple into kalman_vae_b200/_jit/ (nvcc is part of the image; ~0.5-2 min,
# once -- the library is cached on disk and re-used while the sources are unchanged).
# ---------------------------------------------------------------------------------------------------------
JIT_DIR = os.path.join(HERE, "_jit")


def shape_lib_path(n, p, m, k):
    return os.path.join(JIT_DIR, f"libkvae_kalman_{n}_{p}_{m}_{k}.so")


def build_shape_lib(n, p, m, k, force=False, verbose=False):
    lib = shape_lib_path(n, p, m, k)
    if not force and os.path.exists(lib) and all(os.path.getmtime(s) <= os.path.getmtime(lib) for s in _sources()):
        return lib
    if not os.path.exists(NVCC):
        raise RuntimeError(f"shape (n={n}, p={p}, m={m}, K={k}) is not in libkvae_kalman.so and {NVCC} is not available to build it")
    os.makedirs(JIT_DIR, exist_ok=True)
    import fcntl
    with open(os.path.join(JIT_DIR, f".lock_{n}_{p}_{m}_{k}"), "w") as lk:   # ranks of one job build it once
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if force or not os.path.exists(lib) or any(os.path.getmtime(s) > os.path.getmtime(lib) for s in _sources()):
                sys.stderr.write(f"[kalman_vae_b200] compiling the kernels for shape (n={n}, p={p}, m={m}, K={k}) -> {lib}\n")
                _compile([(n, p, m, k)], os.path.join(JIT_DIR, f"_obj_{n}_{p}_{m}_{k}"), lib, True, verbose)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

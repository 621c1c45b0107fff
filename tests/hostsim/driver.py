"""TEST TOOLING: ctypes driver of tests/hostsim/libhostsim.so (host build of the device headers)."""
import ctypes, os, subprocess
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libhostsim.so")
CSRC = os.path.normpath(os.path.join(HERE, "..", "..", "kalman_vae_b200", "csrc"))


def build(force=False):
    src = os.path.join(HERE, "hostsim.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if force or not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", SO, src])
    return SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _f(case, k):
    return case[k].to(torch.float32).contiguous()


def fwd(case, smooth=True, force_mem=False):
    Y, U, mask, alpha = _f(case, "Y"), _f(case, "U"), _f(case, "mask"), _f(case, "alpha")
    A, Bm, C, Q, R, mu0, S0 = (_f(case, k) for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0"))
    B, T, p = Y.shape
    K, n, m = Bm.shape
    sw = int(bool(case["q_per_mode"]))
    assert bool(case["q_per_mode"]) == bool(case["c_shared"])
    z = lambda *s: torch.zeros(*s, dtype=torch.float32)
    o = dict(mus_filt=z(B, T, n, 1), Sigmas_filt=z(B, T, n, n), mus_pred=z(B, T, n, 1), Sigmas_pred=z(B, T, n, n),
             A_list=z(B, T, n, n), B_list=z(B, T, n, m), C_list=z(B, T, p, n),
             mus_smooth=z(B, T, n, 1), Sigmas_smooth=z(B, T, n, n))
    info = torch.zeros(1, dtype=torch.int32)
    rc = lib().hostsim_fwd(n, p, m, K, sw, int(force_mem), int(smooth), B, T, _p(Y), _p(U), _p(mask), _p(alpha),
                           _p(A), _p(Bm), _p(C), _p(Q), _p(R), _p(mu0), _p(S0),
                           _p(o["mus_filt"]), _p(o["Sigmas_filt"]), _p(o["mus_pred"]), _p(o["Sigmas_pred"]),
                           _p(o["A_list"]), _p(o["B_list"]), _p(o["C_list"]), _p(o["mus_smooth"]), _p(o["Sigmas_smooth"]),
                           _p(info))
    assert rc == 0, rc
    o["info"] = info
    return o

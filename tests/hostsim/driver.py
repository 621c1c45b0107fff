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


def elbo(case, states, force_mem=False, jitter=1e-6):
    Y, U, mask, alpha, eps = (_f(case, k) for k in ("Y", "U", "mask", "alpha", "eps"))
    A, Bm, C, Q, R, mu0, S0 = (_f(case, k) for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0"))
    B, T, p = Y.shape
    K, n, m = Bm.shape
    sw = int(bool(case["q_per_mode"]))
    ms = states["mus_smooth"].to(torch.float32).contiguous()
    Ss = states["Sigmas_smooth"].to(torch.float32).contiguous()
    acc = (ctypes.c_double * 5)()
    info = torch.zeros(1, dtype=torch.int32)
    rc = lib().hostsim_elbo(n, p, m, K, sw, int(force_mem), B, T, _p(Y), _p(U), _p(mask), _p(alpha), _p(eps),
                            _p(A), _p(Bm), _p(C), _p(Q), _p(R), _p(mu0), _p(S0), _p(ms), _p(Ss),
                            ctypes.c_float(jitter), acc, _p(info))
    assert rc == 0
    a = list(acc)
    return dict(trans=a[0], emiss=a[1], init=a[2], entropy=a[3], num_el=max(a[4], 1.0),
                elbo=(a[0] + a[1] + a[2] + a[3]) / max(a[4], 1.0), info=int(info))


COT_NAMES = ["mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred",
             "A_list", "B_list", "C_list"]


def bwd(case, states, g_elbo=1.0, cot=None, force_mem=False, jitter=1e-6, with_elbo=False):
    """states: the six state tensors (fp32). Returns dY,dU,dalpha,dA,dB,dC[,dQ]."""
    Y, U, mask, alpha, eps = (_f(case, k) for k in ("Y", "U", "mask", "alpha", "eps"))
    A, Bm, C, Q, R, mu0, S0 = (_f(case, k) for k in ("A", "B", "C", "Q", "R", "mu0", "Sigma0"))
    B, T, p = Y.shape
    K, n, m = Bm.shape
    sw = int(bool(case["q_per_mode"]))
    st = {k: states[k].to(torch.float32).contiguous() for k in COT_NAMES[:6]}
    c_elbo = float(g_elbo) / max(float(mask.sum()), 1.0)
    cots = [(cot[k].to(torch.float32).contiguous() if cot and cot.get(k) is not None else None) for k in COT_NAMES]
    cot_arr = (ctypes.c_void_p * 9)(*[(c.data_ptr() if c is not None else None) for c in cots])
    dY, dU, dal = torch.zeros(B, T, p), torch.zeros(B, T, m), torch.zeros(B, T, K)
    psz = K * n * n + K * n * m + K * p * n + (K * n * n if sw else 0)
    gp = (ctypes.c_double * psz)()
    info = torch.zeros(1, dtype=torch.int32)
    dbg = [torch.zeros(B, T, n, n), torch.zeros(B, T, n, n), torch.zeros(B, T, n), torch.zeros(B, T, n)]
    dbg_arr = (ctypes.c_void_p * 4)(*[d.data_ptr() for d in dbg])
    el5 = (ctypes.c_double * 5)()
    rc = lib().hostsim_bwd(n, p, m, K, sw, int(force_mem), B, T, _p(Y), _p(U), _p(mask), _p(alpha), _p(eps),
                           _p(A), _p(Bm), _p(C), _p(Q), _p(R), _p(mu0), _p(S0),
                           _p(st["mus_filt"]), _p(st["Sigmas_filt"]), _p(st["mus_pred"]), _p(st["Sigmas_pred"]),
                           _p(st["mus_smooth"]), _p(st["Sigmas_smooth"]),
                           ctypes.c_float(c_elbo), ctypes.c_float(jitter), cot_arr,
                           _p(dY), _p(dU), _p(dal), gp, _p(info), dbg_arr, el5 if with_elbo else None)
    assert rc == 0
    flat = torch.tensor(list(gp), dtype=torch.float64)
    o = 0
    out = dict(dY=dY, dU=dU, dalpha=dal, info=int(info))
    if with_elbo:   # the ELBO value accumulated by the adjoint sweep itself (fused value + adjoint mode)
        e = list(el5)
        out["elbo_terms"] = e
        out["elbo"] = (e[0] + e[1] + e[2] + e[3]) / max(e[4], 1.0)
    out['_dbg'] = dict(Sf_b=dbg[0], Sp_b=dbg[1], mf_b=dbg[2], mp_b=dbg[3])
    for name, shp in (("dA", (K, n, n)), ("dB", (K, n, m)), ("dC", (K, p, n))) + ((("dQ", (K, n, n)),) if sw else ()):
        sz = shp[0] * shp[1] * shp[2]
        out[name] = flat[o:o + sz].view(*shp).clone()
        o += sz
    return out

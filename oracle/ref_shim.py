"""TEST INFRASTRUCTURE — live import of the UNMODIFIED reference (rodrigo-paganini/kalman-vae).

Usable where the reference sources exist: the checkout of the build container (/root/reference) or the
pip-installed copy under baseline/_ref (oracle/install_reference.sh; git-ignored, it travels to the GPU box).
Everything the GPU parity tests need from the reference is ALSO exported as golden vectors by
`oracle/make_golden.py` into `tests/golden/`, so they do not depend on either.

Two import shims are required:
  * kvae/kalman/kalman_filter.py:5 imports matplotlib (unused, not installed)
  * kvae/vae/losses.py:3 imports the non-existent module kvae.vae.config

Nothing under `kalman_vae_b200/` may import this module.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the read-only checkout of the build container, then the pip-installed copy that travels to the GPU box
# (oracle/install_reference.sh -> baseline/_ref)
REFERENCE_ROOTS = ("/root/reference", os.path.join(_REPO, "baseline", "_ref"))


def reference_root():
    for r in REFERENCE_ROOTS:
        if os.path.isdir(os.path.join(r, "kvae", "kalman")):
            return r
    return None


def available() -> bool:
    return reference_root() is not None


_loaded = {}


def load():
    """Returns a namespace with the reference classes (KalmanFilter, DynamicsParameter, ...)."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (expected /root/reference or baseline/_ref)")
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if root not in sys.path:
        sys.path.insert(0, root)
    import kvae.utils.config as _cfg  # noqa

    sys.modules.setdefault("kvae.vae.config", _cfg)
    from kvae.kalman.kalman_filter import KalmanFilter
    from kvae.kalman.dyn_param import DynamicsParameter
    from kvae.kalman.switch_dyn_param import SwitchingDynamicsParameter
    import kvae.kalman.switch_dyn_param as switch_mod
    from kvae.model.model import KVAE
    from kvae.utils.config import KVAEConfig

    _loaded.update(
        KalmanFilter=KalmanFilter,
        DynamicsParameter=DynamicsParameter,
        SwitchingDynamicsParameter=SwitchingDynamicsParameter,
        switch_mod=switch_mod,
        KVAE=KVAE,
        KVAEConfig=KVAEConfig,
    )
    return types.SimpleNamespace(**_loaded)


class FixedAlphaDynamics(nn.Module):
    """Drives the reference KalmanFilter with externally supplied mixture weights.

    Implements exactly the part of the dyn_params protocol that KalmanFilter touches
    (kalman_filter.py:13-15,135-139,189,343,382-383): `is_switching_dynamics`, `A/B/C/Q`,
    `reset_state`, `compute_batch`, `Q_seq`, `elbo_terms`.  The mixing is done with the same
    einsum expressions as dyn_param.py:58-60 / switch_dyn_param.py:82-86.

      c_shared   : C_t = C[0] for all t (switching variant) instead of sum_k alpha_k C_k
      q_per_mode : Q_t = sum_k alpha_k Q_k (switching) instead of the fixed kf.Q (lstm)
    """

    def __init__(self, A, B, C, Q, alpha, c_shared: bool, q_per_mode: bool):
        super().__init__()
        self.is_switching_dynamics = True
        self.A = nn.Parameter(A.clone())
        self.B = nn.Parameter(B.clone())
        self.C = nn.Parameter(C.clone())
        self.Q = nn.Parameter(Q.clone())  # [K,n,n] if q_per_mode else [n,n]
        self.alpha = nn.Parameter(alpha.clone())  # [B,T,K]
        self.c_shared = c_shared
        self.q_per_mode = q_per_mode
        self.state_seq = None
        self.Q_seq = None

    def reset_state(self):
        self.state_seq = None

    def compute_batch(self, a_seq, is_training=True):
        al = self.alpha
        Bsz, T, _ = al.shape
        A_seq = torch.einsum("btk,kij->btij", al, self.A)
        B_seq = torch.einsum("btk,knm->btnm", al, self.B)
        if self.c_shared:
            C_seq = self.C[0].expand(Bsz, T, -1, -1)
        else:
            C_seq = torch.einsum("btk,kpn->btpn", al, self.C)
        if self.q_per_mode:
            Q_seq = torch.einsum("btk,kij->btij", al, self.Q)
        else:
            Q_seq = self.Q.expand(Bsz, T, -1, -1)
        self.Q_seq = Q_seq
        self.state_seq = al
        self.log_qseq = torch.zeros(Bsz, T, dtype=al.dtype)
        self.log_pseq = torch.zeros(Bsz, T, dtype=al.dtype)
        return A_seq, B_seq, C_seq, Q_seq

    def elbo_terms(self):
        return self.log_qseq, self.log_pseq


class fixed_eps:
    """Context manager: makes MultivariateNormal.rsample use the supplied standard-normal draw
    (kalman_filter.py:351 -> torch.distributions.multivariate_normal._standard_normal)."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        import torch.distributions.multivariate_normal as mvn

        self._mvn = mvn
        self._orig = mvn._standard_normal
        eps = self.eps

        def _fake(shape, dtype, device):
            assert tuple(shape) == tuple(eps.shape), (shape, eps.shape)
            return eps.to(dtype=dtype, device=device)

        mvn._standard_normal = _fake
        return self

    def __exit__(self, *a):
        self._mvn._standard_normal = self._orig
        return False


def run_reference_case(case: dict, dtype=torch.float32, want_grads=True, cotangents=None, with_elbo=True):
    """Runs the unmodified reference KalmanFilter.smooth + .elbo (+ backward) on a case dict
    (see oracle/cases.py) and returns a dict of numpy-convertible tensors."""
    ref = load()
    g = lambda k: case[k].to(dtype) if case.get(k) is not None else None
    Y, U, mask, alpha, eps = g("Y"), g("U"), g("mask"), g("alpha"), g("eps")
    A, Bm, C, Q, R = g("A"), g("B"), g("C"), g("Q"), g("R")
    mu0, Sigma0 = g("mu0"), g("Sigma0")
    q_per_mode, c_shared = bool(case["q_per_mode"]), bool(case["c_shared"])
    dyn = FixedAlphaDynamics(A, Bm, C, Q, alpha, c_shared, q_per_mode)
    # std_dyn/std_obs only build kf.Q / kf.R (kalman_filter.py:22-23); overwrite them with the
    # case's matrices so non-isotropic R/Q are covered too.
    kf = ref.KalmanFilter(1.0, 1.0, mu0, Sigma0, dyn)
    kf.R.copy_(R)
    if not q_per_mode:
        kf.Q.copy_(Q)
    Yv = Y.clone().requires_grad_(want_grads)
    Uv = U.clone().requires_grad_(want_grads)
    dyn.reset_state()
    outs = kf.smooth(Yv, Uv, mask)
    names = ["mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred",
             "A_list", "B_list", "C_list"]
    res = {n: o.detach().clone() for n, o in zip(names, outs)}
    if with_elbo:
        with fixed_eps(eps):
            elbo = kf.elbo(outs[0], outs[1], Yv, Uv, outs[6], outs[7], outs[8], mask=mask)
        res["elbo"] = elbo.detach().clone()
    if want_grads:
        loss = elbo if with_elbo else 0.0
        if cotangents is not None:
            for n, o in zip(names, outs):
                if cotangents.get(n) is not None:
                    loss = loss + (cotangents[n].to(dtype) * o).sum()
        params = [Yv, Uv, dyn.alpha, dyn.A, dyn.B, dyn.C] + ([dyn.Q] if q_per_mode else [])
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        gn = ["dY", "dU", "dalpha", "dA", "dB", "dC"] + (["dQ"] if q_per_mode else [])
        for n, gr, p in zip(gn, grads, params):
            res[n] = torch.zeros_like(p) if gr is None else gr.detach().clone()
    return res

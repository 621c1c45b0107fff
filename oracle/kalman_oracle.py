"""TEST INFRASTRUCTURE — CPU oracle for the Kalman hot path.  NOT a product path.

A restatement (own code, functional style) of the algorithm in the reference's
`kvae/kalman/kalman_filter.py`, `dyn_param.py:58-60` and `switch_dyn_param.py:82-86`, written
with the same torch op sequence (matmul association order, `linalg.solve`, `linalg.cholesky`)
so that in fp32 it reproduces the reference's rounding, and dtype-generic so that an fp64 run
serves as accuracy referee.

Parity pin: `tests/test_oracle.py` checks this file against golden vectors produced by running
the UNMODIFIED reference (`oracle/make_golden.py`, committed under `tests/golden/`) and, where
`/root/reference` exists, against the live reference.  The reference ships no golden vectors of
its own for this path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.
"""
from __future__ import annotations

import math

import torch

LOG2PI = math.log(2.0 * math.pi)


def sym(X):
    return 0.5 * (X + X.mT)


# --------------------------------------------------------------------------------------
# A.0 mixing  (dyn_param.py:58-60; switch_dyn_param.py:82-86)
# --------------------------------------------------------------------------------------
def mix(alpha, A, Bm, C, Q, c_shared: bool, q_per_mode: bool):
    """alpha [B,T,K]; A [K,n,n]; Bm [K,n,m]; C [K,p,n]; Q [K,n,n] | [n,n]."""
    Bsz, T, _ = alpha.shape
    A_seq = torch.einsum("btk,kij->btij", alpha, A)
    B_seq = torch.einsum("btk,knm->btnm", alpha, Bm)
    if c_shared:
        C_seq = C[0].expand(Bsz, T, -1, -1)
    else:
        C_seq = torch.einsum("btk,kpn->btpn", alpha, C)
    if q_per_mode:
        Q_seq = torch.einsum("btk,kij->btij", alpha, Q)
    else:
        Q_seq = Q.expand(Bsz, T, -1, -1)
    return A_seq, B_seq, C_seq, Q_seq


# --------------------------------------------------------------------------------------
# A.1 filter  (kalman_filter.py:31-104 step, :107-201 loop)
# --------------------------------------------------------------------------------------
def filter_step(mu, Sigma, y, u, A, Bm, C, Q, R, m_t):
    """mu [B,n,1], Sigma [B,n,n], y [B,p,1], u [B,m,1], m_t [B]."""
    n = Sigma.shape[-1]
    I = torch.eye(n, dtype=Sigma.dtype, device=Sigma.device)
    mu_p = (A @ mu) + (Bm @ u)                                   # :65
    Sig_p = A @ Sigma @ A.mT + Q                                 # :67 (not symmetrised)
    r = y - C @ mu_p                                             # :73-75
    S = C @ Sig_p @ C.mT + R                                     # :78
    S = 0.5 * (S + S.mT)                                         # :79
    PCT = Sig_p @ C.mT                                           # :82
    Kg = torch.linalg.solve(S, PCT.mT).mT                        # :89
    Kg = m_t.view(-1, 1, 1) * Kg                                 # :92 (float multiply, no branch)
    mu_f = mu_p + Kg @ r                                         # :96
    G = I - Kg @ C                                               # :99
    Sig_f = G @ Sig_p @ G.mT + Kg @ R @ Kg.mT                    # :100
    Sig_f = 0.5 * (Sig_f + Sig_f.mT)                             # :101
    return mu_f, Sig_f, mu_p, Sig_p


def kalman_filter(Y, U, mask, A_seq, B_seq, C_seq, Q_seq, R, mu0, Sigma0):
    Bsz, T, _ = Y.shape
    mu = mu0.expand(Bsz, -1).unsqueeze(-1)
    Sigma = Sigma0.expand(Bsz, -1, -1)
    mf, Sf, mp, Sp = [], [], [], []
    for t in range(T):
        mu, Sigma, mu_p, Sig_p = filter_step(
            mu, Sigma, Y[:, t].unsqueeze(-1), U[:, t].unsqueeze(-1),
            A_seq[:, t], B_seq[:, t], C_seq[:, t], Q_seq[:, t], R.expand(Bsz, -1, -1), mask[:, t])
        mf.append(mu); Sf.append(Sigma); mp.append(mu_p); Sp.append(Sig_p)
    return torch.stack(mf, 1), torch.stack(Sf, 1), torch.stack(mp, 1), torch.stack(Sp, 1)


# --------------------------------------------------------------------------------------
# A.2 RTS smoother  (kalman_filter.py:204-237 step, :240-279 loop; uses A_{t+1}, :258)
# --------------------------------------------------------------------------------------
def rts_smooth(mf, Sf, mp, Sp, A_seq):
    T = mf.shape[1]
    mu_T, Sig_T = mf[:, -1], Sf[:, -1]
    ms, Ss = [None] * T, [None] * T
    ms[-1], Ss[-1] = mu_T, Sig_T                                  # copied, no sym (:251-256)
    for t in range(T - 2, -1, -1):
        A1 = A_seq[:, t + 1]
        J = torch.linalg.solve(Sp[:, t + 1].mT, (Sf[:, t] @ A1.mT).mT).mT      # :229
        mu_T = mf[:, t] + J @ (mu_T - mp[:, t + 1])                            # :232
        Sig_T = Sf[:, t] + J @ (Sig_T - Sp[:, t + 1]) @ J.mT                   # :234
        Sig_T = 0.5 * (Sig_T + Sig_T.mT)                                       # :235
        ms[t], Ss[t] = mu_T, Sig_T
    return torch.stack(ms, 1), torch.stack(Ss, 1)


def smooth(Y, U, mask, alpha, A, Bm, C, Q, R, mu0, Sigma0, c_shared, q_per_mode):
    """Returns the reference's 9-tuple (kalman_filter.py:274-279) plus Q_seq."""
    A_seq, B_seq, C_seq, Q_seq = mix(alpha, A, Bm, C, Q, c_shared, q_per_mode)
    mf, Sf, mp, Sp = kalman_filter(Y, U, mask, A_seq, B_seq, C_seq, Q_seq, R, mu0, Sigma0)
    ms, Ss = rts_smooth(mf, Sf, mp, Sp, A_seq)
    return (ms, Ss, mf, Sf, mp, Sp, A_seq, B_seq, C_seq), Q_seq


# --------------------------------------------------------------------------------------
# A.3 ELBO  (kalman_filter.py:282-302 _safe_cholesky, :305-401 elbo)
# --------------------------------------------------------------------------------------
def safe_cholesky(Sigma, max_tries=5, jitter_init=1e-6):
    n = Sigma.shape[-1]
    Sigma = 0.5 * (Sigma + Sigma.mT)
    eye = torch.eye(n, dtype=Sigma.dtype, device=Sigma.device)
    jitter = jitter_init
    for _ in range(max_tries):
        L, info = torch.linalg.cholesky_ex(Sigma + jitter * eye)
        if int(info.max()) == 0:
            return L
        jitter *= 10.0            # any failure in the batch bumps the jitter for all (:295-296)
    d = torch.diagonal(Sigma, dim1=-2, dim2=-1).clamp(min=1e-6)
    return torch.diag_embed(torch.sqrt(d))


def _mvn_logprob_tril(x, L):
    """log N(x; 0, L L^T); x [...,d], L [...,d,d] (torch MultivariateNormal.log_prob)."""
    d = x.shape[-1]
    w = torch.linalg.solve_triangular(L, x.unsqueeze(-1), upper=False).squeeze(-1)
    half_log_det = torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)
    return -0.5 * (d * LOG2PI + (w * w).sum(-1)) - half_log_det


def elbo(ms, Ss, Y, U, A_seq, B_seq, C_seq, Q_seq, R, mu0, Sigma0, mask, eps,
         log_pseq=None, log_qseq=None, return_terms=False):
    """ms [B,T,n,1] or [B,T,n]; eps [B,T,n] is the standard-normal draw of rsample (:351)."""
    if ms.dim() == 4:
        ms = ms.squeeze(-1)
    Bsz, T, n = ms.shape
    p = Y.shape[-1]
    L = safe_cholesky(Ss)                                                       # :348
    z = ms + (L @ eps.unsqueeze(-1)).squeeze(-1)                                # :349-351
    z_prev = z[:, :-1].unsqueeze(-1)
    Az = A_seq[:, 1:] @ z_prev                                                  # :357
    Bu = B_seq[:, 1:] @ U[:, 1:].unsqueeze(-1)                                  # :358
    mu_trans = (Az + Bu).squeeze(-1)
    L_Q = safe_cholesky(Q_seq[:, 1:])                                           # :364-365
    lp_trans = _mvn_logprob_tril(z[:, 1:] - mu_trans, L_Q)                      # :369
    mu_emiss = (C_seq @ z.unsqueeze(-1)).squeeze(-1)                            # :372
    L_R = torch.linalg.cholesky(R)                                              # :373
    lp_emiss = _mvn_logprob_tril(Y - mu_emiss, L_R) * mask                      # :374-377
    L_0 = torch.linalg.cholesky(Sigma0)                                         # :380
    lp_init = _mvn_logprob_tril(z[:, 0] - mu0, L_0)                             # :381
    entropy = -_mvn_logprob_tril(z - ms, L)                                     # :389
    num_el = mask.sum().clamp(min=1.0)                                          # :392
    tot = lp_trans.sum() + lp_emiss.sum() + lp_init.sum() + entropy.sum()
    if log_pseq is not None:
        tot = tot + log_pseq.sum() - log_qseq.sum()                             # :397-398
    val = tot / num_el
    if return_terms:
        return val, dict(trans=lp_trans.sum(), emiss=lp_emiss.sum(), init=lp_init.sum(),
                         entropy=entropy.sum(), num_el=num_el)
    return val


# --------------------------------------------------------------------------------------
# Case runner: same outputs / names as oracle.ref_shim.run_reference_case
# --------------------------------------------------------------------------------------
OUT_NAMES = ["mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred",
             "A_list", "B_list", "C_list"]


def run_case(case: dict, dtype=torch.float32, want_grads=True, cotangents=None, with_elbo=True):
    g = lambda k: case[k].to(dtype).clone()
    Y, U, mask, alpha, eps = g("Y"), g("U"), g("mask"), g("alpha"), g("eps")
    A, Bm, C, Q, R, mu0, Sigma0 = g("A"), g("B"), g("C"), g("Q"), g("R"), g("mu0"), g("Sigma0")
    q_per_mode, c_shared = bool(case["q_per_mode"]), bool(case["c_shared"])
    leaves = [Y, U, alpha, A, Bm, C] + ([Q] if q_per_mode else [])
    if want_grads:
        for t in leaves:
            t.requires_grad_(True)
    outs, Q_seq = smooth(Y, U, mask, alpha, A, Bm, C, Q, R, mu0, Sigma0, c_shared, q_per_mode)
    res = {n: o.detach().clone() for n, o in zip(OUT_NAMES, outs)}
    val = 0.0
    if with_elbo:
        val = elbo(outs[0], outs[1], Y, U, outs[6], outs[7], outs[8], Q_seq, R, mu0, Sigma0, mask, eps)
        res["elbo"] = val.detach().clone()
    if want_grads:
        loss = val
        if cotangents is not None:
            for n, o in zip(OUT_NAMES, outs):
                if cotangents.get(n) is not None:
                    loss = loss + (cotangents[n].to(dtype) * o).sum()
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        gn = ["dY", "dU", "dalpha", "dA", "dB", "dC"] + (["dQ"] if q_per_mode else [])
        for n, gr, leaf in zip(gn, grads, leaves):
            res[n] = torch.zeros_like(leaf) if gr is None else gr.detach().clone()
    return res


def smooth_elbo_fwd_bwd(case: dict, dtype=torch.float32, backward=True):
    """One pass of the hot path (the bench's CPU 'step'): smooth + elbo (+ backward)."""
    r = run_case(case, dtype=dtype, want_grads=backward)
    return r["elbo"]

"""Per-step forms under autograd: `KalmanFilter.filter_step` / `smooth_step` when a gradient is wanted, and the T-step
loop for lstm dynamics with missing observations under autograd (alpha_{t+1} then depends on the running prediction,
kvae/kalman/kalman_filter.py:159,183-185, and the gradient flows through the LSTM step by step).

Not kernel paths: batched torch operations on the GPU, differentiated by autograd as the reference's own op sequence is.
The kernels cover these calls forward-only (one launch per step / the LSTM cell inside the filter kernel); training runs
through `filter()/smooth()` with fully observed sequences (train.py:41) and never comes here.
"""
from __future__ import annotations

import torch


def filter_step_ops(mu, Sigma, y, u, A, Bm, C, Q, R, mask_t):
    """One predict / update step (kalman_filter.py:31-104).  mu [B,n,1], Sigma [B,n,n], y [B,p,1], u [B,m,1], matrices
    [B,..] (Q may be [n,n]), mask_t [B] or None.  Returns (mu_f, Sigma_f, mu_p, Sigma_p)."""
    n = Sigma.shape[-1]
    mu_p = A @ mu + Bm @ u                                                    # :65
    Sig_p = A @ Sigma @ A.transpose(-1, -2) + Q                               # :67
    innov = y - C @ mu_p                                                      # :73-75
    P = Sig_p @ C.transpose(-1, -2)                                           # :82
    S = C @ P + R                                                             # :78
    S = 0.5 * (S + S.transpose(-1, -2))                                       # :79
    K = torch.linalg.solve(S, P.transpose(-1, -2)).transpose(-1, -2)          # :89
    if mask_t is not None:
        K = mask_t.reshape(-1, 1, 1).to(K.dtype) * K                          # :92
    mu_f = mu_p + K @ innov                                                   # :96
    G = torch.eye(n, dtype=Sigma.dtype, device=Sigma.device) - K @ C          # :99
    Sig_f = G @ Sig_p @ G.transpose(-1, -2) + K @ R @ K.transpose(-1, -2)     # :100 (Joseph form)
    return mu_f, 0.5 * (Sig_f + Sig_f.transpose(-1, -2)), mu_p, Sig_p         # :101


def smooth_step_ops(Sigma_f, Sigma_p1, Sigma_s1, mu_f, mu_p1, mu_s1, A1):
    """One RTS step (kalman_filter.py:204-237): belief at t from the filtered belief at t and the predicted / smoothed
    beliefs at t+1.  Means [B,n,1]."""
    J = torch.linalg.solve(Sigma_p1.transpose(-1, -2), (Sigma_f @ A1.transpose(-1, -2)).transpose(-1, -2)).transpose(-1, -2)   # :229
    mu_s = mu_f + J @ (mu_s1 - mu_p1)                                         # :232
    Sig_s = Sigma_f + J @ (Sigma_s1 - Sigma_p1) @ J.transpose(-1, -2)         # :234
    return mu_s, 0.5 * (Sig_s + Sig_s.transpose(-1, -2))                      # :235


def filter_smooth_stepwise(kf, Y, U, mask, smooth, weights=None):
    """filter() / smooth() step by step in torch ops: the reference's loop (kalman_filter.py:141-191 and :250-271).
    weights = None: lstm dynamics, the dynamics network called between the steps (missing observations under autograd).
    weights = (alpha [B,T,K], A, B, C, Q, q_per_mode, c_shared): mixture weights known up front (switching dynamics, or
    any dynamics in a dtype the kernels do not compute in, e.g. float64)."""
    dyn = kf.dyn_params
    Bsz, T, p = Y.shape
    n, m = kf.n, kf.m
    dev, dt = Y.device, Y.dtype
    if U is None:
        U = torch.zeros(Bsz, T, m, device=dev, dtype=dt)
    if mask is None:
        mask = torch.ones(Bsz, T, device=dev, dtype=dt)
    if weights is not None:
        return _filter_smooth_given_weights(kf, Y, U, mask, smooth, *weights)
    mu = kf.mu0.to(dt).expand(Bsz, n).unsqueeze(-1)
    Sig = kf.Sigma0.to(dt).expand(Bsz, n, n)
    R = kf.R.to(dt)
    if not isinstance(dyn.state_seq, list):
        dyn.state_seq = []
    y_for_dyn = torch.zeros(Bsz, p, device=dev, dtype=dt)                     # :142
    mf, Sf, mp, Sp, Al, Bl, Cl, alphas = [], [], [], [], [], [], [], []
    for t in range(T):
        w = dyn.step_weights(y_for_dyn) if hasattr(dyn, "step_weights") else kf._ref_step_weights(y_for_dyn)
        alphas.append(w)
        A = torch.einsum("bk,kij->bij", w, dyn.A)                             # dyn_param.py:58-60
        Bm = torch.einsum("bk,kij->bij", w, dyn.B)
        C = torch.einsum("bk,kij->bij", w, dyn.C)
        m_t = mask[:, t]
        mu, Sig, mu_p, Sig_p = filter_step_ops(mu, Sig, Y[:, t].unsqueeze(-1), U[:, t].unsqueeze(-1), A, Bm, C, kf.Q.to(dt), R, m_t)
        mf.append(mu); Sf.append(Sig); mp.append(mu_p); Sp.append(Sig_p); Al.append(A); Bl.append(Bm); Cl.append(C)
        y_pred = (C @ mu_p).squeeze(-1)
        mc = m_t.reshape(Bsz, 1).to(dt)
        y_for_dyn = mc * Y[:, t] + (1.0 - mc) * y_pred                        # :183-185
    dyn.state_seq = torch.stack(alphas, 1)                                    # :188-191
    st = lambda xs: torch.stack(xs, 1)
    mf, Sf, mp, Sp, A_list, B_list, C_list = st(mf), st(Sf), st(mp), st(Sp), st(Al), st(Bl), st(Cl)
    if not smooth:
        return mf, Sf, mp, Sp, A_list, B_list, C_list
    ms, Ss = [None] * T, [None] * T
    ms[-1], Ss[-1] = mf[:, -1], Sf[:, -1]                                     # :251-256 (copied, not symmetrised)
    for t in range(T - 2, -1, -1):
        ms[t], Ss[t] = smooth_step_ops(Sf[:, t], Sp[:, t + 1], Ss[t + 1], mf[:, t], mp[:, t + 1], ms[t + 1], A_list[:, t + 1])   # :258
    return st(ms), st(Ss), mf, Sf, mp, Sp, A_list, B_list, C_list


def _filter_smooth_given_weights(kf, Y, U, mask, smooth, alpha, A, Bm, C, Q, q_per_mode, c_shared):
    Bsz, T, p = Y.shape
    n = kf.n
    dt = Y.dtype
    A_list = torch.einsum("btk,kij->btij", alpha, A.to(dt))                   # dyn_param.py:58-60 / switch_dyn_param.py:82-84
    B_list = torch.einsum("btk,kij->btij", alpha, Bm.to(dt))
    C_list = C[0].to(dt).expand(Bsz, T, -1, -1) if c_shared else torch.einsum("btk,kij->btij", alpha, C.to(dt))   # :85-86
    Q_seq = torch.einsum("btk,kij->btij", alpha, Q.to(dt)) if q_per_mode else None
    mu = kf.mu0.to(dt).expand(Bsz, n).unsqueeze(-1)
    Sig = kf.Sigma0.to(dt).expand(Bsz, n, n)
    R = kf.R.to(dt)
    mf, Sf, mp, Sp = [], [], [], []
    for t in range(T):
        Qt = Q_seq[:, t] if q_per_mode else Q.to(dt)
        mu, Sig, mu_p, Sig_p = filter_step_ops(mu, Sig, Y[:, t].unsqueeze(-1), U[:, t].unsqueeze(-1), A_list[:, t], B_list[:, t],
                                               C_list[:, t], Qt, R, mask[:, t])
        mf.append(mu); Sf.append(Sig); mp.append(mu_p); Sp.append(Sig_p)
    st = lambda xs: torch.stack(xs, 1)
    mf, Sf, mp, Sp = st(mf), st(Sf), st(mp), st(Sp)
    if q_per_mode:
        kf.dyn_params.Q_seq = Q_seq                                           # switch_dyn_param.py:84 side effect
    if not smooth:
        return mf, Sf, mp, Sp, A_list, B_list, C_list
    ms, Ss = [None] * T, [None] * T
    ms[-1], Ss[-1] = mf[:, -1], Sf[:, -1]
    for t in range(T - 2, -1, -1):
        ms[t], Ss[t] = smooth_step_ops(Sf[:, t], Sp[:, t + 1], Ss[t + 1], mf[:, t], mp[:, t + 1], ms[t + 1], A_list[:, t + 1])
    return st(ms), st(Ss), mf, Sf, mp, Sp, A_list, B_list, C_list

"""TEST INFRASTRUCTURE — explicit reverse-time adjoint of the Kalman hot path (CPU, torch).

The reference has no hand-written backward: gradients come from autograd replaying
kalman_filter.py:31-401.  The CUDA product implements an explicit adjoint instead (sweep 3:
ELBO + smoother adjoint, forward in time; sweep 4: filter + mixing adjoint, backward in time).
This module is the same adjoint written with batched torch ops in the SAME sweep structure and
with the SAME recomputation of K, S^-1, J, Ls, Qj^-1 from the saved public tensors, so the CUDA
formulas can be checked term by term.  `tests/test_oracle.py` validates it against
torch.autograd of `oracle.kalman_oracle` (i.e. of the reference's op sequence) in fp64.

Equation labels follow SURVEY.md App. A (A.3 ELBO, A.4 smoother adjoint, A.5 filter adjoint,
A.0/A.6 mixing adjoint).
"""
from __future__ import annotations

import torch

from . import kalman_oracle as ko


def _chol_bwd(L, Lbar):
    """torch cholesky_backward: gA = sym( L^-T Phi(L^T Lbar) L^-1 ), Phi = tril with halved diag."""
    phi = torch.tril(L.mT @ Lbar)
    phi = phi - 0.5 * torch.diag_embed(torch.diagonal(phi, dim1=-2, dim2=-1))
    X = torch.linalg.solve_triangular(L.mT, phi, upper=True)                 # L^-T phi
    X = torch.linalg.solve_triangular(L.mT, X.mT, upper=True).mT             # (..) L^-1
    return 0.5 * (X + X.mT)


def smooth_elbo_backward(case, saved, g_elbo=1.0, cot=None, dtype=torch.float64, jitter=1e-6):
    """case: inputs (Y,U,mask,alpha,eps,A,B,C,Q,R,mu0,Sigma0, flags); saved: the six state
    tensors of the forward pass (mus/Sigmas smooth/filt/pred).  cot: optional dense cotangents
    of the nine `smooth` outputs.  Returns dY,dU,dalpha,dA,dB,dC[,dQ]."""
    f = lambda k: case[k].to(dtype)
    Y, U, mask, alpha, eps = f("Y"), f("U"), f("mask"), f("alpha"), f("eps")
    Ak, Bk, Ck, Qk, R, mu0, Sigma0 = f("A"), f("B"), f("C"), f("Q"), f("R"), f("mu0"), f("Sigma0")
    q_per_mode, c_shared = bool(case["q_per_mode"]), bool(case["c_shared"])
    Bsz, T, p = Y.shape
    n = Ak.shape[-1]
    K = Ak.shape[0]
    sv = lambda k: saved[k].to(dtype)
    ms, Ss = sv("mus_smooth").reshape(Bsz, T, n, 1), sv("Sigmas_smooth")
    mf, Sf = sv("mus_filt").reshape(Bsz, T, n, 1), sv("Sigmas_filt")
    mp, Sp = sv("mus_pred").reshape(Bsz, T, n, 1), sv("Sigmas_pred")
    A_seq, B_seq, C_seq, Q_seq = ko.mix(alpha, Ak, Bk, Ck, Qk, c_shared, q_per_mode)   # re-mixed
    I_n = torch.eye(n, dtype=dtype)
    z3 = lambda *s: torch.zeros(Bsz, T, *s, dtype=dtype)
    cget = lambda k, shape: (cot[k].to(dtype).reshape(shape) if cot and cot.get(k) is not None
                             else torch.zeros(shape, dtype=dtype))
    # accumulators seeded with the direct cotangents of the nine outputs
    ms_b, Ss_b = cget("mus_smooth", (Bsz, T, n, 1)).clone(), cget("Sigmas_smooth", (Bsz, T, n, n)).clone()
    mf_b, Sf_b = cget("mus_filt", (Bsz, T, n, 1)).clone(), cget("Sigmas_filt", (Bsz, T, n, n)).clone()
    mp_b, Sp_b = cget("mus_pred", (Bsz, T, n, 1)).clone(), cget("Sigmas_pred", (Bsz, T, n, n)).clone()
    A_b = cget("A_list", (Bsz, T, n, n)).clone()
    B_b = cget("B_list", (Bsz, T, n, Bk.shape[-1])).clone()
    C_b = cget("C_list", (Bsz, T, p, n)).clone()
    Q_b = z3(n, n)
    dY, dU = z3(p, 1), z3(Bk.shape[-1], 1)
    Yc, Uc = Y.unsqueeze(-1), U.unsqueeze(-1)
    epsc = eps.unsqueeze(-1)

    # ------------------------------------------------------------------ A.3 ELBO adjoint
    c = g_elbo / mask.sum().clamp(min=1.0)
    Ls = torch.linalg.cholesky(ko.sym(Ss) + jitter * I_n)
    z = ms + Ls @ epsc                                                      # [B,T,n,1]
    zbar = z3(n, 1)
    if g_elbo != 0.0:
        Qj = ko.sym(Q_seq[:, 1:]) + jitter * I_n
        x = z[:, 1:] - A_seq[:, 1:] @ z[:, :-1] - B_seq[:, 1:] @ Uc[:, 1:]
        Qj_inv = torch.linalg.inv(Qj)
        q = Qj_inv @ x
        xbar = -c * q
        zbar[:, 1:] += xbar
        zbar[:, :-1] += -(A_seq[:, 1:].mT @ xbar)
        A_b[:, 1:] += -(xbar @ z[:, :-1].mT)
        B_b[:, 1:] += -(xbar @ Uc[:, 1:].mT)
        dU[:, 1:] += -(B_seq[:, 1:].mT @ xbar)
        Q_b[:, 1:] += c * 0.5 * (q @ q.mT - Qj_inv)
        e = Yc - C_seq @ z
        ebar = -c * mask.view(Bsz, T, 1, 1) * (torch.linalg.inv(R) @ e)
        dY += ebar
        C_b += -(ebar @ z.mT)
        zbar += -(C_seq.mT @ ebar)
        zbar[:, 0] += -c * (torch.linalg.inv(Sigma0) @ (z[:, 0] - mu0.view(n, 1)))
        Lbar = torch.tril(zbar @ epsc.mT) + c * torch.diag_embed(1.0 / torch.diagonal(Ls, dim1=-2, dim2=-1))
        ms_b += zbar
        Ss_b += _chol_bwd(Ls, Lbar)

    # ------------------------------------------------------------------ A.4 smoother adjoint (t = 0 .. T-2)
    for t in range(T - 1):
        A1 = A_seq[:, t + 1]
        W = Sf[:, t] @ A1.mT
        J = torch.linalg.solve(Sp[:, t + 1].mT, W.mT).mT                     # recomputed
        D = Ss[:, t + 1] - Sp[:, t + 1]
        d = ms[:, t + 1] - mp[:, t + 1]
        Gs = ko.sym(Ss_b[:, t])
        Sf_b[:, t] += Gs
        mf_b[:, t] += ms_b[:, t]
        Jb = Gs @ J @ D.mT + Gs.mT @ J @ D + ms_b[:, t] @ d.mT
        Db = J.mT @ Gs @ J
        db = J.mT @ ms_b[:, t]
        Ss_b[:, t + 1] += Db
        Sp_b[:, t + 1] -= Db
        ms_b[:, t + 1] += db
        mp_b[:, t + 1] -= db
        Wb = torch.linalg.solve(Sp[:, t + 1], Jb.mT).mT                      # Jb Sp^-T
        Sp_b[:, t + 1] += -(J.mT @ Wb)
        Sf_b[:, t] += Wb @ A1
        A_b[:, t + 1] += Wb.mT @ Sf[:, t]
    Sf_b[:, T - 1] += Ss_b[:, T - 1]
    mf_b[:, T - 1] += ms_b[:, T - 1]
    dbg = dict(Sf_b=Sf_b.clone(), Sp_b=Sp_b.clone(), mf_b=mf_b.clone().squeeze(-1), mp_b=mp_b.clone().squeeze(-1),
               A_b3=A_b.clone(), C_b3=C_b.clone(), Q_b3=Q_b.clone())

    # ------------------------------------------------------------------ A.5 filter adjoint (t = T-1 .. 0)
    for t in range(T - 1, -1, -1):
        A, Bm, C, Q = A_seq[:, t], B_seq[:, t], C_seq[:, t], Q_seq[:, t]
        if t > 0:
            Sprev, mprev = Sf[:, t - 1], mf[:, t - 1]
        else:
            Sprev, mprev = Sigma0.expand(Bsz, n, n), mu0.view(1, n, 1).expand(Bsz, n, 1)
        Sig_p, mu_p = Sp[:, t], mp[:, t]
        # recompute gain
        S = ko.sym(C @ Sig_p @ C.mT + R)
        P = Sig_p @ C.mT
        K0 = torch.linalg.solve(S, P.mT).mT
        m_t = mask[:, t].view(Bsz, 1, 1)
        Kg = m_t * K0
        r = Yc[:, t] - C @ mu_p
        G = I_n - Kg @ C
        Gf = ko.sym(Sf_b[:, t])
        Gb = Gf @ G @ Sig_p.mT + Gf.mT @ G @ Sig_p
        Spb = Sp_b[:, t] + G.mT @ Gf @ G
        Kb = Gf @ Kg @ R.mT + Gf.mT @ Kg @ R - Gb @ C.mT + mf_b[:, t] @ r.mT
        Cb = -(Kg.mT @ Gb)
        mpb = mp_b[:, t] + mf_b[:, t]
        rb = Kg.mT @ mf_b[:, t]
        K0b = m_t * Kb
        S_inv = torch.linalg.inv(S)
        Pb = K0b @ S_inv
        Sb = ko.sym(-(S_inv.mT @ K0b.mT @ K0))
        Cb = Cb + Sb @ C @ Sig_p.mT + Sb.mT @ C @ Sig_p + Pb.mT @ Sig_p - rb @ mu_p.mT
        Spb = Spb + C.mT @ Sb @ C + Pb @ C
        dY[:, t] += rb
        mpb = mpb - C.mT @ rb
        A_b[:, t] += Spb @ A @ Sprev.mT + Spb.mT @ A @ Sprev + mpb @ mprev.mT
        Q_b[:, t] += Spb
        B_b[:, t] += mpb @ Uc[:, t].mT
        dU[:, t] += Bm.mT @ mpb
        C_b[:, t] += Cb
        if t > 0:
            Sf_b[:, t - 1] += A.mT @ Spb @ A
            mf_b[:, t - 1] += A.mT @ mpb

    # ------------------------------------------------------------------ A.0 / A.6 mixing adjoint
    dalpha = torch.einsum("btij,kij->btk", A_b, Ak) + torch.einsum("btij,kij->btk", B_b, Bk)
    dA = torch.einsum("btk,btij->kij", alpha, A_b)
    dB = torch.einsum("btk,btij->kij", alpha, B_b)
    if c_shared:
        dC = torch.zeros_like(Ck)
        dC[0] = C_b.sum((0, 1))
    else:
        dalpha = dalpha + torch.einsum("btij,kij->btk", C_b, Ck)
        dC = torch.einsum("btk,btij->kij", alpha, C_b)
    out = dict(dY=dY.squeeze(-1), dU=dU.squeeze(-1), dA=dA, dB=dB, dC=dC)
    if q_per_mode:
        dalpha = dalpha + torch.einsum("btij,kij->btk", Q_b, Qk)
        out["dQ"] = torch.einsum("btk,btij->kij", alpha, Q_b)
    out["dalpha"] = dalpha
    out["_dbg"] = dbg
    return out

"""ELBO for ARBITRARY per-step matrices: `KalmanFilter.elbo` called with `A_list / B_list / C_list` tensors that did not come
from this object's `filter()` / `smooth()` (the kernels re-mix those from alpha), or with an explicit `Q_list`
(kvae/kalman/kalman_filter.py:305-401, Q fallback :342-345).

Not a kernel path: batched torch operations on the GPU (cuBLAS / cuSOLVER through ATen), differentiable by autograd with
respect to every argument exactly as the reference's op sequence is.  It exists so that the call does not raise; the
hot path (lists produced by this object) never comes here.  CUDA tensors only, like everything else in the package.
"""
from __future__ import annotations

import math

import torch

_LOG2PI = math.log(2.0 * math.pi)


def _factor(S, tries=5, jitter=1e-6):
    """Lower Cholesky factor of sym(S) + jitter I with the reference's ladder (:282-302): a failure anywhere in the batch
    multiplies the jitter by ten for the whole batch; after `tries` failures the clamped diagonal's square root."""
    S = 0.5 * (S + S.transpose(-1, -2))
    eye = torch.eye(S.shape[-1], dtype=S.dtype, device=S.device)
    for _ in range(tries):
        L, info = torch.linalg.cholesky_ex(S + jitter * eye)
        if not bool(info.any()):
            return L
        jitter *= 10.0
    return torch.diag_embed(torch.diagonal(S, dim1=-2, dim2=-1).clamp(min=1e-6).sqrt())


def _log_normal_zero_mean(x, L):
    """sum-free log N(x; 0, L L^T) per leading index; x [..., d], L [..., d, d] or [d, d]."""
    w = torch.linalg.solve_triangular(L, x.unsqueeze(-1), upper=False).squeeze(-1)
    return -0.5 * (x.shape[-1] * _LOG2PI + (w * w).sum(-1)) - torch.diagonal(L, dim1=-2, dim2=-1).log().sum(-1)


def elbo_given_lists(mu, Sigma, y, u, A_list, B_list, C_list, Q, R, mu0, Sigma0, mask, eps, extra=None):
    """mu [B,T,n] or [B,T,n,1]; Sigma [B,T,n,n]; y [B,T,p]; u [B,T,m]; lists [B,T,..]; Q [B,T,n,n] or [n,n];
    mask [B,T] or None; eps [B,T,n] standard normal; extra: 0-dim tensor added before the normalisation (log p - log q of
    the regime chain, :382-386).  Returns the scalar of :392-400."""
    if not mu.is_cuda:
        from .capi import KvaeError
        raise KvaeError("KalmanFilter.elbo (B200-native) needs CUDA tensors; there is no CPU path")
    if mu.dim() == 4:
        mu = mu.squeeze(-1)
    if u.dim() == 4:
        u = u.squeeze(-1)
    Bsz, T, n = mu.shape
    if mask is None:
        mask = torch.ones(Bsz, T, dtype=mu.dtype, device=mu.device)
    Ls = _factor(Sigma)                                                        # :348
    z = mu + (Ls @ eps.unsqueeze(-1)).squeeze(-1)                              # reparameterised sample, :349-351
    total = -_log_normal_zero_mean(z - mu, Ls).sum()                           # entropy of q, :389
    total = total + _log_normal_zero_mean(z[:, 0] - mu0, torch.linalg.cholesky(Sigma0)).sum()          # :380-381
    resid_y = y - (C_list @ z.unsqueeze(-1)).squeeze(-1)                       # :372
    total = total + (_log_normal_zero_mean(resid_y, torch.linalg.cholesky(R)) * mask).sum()            # :373-377
    if T > 1:
        drift = (A_list[:, 1:] @ z[:, :-1].unsqueeze(-1) + B_list[:, 1:] @ u[:, 1:].unsqueeze(-1)).squeeze(-1)   # :357-358
        Qt = Q[:, 1:] if Q.dim() == 4 else Q
        total = total + _log_normal_zero_mean(z[:, 1:] - drift, _factor(Qt)).sum()                     # :364-369
    if extra is not None:
        total = total + extra
    return total / mask.sum().clamp(min=1.0)                                   # :392-400

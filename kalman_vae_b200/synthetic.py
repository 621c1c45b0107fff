"""Synthetic workloads for the Kalman hot path (SURVEY.md §8(d) "Synthetic inputs").

Bouncing-ball-shaped observation sequences, softmax mixture weights, the block / Bernoulli
imputation masks of the reference (kvae/train/imputation.py:4-25) and the KVAE parameter
initialisation (kvae/model/model.py:33-78, kvae/utils/config.py:11-26).  Pure torch, device
agnostic; used by bench.py, the tests and the golden-vector generator.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class Shape:
    B: int
    T: int
    n: int
    p: int
    m: int
    K: int
    q_per_mode: bool = False  # switching variant: Q_t = sum_k alpha_k Q_k
    c_shared: bool = False    # switching variant: C_t = C_0


# BASELINE.json configs -> shapes (SURVEY.md §8(d) "Configs -> shapes")
CONFIGS = {
    "cfg1": Shape(32, 20, 4, 2, 4, 3),
    "cfg1_switching": Shape(32, 20, 4, 2, 4, 3, True, True),
    "cfg2": Shape(8192, 20, 4, 2, 4, 3),
    "cfg3": Shape(65536, 1000, 4, 2, 4, 3),
    "cfg4": Shape(16384, 200, 16, 8, 16, 8, True, True),
}


def bouncing_ball(B, T, p, gen, noise_std=0.03, dtype=torch.float32):
    """Per sequence: a point moving at constant speed |v|~U(0.05,0.15)/step with a random
    heading, reflecting in the box [-1,1]^p, plus N(0, noise_std) encoder-like noise."""
    pos = torch.rand(B, p, generator=gen, dtype=torch.float64) * 2 - 1
    d = torch.randn(B, p, generator=gen, dtype=torch.float64)
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    speed = 0.05 + 0.10 * torch.rand(B, 1, generator=gen, dtype=torch.float64)
    vel = d * speed
    t = torch.arange(T, dtype=torch.float64).view(1, T, 1)
    x = pos.unsqueeze(1) + vel.unsqueeze(1) * t  # unfolded straight line
    # reflect into [-1,1]: triangle wave of period 4
    x = (x + 1.0) % 4.0
    x = torch.where(x > 2.0, 4.0 - x, x) - 1.0
    x = x + noise_std * torch.randn(B, T, p, generator=gen, dtype=torch.float64)
    return x.to(dtype)


def block_mask(B, T, t_init=4, t_hide=12, tile=False, dtype=torch.float32):
    """'observe t_init, hide t_hide' (imputation.py:4-12); tile=True repeats the pattern."""
    mask = torch.ones(B, T, dtype=dtype)
    if not tile:
        mask[:, t_init:min(t_init + t_hide, T)] = 0.0
        return mask
    period = t_init + t_hide
    tt = torch.arange(T) % period
    mask[:, tt >= t_init] = 0.0
    return mask


def bernoulli_mask(B, T, gen, drop_prob=0.5, t_init=4, dtype=torch.float32):
    """imputation.py:15-25: first t_init observed, the rest dropped with prob drop_prob."""
    mask = torch.ones(B, T, dtype=dtype)
    if T > t_init:
        keep = torch.rand(B, T - t_init, generator=gen) < (1.0 - drop_prob)
        mask[:, t_init:] = keep.to(dtype)
    return mask


def make_params(shape: Shape, gen, c_std=0.05, a_std=0.05, dtype=torch.float32,
                noise_transition=0.02, noise_emission=0.03, init_cov=20.0, nonsym_q=False):
    n, p, m, K = shape.n, shape.p, shape.m, shape.K
    r = lambda *s: torch.randn(*s, generator=gen, dtype=torch.float64)
    A = torch.eye(n, dtype=torch.float64).repeat(K, 1, 1) + a_std * r(K, n, n)
    Bm = c_std * r(K, n, m)
    C = c_std * r(K, p, n)
    if shape.q_per_mode:
        Q = noise_transition * torch.eye(n, dtype=torch.float64).repeat(K, 1, 1)
        if nonsym_q:
            Q = Q * (1.0 + 0.3 * torch.rand(K, 1, 1, generator=gen, dtype=torch.float64))
            Q = Q + 0.1 * noise_transition * r(K, n, n)
    else:
        Q = noise_transition * torch.eye(n, dtype=torch.float64)
    R = noise_emission * torch.eye(p, dtype=torch.float64)
    mu0 = torch.zeros(n, dtype=torch.float64)
    Sigma0 = init_cov * torch.eye(n, dtype=torch.float64)
    return {k: v.to(dtype) for k, v in dict(A=A, B=Bm, C=C, Q=Q, R=R, mu0=mu0, Sigma0=Sigma0).items()}


def make_case(shape: Shape, seed=10, mask_kind="ones", zero_u=True, c_std=0.05, nonsym_q=False,
              dtype=torch.float32):
    """A complete, seeded input set for the hot path (CPU tensors)."""
    gen = torch.Generator().manual_seed(seed)
    B, T = shape.B, shape.T
    case = make_params(shape, gen, c_std=c_std, dtype=dtype, nonsym_q=nonsym_q)
    case["Y"] = bouncing_ball(B, T, shape.p, gen, dtype=dtype)
    if zero_u:
        case["U"] = torch.zeros(B, T, shape.m, dtype=dtype)
    else:
        case["U"] = (0.5 * torch.randn(B, T, shape.m, generator=gen, dtype=torch.float64)).to(dtype)
    logits = torch.randn(B, T, shape.K, generator=gen, dtype=torch.float64)
    case["alpha"] = torch.softmax(logits, dim=-1).to(dtype)
    if mask_kind == "ones":
        case["mask"] = torch.ones(B, T, dtype=dtype)
    elif mask_kind == "block":
        case["mask"] = block_mask(B, T, tile=T > 16, dtype=dtype)
    elif mask_kind == "bernoulli":
        case["mask"] = bernoulli_mask(B, T, gen, dtype=dtype)
    elif mask_kind == "zeros":
        case["mask"] = torch.zeros(B, T, dtype=dtype)
    elif mask_kind == "fractional":
        case["mask"] = torch.rand(B, T, generator=gen, dtype=torch.float64).to(dtype)
    else:
        raise ValueError(mask_kind)
    case["eps"] = torch.randn(B, T, shape.n, generator=gen, dtype=torch.float64).to(dtype)
    case["q_per_mode"] = shape.q_per_mode
    case["c_shared"] = shape.c_shared
    return case

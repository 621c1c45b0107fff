"""TEST INFRASTRUCTURE — CPU restatement of the SKVAE regime sampler.

Restates SwitchingDynamicsParameter.compute_batch, kvae/kalman/switch_dyn_param.py:51-79 (own code, same torch op
order), with the Gumbel noise as an explicit input: torch.nn.functional.gumbel_softmax draws
`gumbels = -empty_like(logits).exponential_().log()` and returns `((logits + gumbels) / tau).softmax(dim)` (soft) or
`y_hard - y_soft.detach() + y_soft` (hard); `gumbel_softmax_given` is that function with the draw factored out.

Pinned against the live reference by oracle/make_golden_regime.py -> tests/golden/regime_*.npz
(tests/test_oracle.py).  Only tests/ and bench tooling may import this module; kalman_vae_b200/ never does.
"""
from __future__ import annotations

import torch


def gumbel_softmax_given(logits, gumbels, tau, hard):
    y_soft = ((logits + gumbels) / tau).softmax(-1)
    if hard:
        idx = y_soft.max(-1, keepdim=True)[1]
        y_hard = torch.zeros_like(logits).scatter_(-1, idx, 1.0)
        return y_hard - y_soft.detach() + y_soft
    return y_soft


def regime_sample(logits, init_logits, gumbel, trans, tau=0.5, hard=False):
    """logits [B,T,K,K], init_logits [B,K], gumbel [B,T,K], trans [K,K] -> (y_seq [B,T,K], log_q [B,T], log_p [B,T])."""
    B, T, K, _ = logits.shape
    y0 = gumbel_softmax_given(init_logits, gumbel[:, 0], tau, hard)                     # :52
    log_q0 = torch.log_softmax(init_logits, dim=-1)                                     # :53
    log_p0 = torch.full_like(log_q0, 1.0 / K).log()                                     # :54
    ys, lq, lp = [y0], [(y0 * log_q0).sum(-1)], [(y0 * log_p0).sum(-1)]                 # :60-62
    y_prev = y0
    for t in range(1, T):                                                               # :67-79
        l_t = torch.matmul(y_prev.unsqueeze(1), logits[:, t]).squeeze(1)
        y_t = gumbel_softmax_given(l_t, gumbel[:, t], tau, hard)
        lq.append((y_t * torch.log_softmax(l_t, dim=-1)).sum(-1))
        tp = torch.matmul(y_prev.unsqueeze(1), trans).squeeze(1)
        lp.append((y_t * torch.log(tp.clamp_min(1e-8))).sum(-1))
        ys.append(y_t)
        y_prev = y_t
    return torch.stack(ys, 1), torch.stack(lq, 1), torch.stack(lp, 1)


def regime_sample_with_grads(case, dtype=torch.float64):
    """case: dict(logits, init_logits, gumbel, trans, tau, hard, cot_y, cot_q, cot_p) -> outputs and gradients of
    <cot_y, y_seq> + <cot_q, log_q> + <cot_p, log_p> w.r.t. logits / init_logits (autograd of the restatement)."""
    f = lambda k: case[k].to(dtype)
    logits = f("logits").clone().requires_grad_(True)
    init = f("init_logits").clone().requires_grad_(True)
    y, lq, lp = regime_sample(logits, init, f("gumbel"), f("trans"), float(torch.as_tensor(case["tau"]).reshape(-1)[0]), bool(case["hard"]))
    loss = (f("cot_y") * y).sum() + (f("cot_q") * lq).sum() + (f("cot_p") * lp).sum()
    d_logits, d_init = torch.autograd.grad(loss, [logits, init], allow_unused=True)
    if d_logits is None:      # T = 1: the transition logits are never read
        d_logits = torch.zeros_like(logits)
    return dict(y_seq=y.detach(), log_q=lq.detach(), log_p=lp.detach(), d_logits=d_logits, d_init=d_init)

"""TEST INFRASTRUCTURE — freezes outputs of the UNMODIFIED reference regime sampler into tests/golden/regime_*.npz.

Runs kvae.kalman.switch_dyn_param.SwitchingDynamicsParameter.compute_batch (reference, imported through
oracle/ref_shim.py) in fp32 and fp64 with
  * a stand-in `markov_regime_posterior` (constructor argument of the reference class, switch_dyn_param.py:7,27) that
    returns fixed leaf tensors (logits [B,T,K,K], init_logits [B,K]) — the bi-GRU is outside this path, and
  * torch's gumbel_softmax patched at the module-level name the reference imported (switch_dyn_param.py:5) so that it
    consumes pre-drawn noise (same formula as torch.nn.functional.gumbel_softmax, see oracle/regime_oracle.py).
Gradients: autograd of <cot_y, state_seq> + <cot_q, log_qseq> + <cot_p, log_pseq> w.r.t. logits / init_logits.

    python -m oracle.make_golden_regime      (needs /root/reference; run in the build container only)
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle import ref_shim
from oracle.regime_oracle import gumbel_softmax_given

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (B, T, K, hard, p_stay, seed)
    "regime_k3_soft": (37, 20, 3, False, 0.9, 1),
    "regime_k3_hard": (37, 20, 3, True, 0.9, 2),
    "regime_k8_soft": (19, 33, 8, False, 0.8, 3),
    "regime_k2_T1": (5, 1, 2, False, 0.9, 4),
}


def make_inputs(B, T, K, hard, p_stay, seed):
    g = torch.Generator().manual_seed(seed)
    logits = 1.5 * torch.randn(B, T, K, K, generator=g)
    init_logits = torch.randn(B, K, generator=g)
    gumbel = -torch.empty(B, T, K).exponential_(generator=g).log()
    trans = torch.ones(K, K) * ((1 - p_stay) / (K - 1))
    trans.fill_diagonal_(p_stay)                                          # StickyRegimePrior, switch_dyn_param.py:98-103
    return dict(logits=logits, init_logits=init_logits, gumbel=gumbel, trans=trans, tau=0.5, hard=hard,
                cot_y=torch.randn(B, T, K, generator=g), cot_q=torch.randn(B, T, generator=g),
                cot_p=torch.randn(B, T, generator=g), p_stay=p_stay)


class _FixedPosterior(torch.nn.Module):
    def __init__(self, logits, init_logits):
        super().__init__()
        self.logits, self.init_logits = logits, init_logits

    def forward(self, a_seq):
        return self.logits, self.init_logits


def run_reference(case, dtype):
    ref = ref_shim.load()
    B, T, K, _ = case["logits"].shape
    n, m, p = 4, 4, 2
    logits = case["logits"].to(dtype).clone().requires_grad_(True)
    init = case["init_logits"].to(dtype).clone().requires_grad_(True)
    prior = ref.switch_mod.StickyRegimePrior(K, p_stay=case["p_stay"])
    dyn = ref.SwitchingDynamicsParameter(torch.zeros(K, n, n, dtype=dtype), torch.zeros(K, n, m, dtype=dtype),
                                         torch.zeros(K, p, n, dtype=dtype), prior=prior,
                                         markov_regime_posterior=_FixedPosterior(logits, init))
    dyn.tau = case["tau"]
    noise = case["gumbel"].to(dtype)
    calls = {"t": 0}

    def det_gumbel_softmax(lg, tau=1.0, hard=False, dim=-1):
        t = calls["t"]
        calls["t"] += 1
        return gumbel_softmax_given(lg, noise[:, t], tau, hard)

    orig = ref.switch_mod.gumbel_softmax
    ref.switch_mod.gumbel_softmax = det_gumbel_softmax
    try:
        dyn.compute_batch(torch.zeros(B, T, p, dtype=dtype), is_training=not case["hard"])
    finally:
        ref.switch_mod.gumbel_softmax = orig
    y, (lq, lp) = dyn.state_seq, dyn.elbo_terms()
    loss = (case["cot_y"].to(dtype) * y).sum() + (case["cot_q"].to(dtype) * lq).sum() + (case["cot_p"].to(dtype) * lp).sum()
    d_logits, d_init = torch.autograd.grad(loss, [logits, init], allow_unused=True)
    if d_logits is None:
        d_logits = torch.zeros_like(logits)
    return dict(y_seq=y.detach(), log_q=lq.detach(), log_p=lp.detach(), d_logits=d_logits, d_init=d_init)


def main():
    torch.set_num_threads(1)
    for name, spec in CASES.items():
        case = make_inputs(*spec)
        arrs = {}
        for k, v in case.items():
            # floats as 1-element arrays (tests/_util.load_golden reads 0-dim entries as booleans)
            arrs["in_" + k] = v.numpy() if torch.is_tensor(v) else (np.asarray(v) if isinstance(v, bool) else np.asarray([v], dtype=np.float64))
        for tag, dt in (("ref32_", torch.float32), ("ref64_", torch.float64)):
            for k, v in run_reference(case, dt).items():
                arrs[tag + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
        print("wrote", name, {k: v.shape for k, v in arrs.items() if k.startswith("ref32_")})


if __name__ == "__main__":
    main()

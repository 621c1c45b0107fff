"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.make_golden

The reference ships no golden vectors for the Kalman path (its tests/fixtures/*.pt are git-ignored,
SURVEY.md §4), so parity is pinned by executing kvae/kalman/kalman_filter.py itself on seeded inputs
and freezing inputs + outputs here.  Two kinds of fixture:

  kalman_*.npz : KalmanFilter.smooth + .elbo (+ autograd gradients) driven with fixed mixture weights
                 (oracle.ref_shim.FixedAlphaDynamics); fp32 results and an fp64 run of the same code
                 (the accuracy referee).
  kvae_*.npz   : the whole reference KVAE (real LSTM / bi-GRU dynamics networks) on the recipe of the
                 reference's own regression test tests/test_imputation_stability.py:16-53 (weights
                 randn*0.01 seed 42, input seed 123, B=2, T=10, mask hides 4:10); stores the encoder
                 sample `a`, the Kalman block's state dict and every Kalman output, so the drop-in can be
                 checked from `a` onwards.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from kalman_vae_b200.synthetic import Shape, make_case
from oracle import ref_shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

KALMAN_CASES = {
    # name: (shape, make_case kwargs, with output cotangents)
    "kalman_lstm": (Shape(4, 20, 4, 2, 4, 3), dict(seed=11, mask_kind="bernoulli", zero_u=False, c_std=0.3), True),
    "kalman_lstm_default": (Shape(4, 20, 4, 2, 4, 3), dict(seed=10, mask_kind="ones", zero_u=True, c_std=0.05), False),
    "kalman_switch": (Shape(4, 20, 4, 2, 4, 3, True, True),
                      dict(seed=12, mask_kind="block", zero_u=False, c_std=0.3, nonsym_q=True), True),
    "kalman_zero_mask": (Shape(3, 7, 4, 2, 4, 3), dict(seed=13, mask_kind="zeros", zero_u=False, c_std=0.3), False),
    "kalman_T1": (Shape(3, 1, 4, 2, 4, 3, True, True), dict(seed=14, mask_kind="ones", zero_u=False, c_std=0.3), False),
    "kalman_fractional": (Shape(3, 9, 4, 2, 4, 1), dict(seed=15, mask_kind="fractional", zero_u=False, c_std=0.3), True),
    "kalman_n8": (Shape(3, 10, 8, 4, 8, 4, True, True),
                  dict(seed=16, mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), False),
    "kalman_n16": (Shape(2, 12, 16, 8, 16, 8, True, True),
                   dict(seed=17, mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=True), True),
}


def _np(t):
    return t.detach().cpu().numpy()


def rocket_case():
    """The reference's own toy LGSSM (kvae/kalman/test_filter.py:5-57): n=2, p=1, m=1, K=1, T=100,
    switching dynamics object with its default Q = I."""
    dt, g, T = 0.1, -9.81, 10.0
    N = int(T / dt)
    t = np.arange(N) * dt
    std_obs, std_dyn = 4.0, 2.0
    x = np.zeros((N, 2))
    for n in range(N - 1):
        a = (20.0 if t[n] < 6.0 else 0.0) + g
        x[n + 1, 0] = x[n, 0] + x[n, 1] * dt + 0.5 * a * dt * dt
        x[n + 1, 1] = x[n, 1] + a * dt
    rng = np.random.RandomState(0)
    a_spec = np.r_[((x[1:, 1] - x[:-1, 1]) / dt - g)[0], (x[1:, 1] - x[:-1, 1]) / dt - g]
    u_meas = a_spec + g + rng.randn(N) * std_dyn ** 2
    z_meas = x[:, 0] + rng.randn(N) * std_obs ** 2
    f = lambda v: torch.tensor(v, dtype=torch.float32)
    case = dict(
        A=f([[[1.0, dt], [0.0, 1.0]]]), B=f([[[0.5 * dt ** 2], [dt]]]), C=f([[[1.0, 0.0]]]),
        Q=torch.eye(2).unsqueeze(0), R=f([[std_obs ** 2]]), mu0=torch.zeros(2), Sigma0=torch.eye(2),
        Y=f(z_meas).view(1, N, 1), U=f(u_meas).view(1, N, 1), mask=torch.ones(1, N),
        alpha=torch.ones(1, N, 1), q_per_mode=True, c_shared=True)
    case["eps"] = torch.randn(1, N, 2, generator=torch.Generator().manual_seed(3))
    return case


def save_kalman(name, case, with_cot):
    shp = {k: tuple(v.shape) for k, v in case.items() if torch.is_tensor(v)}
    B, T, p = shp["Y"]
    K, n, m = shp["B"]
    cot = None
    if with_cot:
        gen = torch.Generator().manual_seed(99)
        cs = dict(mus_smooth=(B, T, n, 1), Sigmas_smooth=(B, T, n, n), mus_filt=(B, T, n, 1), Sigmas_filt=(B, T, n, n),
                  mus_pred=(B, T, n, 1), Sigmas_pred=(B, T, n, n), A_list=(B, T, n, n), B_list=(B, T, n, m),
                  C_list=(B, T, p, n))
        cot = {k: 0.1 * torch.randn(*s, generator=gen) for k, s in cs.items()}
        if case["c_shared"]:
            cot["C_list"] = None
    arrs = {}
    for k, v in case.items():
        arrs["in_" + k] = _np(v) if torch.is_tensor(v) else np.array(int(v))
    if cot:
        for k, v in cot.items():
            if v is not None:
                arrs["cot_" + k] = _np(v)
    # the reference's elbo() cannot run with T == 1 (empty transition batch, kalman_filter.py:367)
    we = T > 1
    wg = we or cot is not None
    r32 = ref_shim.run_reference_case(case, torch.float32, cotangents=cot, with_elbo=we, want_grads=wg)
    r64 = ref_shim.run_reference_case(case, torch.float64, cotangents=cot, with_elbo=we, want_grads=wg)
    for k, v in r32.items():
        arrs["ref32_" + k] = _np(v)
    for k, v in r64.items():
        arrs["ref64_" + k] = _np(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrs)
    print("wrote", name, {k: v.shape for k, v in arrs.items() if k.startswith("ref32_")})


def save_kvae(kind):
    """tests/test_imputation_stability.py:16-53 recipe; everything the Kalman block saw and produced."""
    ref = ref_shim.load()
    cfg = ref.KVAEConfig(dynamics_model=kind)
    model = ref.KVAE(cfg)
    torch.manual_seed(42)
    for prm in model.parameters():
        if prm.requires_grad:
            prm.data = torch.randn_like(prm.data) * 0.01
    model.eval()
    torch.manual_seed(123)
    x = torch.randn(2, 10, 1, 32, 32)
    B, T = 2, 10
    mask = torch.ones(B, T)
    mask[:, 4:10] = 0.0
    arrs = {}
    # deterministic Gumbel noise for the switching posterior (switch_dyn_param.py:52,69)
    gnoise = -torch.empty(T, B, cfg.num_modes).exponential_(generator=torch.Generator().manual_seed(7)).log()
    arrs["gumbel_noise"] = _np(gnoise)
    calls = {"i": 0}

    def det_gumbel_softmax(logits, tau=1.0, hard=False, dim=-1):
        g = gnoise[calls["i"] % T].to(logits.dtype)
        calls["i"] += 1
        y_soft = ((logits + g) / tau).softmax(dim)
        if hard:
            idx = y_soft.max(dim, keepdim=True)[1]
            y_hard = torch.zeros_like(logits).scatter_(dim, idx, 1.0)
            return y_hard - y_soft.detach() + y_soft
        return y_soft

    orig = ref.switch_mod.gumbel_softmax
    ref.switch_mod.gumbel_softmax = det_gumbel_softmax
    try:
        with torch.no_grad():
            torch.manual_seed(5)
            out = model.forward(x, mask=mask)
            calls["i"] = 0
            torch.manual_seed(5)
            imp = model.impute(x, mask=mask)
    finally:
        ref.switch_mod.gumbel_softmax = orig
    assert torch.equal(out["a_samples"], imp["a_vae"])
    # fp64 referee: the reference's OWN Kalman block in double on the same (a, u, mask, noise): tells the reference's fp32
    # rounding floor from an error of the drop-in (tests/_util.check_close)
    import copy
    kf64 = copy.deepcopy(model.kalman_filter).double()
    ref.switch_mod.gumbel_softmax = det_gumbel_softmax
    try:
        with torch.no_grad():
            calls["i"] = 0
            kf64.dyn_params.reset_state()
            outs64 = kf64.smooth(out["a_samples"].double().clone(), out["u"].double().clone(), mask=mask.double())
    finally:
        ref.switch_mod.gumbel_softmax = orig
    for k, v in zip(("mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred", "A_list", "B_list",
                     "C_list"), outs64):
        arrs["ref64_" + k] = v.detach().contiguous().numpy()
    arrs["a"] = _np(out["a_samples"])
    arrs["mask"] = _np(mask)
    arrs["u"] = _np(out["u"])
    for k in ("mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred"):
        arrs[k] = _np(out[k])
    for k, v in zip(("A_list", "B_list", "C_list"), out["ABC"]):
        arrs[k] = _np(v.contiguous())
    arrs["state_probs"] = _np(out["state_probs"])
    arrs["a_imputed"] = _np(imp["a_imputed"])
    arrs["a_filtered"] = _np(imp["a_filtered"])
    for k, v in model.kalman_filter.state_dict().items():
        arrs["sd_" + k] = _np(v)
    np.savez_compressed(os.path.join(OUT, f"kvae_{kind}.npz"), **arrs)
    print("wrote kvae_" + kind, sorted(k for k in arrs if k.startswith("sd_")))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)  # bit-reproducible reductions
    only_kvae = len(sys.argv) > 1 and sys.argv[1] == "kvae"     # `python oracle/make_golden.py kvae`: the two kvae_* files only
    if not only_kvae:
        for name, (shape, kw, with_cot) in KALMAN_CASES.items():
            save_kalman(name, make_case(shape, **kw), with_cot)
        save_kalman("kalman_rocket", rocket_case(), False)
    for kind in ("lstm", "switching"):
        save_kvae(kind)


if __name__ == "__main__":
    main()

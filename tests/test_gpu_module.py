"""GPU tests of the drop-in boundary: kalman_vae_b200.KalmanFilter used exactly like the reference's
kvae.kalman.kalman_filter.KalmanFilter (same constructor / calls / tuples / layouts), checked against
golden vectors of the unmodified reference (kalman_*: fixed mixture weights incl. autograd gradients;
kvae_*: the whole reference KVAE with its real LSTM / bi-GRU dynamics networks on the recipe of the
reference's own tests/test_imputation_stability.py)."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from kalman_vae_b200 import KalmanFilter, DynamicsParameter, SwitchingDynamicsParameter
from kalman_vae_b200 import dyn_param as dp_mod
from kalman_vae_b200.dyn_param import MarkovVariationalRegimePosterior, StickyRegimePrior
from tests._util import GOLDEN, GRAD_NAMES, OUT_NAMES, check_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class FixedWeights(nn.Module):
    """dyn_params protocol object with externally given mixture weights (test double of the LSTM / GRU)."""

    def __init__(self, case):
        super().__init__()
        self.is_switching_dynamics = bool(case["q_per_mode"])
        self.A, self.B, self.C = (nn.Parameter(case[k].clone()) for k in ("A", "B", "C"))
        if self.is_switching_dynamics:
            self.Q = nn.Parameter(case["Q"].clone())
        self.alpha = nn.Parameter(case["alpha"].clone())
        self.K = self.A.shape[0]
        self.state_seq = None

    def reset_state(self):
        self.state_seq = None

    def compute_weights(self, a_seq, is_training=True):
        B, T, _ = self.alpha.shape
        self.state_seq = self.alpha
        self.log_qseq = torch.zeros(B, T, device=self.alpha.device)
        self.log_pseq = torch.zeros(B, T, device=self.alpha.device)
        return self.alpha

    def elbo_terms(self):
        return self.log_qseq, self.log_pseq


def make_kf(case):
    dyn = FixedWeights(case)
    kf = KalmanFilter(1.0, 1.0, case["mu0"], case["Sigma0"], dyn)
    kf.R.copy_(case["R"])
    if not case["q_per_mode"]:
        kf.Q.copy_(case["Q"])
    return kf.to(DEV), dyn


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_lstm_default", "kalman_switch", "kalman_fractional", "kalman_n16",
                                  "kalman_rocket"])
def test_smooth_elbo_backward_through_module_api(name, monkeypatch):
    case, cot, r32, r64 = load_golden(name)
    kf, dyn = make_kf(case)
    assert (kf.n, kf.m, kf.p) == (case["A"].shape[-1], case["B"].shape[-1], case["C"].shape[-2])
    Y = case["Y"].to(DEV).requires_grad_(True)
    U = case["U"].to(DEV).requires_grad_(True)
    mask = case["mask"].to(DEV)
    dyn.reset_state()
    outs = kf.smooth(Y, U, mask)
    assert len(outs) == 9
    B, T, p = case["Y"].shape
    n = kf.n
    assert outs[0].shape == (B, T, n, 1) and outs[1].shape == (B, T, n, n) and outs[8].shape == (B, T, p, n)
    for k, o in zip(OUT_NAMES, outs):
        check_close(f"{name}.{k}", o, r32[k], r64[k])
    # elbo(): the eps draw is torch.empty(B,T,n).normal_() on the device; inject the golden's draw
    eps = case["eps"].to(DEV)
    monkeypatch.setattr(kf, "_draw_eps", lambda B, T, n, like: eps)
    val = kf.elbo(outs[0], outs[1], Y, U, outs[6], outs[7], outs[8], mask=mask)
    check_close(f"{name}.elbo", val, r32["elbo"], r64["elbo"])
    loss = val
    if cot:
        for k, o in zip(OUT_NAMES, outs):
            if k in cot:
                loss = loss + (cot[k].to(DEV) * o).sum()
    params = [Y, U, dyn.alpha, dyn.A, dyn.B, dyn.C] + ([dyn.Q] if case["q_per_mode"] else [])
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, g in zip(GRAD_NAMES, grads):
        assert g is not None, k
        check_close(f"{name}.{k}", g, r32[k], r64[k])


def test_filter_returns_seven_tuple_and_matches():
    case, _, r32, r64 = load_golden("kalman_lstm")
    kf, dyn = make_kf(case)
    with torch.no_grad():
        outs = kf.filter(case["Y"].to(DEV), case["U"].to(DEV), case["mask"].to(DEV))
    assert len(outs) == 7
    for k, o in zip(OUT_NAMES[2:], outs):
        check_close(k, o, r32[k], r64[k])


def test_cpu_input_raises():
    case = load_golden("kalman_lstm")[0]
    kf, _ = make_kf(case)
    with pytest.raises(Exception):
        kf.smooth(case["Y"], case["U"], case["mask"])


def _load_kvae(kind):
    z = np.load(os.path.join(GOLDEN, f"kvae_{kind}.npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _kvae_kalman_block(kind, g):
    """The Kalman block of the reference KVAE (model.py:33-78) built from this package's classes."""
    K, n, p, m, hidden = 3, 4, 2, 4, 50
    A, Bm, C = torch.zeros(K, n, n), torch.zeros(K, n, m), torch.zeros(K, p, n)
    if kind == "switching":
        dyn = SwitchingDynamicsParameter(A, Bm, C, Q=torch.zeros(K, n, n), prior=StickyRegimePrior(K, p_stay=0.8),
                                         hidden_lstm=hidden,
                                         markov_regime_posterior=MarkovVariationalRegimePosterior(K, input_dim=p, hidden_size=hidden))
        dyn.tau = 1.0                                                      # config.tau_init (model.py:62)
    else:
        dyn = DynamicsParameter(A, Bm, C, hidden_lstm=hidden)
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, torch.zeros(n), 20.0 * torch.eye(n), dyn)
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd_")}
    missing, unexpected = kf.load_state_dict(sd, strict=True)              # same keys as the reference's state dict
    assert not missing and not unexpected
    return kf.to(DEV).eval(), dyn


@pytest.mark.parametrize("kind", ["lstm", "switching"])
def test_kvae_imputation_recipe_matches_reference(kind, monkeypatch):
    """tests/test_imputation_stability.py recipe, from the encoder sample `a` onwards: masked frames 4:10.
    lstm: alpha_t depends on the running prediction -> the stepwise (per-step launch) path is exercised."""
    g = _load_kvae(kind)
    kf, dyn = _kvae_kalman_block(kind, g)
    a, u, mask = g["a"].to(DEV), g["u"].to(DEV), g["mask"].to(DEV)
    if kind == "switching":
        noise = g["gumbel_noise"].to(DEV)                       # [T,B,K]: one draw per gumbel_softmax call of the reference
        dyn._draw_gumbel = lambda b, t, k, like: noise.permute(1, 0, 2).contiguous()
    with torch.no_grad():
        dyn.reset_state()
        outs = kf.smooth(a.clone(), u.clone(), mask)
    names = OUT_NAMES
    for k, o in zip(names, outs):   # the reference's fp32 outputs, with its own Kalman block in fp64 as the referee
        e32, e64, floor = check_close(f"kvae_{kind}.{k}", o, g[k], g["ref64_" + k])
        print(f"kvae_{kind}.{k}: e32 {e32:.1e} e64 {e64:.1e} floor {floor:.1e}")
    assert torch.allclose(dyn.state_seq.cpu(), g["state_probs"], atol=2e-6)
    a_imputed = (outs[8] @ outs[0]).squeeze(-1)                             # model.py:280-281
    a_filtered = (outs[8] @ outs[2]).squeeze(-1)                            # model.py:287-288
    assert torch.allclose(a_imputed.cpu(), g["a_imputed"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(a_filtered.cpu(), g["a_filtered"], rtol=1e-4, atol=1e-6)
    # the same two quantities without materialising A_list / B_list / C_list
    dyn.reset_state()
    ai, af, ms2, mf2 = kf.impute_observations(a.clone(), u.clone(), mask)
    assert torch.allclose(ai.cpu(), g["a_imputed"], rtol=1e-4, atol=1e-6)
    assert torch.allclose(af.cpu(), g["a_filtered"], rtol=1e-4, atol=1e-6)
    assert torch.equal(ms2, outs[0]) and torch.equal(mf2, outs[2])


def test_filter_step_and_smooth_step_match_reference_steps():
    """Per-step forms (kalman_filter.py:31-104, :204-237) with explicit per-sample matrices, against the oracle's
    step functions (same op sequence as the reference)."""
    from oracle import kalman_oracle as ko
    case, _, _, _ = load_golden("kalman_lstm")
    kf, dyn = make_kf(case)
    B, T, p = case["Y"].shape
    n, m = kf.n, kf.m
    gen = torch.Generator().manual_seed(3)
    A = torch.eye(n).expand(B, n, n) + 0.1 * torch.randn(B, n, n, generator=gen)
    Bm = 0.3 * torch.randn(B, n, m, generator=gen)
    C = 0.3 * torch.randn(B, p, n, generator=gen)
    L = torch.randn(B, n, n, generator=gen)
    Sigma = L @ L.mT + 0.5 * torch.eye(n)
    mu = torch.randn(B, n, generator=gen)
    y, u = torch.randn(B, p, generator=gen), torch.randn(B, m, generator=gen)
    mk = (torch.rand(B, generator=gen) > 0.3).float()
    Q, R = case["Q"], case["R"]
    ref = {dt: ko.filter_step(mu.to(dt).unsqueeze(-1), Sigma.to(dt), y.to(dt).unsqueeze(-1), u.to(dt).unsqueeze(-1), A.to(dt),
                              Bm.to(dt), C.to(dt), Q.to(dt), R.to(dt).expand(B, -1, -1), mk.to(dt))
           for dt in (torch.float32, torch.float64)}
    d = lambda x: x.to(DEV)
    with torch.no_grad():
        out = kf.filter_step(d(mu), d(Sigma), d(y), d(u), d(A), d(Bm), d(C), d(Q), mask_t=d(mk))
    assert len(out) == 7 and out[0].shape == (B, n, 1) and out[1].shape == (B, n, n)
    for i, name in enumerate(("mu_f", "Sigma_f", "mu_p", "Sigma_p")):
        check_close("filter_step." + name, out[i], ref[torch.float32][i], ref[torch.float64][i])
    # smooth_step: t -> t+1 quantities from a second filter step
    mu_f, Sig_f = ref[torch.float64][0], ref[torch.float64][1]
    A1 = torch.eye(n).expand(B, n, n) + 0.1 * torch.randn(B, n, n, generator=gen)
    Sig_p1 = A1.double() @ Sig_f @ A1.double().mT + Q.double()
    mu_p1 = A1.double() @ mu_f
    Sig_s1 = 0.7 * Sig_p1
    Sig_s1 = 0.5 * (Sig_s1 + Sig_s1.mT)
    mu_s1 = mu_p1 + 0.1
    def ref_smooth(dt):
        J = torch.linalg.solve(Sig_p1.to(dt).mT, (Sig_f.to(dt) @ A1.to(dt).mT).mT).mT
        m_ = mu_f.to(dt) + J @ (mu_s1.to(dt) - mu_p1.to(dt))
        S_ = Sig_f.to(dt) + J @ (Sig_s1.to(dt) - Sig_p1.to(dt)) @ J.mT
        return m_, 0.5 * (S_ + S_.mT)
    r32, r64 = ref_smooth(torch.float32), ref_smooth(torch.float64)
    with torch.no_grad():
        ms, Ss = kf.smooth_step(d(Sig_f.float()), d(Sig_p1.float()), d(Sig_s1.float()), d(mu_f.float()), d(mu_p1.float()),
                                d(mu_s1.float()), d(A1))
    assert ms.shape == (B, n, 1)
    check_close("smooth_step.mu", ms, r32[0], r64[0])
    check_close("smooth_step.Sigma", Ss, r32[1], r64[1])


def test_safe_cholesky_matches_reference_ladder():
    case = load_golden("kalman_lstm")[0]
    kf, _ = make_kf(case)
    S = torch.eye(4, device=DEV).expand(3, 4, 4).clone()
    S[1, 0, 0] = -1e-4          # needs a larger jitter than 1e-6
    L = kf._safe_cholesky(S)
    assert torch.isfinite(L).all() and torch.allclose(L[0], torch.linalg.cholesky(torch.eye(4, device=DEV) * (1 + 1e-3)), atol=1e-6)


def test_elbo_general_form_with_filtered_states_and_lists_from_filter():
    """kvae/kalman/test_filter.py usage: lists from filter(), states from a separate smooth() call; and the ELBO
    of arbitrary (here: filtered) states.  Values and gradients w.r.t. mu/Sigma/y/alpha/A/C against the oracle."""
    from oracle import kalman_oracle as ko
    case, _, _, _ = load_golden("kalman_lstm")
    kf, dyn = make_kf(case)
    Y, U, mask = case["Y"].to(DEV), case["U"].to(DEV), case["mask"].to(DEV)
    eps = case["eps"].to(DEV)
    kf._draw_eps = lambda B, T, n, like: eps
    mf, Sf, mp, Sp, A_list, B_list, C_list = kf.filter(Y, U, mask)      # with autograd: the lists carry gradients to alpha/A/C
    mu_in = mf.detach().clone().requires_grad_(True)
    Sig_in = Sf.detach().clone().requires_grad_(True)
    Yg = Y.clone().requires_grad_(True)
    val = kf.elbo(mu_in, Sig_in, Yg, U, A_list, B_list, C_list, mask=mask)
    gmu, gSig, gY, gal, gA, gC = torch.autograd.grad(val, [mu_in, Sig_in, Yg, dyn.alpha, dyn.A, dyn.C])
    # oracle: same ELBO on the same (fp64) inputs with autograd
    d = lambda k: case[k].double()
    mu64 = mf.detach().cpu().double().requires_grad_(True)
    Sg64 = Sf.detach().cpu().double().requires_grad_(True)
    Y64 = d("Y").clone().requires_grad_(True)
    al64, A64, C64 = d("alpha").clone().requires_grad_(True), d("A").clone().requires_grad_(True), d("C").clone().requires_grad_(True)
    A_seq, B_seq, C_seq, Q_seq = ko.mix(al64, A64, d("B"), C64, d("Q"), False, False)
    ref = ko.elbo(mu64, Sg64, Y64, d("U"), A_seq, B_seq, C_seq, Q_seq, d("R"), d("mu0"), d("Sigma0"), d("mask"), d("eps"))
    rg = torch.autograd.grad(ref, [mu64, Sg64, Y64, al64, A64, C64])
    rel = lambda a, b: float((a.detach().cpu().double() - b).norm() / b.norm())
    assert rel(val, ref) < 2e-6
    for name, got, want in zip(("dmu", "dSigma", "dY", "dalpha", "dA", "dC"), (gmu, gSig, gY, gal, gA, gC), rg):
        want = want.reshape(got.shape) if name != "dSigma" else 0.5 * (want + want.mT)   # kernel returns the symmetric gradient
        assert rel(got, want) < 2e-4, (name, rel(got, want))


def test_elbo_after_no_grad_smooth_treats_states_and_lists_as_constants():
    """Reference semantics: smooth() under torch.no_grad() returns constants, so a later elbo() differentiates
    only through what it is handed -- y_t (emission term) here -- and nothing reaches alpha / A / B / C."""
    from oracle import kalman_oracle as ko
    case, _, _, _ = load_golden("kalman_lstm")
    kf, dyn = make_kf(case)
    Y, U, mask = case["Y"].to(DEV), case["U"].to(DEV), case["mask"].to(DEV)
    eps = case["eps"].to(DEV)
    kf._draw_eps = lambda B, T, n, like: eps
    with torch.no_grad():
        outs = kf.smooth(Y, U, mask)
    Yg = Y.clone().requires_grad_(True)
    val = kf.elbo(outs[0], outs[1], Yg, U, outs[6], outs[7], outs[8], mask=mask)
    gY, gal, gA = torch.autograd.grad(val, [Yg, dyn.alpha, dyn.A], allow_unused=True)
    assert gal is None and gA is None
    d = lambda k: case[k].double()
    Y64 = d("Y").clone().requires_grad_(True)
    A_seq, B_seq, C_seq, Q_seq = ko.mix(d("alpha"), d("A"), d("B"), d("C"), d("Q"), False, False)
    ref = ko.elbo(outs[0].detach().cpu().double(), outs[1].detach().cpu().double(), Y64, d("U"), A_seq, B_seq, C_seq, Q_seq,
                  d("R"), d("mu0"), d("Sigma0"), d("mask"), d("eps"))
    (rY,) = torch.autograd.grad(ref, [Y64])
    rel = lambda a, b: float((a.detach().cpu().double() - b).norm() / b.norm())
    assert rel(val, ref) < 2e-6 and rel(gY, rY) < 2e-4, (rel(val, ref), rel(gY, rY))


@pytest.mark.parametrize("shape,hidden", [((4, 2, 4, 3), 50), ((8, 4, 8, 4), 50), ((4, 2, 4, 3), 17)])
def test_lstm_in_the_loop_kernel_matches_stepwise_path(shape, hidden):
    """lstm dynamics + missing observations: the fused launch (LSTM cell, head, softmax and y_for_dyn inside the filter
    kernel, kvae_kf_filter_lstm_fwd) against the per-step path (cuDNN LSTM step + one filter launch per time step),
    which is itself pinned to the reference by the kvae_lstm golden above.  Also: state carried across calls."""
    torch.manual_seed(3)
    B, T = 67, 33
    n, p, m, K = shape
    A = torch.eye(n).repeat(K, 1, 1) + 0.05 * torch.randn(K, n, n)
    Bm, C = 0.05 * torch.randn(K, n, m), 0.3 * torch.randn(K, p, n)
    dyn = DynamicsParameter(A, Bm, C, hidden_lstm=hidden)
    with torch.no_grad():
        dyn.head_w.bias.copy_(torch.randn(K))          # the default bias (0,-10,-10) would pin alpha to mode 0
        dyn.head_w.weight.mul_(3.0)
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, torch.zeros(n), 20.0 * torch.eye(n), dyn).to(DEV).eval()
    Y = torch.randn(B, T, p, device=DEV)
    U = 0.1 * torch.randn(B, T, m, device=DEV)
    mask = (torch.rand(B, T, device=DEV) < 0.6).float()
    outs = {}
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False             # cuDNN's LSTM would otherwise run its GEMMs in TF32 (1e-3 gates)
    for mode in ("fused", "stepwise"):
        if mode == "stepwise":
            kf._run_fused_lstm = lambda *a, **k: None
        with torch.no_grad():
            dyn.reset_state()
            o1 = kf.smooth(Y, U, mask)
            alpha1 = dyn.state_seq.clone()
            dyn.state_seq = []                          # second call continues from the carried LSTM state
            o2 = kf.filter(Y, U, mask)
            alpha2 = dyn.state_seq.clone()
        outs[mode] = (o1, alpha1, o2, alpha2, dyn.lstm_state[0].clone(), dyn.lstm_state[1].clone())
    torch.backends.cudnn.allow_tf32 = tf32
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    f, s = outs["fused"], outs["stepwise"]
    assert (f[1] - s[1]).abs().max() < 2e-5 and (f[3] - s[3]).abs().max() < 2e-5          # alpha
    assert float(f[1].std()) > 0.05                                                        # weights really vary
    for a, b in zip(f[0], s[0]):
        assert rel(a, b) < 2e-5
    for a, b in zip(f[2], s[2]):
        assert rel(a, b) < 2e-5
    assert rel(f[4], s[4]) < 2e-5 and rel(f[5], s[5]) < 2e-5                               # carried (h, c)
    miss = mask == 0                                                                        # mask handling stays bit-exact
    assert torch.equal(f[0][2][miss], f[0][4][miss])


# --------------------------------------------------------------------------------------------------------------------
# _safe_cholesky ladder THROUGH THE KERNELS (kalman_filter.py:282-302): the kernels report which family failed in the
# status word, KalmanFilter(check_info=True).elbo re-launches with the reference's next rung (10x jitter for the whole
# batch, separately for Sigma_smooth and Q; clamped-diagonal fallback after five failures).  Expected values: the oracle
# (same ladder, pinned to the live reference by tests/test_oracle.py::test_safe_cholesky_ladder_matches_reference).
# --------------------------------------------------------------------------------------------------------------------
def _ladder_case(min_eig):
    """kalman_lstm golden with ONE smoothed covariance pushed to smallest eigenvalue `min_eig`."""
    from oracle import kalman_oracle as ko
    case = load_golden("kalman_lstm")[0]
    g = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in case.items()}
    outs, Q_seq = ko.smooth(g["Y"], g["U"], g["mask"], g["alpha"], g["A"], g["B"], g["C"], g["Q"], g["R"], g["mu0"], g["Sigma0"],
                            bool(g["c_shared"]), bool(g["q_per_mode"]))
    ms, Ss = outs[0].squeeze(-1).clone(), outs[1].clone()
    S = 0.5 * (Ss[1, 3] + Ss[1, 3].T)
    w, V = torch.linalg.eigh(S.double())
    v = V[:, 0]
    Ss[1, 3] = (S.double() - (w[0] - min_eig) * torch.outer(v, v)).float()
    return g, outs, Q_seq, ms, Ss


@pytest.mark.parametrize("min_eig,rung", [(-5e-6, "1e-5"), (-1.0, "diag")])
def test_safe_cholesky_ladder_through_the_kernels(min_eig, rung):
    from oracle import kalman_oracle as ko
    g, outs, Q_seq, ms, Ss = _ladder_case(min_eig)
    ref = {}
    for dt in (torch.float32, torch.float64):
        c = lambda x: x.to(dt)
        m_, S_ = c(ms).requires_grad_(True), c(Ss).requires_grad_(True)
        val = ko.elbo(m_, S_, c(g["Y"]), c(g["U"]), c(outs[6]), c(outs[7]), c(outs[8]), c(Q_seq), c(g["R"]), c(g["mu0"]),
                      c(g["Sigma0"]), c(g["mask"]), c(g["eps"]))
        gm, gS = torch.autograd.grad(val, [m_, S_])
        ref[dt] = (val.detach(), gm, gS)
    kf, dyn = make_kf(g)
    kf.check_info = True
    d = lambda x: x.to(DEV)
    res = kf.smooth(d(g["Y"]), d(g["U"]), d(g["mask"]))
    mu_in, S_in = d(ms).requires_grad_(True), d(Ss).requires_grad_(True)
    kf._draw_eps = lambda B, T, n, like: d(g["eps"])
    val = kf.elbo(mu_in, S_in, d(g["Y"]), d(g["U"]), res[6], res[7], res[8], mask=d(g["mask"]))
    gm, gS = torch.autograd.grad(val, [mu_in, S_in])
    if rung == "diag":
        assert kf.last_chol["diag_smooth"] is True and kf.last_chol["diag_q"] is False
    else:
        assert abs(kf.last_chol["jitter_smooth"] - 1e-5) < 1e-12 and kf.last_chol["jitter_q"] == 1e-6, kf.last_chol
        assert kf.last_chol["diag_smooth"] is False
    check_close(f"ladder[{rung}].elbo", val, ref[torch.float32][0], ref[torch.float64][0])
    check_close(f"ladder[{rung}].dmu", gm, ref[torch.float32][1], ref[torch.float64][1], rtol=5e-5)
    check_close(f"ladder[{rung}].dSigma", gS, ref[torch.float32][2], ref[torch.float64][2], rtol=5e-5)


def test_safe_cholesky_q_ladder_fused_path():
    """Switching dynamics with a learnable Q_k that is not positive definite (smallest eigenvalue of the mixture -5e-6):
    the Q ladder climbs to 1e-5 while the Sigma_smooth ladder stays at 1e-6; fused value + adjoint launch."""
    from oracle import kalman_oracle as ko
    case = load_golden("kalman_switch")[0]
    g = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in case.items()}
    n = g["A"].shape[-1]
    v = torch.ones(n, dtype=torch.float64) / n ** 0.5                   # every Q_k gets the SAME eigenvector v with
    Pm = torch.eye(n, dtype=torch.float64) - torch.outer(v, v)          # eigenvalue -5e-6, so every mixture Q_t has it too
    g["Q"] = (Pm @ g["Q"].double() @ Pm - 5e-6 * torch.outer(v, v)).float()
    r32 = ko.run_case(g, torch.float32, want_grads=True)
    r64 = ko.run_case(g, torch.float64, want_grads=True)
    kf, dyn = make_kf(g)
    kf.check_info = True
    d = lambda x: x.to(DEV)
    Y = d(g["Y"]).requires_grad_(True)
    res = kf.smooth(Y, d(g["U"]), d(g["mask"]))
    kf._draw_eps = lambda B, T, n_, like: d(g["eps"])
    val = kf.elbo(res[0], res[1], Y, d(g["U"]), res[6], res[7], res[8], mask=d(g["mask"]))
    gY, gQ, gA = torch.autograd.grad(val, [Y, dyn.Q, dyn.A])
    assert kf.last_chol["jitter_q"] > 5e-6 and kf.last_chol["jitter_smooth"] == 1e-6, kf.last_chol
    check_close("qladder.elbo", val, r32["elbo"], r64["elbo"])
    check_close("qladder.dQ", gQ, r32["dQ"], r64["dQ"], rtol=5e-5)
    check_close("qladder.dA", gA, r32["dA"], r64["dA"], rtol=5e-5)
    check_close("qladder.dY", gY, r32["dY"], r64["dY"], rtol=5e-5)


def test_lazy_status_word_reports_one_call_late():
    """check_info='lazy' (the default): no host synchronisation in elbo(); a failed factorisation raises when the next
    call of the object starts."""
    g, outs, Q_seq, ms, Ss = _ladder_case(-1.0)
    kf, dyn = make_kf(g)
    assert kf.check_info == "lazy"
    d = lambda x: x.to(DEV)
    res = kf.smooth(d(g["Y"]), d(g["U"]), d(g["mask"]))
    kf._draw_eps = lambda B, T, n, like: d(g["eps"])
    kf.elbo(d(ms), d(Ss), d(g["Y"]), d(g["U"]), res[6], res[7], res[8], mask=d(g["mask"]))   # does not raise
    torch.cuda.synchronize()
    with pytest.raises(torch.linalg.LinAlgError):
        kf.smooth(d(g["Y"]), d(g["U"]), d(g["mask"]))


def test_lstm_mask_after_ones_mask_takes_the_in_kernel_lstm_path():
    """an all-ones mask followed by a FRESH mask with zeros of the same shape (possibly at the same
    address) must not be treated as all-ones: under no_grad the LSTM runs in the filter kernel for any mask."""
    z = np.load(os.path.join(GOLDEN, "kvae_lstm.npz"))
    K, n, p, m = int(z["K"]) if "K" in z.files else 3, 4, 2, 4
    torch.manual_seed(0)
    A = torch.eye(n).repeat(3, 1, 1) + 0.05 * torch.randn(3, n, n)
    dyn = DynamicsParameter(A, 0.05 * torch.randn(3, n, m), 0.3 * torch.randn(3, p, n)).to(DEV)
    with torch.no_grad():
        dyn.head_w.bias.zero_()
    kf = KalmanFilter(0.02 ** 0.5, 0.03 ** 0.5, torch.zeros(n), 20 * torch.eye(n), dyn).to(DEV)
    B, T = 64, 20
    Y = torch.randn(B, T, p, device=DEV) * 0.5
    U = torch.zeros(B, T, m, device=DEV)
    with torch.no_grad():
        dyn.reset_state()
        ones = torch.ones(B, T, device=DEV)
        o1 = kf.smooth(Y, U, ones)
        a1 = dyn.state_seq.clone()
        del ones
        dyn.reset_state()
        holes = torch.ones(B, T, device=DEV)
        holes[:, 4:16] = 0.0
        o2 = kf.smooth(Y, U, holes)
        a2 = dyn.state_seq.clone()
        # reference semantics for the masked call: the per-step path (LSTM cell + one filter launch per step)
        dyn.reset_state()
        kf.lanes = 2          # lanes != n disables the fused LSTM launch -> per-step fallback
        o3 = kf.smooth(Y, U, holes)
        a3 = dyn.state_seq.clone()
    assert not torch.allclose(a1[:, 6:], a2[:, 6:], atol=1e-4)          # the masked call did NOT reuse the batched alphas
    assert float((a2 - a3).abs().max()) < 5e-5                           # ... it fed C mu_pred to the LSTM at missing steps
    # (the per-step path runs its LSTM through cuDNN, TF32 GEMMs by default; the in-kernel cell is fp32)
    assert float((o2[0] - o3[0]).norm() / o3[0].norm()) < 5e-4


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch"])
def test_reference_shaped_calls_replay_from_a_cuda_graph(name, monkeypatch):
    """smooth() -> elbo() -> autograd.grad captured with torch.cuda.graph (the module makes no host reads under capture:
    the lazy status checks are skipped) and replayed on NEW input values copied into the static buffers: bit-identical
    to the eager calls on the same values (bench.py's `e2e` is this route)."""
    case, _, r32, r64 = load_golden(name)
    kf, dyn = make_kf(case)
    stat = {k: case[k].to(DEV).clone() for k in ("Y", "U", "mask", "eps")}
    monkeypatch.setattr(kf, "_draw_eps", lambda B, T, n, like: stat["eps"])

    def calls():
        Y = stat["Y"].requires_grad_(True)
        dyn.reset_state()
        outs = kf.smooth(Y, stat["U"], stat["mask"])
        val = kf.elbo(outs[0], outs[1], Y, stat["U"], outs[6], outs[7], outs[8], mask=stat["mask"])
        gr = torch.autograd.grad(val, [Y, dyn.alpha, dyn.A, dyn.B, dyn.C])
        return [val, outs[0], outs[1]] + list(gr)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        calls()                                   # warm-up outside the capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static_out = calls()
    stat["Y"].requires_grad_(False)
    gen = torch.Generator().manual_seed(5)
    for _ in range(2):                            # new values in the same buffers, then replay
        with torch.no_grad():
            stat["Y"].copy_(case["Y"] + 0.05 * torch.randn(case["Y"].shape, generator=gen))
            stat["eps"].copy_(torch.randn(case["eps"].shape, generator=gen))
        g.replay()
        got = [t.clone() for t in static_out]
        want = calls()
        stat["Y"].requires_grad_(False)
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b.detach())


@pytest.mark.parametrize("name,with_q", [("kalman_lstm", False), ("kalman_lstm", True), ("kalman_switch", False)])
def test_elbo_with_foreign_lists_and_explicit_q_list(name, with_q):
    """elbo() with list tensors that are NOT this object's outputs (here: perturbed copies) and with an explicit Q_list
    (kalman_filter.py:342-345): the general form (general_elbo.py), value and gradients w.r.t. the lists, mu, Sigma, y
    against the oracle's restatement of :305-401 in fp64."""
    from oracle import kalman_oracle as ko
    case, _, _, _ = load_golden(name)
    kf, dyn = make_kf(case)
    Y, U, mask = case["Y"].to(DEV), case["U"].to(DEV), case["mask"].to(DEV)
    eps = case["eps"].to(DEV)
    kf._draw_eps = lambda B, T, n, like: eps
    with torch.no_grad():
        dyn.reset_state()
        outs = kf.smooth(Y, U, mask)
    gen = torch.Generator().manual_seed(3)
    pert = lambda t: (t.detach().cpu() * (1.0 + 0.05 * torch.randn(t.shape, generator=gen))).contiguous()
    host = dict(mu=outs[0].detach().cpu(), Sig=outs[1].detach().cpu(), A=pert(outs[6]), B=pert(outs[7]), C=pert(outs[8].contiguous()))
    Bsz, T, n = host["mu"].shape[:3]
    Qh = None
    if with_q:
        W = 0.1 * torch.randn(Bsz, T, n, n, generator=gen)
        Qh = 0.02 * torch.eye(n) + W @ W.mT
    dev_in = {k: v.to(DEV).requires_grad_(True) for k, v in host.items()}
    Yg = Y.clone().requires_grad_(True)
    Qd = None if Qh is None else Qh.to(DEV).requires_grad_(True)
    val = kf.elbo(dev_in["mu"], dev_in["Sig"], Yg, U, dev_in["A"], dev_in["B"], dev_in["C"], Q_list=Qd, mask=mask)
    leaves = [dev_in[k] for k in ("mu", "Sig", "A", "B", "C")] + [Yg] + ([Qd] if with_q else [])
    got = torch.autograd.grad(val, leaves)
    # oracle, fp64
    d = lambda k: case[k].double()
    o = {k: v.double().requires_grad_(True) for k, v in host.items()}
    Y64 = d("Y").clone().requires_grad_(True)
    if with_q:
        Q64 = Qh.double().requires_grad_(True)
    elif case["q_per_mode"]:
        Q64 = torch.einsum("btk,kij->btij", d("alpha"), d("Q"))
    else:
        Q64 = d("Q").expand(Bsz, T, n, n)
    ref = ko.elbo(o["mu"], o["Sig"], Y64, d("U"), o["A"], o["B"], o["C"], Q64, d("R"), d("mu0"), d("Sigma0"), d("mask"), d("eps"))
    want = torch.autograd.grad(ref, [o[k] for k in ("mu", "Sig", "A", "B", "C")] + [Y64] + ([Q64] if with_q else []))
    rel = lambda a, b: float((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30))
    assert rel(val, ref) < 5e-6, rel(val, ref)
    for nm, a, b in zip(("dmu", "dSigma", "dA", "dB", "dC", "dY", "dQ"), got, want):
        assert rel(a, b.reshape(a.shape)) < 2e-4, (nm, rel(a, b.reshape(a.shape)))


def test_filter_step_and_smooth_step_under_autograd():
    """filter_step / smooth_step with tensors that require a gradient (general_steps.py): values equal the kernel forms
    (forward-only) and the gradients equal autograd of the oracle's restatement in fp64."""
    from oracle import kalman_oracle as ko
    case, _, r32, r64 = load_golden("kalman_lstm")
    kf, dyn = make_kf(case)
    B, T, n = case["Y"].shape[0], case["Y"].shape[1], kf.n
    t = 3
    d = lambda x: x.to(DEV).float()
    A_t, B_t, C_t = d(r32["A_list"][:, t]), d(r32["B_list"][:, t]), d(r32["C_list"][:, t])
    mu_prev, Sig_prev = d(r32["mus_filt"][:, t - 1]), d(r32["Sigmas_filt"][:, t - 1])
    y_t, u_t, m_t = d(case["Y"][:, t]), d(case["U"][:, t]), d(case["mask"][:, t])
    with torch.no_grad():
        k_out = kf.filter_step(mu_prev, Sig_prev, y_t, u_t, A_t, B_t, C_t, kf.Q, mask_t=m_t)
    leaves = [x.clone().requires_grad_(True) for x in (mu_prev, Sig_prev, y_t, A_t, C_t)]
    g_out = kf.filter_step(leaves[0], leaves[1], leaves[2], u_t, leaves[3], B_t, leaves[4], kf.Q, mask_t=m_t)
    for a, b in zip(g_out[:4], k_out[:4]):
        assert float((a - b).norm() / b.norm()) < 2e-6
    w = [torch.randn_like(o) for o in g_out[:4]]
    got = torch.autograd.grad(sum((wi * o).sum() for wi, o in zip(w, g_out[:4])), leaves)
    l64 = [x.detach().cpu().double().requires_grad_(True) for x in (mu_prev, Sig_prev, y_t, A_t, C_t)]
    o64 = ko.filter_step(l64[0], l64[1], l64[2].unsqueeze(-1), u_t.cpu().double().unsqueeze(-1), l64[3], B_t.cpu().double(),
                         l64[4], kf.Q.cpu().double(), kf.R.cpu().double().expand(B, -1, -1), m_t.cpu().double())
    want = torch.autograd.grad(sum((wi.cpu().double() * o).sum() for wi, o in zip(w, o64)), l64)
    rel = lambda a, b: float((a.detach().cpu().double() - b).norm() / b.norm().clamp_min(1e-30))
    for nm, a, b in zip(("dmu", "dSigma", "dy", "dA", "dC"), got, want):
        assert rel(a, b) < 2e-4, (nm, rel(a, b))
    # smooth_step
    Sf, Sp1, Ss1 = d(r32["Sigmas_filt"][:, t]), d(r32["Sigmas_pred"][:, t + 1]), d(r32["Sigmas_smooth"][:, t + 1])
    mf, mp1, ms1 = d(r32["mus_filt"][:, t]), d(r32["mus_pred"][:, t + 1]), d(r32["mus_smooth"][:, t + 1])
    A1 = d(r32["A_list"][:, t + 1])
    with torch.no_grad():
        km, kS = kf.smooth_step(Sf, Sp1, Ss1, mf, mp1, ms1, A1)
    Sfg = Sf.clone().requires_grad_(True)
    gm, gS = kf.smooth_step(Sfg, Sp1, Ss1, mf, mp1, ms1, A1)
    assert float((gm - km).norm() / km.norm()) < 2e-6 and float((gS - kS).norm() / kS.norm()) < 2e-6
    assert torch.autograd.grad(gS.sum() + gm.sum(), [Sfg])[0].isfinite().all()
    check_close("smooth_step.Sigma", gS, r32["Sigmas_smooth"][:, t], r64["Sigmas_smooth"][:, t])


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch"])
def test_float64_module_takes_the_torch_route(name):
    """The reference computes in the dtype of its buffers (float64 works there).  The kernels are fp32: a float64 module /
    float64 inputs take the step-by-step torch-op route (general_steps.py, general_elbo.py) and return float64 results that
    match the reference's fp64 goldens to 1e-9 (outputs, ELBO, gradients)."""
    case, cot, r32, r64 = load_golden(name)
    kf, dyn = make_kf(case)
    kf, dyn = kf.double(), dyn.double()
    Y = case["Y"].to(DEV).double().requires_grad_(True)
    U, mask, eps = case["U"].to(DEV).double(), case["mask"].to(DEV).double(), case["eps"].to(DEV).double()
    dyn.reset_state()
    outs = kf.smooth(Y, U, mask)
    rel = lambda a, b: float((a.detach().cpu().double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
    for k, o in zip(OUT_NAMES, outs):
        assert o.dtype == torch.float64 and rel(o, r64[k]) < 1e-9, (k, rel(o, r64[k]))
    kf._draw_eps = lambda B, T, n, like: eps
    val = kf.elbo(outs[0], outs[1], Y, U, outs[6], outs[7], outs[8], mask=mask)
    assert val.dtype == torch.float64 and rel(val, r64["elbo"]) < 1e-9
    loss = val
    if cot:   # the goldens' gradients are those of elbo + sum <cotangent, output>
        for k, o in zip(OUT_NAMES, outs):
            if k in cot:
                loss = loss + (cot[k].to(DEV).double() * o).sum()
    gY, gA = torch.autograd.grad(loss, [Y, dyn.A])
    assert rel(gY, r64["dY"]) < 1e-7 and rel(gA, r64["dA"]) < 1e-7, (rel(gY, r64["dY"]), rel(gA, r64["dA"]))

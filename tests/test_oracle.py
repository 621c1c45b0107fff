"""CPU tests: the oracle restatement is pinned against golden vectors produced by the unmodified
reference (and against the live reference where /root/reference exists); the explicit adjoint is
pinned against autograd of the reference's op sequence."""
import pytest
import torch

from kalman_vae_b200.synthetic import Shape, make_case
from oracle import adjoint, kalman_oracle as ko, ref_shim
from tests._util import GRAD_NAMES, OUT_NAMES, golden_names, load_golden, rel


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden_fp32(name):
    case, cot, r32, _ = load_golden(name)
    T = case["Y"].shape[1]
    want_grads = "dY" in r32
    got = ko.run_case(case, torch.float32, want_grads=want_grads, cotangents=cot, with_elbo=T > 1)
    for k in OUT_NAMES:
        # same op sequence as the reference -> forward values agree to the last few ulps
        assert rel(got[k], r32[k]) <= 2e-7, (k, rel(got[k], r32[k]))
    if "elbo" in r32:
        assert rel(got["elbo"], r32["elbo"]) <= 1e-6
    for k in GRAD_NAMES:
        if k in r32:
            assert rel(got[k], r32[k]) <= 2e-4, (k, rel(got[k], r32[k]))   # fp32 autograd noise


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden_fp64(name):
    case, cot, _, r64 = load_golden(name)
    T = case["Y"].shape[1]
    want_grads = "dY" in r64
    got = ko.run_case(case, torch.float64, want_grads=want_grads, cotangents=cot, with_elbo=T > 1)
    for k in OUT_NAMES + ["elbo"] + GRAD_NAMES:
        if k in r64:
            assert rel(got[k], r64[k]) <= 1e-11, (k, rel(got[k], r64[k]))


@pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")
@pytest.mark.parametrize("variant", ["lstm", "switching"])
def test_oracle_matches_live_reference(variant):
    sw = variant == "switching"
    case = make_case(Shape(3, 8, 4, 2, 4, 3, sw, sw), seed=21, mask_kind="bernoulli", zero_u=False, c_std=0.3, nonsym_q=sw)
    ref = ref_shim.run_reference_case(case, torch.float64)
    got = ko.run_case(case, torch.float64)
    for k in ref:
        assert rel(got[k], ref[k]) <= 1e-11, k


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch", "kalman_n16", "kalman_fractional", "kalman_rocket"])
def test_explicit_adjoint_matches_autograd_fp64(name):
    case, cot, _, r64 = load_golden(name)
    got = adjoint.smooth_elbo_backward(case, r64, 1.0, cot, dtype=torch.float64)
    for k in GRAD_NAMES:
        if k in r64:
            assert rel(got[k], r64[k]) <= 1e-9, (k, rel(got[k], r64[k]))


def test_mask_zero_is_bit_exact_prediction():
    """mask = 0 -> mu_filt == mu_pred and Sigma_filt == sym(Sigma_pred), bit-exact."""
    case, _, r32, _ = load_golden("kalman_zero_mask")
    assert torch.equal(r32["mus_filt"], r32["mus_pred"])
    assert torch.equal(r32["Sigmas_filt"], 0.5 * (r32["Sigmas_pred"] + r32["Sigmas_pred"].mT))
    got = ko.run_case(case, torch.float32, want_grads=False)
    assert torch.equal(got["mus_filt"], got["mus_pred"])


def test_steady_state_riccati_known_answer():
    """Analytic check: scalar state, K=1, time-invariant -> Sigma_pred converges to the positive root of
    the discrete Riccati equation  P = a^2 P r/(c^2 P + r) + q."""
    a, c, q, r = 0.9, 1.3, 0.02, 0.03
    n = 2  # smallest instantiated state dim: two decoupled copies of the scalar system
    case = dict(A=a * torch.eye(n).unsqueeze(0), B=torch.zeros(1, n, 1), C=torch.tensor([[[c, 0.0]]]),
                Q=q * torch.eye(n).unsqueeze(0), R=torch.tensor([[r]]), mu0=torch.zeros(n), Sigma0=torch.eye(n),
                Y=torch.zeros(1, 200, 1), U=torch.zeros(1, 200, 1), mask=torch.ones(1, 200), alpha=torch.ones(1, 200, 1),
                eps=torch.zeros(1, 200, n), q_per_mode=True, c_shared=True)
    out = ko.run_case(case, torch.float64, want_grads=False)
    P = out["Sigmas_pred"][0, -1, 0, 0].item()
    a, c, q, r = (float(torch.tensor(v, dtype=torch.float32)) for v in (a, c, q, r))   # inputs are stored in fp32
    assert abs(P - (a * a * P * r / (c * c * P + r) + q)) < 1e-12


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
@pytest.mark.parametrize("min_eig", [1e-3, -5e-6, -2e-4, -1.0])
def test_safe_cholesky_ladder_matches_reference(min_eig):
    """The oracle's ladder (which the GPU ladder tests are checked against) against the live reference's
    KalmanFilter._safe_cholesky (kalman_filter.py:282-302): first rung, a higher rung, the clamped-diagonal fallback."""
    ns = ref_shim.load()
    torch.manual_seed(3)
    X = torch.randn(5, 7, 4, 4, dtype=torch.float64)
    S = X @ X.mT + 0.5 * torch.eye(4, dtype=torch.float64)
    w, V = torch.linalg.eigh(S[2, 3])
    S[2, 3] = S[2, 3] - (w[0] - min_eig) * torch.outer(V[:, 0], V[:, 0])
    S = S + 1e-3 * torch.randn(5, 7, 4, 4, dtype=torch.float64)          # not exactly symmetric, as Sigma_pred is
    kf = ns.KalmanFilter.__new__(ns.KalmanFilter)                         # the method uses no instance state
    for dt in (torch.float64, torch.float32):
        L_ref = ns.KalmanFilter._safe_cholesky(kf, S.to(dt))
        L = ko.safe_cholesky(S.to(dt))
        assert torch.equal(L, L_ref), (min_eig, dt)

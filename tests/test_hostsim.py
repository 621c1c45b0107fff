"""CPU tests of the DEVICE ARITHMETIC: kalman_vae_b200/csrc/*.cuh compiled for the host with one lane
per sequence (tests/hostsim, test tooling only) and compared with the oracle / the reference goldens.
Both the register path (L=1) and the shared-memory 'publish' path (forced) are exercised, so tile
life-time bugs show up here, before any GPU time is spent."""
import pytest
import torch

from tests._util import GRAD_NAMES, OUT_NAMES, check_close, load_golden, rel
from tests.hostsim import driver

CASES = ["kalman_lstm", "kalman_switch", "kalman_fractional", "kalman_zero_mask", "kalman_T1", "kalman_n8", "kalman_rocket"]


@pytest.fixture(scope="module", autouse=True)
def _build():
    driver.build()


@pytest.mark.parametrize("force_mem", [False, True])
@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name, force_mem):
    case, _, r32, r64 = load_golden(name)
    got = driver.fwd(case, smooth=True, force_mem=force_mem)
    assert int(got["info"]) == 0
    for k in OUT_NAMES:
        check_close(f"{name}.{k}", got[k], r32[k], r64[k])


def test_mask_zero_bit_exact():
    case, _, r32, _ = load_golden("kalman_zero_mask")
    got = driver.fwd(case, smooth=True)
    assert torch.equal(got["mus_filt"], got["mus_pred"])
    assert torch.equal(got["Sigmas_filt"], 0.5 * (got["Sigmas_pred"] + got["Sigmas_pred"].mT))


@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch", "kalman_rocket", "kalman_n8"])
def test_elbo_matches_reference(name):
    case, _, r32, r64 = load_golden(name)
    got = driver.elbo(case, r32)
    assert got["info"] == 0
    check_close(name + ".elbo", torch.tensor(got["elbo"]), r32["elbo"], r64["elbo"])


@pytest.mark.parametrize("force_mem", [False, True])
@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch", "kalman_fractional", "kalman_rocket"])
def test_adjoint_matches_reference_autograd(name, force_mem):
    case, cot, r32, r64 = load_golden(name)
    got = driver.bwd(case, r32, 1.0, cot, force_mem=force_mem)
    assert got["info"] == 0
    for k in GRAD_NAMES:
        if k in r32:
            check_close(f"{name}.{k}", got[k], r32[k], r64[k])


@pytest.mark.parametrize("force_mem", [False, True])
@pytest.mark.parametrize("name", ["kalman_lstm", "kalman_switch", "kalman_fractional", "kalman_rocket", "kalman_n8"])
def test_fused_elbo_value_in_adjoint_sweep(name, force_mem):
    """KVAE_FLAG_WITH_ELBO: the adjoint sweep accumulates the ELBO value itself; it must equal the ELBO kernel's
    sums term by term and the reference's value, and leave the gradients unchanged."""
    case, cot, r32, r64 = load_golden(name)
    sep = driver.elbo(case, r32, force_mem=force_mem)
    got = driver.bwd(case, r32, 1.0, None, force_mem=force_mem, with_elbo=True)
    plain = driver.bwd(case, r32, 1.0, None, force_mem=force_mem)
    assert got["info"] == 0
    for i, k in enumerate(("trans", "emiss", "init", "entropy")):
        assert abs(got["elbo_terms"][i] - sep[k]) <= 2e-6 * max(1.0, abs(sep[k])), (k, got["elbo_terms"][i], sep[k])
    assert abs(got["elbo_terms"][4] - float(case["mask"].double().sum())) <= 1e-9 * case["mask"].numel()
    check_close(name + ".elbo", torch.tensor(got["elbo"]), r32["elbo"], r64["elbo"])
    for k in ("dY", "dalpha"):
        assert torch.equal(got[k], plain[k])

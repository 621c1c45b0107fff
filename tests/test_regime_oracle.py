"""CPU: the regime-sampler oracle (oracle/regime_oracle.py) against outputs of the UNMODIFIED reference
(tests/golden/regime_*.npz, written by oracle/make_golden_regime.py), and the host-side argument checks of the C ABI."""
import pytest
import torch

from oracle import regime_oracle as ro
from tests._util import golden_names, load_golden, rel

NAMES = golden_names("regime_")


def test_goldens_present():
    assert set(NAMES) >= {"regime_k3_soft", "regime_k3_hard", "regime_k8_soft", "regime_k2_T1"}


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_reference(name):
    case, cot, r32, r64 = load_golden(name)
    got32 = ro.regime_sample_with_grads(case, torch.float32)
    got64 = ro.regime_sample_with_grads(case, torch.float64)
    for k in ("y_seq", "log_q", "log_p", "d_logits", "d_init"):
        if k in ("y_seq", "log_q", "log_p"):
            assert torch.equal(got32[k], r32[k]), (name, k)      # same forward op sequence -> same bits
        else:                                                    # autograd accumulates in another order (stack vs slice writes)
            assert rel(got32[k], r32[k]) < 2e-6, (name, k)
        assert rel(got64[k], r64[k]) < 1e-14, (name, k)
    if case["hard"]:
        y = r64["y_seq"]
        assert torch.all((y == 0) | ((y - 1).abs() < 1e-12))      # straight-through samples are one-hot
        assert torch.all((y.sum(-1) - 1).abs() < 1e-12)


def test_abi_rejects_bad_arguments_without_a_gpu():
    from kalman_vae_b200 import capi
    L = capi.lib()
    assert L.kvae_regime_supported(3) == 1 and L.kvae_regime_supported(8) == 1
    assert L.kvae_regime_supported(1) == 0 and L.kvae_regime_supported(9) == 0
    rc = L.kvae_regime_sample_fwd(None, None, None, None, None, None, None, None, 0, None)
    assert rc < 0 and b"null" in L.kvae_regime_last_error()

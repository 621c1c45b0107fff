"""GPU parity of the SKVAE regime sampler kernels (csrc/kvae_regime.cu) against the reference's own
outputs (tests/golden/regime_*.npz): forward chain, explicit adjoint, and the drop-in SwitchingDynamicsParameter."""
import pytest
import torch

from kalman_vae_b200.functional import RegimeSampleFunction
from tests._util import check_close, golden_names, load_golden

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("name", golden_names("regime_"))
def test_kernels_match_reference(name):
    case, cot, r32, r64 = load_golden(name)
    logits = case["logits"].to(DEV).requires_grad_(True)
    init = case["init_logits"].to(DEV).requires_grad_(True)
    y, lq, lp = RegimeSampleFunction.apply(logits, init, case["gumbel"].to(DEV), case["trans"].to(DEV), float(case["tau"][0]),
                                           bool(case["hard"]))
    for k, v in (("y_seq", y), ("log_q", lq), ("log_p", lp)):
        check_close(f"{name}.{k}", v, r32[k], r64[k])
    loss = (case["cot_y"].to(DEV) * y).sum() + (case["cot_q"].to(DEV) * lq).sum() + (case["cot_p"].to(DEV) * lp).sum()
    d_logits, d_init = torch.autograd.grad(loss, [logits, init])
    check_close(f"{name}.d_init", d_init, r32["d_init"], r64["d_init"])
    if r64["d_logits"].abs().max() > 0:
        check_close(f"{name}.d_logits", d_logits, r32["d_logits"], r64["d_logits"])
    else:
        assert float(d_logits.abs().max()) == 0.0                   # T = 1: logits are never read
    assert float(d_logits[:, 0].abs().max()) == 0.0                 # slice t = 0 is never read (switch_dyn_param.py:67)
    if case["hard"]:
        assert torch.all((y == 0) | ((y - 1).abs() <= 2e-7))
        assert torch.equal(y.argmax(-1).cpu(), r32["y_seq"].argmax(-1))   # same regimes picked: index work is exact


def test_dropin_switching_dynamics_uses_the_kernel_and_matches():
    """SwitchingDynamicsParameter.compute_weights (mirror of compute_batch) with its bi-GRU replaced by fixed logits."""
    from kalman_vae_b200 import SwitchingDynamicsParameter
    from kalman_vae_b200.dyn_param import StickyRegimePrior
    case, cot, r32, r64 = load_golden("regime_k3_soft")
    B, T, K, _ = case["logits"].shape
    n, m, p = 4, 4, 2

    class Fixed(torch.nn.Module):
        def forward(self, a_seq):
            return case["logits"].to(DEV), case["init_logits"].to(DEV)

    dyn = SwitchingDynamicsParameter(torch.zeros(K, n, n), torch.zeros(K, n, m), torch.zeros(K, p, n),
                                     prior=StickyRegimePrior(K, p_stay=float(case["p_stay"][0])),
                                     markov_regime_posterior=Fixed()).to(DEV)
    dyn._draw_gumbel = lambda b, t, k, like: case["gumbel"].to(DEV)
    y = dyn.compute_weights(torch.zeros(B, T, p, device=DEV), is_training=True)
    lq, lp = dyn.elbo_terms()
    check_close("dyn.state_seq", y, r32["y_seq"], r64["y_seq"])
    check_close("dyn.log_qseq", lq, r32["log_q"], r64["log_q"])
    check_close("dyn.log_pseq", lp, r32["log_p"], r64["log_p"])
    with pytest.raises(Exception):
        dyn.compute_weights(torch.zeros(B, T, p), is_training=True)   # CPU tensors: no CPU path

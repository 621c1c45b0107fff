"""CPU: the VAE-loss oracle against the live reference's kvae.vae.losses.vae_loss (both output distributions, masks)."""
import pytest
import torch

from oracle import ref_shim, vae_loss_oracle as vo


def _inputs(seed, dtype):
    g = torch.Generator().manual_seed(seed)
    B, T = 3, 5
    x = (torch.rand(B, T, 1, 8, 8, generator=g) > 0.7).to(dtype)
    x_mu = torch.randn(B, T, 1, 8, 8, generator=g, dtype=torch.float64).to(dtype)
    a = torch.randn(B, T, 2, generator=g, dtype=torch.float64).to(dtype)
    a_mu = torch.randn(B, T, 2, generator=g, dtype=torch.float64).to(dtype)
    a_var = (0.05 + torch.rand(B, T, 2, generator=g, dtype=torch.float64)).to(dtype)
    mask = (torch.rand(B, T, generator=g) > 0.3).to(dtype)
    return x, x_mu, a, a_mu, a_var, mask


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
@pytest.mark.parametrize("distr", ["bernoulli", "gaussian"])
@pytest.mark.parametrize("masked", [False, True])
def test_vae_loss_oracle_matches_reference(distr, masked):
    ref_shim.load()
    from kvae.vae.losses import vae_loss as ref_vae_loss
    for dtype in (torch.float64, torch.float32):
        x, x_mu, a, a_mu, a_var, mask = _inputs(5, dtype)
        xv = torch.tensor(0.1, dtype=dtype)
        kw = dict(scale_reconstruction=0.3, beta=0.7, mask=mask if masked else None, out_distr=distr)
        leaves = [t.clone().requires_grad_(True) for t in (x_mu, a, a_mu, a_var)]
        ref = ref_vae_loss(x, leaves[0], xv, leaves[1], leaves[2], leaves[3], **kw)
        gref = torch.autograd.grad(ref[0] + 0.5 * ref[1] - 0.25 * ref[2], leaves)
        mine_l = [t.clone().requires_grad_(True) for t in (x_mu, a, a_mu, a_var)]
        got = vo.vae_loss(x, mine_l[0], xv, mine_l[1], mine_l[2], mine_l[3], **kw)
        ggot = torch.autograd.grad(got[0] + 0.5 * got[1] - 0.25 * got[2], mine_l)
        tol = 1e-12 if dtype == torch.float64 else 2e-6
        for r, g_ in zip(ref, got):
            assert abs(float(r) - float(g_)) <= tol * max(1.0, abs(float(r)))
        for r, g_ in zip(gref, ggot):
            assert float((r - g_).norm()) <= tol * max(1.0, float(r.norm()))

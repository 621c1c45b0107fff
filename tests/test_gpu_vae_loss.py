"""GPU: fused VAE-side reductions against the oracle restatement of kvae/vae/losses.py:62-111 (fp32 and
fp64), value and gradients, both output distributions, with and without a mask; reparameterisation (model.py:81-84)."""
import pytest
import torch

from kalman_vae_b200.vae_loss import reparameterize, vae_loss
from oracle import vae_loss_oracle as vo
from tests._util import check_close

pytestmark = pytest.mark.gpu


def _inputs(seed, B=32, T=20, H=32):
    g = torch.Generator().manual_seed(seed)
    x = (torch.rand(B, T, 1, H, H, generator=g) > 0.9).float()
    x_mu = 2.0 * torch.randn(B, T, 1, H, H, generator=g)
    a = torch.randn(B, T, 2, generator=g)
    a_mu = torch.randn(B, T, 2, generator=g)
    a_var = 0.05 + torch.rand(B, T, 2, generator=g)
    mask = (torch.rand(B, T, generator=g) > 0.3).float()
    return x, x_mu, a, a_mu, a_var, mask


@pytest.mark.parametrize("distr", ["bernoulli", "gaussian"])
@pytest.mark.parametrize("masked", [False, True])
def test_vae_loss_matches_oracle(distr, masked):
    dev = torch.device("cuda:0")
    x, x_mu, a, a_mu, a_var, mask = _inputs(3)
    kw = dict(scale_reconstruction=0.3, beta=0.7, out_distr=distr)
    ref = {}
    for dt in (torch.float32, torch.float64):
        leaves = [t.to(dt).requires_grad_(True) for t in (x_mu, a, a_mu, a_var)]
        out = vo.vae_loss(x.to(dt), leaves[0], torch.tensor(0.1, dtype=dt), leaves[1], leaves[2], leaves[3],
                          mask=mask.to(dt) if masked else None, **kw)
        grads = torch.autograd.grad(out[0] + 0.5 * out[1] - 0.25 * out[2], leaves)
        ref[dt] = ([o.detach() for o in out], grads)
    leaves = [t.to(dev).requires_grad_(True) for t in (x_mu, a, a_mu, a_var)]
    out = vae_loss(x.to(dev), leaves[0], torch.tensor(0.1, device=dev), leaves[1], leaves[2], leaves[3],
                   mask=mask.to(dev) if masked else None, **kw)
    grads = torch.autograd.grad(out[0] + 0.5 * out[1] - 0.25 * out[2], leaves)
    for i, nm in enumerate(("vae_elbo", "recon", "reg")):
        check_close(f"{distr}.{nm}", out[i], ref[torch.float32][0][i], ref[torch.float64][0][i])
    for i, nm in enumerate(("d_x_mu", "d_a", "d_a_mu", "d_a_var")):
        check_close(f"{distr}.{nm}", grads[i], ref[torch.float32][1][i], ref[torch.float64][1][i])


def test_reparameterize_matches_model():
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    mu, var, eps = torch.randn(640, 2, generator=g), 0.01 + torch.rand(640, 2, generator=g), torch.randn(640, 2, generator=g)
    m, v = mu.to(dev).requires_grad_(True), var.to(dev).requires_grad_(True)
    a = reparameterize(m, v, eps.to(dev))
    w = torch.randn(640, 2, generator=g)
    gm, gv = torch.autograd.grad((a * w.to(dev)).sum(), [m, v])
    m64, v64 = mu.double().requires_grad_(True), var.double().requires_grad_(True)
    a64 = vo.reparameterize(m64, v64, eps.double())
    gm64, gv64 = torch.autograd.grad((a64 * w.double()).sum(), [m64, v64])
    assert float((a.cpu().double() - a64).norm() / a64.norm()) < 1e-6
    assert float((gm.cpu().double() - gm64).norm() / gm64.norm()) < 1e-6
    assert float((gv.cpu().double() - gv64).norm() / gv64.norm()) < 1e-6


def test_vae_loss_per_pixel_variance_takes_the_torch_route():
    """A tensor-valued x_var (one variance per pixel) is outside the reduction kernel: torch ops under autograd, against the
    oracle restatement in fp64."""
    dev = torch.device("cuda:0")
    x, x_mu, a, a_mu, a_var, mask = _inputs(5, B=4, T=6)
    g = torch.Generator().manual_seed(9)
    x_var = 0.05 + torch.rand(x.shape, generator=g)
    kw = dict(scale_reconstruction=0.3, beta=0.7, out_distr="gaussian")
    l64 = [t.double().requires_grad_(True) for t in (x_mu, a, x_var)]
    want = vo.vae_loss(x.double(), l64[0], l64[2], l64[1], a_mu.double(), a_var.double(), mask=mask.double(), **kw)
    gw = torch.autograd.grad(want[0], l64)
    lv = [t.to(dev).requires_grad_(True) for t in (x_mu, a, x_var)]
    got = vae_loss(x.to(dev), lv[0], lv[2], lv[1], a_mu.to(dev), a_var.to(dev), mask=mask.to(dev), **kw)
    gg = torch.autograd.grad(got[0], lv)
    rel = lambda p, q: float((p.detach().cpu().double() - q).norm() / q.norm().clamp_min(1e-30))
    for p, q in zip(list(got) + list(gg), list(want) + list(gw)):
        assert rel(p, q) < 1e-5, rel(p, q)

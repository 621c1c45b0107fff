"""TEST INFRASTRUCTURE — CPU restatement (own code, torch ops) of the reference's VAE loss and reparameterisation:
kvae/vae/losses.py:5-17 (log_gaussian), :62-111 (vae_loss) and kvae/model/model.py:81-84.  Pinned to the live reference by
tests/test_vae_oracle.py.  Only tests/ may import this module."""
import math

import torch
import torch.nn.functional as F


def log_gaussian(x, mean, var):                                           # losses.py:5-17
    return -0.5 * math.log(2.0 * math.pi) - torch.log(var) / 2 - torch.square(x - mean) / (2 * var)


def vae_loss(x, x_mu, x_var, a, a_mu, a_var, scale_reconstruction=0.3, beta=1.0, mask=None, out_distr="gaussian"):
    B, T = x.shape[:2]
    m = torch.ones(B, T, dtype=x.dtype, device=x.device) if mask is None else mask.to(x.dtype).view(B, T)   # :74-80
    denom = m.sum().clamp(min=1.0)                                                                           # :81
    if out_distr.lower() == "bernoulli":
        log_px = -F.binary_cross_entropy_with_logits(x_mu, x, reduction="none").sum(dim=(2, 3, 4))           # :83-85
    else:
        log_px = log_gaussian(x, x_mu, x_var).sum(dim=(2, 3, 4))                                             # :46
    log_q = log_gaussian(a, a_mu, a_var).sum(dim=-1)
    log_p = log_gaussian(a, torch.zeros_like(a), torch.ones_like(a)).sum(dim=-1)                             # :96-99
    recon = (log_px * m).sum() / denom
    reg = ((log_p * m).sum() - (log_q * m).sum()) / denom                                                    # :103-105
    return scale_reconstruction * recon + beta * reg, recon, reg                                             # :107-109


def reparameterize(mu, var, eps):                                                                            # model.py:81-84
    return mu + eps * torch.sqrt(var + 1e-6)

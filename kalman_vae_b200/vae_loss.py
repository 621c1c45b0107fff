"""VAE-side reductions of the KVAE loss as fused CUDA launches.

`vae_loss(...)` has the signature and return values of the reference's `kvae.vae.losses.vae_loss`
(kvae/vae/losses.py:62-111): one pass over the frames for the value (masked pixel log-likelihood, log q(a|x), log p(a),
normaliser clamp(sum(mask), 1)) and one elementwise pass for the gradient, instead of ~25 ATen ops and their autograd
replay.  `reparameterize(mu, var)` is kvae/model/model.py:81-84.  CUDA only; no fallback.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, byref, c_float, c_int, c_int32, c_long, c_size_t, c_void_p

import torch

from . import capi
from .functional import prep, workspace


class KvaeVaeDims(Structure):
    _fields_ = [("frames", c_int32), ("pixels", c_int32), ("a_dim", c_int32), ("bernoulli", c_int32),
                ("x_var", c_float), ("scale_reconstruction", c_float), ("beta", c_float)]


_bound = False


def _lib():
    global _bound
    L = capi.lib()
    if not _bound:
        L.kvae_vae_last_error.restype = ctypes.c_char_p
        L.kvae_vae_loss_workspace_bytes.argtypes = [POINTER(KvaeVaeDims)]
        L.kvae_vae_loss_workspace_bytes.restype = c_size_t
        L.kvae_vae_loss_fwd.argtypes = [POINTER(KvaeVaeDims)] + [c_void_p] * 8 + [c_int, c_void_p]
        L.kvae_vae_loss_bwd.argtypes = [POINTER(KvaeVaeDims)] + [c_void_p] * 12 + [c_int, c_void_p]
        L.kvae_vae_reparam_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p]
        L.kvae_vae_reparam_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p]
        _bound = True
    return L


def _check(rc, what):
    if rc != 0:
        raise capi.KvaeError(f"{what} failed (status {rc}): {_lib().kvae_vae_last_error().decode()}")


class _VaeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, x_mu, a, a_mu, a_var, mask, dims):
        dev = x.device
        xs, ls = prep(x).reshape(dims.frames, dims.pixels), prep(x_mu).reshape(dims.frames, dims.pixels)
        av, am, avr = prep(a), prep(a_mu), prep(a_var)
        mk = None if mask is None else prep(mask).reshape(-1)
        out = torch.empty(8, dtype=torch.float32, device=dev)
        L = _lib()
        ws = workspace(dev, "vae", int(L.kvae_vae_loss_workspace_bytes(byref(dims))))
        _check(L.kvae_vae_loss_fwd(byref(dims), capi._ptr(xs, "x"), capi._ptr(ls, "x_mu"), capi._ptr(av, "a"), capi._ptr(am, "a_mu"),
                                   capi._ptr(avr, "a_var"), capi._ptr(mk, "mask"), capi._ptr(out, "out"), ws.data_ptr(),
                                   dev.index, capi._stream(dev)), "kvae_vae_loss_fwd")
        ctx.save_for_backward(xs, ls, av, am, avr, out) if mk is None else ctx.save_for_backward(xs, ls, av, am, avr, out, mk)
        ctx.dims, ctx.shapes = dims, (x_mu.shape, a.shape)
        return out[0].clone(), out[1].clone(), out[2].clone()

    @staticmethod
    def backward(ctx, g_elbo, g_recon, g_reg):
        saved = ctx.saved_tensors
        xs, ls, av, am, avr, out = saved[:6]
        mk = saved[6] if len(saved) > 6 else None
        dev = xs.device
        z = lambda g: torch.zeros((), device=dev) if g is None else g.detach().to(torch.float32).reshape(())
        g3 = torch.stack([z(g_elbo), z(g_recon), z(g_reg)]).contiguous()
        d_l, d_a, d_am, d_av = torch.empty_like(ls), torch.empty_like(av), torch.empty_like(am), torch.empty_like(avr)
        dims = ctx.dims
        _check(_lib().kvae_vae_loss_bwd(byref(dims), capi._ptr(xs, "x"), capi._ptr(ls, "x_mu"), capi._ptr(av, "a"), capi._ptr(am, "a_mu"),
                                        capi._ptr(avr, "a_var"), capi._ptr(mk, "mask"), capi._ptr(g3, "g"), capi._ptr(out, "out"),
                                        capi._ptr(d_l, "d_x_mu"), capi._ptr(d_a, "d_a"), capi._ptr(d_am, "d_a_mu"),
                                        capi._ptr(d_av, "d_a_var"), dev.index, capi._stream(dev)), "kvae_vae_loss_bwd")
        return None, d_l.view(ctx.shapes[0]), d_a.view(ctx.shapes[1]), d_am.view(ctx.shapes[1]), d_av.view(ctx.shapes[1]), None, None


def vae_loss(x, x_mu, x_var, a, a_mu, a_var, scale_reconstruction: float = 0.3, beta: float = 1.0, mask=None,
             out_distr: str = "gaussian"):
    """kvae/vae/losses.py:62-111 -> (vae_elbo, recon_term, regularization_term).  x_var: a scalar (float or 0-dim tensor,
    as KVAE.compute_loss passes it, model.py:205); per-pixel variances are not supported."""
    if not x.is_cuda:
        raise capi.KvaeError("vae_loss (B200-native) needs CUDA tensors; there is no CPU path")
    B, T = x.shape[:2]
    pixels = 1
    for s in x.shape[2:]:
        pixels *= int(s)
    xv = float(x_var) if not torch.is_tensor(x_var) else (float(x_var.item()) if x_var.numel() == 1 else None)
    if xv is None:
        return _vae_loss_per_pixel_variance(x, x_mu, x_var, a, a_mu, a_var, scale_reconstruction, beta, mask, out_distr)
    m = None
    if mask is not None:
        m = mask.to(device=x.device, dtype=torch.float32)
        if m.shape != (B, T):
            m = m.view(B, T)                                              # losses.py:78-80
    dims = KvaeVaeDims(B * T, pixels, int(a.shape[-1]), 1 if out_distr.lower() == "bernoulli" else 0, xv,
                       float(scale_reconstruction), float(beta))
    return _VaeLoss.apply(x, x_mu, a, a_mu, a_var, m, dims)


class _Reparam(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, var, eps):
        dev = mu.device
        m, v, e = prep(mu), prep(var), prep(eps)
        a = torch.empty_like(m)
        _check(_lib().kvae_vae_reparam_fwd(capi._ptr(m, "mu"), capi._ptr(v, "var"), capi._ptr(e, "eps"), m.numel(), capi._ptr(a, "a"),
                                           dev.index, capi._stream(dev)), "kvae_vae_reparam_fwd")
        ctx.save_for_backward(v, e)
        return a.view(mu.shape)

    @staticmethod
    def backward(ctx, g):
        v, e = ctx.saved_tensors
        gc = prep(g)
        dv = torch.empty_like(v)
        _check(_lib().kvae_vae_reparam_bwd(capi._ptr(v, "var"), capi._ptr(e, "eps"), capi._ptr(gc, "g"), v.numel(), capi._ptr(dv, "d_var"),
                                           v.device.index, capi._stream(v.device)), "kvae_vae_reparam_bwd")
        return g, dv.view(g.shape), None


def reparameterize(mu, var, eps=None):
    """kvae/model/model.py:81-84: a = mu + eps * sqrt(var + 1e-6), eps ~ N(0, I) drawn with the reference's call
    (torch.randn_like) unless given."""
    if eps is None:
        eps = torch.randn_like(var)
    return _Reparam.apply(mu, var, eps)


def _vae_loss_per_pixel_variance(x, x_mu, x_var, a, a_mu, a_var, scale_reconstruction, beta, mask, out_distr):
    """losses.py:62-111 for a TENSOR-valued x_var (one variance per pixel; KVAE.compute_loss passes a scalar and runs in
    the reduction kernel above): plain torch ops under autograd -- a library route beside the kernel path."""
    import math
    B, T = x.shape[:2]
    m = torch.ones(B, T, dtype=x.dtype, device=x.device) if mask is None else mask.to(device=x.device, dtype=x.dtype).view(B, T)
    logn = lambda v, mean, var: -0.5 * math.log(2.0 * math.pi) - 0.5 * torch.log(var) - (v - mean) ** 2 / (2.0 * var)   # :5-17
    if out_distr.lower() == "bernoulli":
        per_frame = -torch.nn.functional.binary_cross_entropy_with_logits(x_mu, x, reduction="none").flatten(2).sum(-1)   # :83-85
    else:
        per_frame = logn(x, x_mu, x_var).flatten(2).sum(-1)
    denom = m.sum().clamp(min=1.0)                                                   # :81
    recon = (per_frame * m).sum() / denom
    log_q = logn(a, a_mu, a_var).sum(-1)
    log_p = (-0.5 * math.log(2.0 * math.pi) - 0.5 * a * a).sum(-1)                   # standard-normal prior, :96-99
    reg = ((log_p - log_q) * m).sum() / denom                                        # :103-105
    return scale_reconstruction * recon + beta * reg, recon, reg                     # :107-109

"""Autograd plumbing over the C ABI: forward recursion, ELBO and the explicit adjoint.

Three entry points:
  smooth_fwd(...)            filter (+ RTS smoother) -> the reference's 7-/9-tuple tensors
  SmoothFunction             autograd node for dense cotangents of those outputs (general path)
  FusedElboFunction          elbo as a function of the ORIGINAL inputs (Y,U,alpha,A,B,C,Q): its
                             backward runs the complete adjoint (ELBO + smoother + filter + mixing)
                             in one launch, with the states saved by smooth_fwd re-used as is.

torch is used for memory, streams and autograd bookkeeping only; all arithmetic of the path is in
libkvae_kalman.so.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from . import capi


_supported_cache = {}   # (n,p,m,K,variant,lanes) -> bool: kvae_supported() is pure
_size_cache = {}        # (kind, dims fields) -> bytes / counts: the workspace-size entry points are pure


def _dims_key(kind, d):
    return (kind, d.B, d.T, d.n, d.p, d.m, d.K, d.q_per_mode, d.c_shared, d.lanes, d.flags)


def bwd_workspace_bytes(dims):
    k = _dims_key("bwd", dims)
    v = _size_cache.get(k)
    if v is None:
        v = _size_cache[k] = capi.bwd_workspace_bytes(dims)
    return v


def mask_partials_count(dims):
    k = _dims_key("mp", dims)
    v = _size_cache.get(k)
    if v is None:
        v = _size_cache[k] = capi.mask_partials_count(dims)
    return v


@dataclass
class Problem:
    """Inputs of one call, normalised to contiguous fp32 CUDA tensors."""
    Y: torch.Tensor
    U: Optional[torch.Tensor]
    mask: Optional[torch.Tensor]
    alpha: torch.Tensor
    A: torch.Tensor
    Bm: torch.Tensor
    C: torch.Tensor
    Q: torch.Tensor
    R: torch.Tensor
    mu0: torch.Tensor
    Sigma0: torch.Tensor
    q_per_mode: bool
    c_shared: bool
    lanes: int = 0
    mu_init: Optional[torch.Tensor] = None
    Sigma_init: Optional[torch.Tensor] = None
    flags: int = 0
    dense: Optional[tuple] = None      # (A [B,T,n,n], B [B,T,n,m], C [B,T,p,n], Q [B,T,n,n] | None): forward only
    mask_partials: Optional[torch.Tensor] = None   # left by the forward launch of this problem (States.mask_partials)
    dims: object = field(default=None, repr=False)

    def __post_init__(self):
        B, T, p = self.Y.shape
        K, n, m = self.Bm.shape
        self.dims = capi.make_dims(B, T, n, p, m, K, self.q_per_mode, self.c_shared, self.lanes, self.flags)
        self._inputs_cache = None
        key = (n, p, m, K, bool(self.q_per_mode), bool(self.c_shared), int(self.lanes))   # lanes = 0: every count the library may pick is instantiated
        ok = _supported_cache.get(key)
        if ok is None:
            ok = _supported_cache[key] = capi.supported(self.dims)
        if not ok:
            raise capi.KvaeError(
                f"shape (n={n}, p={p}, m={m}, K={K}, switching={self.q_per_mode}, lanes={self.lanes}) is not "
                "instantiated (libkvae_kalman.so holds the tuples of kalman_vae_b200/csrc/kvae_configs.h; any other (n <= 16, p, m, "
                "K) is compiled on demand unless KVAE_JIT=0; lanes must be a power of two dividing n)")

    @property
    def shape(self):
        d = self.dims
        return d.B, d.T, d.n, d.p, d.m, d.K

    def inputs(self, Y=None, U=None):
        """kvae_inputs for this problem (built once: the tensors of a Problem never change)."""
        if Y is None and U is None and self._inputs_cache is not None:
            return self._inputs_cache
        d = self.dense or (None, None, None, None)
        c = capi.make_inputs(self.Y if Y is None else Y, self.U if U is None else U, self.mask, self.alpha,
                             self.A, self.Bm, self.C, self.Q, self.R, self.mu0, self.Sigma0,
                             self.mu_init, self.Sigma_init, d[0], d[1], d[2], d[3])
        if Y is None and U is None:
            self._inputs_cache = c
        return c


def prep(t, device=None):
    """contiguous fp32 CUDA tensor with a 16-byte aligned base (no copy when already so)."""
    if t is None:
        return None
    t = t.detach()
    if device is not None and t.device != device:
        t = t.to(device)
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16 != 0:
        t = t.clone()
    return t


_info_cache = {}
_ws_cache = {}


def workspace(device, kind, nbytes):
    """Persistent scratch per (device, stream, kind), grown on demand: kernels are stream-ordered, so the
    same scratch can serve every call on that stream (no allocator traffic on the hot path)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream, kind)
    w = _ws_cache.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(max(int(nbytes * 1.25), 256), dtype=torch.uint8, device=device)
        _ws_cache[key] = w
    return w


def info_word(device):
    """One persistent int32 flag per device (zeroed by the caller when it wants to read it)."""
    w = _info_cache.get(device)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=device)
        _info_cache[device] = w
    return w


@dataclass
class States:
    mus_filt: torch.Tensor
    Sigmas_filt: torch.Tensor
    mus_pred: torch.Tensor
    Sigmas_pred: torch.Tensor
    mus_smooth: Optional[torch.Tensor] = None
    Sigmas_smooth: Optional[torch.Tensor] = None
    mask_partials: Optional[torch.Tensor] = None   # per-CTA mask sums of the forward launch (kvae_states.mask_partials)
    a_filt: Optional[torch.Tensor] = None          # [B,T,p] C_t mu_{t|t}  (optional forward output)
    a_smooth: Optional[torch.Tensor] = None        # [B,T,p] C_t mu_{t|T}  (optional forward output)

    def c_struct(self):
        return capi.make_states(self.mus_filt, self.Sigmas_filt, self.mus_pred, self.Sigmas_pred,
                                self.mus_smooth, self.Sigmas_smooth, self.mask_partials, self.a_filt, self.a_smooth)


def smooth_fwd(pb: Problem, smooth=True, lists=True, projections=False):
    """Runs the forward recursion; returns (States, A_list, B_list, C_list).  projections: the launch also emits
    States.a_filt = C_t mu_{t|t} and (smooth) States.a_smooth = C_t mu_{t|T}."""
    B, T, n, p, m, K = pb.shape
    dev = pb.Y.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    st = States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n),
                e(B, T, n, 1) if smooth else None, e(B, T, n, n) if smooth else None)
    if projections:
        st.a_filt = e(B, T, p)
        st.a_smooth = e(B, T, p) if smooth else None
    if smooth and pb.dense is None and not (pb.dims.flags & capi.FLAG_SMOOTH_ONLY):
        st.mask_partials = e(max(mask_partials_count(pb.dims), 4))
    A_list = e(B, T, n, n) if lists else None
    B_list = e(B, T, n, m) if lists else None
    # shared emission matrix: the reference returns a stack of C[0] (switch_dyn_param.py:85-86);
    # an expanded view has the same values without B*T copies of it
    C_list = None
    if lists and not pb.c_shared:
        C_list = e(B, T, p, n)
    capi.filter_smooth_fwd(pb.dims, pb.inputs(), st.c_struct(), A_list, B_list, C_list, info_word(dev), dev)
    if lists and pb.c_shared:
        C_list = pb.C[0].expand(B, T, p, n)
    return st, A_list, B_list, C_list


def elbo_terms(pb: Problem, st: States, eps, jitter=1e-6, Y=None, U=None):
    """terms[8] (see include/kvae_kalman.h)."""
    dev = pb.Y.device
    terms = torch.empty(8, dtype=torch.float32, device=dev)
    ws = workspace(dev, "elbo", capi.elbo_workspace_bytes(pb.dims))
    capi.elbo_fwd(pb.dims, pb.inputs(Y, U), st.c_struct(), eps, jitter, terms, ws, info_word(dev), dev)
    return terms


def adjoint(pb: Problem, st: States, eps=None, jitter=1e-6, g_elbo=None, terms=None, cot=None, need_dU=True,
            elbo_only=False, with_elbo=False):
    """Explicit adjoint.  Returns dict(dY,dU,dalpha,dA,dBm,dC,dQ[,dmus,dSigmas]).
    with_elbo: fused value + adjoint (KVAE_FLAG_WITH_ELBO): `terms` is WRITTEN by the launch."""
    B, T, n, p, m, K = pb.shape
    dev = pb.Y.device
    e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    grads = dict(dY=e(B, T, p), dU=e(B, T, m) if need_dU else None, dalpha=e(B, T, K),
                 dA=e(K, n, n), dBm=e(K, n, m), dC=e(K, p, n), dQ=e(K, n, n) if pb.q_per_mode else None)
    dims = pb.dims
    if elbo_only:
        grads["dmus"], grads["dSigmas"] = e(B, T, n), e(B, T, n, n)
        dims = capi.make_dims(B, T, n, p, m, K, pb.q_per_mode, pb.c_shared, pb.lanes, capi.FLAG_ELBO_ONLY)
    if with_elbo:
        dims = capi.make_dims(B, T, n, p, m, K, pb.q_per_mode, pb.c_shared, pb.lanes, capi.FLAG_WITH_ELBO)
    ws = workspace(dev, "bwd", bwd_workspace_bytes(dims))
    capi.bwd(dims, pb.inputs(), st.c_struct(), eps, jitter, g_elbo, terms, cot, grads, ws, info_word(dev), dev)
    return grads


_ones_cache = {}


def _one(device):
    w = _ones_cache.get(device)
    if w is None:
        w = torch.ones(1, dtype=torch.float32, device=device)
        _ones_cache[device] = w
    return w


_COT_ORDER = ("mus_smooth", "Sigmas_smooth", "mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred",
              "A_list", "B_list", "C_list")


class SmoothFunction(torch.autograd.Function):
    """(Y, U, alpha, A, Bm, C, Q) -> the outputs of filter()/smooth(); dense-cotangent backward."""

    @staticmethod
    def forward(ctx, pb: Problem, smooth: bool, Y, U, alpha, A, Bm, C, Q):
        st, A_list, B_list, C_list = smooth_fwd(pb, smooth=smooth, lists=True)
        pb.mask_partials = st.mask_partials
        ctx.pb, ctx.smooth = pb, smooth
        # outputs are saved through save_for_backward: holding them on ctx directly would create a
        # ctx -> output -> grad_fn -> ctx cycle that only the cyclic GC frees (device memory would pile up)
        ctx.save_for_backward(*[t for t in (st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred,
                                            st.mus_smooth, st.Sigmas_smooth) if t is not None])
        ctx.has_U = U is not None
        outs = [st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred, A_list, B_list]
        ctx.c_materialised = not pb.c_shared
        if ctx.c_materialised:
            outs.append(C_list)
        if smooth:
            outs = [st.mus_smooth, st.Sigmas_smooth] + outs
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        pb = ctx.pb
        st = States(*ctx.saved_tensors)
        gouts = list(gouts)
        names = (["mus_smooth", "Sigmas_smooth"] if ctx.smooth else []) + \
                ["mus_filt", "Sigmas_filt", "mus_pred", "Sigmas_pred", "A_list", "B_list"] + \
                (["C_list"] if ctx.c_materialised else [])
        cot = {k: prep(g) for k, g in zip(names, gouts) if g is not None}
        st_b = st
        if not ctx.smooth:
            # filter-only call: the adjoint kernel still wants smoothed-state buffers; with no
            # cotangent on them and T handled generically they are only read for D = Ss1 - Sp1
            # whose adjoint is multiplied by zero, so alias the filtered ones.
            st_b = States(st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred, st.mus_filt, st.Sigmas_filt)
        g = adjoint(pb, st_b, cot=cot, need_dU=ctx.has_U)
        return (None, None, g["dY"], g["dU"] if ctx.has_U else None, g["dalpha"], g["dA"], g["dBm"], g["dC"],
                g["dQ"] if pb.q_per_mode else None)


class FusedElboFunction(torch.autograd.Function):
    """elbo(Y, U, alpha, A, Bm, C, Q) with the smoothed states of a previous smooth_fwd on the same
    inputs passed as constants.  forward: ELBO sweep; backward: sweeps 3+4 in one launch."""

    @staticmethod
    def forward(ctx, pb: Problem, st: States, eps, jitter, extra, Y, U, alpha, A, Bm, C, Q):
        ctx.has_U = U is not None
        ctx.eager = None
        if any(ctx.needs_input_grad[5:]):
            # a gradient will be asked for: evaluate the ELBO and its complete adjoint in ONE launch now
            # (the adjoint recomputes every quantity of the ELBO anyway); backward() only scales by the upstream factor
            terms = torch.empty(8, dtype=torch.float32, device=pb.Y.device)
            ctx.eager = adjoint(pb, st, eps=eps, jitter=jitter, g_elbo=_one(pb.Y.device), terms=terms, need_dU=ctx.has_U,
                                with_elbo=True)
        else:
            terms = elbo_terms(pb, st, eps, jitter)
        ctx.pb, ctx.st, ctx.eps, ctx.jitter, ctx.terms = pb, st, eps, jitter, terms
        # extra: optional 0-dim tensor added inside the normalisation: (log_p - log_q).sum()
        val = terms[5]
        if extra is not None:
            val = val + extra * terms[6]
        ctx.has_extra = extra is not None
        return val.clone()

    @staticmethod
    def backward(ctx, g):
        pb = ctx.pb
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        if ctx.eager is not None:
            names = ["dY", "dalpha", "dA", "dBm", "dC"] + (["dU"] if ctx.has_U else []) + (["dQ"] if pb.q_per_mode else [])
            gr = dict(zip(names, torch._foreach_mul([ctx.eager[k] for k in names], g.reshape(()))))
        else:
            gr = adjoint(pb, ctx.st, eps=ctx.eps, jitter=ctx.jitter, g_elbo=g, terms=ctx.terms, need_dU=ctx.has_U)
        g_extra = (g * ctx.terms[6]).reshape(()) if ctx.has_extra else None
        return (None, None, None, None, g_extra, gr["dY"], gr["dU"] if ctx.has_U else None, gr["dalpha"], gr["dA"],
                gr["dBm"], gr["dC"], gr["dQ"] if pb.q_per_mode else None)


class ElboFunction(torch.autograd.Function):
    """elbo(mu, Sigma, Y, U, alpha, A, Bm, C, Q) for ARBITRARY (mu, Sigma) tensors (e.g. filtered states, or
    smoothed states of another call): the general form of KalmanFilter.elbo.  backward = ELBO adjoint only."""

    @staticmethod
    def forward(ctx, pb: Problem, eps, jitter, extra, mu, Sigma, Y, U, alpha, A, Bm, C, Q):
        B, T, n, p, m, K = pb.shape
        mu_c, Sig_c = prep(mu).reshape(B, T, n, 1), prep(Sigma)
        st = States(mu_c, Sig_c, mu_c, Sig_c, mu_c, Sig_c)   # only the "smoothed" slots are read
        terms = elbo_terms(pb, st, eps, jitter)
        ctx.pb, ctx.st, ctx.eps, ctx.jitter, ctx.terms = pb, st, eps, jitter, terms
        ctx.has_U, ctx.has_extra, ctx.mu_shape = U is not None, extra is not None, mu.shape
        val = terms[5]
        if extra is not None:
            val = val + extra * terms[6]
        return val.clone()

    @staticmethod
    def backward(ctx, g):
        pb = ctx.pb
        g = g.detach().to(torch.float32).reshape(1).contiguous()
        gr = adjoint(pb, ctx.st, eps=ctx.eps, jitter=ctx.jitter, g_elbo=g, terms=ctx.terms, need_dU=ctx.has_U,
                     elbo_only=True)
        g_extra = (g * ctx.terms[6]).reshape(()) if ctx.has_extra else None
        return (None, None, None, g_extra, gr["dmus"].reshape(ctx.mu_shape), gr["dSigmas"], gr["dY"],
                gr["dU"] if ctx.has_U else None, gr["dalpha"], gr["dA"], gr["dBm"], gr["dC"],
                gr["dQ"] if pb.q_per_mode else None)


class RegimeSampleFunction(torch.autograd.Function):
    """(logits [B,T,K,K], init_logits [B,K]) -> (y_seq [B,T,K], log_q [B,T], log_p [B,T]): the Gumbel-softmax regime
    chain of SwitchingDynamicsParameter.compute_batch (switch_dyn_param.py:51-79) as one launch; backward = one
    explicit-adjoint launch (csrc/kvae_regime.cu)."""

    @staticmethod
    def forward(ctx, logits, init_logits, gumbel, trans, tau, hard):
        B, T, K, _ = logits.shape
        dev = logits.device
        lg, il, gn, tr = prep(logits), prep(init_logits), prep(gumbel), prep(trans, dev)
        y = torch.empty(B, T, K, dtype=torch.float32, device=dev)
        lq = torch.empty(B, T, dtype=torch.float32, device=dev)
        lp = torch.empty(B, T, dtype=torch.float32, device=dev)
        capi.regime_fwd(B, T, K, hard, tau, lg, il, gn, tr, y, lq, lp, dev)
        ctx.save_for_backward(lg, il, gn, tr, y)
        ctx.tau, ctx.hard = tau, hard
        ctx.mark_non_differentiable()
        return y, lq, lp

    @staticmethod
    def backward(ctx, g_y, g_lq, g_lp):
        lg, il, gn, tr, y = ctx.saved_tensors
        B, T, K, _ = lg.shape
        dev = lg.device
        d_logits = torch.empty_like(lg)
        d_init = torch.empty_like(il)
        capi.regime_bwd(B, T, K, ctx.hard, ctx.tau, lg, il, gn, tr, y, prep(g_y), prep(g_lq), prep(g_lp), d_logits, d_init, dev)
        return d_logits, d_init, None, None, None, None

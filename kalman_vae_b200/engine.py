"""KalmanStep — the hot path as one pre-planned unit of work per rank.

One step = smooth (filter + RTS smoother) -> ELBO -> explicit adjoint, on this rank's shard of the
batch, with every buffer (states, lists, gradients, workspaces) allocated once and the kernel
sequence captured in CUDA graphs.

Data parallelism (one process per GPU, batch sharded): the adjoint is linear in the upstream factor
c = g / max(sum(mask), 1), so every rank runs its forward + adjoint with the normaliser left out
(c = 1), and ONE all-reduce per step sums a flat buffer [dA | dB | dC | dQ | 5 ELBO sums] over the
ranks; the global 1/max(sum(mask),1) is applied afterwards to the reduced parameter gradients, to the
local dY / dalpha and to the ELBO.  This is the only collective of the path; forward-only use
(imputation) has none.

This is what bench.py times for the device-resident number, and what a trainer that does not need
autograd in between can call directly.  The autograd route (KalmanFilter.smooth / .elbo / backward)
launches exactly the same kernels.
"""
from __future__ import annotations

import torch

from . import capi
from . import dist as kdist
from .functional import Problem, States, info_word


class KalmanStep:
    def __init__(self, pb: Problem, eps: torch.Tensor, jitter: float = 1e-6, use_graphs: bool = True, group=None,
                 lists: bool = True, need_dU: bool = False):
        self.pb, self.eps, self.jitter, self.group = pb, eps, jitter, group
        B, T, n, p, m, K = pb.shape
        dev = pb.Y.device
        self.dev = dev
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.st = States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
        self.A_list = e(B, T, n, n) if lists else None
        self.B_list = e(B, T, n, m) if lists else None
        self.C_list = e(B, T, p, n) if (lists and not pb.c_shared) else None
        # flat reduction buffer: parameter gradients followed by the ELBO terms (views into it)
        sizes = [K * n * n, K * n * m, K * p * n] + ([K * n * n] if pb.q_per_mode else [])
        psz = sum(sizes)
        psz_pad = (psz + 3) & ~3
        self.flat = torch.zeros(psz_pad + 8, dtype=torch.float32, device=dev)
        self.n_reduce = psz_pad + 5
        o, views = 0, []
        for s_, shp in zip(sizes, [(K, n, n), (K, n, m), (K, p, n), (K, n, n)]):
            views.append(self.flat[o:o + s_].view(*shp))
            o += s_
        self.terms = self.flat[psz_pad:psz_pad + 8]
        self.g_elbo = torch.ones(1, dtype=torch.float32, device=dev)
        self.grads = dict(dY=e(B, T, p), dU=e(B, T, m) if need_dU else None, dalpha=e(B, T, K), dA=views[0],
                          dBm=views[1], dC=views[2], dQ=views[3] if pb.q_per_mode else None)
        self.ws_elbo = torch.empty(max(capi.elbo_workspace_bytes(pb.dims), 16), dtype=torch.uint8, device=dev)
        self.ws_bwd = torch.empty(max(capi.bwd_workspace_bytes(pb.dims), 16), dtype=torch.uint8, device=dev)
        self.info = info_word(dev)
        self._inputs = pb.inputs()
        self._states = self.st.c_struct()
        self.kernel_launches_per_step = 5  # k_filter_smooth, k_elbo, k_elbo_final, k_bwd, k_param_final
        self.world = torch.distributed.get_world_size(group) if kdist._active(group) else 1
        self.graph_main = self.graph_post = None
        if use_graphs:
            self._capture()

    # ------------------------------------------------------------------ the C-ABI calls of one step
    def _compute(self):
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)
        capi.elbo_fwd(self.pb.dims, self._inputs, self._states, self.eps, self.jitter, self.terms, self.ws_elbo,
                      self.info, self.dev)
        if self.world > 1:
            # leave the normaliser out of the adjoint: c = g_elbo * terms[6] = max(sum mask,1) / max(sum mask,1) = 1
            torch.clamp(self.terms[4:5], min=1.0, out=self.g_elbo)
        capi.bwd(self.pb.dims, self._inputs, self._states, self.eps, self.jitter, self.g_elbo, self.terms, None,
                 self.grads, self.ws_bwd, self.info, self.dev)

    def _post(self):
        """after the all-reduce: apply the GLOBAL normaliser"""
        t = self.terms
        inv = torch.reciprocal(torch.clamp(t[4:5], min=1.0))
        self.flat[:self.n_reduce - 5].mul_(inv)
        self.grads["dY"].mul_(inv)
        self.grads["dalpha"].mul_(inv)
        if self.grads["dU"] is not None:
            self.grads["dU"].mul_(inv)
        t[5:6].copy_((t[0:1] + t[1:2] + t[2:3] + t[3:4]) * inv)
        t[6:7].copy_(inv)

    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):   # warm-up outside capture (sets kernel attributes, loads modules)
            self._compute()
            if self.world > 1:
                self._post()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph_main = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_main):
            self._compute()
        if self.world > 1:
            self.graph_post = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_post):
                self._post()

    def step(self):
        """Enqueues one fwd+ELBO+bwd pass (no host sync).  Returns the device tensor `terms` (terms[5] = elbo)."""
        if self.graph_main is not None:
            self.graph_main.replay()
        else:
            self._compute()
        if self.world > 1:
            torch.distributed.all_reduce(self.flat[:self.n_reduce], op=torch.distributed.ReduceOp.SUM, group=self.group)
            if self.graph_post is not None:
                self.graph_post.replay()
            else:
                self._post()
        return self.terms

    def forward_only(self):
        """smooth only (the imputation path): one launch, no collective."""
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)

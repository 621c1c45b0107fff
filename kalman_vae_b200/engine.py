"""KalmanStep — the hot path as one pre-planned unit of work per rank.

One step = smooth (filter + RTS smoother) -> ELBO -> explicit adjoint, on this rank's shard of the
batch, with every buffer (states, lists, gradients, workspaces) allocated once and the kernel
sequence captured in two CUDA graphs (forward+ELBO | adjoint).  Between the two graphs sits the only
collective of the path: the all-reduce of the five ELBO partial sums (global mask normalisation);
after the second the all-reduce of the flat parameter-gradient buffer (kalman_vae_b200.dist).

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
        self.terms = torch.zeros(8, dtype=torch.float32, device=dev)
        self.g_elbo = torch.ones(1, dtype=torch.float32, device=dev)
        self.grads = dict(dY=e(B, T, p), dU=e(B, T, m) if need_dU else None, dalpha=e(B, T, K), dA=e(K, n, n),
                          dBm=e(K, n, m), dC=e(K, p, n), dQ=e(K, n, n) if pb.q_per_mode else None)
        self.ws_elbo = torch.empty(max(capi.elbo_workspace_bytes(pb.dims), 16), dtype=torch.uint8, device=dev)
        self.ws_bwd = torch.empty(max(capi.bwd_workspace_bytes(pb.dims), 16), dtype=torch.uint8, device=dev)
        self.info = info_word(dev)
        self._inputs = pb.inputs()
        self._states = self.st.c_struct()
        self.kernel_launches_per_step = 5  # k_filter_smooth, k_elbo, k_elbo_final, k_bwd, k_param_final
        self.graph_fwd = self.graph_bwd = None
        self.world = torch.distributed.get_world_size(group) if kdist._active(group) else 1
        if use_graphs:
            self._capture()

    # the three C-ABI calls
    def _fwd(self):
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)
        capi.elbo_fwd(self.pb.dims, self._inputs, self._states, self.eps, self.jitter, self.terms, self.ws_elbo,
                      self.info, self.dev)

    def _bwd(self):
        capi.bwd(self.pb.dims, self._inputs, self._states, self.eps, self.jitter, self.g_elbo, self.terms, None,
                 self.grads, self.ws_bwd, self.info, self.dev)

    def _capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):   # warm-up outside capture (sets kernel attributes, loads modules)
            self._fwd()
            self._bwd()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph_fwd, self.graph_bwd = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_fwd):
            self._fwd()
        with torch.cuda.graph(self.graph_bwd):
            self._bwd()

    def step(self):
        """Enqueues one fwd+ELBO+bwd pass (no host sync).  Returns the device tensor `terms` (terms[5] = elbo)."""
        if self.graph_fwd is not None:
            self.graph_fwd.replay()
        else:
            self._fwd()
        if self.world > 1:
            kdist.globalize_elbo_terms(self.terms, self.group)
        if self.graph_bwd is not None:
            self.graph_bwd.replay()
        else:
            self._bwd()
        if self.world > 1:
            g = self.grads
            kdist.allreduce_param_grads([g["dA"], g["dBm"], g["dC"], g["dQ"]], self.group)
        return self.terms

    def forward_only(self):
        """smooth only (the imputation path): one launch, no collective."""
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)

"""KalmanStep — the hot path as one pre-planned unit of work per rank.

One step = smooth (filter + RTS smoother) -> ELBO value + explicit adjoint (one fused launch,
KVAE_FLAG_WITH_ELBO), on this rank's shard of the batch, with every buffer (states, lists, gradients,
workspaces) allocated once and the kernel sequence captured in a CUDA graph.

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

import os

import torch

from . import capi
from . import dist as kdist
from .functional import Problem, States, info_word


class KalmanStep:
    def __init__(self, pb: Problem, eps: torch.Tensor, jitter: float = 1e-6, use_graphs: bool = True, group=None,
                 lists: bool = True, need_dU: bool = False, collective: str | None = None):
        self.pb, self.eps, self.jitter, self.group = pb, eps, jitter, group
        B, T, n, p, m, K = pb.shape
        dev = pb.Y.device
        self.dev = dev
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.st = States(e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n))
        self.st.mask_partials = e(max(capi.mask_partials_count(pb.dims), 4))
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
        self.world = torch.distributed.get_world_size(group) if kdist._active(group) else 1
        # fused value + adjoint launch; under data parallelism the normaliser is applied after the all-reduce
        self.dims_bwd = capi.make_dims(B, T, n, p, m, K, pb.q_per_mode, pb.c_shared, pb.dims.lanes,
                                       capi.FLAG_WITH_ELBO | (capi.FLAG_RAW_SUMS if self.world > 1 else 0))
        self.ws_bwd = torch.empty(max(capi.bwd_workspace_bytes(self.dims_bwd), 16), dtype=torch.uint8, device=dev)
        self.info = info_word(dev)
        self._inputs = pb.inputs()
        self._states = self.st.c_struct()
        self.kernel_launches_per_step = 3  # k_filter_smooth, k_bwd (ELBO value + adjoint), k_bwd_final
        # data parallel: the exchange runs in this library's own kernels over NVLink peer memory (dist.PeerExchange);
        # KVAE_DP_COLLECTIVE=nccl selects the torch.distributed all-reduce + scaling kernels instead
        self.peer = None
        self.peer_two_launch = False
        self.collective = "none"
        if self.world > 1:
            self.collective = "nccl"
            mode = collective or os.environ.get("KVAE_DP_COLLECTIVE", "peer")   # peer: fused into the adjoint's final kernel (kvae_kf_bwd_dp);
            self.peer_two_launch = (mode == "peer2")                # peer2: kvae_kf_bwd + kvae_dp_finalize; nccl: torch.distributed
            if mode in ("peer", "peer2"):
                try:
                    self.peer = kdist.PeerExchange(dev, psz, group)
                    self.collective = "nvlink-peer-memory"
                except RuntimeError as err:
                    import warnings
                    warnings.warn(f"{err}; using the NCCL all-reduce")
        self.graph_main = self.graph_post = None
        if use_graphs:
            self._capture()

    # ------------------------------------------------------------------ the C-ABI calls of one step
    def _compute(self):
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)
        if self.peer is not None and not self.peer_two_launch:
            # the final kernel of the adjoint also does the cross-rank exchange (NVLink peer memory)
            capi.bwd_dp(self.dims_bwd, self._inputs, self._states, self.eps, self.jitter, self.g_elbo, self.terms,
                        self.grads, self.ws_bwd, self.info, self.dev, self.peer.comm)
        else:
            capi.bwd(self.dims_bwd, self._inputs, self._states, self.eps, self.jitter, self.g_elbo, self.terms, None,
                     self.grads, self.ws_bwd, self.info, self.dev)
            if self.peer is not None:
                capi.dp_finalize(self.pb.dims, self.peer.comm, self.grads, self.terms, self.info, self.dev)

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
            if self.world > 1 and self.peer is None:
                self._post()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph_main = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_main):
            self._compute()
        if self.world > 1 and self.peer is None:
            self.graph_post = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_post):
                self._post()

    def step(self):
        """Enqueues one fwd+ELBO+bwd pass (no host sync).  Returns the device tensor `terms` (terms[5] = elbo)."""
        if self.graph_main is not None:
            self.graph_main.replay()
        else:
            self._compute()
        if self.world > 1 and self.peer is None:
            torch.distributed.all_reduce(self.flat[:self.n_reduce], op=torch.distributed.ReduceOp.SUM, group=self.group)
            if self.graph_post is not None:
                self.graph_post.replay()
            else:
                self._post()
        return self.terms

    def check(self):
        """Host-side look at the device status word (ONE synchronising read: call it when the results are consumed, not
        per launch).  Raises if a factorisation met a non-positive pivot (1: the reference would raise or retry with a
        larger jitter, kalman_filter.py:282-302) or if the data-parallel exchange gave up waiting for a peer (2: the
        gradients of that step were NOT written)."""
        code = int(self.info.item())
        if code & capi.INFO_PEER:
            raise RuntimeError("kvae: the data-parallel exchange timed out waiting for a peer; gradients of this step are invalid")
        if code:
            raise torch.linalg.LinAlgError(f"kvae: a Cholesky / LU pivot was not positive in this step (status word {code}: "
                                           "1 filter/smoother, 4 Sigma_smooth, 8 Q)")

    def close(self):
        """Releases the peer-memory exchange (CUDA IPC mappings); the object must not be stepped afterwards."""
        if self.peer is not None:
            self.peer.close()
            self.peer = None

    def forward_only(self):
        """smooth only (the imputation path): one launch, no collective."""
        capi.filter_smooth_fwd(self.pb.dims, self._inputs, self._states, self.A_list, self.B_list, self.C_list,
                               self.info, self.dev)


class HostPipeline:
    """KalmanStep for inputs that live in (pinned) HOST memory: the end-to-end form of the training step.

    `slots` device-side copies of the per-step inputs (Y, U, mask, alpha, eps) are kept, each with its own
    KalmanStep (states, gradients, CUDA graph).  A copy stream uploads step i+1 while the compute stream runs
    step i; the compute of consecutive steps stays strictly ordered (a trainer updates the parameters in
    between), only the host->device transfer overlaps.  Per step the pipeline
        1. waits until the slot's previous step no longer reads its inputs,
        2. copies the five input tensors host->device (cudaMemcpyAsync from pinned memory, copy stream),
        3. replays the step graph (k_filter_smooth, k_bwd with the fused ELBO value, k_bwd_final) on the compute
           stream [+ the one all-reduce under data parallelism],
        4. copies [dA | dB | dC | dQ | terms] (one flat buffer) device->host into a pinned result buffer.
    dY / dalpha (the gradients that flow on to the encoder and the dynamics network) stay on the device.
    """

    def __init__(self, shape, params, q_per_mode=False, c_shared=False, has_U=True, has_mask=True, lanes=0, slots=2,
                 device=None, group=None, jitter=1e-6):
        B, T, n, p, m, K = shape
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.dev, self.slots = dev, slots
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.inputs, self.steps, self.out_host, self.info_host = [], [], [], []
        self.ev_in = [torch.cuda.Event() for _ in range(slots)]
        self.ev_free = [torch.cuda.Event() for _ in range(slots)]
        self.ev_out = [torch.cuda.Event() for _ in range(slots)]
        self._used = [False] * slots
        with torch.cuda.stream(self.compute_stream):
            # the per-step inputs of a slot are views into ONE device slab (256-byte aligned pieces), mirrored by a pinned
            # host slab (host_slab()): a caller that fills the host views moves a whole step with one cudaMemcpyAsync
            self._layout, off = {}, 0
            for name, shp in (("Y", (B, T, p)), ("U", (B, T, m) if has_U else None), ("mask", (B, T) if has_mask else None),
                              ("alpha", (B, T, K)), ("eps", (B, T, n))):
                if shp is None:
                    continue
                cnt = 1
                for v in shp:
                    cnt *= v
                self._layout[name] = (off, cnt, shp)
                off += (cnt + 63) & ~63
            self._slab_floats = off
            self.slabs = []
            for _ in range(slots):
                slab = e(off)
                self.slabs.append(slab)
                view = lambda nm: slab[self._layout[nm][0]:self._layout[nm][0] + self._layout[nm][1]].view(*self._layout[nm][2])
                d = dict(Y=view("Y"), U=view("U") if has_U else None, mask=view("mask") if has_mask else None,
                         alpha=view("alpha"), eps=view("eps"))
                for t in d.values():   # defined contents for the graph-capture warm-up run
                    if t is not None:
                        t.zero_()
                if d["mask"] is not None:
                    d["mask"].fill_(1.0)
                d["alpha"].fill_(1.0 / K)
                pb = Problem(d["Y"], d["U"], d["mask"], d["alpha"], params["A"], params["B"], params["C"], params["Q"],
                             params["R"], params["mu0"], params["Sigma0"], q_per_mode, c_shared, lanes=lanes)
                self.inputs.append(d)
                self.steps.append(KalmanStep(pb, d["eps"], jitter=jitter, use_graphs=True, group=group, need_dU=False))
                self.out_host.append(torch.empty(self.steps[-1].flat.numel(), dtype=torch.float32).pin_memory())
                self.info_host.append(torch.zeros(1, dtype=torch.int32).pin_memory())
        self.compute_stream.synchronize()
        self.h2d_bytes_per_step = sum(t.numel() * 4 for t in self.inputs[0].values() if t is not None)
        self.d2h_bytes_per_step = self.out_host[0].numel() * 4 + 4
        self.i = 0

    def step(self, Y, U, mask, alpha, eps):
        """Enqueues one step on host tensors (pinned for a truly asynchronous copy).  Returns the slot index to
        pass to `result()`; does not synchronise.  U=None / mask=None: nothing is copied for that input and the slot's
        device-resident tensor is used as it is (zeros / ones from construction -- what the reference creates on the
        device itself every step, model.py:149-150 and train.py:41)."""
        k = self.i % self.slots
        self.i += 1
        d = self.inputs[k]
        with torch.cuda.stream(self.copy_stream):
            if self._used[k]:
                self.copy_stream.wait_event(self.ev_free[k])
            for name, src in (("Y", Y), ("U", U), ("mask", mask), ("alpha", alpha), ("eps", eps)):
                if d[name] is not None and src is not None:
                    d[name].copy_(src, non_blocking=True)
            self.ev_in[k].record(self.copy_stream)
        return self._launch(k)

    def host_slab(self):
        """A pinned host buffer with the slot layout and its named views {Y, U, mask, alpha, eps}: fill the views
        (e.g. as the collate buffer of a data loader), then pass the slab to step_packed()."""
        slab = torch.empty(self._slab_floats, dtype=torch.float32).pin_memory()
        views = {nm: slab[o:o + c].view(*shp) for nm, (o, c, shp) in self._layout.items()}
        return slab, views

    def step_packed(self, slab):
        """step() for a host_slab(): the five inputs cross PCIe as ONE copy."""
        k = self.i % self.slots
        self.i += 1
        with torch.cuda.stream(self.copy_stream):
            if self._used[k]:
                self.copy_stream.wait_event(self.ev_free[k])
            self.slabs[k].copy_(slab, non_blocking=True)
            self.ev_in[k].record(self.copy_stream)
        return self._launch(k)

    def _launch(self, k):
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.ev_in[k])
            self.steps[k].step()
            self.ev_free[k].record(self.compute_stream)
            self.out_host[k].copy_(self.steps[k].flat, non_blocking=True)
            self.info_host[k].copy_(self.steps[k].info, non_blocking=True)   # status word travels with the results
            self.ev_out[k].record(self.compute_stream)
        self._used[k] = True
        return k

    def result(self, k):
        """Blocks until step `k`'s results are in host memory.  Returns (elbo, flat_host) where flat_host is the pinned
        buffer [dA | dB | dC | dQ | pad | terms(8)] of that slot (valid until the slot is used again)."""
        self.ev_out[k].synchronize()
        code = int(self.info_host[k][0])
        if code:   # bit 2: a peer never arrived and the gradients of this step were not written; else a non-positive pivot
            self.steps[k].info.zero_()
            raise (RuntimeError if code & capi.INFO_PEER else torch.linalg.LinAlgError)(
                "kvae: " + ("the data-parallel exchange timed out waiting for a peer; this step's gradients are invalid"
                            if code & capi.INFO_PEER else f"a Cholesky / LU pivot was not positive (status word {code})"))
        out = self.out_host[k]
        return float(out[out.numel() - 3]), out   # terms[5]

    def close(self):
        for s in self.steps:
            s.close()

    def device_grads(self, k):
        """dY, dalpha (+ parameter-gradient views) of slot k on the device, ordered on the compute stream."""
        return self.steps[k].grads


class ImputePipeline:
    """`KalmanFilter.impute_observations` (the Kalman part of KVAE.impute, model.py:267-288) for inputs that live in pinned
    HOST memory, chunked over the batch so that the three engines of the step overlap:

        copy stream  : host -> device of chunk c+1 (Y, mask, and alpha / U when given)
        main stream  : filter + smoother launch of chunk c (forward only, no list tensors, projections emitted by the sweeps)
        drain stream : device -> host of chunk c-1's `a_imputed` (and `a_filtered`)

    Sequences are independent, so chunking changes no value.  A monolithic call moves inputs, computes and moves results
    one after the other (BASELINE cfg3: 1.57 GB up, 11 ms of kernels, 0.52 GB down = 48 ms per step on one B200); the
    pipeline is bound by the larger of the two PCIe directions alone.  `chunk` should stay >= 12 288 sequences where the
    batch allows it (the thread-per-sequence kernels, see kvae_pick_lanes)."""

    def __init__(self, kf, chunk=16384, want_filtered=False, device=None):
        self.kf, self.chunk, self.want_filtered = kf, int(chunk), want_filtered
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.drain_stream = torch.cuda.Stream(device=self.dev)
        self._slots = None

    def _alloc(self, T, p, m, K, has_u, has_alpha):
        key = (T, p, m, K, has_u, has_alpha)
        if self._slots is not None and self._key == key:
            return
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=self.dev)
        self._slots = [dict(Y=e(self.chunk, T, p), mask=e(self.chunk, T), U=e(self.chunk, T, m) if has_u else None,
                            alpha=e(self.chunk, T, K) if has_alpha else None, ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(),
                            ev_out=torch.cuda.Event(), keep=None, used=False) for _ in range(2)]
        self._key = key

    @torch.no_grad()
    def run(self, Y, mask, alpha=None, U=None, out_imputed=None, out_filtered=None):
        """Y [B,T,p], mask [B,T] (and alpha [B,T,K] for a PrecomputedWeights dynamics object, U [B,T,m]): pinned host tensors.
        Returns (a_imputed, a_filtered | None) in pinned host memory (the given `out_*` buffers or new ones); blocks until
        the last chunk has arrived."""
        kf, dev = self.kf, self.dev
        B, T, p = Y.shape
        K = kf.dyn_params.A.size(0)
        self._alloc(T, p, kf.m, K, U is not None, alpha is not None)
        if out_imputed is None:
            out_imputed = torch.empty(B, T, p).pin_memory()
        if self.want_filtered and out_filtered is None:
            out_filtered = torch.empty(B, T, p).pin_memory()
        main = torch.cuda.current_stream(dev)
        bounds = [(lo, min(lo + self.chunk, B)) for lo in range(0, B, self.chunk)]

        def upload(c):
            lo, hi = bounds[c]
            s = self._slots[c % 2]
            with torch.cuda.stream(self.copy_stream):
                if s["used"]:
                    self.copy_stream.wait_event(s["ev_done"])      # the launch that last read this slot's inputs
                for name, src in (("Y", Y), ("mask", mask), ("U", U), ("alpha", alpha)):
                    if src is not None:
                        s[name][:hi - lo].copy_(src[lo:hi], non_blocking=True)
                s["ev_in"].record(self.copy_stream)

        upload(0)
        for c, (lo, hi) in enumerate(bounds):
            s = self._slots[c % 2]
            if c + 1 < len(bounds):
                upload(c + 1)
            main.wait_event(s["ev_in"])
            nb = hi - lo
            if alpha is not None:
                kf.dyn_params.set_weights(s["alpha"][:nb])
            else:
                kf.dyn_params.reset_state()
            a_imp, a_filt, _, _ = kf.impute_observations(s["Y"][:nb], None if U is None else s["U"][:nb], s["mask"][:nb])
            s["ev_done"].record(main)
            s["keep"] = (a_imp, a_filt)                            # alive until the drain stream is done with them
            with torch.cuda.stream(self.drain_stream):
                self.drain_stream.wait_event(s["ev_done"])
                out_imputed[lo:hi].copy_(a_imp, non_blocking=True)
                if self.want_filtered:
                    out_filtered[lo:hi].copy_(a_filt, non_blocking=True)
                a_imp.record_stream(self.drain_stream)
                a_filt.record_stream(self.drain_stream)
                s["ev_out"].record(self.drain_stream)
            s["used"] = True
        main.wait_stream(self.drain_stream)
        self.drain_stream.synchronize()
        return out_imputed, (out_filtered if self.want_filtered else None)

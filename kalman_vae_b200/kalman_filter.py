"""Drop-in `KalmanFilter` for rodrigo-paganini/kalman-vae backed by the sm_100a kernels.

Mirrors kvae/kalman/kalman_filter.py: same constructor, buffers (`Q,R,I,mu0,Sigma0`), attributes
(`n,m,p,dyn_params`), call signatures, return tuples and tensor layouts (means are [B,T,n,1]),
and the same side effects on `dyn_params` (`state_seq`, `Q_seq`, LSTM state).  Differences a
caller can observe:

  * CUDA only.  There is no CPU implementation; a CPU tensor raises.
  * With shared emission (switching dynamics) `C_list` is the expanded view `C[0].expand(B,T,p,n)`
    instead of a materialised stack (same values).
  * `elbo()` needs `A_list/B_list/C_list` returned by this object's own `filter()/smooth()` (it re-mixes
    A_t/B_t/C_t/Q_t from alpha inside the kernel, so gradients reach alpha and the base matrices
    directly); `mu_t_T/Sigma_t_T`, `y_t`, `u_t`, `mask` may be arbitrary tensors.  With the smoothed states
    of the same `smooth()` call the whole backward pass is one fused adjoint launch.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import functional as F
from .functional import Problem, States, prep


class _Provenance:
    """What the list tensors returned by filter()/smooth() were made from."""

    def __init__(self, pb, st, diff_inputs, smooth, with_grad=False):
        self.pb, self.st, self.diff_inputs, self.smooth = pb, st, diff_inputs, smooth
        # False: filter()/smooth() ran without autograd (torch.no_grad, or nothing required a gradient): the states and the
        # list tensors are constants for a later elbo(), exactly as they are in the reference
        self.with_grad = with_grad


def _tag(t, prov):
    try:
        t._kvae_prov = prov
    except Exception:  # pragma: no cover
        pass
    return t


class KalmanFilter(nn.Module):
    def __init__(self, std_dyn, std_obs, mu0, Sigma0, dyn_params, lanes=0, check_info="lazy"):
        super().__init__()
        self.dyn_params = dyn_params
        n = dyn_params.A.size(1)
        m = dyn_params.B.size(2)
        p = dyn_params.C.size(1)
        self.n, self.m, self.p = n, m, p
        dev, dtp = Sigma0.device, Sigma0.dtype
        self.register_buffer("Q", (std_dyn ** 2) * torch.eye(n, dtype=dtp, device=dev))
        self.register_buffer("R", (std_obs ** 2) * torch.eye(p, dtype=dtp, device=dev))
        self.register_buffer("I", torch.eye(n, dtype=dtp, device=dev))
        self.register_buffer("mu0", mu0.clone())
        self.register_buffer("Sigma0", Sigma0.clone())
        self.lanes = lanes              # 0: library picks lanes per sequence
        # what happens with the device status word ('a factorisation met a non-positive pivot') after elbo():
        #   "lazy" (default): no host synchronisation; the word is copied to pinned host memory and looked at when the
        #                     NEXT call of this object starts -- a failure raises one call late
        #   True            : read it right away (one host sync per elbo()) and run the reference's _safe_cholesky
        #                     ladder (kalman_filter.py:282-302: 10x jitter per retry, separately for Sigma_smooth and Q,
        #                     clamped-diagonal fallback after five tries)
        #   False           : never look
        self.check_info = check_info
        self.strict = False             # True: verify (host syncs) that elbo() sees the same y / u / mask VALUES as smooth()
        self._prep_cache = {}
        self._deferred = []             # [(event, pinned int32 word, kind)] device-side checks read one call late
        self._ring = None

    # ------------------------------------------------------------------ helpers
    def _mask(self, mask, B, T, ref):
        if mask is None:
            return None
        m = mask.to(device=ref.device, dtype=ref.dtype)
        if m.shape != (B, T):
            m = m.view(B, T)                                         # kalman_filter.py:131-133
        return m

    # ------------------------------------------------------------------ deferred device-side checks
    def _defer(self, flag, kind):
        """Queues a device int32 `flag` (non-zero = failure) for a look at the start of the next public call.
        (Not while a CUDA graph is being captured: a replay runs no host code, so there is nobody to look.)"""
        if torch.cuda.is_current_stream_capturing():
            return
        if self._ring is None:     # pinned words are allocated ONCE (cudaHostAlloc synchronises the device)
            self._ring = torch.zeros(16, dtype=torch.int32).pin_memory()
            self._ring_i = 0
        if len(self._deferred) >= 8:      # never let the queue catch up with the ring: wait for the oldest entry
            self._poll_deferred(block_oldest=True)
        host = self._ring[self._ring_i:self._ring_i + 1]
        self._ring_i = (self._ring_i + 1) % 16
        host.copy_(flag if flag.dtype == torch.int32 else flag.reshape(1).to(torch.int32), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(flag.device))
        self._deferred.append((ev, host, kind))

    def _poll_deferred(self, block_oldest=False):
        if not self._deferred or torch.cuda.is_current_stream_capturing():
            return
        keep = []
        for i, (ev, host, kind) in enumerate(self._deferred):
            if block_oldest and i == 0:
                ev.synchronize()
            if not ev.query():
                keep.append((ev, host, kind))
                continue
            code = int(host[0])
            if code:
                self._deferred = []
                if kind == "mask":
                    raise NotImplementedError(
                        "an earlier filter()/smooth() call with lstm dynamics and gradients enabled had missing observations "
                        "(mask != 1): the batched-alpha kernels equal the reference only for fully observed sequences (it feeds "
                        "C mu_pred to the LSTM at missing steps, kalman_filter.py:183-185).  Set kf.strict = True to have such "
                        "calls take the step-by-step autograd path (general_steps.py), or wrap them in torch.no_grad()")
                raise torch.linalg.LinAlgError(
                    "kvae: a factorisation in an earlier elbo() call met a non-positive pivot (status word "
                    f"{code}); construct KalmanFilter(..., check_info=True) to get the reference's jitter ladder "
                    "(kalman_filter.py:282-302) instead of this late report")
        self._deferred = keep

    def _lstm_alpha_batched(self, Y):
        """alpha [B,T,K] of the LSTM dynamics network for a fully observed sequence, one cuDNN call
        (dyn_param.py:50-56 stepped T times == one call over [0, a_0 .. a_{T-2}])."""
        dyn = self.dyn_params
        if hasattr(dyn, "compute_weights"):
            return dyn.compute_weights(Y)
        B, T, _ = Y.shape
        if dyn.K == 1:
            alpha = torch.ones(B, T, 1, device=Y.device, dtype=Y.dtype)
        else:
            shifted = torch.cat([torch.zeros_like(Y[:, :1]), Y[:, :-1]], dim=1)
            h, dyn.lstm_state = dyn.lstm(shifted, dyn.lstm_state)
            alpha = torch.softmax(dyn.head_w(h), dim=-1)
        dyn.state_seq = alpha
        return alpha

    def _weights(self, Y, mask_t):
        """(alpha, A, B, C, Q, q_per_mode, c_shared) for this call."""
        dyn = self.dyn_params
        if dyn.is_switching_dynamics:
            if hasattr(dyn, "compute_weights"):
                alpha = dyn.compute_weights(Y, is_training=self.training)
            else:  # reference object: run its own compute_batch and take the weights it leaves behind
                dyn.compute_batch(Y, is_training=self.training)       # kalman_filter.py:135-139
                alpha = dyn.state_seq
            return alpha, dyn.A, dyn.B, dyn.C, dyn.Q, True, True
        alpha = self._lstm_alpha_batched(Y)
        return alpha, dyn.A, dyn.B, dyn.C, self.Q, False, False

    def _prep_const(self, t, dev):
        """prep() of a parameter / buffer, cached until the tensor is modified in place or replaced."""
        key = id(t)
        hit = self._prep_cache.get(key)
        if (hit is not None and hit[0] == t._version and hit[1] == t.data_ptr() and hit[2].device == dev
                and hit[2].shape == t.shape and t.dtype == torch.float32):
            return hit[2]
        v = prep(t, dev)
        self._prep_cache[key] = (t._version, t.data_ptr(), v)
        return v

    def _problem(self, Y, U, mask_t, alpha, A, Bm, C, Q, qpm, csh, mu_init=None, Sigma_init=None):
        dev = Y.device
        if not Y.is_cuda:
            raise F.capi.KvaeError("KalmanFilter (B200-native) needs CUDA tensors; there is no CPU path")
        pc = self._prep_const
        return Problem(prep(Y), prep(U), prep(mask_t), prep(alpha), pc(A, dev), pc(Bm, dev), pc(C, dev),
                       pc(Q, dev), pc(self.R, dev), pc(self.mu0, dev), pc(self.Sigma0, dev), qpm, csh,
                       lanes=self.lanes, mu_init=prep(mu_init), Sigma_init=prep(Sigma_init))

    def _run(self, Y, U, mask, smooth):
        B, T, _ = Y.shape
        mask_t = self._mask(mask, B, T, Y)
        dyn = self.dyn_params
        self._poll_deferred()
        if Y.dtype != torch.float32:
            # the kernels compute in fp32; the reference runs in whatever dtype its buffers have (float64 works there,
            # SURVEY 8 a1): other dtypes take the step-by-step torch-op route in that dtype (general_steps.py)
            if not Y.is_cuda:
                raise F.capi.KvaeError("KalmanFilter (B200-native) needs CUDA tensors; there is no CPU path")
            from .general_steps import filter_smooth_stepwise
            if (not dyn.is_switching_dynamics) and dyn.K > 1 and hasattr(dyn, "lstm") and mask_t is not None:
                return filter_smooth_stepwise(self, Y, U, mask_t, smooth)      # alpha depends on the running prediction
            alpha, A, Bm, C, Q, qpm, csh = self._weights(Y, mask_t)
            return filter_smooth_stepwise(self, Y, U, mask_t, smooth, weights=(alpha.to(Y.dtype), A, Bm, C, Q, qpm, csh))
        if (not dyn.is_switching_dynamics) and dyn.K > 1 and hasattr(dyn, "lstm") and mask_t is not None:
            # lstm dynamics + a mask: alpha_{t+1} depends on the running prediction wherever mask_t = 0
            # (kalman_filter.py:159,183-185).  No host look at the mask's values:
            #  * no gradient wanted (imputation, evaluation): the LSTM runs inside the filter kernel -- exact for ANY mask;
            #  * gradient wanted (training): the batched-alpha path, which equals the reference iff mask == 1 everywhere
            #    (what the reference's trainer passes, train.py:41); a device-side test of that is read one call late
            #    (immediately with strict=True).
            wants_grad = torch.is_grad_enabled() and (Y.requires_grad or any(p.requires_grad for p in dyn.parameters()))
            if not wants_grad:
                return self._run_stepwise_lstm(Y, U, mask_t, smooth)
            bad = (mask_t != 1).any()
            if self.strict:
                if bool(bad.item()):
                    # missing observations under autograd: the reference's own step-by-step loop in torch ops (the gradient
                    # has to pass through the LSTM between the steps); not a kernel path (general_steps.py)
                    from .general_steps import filter_smooth_stepwise
                    return filter_smooth_stepwise(self, Y, U, mask_t, smooth)
            else:
                self._defer(bad, "mask")
        alpha, A, Bm, C, Q, qpm, csh = self._weights(Y, mask_t)
        pb = self._problem(Y, U, mask_t, alpha, A, Bm, C, Q, qpm, csh)
        diff = (Y, U, alpha, A, Bm, C, Q if qpm else None)
        needs_grad = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in diff)
        if needs_grad:
            outs = list(F.SmoothFunction.apply(pb, smooth, *diff))
            if smooth:
                ms, Ss = outs[0], outs[1]
                outs = outs[2:]
            mf, Sf, mp, Sp, A_list, B_list = outs[:6]
            C_list = outs[6] if not csh else C[0].expand(B, T, self.p, self.n)
            st = States(mf.detach(), Sf.detach(), mp.detach(), Sp.detach(),
                        ms.detach() if smooth else None, Ss.detach() if smooth else None, pb.mask_partials)
        else:
            st, A_list, B_list, C_list = F.smooth_fwd(pb, smooth=smooth, lists=True)
            mf, Sf, mp, Sp = st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred
            ms, Ss = st.mus_smooth, st.Sigmas_smooth
        prov = _Provenance(pb, st, diff, smooth, with_grad=needs_grad)
        prov.mus_smooth_ref = weakref.ref(ms) if smooth else None
        prov.Sigmas_smooth_ref = weakref.ref(Ss) if smooth else None
        for t in (A_list, B_list, C_list):
            _tag(t, prov)
        if smooth:
            return ms, Ss, mf, Sf, mp, Sp, A_list, B_list, C_list
        return mf, Sf, mp, Sp, A_list, B_list, C_list

    @torch.no_grad()
    def _run_stepwise_lstm(self, Y, U, mask_t, smooth):
        """lstm dynamics with missing observations: alpha_t depends on the running prediction
        (kalman_filter.py:159,183-185), so the filter advances one step per launch with the LSTM cell
        in between; the smoother then runs as one launch over the stored states."""
        dyn = self.dyn_params
        B, T, p = Y.shape
        dev = Y.device
        n, m = self.n, self.m
        fused = self._run_fused_lstm(Y, U, mask_t, smooth)
        if fused is not None:
            return fused
        y_for_dyn = torch.zeros(B, p, device=dev, dtype=Y.dtype)             # kalman_filter.py:142
        mu = prep(self.mu0, dev).expand(B, n).contiguous()
        Sig = prep(self.Sigma0, dev).expand(B, n, n).contiguous()
        alphas = []
        tm = lambda *s: torch.empty(T, B, *s, dtype=torch.float32, device=dev)   # time-major staging
        mf, Sf, mp, Sp, Al, Bl, Cl = tm(n, 1), tm(n, n), tm(n, 1), tm(n, n), tm(n, n), tm(n, m), tm(p, n)
        Yc, Uc = prep(Y), prep(U)
        for t in range(T):
            w = dyn.step_weights(y_for_dyn) if hasattr(dyn, "step_weights") else self._ref_step_weights(y_for_dyn)
            alphas.append(w)
            pb = self._problem(Yc[:, t:t + 1], None if Uc is None else Uc[:, t:t + 1], mask_t[:, t:t + 1],
                               w.unsqueeze(1), dyn.A, dyn.B, dyn.C, self.Q, False, False, mu_init=mu, Sigma_init=Sig)
            st = States(mf[t].view(B, 1, n, 1), Sf[t].view(B, 1, n, n), mp[t].view(B, 1, n, 1), Sp[t].view(B, 1, n, n))
            F.capi.filter_smooth_fwd(pb.dims, pb.inputs(), st.c_struct(), Al[t].view(B, 1, n, n), Bl[t].view(B, 1, n, m),
                                     Cl[t].view(B, 1, p, n), F.info_word(dev), dev)
            mu, Sig = mf[t].view(B, n), Sf[t]
            y_pred = (Cl[t] @ mp[t]).squeeze(-1)
            m_col = mask_t[:, t].view(B, 1)
            y_for_dyn = m_col * Yc[:, t] + (1.0 - m_col) * y_pred            # kalman_filter.py:183-185
        alpha = torch.stack(alphas, 1)
        dyn.state_seq = alpha                                                # kalman_filter.py:188-191
        bt = lambda x: x.transpose(0, 1).contiguous()
        mf, Sf, mp, Sp, A_list, B_list, C_list = (bt(x) for x in (mf, Sf, mp, Sp, Al, Bl, Cl))
        if not smooth:
            return mf, Sf, mp, Sp, A_list, B_list, C_list
        return self._smooth_from_filtered(Y, U, mask_t, alpha, mf, Sf, mp, Sp, A_list, B_list, C_list)

    def _run_fused_lstm(self, Y, U, mask_t, smooth):
        """The same recursion as the stepwise loop below in ONE launch: the LSTM cell, the head and the softmax run
        inside the filter kernel (csrc/kvae_lstm.cuh, kvae_kf_filter_lstm_fwd).  Returns None when the dynamics network
        is not a plain 1-layer nn.LSTM (hidden <= 52) + Linear head or the shape is not instantiated for it."""
        dyn = self.dyn_params
        lstm, head = getattr(dyn, "lstm", None), getattr(dyn, "head_w", None)
        B, T, p = Y.shape
        n, m, K = self.n, self.m, dyn.A.size(0)
        if not (isinstance(lstm, nn.LSTM) and isinstance(head, nn.Linear) and lstm.num_layers == 1 and not lstm.bidirectional
                and lstm.bias and getattr(lstm, "proj_size", 0) == 0 and lstm.input_size == p and lstm.hidden_size <= 52
                and head.bias is not None and n <= 8 and (n & (n - 1)) == 0 and K > 1 and self.lanes in (0, n)):
            return None
        dev = Y.device
        H = lstm.hidden_size
        pb = self._problem(Y, U, mask_t, None, dyn.A, dyn.B, dyn.C, self.Q, False, False)
        pb.dims.lanes = n
        e = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
        mf, Sf, mp, Sp = e(B, T, n, 1), e(B, T, n, n), e(B, T, n, 1), e(B, T, n, n)
        A_list, B_list, C_list, alpha = e(B, T, n, n), e(B, T, n, m), e(B, T, p, n), e(B, T, K)
        w = dict(w_ih=prep(lstm.weight_ih_l0, dev), w_hh=prep(lstm.weight_hh_l0, dev), b_ih=prep(lstm.bias_ih_l0, dev),
                 b_hh=prep(lstm.bias_hh_l0, dev), w_head=prep(head.weight, dev), b_head=prep(head.bias, dev),
                 h_out=e(B, H), c_out=e(B, H))
        if dyn.lstm_state is not None:
            w["h0"] = prep(dyn.lstm_state[0].reshape(B, H), dev)
            w["c0"] = prep(dyn.lstm_state[1].reshape(B, H), dev)
        st = States(mf, Sf, mp, Sp)
        try:
            F.capi.filter_lstm_fwd(pb.dims, pb.inputs(), st.c_struct(), A_list, B_list, C_list, w, H, alpha,
                                   F.info_word(dev), dev)
        except F.capi.KvaeError as err:
            if "not instantiated" in str(err) or "unsupported shape" in str(err):
                return None
            raise
        dyn.lstm_state = (w["h_out"].unsqueeze(0), w["c_out"].unsqueeze(0))
        dyn.state_seq = alpha                                                # kalman_filter.py:188-191
        if not smooth:
            return mf, Sf, mp, Sp, A_list, B_list, C_list
        return self._smooth_from_filtered(Y, U, mask_t, alpha, mf, Sf, mp, Sp, A_list, B_list, C_list)

    def _smooth_from_filtered(self, Y, U, mask_t, alpha, mf, Sf, mp, Sp, A_list, B_list, C_list):
        """smoother sweep over stored filter states (one launch, no refiltering) + provenance for elbo()"""
        dyn = self.dyn_params
        dev = Y.device
        pb = Problem(prep(Y), prep(U), prep(mask_t), prep(alpha), prep(dyn.A, dev), prep(dyn.B, dev), prep(dyn.C, dev),
                     prep(self.Q, dev), prep(self.R, dev), prep(self.mu0, dev), prep(self.Sigma0, dev), False, False,
                     lanes=self.lanes, flags=F.capi.FLAG_SMOOTH_ONLY)
        st = States(mf, Sf, mp, Sp, torch.empty_like(mf), torch.empty_like(Sf))
        F.capi.filter_smooth_fwd(pb.dims, pb.inputs(), st.c_struct(), None, None, None, F.info_word(dev), dev)
        pb.flags = 0
        pb.dims.flags = 0
        prov = _Provenance(pb, st, (Y, U, alpha, dyn.A, dyn.B, dyn.C, None), True)
        prov.mus_smooth_ref = weakref.ref(st.mus_smooth)
        prov.Sigmas_smooth_ref = weakref.ref(st.Sigmas_smooth)
        for t_ in (A_list, B_list, C_list):
            _tag(t_, prov)
        return (st.mus_smooth, st.Sigmas_smooth, mf, Sf, mp, Sp, A_list, B_list, C_list)

    def _ref_step_weights(self, a_tprev):
        dyn = self.dyn_params
        batch = a_tprev.size(0)
        if dyn.K == 1:
            w = torch.ones(batch, 1, device=a_tprev.device, dtype=a_tprev.dtype)
        else:
            h, dyn.lstm_state = dyn.lstm(a_tprev.unsqueeze(1), dyn.lstm_state)
            w = torch.softmax(dyn.head_w(h.squeeze(1)), dim=-1)
        if isinstance(dyn.state_seq, list):
            dyn.state_seq.append(w)
        return w

    # ------------------------------------------------------------------ public API (reference signatures)
    def filter(self, Y, U, mask=None):
        """kalman_filter.py:107-201 -> (mus_filt, Sigmas_filt, mus_pred, Sigmas_pred, A_list, B_list, C_list)"""
        return self._run(Y, U, mask, smooth=False)

    def smooth(self, Y, U, mask=None):
        """kalman_filter.py:240-279 -> 9-tuple (mus_smooth, Sigmas_smooth, mus_filt, ..., C_list)"""
        return self._run(Y, U, mask, smooth=True)

    @torch.no_grad()
    def impute_observations(self, Y, U, mask=None):
        """The Kalman part of KVAE.impute (model.py:267-288) without materialising the per-step
        matrices: smooth, then  a_imputed = C_t mu_{t|T},  a_filtered = C_t mu_{t|t}  straight from the mixture weights
        (`(C_list @ mus).squeeze(-1)` in the reference): the filter / smoother sweeps emit C_t mu themselves
        (kvae_states.a_filt / a_smooth) and skip A_list / B_list / C_list (160 of the 440 output bytes per sequence-step).  Returns (a_imputed [B,T,p], a_filtered [B,T,p],
        mus_smooth [B,T,n,1], mus_filt [B,T,n,1]).  Forward only, like the reference's impute()."""
        B, T, _ = Y.shape
        mask_t = self._mask(mask, B, T, Y)
        dyn = self.dyn_params
        self._poll_deferred()
        if (not dyn.is_switching_dynamics) and dyn.K > 1 and hasattr(dyn, "lstm") and mask_t is not None:
            outs = self._run_stepwise_lstm(Y, U, mask_t, True)          # fused LSTM launch (or the per-step fallback)
            ms, mf, alpha, csh = outs[0], outs[2], dyn.state_seq, False
        else:
            alpha, A, Bm, C, Q, qpm, csh = self._weights(Y, mask_t)
            pb = self._problem(Y, U, mask_t, alpha, A, Bm, C, Q, qpm, csh)
            # the sweeps emit the two projections themselves (kvae_states.a_filt / a_smooth): no list tensors, no GEMM
            st, _, _, _ = F.smooth_fwd(pb, smooth=True, lists=False, projections=True)
            return st.a_smooth, st.a_filt, st.mus_smooth, st.mus_filt
        # (lstm dynamics in the filter loop: the fused LSTM launch does not emit the projections; two small GEMMs)
        Cp = dyn.C.detach().to(torch.float32)
        K, p, n = Cp.shape
        Cflat = Cp.reshape(K * p, n).T
        al = alpha.to(torch.float32).unsqueeze(-1)
        proj = lambda mu: ((mu.squeeze(-1) @ Cflat).view(B, T, K, p) * al).sum(2)
        return proj(ms), proj(mf), ms, mf

    def elbo(self, mu_t_T, Sigma_t_T, y_t, u_t, A_list, B_list, C_list, Q_list=None, mask=None):
        """kalman_filter.py:305-401.  The standard-normal draw of `rsample` (:351) is made here with
        the same torch call the reference ends up in (`torch.empty(B,T,n).normal_()`)."""
        prov = getattr(A_list, "_kvae_prov", None)
        if prov is None or getattr(B_list, "_kvae_prov", None) is not prov or Q_list is not None:
            # list tensors that did not come from this object's filter()/smooth() (the kernels re-mix those from alpha), or
            # an explicit Q_list: the general form, batched torch ops on the GPU (general_elbo.py), autograd as the reference
            return self._elbo_given_lists(mu_t_T, Sigma_t_T, y_t, u_t, A_list, B_list, C_list, Q_list, mask)
        pb, st = prov.pb, prov.st
        B, T, n = pb.dims.B, pb.dims.T, pb.dims.n
        mask_t = self._mask(mask, B, T, y_t)
        # fast path: the states are the smoothed states of the SAME smooth() call as the lists, on the same y/mask
        fused = (prov.smooth and prov.mus_smooth_ref() is mu_t_T and prov.Sigmas_smooth_ref() is Sigma_t_T)
        if fused and self.strict:   # value checks cost a host sync each; switch off with kf.strict = False
            same_mask = (mask_t is None) == (pb.mask is None) and (
                mask_t is None or mask_t.data_ptr() == pb.mask.data_ptr() or torch.equal(mask_t.float(), pb.mask))
            same_y = y_t.data_ptr() == pb.Y.data_ptr() or torch.equal(y_t.detach().float(), pb.Y)
            fused = same_mask and same_y
        eps = prep(self._draw_eps(B, T, n, y_t))
        dyn = self.dyn_params
        extra = None
        if dyn.is_switching_dynamics:
            log_q, log_p = dyn.elbo_terms()                                        # :382-383
            extra = (log_p.sum() - log_q.sum()).to(torch.float32)
        Ys, Us, alpha, A, Bm, C, Q = prov.diff_inputs
        if not prov.with_grad:
            # the lists (and states) were produced without autograd: constants, as in the reference -- the ELBO then
            # differentiates only with respect to what THIS call is handed (mu, Sigma, y_t, u_t)
            fused = False
            alpha, A, Bm, C, Q = (t.detach() if t is not None else None for t in (alpha, A, Bm, C, Q))
        dev = y_t.device
        if fused:
            # y_t / u_t of this call are the same values as smooth()'s inputs: route the gradient of BOTH uses
            # (filter innovation and ELBO emission) to the tensor handed to elbo()
            y_in = y_t if y_t.requires_grad or not (Ys is not None and Ys.requires_grad) else Ys
            run = lambda jit: F.FusedElboFunction.apply(pb, st, eps, jit, extra, y_in, Us if Us is not None else None,
                                                        alpha, A, Bm, C, Q)
        else:
            # general form: (mu, Sigma) are arbitrary tensors (filtered states, another call's smoothed states, ...)
            u3 = u_t.squeeze(-1) if (u_t is not None and u_t.dim() == 4) else u_t
            pb2 = Problem(prep(y_t), prep(u3), prep(mask_t), pb.alpha, pb.A, pb.Bm, pb.C, pb.Q, pb.R, pb.mu0, pb.Sigma0,
                          pb.q_per_mode, pb.c_shared, lanes=self.lanes)
            run = lambda jit: F.ElboFunction.apply(pb2, eps, jit, extra, mu_t_T, Sigma_t_T, y_t, u3, alpha, A, Bm, C, Q)
        info = F.info_word(dev)
        if self.check_info is not True:
            # one launch with the reference's first rung (jitter 1e-6 on both factorisations); no host synchronisation:
            # the status word is looked at when the next call of this object starts ("lazy") or never (False)
            if self.check_info:
                info.zero_()
            val = run(1e-6)
            if self.check_info:
                self._defer(info, "chol")   # stream-ordered copy: later launches cannot overtake it
            return val
        # the reference's _safe_cholesky ladder (kalman_filter.py:282-302), one ladder per factorised family as there:
        # any failing matrix bumps the jitter of its family 10x for the WHOLE batch (:295-296); after five failed
        # attempts the family falls back to L = diag(sqrt(clamp(diag, 1e-6))) (:298-302)
        js = jq = 1e-6
        ds = dq = False
        fails_s = fails_q = 0
        while True:
            info.zero_()
            val = run((js, jq, ds, dq))
            code = int(info.item())
            if code & F.capi.INFO_PIVOT:
                raise torch.linalg.LinAlgError("kvae elbo: a pivot of the filter / smoother / R / Sigma0 factorisations was not positive")
            retry = False
            if (code & F.capi.INFO_CHOL_S) and not ds:
                fails_s += 1
                ds, js = (True, js) if fails_s >= 5 else (False, js * 10.0)
                retry = True
            if (code & F.capi.INFO_CHOL_Q) and not dq:
                fails_q += 1
                dq, jq = (True, jq) if fails_q >= 5 else (False, jq * 10.0)
                retry = True
            if not retry:
                break
        self.last_chol = dict(jitter_smooth=js, jitter_q=jq, diag_smooth=ds, diag_q=dq)
        return val

    def _elbo_given_lists(self, mu, Sigma, y_t, u_t, A_list, B_list, C_list, Q_list, mask):
        from .general_elbo import elbo_given_lists
        dyn = self.dyn_params
        Bsz, T = y_t.shape[0], y_t.shape[1]
        if Q_list is None:                                                         # kalman_filter.py:342-345
            Q_list = getattr(dyn, "Q_seq", None)
            if Q_list is None and dyn.is_switching_dynamics and getattr(dyn, "state_seq", None) is not None:
                # the mirrors do not materialise Q_seq (switch_dyn_param.py:84): mix it from the last regime weights
                Q_list = torch.einsum("btk,kij->btij", dyn.state_seq, dyn.Q)
            if Q_list is None:
                Q_list = self.Q
        extra = None
        if dyn.is_switching_dynamics:
            log_q, log_p = dyn.elbo_terms()                                        # :382-383
            extra = log_p.sum() - log_q.sum()
        eps = self._draw_eps(Bsz, T, self.n, y_t)
        return elbo_given_lists(mu, Sigma, y_t, u_t, A_list, B_list, C_list, Q_list, self.R, self.mu0, self.Sigma0,
                                self._mask(mask, Bsz, T, y_t), eps, extra)

    def _draw_eps(self, B, T, n, like):
        """The standard-normal draw behind MultivariateNormal.rsample (kalman_filter.py:351)."""
        return torch.empty(B, T, n, dtype=like.dtype, device=like.device).normal_()

    # ------------------------------------------------------------------ per-step forms (explicit matrices)
    def _dense_problem(self, Y, U, mask, T, A, Bm, C, Q, mu_init=None, Sigma_init=None, flags=0):
        """A forward launch that reads explicit per-step matrices [B,T,..] instead of mixing (forward only)."""
        dev = Y.device
        n, m, p = self.n, self.m, self.p
        dyn = self.dyn_params
        K = dyn.A.size(0)
        sw = bool(dyn.is_switching_dynamics)
        base_Q = dyn.Q if sw else self.Q
        pb = Problem(prep(Y), prep(U), prep(mask), None, prep(dyn.A, dev), prep(dyn.B, dev), prep(dyn.C, dev),
                     prep(base_Q, dev), prep(self.R, dev), prep(self.mu0, dev), prep(self.Sigma0, dev), sw, sw,
                     lanes=self.lanes, mu_init=prep(mu_init), Sigma_init=prep(Sigma_init), flags=flags,
                     dense=(prep(A), prep(Bm), prep(C), prep(Q)))
        return pb

    def filter_step(self, mu_t_t, Sigma_t_t, y_t, u_t, A, B, C, Q, mask_t=None):
        """kalman_filter.py:31-104 — one predict/update step with explicit per-sample matrices.
        Returns (mu_t_t [B,n,1], Sigma_t_t, mu_t_tprev [B,n,1], Sigma_t_tprev, A, B, C).  One kernel launch; when a
        gradient is wanted: batched torch ops under autograd (general_steps.py)."""
        batch = y_t.size(0)
        n, m, p = self.n, self.m, self.p
        if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in (mu_t_t, Sigma_t_t, y_t, u_t, A, B, C, Q)):
            from .general_steps import filter_step_ops
            if not y_t.is_cuda:
                raise F.capi.KvaeError("KalmanFilter (B200-native) needs CUDA tensors; there is no CPU path")
            mk = None if mask_t is None else (mask_t.expand(batch) if mask_t.dim() == 0 else mask_t.reshape(batch))
            mu_f, Sig_f, mu_p, Sig_p = filter_step_ops(mu_t_t.reshape(batch, n, 1), Sigma_t_t, y_t.reshape(batch, p, 1),
                                                      u_t.reshape(batch, m, 1), A, B, C, Q, self.R.to(y_t.dtype), mk)
            return mu_f, Sig_f, mu_p, Sig_p, A, B, C
        dev = y_t.device
        f = lambda x, *shape: x.detach().to(torch.float32).expand(*shape).contiguous().view(*shape)
        Qd = f(Q, batch, n, n).view(batch, 1, n, n)                                  # :35-36 (2-d Q is expanded)
        if mask_t is None:
            mask = None
        else:
            mk = mask_t.to(device=dev, dtype=torch.float32)
            mask = (mk.expand(batch) if mk.dim() == 0 else mk).contiguous().view(batch, 1)   # :54-60
        pb = self._dense_problem(y_t.reshape(batch, 1, p), u_t.reshape(batch, 1, m), mask, 1,
                                 f(A, batch, n, n).view(batch, 1, n, n), f(B, batch, n, m).view(batch, 1, n, m),
                                 f(C, batch, p, n).view(batch, 1, p, n), Qd,
                                 mu_init=mu_t_t.reshape(batch, n), Sigma_init=f(Sigma_t_t, batch, n, n))
        st, _, _, _ = F.smooth_fwd(pb, smooth=False, lists=False)
        return (st.mus_filt.view(batch, n, 1), st.Sigmas_filt.view(batch, n, n), st.mus_pred.view(batch, n, 1),
                st.Sigmas_pred.view(batch, n, n), A, B, C)

    def smooth_step(self, Sigma_t_t, Sigma_tpost_t, Sigma_tpost_T, mu_t_t, mu_tpost_t, mu_tpost_T, A):
        """kalman_filter.py:204-237 — one RTS step.  Returns (mu_t_T [B,n,1], Sigma_t_T).  One kernel launch; when a
        gradient is wanted: batched torch ops under autograd (general_steps.py)."""
        batch = Sigma_t_t.size(0)
        n, m, p = self.n, self.m, self.p
        if torch.is_grad_enabled() and any(t.requires_grad for t in (Sigma_t_t, Sigma_tpost_t, Sigma_tpost_T, mu_t_t,
                                                                     mu_tpost_t, mu_tpost_T, A)):
            from .general_steps import smooth_step_ops
            if not Sigma_t_t.is_cuda:
                raise F.capi.KvaeError("KalmanFilter (B200-native) needs CUDA tensors; there is no CPU path")
            r3 = lambda v: v.reshape(batch, n, 1)
            return smooth_step_ops(Sigma_t_t, Sigma_tpost_t, Sigma_tpost_T, r3(mu_t_t), r3(mu_tpost_t), r3(mu_tpost_T), A)
        dev = Sigma_t_t.device
        f32 = lambda x: x.detach().to(torch.float32)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        # a two-step "sequence": slot 0 = time t, slot 1 = time t+1.  The smoother sweep starts from the belief
        # stored at the last FILTERED slot, so the smoothed belief at t+1 is placed there.
        Sf = torch.stack([f32(Sigma_t_t), f32(Sigma_tpost_T)], 1).contiguous()
        mf = torch.stack([f32(mu_t_t).reshape(batch, n), f32(mu_tpost_T).reshape(batch, n)], 1).contiguous()
        Sp = torch.stack([z(batch, n, n), f32(Sigma_tpost_t)], 1).contiguous()
        mp = torch.stack([z(batch, n), f32(mu_tpost_t).reshape(batch, n)], 1).contiguous()
        Ad = torch.stack([z(batch, n, n), f32(A).expand(batch, n, n)], 1).contiguous()
        pb = self._dense_problem(z(batch, 2, p), None, None, 2, Ad, z(batch, 2, n, m), z(batch, 2, p, n), None,
                                 flags=F.capi.FLAG_SMOOTH_ONLY)
        st = States(mf.view(batch, 2, n, 1), Sf, mp.view(batch, 2, n, 1), Sp, z(batch, 2, n, 1), z(batch, 2, n, n))
        F.capi.filter_smooth_fwd(pb.dims, pb.inputs(), st.c_struct(), None, None, None, F.info_word(dev), dev)
        return st.mus_smooth[:, 0].reshape(batch, n, 1), st.Sigmas_smooth[:, 0]

    def _safe_cholesky(self, Sigma, max_tries=5, jitter_init=1e-6):
        """kalman_filter.py:282-302 for callers that use it directly (the ELBO kernel factorises in-kernel):
        chol(sym(Sigma) + jitter I) with the reference's 10x retry ladder and diagonal fallback."""
        n = Sigma.size(-1)
        Sigma = 0.5 * (Sigma + Sigma.mT)
        eye = torch.eye(n, device=Sigma.device, dtype=Sigma.dtype)
        jitter = jitter_init
        for _ in range(max_tries):
            L, info = torch.linalg.cholesky_ex(Sigma + jitter * eye)
            if int(info.max()) == 0:
                return L
            jitter *= 10.0
        diag = torch.clamp(torch.diagonal(Sigma, dim1=-2, dim2=-1), min=1e-6)
        return torch.diag_embed(torch.sqrt(diag))

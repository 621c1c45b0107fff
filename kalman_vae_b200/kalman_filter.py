"""Drop-in `KalmanFilter` for rodrigo-paganini/kalman-vae backed by the sm_100a kernels.

Mirrors kvae/kalman/kalman_filter.py: same constructor, buffers (`Q,R,I,mu0,Sigma0`), attributes
(`n,m,p,dyn_params`), call signatures, return tuples and tensor layouts (means are [B,T,n,1]),
and the same side effects on `dyn_params` (`state_seq`, `Q_seq`, LSTM state).  Differences a
caller can observe:

  * CUDA only.  There is no CPU implementation; a CPU tensor raises.
  * With shared emission (switching dynamics) `C_list` is the expanded view `C[0].expand(B,T,p,n)`
    instead of a materialised stack (same values).
  * `elbo()` accepts the tensors returned by this object's own `filter()/smooth()` (it re-mixes
    A_t/B_t/C_t/Q_t from alpha inside the kernel, so gradients reach alpha and the base matrices
    directly); hand-made `A_list/B_list/C_list` tensors are not supported yet.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import functional as F
from .functional import Problem, States, prep


class _Provenance:
    """What the list tensors returned by filter()/smooth() were made from."""

    def __init__(self, pb, st, diff_inputs, smooth):
        self.pb, self.st, self.diff_inputs, self.smooth = pb, st, diff_inputs, smooth


def _tag(t, prov):
    try:
        t._kvae_prov = prov
    except Exception:  # pragma: no cover
        pass
    return t


class KalmanFilter(nn.Module):
    def __init__(self, std_dyn, std_obs, mu0, Sigma0, dyn_params, lanes=0, check_info=True):
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
        self.check_info = check_info    # read the device 'non-positive pivot' flag after elbo()
        self.strict = True              # verify that elbo() sees the same y/mask values as smooth()
        self._mask_cache = None

    # ------------------------------------------------------------------ helpers
    def _mask(self, mask, B, T, ref):
        if mask is None:
            return None
        m = mask.to(device=ref.device, dtype=ref.dtype)
        if m.shape != (B, T):
            m = m.view(B, T)                                         # kalman_filter.py:131-133
        return m

    def _mask_is_ones(self, mask):
        """True when every entry is 1 (one host sync, cached on the tensor's identity/version)."""
        if mask is None:
            return True
        key = (mask.data_ptr(), mask._version, tuple(mask.shape))
        if self._mask_cache is not None and self._mask_cache[0] == key:
            return self._mask_cache[1]
        val = bool((mask == 1).all().item())
        self._mask_cache = (key, val)
        return val

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

    def _problem(self, Y, U, mask_t, alpha, A, Bm, C, Q, qpm, csh, mu_init=None, Sigma_init=None):
        dev = Y.device
        if not Y.is_cuda:
            raise F.capi.KvaeError("KalmanFilter (B200-native) needs CUDA tensors; there is no CPU path")
        return Problem(prep(Y), prep(U), prep(mask_t), prep(alpha), prep(A, dev), prep(Bm, dev), prep(C, dev),
                       prep(Q, dev), prep(self.R, dev), prep(self.mu0, dev), prep(self.Sigma0, dev), qpm, csh,
                       lanes=self.lanes, mu_init=prep(mu_init), Sigma_init=prep(Sigma_init))

    def _run(self, Y, U, mask, smooth):
        B, T, _ = Y.shape
        mask_t = self._mask(mask, B, T, Y)
        dyn = self.dyn_params
        if (not dyn.is_switching_dynamics) and dyn.K > 1 and hasattr(dyn, "lstm") and not self._mask_is_ones(mask_t):
            if torch.is_grad_enabled() and (Y.requires_grad or any(p.requires_grad for p in dyn.parameters())):
                raise NotImplementedError(
                    "lstm dynamics with missing observations is forward-only here (imputation, as in "
                    "KVAE.impute); wrap the call in torch.no_grad()")
            return self._run_stepwise_lstm(Y, U, mask_t, smooth)
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
                        ms.detach() if smooth else None, Ss.detach() if smooth else None)
        else:
            st, A_list, B_list, C_list = F.smooth_fwd(pb, smooth=smooth, lists=True)
            mf, Sf, mp, Sp = st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred
            ms, Ss = st.mus_smooth, st.Sigmas_smooth
        prov = _Provenance(pb, st, diff, smooth)
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
        if torch.is_grad_enabled() and (Y.requires_grad or any(p.requires_grad for p in self.dyn_params.parameters())):
            pass  # under torch.no_grad() this is never reached with grad enabled
        dyn = self.dyn_params
        B, T, p = Y.shape
        dev = Y.device
        n, m = self.n, self.m
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
        # smoother sweep over the stored filter states: rerun the fused forward with the alphas now known
        pb = self._problem(Y, U, mask_t, alpha, dyn.A, dyn.B, dyn.C, self.Q, False, False)
        st, A_list, B_list, C_list = F.smooth_fwd(pb, smooth=True, lists=True)
        prov = _Provenance(pb, st, (Y, U, alpha, dyn.A, dyn.B, dyn.C, None), True)
        prov.mus_smooth_ref = weakref.ref(st.mus_smooth)
        prov.Sigmas_smooth_ref = weakref.ref(st.Sigmas_smooth)
        for t_ in (A_list, B_list, C_list):
            _tag(t_, prov)
        return (st.mus_smooth, st.Sigmas_smooth, st.mus_filt, st.Sigmas_filt, st.mus_pred, st.Sigmas_pred,
                A_list, B_list, C_list)

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

    def elbo(self, mu_t_T, Sigma_t_T, y_t, u_t, A_list, B_list, C_list, Q_list=None, mask=None):
        """kalman_filter.py:305-401.  The standard-normal draw of `rsample` (:351) is made here with
        the same torch call the reference ends up in (`torch.empty(B,T,n).normal_()`)."""
        prov = getattr(A_list, "_kvae_prov", None)
        if prov is None or getattr(B_list, "_kvae_prov", None) is not prov:
            raise NotImplementedError(
                "elbo(): A_list/B_list/C_list must be the tensors returned by this KalmanFilter's "
                "filter()/smooth() (the kernels re-mix them from alpha); arbitrary list tensors are not supported yet")
        pb, st = prov.pb, prov.st
        B, T, n = pb.dims.B, pb.dims.T, pb.dims.n
        same_states = (prov.smooth and prov.mus_smooth_ref() is mu_t_T and prov.Sigmas_smooth_ref() is Sigma_t_T)
        if not same_states:
            raise NotImplementedError(
                "elbo(): mu_t_T/Sigma_t_T must be the smoothed states returned by the same smooth() call as the lists")
        if Q_list is not None:
            raise NotImplementedError("elbo(): explicit Q_list is not supported; Q is mixed from alpha in the kernel")
        mask_t = self._mask(mask, B, T, y_t)
        if self.strict:   # value checks cost a host sync each; switch off with kf.strict = False
            if (mask_t is None) != (pb.mask is None) or (mask_t is not None and mask_t.data_ptr() != pb.mask.data_ptr()
                                                           and not torch.equal(mask_t.float(), pb.mask)):
                raise NotImplementedError("elbo(): mask differs from the one given to smooth()")
            if y_t.data_ptr() != pb.Y.data_ptr() and not torch.equal(y_t.detach().float(), pb.Y):
                raise NotImplementedError("elbo(): y_t differs from the observations given to smooth()")
        eps = prep(self._draw_eps(B, T, n, y_t))
        dyn = self.dyn_params
        extra = None
        if dyn.is_switching_dynamics:
            log_q, log_p = dyn.elbo_terms()                                        # :382-383
            extra = (log_p.sum() - log_q.sum()).to(torch.float32)
        Ys, Us, alpha, A, Bm, C, Q = prov.diff_inputs
        # y_t / u_t of this call are the same values as smooth()'s inputs (checked above): route the
        # gradient to the tensors the caller handed to elbo() AND smooth() by summing over both uses.
        y_in = y_t if y_t.requires_grad or not (Ys is not None and Ys.requires_grad) else Ys
        jitter = 1e-6
        dev = y_t.device
        if self.check_info:
            F.info_word(dev).zero_()
        val = F.FusedElboFunction.apply(pb, st, eps, jitter, extra, y_in, Us if Us is not None else None,
                                        alpha, A, Bm, C, Q)
        if self.check_info and int(F.info_word(dev).item()) != 0:
            raise torch.linalg.LinAlgError(
                "kvae elbo: a Cholesky factorisation met a non-positive pivot (the reference's "
                "_safe_cholesky would retry with 10x jitter, kalman_filter.py:291-296)")
        return val

    def _draw_eps(self, B, T, n, like):
        """The standard-normal draw behind MultivariateNormal.rsample (kalman_filter.py:351)."""
        return torch.empty(B, T, n, dtype=like.dtype, device=like.device).normal_()

    def filter_step(self, mu_t_t, Sigma_t_t, y_t, u_t, A, B, C, Q, mask_t=None):
        raise NotImplementedError("per-step form with explicit matrices: use filter() (single-step launches are "
                                  "used internally for masked lstm dynamics)")

    def smooth_step(self, *a, **k):
        raise NotImplementedError("per-step form: use smooth()")

"""Host-side mirrors of the reference's dynamics-parameter modules.

They keep the reference's constructor signatures, parameter / sub-module names (so state dicts
load unchanged: `A,B,C[,Q]`, `lstm.*`, `head_w.*`, `markov_regime_posterior.{bigru,linear_head,
init_head}.*`) and the protocol `KalmanFilter` relies on (`is_switching_dynamics`, `reset_state`,
`state_seq`, `Q_seq`, `elbo_terms`, `compute_step`, `compute_batch`), and add one method the CUDA
path uses: `compute_weights(...) -> alpha [B,T,K]`.  The mixing of the K base matrices by alpha
(dyn_param.py:58-60, switch_dyn_param.py:82-86) is NOT done here: it happens inside the kernels.

The recurrent networks themselves (LSTM / bi-GRU, cuDNN) stay in PyTorch: they are out of the
hot path's scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class DynamicsParameter(nn.Module):
    """Mirror of kvae/kalman/dyn_param.py:5-63 (LSTM 'dynamics parameter network')."""

    def __init__(self, A, B, C, hidden_lstm=50):
        super().__init__()
        self.is_switching_dynamics = False
        self.K = A.size(0)
        self.n, self.m, self.p = A.size(1), B.size(2), C.size(1)
        self.A = nn.Parameter(A.clone())
        self.B = nn.Parameter(B.clone())
        self.C = nn.Parameter(C.clone())
        self.lstm_state = None
        self.state_seq = None
        if self.K > 1:
            self.lstm = nn.LSTM(input_size=self.p, hidden_size=hidden_lstm, num_layers=1, batch_first=True)
            self.head_w = nn.Linear(hidden_lstm, self.K)
            with torch.no_grad():  # bias alpha towards mode 0 at initialisation (dyn_param.py:31-33)
                self.head_w.bias.fill_(-10.0)
                self.head_w.bias[0] = 0.0

    def reset_state(self):
        self.lstm_state = None
        self.state_seq = []

    def step_weights(self, a_tprev):
        """alpha_t [B,K] from the previous (observed or predicted) encoding; advances the LSTM."""
        batch = a_tprev.size(0)
        if self.K == 1:
            w = torch.ones(batch, 1, device=a_tprev.device, dtype=a_tprev.dtype)
        else:
            h, self.lstm_state = self.lstm(a_tprev.unsqueeze(1), self.lstm_state)
            w = torch.softmax(self.head_w(h.squeeze(1)), dim=-1)
        self.state_seq.append(w)
        return w

    def compute_step(self, a_tprev):
        """Reference protocol (dyn_param.py:39-63): returns the mixed (A,B,C) for one step."""
        w = self.step_weights(a_tprev)
        A = torch.einsum("bk,kij->bij", w, self.A)
        B = torch.einsum("bk,knm->bnm", w, self.B)
        C = torch.einsum("bk,kpn->bpn", w, self.C)
        return A, B, C

    def compute_weights(self, a_seq):
        """alpha [B,T,K] for a fully observed sequence in ONE recurrent call.

        With mask == 1 the filter feeds the network y_for_dyn = a_{t-1} exactly
        (kalman_filter.py:142,183-185), i.e. the input sequence [0, a_0, ..., a_{T-2}]; running the
        LSTM once over it equals the reference's T single-step calls.
        """
        batch, T, _ = a_seq.shape
        if self.K == 1:
            alpha = torch.ones(batch, T, 1, device=a_seq.device, dtype=a_seq.dtype)
        else:
            shifted = torch.cat([torch.zeros_like(a_seq[:, :1]), a_seq[:, :-1]], dim=1)
            h, self.lstm_state = self.lstm(shifted, self.lstm_state)
            alpha = torch.softmax(self.head_w(h), dim=-1)
        self.state_seq = alpha  # what the reference leaves after filter() (kalman_filter.py:188-191)
        return alpha


class PrecomputedWeights(nn.Module):
    """dyn_params protocol object whose mixture weights alpha [B,T,K] are supplied by the caller
    (e.g. computed elsewhere or streamed from the host); base matrices are ordinary parameters.
    switching=True selects the SKVAE conventions (Q per mode, shared C)."""

    def __init__(self, A, B, C, Q=None, switching=False):
        super().__init__()
        self.is_switching_dynamics = bool(switching)
        self.K = A.size(0)
        self.n, self.m, self.p = A.size(1), B.size(2), C.size(1)
        self.A = nn.Parameter(A.clone())
        self.B = nn.Parameter(B.clone())
        self.C = nn.Parameter(C.clone())
        if switching:
            self.Q = nn.Parameter(Q.clone())
        self.alpha = None
        self.state_seq = None
        self.lstm_state = None
        self._zeros = None

    def set_weights(self, alpha):
        object.__setattr__(self, "alpha", alpha)      # plain per-call tensors: skip nn.Module's attribute bookkeeping

    def reset_state(self):
        object.__setattr__(self, "state_seq", None)

    def compute_weights(self, a_seq, is_training=True):
        B, T, _ = a_seq.shape
        object.__setattr__(self, "state_seq", self.alpha)
        z = self._zeros
        if z is None or z.shape != (B, T) or z.device != a_seq.device or z.dtype != a_seq.dtype:
            z = torch.zeros(B, T, device=a_seq.device, dtype=a_seq.dtype)
            object.__setattr__(self, "_zeros", z)
        object.__setattr__(self, "log_qseq", z)      # no regime chain here: log q = log p = 0
        object.__setattr__(self, "log_pseq", z)
        return self.alpha

    def elbo_terms(self):
        return self.log_qseq, self.log_pseq


class StickyRegimePrior:
    """Mirror of switch_dyn_param.py:98-110."""

    def __init__(self, K, p_stay=0.9):
        self.K = K
        self.p_stay = p_stay
        self.transition_matrix = torch.ones((K, K)) * ((1 - p_stay) / (K - 1))
        self.transition_matrix.fill_diagonal_(p_stay)


class MarkovVariationalRegimePosterior(nn.Module):
    """Mirror of switch_dyn_param.py:113-129 (bi-GRU + transition / initial heads)."""

    def __init__(self, K, input_dim, hidden_size=32):
        super().__init__()
        self.K = K
        self.hidden_size = hidden_size
        self.bigru = nn.GRU(input_size=input_dim, hidden_size=hidden_size, num_layers=1, batch_first=True,
                            bidirectional=True)
        self.linear_head = nn.Linear(2 * hidden_size, K * K)
        self.init_head = nn.Linear(2 * hidden_size, K)

    def forward(self, a_seq):
        h_seq, _ = self.bigru(a_seq)
        logits = self.linear_head(h_seq)
        B, T, _ = logits.shape
        return logits.view(B, T, self.K, self.K), self.init_head(h_seq[:, 0])


class SwitchingDynamicsParameter(nn.Module):
    """Mirror of kvae/kalman/switch_dyn_param.py:7-95 (SKVAE regime posterior)."""

    def __init__(self, A, B, C, Q=None, prior=None, hidden_lstm=32, markov_regime_posterior=None, reference_rng=False):
        super().__init__()
        self.is_switching_dynamics = True
        # reference_rng=True: draw the Gumbel noise as the reference does -- one [B,K] exponential draw per time step
        # (torch.nn.functional.gumbel_softmax called T times, switch_dyn_param.py:52,69) -- so that a run seeded like a
        # reference run samples the SAME regimes; the default draws the whole [B,T,K] chain in one call (one launch
        # instead of T, a different stream of the same distribution)
        self.reference_rng = bool(reference_rng)
        self.K = A.size(0)
        self.n, self.m, self.p = A.size(1), B.size(2), C.size(1)
        self.tau = 0.5
        if Q is None:
            Q = torch.eye(self.n, device=A.device, dtype=A.dtype).unsqueeze(0).repeat(self.K, 1, 1)
        self.A = nn.Parameter(A.clone())
        self.B = nn.Parameter(B.clone())
        self.C = nn.Parameter(C.clone())
        self.Q = nn.Parameter(Q.clone())
        self.s_tprev = None
        self.prior = prior if prior is not None else StickyRegimePrior(self.K)
        self.markov_regime_posterior = markov_regime_posterior or MarkovVariationalRegimePosterior(
            self.K, input_dim=self.p, hidden_size=hidden_lstm)
        self.hidden_size = hidden_lstm
        self.state_seq = None
        self.Q_seq = None

    def reset_state(self):
        self.state_seq = None

    def compute_weights(self, a_seq, is_training=True):
        """Regime weights y_seq [B,T,K] plus the log q / log p bookkeeping of
        switch_dyn_param.py:51-79; no mixed matrices are materialised."""
        batch, T, _ = a_seq.size()
        dev, dt = a_seq.device, a_seq.dtype
        if self.K == 1:
            self.log_qseq = torch.zeros(batch, T, device=dev, dtype=dt)
            self.log_pseq = torch.zeros(batch, T, device=dev, dtype=dt)
            self.state_seq = torch.ones(batch, T, 1, device=dev, dtype=dt)
            return self.state_seq
        if not a_seq.is_cuda:
            from .capi import KvaeError
            raise KvaeError("SwitchingDynamicsParameter (B200-native) needs CUDA tensors; there is no CPU path")
        from .functional import RegimeSampleFunction
        logits, init_logits = self.markov_regime_posterior(a_seq)
        gumbel = self._draw_gumbel(batch, T, self.K, logits)
        # the prior's transition matrix is a plain (CPU) tensor attribute, as in the reference: keep a device copy instead
        # of a pageable host-to-device copy per call (which synchronises, and cannot be part of a CUDA-graph capture)
        tm = self.prior.transition_matrix
        c = getattr(self, "_trans_dev", None)
        if c is None or c[0] is not tm or c[1] != tm._version or c[2].device != dev:
            c = (tm, tm._version, tm.detach().to(device=dev, dtype=torch.float32).contiguous())
            object.__setattr__(self, "_trans_dev", c)
        trans = c[2]
        # one launch for the whole chain (and one for its adjoint) instead of ~12 ops per time step
        y_seq, log_q, log_p = RegimeSampleFunction.apply(logits, init_logits, gumbel, trans, float(self.tau), not is_training)
        self.state_seq, self.log_qseq, self.log_pseq = y_seq, log_q, log_p
        return self.state_seq

    def _draw_gumbel(self, batch, T, K, like):
        """Gumbel(0,1) noise for every step: the draw torch.nn.functional.gumbel_softmax makes per call
        (`-empty_like(logits).exponential_().log()`), here for the whole [B,T,K] chain at once."""
        if self.reference_rng:
            # gumbel_softmax's own draw, per step and in the reference's order: -empty_like(logits_t).exponential_().log()
            return torch.stack([-torch.empty(batch, K, dtype=like.dtype, device=like.device).exponential_().log()
                                for _ in range(T)], dim=1)
        return -torch.empty(batch, T, K, dtype=like.dtype, device=like.device).exponential_().log()

    def compute_batch(self, a_seq, is_training=True):
        """Reference protocol (switch_dyn_param.py:37-92): mixed sequences for callers that want them."""
        y_seq = self.compute_weights(a_seq, is_training)
        batch, T, _ = a_seq.size()
        A_seq = torch.einsum("btk,kij->btij", y_seq, self.A)
        B_seq = torch.einsum("btk,knm->btnm", y_seq, self.B)
        Q_seq = torch.einsum("btk,kij->btij", y_seq, self.Q)
        C_seq = self.C[0].expand(batch, T, -1, -1)
        self.Q_seq = Q_seq
        return A_seq, B_seq, C_seq, Q_seq

    def elbo_terms(self):
        return self.log_qseq, self.log_pseq

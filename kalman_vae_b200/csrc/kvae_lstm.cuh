// kvae_lstm.cuh — the LSTM "dynamics parameter network" INSIDE the filter loop.
//
// With lstm dynamics and missing observations the mixture weights alpha_t depend on the running prediction:
//     y_for_dyn_t = mask_t * y_t + (1 - mask_t) * C_t mu_{t|t-1}            (kalman_filter.py:183-185)
//     alpha_{t+1} = softmax(head(LSTM_step(y_for_dyn_t)))                   (dyn_param.py:50-56), alpha_0 from zeros (:142)
// so the filter cannot be given alpha up front.  The reference (and the stepwise fallback of this package) runs one
// cuDNN LSTM step + ~10 torch ops + one filter launch per time step.  Here the cell runs in the filter kernel: the lane
// group that owns a sequence also owns its LSTM state.  Weights are staged once per CTA in shared memory as rows
// [ W_hh row (H, zero padded to HP) | W_ih row (p) | b_ih + b_hh | pad ] of LDW = 4 x odd floats (60 for p = 2): lane l
// owns the hidden units j = u*L + l, so the lanes of a group read rows LDW floats apart = disjoint banks for 128-bit loads.
#pragma once
#include "kvae_fwd.cuh"

namespace kvae {

struct LstmPtrs {
  const float *w_ih, *w_hh, *b_ih, *b_hh, *w_head, *b_head;   // [4H,p] [4H,H] [4H] [4H] [K,H] [K]   (gate order i,f,g,o)
  const float *h0, *c0;                                        // [B,H] or null (zeros)
  float *h_out, *c_out;                                        // [B,H] final state (nullable)
  float* alpha_out;                                            // [B,T,K] (= state_seq)
  int H;
};

constexpr int lstm_ldw(int need) {   // smallest row stride >= need that is 4 x odd floats (disjoint banks for the lanes)
  int y = (need + 3) & ~3;
  while ((y / 4) % 2 == 0) y += 4;
  return y;
}
template <class C> struct LstmGeo {
  static constexpr int L = C::L, U = (52 + L - 1) / L, HP = U * L, LDW = lstm_ldw(HP + C::P + 1);
  static_assert(HP % 4 == 0 && HP + C::P + 1 <= LDW, "row layout");
  static constexpr int oHead = 4 * HP * LDW;
  static constexpr int total = oHead + C::K * LDW;
};

template <class C> KV_FN float lstm_weight_at(const LstmPtrs& w, int idx) {
  using G = LstmGeo<C>;
  const int H = w.H;
  if (idx < G::oHead) {
    const int gate = idx / (G::HP * G::LDW), rem = idx % (G::HP * G::LDW), j = rem / G::LDW, c = rem % G::LDW;
    if (j >= H) return 0.f;
    const int row = gate * H + j;
    if (c < H) return w.w_hh[(size_t)row * H + c];
    if (c >= G::HP && c < G::HP + C::P) return w.w_ih[(size_t)row * C::P + (c - G::HP)];
    if (c == G::HP + C::P) return w.b_ih[row] + w.b_hh[row];
    return 0.f;
  }
  const int k = (idx - G::oHead) / G::LDW, c = (idx - G::oHead) % G::LDW;
  if (c < H) return w.w_head[(size_t)k * H + c];
  if (c == G::HP) return w.b_head[k];
  return 0.f;
}

KV_FN float kv_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

template <class C> struct LstmHook {
  static constexpr bool ON = true;
  using G = LstmGeo<C>;
  const float* W;        // staged weights (shared memory)
  float* hbuf;           // this group's hidden state [HP] (shared memory)
  float* alpha_out;      // [B,T,K]
  bool on;               // false: tail group mirroring a valid sequence (no stores)
  float c_own[G::U];
  float ydyn[C::P];

  KV_FN void init(const Group<C::L, C::R>& g, const LstmPtrs& w, int b) {
    KV_UNROLL for (int u = 0; u < G::U; ++u) {
      const int j = u * C::L + g.lane;
      c_own[u] = (w.c0 && j < w.H) ? w.c0[(size_t)b * w.H + j] : 0.f;
      hbuf[j] = (w.h0 && j < w.H) ? w.h0[(size_t)b * w.H + j] : 0.f;
    }
    KV_UNROLL for (int q = 0; q < C::P; ++q) ydyn[q] = 0.f;                       // kalman_filter.py:142
    g.sync();
  }
  // alpha_t = softmax(head(LSTM_step(y_for_dyn_{t-1})))                            (dyn_param.py:50-56)
  KV_FN void before_step(const Group<C::L, C::R>& g, long bt, StepIn<C>& cur) {
    float hv[G::HP];
    load_row<G::HP>(hbuf, hv);
    float hn[G::U];
    KV_UNROLL for (int u = 0; u < G::U; ++u) {
      const int j = u * C::L + g.lane;
      float acc[4];
      KV_UNROLL for (int gate = 0; gate < 4; ++gate) {
        const float* row = W + (size_t)(gate * G::HP + j) * G::LDW;
        float r[G::LDW];
        load_row<G::LDW>(row, r);
        float s = r[G::HP + C::P];
        KV_UNROLL for (int c = 0; c < G::HP; ++c) s = fmaf(r[c], hv[c], s);
        KV_UNROLL for (int q = 0; q < C::P; ++q) s = fmaf(r[G::HP + q], ydyn[q], s);
        acc[gate] = s;
      }
      const float ig = kv_sigmoid(acc[0]), fg = kv_sigmoid(acc[1]), gg = tanhf(acc[2]), og = kv_sigmoid(acc[3]);
      c_own[u] = fmaf(fg, c_own[u], ig * gg);
      hn[u] = og * tanhf(c_own[u]);
    }
    g.sync();   // everyone has read the old hidden state
    KV_UNROLL for (int u = 0; u < G::U; ++u) hbuf[u * C::L + g.lane] = hn[u];
    g.sync();
    load_row<G::HP>(hbuf, hv);
    float logit[C::K], mx = -INFINITY;
    KV_UNROLL for (int k = 0; k < C::K; ++k) {
      float r[G::LDW];
      load_row<G::LDW>(W + G::oHead + k * G::LDW, r);
      float s = r[G::HP];
      KV_UNROLL for (int c = 0; c < G::HP; ++c) s = fmaf(r[c], hv[c], s);
      logit[k] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
    KV_UNROLL for (int k = 0; k < C::K; ++k) { logit[k] = expf(logit[k] - mx); sum += logit[k]; }
    KV_UNROLL for (int k = 0; k < C::K; ++k) cur.al[k] = logit[k] / sum;
    if (on && g.lane == 0) { KV_UNROLL for (int k = 0; k < C::K; ++k) alpha_out[bt * C::K + k] = cur.al[k]; }
  }
  // y_for_dyn_t = m y_t + (1 - m) C_t mu_p                                          (kalman_filter.py:183-185)
  KV_FN void after_gain(const StepIn<C>& cur, const GainOut<C>& go) {
    KV_UNROLL for (int q = 0; q < C::P; ++q) ydyn[q] = cur.m * cur.y[q] + (1.0f - cur.m) * go.yp[q];
  }
  KV_FN void finish(const Group<C::L, C::R>& g, const LstmPtrs& w, int b) {
    if (!on) return;
    KV_UNROLL for (int u = 0; u < G::U; ++u) {
      const int j = u * C::L + g.lane;
      if (j < w.H) {
        if (w.h_out) w.h_out[(size_t)b * w.H + j] = hbuf[j];
        if (w.c_out) w.c_out[(size_t)b * w.H + j] = c_own[u];
      }
    }
  }
};

}  // namespace kvae

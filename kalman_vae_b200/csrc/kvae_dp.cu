// kvae_dp.cu — the ONE exchange of the data-parallel Kalman step, done over NVLink peer memory instead of NCCL.
//
// Every rank (one process per GPU) owns a small exchange buffer allocated with cudaMalloc and exported through CUDA
// IPC; every rank maps every peer's buffer.  After the fused adjoint launch (kvae_kf_bwd with WITH_ELBO|RAW_SUMS) a
// rank holds its LOCAL sums  v = [dA | dB | dC | dQ | trans, emiss, init, entropy, sum(mask)]  (a few hundred floats).
// kvae_dp_finalize then runs two launches on the caller's stream:
//
//   k_dp_publish  (1 CTA)   PUSHES v into every rank's exchange buffer (its own included) as 8-byte words
//                           {value, step}: one remote store per element and peer, no fence and no separate flag --
//                           an aligned 8-byte store arrives whole, so the step tag IS the ready flag (the "LL"
//                           protocol of NCCL).  Latency = one one-way NVLink store.
//   k_dp_final    (grid)    polls LOCAL memory only: every CTA waits for the five ELBO sums of every rank and derives
//                           the GLOBAL normaliser 1/max(sum_r sum(mask)_r, 1); then
//                             - parameter CTAs: wait for element i of every rank, sum IN RANK ORDER in fp64 (identical
//                               bits on every rank), multiply by the normaliser, scatter into dA | dB | dC | dQ;
//                             - CTA 0 also writes terms[0..7] (global sums, elbo, normaliser);
//                             - remaining CTAs scale this rank's dY / dalpha / dU in place.
//
// So the reduction, the normalisation (kalman_filter.py:392-400, a GLOBAL mask count) and the scaling of the local
// per-step gradients are one kernel; there is no NCCL launch, no separate "post" kernels, and the whole step
// (k_filter_smooth, k_bwd, k_bwd_final, k_dp_publish, k_dp_final) is one CUDA graph.
//
// Slot reuse: the words of step s live in slot s&1 and are overwritten by the peers' pushes of step s+2.  A peer
// reaches step s+2's publish only after its own k_dp_final of step s+1 consumed MY pushes of step s+1, which I issue
// after my k_dp_final of step s finished (stream order): nobody overwrites a slot that is still being read.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <new>
#include "../../include/kvae_kalman.h"
#include "kvae_ops.h"

namespace {

constexpr int MAX_WORLD = 16;
constexpr int HDR_WORDS = 16;     // u64 words: [0] = step counter of this rank (device side, advanced by k_dp_publish)

struct f4 { float x, y, z, w; };
typedef unsigned long long u64;

struct DpPeers { u64* buf[MAX_WORLD]; };   // rank r's exchange buffer as mapped in this process

// exchange buffer (u64 words): [header | slot 0: world x nf_pad | slot 1: world x nf_pad]
__host__ __device__ inline size_t ll_off(int s, int world, int src, size_t nf_pad, size_t i) {
  return HDR_WORDS + ((size_t)(s * world + src)) * nf_pad + i;
}
__device__ __forceinline__ void st_ll(u64* p, float v, unsigned step) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(step) : "memory");
}
// spins until the word carries `step`; false on time-out (~20 s: a peer died -- do not hang the GPU for ever)
__device__ __forceinline__ bool ld_ll(const u64* p, unsigned step, float& v) {
  unsigned lo, hi;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "l"(p) : "memory");
    if (hi == step) break;
    if (clock64() - t0 > 40000000000LL) { v = 0.f; return false; }
  }
  v = __uint_as_float(lo);
  return true;
}

struct LocalVec { const float* p[5]; int n[5]; };   // dA, dB, dC, dQ, terms[0..4]
struct OutVec { float* p[4]; int n[4]; };
struct ScaleJob { float* p[3]; long n[3]; };

__global__ void __launch_bounds__(1024) k_dp_publish(DpPeers peers, int rank, int world, size_t nf_pad, LocalVec lv) {
  u64* mine = peers.buf[rank];
  const unsigned step = (unsigned)mine[0] + 1u;          // graph replays advance the device-side counter
  const int s = (int)(step & 1u);
  int off = 0;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    if (lv.p[j]) {
      for (int i = threadIdx.x; i < lv.n[j]; i += blockDim.x) {
        const float v = lv.p[j][i];
        for (int q = 0; q < world; ++q) st_ll(peers.buf[q] + ll_off(s, world, rank, nf_pad, off + i), v, step);
      }
    }
    off += lv.n[j];
  }
  __syncthreads();                                        // everyone has read the old counter
  if (threadIdx.x == 0) mine[0] = step;
}

__global__ void __launch_bounds__(128) k_dp_final(DpPeers peers, int rank, int world, size_t nf_pad, int nparam, OutVec ov,
                                                  float* __restrict__ terms, int nparam_blocks, ScaleJob sj,
                                                  int32_t* __restrict__ info) {
  __shared__ float part[MAX_WORLD][5];
  __shared__ double tot[5];
  __shared__ int bad;
  const u64* mine = peers.buf[rank];
  const unsigned step = (unsigned)mine[0];                // written by k_dp_publish (previous launch on this stream)
  const int s = (int)(step & 1u);
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < 5 * world) {                     // the five ELBO sums of every rank (local polling)
    const int q = threadIdx.x / 5, j = threadIdx.x % 5;
    float v;
    if (!ld_ll(mine + ll_off(s, world, q, nf_pad, nparam + j), step, v)) bad = 1;
    part[q][j] = v;
  }
  __syncthreads();
  if (bad) { if (threadIdx.x == 0) atomicOr(info, 2); return; }
  if (threadIdx.x < 5) {
    double v = 0.0;
    for (int r = 0; r < world; ++r) v += (double)part[r][threadIdx.x];   // rank order, fp64
    tot[threadIdx.x] = v;
  }
  __syncthreads();
  const double nrm = tot[4] < 1.0 ? 1.0 : tot[4];
  const float scale = (float)(1.0 / nrm);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int j = 0; j < 5; ++j) terms[j] = (float)tot[j];
    terms[5] = (float)((tot[0] + tot[1] + tot[2] + tot[3]) / nrm);
    terms[6] = scale;
    terms[7] = 0.f;
  }
  if ((int)blockIdx.x < nparam_blocks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nparam) return;
    double v = 0.0;
    bool ok = true;
    for (int r = 0; r < world; ++r) {
      float x;
      ok = ld_ll(mine + ll_off(s, world, r, nf_pad, i), step, x) && ok;
      v += (double)x;
    }
    if (!ok) { atomicOr(info, 2); return; }
    const float f = (float)v * scale;
    int o = i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (o < ov.n[j]) { if (ov.p[j]) ov.p[j][o] = f; return; }
      o -= ov.n[j];
    }
    return;
  }
  const long nthreads = (long)(gridDim.x - nparam_blocks) * blockDim.x;
  const long tid = (long)(blockIdx.x - nparam_blocks) * blockDim.x + threadIdx.x;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float* p = sj.p[j];
    if (!p) continue;
    const long n4 = sj.n[j] >> 2;
    f4* p4 = reinterpret_cast<f4*>(p);
    for (long i = tid; i < n4; i += nthreads) { f4 v = p4[i]; v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale; p4[i] = v; }
    for (long i = (n4 << 2) + tid; i < sj.n[j]; i += nthreads) p[i] *= scale;
  }
}

thread_local char g_dp_err[256] = "";
int dp_fail(int code, const char* msg) { snprintf(g_dp_err, sizeof(g_dp_err), "%s", msg); return code; }

}  // namespace

struct kvae_dp_comm {
  int device, rank, world;
  size_t nfloats, nf_pad, bytes;
  u64* mine;
  void* opened[MAX_WORLD];
  DpPeers peers;
  bool connected;
};

namespace kvae {
bool kvae_dp_get_view(kvae_dp_comm* c, DpView* out) {
  if (!c || !c->connected) return false;
  static_assert(HDR_WORDS == KV_DP_HDR_WORDS, "header size");
  for (int r = 0; r < 16; ++r) out->buf[r] = (r < c->world) ? c->peers.buf[r] : nullptr;
  out->rank = c->rank; out->world = c->world; out->nf_pad = c->nf_pad; out->nparam = (int)c->nfloats;
  return true;
}
}  // namespace kvae

extern "C" {

const char* kvae_dp_last_error(void) { return g_dp_err; }

size_t kvae_dp_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int kvae_dp_create(int device, int rank, int world, size_t nfloats, kvae_dp_comm** out, void* handle_out) {
  if (!out || !handle_out || world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || nfloats == 0)
    return dp_fail(-1, "kvae_dp_create: bad argument");
  int prev = -1;
  cudaGetDevice(&prev);
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return dp_fail(-1, "kvae_dp_create: bad device");
  kvae_dp_comm* c = new (std::nothrow) kvae_dp_comm();
  if (!c) return dp_fail(-1, "out of host memory");
  cudaGetDevice(&c->device);
  c->rank = rank; c->world = world; c->nfloats = nfloats; c->connected = false;
  c->nf_pad = (nfloats + 5 + 15) & ~(size_t)15;
  c->bytes = sizeof(u64) * (HDR_WORDS + 2 * (size_t)world * c->nf_pad);
  cudaError_t e = cudaMalloc(&c->mine, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->mine, 0, c->bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->mine);
  if (prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) { if (c->mine) cudaFree(c->mine); delete c; return dp_fail((int)e, cudaGetErrorString(e)); }
  memcpy(handle_out, &h, sizeof(h));
  for (int r = 0; r < MAX_WORLD; ++r) { c->opened[r] = nullptr; c->peers.buf[r] = nullptr; }
  *out = c;
  return 0;
}

/* handles: world * kvae_dp_handle_bytes() bytes in rank order (this rank's own entry is ignored) */
int kvae_dp_connect(kvae_dp_comm* c, const void* handles) {
  if (!c || !handles) return dp_fail(-1, "kvae_dp_connect: null argument");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  int rc = 0;
  for (int r = 0; r < c->world && rc == 0; ++r) {
    if (r == c->rank) { c->peers.buf[r] = c->mine; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { rc = dp_fail((int)e, cudaGetErrorString(e)); (void)cudaGetLastError(); break; }
    c->opened[r] = p;
    c->peers.buf[r] = reinterpret_cast<u64*>(p);
  }
  if (prev >= 0) cudaSetDevice(prev);
  c->connected = (rc == 0);
  return rc;
}

int kvae_dp_destroy(kvae_dp_comm* c) {
  if (!c) return 0;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  for (int r = 0; r < c->world; ++r) if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
  if (c->mine) cudaFree(c->mine);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
  return 0;
}

/* see include/kvae_kalman.h */
int kvae_dp_finalize(const kvae_dims* d, kvae_dp_comm* c, const kvae_grads* g, float* terms, int32_t* info, void* stream) {
  if (!d || !c || !g || !terms || !info) return dp_fail(-1, "kvae_dp_finalize: null argument");
  if (!c->connected) return dp_fail(-1, "kvae_dp_finalize: communicator not connected");
  if (!g->dY || !g->dalpha || !g->dA || !g->dBm || !g->dC) return dp_fail(-1, "kvae_dp_finalize: null gradient buffer");
  const int nA = d->K * d->n * d->n, nB = d->K * d->n * d->m, nC = d->K * d->p * d->n, nQ = d->q_per_mode ? nA : 0;
  if (d->q_per_mode && !g->dQ) return dp_fail(-1, "kvae_dp_finalize: dQ required when q_per_mode");
  const int nparam = nA + nB + nC + nQ;
  if ((size_t)nparam != c->nfloats) return dp_fail(-1, "kvae_dp_finalize: communicator was created for another parameter count");
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != c->device) cudaSetDevice(c->device);
  cudaStream_t s = (cudaStream_t)stream;
  LocalVec lv{{g->dA, g->dBm, g->dC, g->dQ, terms}, {nA, nB, nC, nQ, 5}};
  k_dp_publish<<<1, 1024, 0, s>>>(c->peers, c->rank, c->world, c->nf_pad, lv);
  OutVec ov{{g->dA, g->dBm, g->dC, g->dQ}, {nA, nB, nC, nQ}};
  const long BT = (long)d->B * d->T;
  ScaleJob sj{{g->dY, g->dalpha, g->dU}, {BT * d->p, BT * d->K, g->dU ? BT * d->m : 0}};
  const int nparam_blocks = (nparam + 127) / 128;
  const long total4 = (sj.n[0] + sj.n[1] + sj.n[2]) / 4;
  int scale_blocks = (int)((total4 + 128 * 8 - 1) / (128 * 8));
  if (scale_blocks < 1) scale_blocks = 1;
  if (scale_blocks > 148 * 8) scale_blocks = 148 * 8;
  k_dp_final<<<nparam_blocks + scale_blocks, 128, 0, s>>>(c->peers, c->rank, c->world, c->nf_pad, nparam, ov, terms,
                                                          nparam_blocks, sj, info);
  const cudaError_t e = cudaGetLastError();
  if (prev != c->device && prev >= 0) cudaSetDevice(prev);
  if (e != cudaSuccess) return dp_fail((int)e, cudaGetErrorString(e));
  return 0;
}

}  // extern "C"

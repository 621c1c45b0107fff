// kvae_vae.cu — the VAE-side reductions that feed the same loss as the Kalman ELBO:
//   kvae/vae/losses.py:62-111  vae_loss  (masked Bernoulli / Gaussian pixel log-likelihood, log q(a|x), log p(a),
//                                         normalised by clamp(sum(mask), 1))
//   kvae/model/model.py:81-84  reparameterize  a = mu + eps * sqrt(var + 1e-6)
// The reference runs ~25 ATen ops (each a pass over the [B,T,C,H,W] frames or a tiny launch) and their autograd replay;
// here the value is ONE pass over the frames (128-bit loads, per-CTA fp64 partials, fixed-order final reduction ->
// deterministic) and the gradient ONE elementwise pass.  HBM bound: 8 B read per pixel forward, 8 B read + 4 B written
// backward.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <cstdio>
#include "../../include/kvae_kalman.h"

namespace {
thread_local char g_verr[256] = "";
int vfail(int code, const char* msg) { snprintf(g_verr, sizeof(g_verr), "%s", msg); return code; }

constexpr float KV_HALF_LOG2PI = 0.9189385332046727f;

__device__ __forceinline__ float px_bernoulli(float logit, float x) {   // -BCEWithLogits (losses.py:83-84)
  return -(fmaxf(logit, 0.f) - logit * x + log1pf(expf(-fabsf(logit))));
}
__device__ __forceinline__ float px_gauss(float mu, float x, float var, float half_log_var) {   // log_gaussian (losses.py:5-17)
  const float d = x - mu;
  return -KV_HALF_LOG2PI - half_log_var - d * d / (2.f * var);
}

// one CTA per block of frames; sums[4] per CTA: sum m*log p(x|a), sum m*log q(a|x), sum m*log p(a), sum m
template <bool BERN>
__global__ void __launch_bounds__(256) k_vae_fwd(const float* __restrict__ x, const float* __restrict__ xmu, float x_var,
                                                 const float* __restrict__ a, const float* __restrict__ amu,
                                                 const float* __restrict__ avar, const float* __restrict__ mask, int frames,
                                                 int D, int adim, double* __restrict__ partials) {
  __shared__ double red[8][4];
  const float hlv = 0.5f * logf(x_var);
  double s_px = 0.0, s_q = 0.0, s_p = 0.0, s_m = 0.0;
  for (int f = blockIdx.x; f < frames; f += gridDim.x) {
    const float m = mask ? mask[f] : 1.0f;
    const float* xf = x + (size_t)f * D;
    const float* lf = xmu + (size_t)f * D;
    float acc = 0.f;
    if ((D & 3) == 0) {
      const float4* x4 = reinterpret_cast<const float4*>(xf);
      const float4* l4 = reinterpret_cast<const float4*>(lf);
      for (int i = threadIdx.x; i < D / 4; i += blockDim.x) {
        const float4 xv = x4[i], lv = l4[i];
        if (BERN) acc += (px_bernoulli(lv.x, xv.x) + px_bernoulli(lv.y, xv.y)) + (px_bernoulli(lv.z, xv.z) + px_bernoulli(lv.w, xv.w));
        else acc += (px_gauss(lv.x, xv.x, x_var, hlv) + px_gauss(lv.y, xv.y, x_var, hlv)) + (px_gauss(lv.z, xv.z, x_var, hlv) + px_gauss(lv.w, xv.w, x_var, hlv));
      }
    } else {
      for (int i = threadIdx.x; i < D; i += blockDim.x) acc += BERN ? px_bernoulli(lf[i], xf[i]) : px_gauss(lf[i], xf[i], x_var, hlv);
    }
    s_px += (double)(acc * m);
    if (threadIdx.x < adim) {
      const size_t k = (size_t)f * adim + threadIdx.x;
      const float av = a[k], d = av - amu[k], v = avar[k];
      s_q += (double)(m * (-KV_HALF_LOG2PI - 0.5f * logf(v) - d * d / (2.f * v)));
      s_p += (double)(m * (-KV_HALF_LOG2PI - 0.5f * av * av));
    }
    if (threadIdx.x == 0) s_m += (double)m;
  }
  double v[4] = {s_px, s_q, s_p, s_m};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
  }
}

// out[0] = vae_elbo, [1] = recon_term, [2] = regularization_term, [3] = 1/denom, [4..7] = the four raw sums
__global__ void k_vae_final(const double* __restrict__ partials, int n, float scale_rec, float beta, float* __restrict__ out) {
  __shared__ double tot[4];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w < 4) {
    double v = 0.0;
    for (int i = lane; i < n; i += 32) v += partials[(size_t)i * 4 + w];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) tot[w] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double denom = tot[3] < 1.0 ? 1.0 : tot[3];                       // losses.py:81
    const double recon = tot[0] / denom, reg = (tot[2] - tot[1]) / denom;   // :103-105
    out[0] = (float)(scale_rec * recon + beta * reg);                       // :107
    out[1] = (float)recon; out[2] = (float)reg; out[3] = (float)(1.0 / denom);
    out[4] = (float)tot[0]; out[5] = (float)tot[1]; out[6] = (float)tot[2]; out[7] = (float)tot[3];
  }
}

// gradient of  g[0]*vae_elbo + g[1]*recon + g[2]*reg  w.r.t. x_mu, a, a_mu, a_var
template <bool BERN>
__global__ void __launch_bounds__(256) k_vae_bwd(const float* __restrict__ x, const float* __restrict__ xmu, float x_var,
                                                 const float* __restrict__ a, const float* __restrict__ amu,
                                                 const float* __restrict__ avar, const float* __restrict__ mask,
                                                 const float* __restrict__ g, const float* __restrict__ fwd_out, float scale_rec,
                                                 float beta, long npix, int D, long nlat, int adim, float* __restrict__ d_xmu,
                                                 float* __restrict__ d_a, float* __restrict__ d_amu, float* __restrict__ d_avar) {
  const float inv = fwd_out[3];
  const float c_px = (scale_rec * g[0] + g[1]) * inv;       // coefficient of sum m*log p(x|a)
  const float c_rg = (beta * g[0] + g[2]) * inv;            // coefficient of sum m*(log p(a) - log q(a|x))
  const long stride = (long)gridDim.x * blockDim.x;
  const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((D & 3) == 0) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* l4 = reinterpret_cast<const float4*>(xmu);
    float4* o4 = reinterpret_cast<float4*>(d_xmu);
    const int D4 = D / 4;
    for (long i = tid; i < npix / 4; i += stride) {
      const float cm = c_px * (mask ? mask[i / D4] : 1.0f);
      const float4 xv = x4[i], lv = l4[i];
      float4 o;
      if (BERN) {
        o.x = cm * (xv.x - 1.f / (1.f + expf(-lv.x))); o.y = cm * (xv.y - 1.f / (1.f + expf(-lv.y)));
        o.z = cm * (xv.z - 1.f / (1.f + expf(-lv.z))); o.w = cm * (xv.w - 1.f / (1.f + expf(-lv.w)));
      } else {
        const float iv = cm / x_var;
        o.x = iv * (xv.x - lv.x); o.y = iv * (xv.y - lv.y); o.z = iv * (xv.z - lv.z); o.w = iv * (xv.w - lv.w);
      }
      o4[i] = o;
    }
  } else {
    for (long i = tid; i < npix; i += stride) {
      const float cm = c_px * (mask ? mask[i / D] : 1.0f);
      d_xmu[i] = BERN ? cm * (x[i] - 1.f / (1.f + expf(-xmu[i]))) : cm * (x[i] - xmu[i]) / x_var;
    }
  }
  for (long k = tid; k < nlat; k += stride) {
    const float cm = c_rg * (mask ? mask[k / adim] : 1.0f);
    const float av = a[k], d = av - amu[k], v = avar[k];
    // d/d. of m*(log p(a) - log q(a|x)),  log q = c - log(v)/2 - d^2/(2v),  log p = c - a^2/2
    d_a[k] = cm * (-av + d / v);
    d_amu[k] = cm * (-d / v);
    d_avar[k] = cm * (0.5f / v - d * d / (2.f * v * v));
  }
}

// a = mu + eps * sqrt(var + 1e-6)  (model.py:81-84);  backward: dmu = g, dvar = g * eps / (2 sqrt(var + 1e-6))
__global__ void k_reparam_fwd(const float* __restrict__ mu, const float* __restrict__ var, const float* __restrict__ eps, long n,
                              float* __restrict__ a) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    a[i] = mu[i] + eps[i] * sqrtf(var[i] + 1e-6f);
}
__global__ void k_reparam_bwd(const float* __restrict__ var, const float* __restrict__ eps, const float* __restrict__ g, long n,
                              float* __restrict__ dvar) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
    dvar[i] = g[i] * eps[i] * 0.5f * rsqrtf(var[i] + 1e-6f);
}

int grid_for(long work, int tpb, int cap) {
  long g = (work + tpb - 1) / tpb;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}
}  // namespace

extern "C" {

const char* kvae_vae_last_error(void) { return g_verr; }

size_t kvae_vae_loss_workspace_bytes(const kvae_vae_dims* d) {
  if (!d) return 0;
  return sizeof(double) * 4 * (size_t)grid_for(d->frames, 1, 148 * 8);
}

int kvae_vae_loss_fwd(const kvae_vae_dims* d, const float* x, const float* x_mu, const float* a, const float* a_mu,
                      const float* a_var, const float* mask, float* out8, void* workspace, int device, void* stream) {
  if (!d || !x || !x_mu || !a || !a_mu || !a_var || !out8 || !workspace) return vfail(-1, "null argument");
  if (d->frames <= 0 || d->pixels <= 0 || d->a_dim <= 0 || d->a_dim > 256) return vfail(-1, "bad dims (a_dim must be 1..256)");
  if (!(d->x_var > 0.f) && !d->bernoulli) return vfail(-1, "x_var must be positive");
  int prev = -1;
  if (device >= 0) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); }
  (void)cudaGetLastError();
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = grid_for(d->frames, 1, 148 * 8);
  double* part = reinterpret_cast<double*>(workspace);
  const float xv = d->bernoulli ? 1.0f : d->x_var;
  if (d->bernoulli) k_vae_fwd<true><<<grid, 256, 0, s>>>(x, x_mu, xv, a, a_mu, a_var, mask, d->frames, d->pixels, d->a_dim, part);
  else k_vae_fwd<false><<<grid, 256, 0, s>>>(x, x_mu, xv, a, a_mu, a_var, mask, d->frames, d->pixels, d->a_dim, part);
  k_vae_final<<<1, 128, 0, s>>>(part, grid, d->scale_reconstruction, d->beta, out8);
  const cudaError_t e = cudaGetLastError();
  if (device >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? 0 : vfail((int)e, cudaGetErrorString(e));
}

int kvae_vae_loss_bwd(const kvae_vae_dims* d, const float* x, const float* x_mu, const float* a, const float* a_mu,
                      const float* a_var, const float* mask, const float* g3, const float* out8, float* d_x_mu, float* d_a,
                      float* d_a_mu, float* d_a_var, int device, void* stream) {
  if (!d || !x || !x_mu || !a || !a_mu || !a_var || !g3 || !out8 || !d_x_mu || !d_a || !d_a_mu || !d_a_var) return vfail(-1, "null argument");
  int prev = -1;
  if (device >= 0) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); }
  (void)cudaGetLastError();
  cudaStream_t s = (cudaStream_t)stream;
  const long npix = (long)d->frames * d->pixels, nlat = (long)d->frames * d->a_dim;
  const int grid = grid_for(npix / 4 + 1, 256, 148 * 16);
  const float xv = d->bernoulli ? 1.0f : d->x_var;
  if (d->bernoulli) k_vae_bwd<true><<<grid, 256, 0, s>>>(x, x_mu, xv, a, a_mu, a_var, mask, g3, out8, d->scale_reconstruction, d->beta, npix, d->pixels, nlat, d->a_dim, d_x_mu, d_a, d_a_mu, d_a_var);
  else k_vae_bwd<false><<<grid, 256, 0, s>>>(x, x_mu, xv, a, a_mu, a_var, mask, g3, out8, d->scale_reconstruction, d->beta, npix, d->pixels, nlat, d->a_dim, d_x_mu, d_a, d_a_mu, d_a_var);
  const cudaError_t e = cudaGetLastError();
  if (device >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? 0 : vfail((int)e, cudaGetErrorString(e));
}

int kvae_vae_reparam_fwd(const float* mu, const float* var, const float* eps, long n, float* a, int device, void* stream) {
  if (!mu || !var || !eps || !a || n <= 0) return vfail(-1, "null argument");
  int prev = -1;
  if (device >= 0) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); }
  (void)cudaGetLastError();
  k_reparam_fwd<<<grid_for(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(mu, var, eps, n, a);
  const cudaError_t e = cudaGetLastError();
  if (device >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? 0 : vfail((int)e, cudaGetErrorString(e));
}

int kvae_vae_reparam_bwd(const float* var, const float* eps, const float* g, long n, float* d_var, int device, void* stream) {
  if (!var || !eps || !g || !d_var || n <= 0) return vfail(-1, "null argument");
  int prev = -1;
  if (device >= 0) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); }
  (void)cudaGetLastError();
  k_reparam_bwd<<<grid_for(n, 256, 148 * 8), 256, 0, (cudaStream_t)stream>>>(var, eps, g, n, d_var);
  const cudaError_t e = cudaGetLastError();
  if (device >= 0 && prev != device) cudaSetDevice(prev);
  return e == cudaSuccess ? 0 : vfail((int)e, cudaGetErrorString(e));
}

}  // extern "C"

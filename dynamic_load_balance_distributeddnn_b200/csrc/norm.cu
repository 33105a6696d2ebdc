// GroupNorm (+ReLU, +residual) forward / backward for NHWC activations, bf16 or fp32.
//
// Replaces the reference's per-layer ATen native_group_norm + relu (+ add) launches
// (reference Net/Densenet.py:18-19,31; Net/Resnet.py:50-54).
//
// Both directions share one structure:
//   (1) nc_reduce2 : per-(sample, channel) reduction of two quantities over H*W
//         fwd: (sum x, sum x^2)          bwd: (sum dz, sum dz*x),  dz = dy * [y > 0]
//       -> a tiny fp32 table [N][C][2]; group statistics for ANY grouping of channels are sums
//       of table entries, so DenseNet's growing concat buffer never has to be re-read for stats.
//   (2) an apply pass with per-(sample, channel) affine coefficients precomputed in shared memory
//         fwd: y  = act(a[c]*x + b[c] (+ residual))
//         bwd: dx = k1[c]*dz + k2[c]*x + k3[c]   (optionally accumulated into dx)
// Rows are pixels (n, h, w); `ld*` are row strides in elements so channel slices of a wider
// buffer can be read / written in place.
#include "common.cuh"
#include <stdlib.h>

static int g_skip_zero = 0;      // callers that pre-zero one big arena set this to skip the per-call memsets
// rows in flight per thread (independent 16-byte loads issued before the first use) of the streaming kernels; tunable at run
// time for sweeps (dlb_norm_tune / DLB_GN_UNR_{RED,FWD,BWD}); defaults are the measured optimum on B200
static int g_unr_red = 0, g_unr_fwd = 0, g_unr_bwd = 0, g_min_kb = 0;
static int g_bulk_on = -1, g_bulk_kb = 64;   // fused backward: bulk-copy flavour (DLB_GN_BWD_BULK / DLB_GN_BULK_KB, dlb_norm_bulk)
static int unr_env(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return (v == 1 || v == 2 || v == 4) ? v : dflt;
}
static void unr_init() {
  if (g_unr_red) return;
  // sweep on B200 (profiles/r2_03_gn_sweep.txt): one row in flight per thread wins everywhere -- more rows cost registers,
  // i.e. resident warps, and the kernels are occupancy- not MLP-limited (backward reduce 97 us at 1 row, 124 at 2, 183 at 4)
  g_unr_red = unr_env("DLB_GN_UNR_RED", 1);
  g_unr_fwd = unr_env("DLB_GN_UNR_FWD", 1);
  g_unr_bwd = unr_env("DLB_GN_UNR_BWD", 1);
}

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------------------
// (1) per-(n, c) reduction
// MODE 0: q0 = x, q1 = x*x                    (inputs: x)
// MODE 1: q0 = dz, q1 = dz*x, dz = dy*(y>0)   (inputs: x, dy, y)  [relu]
// MODE 2: q0 = dy, q1 = dy*x                  (inputs: x, dy)     [no relu]
// MODE 3: as MODE 1 but the ReLU mask is recomputed as (ca[n,c]*x + cb[n,c] > 0) from the forward's affine
//         coefficients, so the normalised activation y never has to exist in memory
template <typename T, int V, int MODE, int UNR>
__global__ void __launch_bounds__(kThreads)
nc_reduce2_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                  const T* __restrict__ y, int64_t ldy, float* __restrict__ table, int64_t table_ns,
                  int HW, int C, int rows_per_block, const float* __restrict__ mean,
                  const float* __restrict__ rstd, int G, float* __restrict__ dgamma, float* __restrict__ dbeta,
                  const float* __restrict__ coef_a, const float* __restrict__ coef_b, int64_t coef_ld,
                  T* __restrict__ copy_dst, int64_t ldc) {
  dlb_pdl_wait();
  extern __shared__ float smem[];            // [2*C]
  const int n = blockIdx.y;
  const int lanes = C / V;                   // channel-vector lanes
  const int row_lanes = max(1, kThreads / lanes);
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) smem[i] = 0.f;
  __syncthreads();
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  for (int lane0 = 0; lane0 < lanes; lane0 += kThreads) {      // loop only when C/V > 256
    const int lane = lane0 + (threadIdx.x % min(lanes, kThreads));
    const int rl = threadIdx.x / min(lanes, kThreads);
    if (lane < lanes && rl < row_lanes) {
      float a0[V], a1[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
      const int c = lane * V;
      float ka[V], kb[V];
      if constexpr (MODE == 3) {
#pragma unroll
        for (int i = 0; i < V; ++i) { ka[i] = coef_a[(int64_t)n * coef_ld + c + i]; kb[i] = coef_b[(int64_t)n * coef_ld + c + i]; }
      }
      for (int rb = r0 + rl; rb < r1; rb += row_lanes * UNR) {
        float xv[UNR][V], gv[UNR][V], yv[UNR][V];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            const int64_t row = (int64_t)n * HW + r;
            load_vec<T, V>(x + row * ldx + c, xv[u]);
            if constexpr (MODE == 0) { if (copy_dst != nullptr) store_vec<T, V>(copy_dst + row * ldc + c, xv[u]); }
            if constexpr (MODE != 0) load_vec<T, V>(dy + row * lddy + c, gv[u]);
            if constexpr (MODE == 1) load_vec<T, V>(y + row * ldy + c, yv[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            if constexpr (MODE == 0) {
#pragma unroll
              for (int i = 0; i < V; ++i) { a0[i] += xv[u][i]; a1[i] += xv[u][i] * xv[u][i]; }
            } else {
#pragma unroll
              for (int i = 0; i < V; ++i) {
                float gz = gv[u][i];
                if constexpr (MODE == 1) gz = yv[u][i] > 0.f ? gz : 0.f;
                if constexpr (MODE == 3) gz = fmaf(ka[i], xv[u][i], kb[i]) > 0.f ? gz : 0.f;
                a0[i] += gz; a1[i] += gz * xv[u][i];
              }
            }
          }
        }
      }
      // lanes of a warp that own the same channels sit `lanes` apart: fold them first
      bool owner = true;
      if (lanes < 32 && (lanes & (lanes - 1)) == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          for (int o = lanes; o < 32; o <<= 1) {
            a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
            a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
          }
        }
        owner = (threadIdx.x & 31) < lanes;
      }
      if (owner) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
          atomicAdd(&smem[2 * (c + i)], a0[i]);
          atomicAdd(&smem[2 * (c + i) + 1], a1[i]);
        }
      }
    }
  }
  __syncthreads();
  float* dst = table + (int64_t)n * table_ns;
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) atomicAdd(&dst[i], smem[i]);
  if constexpr (MODE != 0) {
    // affine-parameter gradients are linear in the per-(n,c) sums: fold them in here instead of a
    // separate pass over the table (dbeta = sum A, dgamma = sum rstd*(B - mu*A))
    if (dgamma != nullptr) {
      const int cpg = C / G;
      for (int c = threadIdx.x; c < C; c += kThreads) {
        const float A = smem[2 * c], B = smem[2 * c + 1];
        const int g = c / cpg;
        atomicAdd(&dbeta[c], A);
        atomicAdd(&dgamma[c], rstd[n * G + g] * (B - mean[n * G + g] * A));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// group statistics from the table:  mean/rstd [N][G]
__global__ void gn_finalize_kernel(const float* __restrict__ table, int64_t table_ns, float* __restrict__ mean,
                                   float* __restrict__ rstd, int N, int C, int G, int HW, float eps) {
  dlb_pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * G) return;
  const int n = idx / G, g = idx % G, cpg = C / G;
  const float* t = table + (int64_t)n * table_ns + (int64_t)g * cpg * 2;
  float s = 0.f, ss = 0.f;
  for (int i = 0; i < cpg; ++i) { s += t[2 * i]; ss += t[2 * i + 1]; }
  const float m = 1.f / ((float)cpg * (float)HW);
  const float mu = s * m;
  const float var = fmaxf(ss * m - mu * mu, 0.f);
  mean[idx] = mu;
  rstd[idx] = rsqrtf(var + eps);
}

// group statistics + per-(sample, channel) affine coefficients for a fused GN prologue:
//   a[n][c] = gamma[c]*rstd,  b[n][c] = beta[c] - mean*a   (rows padded to `ld` floats, zero filled)
__global__ void gn_coeff_kernel(const float* __restrict__ table, int64_t table_ns, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ mean, float* __restrict__ rstd,
                                float* __restrict__ ca, float* __restrict__ cb, int64_t ld, int C, int G, int HW, float eps) {
  dlb_pdl_wait();
  extern __shared__ float sm[];              // mu[G], rs[G]
  const int n = blockIdx.x, cpg = C / G;
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float* t = table + (int64_t)n * table_ns + (int64_t)g * cpg * 2;
    float s = 0.f, ss = 0.f;
    for (int i = 0; i < cpg; ++i) { s += t[2 * i]; ss += t[2 * i + 1]; }
    const float m = 1.f / ((float)cpg * (float)HW);
    const float mu = s * m;
    const float r = rsqrtf(fmaxf(ss * m - mu * mu, 0.f) + eps);
    sm[g] = mu; sm[G + g] = r;
    mean[n * G + g] = mu; rstd[n * G + g] = r;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    if (c < C) { const int g = c / cpg; a = gamma[c] * sm[G + g]; b = beta[c] - sm[g] * a; }
    ca[(int64_t)n * ld + c] = a;
    cb[(int64_t)n * ld + c] = b;
  }
}

// ---------------------------------------------------------------------------------------------
// (2a) forward apply: y = act(gamma*(x-mu)*rstd + beta (+ res))
// `tab` (optional): derive the group statistics from the per-(sample, channel) (sum, sumsq) table inside the kernel instead of
// reading precomputed mean/rstd -- folds the finalize/coefficient kernel into the apply pass; block x == 0 of each sample then
// also publishes mean/rstd and the affine coefficient rows (coef_a, coef_b) the backward pass needs.
struct FwdFromTable {
  const float* tab; int64_t tab_ns; float eps;
  float* mean_out; float* rstd_out; float* coef_a; float* coef_b; int64_t coef_ld;
};

template <typename T, int V, bool RELU, bool RES, int UNR>
__global__ void __launch_bounds__(kThreads)
gn_fwd_apply_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ res, int64_t ldr,
                    T* __restrict__ y, int64_t ldy, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ rstd, int HW, int C, int G, int rows_per_block, const FwdFromTable ft) {
  dlb_pdl_wait();
  extern __shared__ float smem[];            // a[C], b[C] (+ mu[G], rs[G] with ft.tab)
  float* sa = smem;
  float* sb = smem + C;
  const int n = blockIdx.y, cpg = C / G;
  if (ft.tab != nullptr) {
    float* smu = smem + 2 * C;
    float* srs = smu + G;
    for (int g = threadIdx.x; g < G; g += kThreads) {
      const float* t = ft.tab + (int64_t)n * ft.tab_ns + (int64_t)g * cpg * 2;
      float s = 0.f, ss = 0.f;
      for (int i = 0; i < cpg; ++i) { s += t[2 * i]; ss += t[2 * i + 1]; }
      const float m = 1.f / ((float)cpg * (float)HW);
      const float mu = s * m;
      const float r = rsqrtf(fmaxf(ss * m - mu * mu, 0.f) + ft.eps);
      smu[g] = mu; srs[g] = r;
      if (blockIdx.x == 0) { ft.mean_out[n * G + g] = mu; ft.rstd_out[n * G + g] = r; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += kThreads) {
      const int g = c / cpg;
      const float a = gamma[c] * srs[g];
      sa[c] = a;
      sb[c] = beta[c] - smu[g] * a;
    }
    if (blockIdx.x == 0 && ft.coef_a != nullptr) {
      for (int c = threadIdx.x; c < ft.coef_ld; c += kThreads) {
        float a = 0.f, b = 0.f;
        if (c < C) { const int g = c / cpg; a = gamma[c] * srs[g]; b = beta[c] - smu[g] * a; }
        ft.coef_a[(int64_t)n * ft.coef_ld + c] = a;
        ft.coef_b[(int64_t)n * ft.coef_ld + c] = b;
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += kThreads) {
      const int g = c / cpg;
      const float r = rstd[n * G + g], mu = mean[n * G + g];
      const float a = gamma[c] * r;
      sa[c] = a;
      sb[c] = beta[c] - mu * a;
    }
  }
  __syncthreads();
  const int lanes = C / V;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  if (lanes <= kThreads) {
    // one channel-vector per thread for the whole block: coefficients live in registers, rows are
    // streamed UNR at a time so several 16-byte loads are in flight per thread
    const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes, row_lanes = kThreads / lanes;
    if (rl < row_lanes) {
      const int c = lane * V;
      float a[V], b[V];
#pragma unroll
      for (int k = 0; k < V; ++k) { a[k] = sa[c + k]; b[k] = sb[c + k]; }
      for (int rb = r0 + rl; rb < r1; rb += row_lanes * UNR) {
        float xv[UNR][V], rv[UNR][V];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            const int64_t row = (int64_t)n * HW + r;
            load_vec<T, V>(x + row * ldx + c, xv[u]);
            if constexpr (RES) load_vec<T, V>(res + row * ldr + c, rv[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            float out[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
              out[k] = fmaf(a[k], xv[u][k], b[k]);
              if constexpr (RES) out[k] += rv[u][k];
              if constexpr (RELU) out[k] = fmaxf(out[k], 0.f);
            }
            store_vec<T, V>(y + ((int64_t)n * HW + r) * ldy + c, out);
          }
        }
      }
    }
    return;
  }
  const int64_t total = (int64_t)(r1 - r0) * lanes;
  for (int64_t i = threadIdx.x; i < total; i += kThreads) {
    const int r = r0 + (int)(i / lanes);
    const int c = (int)(i % lanes) * V;
    const int64_t row = (int64_t)n * HW + r;
    float xv[V], out[V];
    load_vec<T, V>(x + row * ldx + c, xv);
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = fmaf(sa[c + k], xv[k], sb[c + k]);
    if constexpr (RES) {
      float rv[V];
      load_vec<T, V>(res + row * ldr + c, rv);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] += rv[k];
    }
    if constexpr (RELU) {
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] = fmaxf(out[k], 0.f);
    }
    store_vec<T, V>(y + row * ldy + c, out);
  }
}

// ---------------------------------------------------------------------------------------------
// (2b) backward apply: dx (+)= k1[c]*dz + k2[c]*x + k3[c];  optional dres = dz
// RELU: 0 = no activation, 1 = mask from the saved output y, 2 = mask recomputed from (coef_a, coef_b)
template <typename T, int V, int RELU, bool RES, bool ACC, int UNR>
__global__ void __launch_bounds__(kThreads)
gn_bwd_apply_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy,
                    const T* __restrict__ y, int64_t ldy, T* __restrict__ dx, int64_t lddx,
                    T* __restrict__ dres, int64_t lddr, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ table, int64_t table_ns, int HW, int C, int G, int rows_per_block,
                    const float* __restrict__ coef_a, const float* __restrict__ coef_b, int64_t coef_ld,
                    float* __restrict__ dgamma, float* __restrict__ dbeta) {
  dlb_pdl_wait();
  extern __shared__ float smem[];            // k1[C], k2[C], k3[C], s1[G], s2[G]
  float* k1 = smem;
  float* k2 = smem + C;
  float* k3 = smem + 2 * C;
  float* s1 = smem + 3 * C;
  float* s2 = s1 + G;
  const int n = blockIdx.y, cpg = C / G;
  const float* t = table + (int64_t)n * table_ns;
  if (dgamma != nullptr && blockIdx.x == 0) {
    // affine-parameter gradients from the sample's complete sums: one block per sample (N atomics per channel)
    for (int c = threadIdx.x; c < C; c += kThreads) {
      const float A = t[2 * c], B = t[2 * c + 1];
      const int g = c / cpg;
      atomicAdd(&dbeta[c], A);
      atomicAdd(&dgamma[c], rstd[n * G + g] * (B - mean[n * G + g] * A));
    }
  }
  for (int g = threadIdx.x; g < G; g += kThreads) {
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float a = 0.f, b = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int c = g * cpg + i;
      const float A = t[2 * c], B = t[2 * c + 1];
      a += gamma[c] * A;
      b += gamma[c] * r * (B - mu * A);
    }
    s1[g] = a;
    s2[g] = b;
  }
  __syncthreads();
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  for (int c = threadIdx.x; c < C; c += kThreads) {
    const int g = c / cpg;
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    k1[c] = gamma[c] * r;
    const float q = r * r * s2[g] * inv_m;
    k2[c] = -q;
    k3[c] = -r * s1[g] * inv_m + q * mu;
  }
  __syncthreads();
  const int lanes = C / V;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  if (lanes <= kThreads) {
    const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes, row_lanes = kThreads / lanes;
    if (rl < row_lanes) {
      const int c = lane * V;
      float q1[V], q2[V], q3[V], ka[V], kb[V];
#pragma unroll
      for (int k = 0; k < V; ++k) { q1[k] = k1[c + k]; q2[k] = k2[c + k]; q3[k] = k3[c + k]; }
      if constexpr (RELU == 2) {
#pragma unroll
        for (int k = 0; k < V; ++k) { ka[k] = coef_a[(int64_t)n * coef_ld + c + k]; kb[k] = coef_b[(int64_t)n * coef_ld + c + k]; }
      }
      for (int rb = r0 + rl; rb < r1; rb += row_lanes * UNR) {
        float xv[UNR][V], gv[UNR][V], yv[UNR][V], old[UNR][V];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            const int64_t row = (int64_t)n * HW + r;
            load_vec<T, V>(x + row * ldx + c, xv[u]);
            load_vec<T, V>(dy + row * lddy + c, gv[u]);
            if constexpr (RELU == 1) load_vec<T, V>(y + row * ldy + c, yv[u]);
            if constexpr (ACC) load_vec<T, V>(dx + row * lddx + c, old[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
          const int r = rb + u * row_lanes;
          if (r < r1) {
            const int64_t row = (int64_t)n * HW + r;
            float out[V];
#pragma unroll
            for (int k = 0; k < V; ++k) {
              float gz = gv[u][k];
              if constexpr (RELU == 1) gz = yv[u][k] > 0.f ? gz : 0.f;
              if constexpr (RELU == 2) gz = fmaf(ka[k], xv[u][k], kb[k]) > 0.f ? gz : 0.f;
              gv[u][k] = gz;
              out[k] = fmaf(q1[k], gz, fmaf(q2[k], xv[u][k], q3[k]));
              if constexpr (ACC) out[k] += old[u][k];
            }
            if constexpr (RES) store_vec<T, V>(dres + row * lddr + c, gv[u]);
            store_vec<T, V>(dx + row * lddx + c, out);
          }
        }
      }
    }
    return;
  }
  const int64_t total = (int64_t)(r1 - r0) * lanes;
  for (int64_t i = threadIdx.x; i < total; i += kThreads) {
    const int r = r0 + (int)(i / lanes);
    const int c = (int)(i % lanes) * V;
    const int64_t row = (int64_t)n * HW + r;
    float xv[V], gv[V], out[V];
    load_vec<T, V>(x + row * ldx + c, xv);
    load_vec<T, V>(dy + row * lddy + c, gv);
    if constexpr (RELU == 1) {
      float yv[V];
      load_vec<T, V>(y + row * ldy + c, yv);
#pragma unroll
      for (int k = 0; k < V; ++k) gv[k] = yv[k] > 0.f ? gv[k] : 0.f;
    }
    if constexpr (RELU == 2) {
#pragma unroll
      for (int k = 0; k < V; ++k)
        gv[k] = fmaf(coef_a[(int64_t)n * coef_ld + c + k], xv[k], coef_b[(int64_t)n * coef_ld + c + k]) > 0.f ? gv[k] : 0.f;
    }
    if constexpr (RES) store_vec<T, V>(dres + row * lddr + c, gv);
#pragma unroll
    for (int k = 0; k < V; ++k) out[k] = fmaf(k1[c + k], gv[k], fmaf(k2[c + k], xv[k], k3[c + k]));
    if constexpr (ACC) {
      float old[V];
      load_vec<T, V>(dx + row * lddx + c, old);
#pragma unroll
      for (int k = 0; k < V; ++k) out[k] += old[k];
    }
    store_vec<T, V>(dx + row * lddx + c, out);
  }
}

// ---------------------------------------------------------------------------------------------
// (1)+(2b) in ONE launch: GroupNorm(+ReLU) backward with the ReLU mask recomputed from the forward coefficients.
//   phase 1: per-(n, c) sums (sum dz, sum dz*x) of this block's rows -> table[n] (atomics), affine-parameter gradients
//   per-sample barrier: the `gridDim.x` blocks of sample n count themselves in done[n] and wait for each other (they are
//            launched back to back and are all resident long before the first one gets here; watchdog instead of a hang)
//   phase 2: dx (+)= k1*dz + k2*x + k3 over the same rows -- x and dy were read a few microseconds ago by this very block
// Saves a kernel launch per GroupNorm backward (two per DenseNet layer) and turns the second pass into cache hits.
// STAGE: the block's rows of x and dy stay in shared memory between the two phases (16-byte vectors, [row][C]); phase 2 then
// reads nothing but dX from global memory -- 4 instead of 6 passes over |x| once the working set no longer fits L2.
template <typename T, int V, bool ACC, bool STAGE>
__global__ void __launch_bounds__(kThreads)
gn_bwd_fused_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy, T* __restrict__ dx, int64_t lddx,
                    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                    float* __restrict__ table, int64_t table_ns, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    const float* __restrict__ coef_a, const float* __restrict__ coef_b, int64_t coef_ld,
                    unsigned* __restrict__ done, int HW, int C, int G, int rows_per_block) {
  dlb_pdl_wait();
  extern __shared__ float smem[];            // phase 1: acc[2C];  phase 2: k1[C], k2[C], k3[C], s1[G], s2[G]
  const int n = blockIdx.y, cpg = C / G;
  const int lanes = C / V;                   // <= kThreads (checked by the host)
  const int row_lanes = kThreads / lanes;
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const bool active = rl < row_lanes;
  const int c = lane * V;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  // staging area behind the coefficient arrays: x rows, then dy rows, C * sizeof(T) bytes each
  uint8_t* stage_x = reinterpret_cast<uint8_t*>(smem) + (((3 * C + 2 * G) * 4 + 15) & ~15);
  uint8_t* stage_g = stage_x + (size_t)rows_per_block * C * sizeof(T);
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) smem[i] = 0.f;
  float ka[V], kb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { ka[i] = coef_a[(int64_t)n * coef_ld + c + i]; kb[i] = coef_b[(int64_t)n * coef_ld + c + i]; }
  __syncthreads();
  // ---- phase 1 ----
  if (active) {
    float a0[V], a1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
    constexpr int UNR = 1;                   // one row in flight per thread: registers -> occupancy beats MLP (profiles/r2_03_gn_sweep.txt)
    for (int rb = r0 + rl; rb < r1; rb += row_lanes * UNR) {
      float xv[UNR][V], gv[UNR][V];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int r = rb + u * row_lanes;
        if (r < r1) {
          const int64_t row = (int64_t)n * HW + r;
          if constexpr (STAGE) {
            const uint4 xr = *reinterpret_cast<const uint4*>(x + row * ldx + c);
            const uint4 gr = *reinterpret_cast<const uint4*>(dy + row * lddy + c);
            const size_t so = ((size_t)(r - r0) * C + c) * sizeof(T);
            *reinterpret_cast<uint4*>(stage_x + so) = xr;
            *reinterpret_cast<uint4*>(stage_g + so) = gr;
            load_vec<T, V>(reinterpret_cast<const T*>(&xr), xv[u]);
            load_vec<T, V>(reinterpret_cast<const T*>(&gr), gv[u]);
          } else {
            load_vec<T, V>(x + row * ldx + c, xv[u]);
            load_vec<T, V>(dy + row * lddy + c, gv[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int r = rb + u * row_lanes;
        if (r < r1) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float gz = fmaf(ka[i], xv[u][i], kb[i]) > 0.f ? gv[u][i] : 0.f;
            a0[i] += gz; a1[i] += gz * xv[u][i];
          }
        }
      }
    }
    bool owner = true;
    if (lanes < 32 && (lanes & (lanes - 1)) == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        for (int o = lanes; o < 32; o <<= 1) {
          a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
          a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
        }
      }
      owner = (threadIdx.x & 31) < lanes;
    }
    if (owner) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        atomicAdd(&smem[2 * (c + i)], a0[i]);
        atomicAdd(&smem[2 * (c + i) + 1], a1[i]);
      }
    }
  }
  __syncthreads();
  float* tab = table + (int64_t)n * table_ns;
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) atomicAdd(&tab[i], smem[i]);
  // ---- per-sample barrier ----
  __shared__ unsigned s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_ticket = atomicAdd(&done[n], 1u);
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done + n) : "memory");
      if (v >= gridDim.x) break;
      if (++spins > (1u << 24)) __trap();      // watchdog: a sibling block never arrived
    }
  }
  __syncthreads();
  // affine-parameter gradients: dbeta[c] += A, dgamma[c] += rstd*(B - mean*A) with the sample's COMPLETE sums, by the one block
  // that arrived last -- N atomics per channel instead of N x chunks (every block adding its partials made these 2C addresses
  // the bottleneck at small batches: 300+ blocks hammering ~1 000 floats; ncu r2_03: 32 us for a 20 MB problem)
  if (dgamma != nullptr && s_ticket == gridDim.x - 1) {
    for (int cc = threadIdx.x; cc < C; cc += kThreads) {
      const float A = __ldcg(tab + 2 * cc), B = __ldcg(tab + 2 * cc + 1);
      const int g = cc / cpg;
      atomicAdd(&dbeta[cc], A);
      atomicAdd(&dgamma[cc], rstd[n * G + g] * (B - mean[n * G + g] * A));
    }
  }
  // ---- phase 2 ----
  float* k1 = smem;
  float* k2 = smem + C;
  float* k3 = smem + 2 * C;
  float* s1 = smem + 3 * C;
  float* s2 = s1 + G;
  for (int g = threadIdx.x; g < G; g += kThreads) {
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float a = 0.f, b = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int cc = g * cpg + i;
      const float A = __ldcg(tab + 2 * cc), B = __ldcg(tab + 2 * cc + 1);
      a += gamma[cc] * A;
      b += gamma[cc] * r * (B - mu * A);
    }
    s1[g] = a;
    s2[g] = b;
  }
  __syncthreads();
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  for (int cc = threadIdx.x; cc < C; cc += kThreads) {
    const int g = cc / cpg;
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    k1[cc] = gamma[cc] * r;
    const float q = r * r * s2[g] * inv_m;
    k2[cc] = -q;
    k3[cc] = -r * s1[g] * inv_m + q * mu;
  }
  __syncthreads();
  if (!active) return;
  float q1[V], q2[V], q3[V];
#pragma unroll
  for (int k = 0; k < V; ++k) { q1[k] = k1[c + k]; q2[k] = k2[c + k]; q3[k] = k3[c + k]; }
  constexpr int UNR2 = 1;
  for (int rb = r0 + rl; rb < r1; rb += row_lanes * UNR2) {
    float xv[UNR2][V], gv[UNR2][V], old[UNR2][V];
#pragma unroll
    for (int u = 0; u < UNR2; ++u) {
      const int r = rb + u * row_lanes;
      if (r < r1) {
        const int64_t row = (int64_t)n * HW + r;
        if constexpr (STAGE) {
          const size_t so = ((size_t)(r - r0) * C + c) * sizeof(T);
          load_vec<T, V>(reinterpret_cast<const T*>(stage_x + so), xv[u]);
          load_vec<T, V>(reinterpret_cast<const T*>(stage_g + so), gv[u]);
        } else {
          load_vec<T, V>(x + row * ldx + c, xv[u]);
          load_vec<T, V>(dy + row * lddy + c, gv[u]);
        }
        if constexpr (ACC) load_vec<T, V>(dx + row * lddx + c, old[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR2; ++u) {
      const int r = rb + u * row_lanes;
      if (r < r1) {
        float out[V];
#pragma unroll
        for (int k = 0; k < V; ++k) {
          const float gz = fmaf(ka[k], xv[u][k], kb[k]) > 0.f ? gv[u][k] : 0.f;
          out[k] = fmaf(q1[k], gz, fmaf(q2[k], xv[u][k], q3[k]));
          if constexpr (ACC) out[k] += old[u][k];
        }
        store_vec<T, V>(dx + ((int64_t)n * HW + r) * lddx + c, out);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bulk-copy (TMA unit, cp.async.bulk) flavour of the fused backward: the block's rows of x, dy (and dX when accumulating) are
// fetched by asynchronous bulk copies into shared memory -- one copy per row per tensor, issued up front by one warp and
// tracked by a single mbarrier -- so the bytes in flight are bounded by the tile (tens of KB per block) instead of by
// registers x resident warps (2 x 16 B per thread in the register flavour: ~64 KB per SM at full occupancy, less than the
// ~52 KB x latency product HBM3e needs once anything else limits occupancy).  Both phases then compute out of shared memory,
// the dX tile is updated in place and written back with bulk stores: x and dy cross the memory system once, not twice.
__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gn_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gn_bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <typename T, int V, bool ACC>
__global__ void __launch_bounds__(kThreads)
gn_bwd_bulk_kernel(const T* __restrict__ x, int64_t ldx, const T* __restrict__ dy, int64_t lddy, T* __restrict__ dx, int64_t lddx,
                   const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                   float* __restrict__ table, int64_t table_ns, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   const float* __restrict__ coef_a, const float* __restrict__ coef_b, int64_t coef_ld,
                   unsigned* __restrict__ done, int HW, int C, int G, int rows_per_block) {
  extern __shared__ __align__(16) float smem[];   // [acc 2C | k1 k2 k3 s1 s2][mbarrier][x tile][dy tile][dX tile]
  const int n = blockIdx.y, cpg = C / G;
  const int lanes = C / V;                        // <= kThreads (checked by the host)
  const int row_lanes = kThreads / lanes;
  const int lane = threadIdx.x % lanes, rl = threadIdx.x / lanes;
  const bool active = rl < row_lanes;
  const int c = lane * V;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int nrows = r1 - r0;
  const uint32_t row_bytes = (uint32_t)C * sizeof(T);
  uint8_t* base = reinterpret_cast<uint8_t*>(smem) + ((((size_t)(3 * C + 2 * G) * 4) + 15) & ~(size_t)15);
  uint64_t* bar = reinterpret_cast<uint64_t*>(base);
  uint8_t* tile_x = base + 16;
  uint8_t* tile_g = tile_x + (size_t)rows_per_block * row_bytes;
  uint8_t* tile_d = tile_g + (size_t)rows_per_block * row_bytes;
  const uint32_t bar_a = gn_smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) smem[i] = 0.f;
  __syncthreads();
  dlb_pdl_wait();                                 // nothing above reads what the previous kernel wrote
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)nrows * row_bytes * (ACC ? 3u : 2u)) : "memory");
    __syncwarp();
    for (int r = threadIdx.x; r < nrows; r += 32) {
      const int64_t row = (int64_t)n * HW + r0 + r;
      gn_bulk_load(gn_smem_u32(tile_x + (size_t)r * row_bytes), x + row * ldx, row_bytes, bar_a);
      gn_bulk_load(gn_smem_u32(tile_g + (size_t)r * row_bytes), dy + row * lddy, row_bytes, bar_a);
      if constexpr (ACC) gn_bulk_load(gn_smem_u32(tile_d + (size_t)r * row_bytes), dx + row * lddx, row_bytes, bar_a);
    }
  }
  float ka[V], kb[V];
#pragma unroll
  for (int i = 0; i < V; ++i) { ka[i] = coef_a[(int64_t)n * coef_ld + c + i]; kb[i] = coef_b[(int64_t)n * coef_ld + c + i]; }
  {                                               // wait for the tiles (watchdog instead of a hang)
    uint32_t ok = 0, spins = 0;
    while (true) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar_a), "r"(0u) : "memory");
      if (ok) break;
      if (++spins > (1u << 22)) __trap();
    }
  }
  // ---- phase 1: per-(n, c) sums of dz and dz*x over this block's rows ----
  if (active) {
    float a0[V], a1[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { a0[i] = 0.f; a1[i] = 0.f; }
    for (int r = rl; r < nrows; r += row_lanes) {
      float xv[V], gv[V];
      load_vec<T, V>(reinterpret_cast<const T*>(tile_x + (size_t)r * row_bytes) + c, xv);
      load_vec<T, V>(reinterpret_cast<const T*>(tile_g + (size_t)r * row_bytes) + c, gv);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float gz = fmaf(ka[i], xv[i], kb[i]) > 0.f ? gv[i] : 0.f;
        a0[i] += gz; a1[i] += gz * xv[i];
      }
    }
    bool owner = true;
    if (lanes < 32 && (lanes & (lanes - 1)) == 0) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        for (int o = lanes; o < 32; o <<= 1) {
          a0[i] += __shfl_xor_sync(0xffffffffu, a0[i], o);
          a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
        }
      }
      owner = (threadIdx.x & 31) < lanes;
    }
    if (owner) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        atomicAdd(&smem[2 * (c + i)], a0[i]);
        atomicAdd(&smem[2 * (c + i) + 1], a1[i]);
      }
    }
  }
  __syncthreads();
  float* tab = table + (int64_t)n * table_ns;
  for (int i = threadIdx.x; i < 2 * C; i += kThreads) atomicAdd(&tab[i], smem[i]);
  // ---- per-sample barrier (same protocol as gn_bwd_fused_kernel) ----
  __shared__ unsigned s_ticket;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_ticket = atomicAdd(&done[n], 1u);
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done + n) : "memory");
      if (v >= gridDim.x) break;
      if (++spins > (1u << 24)) __trap();
    }
  }
  __syncthreads();
  if (dgamma != nullptr && s_ticket == gridDim.x - 1) {
    for (int cc = threadIdx.x; cc < C; cc += kThreads) {
      const float A = __ldcg(tab + 2 * cc), B = __ldcg(tab + 2 * cc + 1);
      const int g = cc / cpg;
      atomicAdd(&dbeta[cc], A);
      atomicAdd(&dgamma[cc], rstd[n * G + g] * (B - mean[n * G + g] * A));
    }
  }
  // ---- phase 2: dX tile (+)= k1*dz + k2*x + k3, in shared memory, then bulk stores ----
  float* k1 = smem;
  float* k2 = smem + C;
  float* k3 = smem + 2 * C;
  float* s1 = smem + 3 * C;
  float* s2 = s1 + G;
  for (int g = threadIdx.x; g < G; g += kThreads) {
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    float a = 0.f, b = 0.f;
    for (int i = 0; i < cpg; ++i) {
      const int cc = g * cpg + i;
      const float A = __ldcg(tab + 2 * cc), B = __ldcg(tab + 2 * cc + 1);
      a += gamma[cc] * A;
      b += gamma[cc] * r * (B - mu * A);
    }
    s1[g] = a;
    s2[g] = b;
  }
  __syncthreads();
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  for (int cc = threadIdx.x; cc < C; cc += kThreads) {
    const int g = cc / cpg;
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    k1[cc] = gamma[cc] * r;
    const float q = r * r * s2[g] * inv_m;
    k2[cc] = -q;
    k3[cc] = -r * s1[g] * inv_m + q * mu;
  }
  __syncthreads();
  if (active) {
    float q1[V], q2[V], q3[V];
#pragma unroll
    for (int k = 0; k < V; ++k) { q1[k] = k1[c + k]; q2[k] = k2[c + k]; q3[k] = k3[c + k]; }
    for (int r = rl; r < nrows; r += row_lanes) {
      float xv[V], gv[V], out[V];
      load_vec<T, V>(reinterpret_cast<const T*>(tile_x + (size_t)r * row_bytes) + c, xv);
      load_vec<T, V>(reinterpret_cast<const T*>(tile_g + (size_t)r * row_bytes) + c, gv);
      T* drow = reinterpret_cast<T*>(tile_d + (size_t)r * row_bytes) + c;
      if constexpr (ACC) load_vec<T, V>(drow, out);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const float gz = fmaf(ka[k], xv[k], kb[k]) > 0.f ? gv[k] : 0.f;
        const float v = fmaf(q1[k], gz, fmaf(q2[k], xv[k], q3[k]));
        out[k] = ACC ? out[k] + v : v;
      }
      store_vec<T, V>(drow, out);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the bulk-copy engine
  __syncthreads();
  if (threadIdx.x < 32) {
    for (int r = threadIdx.x; r < nrows; r += 32)
      gn_bulk_store(dx + ((int64_t)n * HW + r0 + r) * lddx, gn_smem_u32(tile_d + (size_t)r * row_bytes), row_bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory must outlive the reads of the copy engine
  }
}

// dgamma[c] = sum_n rstd*(B - mu*A), dbeta[c] = sum_n A.   Block = 32 channels x 8 sample-lanes.
__global__ void __launch_bounds__(256) gn_param_grad_kernel(const float* __restrict__ table, int64_t table_ns, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int N, int C, int G) {
  __shared__ float sg[8][33], sb[8][33];
  const int cx = threadIdx.x & 31, ny = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float dg = 0.f, db = 0.f;
  if (c < C) {
    const int g = c / (C / G);
    for (int n = ny; n < N; n += 8) {
      const float2 t = *reinterpret_cast<const float2*>(table + (int64_t)n * table_ns + 2 * c);
      dg += rstd[n * G + g] * (t.y - mean[n * G + g] * t.x);
      db += t.x;
    }
  }
  sg[ny][cx] = dg; sb[ny][cx] = db;
  __syncthreads();
  if (ny == 0 && c < C) {
#pragma unroll
    for (int k = 1; k < 8; ++k) { dg += sg[k][cx]; db += sb[k][cx]; }
    dgamma[c] = dg;
    dbeta[c] = db;
  }
}

inline void grid_for(int N, int HW, int C, int V, dim3& grid, int& rows_per_block) {
  const int lanes = C / V;
  const int row_lanes = lanes >= kThreads ? 1 : kThreads / lanes;
  int max_chunks = HW / (row_lanes * 4);                  // keep >= 4 rows per row-lane per block
  if (max_chunks < 1) max_chunks = 1;
  // Every block pays a per-sample preamble (coefficients / zeroing / 2C atomics).  Large problems: give each block
  // >= ~192 KB of rows to stream (48/96/192 KB sweep on B200 at batch 512: 70.6/62.5/56.3 us), at most about one full wave of
  // resident blocks.  Small problems (per-rank batches of 32..128 after the DBS split): the 96 KB rule would leave one block
  // per sample -- 64 blocks on 148 SMs, each streaming its sample serially (measured 14-19 us for kernels that move < 10 MB)
  // -- so parallelism wins: aim for two blocks per SM as long as a block still gets >= 16 KB.
  const long long bytes_per_sample = (long long)HW * C * (V == 4 ? 4 : 2);
  if (!g_min_kb) { const char* e = getenv("DLB_GN_MIN_KB"); g_min_kb = e ? atoi(e) : 192; if (g_min_kb < 1) g_min_kb = 192; }
  const long long total = (long long)N * bytes_per_sample;
  long long want = total / ((long long)g_min_kb * 1024);
  long long floor_ctas = total / (16 * 1024);
  if (floor_ctas > 2 * 148) floor_ctas = 2 * 148;
  if (want < floor_ctas) want = floor_ctas;
  if (want > 148 * 8) want = 148 * 8;
  if (want < 1) want = 1;
  int chunks = (int)((want + N - 1) / N);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  rows_per_block = (HW + chunks - 1) / chunks;
  chunks = (HW + rows_per_block - 1) / rows_per_block;
  grid = dim3(chunks, N, 1);
}

template <typename T, int V>
int reduce2_launch(int mode, const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* y,
                   int64_t ldy, float* table, int64_t table_ns, int N, int HW, int C, cudaStream_t st,
                   const float* mean = nullptr, const float* rstd = nullptr, int G = 1, float* dgamma = nullptr,
                   float* dbeta = nullptr, const float* ca = nullptr, const float* cb = nullptr, int64_t cld = 0,
                   void* copy_dst = nullptr, int64_t ldc = 0) {
  dim3 grid; int rpb;
  grid_for(N, HW, C, V, grid, rpb);
  const size_t sm = 2 * C * sizeof(float);
  const T* X = (const T*)x; const T* DY = (const T*)dy; const T* Y = (const T*)y;
  unr_init();
#define RGO(MD, U) dlb_launch(nc_reduce2_kernel<T, V, MD, U>, grid, dim3(kThreads), sm, st, X, (int64_t)ldx, DY, (int64_t)lddy, Y, (int64_t)ldy, table, (int64_t)table_ns, HW, C, rpb, mean, rstd, G, dgamma, dbeta, ca, cb, (int64_t)cld, (T*)copy_dst, (int64_t)ldc)
#define RGU(MD) do { if (g_unr_red == 4) RGO(MD, 4); else if (g_unr_red == 2) RGO(MD, 2); else RGO(MD, 1); } while (0)
  if (mode == 0) RGU(0);
  else if (mode == 1) RGU(1);
  else if (mode == 3) RGU(3);
  else RGU(2);
#undef RGU
#undef RGO
  return dlb_post_launch();
}

template <typename T, int V>
int fwd_apply_launch(const void* x, int64_t ldx, const void* res, int64_t ldr, void* y, int64_t ldy,
                     const float* gamma, const float* beta, const float* mean, const float* rstd,
                     int N, int HW, int C, int G, int relu, cudaStream_t st, FwdFromTable ft = FwdFromTable{}) {
  dim3 grid; int rpb;
  grid_for(N, HW, C, V, grid, rpb);
  const size_t sm = (2 * C + 2 * G) * sizeof(float);
  const T* X = (const T*)x; const T* R = (const T*)res; T* Y = (T*)y;
  unr_init();
#define GOU(RL, RS, U) dlb_launch(gn_fwd_apply_kernel<T, V, RL, RS, U>, grid, dim3(kThreads), sm, st, X, (int64_t)ldx, R, (int64_t)ldr, Y, (int64_t)ldy, gamma, beta, mean, rstd, HW, C, G, rpb, ft)
#define GO(RL, RS) do { if (g_unr_fwd == 4) GOU(RL, RS, 4); else if (g_unr_fwd == 2) GOU(RL, RS, 2); else GOU(RL, RS, 1); } while (0)
  if (relu) { if (res) GO(true, true); else GO(true, false); }
  else { if (res) GO(false, true); else GO(false, false); }
#undef GO
#undef GOU
  return dlb_post_launch();
}

template <typename T, int V>
int bwd_apply_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* y, int64_t ldy,
                     void* dx, int64_t lddx, void* dres, int64_t lddr, const float* gamma,
                     const float* mean, const float* rstd, const float* table, int64_t table_ns, int N, int HW, int C,
                     int G, int relu, int acc, cudaStream_t st, const float* ca = nullptr, const float* cb = nullptr,
                     int64_t cld = 0, float* dgamma = nullptr, float* dbeta = nullptr) {
  dim3 grid; int rpb;
  grid_for(N, HW, C, V, grid, rpb);
  const size_t sm = (3 * C + 2 * G) * sizeof(float);
  const T* X = (const T*)x; const T* DY = (const T*)dy; const T* Y = (const T*)y;
  T* DX = (T*)dx; T* DR = (T*)dres;
  const int msrc = !relu ? 0 : (ca ? 2 : 1);
  unr_init();
#define GOU(RL, RS, AC, U) dlb_launch(gn_bwd_apply_kernel<T, V, RL, RS, AC, U>, grid, dim3(kThreads), sm, st, X, (int64_t)ldx, DY, (int64_t)lddy, Y, (int64_t)ldy, DX, (int64_t)lddx, DR, (int64_t)lddr, gamma, mean, rstd, table, (int64_t)table_ns, HW, C, G, rpb, ca, cb, (int64_t)cld, dgamma, dbeta)
#define GO(RL, RS, AC) do { if (g_unr_bwd == 4) GOU(RL, RS, AC, 4); else if (g_unr_bwd == 2) GOU(RL, RS, AC, 2); else GOU(RL, RS, AC, 1); } while (0)
#define GO2(RL) do { if (dres) { if (acc) GO(RL, true, true); else GO(RL, true, false); } else { if (acc) GO(RL, false, true); else GO(RL, false, false); } } while (0)
  if (msrc == 0) GO2(0); else if (msrc == 1) GO2(1); else GO2(2);
#undef GO2
#undef GO
#undef GOU
  return dlb_post_launch();
}

inline bool vec_ok(int dtype, int C, std::initializer_list<int64_t> lds, std::initializer_list<const void*> ptrs) {
  const int V = dtype == DLB_BF16 ? 8 : 4;
  if (C % V) return false;
  for (int64_t l : lds) if (l % V) return false;
  for (const void* p : ptrs) if (p && ((uintptr_t)p & 15)) return false;
  return true;
}

}  // namespace

#define DISPATCH(dtype, vec, CALL)                                                         \
  do {                                                                                      \
    if (dtype == DLB_BF16) { if (vec) { using T = __nv_bfloat16; constexpr int V = 8; CALL; } \
                             else { using T = __nv_bfloat16; constexpr int V = 1; CALL; } }  \
    else { if (vec) { using T = float; constexpr int V = 4; CALL; }                          \
           else { using T = float; constexpr int V = 1; CALL; } }                            \
  } while (0)

DLB_API void dlb_norm_skip_zero(int flag) { g_skip_zero = flag; }
// rows in flight per thread of the reduce / forward-apply / backward-apply kernels (1, 2 or 4; 0 keeps the current value)
DLB_API void dlb_norm_tune(int unr_red, int unr_fwd, int unr_bwd) {
  unr_init();
  if (unr_red == 1 || unr_red == 2 || unr_red == 4) g_unr_red = unr_red;
  if (unr_fwd == 1 || unr_fwd == 2 || unr_fwd == 4) g_unr_fwd = unr_fwd;
  if (unr_bwd == 1 || unr_bwd == 2 || unr_bwd == 4) g_unr_bwd = unr_bwd;
}
DLB_API void dlb_norm_tune_kb(int min_kb) { if (min_kb > 0) g_min_kb = min_kb; }
// fused GroupNorm backward: bulk-copy (TMA unit) flavour on/off (-1 keeps) and its tile budget in KB per block (0 keeps)
DLB_API void dlb_norm_bulk(int on, int tile_kb) {
  if (on == 0 || on == 1) g_bulk_on = on;
  if (tile_kb >= 8 && tile_kb <= 190) g_bulk_kb = tile_kb;
}

// table: fp32 [N][table_ns] with (q0,q1) pairs for C channels starting at `table`; zeroed here.
DLB_API int dlb_nc_reduce2(int mode, int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy,
                           const void* y, int64_t ldy, float* table, int64_t table_ns, int N, int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 6000) return -2;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  if (!g_skip_zero) cudaMemset2DAsync(table, (size_t)table_ns * sizeof(float), 0, (size_t)C * 2 * sizeof(float), (size_t)N, st);
  const bool vec = vec_ok(dtype, C, {ldx, dy ? lddy : 0, y ? ldy : 0}, {x, dy, y});
  int rc = 0;
  DISPATCH(dtype, vec, (rc = reduce2_launch<T, V>(mode, x, ldx, dy, lddy, y, ldy, table, table_ns, N, HW, C, st)));
  return rc;
}

// Backward reduction with the affine-parameter gradients folded in (dgamma/dbeta are zeroed here).
DLB_API int dlb_nc_reduce2_bwd(int relu, int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy,
                               const void* y, int64_t ldy, float* table, int64_t table_ns, const float* mean,
                               const float* rstd, float* dgamma, float* dbeta, int N, int HW, int C, int G, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 6000) return -2;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  if (!g_skip_zero) {
    cudaMemset2DAsync(table, (size_t)table_ns * sizeof(float), 0, (size_t)C * 2 * sizeof(float), (size_t)N, st);
    if (dgamma) { cudaMemsetAsync(dgamma, 0, C * sizeof(float), st); cudaMemsetAsync(dbeta, 0, C * sizeof(float), st); }
  }
  const bool vec = vec_ok(dtype, C, {ldx, lddy, relu ? ldy : 0}, {x, dy, relu ? y : nullptr});
  int rc = 0;
  DISPATCH(dtype, vec, (rc = reduce2_launch<T, V>(relu ? 1 : 2, x, ldx, dy, lddy, y, ldy, table, table_ns, N, HW, C, st, mean, rstd, G, dgamma, dbeta)));
  return rc;
}

// Same, with the ReLU mask recomputed from the forward coefficients (no saved activation needed).
DLB_API int dlb_nc_reduce2_bwd_coef(int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy, float* table,
                                    int64_t table_ns, const float* mean, const float* rstd, float* dgamma, float* dbeta,
                                    const float* ca, const float* cb, int64_t cld, int N, int HW, int C, int G, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 6000) return -2;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  if (!g_skip_zero) {
    cudaMemset2DAsync(table, (size_t)table_ns * sizeof(float), 0, (size_t)C * 2 * sizeof(float), (size_t)N, st);
    if (dgamma) { cudaMemsetAsync(dgamma, 0, C * sizeof(float), st); cudaMemsetAsync(dbeta, 0, C * sizeof(float), st); }
  }
  const bool vec = vec_ok(dtype, C, {ldx, lddy}, {x, dy});
  int rc = 0;
  DISPATCH(dtype, vec, (rc = reduce2_launch<T, V>(3, x, ldx, dy, lddy, nullptr, 0, table, table_ns, N, HW, C, st, mean, rstd, G, dgamma, dbeta, ca, cb, cld)));
  return rc;
}

DLB_API int dlb_gn_bwd_apply_coef(int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy, void* dx, int64_t lddx,
                                  const float* gamma, const float* mean, const float* rstd, const float* table, int64_t table_ns,
                                  const float* ca, const float* cb, int64_t cld, int N, int HW, int C, int G, int acc, void* stream) {
  int rc = 0;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  const bool vec = vec_ok(dtype, C, {ldx, lddy, lddx}, {x, dy, dx});
  DISPATCH(dtype, vec, (rc = bwd_apply_launch<T, V>(x, ldx, dy, lddy, nullptr, 0, dx, lddx, nullptr, 0, gamma, mean, rstd, table, table_ns, N, HW, C, G, 1, acc, (cudaStream_t)stream, ca, cb, cld)));
  return rc;
}


// GroupNorm(+ReLU) backward, reduce + apply in one launch (gn_bwd_fused_kernel).  `table` [N][table_ns], `dgamma`/`dbeta` [C] and
// `done` [N] (32-bit counters) must be zero on entry.  Returns 1 when the shape is not supported (caller falls back to
// dlb_nc_reduce2_bwd_coef + dlb_gn_bwd_apply_coef), 0 on success.
DLB_API int dlb_gn_bwd_fused(int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy, void* dx, int64_t lddx,
                             const float* gamma, const float* mean, const float* rstd, float* table, int64_t table_ns,
                             float* dgamma, float* dbeta, const float* ca, const float* cb, int64_t cld, void* done,
                             int N, int HW, int C, int G, int acc, void* stream) {
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  const int V = dtype == DLB_BF16 ? 8 : 4;
  if (!vec_ok(dtype, C, {ldx, lddy, lddx}, {x, dy, dx}) || C / V > kThreads || C > 6000 || (C % G)) return 1;
  dim3 grid; int rpb;
  grid_for(N, HW, C, V, grid, rpb);
  // staged flavour: once x + dy of the whole problem no longer fit comfortably in L2 (126 MB), keep each block's rows in
  // shared memory between the phases (<= 96 KB per block -> two blocks per SM)
  // (measured on B200, batch 512: the staged flavour LOSES -- tf32 32.6 vs 28.4 ms/step, bf16 23.3 vs 21.6: shared-memory
  // residency caps the blocks per SM and L2 already serves most of the phase-2 re-reads.  Opt-in: DLB_GN_BWD_STAGE=1.)
  static int stage_on = -1;
  if (stage_on < 0) { const char* e = getenv("DLB_GN_BWD_STAGE"); stage_on = (e && atoi(e) == 1) ? 1 : 0; }
  const size_t esz = dtype == DLB_BF16 ? 2 : 4;
  const size_t row_pair = 2 * (size_t)C * esz;
  const bool stage = stage_on && (size_t)N * HW * row_pair > (size_t)96 * 1024 * 1024 && row_pair * 8 <= 96 * 1024;
  if (stage) {
    int max_rows = (int)((96 * 1024) / row_pair);
    if (rpb > max_rows) rpb = max_rows;
    const int chunks = (HW + rpb - 1) / rpb;
    rpb = (HW + chunks - 1) / chunks;
    grid = dim3(chunks, N, 1);
  }
  cudaStream_t st = (cudaStream_t)stream;
  // bulk-copy flavour (opt-in, DLB_GN_BWD_BULK=1): tiles of <= DLB_GN_BULK_KB (default 64) KB per block
  if (g_bulk_on < 0) {
    // measured on B200 (profiles/r2_09_gn_bwd_flavours.txt): the bulk flavour LOSES to the register flavour at every DenseNet
    // shape (tf32 batch 512: 27-45 % vs 52-57 % of the copy peak; step 28.9 vs 26.9 ms) -- one bulk copy per 0.5-4 KB row is
    // too fine-grained for the copy engine and the small tiles multiply the per-block preamble and per-sample barrier
    // (ncu: 6 400 blocks, 14 waves, DRAM 22 %).  Opt-in: DLB_GN_BWD_BULK=1.
    const char* e = getenv("DLB_GN_BWD_BULK"); g_bulk_on = (e && atoi(e) == 1) ? 1 : 0;
    const char* k = getenv("DLB_GN_BULK_KB"); if (k && atoi(k) >= 8 && atoi(k) <= 190) g_bulk_kb = atoi(k);
  }
  const int bulk_on = g_bulk_on, bulk_kb = g_bulk_kb;
  if (bulk_on && !stage) {
    const size_t row_all = (size_t)C * esz * (acc ? 3 : 3);          // the dX tile exists in both flavours (written in place)
    int R = (int)(((size_t)bulk_kb * 1024) / row_all);
    if (R > HW) R = HW;
    if (R >= 1) {
      int chunks = (HW + R - 1) / R;
      // small problems: at least ~2 blocks per SM as long as a block keeps >= 4 rows
      const int want = (2 * 148 + N - 1) / N;
      if (chunks < want) chunks = want;
      if (chunks > HW / 4) chunks = HW / 4 > 0 ? HW / 4 : 1;
      R = (HW + chunks - 1) / chunks;
      chunks = (HW + R - 1) / R;
      const size_t coef_b = (((size_t)(3 * C + 2 * G) * sizeof(float)) + 15) & ~(size_t)15;
      const size_t smb = coef_b + 16 + 3 * (size_t)R * C * esz;
      if (chunks <= 64 && smb <= 196 * 1024) {
        dim3 g2(chunks, N, 1);
#define BGO(TT, VV, AC)                                                                                                          \
  do {                                                                                                                           \
    static bool cfgd = false;                                                                                                    \
    if (!cfgd) { cudaFuncSetAttribute(gn_bwd_bulk_kernel<TT, VV, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); cfgd = true; } \
    dlb_launch(gn_bwd_bulk_kernel<TT, VV, AC>, g2, dim3(kThreads), smb, st, (const TT*)x, (int64_t)ldx, (const TT*)dy, (int64_t)lddy, (TT*)dx, (int64_t)lddx, gamma, mean, rstd, table, (int64_t)table_ns, dgamma, dbeta, ca, cb, (int64_t)cld, (unsigned*)done, HW, C, G, R); \
  } while (0)
        if (dtype == DLB_BF16) { if (acc) BGO(__nv_bfloat16, 8, true); else BGO(__nv_bfloat16, 8, false); }
        else { if (acc) BGO(float, 4, true); else BGO(float, 4, false); }
#undef BGO
        return dlb_post_launch();
      }
    }
  }
  if (grid.x > 64) return 1;                 // keep a sample's blocks trivially co-resident
  const size_t coef_bytes = (((size_t)(3 * C + 2 * G) * sizeof(float)) + 15) & ~(size_t)15;
  const size_t sm = coef_bytes + (stage ? (size_t)rpb * row_pair : 0);
#define FGO(TT, VV, AC)                                                                                                          \
  do {                                                                                                                           \
    if (stage) {                                                                                                                 \
      static bool cfgd = false;                                                                                                  \
      if (!cfgd) { cudaFuncSetAttribute(gn_bwd_fused_kernel<TT, VV, AC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024); cfgd = true; } \
      dlb_launch(gn_bwd_fused_kernel<TT, VV, AC, true>, grid, dim3(kThreads), sm, st, (const TT*)x, (int64_t)ldx, (const TT*)dy, (int64_t)lddy, (TT*)dx, (int64_t)lddx, gamma, mean, rstd, table, (int64_t)table_ns, dgamma, dbeta, ca, cb, (int64_t)cld, (unsigned*)done, HW, C, G, rpb); \
    } else {                                                                                                                     \
      dlb_launch(gn_bwd_fused_kernel<TT, VV, AC, false>, grid, dim3(kThreads), sm, st, (const TT*)x, (int64_t)ldx, (const TT*)dy, (int64_t)lddy, (TT*)dx, (int64_t)lddx, gamma, mean, rstd, table, (int64_t)table_ns, dgamma, dbeta, ca, cb, (int64_t)cld, (unsigned*)done, HW, C, G, rpb); \
    }                                                                                                                            \
  } while (0)
  if (dtype == DLB_BF16) { if (acc) FGO(__nv_bfloat16, 8, true); else FGO(__nv_bfloat16, 8, false); }
  else { if (acc) FGO(float, 4, true); else FGO(float, 4, false); }
#undef FGO
  return dlb_post_launch();
}

// Per-(sample, channel) statistics of x fused with a strided copy of x into `dst` (moves a conv output into its
// channel slice of the dense-block buffer and produces its GroupNorm statistics in the same pass).
DLB_API int dlb_copy_stats(int dtype, const void* x, int64_t ldx, void* dst, int64_t ldd, float* table, int64_t table_ns,
                           int N, int HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (C > 6000) return -2;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  if (!g_skip_zero) cudaMemset2DAsync(table, (size_t)table_ns * sizeof(float), 0, (size_t)C * 2 * sizeof(float), (size_t)N, st);
  const bool vec = vec_ok(dtype, C, {ldx, ldd}, {x, dst});
  int rc = 0;
  DISPATCH(dtype, vec, (rc = reduce2_launch<T, V>(0, x, ldx, nullptr, 0, nullptr, 0, table, table_ns, N, HW, C, st, nullptr, nullptr, 1, nullptr, nullptr, nullptr, nullptr, 0, dst, ldd)));
  return rc;
}

DLB_API int dlb_gn_finalize(const float* table, int64_t table_ns, float* mean, float* rstd, int N, int C, int G, int HW,
                            float eps, void* stream) {
  const int total = N * G;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  dlb_launch(gn_finalize_kernel, dim3((total + 127) / 128), dim3(128), 0, (cudaStream_t)stream, table, (int64_t)table_ns, mean, rstd, N, C, G, HW, eps);
  return dlb_post_launch();
}

DLB_API int dlb_gn_coeff(const float* table, int64_t table_ns, const float* gamma, const float* beta, float* mean, float* rstd,
                         float* ca, float* cb, int64_t ld, int N, int C, int G, int HW, float eps, void* stream) {
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  dlb_launch(gn_coeff_kernel, dim3(N), dim3(256), 2 * G * sizeof(float), (cudaStream_t)stream, table, (int64_t)table_ns, gamma, beta, mean, rstd, ca, cb, (int64_t)ld, C, G, HW, eps);
  return dlb_post_launch();
}

DLB_API int dlb_gn_fwd_apply(int dtype, const void* x, int64_t ldx, const void* res, int64_t ldr, void* y, int64_t ldy,
                             const float* gamma, const float* beta, const float* mean, const float* rstd,
                             int N, int HW, int C, int G, int relu, void* stream) {
  int rc = 0;
  const bool vec = vec_ok(dtype, C, {ldx, res ? ldr : 0, ldy}, {x, res, y});
  DISPATCH(dtype, vec, (rc = fwd_apply_launch<T, V>(x, ldx, res, ldr, y, ldy, gamma, beta, mean, rstd, N, HW, C, G, relu, (cudaStream_t)stream)));
  return rc;
}

// GroupNorm(+ReLU) apply with the group statistics derived in-kernel from the (sum, sumsq) table; also writes mean/rstd [N*G]
// and the coefficient rows ca/cb [N][cld] (zero padded) for the backward pass.  Replaces dlb_gn_coeff + dlb_gn_fwd_apply.
DLB_API int dlb_gn_fwd_apply_table(int dtype, const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma, const float* beta,
                                   const float* table, int64_t table_ns, float* mean, float* rstd, float* ca, float* cb, int64_t cld,
                                   int N, int HW, int C, int G, float eps, int relu, void* stream) {
  int rc = 0;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  FwdFromTable ft;
  ft.tab = table; ft.tab_ns = table_ns; ft.eps = eps; ft.mean_out = mean; ft.rstd_out = rstd; ft.coef_a = ca; ft.coef_b = cb; ft.coef_ld = cld;
  const bool vec = vec_ok(dtype, C, {ldx, ldy}, {x, y});
  DISPATCH(dtype, vec, (rc = fwd_apply_launch<T, V>(x, ldx, nullptr, 0, y, ldy, gamma, beta, mean, rstd, N, HW, C, G, relu, (cudaStream_t)stream, ft)));
  return rc;
}

DLB_API int dlb_gn_bwd_apply(int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy, const void* y, int64_t ldy,
                             void* dx, int64_t lddx, void* dres, int64_t lddr, const float* gamma, const float* mean,
                             const float* rstd, const float* table, int64_t table_ns, int N, int HW, int C, int G,
                             int relu, int acc, void* stream) {
  int rc = 0;
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  const bool vec = vec_ok(dtype, C, {ldx, lddy, relu ? ldy : 0, lddx, dres ? lddr : 0}, {x, dy, relu ? y : nullptr, dx, dres});
  DISPATCH(dtype, vec, (rc = bwd_apply_launch<T, V>(x, ldx, dy, lddy, y, ldy, dx, lddx, dres, lddr, gamma, mean, rstd, table, table_ns, N, HW, C, G, relu, acc, (cudaStream_t)stream)));
  return rc;
}

DLB_API int dlb_gn_param_grad(const float* table, int64_t table_ns, const float* mean, const float* rstd, float* dgamma,
                              float* dbeta, int N, int C, int G, void* stream) {
  if (table_ns <= 0) table_ns = 2 * (int64_t)C;
  gn_param_grad_kernel<<<(C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(table, table_ns, mean, rstd, dgamma, dbeta, N, C, G);
  return dlb_post_launch();
}

// Forward: stats (reduce + finalize) + apply. `table` is scratch [N*C*2] fp32.
DLB_API int dlb_gn_forward(int dtype, const void* x, int64_t ldx, const void* res, int64_t ldr, void* y,
                           int64_t ldy, const float* gamma, const float* beta, float* mean, float* rstd,
                           float* table, int N, int HW, int C, int G, float eps, int relu,
                           int stats_ready, void* stream) {
  int rc = 0;
  if (!stats_ready) {
    rc = dlb_nc_reduce2(0, dtype, x, ldx, nullptr, 0, nullptr, 0, table, 0, N, HW, C, stream);
    if (rc) return rc;
    rc = dlb_gn_finalize(table, 0, mean, rstd, N, C, G, HW, eps, stream);
    if (rc) return rc;
  }
  return dlb_gn_fwd_apply(dtype, x, ldx, res, ldr, y, ldy, gamma, beta, mean, rstd, N, HW, C, G, relu, stream);
}

// Backward: reduce (dz, dz*x) + apply (+ param grads). dres may be null; acc!=0 accumulates into dx.
DLB_API int dlb_gn_backward(int dtype, const void* x, int64_t ldx, const void* dy, int64_t lddy,
                            const void* y, int64_t ldy, void* dx, int64_t lddx, void* dres, int64_t lddr,
                            const float* gamma, const float* mean, const float* rstd, float* table,
                            float* dgamma, float* dbeta, int N, int HW, int C, int G, int relu, int acc,
                            void* stream) {
  // the per-(n, c) sums first; the affine-parameter gradients are then added by ONE block per sample of the apply kernel
  // (they used to be added by every block of the reduce kernel: N x chunks atomics per channel)
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_skip_zero && dgamma) { cudaMemsetAsync(dgamma, 0, C * sizeof(float), st); cudaMemsetAsync(dbeta, 0, C * sizeof(float), st); }
  int rc = dlb_nc_reduce2_bwd(relu, dtype, x, ldx, dy, lddy, y, ldy, table, 0, mean, rstd, nullptr, nullptr, N, HW, C, G, stream);
  if (rc) return rc;
  int64_t table_ns = 2 * (int64_t)C;
  const bool vec = vec_ok(dtype, C, {ldx, lddy, relu ? ldy : 0, lddx, dres ? lddr : 0}, {x, dy, relu ? y : nullptr, dx, dres});
  DISPATCH(dtype, vec, (rc = bwd_apply_launch<T, V>(x, ldx, dy, lddy, y, ldy, dx, lddx, dres, lddr, gamma, mean, rstd, table, table_ns, N, HW, C, G, relu, acc, st, nullptr, nullptr, 0, dgamma, dbeta)));
  return rc;
}

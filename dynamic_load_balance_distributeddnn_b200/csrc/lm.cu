// Transformer-LM kernels: fused residual-add + LayerNorm (forward/backward) and fused causal self-attention
// for short sequences (the whole S <= 64 sequence of one (batch, head) lives in one CTA's shared memory).
//
// Reference: nn.TransformerEncoderLayer post-norm blocks (reference Net/Transformer.py:63-64 via torch.nn):
// `norm(x + dropout(sublayer(x)))` = 2 LayerNorms/layer, and MultiheadAttention with an additive causal float
// mask (Net/Transformer.py:71-74) + attention dropout 0.2, d_model 200, 2 heads => head_dim 100, S = 35
// (SURVEY K13/K14).  The reference path is bmm + softmax + dropout + bmm (+ mask add) = 5+ launches with a
// [B*H, S, S] round trip through HBM; here scores, softmax, dropout and the PV product never leave the SM.
#include "common.cuh"

namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned hash4(unsigned a, unsigned b, unsigned c, unsigned d) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ (d + 0x27D4EB2Fu) * 0x165667B1u;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

// ------------------------------------------------------------------------------------------------
// y = LayerNorm(x + r) * gamma + beta ; z = x + r saved for the backward.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ r, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, T* __restrict__ y, T* __restrict__ z,
                                                         float* __restrict__ mean, float* __restrict__ rstd, int rows, int d, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* xp = x + (int64_t)row * d;
  const T* rp = r + (int64_t)row * d;
  float s = 0.f, ss = 0.f;
  for (int i = lane; i < d; i += 32) {
    const float v = (float)xp[i] + (float)rp[i];
    s += v; ss += v * v;
  }
  s = warp_sum(s); ss = warp_sum(ss);
  const float mu = s / d;
  const float rs = rsqrtf(fmaxf(ss / d - mu * mu, 0.f) + eps);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  for (int i = lane; i < d; i += 32) {
    const float v = (float)xp[i] + (float)rp[i];
    z[(int64_t)row * d + i] = (T)v;
    y[(int64_t)row * d + i] = (T)((v - mu) * rs * gamma[i] + beta[i]);
  }
}

// dz = rstd * (g*dy - mean(g*dy) - xhat*mean(g*dy*xhat));  dgamma += dy*xhat, dbeta += dy (block-reduced atomics)
template <typename T>
__global__ void __launch_bounds__(256) add_ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ z, const float* __restrict__ mean,
                                                         const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                         T* __restrict__ dz, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                         int rows, int d) {
  extern __shared__ float sm[];            // dgamma[d], dbeta[d] partials of this block
  for (int i = threadIdx.x; i < 2 * d; i += 256) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int row = blockIdx.x * 8 + w; row < rows; row += gridDim.x * 8) {
    const float mu = mean[row], rs = rstd[row];
    const T* dyp = dy + (int64_t)row * d;
    const T* zp = z + (int64_t)row * d;
    float a = 0.f, b = 0.f;
    for (int i = lane; i < d; i += 32) {
      const float g = (float)dyp[i] * gamma[i];
      const float xh = ((float)zp[i] - mu) * rs;
      a += g; b += g * xh;
    }
    a = warp_sum(a) / d; b = warp_sum(b) / d;
    for (int i = lane; i < d; i += 32) {
      const float dyv = (float)dyp[i];
      const float xh = ((float)zp[i] - mu) * rs;
      dz[(int64_t)row * d + i] = (T)(rs * (dyv * gamma[i] - a - xh * b));
      atomicAdd(&sm[i], dyv * xh);
      atomicAdd(&sm[d + i], dyv);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d; i += 256) { atomicAdd(&dgamma[i], sm[i]); atomicAdd(&dbeta[i], sm[d + i]); }
}

// ------------------------------------------------------------------------------------------------
// Causal attention, one CTA per (batch, head).  qkv: [S, B, 3, H, hd] (the in-proj output, read in place),
// out: [S, B, H*hd].  probs (post-softmax, pre-dropout) are saved [B, H, S, S] fp32 for the backward.
struct AttnDims { int S, B, H, hd; float scale; float p_drop; };

template <typename T>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ probs,
                                                       AttnDims dm, unsigned seed, const long long* __restrict__ step_ptr) {
  extern __shared__ float sm[];
  const int S = dm.S, hd = dm.hd, H = dm.H, B = dm.B;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  float* q = sm;                       // [S][hd]
  float* k = q + S * hd;
  float* v = k + S * hd;
  float* p = v + S * hd;               // [S][S]
  const int64_t row_stride = (int64_t)B * 3 * H * hd;           // elements between consecutive s
  const T* base = qkv + (int64_t)b * 3 * H * hd + (int64_t)h * hd;
  for (int i = threadIdx.x; i < S * hd; i += 128) {
    const int s = i / hd, e = i % hd;
    const T* ptr = base + s * row_stride + e;
    q[i] = (float)ptr[0] * dm.scale;
    k[i] = (float)ptr[(int64_t)H * hd];
    v[i] = (float)ptr[(int64_t)2 * H * hd];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * S; i += 128) {
    const int r = i / S, c = i % S;
    float acc = -INFINITY;
    if (c <= r) {
      acc = 0.f;
      for (int e = 0; e < hd; ++e) acc = fmaf(q[r * hd + e], k[c * hd + e], acc);
    }
    p[i] = acc;
  }
  __syncthreads();
  const unsigned step = step_ptr ? (unsigned)(*step_ptr) : 0u;
  const float keep = 1.f - dm.p_drop, inv_keep = dm.p_drop > 0.f ? 1.f / keep : 1.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < S; r += 4) {
    float m = -INFINITY;
    for (int c = lane; c <= r; c += 32) m = fmaxf(m, p[r * S + c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int c = lane; c < S; c += 32) {
      const float e = c <= r ? __expf(p[r * S + c] - m) : 0.f;
      p[r * S + c] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int c = lane; c < S; c += 32) {
      const float pr = p[r * S + c] * inv;
      probs[(((int64_t)b * H + h) * S + r) * S + c] = pr;
      float pd = pr;
      if (dm.p_drop > 0.f) {
        const unsigned rnd = hash4(seed, step, (unsigned)(blockIdx.x), (unsigned)(r * S + c));
        pd = ((rnd >> 8) * (1.f / 16777216.f)) < keep ? pr * inv_keep : 0.f;
      }
      p[r * S + c] = pd;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * hd; i += 128) {
    const int r = i / hd, e = i % hd;
    float acc = 0.f;
    for (int c = 0; c <= r; ++c) acc = fmaf(p[r * S + c], v[c * hd + e], acc);
    out[((int64_t)r * B + b) * H * hd + h * hd + e] = (T)acc;
  }
}

// dqkv: [S, B, 3, H, hd] gradient, written in place of the three projections' slots.
template <typename T>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, const float* __restrict__ probs,
                                                       T* __restrict__ dqkv, AttnDims dm, unsigned seed, const long long* __restrict__ step_ptr) {
  extern __shared__ float sm[];
  const int S = dm.S, hd = dm.hd, H = dm.H, B = dm.B;
  const int b = blockIdx.x / H, h = blockIdx.x % H;
  float* q = sm;
  float* k = q + S * hd;
  float* v = k + S * hd;
  float* dO = v + S * hd;
  float* p = dO + S * hd;              // dropped probabilities Pd, later dS
  float* ps = p + S * S;               // softmax probabilities P
  const int64_t row_stride = (int64_t)B * 3 * H * hd;
  const T* base = qkv + (int64_t)b * 3 * H * hd + (int64_t)h * hd;
  for (int i = threadIdx.x; i < S * hd; i += 128) {
    const int s = i / hd, e = i % hd;
    const T* ptr = base + s * row_stride + e;
    q[i] = (float)ptr[0] * dm.scale;
    k[i] = (float)ptr[(int64_t)H * hd];
    v[i] = (float)ptr[(int64_t)2 * H * hd];
    dO[i] = (float)dout[((int64_t)s * B + b) * H * hd + h * hd + e];
  }
  const unsigned step = step_ptr ? (unsigned)(*step_ptr) : 0u;
  const float keep = 1.f - dm.p_drop, inv_keep = dm.p_drop > 0.f ? 1.f / keep : 1.f;
  for (int i = threadIdx.x; i < S * S; i += 128) {
    const float pr = probs[((int64_t)b * H + h) * S * S + i];
    float mask = 1.f;
    if (dm.p_drop > 0.f) {
      const unsigned rnd = hash4(seed, step, (unsigned)(blockIdx.x), (unsigned)i);
      mask = ((rnd >> 8) * (1.f / 16777216.f)) < keep ? inv_keep : 0.f;
    }
    ps[i] = pr;
    p[i] = pr * mask;                  // Pd
  }
  __syncthreads();
  T* dbase = dqkv + (int64_t)b * 3 * H * hd + (int64_t)h * hd;
  // dV[c,e] = sum_r Pd[r,c] dO[r,e]
  for (int i = threadIdx.x; i < S * hd; i += 128) {
    const int c = i / hd, e = i % hd;
    float acc = 0.f;
    for (int r = c; r < S; ++r) acc = fmaf(p[r * S + c], dO[r * hd + e], acc);
    dbase[c * row_stride + (int64_t)2 * H * hd + e] = (T)acc;
  }
  __syncthreads();
  // dPd[r,c] = dO[r,:].V[c,:] ; dP = dPd*mask ; dS = P*(dP - sum_c dP*P)
  for (int i = threadIdx.x; i < S * S; i += 128) {
    const int r = i / S, c = i % S;
    float acc = 0.f;
    if (c <= r) {
      for (int e = 0; e < hd; ++e) acc = fmaf(dO[r * hd + e], v[c * hd + e], acc);
      const float pr = ps[i];
      const float mask = pr > 0.f ? p[i] / pr : 0.f;       // recovers inv_keep or 0 (pr > 0 on the causal part)
      acc *= mask;
    }
    p[i] = acc;                        // dP
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < S; r += 4) {
    float dot = 0.f;
    for (int c = lane; c <= r; c += 32) dot = fmaf(p[r * S + c], ps[r * S + c], dot);
    dot = warp_sum(dot);
    for (int c = lane; c < S; c += 32) p[r * S + c] = c <= r ? ps[r * S + c] * (p[r * S + c] - dot) : 0.f;   // dS
  }
  __syncthreads();
  // dQ[r,e] = scale * sum_c dS[r,c] K[c,e] ; dK[c,e] = sum_r dS[r,c] Qs[r,e]   (q already carries the scale)
  for (int i = threadIdx.x; i < S * hd; i += 128) {
    const int r = i / hd, e = i % hd;
    float aq = 0.f, ak = 0.f;
    for (int c = 0; c <= r; ++c) aq = fmaf(p[r * S + c], k[c * hd + e], aq);
    for (int rr = r; rr < S; ++rr) ak = fmaf(p[rr * S + r], q[rr * hd + e], ak);
    dbase[r * row_stride + e] = (T)(aq * dm.scale);
    dbase[r * row_stride + (int64_t)H * hd + e] = (T)ak;
  }
}

}  // namespace

DLB_API int dlb_add_layer_norm_fwd(int dtype, const void* x, const void* r, const float* gamma, const float* beta, void* y, void* z,
                                   float* mean, float* rstd, int rows, int d, float eps, void* stream) {
  const int blocks = (rows + 7) / 8;
  if (blocks == 0) return 0;
  if (dtype == DLB_BF16)
    add_ln_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)r, gamma, beta, (__nv_bfloat16*)y, (__nv_bfloat16*)z, mean, rstd, rows, d, eps);
  else
    add_ln_fwd_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)r, gamma, beta, (float*)y, (float*)z, mean, rstd, rows, d, eps);
  return dlb_post_launch();
}

// dgamma / dbeta are zeroed here.
DLB_API int dlb_add_layer_norm_bwd(int dtype, const void* dy, const void* z, const float* mean, const float* rstd, const float* gamma,
                                   void* dz, float* dgamma, float* dbeta, int rows, int d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(dgamma, 0, d * sizeof(float), st);
  cudaMemsetAsync(dbeta, 0, d * sizeof(float), st);
  int blocks = (rows + 7) / 8;
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks == 0) return 0;
  const size_t smb = 2 * d * sizeof(float);
  if (dtype == DLB_BF16)
    add_ln_bwd_kernel<__nv_bfloat16><<<blocks, 256, smb, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)z, mean, rstd, gamma, (__nv_bfloat16*)dz, dgamma, dbeta, rows, d);
  else
    add_ln_bwd_kernel<float><<<blocks, 256, smb, st>>>((const float*)dy, (const float*)z, mean, rstd, gamma, (float*)dz, dgamma, dbeta, rows, d);
  return dlb_post_launch();
}

static size_t attn_smem(int S, int hd, bool bwd) { return (size_t)((bwd ? 4 : 3) * S * hd + (bwd ? 2 : 1) * S * S) * sizeof(float); }

DLB_API int dlb_attention_fwd(int dtype, const void* qkv, void* out, float* probs, int S, int B, int H, int hd, float scale, float p_drop,
                              unsigned seed, const long long* step_ptr, void* stream) {
  if (S > 64 || hd > 128) return -2;
  AttnDims dm{S, B, H, hd, scale, p_drop};
  const size_t smb = attn_smem(S, hd, false);
  if (dtype == DLB_BF16) {
    auto kern = attn_fwd_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    kern<<<B * H, 128, smb, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, probs, dm, seed, step_ptr);
  } else {
    auto kern = attn_fwd_kernel<float>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    kern<<<B * H, 128, smb, (cudaStream_t)stream>>>((const float*)qkv, (float*)out, probs, dm, seed, step_ptr);
  }
  return dlb_post_launch();
}

DLB_API int dlb_attention_bwd(int dtype, const void* qkv, const void* dout, const float* probs, void* dqkv, int S, int B, int H, int hd,
                              float scale, float p_drop, unsigned seed, const long long* step_ptr, void* stream) {
  if (S > 64 || hd > 128) return -2;
  AttnDims dm{S, B, H, hd, scale, p_drop};
  const size_t smb = attn_smem(S, hd, true);
  if (dtype == DLB_BF16) {
    auto kern = attn_bwd_kernel<__nv_bfloat16>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    kern<<<B * H, 128, smb, (cudaStream_t)stream>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)dout, probs, (__nv_bfloat16*)dqkv, dm, seed, step_ptr);
  } else {
    auto kern = attn_bwd_kernel<float>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    kern<<<B * H, 128, smb, (cudaStream_t)stream>>>((const float*)qkv, (const float*)dout, probs, (float*)dqkv, dm, seed, step_ptr);
  }
  return dlb_post_launch();
}

// ------------------------------------------------------------------------------------------------
// Vocabulary cross-entropy over materialised logits, forward AND backward in one pass, in place:
//   row r:  z = logits[r,:] + bias ;  loss_r = logsumexp(z) - z[target_r]
//           logits[r,:] <- (softmax(z) - onehot(target_r)) * scale          (= d loss / d z, scale = 1/T)
// One CTA per row; the row (V = 33 278 for wikitext-2) is staged in shared memory so global memory sees exactly
// one read and one write of the [T, V] matrix (reference: log_softmax + nll_loss + their backward kernels =
// ~5 passes over a 2.4 GB fp32 matrix; Net/Transformer.py:94-95, dbs.py:270-271).
namespace {
// VEC: the row stride `ld` is a multiple of the 16-byte vector width and rows are 16-byte aligned, so the row is
// streamed with 16-byte loads/stores, several in flight per thread.  Columns in [V, ld) are padding: they are
// excluded from the softmax and get a zero gradient.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) softmax_ce_inplace_kernel(T* __restrict__ logits, int64_t ld, const float* __restrict__ bias,
                                                                 const long long* __restrict__ target, float* __restrict__ loss_sum,
                                                                 int V, float scale) {
  extern __shared__ float row[];            // [V rounded up to the vector width] fp32
  __shared__ float red[8];
  __shared__ float bc[2];
  constexpr int W = 16 / sizeof(T);
  const int r = blockIdx.x;
  T* p = logits + (int64_t)r * ld;
  float m = -INFINITY;
  if constexpr (VEC) {
    const int nvec = (V + W - 1) / W;
    for (int v0 = threadIdx.x; v0 < nvec; v0 += 256 * 4) {
      float x[4][W];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * 256;
        if (v < nvec) load_vec<T, W>(p + (int64_t)v * W, x[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int v = v0 + u * 256;
        if (v < nvec) {
#pragma unroll
          for (int k = 0; k < W; ++k) {
            const int i = v * W + k;
            const float z = i < V ? x[u][k] + (bias ? bias[i] : 0.f) : -INFINITY;
            row[i] = z;
            m = fmaxf(m, z);
          }
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < V; i += 256) {
      const float z = (float)p[i] + (bias ? bias[i] : 0.f);
      row[i] = z;
      m = fmaxf(m, z);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {
    m = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0) bc[0] = m;
  }
  __syncthreads();
  m = bc[0];
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += 256) { const float e = __expf(row[i] - m); row[i] = e; s += e; }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) bc[1] = s;
  }
  __syncthreads();
  s = bc[1];
  const float inv = 1.f / s;
  const int t = (int)target[r];
  if (threadIdx.x == 0) atomicAdd(loss_sum, (__logf(s) - __logf(fmaxf(row[t], 1e-38f))) * scale);
  if constexpr (VEC) {
    const int nvec = (int)(ld / W);
    for (int v = threadIdx.x; v < nvec; v += 256) {
      float g[W];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        const int i = v * W + k;
        float gi = i < V ? row[i] * inv : 0.f;
        if (i == t) gi -= 1.f;
        g[k] = gi * scale;
      }
      store_vec<T, W>(p + (int64_t)v * W, g);
    }
  } else {
    for (int i = threadIdx.x; i < V; i += 256) {
      float g = row[i] * inv;
      if (i == t) g -= 1.f;
      p[i] = (T)(g * scale);
    }
  }
}
}  // namespace

DLB_API int dlb_softmax_ce_inplace(int dtype, void* logits, long long ld, const float* bias, const long long* target, float* loss_sum,
                                   int T_rows, int V, float scale, void* stream) {
  const size_t smb = (size_t)(V + 8) * sizeof(float);
  if (smb > 200 * 1024) return -2;
  const int W = dtype == DLB_BF16 ? 8 : 4;
  const bool vec = (ld % W == 0) && (((uintptr_t)logits & 15) == 0) && ld >= (long long)((V + W - 1) / W) * W;
#define GO(TT, VV)                                                                                         \
  do {                                                                                                     \
    auto kern = softmax_ce_inplace_kernel<TT, VV>;                                                         \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);                     \
    kern<<<T_rows, 256, smb, (cudaStream_t)stream>>>((TT*)logits, ld, bias, target, loss_sum, V, scale);   \
  } while (0)
  if (dtype == DLB_BF16) { if (vec) GO(__nv_bfloat16, true); else GO(__nv_bfloat16, false); }
  else { if (vec) GO(float, true); else GO(float, false); }
#undef GO
  return dlb_post_launch();
}

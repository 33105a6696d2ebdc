// Token embedding front-end of the language model in one kernel per direction (reference Net/Transformer.py:91-92 and the
// PositionalEncoding module :48-49):
//
//     out[s, b, :] = dropout( E[token[s, b], :] * sqrt(d) + pe[s, :] )
//
// instead of gather (nn.Embedding) -> scale -> add -> dropout = 4 launches and 3 intermediate [S, B, d] tensors.  The dropout
// mask is a counter-based hash of (seed, step, element), regenerated in the backward pass, which scatter-adds
//     dE[token[s, b], :] += sqrt(d) / (1 - p) * mask * dout[s, b, :]
// with fp32 reductions into the (zero-initialised) gradient table -- e.g. the parameter's slice of the flat gradient buffer.
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned hash3e(unsigned a, unsigned b, unsigned c) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}
__device__ __forceinline__ bool keep_elem(unsigned seed, unsigned step, unsigned idx, float p) {
  return (hash3e(seed, step, idx) >> 8) * (1.f / 16777216.f) >= p;
}

template <typename T>
__global__ void __launch_bounds__(256) embed_fwd_kernel(const long long* __restrict__ tok, const T* __restrict__ table, long long ldt,
                                                        const float* __restrict__ pe, long long ldpe, T* __restrict__ out, int S, int B,
                                                        int D, int V, float scale, float p, unsigned seed, const long long* __restrict__ step_ptr) {
  dlb_pdl_wait();
  const unsigned step = step_ptr ? (unsigned)(*step_ptr) : 0u;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (long long t = blockIdx.x; t < (long long)S * B; t += gridDim.x) {
    const int s = (int)(t / B);
    long long id = tok[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    const T* row = table + id * ldt;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float v = (float)row[c] * scale + pe[(long long)s * ldpe + c];
      if (p > 0.f) v = keep_elem(seed, step, (unsigned)(t * D + c), p) ? v * inv_keep : 0.f;
      out[t * D + c] = (T)v;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) embed_bwd_kernel(const long long* __restrict__ tok, const T* __restrict__ dout, float* __restrict__ dtable,
                                                        long long ldt, int S, int B, int D, int V, float scale, float p, unsigned seed,
                                                        const long long* __restrict__ step_ptr) {
  dlb_pdl_wait();
  const unsigned step = step_ptr ? (unsigned)(*step_ptr) : 0u;
  const float k = scale * (p > 0.f ? 1.f / (1.f - p) : 1.f);
  for (long long t = blockIdx.x; t < (long long)S * B; t += gridDim.x) {
    long long id = tok[t];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    float* row = dtable + id * ldt;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      if (p > 0.f && !keep_elem(seed, step, (unsigned)(t * D + c), p)) continue;
      atomicAdd(row + c, k * (float)dout[t * D + c]);
    }
  }
}

}  // namespace

DLB_API int dlb_embed_fwd(int dtype, const long long* tok, const void* table, long long ldt, const float* pe, long long ldpe, void* out,
                          int S, int B, int D, int V, float scale, float p, unsigned seed, const long long* step_ptr, void* stream) {
  const long long rows = (long long)S * B;
  if (rows <= 0) return 0;
  const int grid = (int)(rows < 148 * 16 ? rows : 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16)
    dlb_launch(embed_fwd_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, tok, (const __nv_bfloat16*)table, ldt, pe, ldpe,
               (__nv_bfloat16*)out, S, B, D, V, scale, p, seed, step_ptr);
  else
    dlb_launch(embed_fwd_kernel<float>, dim3(grid), dim3(256), 0, st, tok, (const float*)table, ldt, pe, ldpe, (float*)out, S, B, D, V,
               scale, p, seed, step_ptr);
  return dlb_post_launch();
}

// dtable: fp32 [V][ldt], pre-zeroed (or a gradient sink being accumulated into)
DLB_API int dlb_embed_bwd(int dtype, const long long* tok, const void* dout, float* dtable, long long ldt, int S, int B, int D, int V,
                          float scale, float p, unsigned seed, const long long* step_ptr, void* stream) {
  const long long rows = (long long)S * B;
  if (rows <= 0) return 0;
  const int grid = (int)(rows < 148 * 16 ? rows : 148 * 16);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16)
    dlb_launch(embed_bwd_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, st, tok, (const __nv_bfloat16*)dout, dtable, ldt, S, B, D, V, scale, p,
               seed, step_ptr);
  else
    dlb_launch(embed_bwd_kernel<float>, dim3(grid), dim3(256), 0, st, tok, (const float*)dout, dtable, ldt, S, B, D, V, scale, p, seed, step_ptr);
  return dlb_post_launch();
}

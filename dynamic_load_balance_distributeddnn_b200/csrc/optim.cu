// Flat-buffer optimizer path: multi-tensor gradient pack (+DBS weight, +local clip), global
// sum-of-squares, fused momentum-SGD with bf16 shadow refresh.
//
// Replaces, per step: the reference's 362 `weighted * param.grad` launches (dbs.py:295),
// torch.nn.utils.clip_grad_norm_ (dbs.py:274, K19) and the per-tensor torch.optim.SGD loop
// (dbs.py:238,369, K20) with three launches over flat buffers.
#include "common.cuh"

unsigned long long g_dlb_launches = 0;
int g_dlb_pdl = 0;
DLB_API void dlb_set_pdl(int on) { g_dlb_pdl = on; }
DLB_API int dlb_get_pdl() { return g_dlb_pdl; }

DLB_API unsigned long long dlb_launch_count() { return g_dlb_launches; }
DLB_API void dlb_launch_count_add(unsigned long long n) { g_dlb_launches += n; }

namespace {

constexpr int kMaxTensors = 96;
struct TensorList {
  const void* ptr[kMaxTensors];
  long long offset[kMaxTensors];     // destination offset (elements) in the flat buffer
  int numel[kMaxTensors];
  unsigned char dtype[kMaxTensors];
  int blk_start[kMaxTensors + 1];    // first block id of each tensor
  int count;
};
constexpr int kChunk = 256 * 8;      // elements per block

__device__ __forceinline__ float load_any(const void* p, int dtype, long long i) {
  return dtype == DLB_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}

__device__ __forceinline__ int find_tensor(const TensorList& tl, int blk) {
  int lo = 0, hi = tl.count - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tl.blk_start[mid] <= blk) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// sumsq += sum_i g_i^2 over all listed tensors
__global__ void __launch_bounds__(256) mt_sumsq_kernel(const __grid_constant__ TensorList tl, float* __restrict__ out) {
  const int t = find_tensor(tl, blockIdx.x);
  const int base = (blockIdx.x - tl.blk_start[t]) * kChunk;
  const int n = tl.numel[t];
  float acc = 0.f;
#pragma unroll 4
  for (int i = base + threadIdx.x; i < min(n, base + kChunk); i += 256) {
    const float v = load_any(tl.ptr[t], tl.dtype[t], i);
    acc += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ float warp_sums[8];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    acc = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffu, acc, o);
    if (threadIdx.x == 0) atomicAdd(out, acc);
  }
}

// flat[offset + i] = scale * g_i, scale = weights[rank] * clip_coef(sumsq)
template <typename TOut>
__global__ void __launch_bounds__(256) mt_pack_kernel(const __grid_constant__ TensorList tl, TOut* __restrict__ flat,
                                                      const float* __restrict__ weights, int rank,
                                                      const float* __restrict__ sumsq, float max_norm) {
  float scale = weights ? weights[rank] : 1.f;
  if (sumsq && max_norm > 0.f) {
    const float nrm = sqrtf(*sumsq);
    scale *= fminf(1.f, max_norm / (nrm + 1e-6f));       // torch.nn.utils.clip_grad_norm_ semantics
  }
  const int t = find_tensor(tl, blockIdx.x);
  const int base = (blockIdx.x - tl.blk_start[t]) * kChunk;
  const int n = tl.numel[t];
  TOut* dst = flat + tl.offset[t];
#pragma unroll 4
  for (int i = base + threadIdx.x; i < min(n, base + kChunk); i += 256)
    dst[i] = (TOut)(scale * load_any(tl.ptr[t], tl.dtype[t], i));
}

// v = mu*v + g (+wd*p);  p -= lr*v;  shadow = bf16(p)
__global__ void __launch_bounds__(256) sgd_flat_kernel(float* __restrict__ p, float* __restrict__ v,
                                                       const float* __restrict__ g, __nv_bfloat16* __restrict__ shadow,
                                                       long long n, const float* __restrict__ lr_ptr, float mu, float wd,
                                                       const float* __restrict__ sumsq, float max_norm, float* __restrict__ zero_buf) {
  const float lr = *lr_ptr;
  // post-reduce ("global") gradient clipping: the coefficient of clip_grad_norm_ applied to the REDUCED gradient,
  // folded into the update so no separate scale pass runs (--clip_mode global; the reference clips locally, dbs.py:274)
  const float gs = (sumsq && max_norm > 0.f) ? fminf(1.f, max_norm / (sqrtf(*sumsq) + 1e-6f)) : 1.f;
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 gv = reinterpret_cast<const float4*>(g)[i];
    gv.x *= gs; gv.y *= gs; gv.z *= gs; gv.w *= gs;
    vv.x = fmaf(mu, vv.x, fmaf(wd, pv.x, gv.x)); vv.y = fmaf(mu, vv.y, fmaf(wd, pv.y, gv.y));
    vv.z = fmaf(mu, vv.z, fmaf(wd, pv.z, gv.z)); vv.w = fmaf(mu, vv.w, fmaf(wd, pv.w, gv.w));
    pv.x = fmaf(-lr, vv.x, pv.x); pv.y = fmaf(-lr, vv.y, pv.y);
    pv.z = fmaf(-lr, vv.z, pv.z); pv.w = fmaf(-lr, vv.w, pv.w);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(v)[i] = vv;
    // the gradient accumulation buffer (gradient sinks write into it with red.add) is cleared for the next step here
    if (zero_buf) reinterpret_cast<float4*>(zero_buf)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 raw;
      raw.x = *reinterpret_cast<unsigned*>(&lo);
      raw.y = *reinterpret_cast<unsigned*>(&hi);
      reinterpret_cast<uint2*>(shadow)[i] = raw;
    }
  }
  // tail (n is padded to a multiple of 4 by the Python side, kept for safety)
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float vv = fmaf(mu, v[i], fmaf(wd, p[i], gs * g[i]));
    const float pv = fmaf(-lr, vv, p[i]);
    v[i] = vv; p[i] = pv;
    if (zero_buf) zero_buf[i] = 0.f;
    if (shadow) shadow[i] = __float2bfloat16(pv);
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16(src[i]);
}

int build_list(TensorList& tl, int count, const void* const* ptrs, const long long* offsets,
               const int* numels, const int* dtypes) {
  tl.count = count;
  int blk = 0;
  for (int i = 0; i < count; ++i) {
    tl.ptr[i] = ptrs[i];
    tl.offset[i] = offsets ? offsets[i] : 0;
    tl.numel[i] = numels[i];
    tl.dtype[i] = (unsigned char)dtypes[i];
    tl.blk_start[i] = blk;
    blk += (numels[i] + kChunk - 1) / kChunk;
  }
  tl.blk_start[count] = blk;
  return blk;
}

}  // namespace

// sumsq must be zeroed by the caller (dlb_zero_f32) before the first chunk of a step.
DLB_API int dlb_mt_sumsq(int count, const void* const* ptrs, const int* numels, const int* dtypes,
                         float* sumsq, void* stream) {
  int launched = 0;
  for (int s = 0; s < count; s += kMaxTensors) {
    TensorList tl;
    const int c = count - s < kMaxTensors ? count - s : kMaxTensors;
    const int blocks = build_list(tl, c, ptrs + s, nullptr, numels + s, dtypes + s);
    if (blocks == 0) continue;
    mt_sumsq_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tl, sumsq);
    ++launched;
  }
  return dlb_post_launch(launched);
}

DLB_API int dlb_mt_pack(int count, const void* const* ptrs, const long long* offsets, const int* numels,
                        const int* dtypes, void* flat, int flat_dtype, const float* weights, int rank,
                        const float* sumsq, float max_norm, void* stream) {
  int launched = 0;
  for (int s = 0; s < count; s += kMaxTensors) {
    TensorList tl;
    const int c = count - s < kMaxTensors ? count - s : kMaxTensors;
    const int blocks = build_list(tl, c, ptrs + s, offsets + s, numels + s, dtypes + s);
    if (blocks == 0) continue;
    if (flat_dtype == DLB_BF16)
      mt_pack_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(tl, (__nv_bfloat16*)flat, weights, rank, sumsq, max_norm);
    else
      mt_pack_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(tl, (float*)flat, weights, rank, sumsq, max_norm);
    ++launched;
  }
  return dlb_post_launch(launched);
}

DLB_API int dlb_sgd_flat_clip(float* p, float* v, const float* g, void* shadow, long long n, const float* lr_ptr,
                              float momentum, float weight_decay, const float* sumsq, float max_norm, float* zero_buf, void* stream) {
  if (n <= 0) return 0;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  sgd_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(p, v, g, (__nv_bfloat16*)shadow, n, lr_ptr, momentum, weight_decay,
                                                                  sumsq, max_norm, zero_buf);
  return dlb_post_launch();
}

DLB_API int dlb_sgd_flat(float* p, float* v, const float* g, void* shadow, long long n, const float* lr_ptr,
                         float momentum, float weight_decay, void* stream) {
  return dlb_sgd_flat_clip(p, v, g, shadow, n, lr_ptr, momentum, weight_decay, nullptr, 0.f, nullptr, stream);
}

DLB_API int dlb_cast_f32_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cast_f32_bf16_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
  return dlb_post_launch();
}

DLB_API int dlb_zero_f32(float* p, long long n, void* stream) {
  return (int)cudaMemsetAsync(p, 0, (size_t)n * sizeof(float), (cudaStream_t)stream);
}

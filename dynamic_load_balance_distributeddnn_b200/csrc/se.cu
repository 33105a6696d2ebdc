// Squeeze-and-excitation block (RegNetY; reference Net/RegNet.py:10-25) as one kernel per direction, NHWC:
//
//     s = mean_hw(x);  h = relu(W1 s + b1);  z = sigmoid(W2 h + b2);  out = x * z            (per sample)
//
// The reference runs adaptive_avg_pool2d + two 1x1 convolutions on 1x1 maps (cuDNN) + relu + sigmoid + a broadcast multiply
// (7 launches, each far below one wave).  Here one CTA owns one sample: the pooled vector, both tiny matrix-vector products
// and the gate live in shared memory; x is read twice (the second pass hits L2), out written once.
// Backward, same structure: dz = sum_hw(dout * x) -> through sigmoid / W2 / relu / W1 -> ds; dx = dout * z + ds / HW.
// The pre-activation gradients are written out per sample so the (tiny) weight gradients are two small matrix products.
#include "common.cuh"

namespace {

constexpr int kMaxC = 1024, kMaxS = 256;

template <typename T>
__global__ void __launch_bounds__(256) se_fwd_kernel(const T* __restrict__ x, long long ldx, T* __restrict__ out, long long ldo,
                                                     const T* __restrict__ w1, const T* __restrict__ b1, const T* __restrict__ w2,
                                                     const T* __restrict__ b2, float* __restrict__ s_out, float* __restrict__ h_out,
                                                     float* __restrict__ z_out, int HW, int C, int CS) {
  dlb_pdl_wait();
  __shared__ float s[kMaxC], z[kMaxC], h[kMaxS];
  const int n = blockIdx.x;
  const T* xn = x + (long long)n * HW * ldx;
  // 1. squeeze: thread -> channel (coalesced across the warp), strided over pixels by channel-group
  for (int c = threadIdx.x; c < C; c += blockDim.x) s[c] = 0.f;
  __syncthreads();
  {
    const int lanes = min(C, (int)blockDim.x);          // threads along channels
    const int rows = blockDim.x / lanes;                // pixel lanes
    const int c_l = threadIdx.x % lanes, r_l = threadIdx.x / lanes;
    if (r_l < rows) {
      for (int c = c_l; c < C; c += lanes) {
        float a = 0.f;
        for (int r = r_l; r < HW; r += rows) a += (float)xn[(long long)r * ldx + c];
        atomicAdd(&s[c], a);
      }
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s[c] *= inv; s_out[(long long)n * C + c] = s[c]; }
  __syncthreads();
  // 2. excitation: one warp per output row (dot product over the warp)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < CS; j += nw) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a += (float)w1[(long long)j * C + c] * s[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { a = fmaxf(a + (float)b1[j], 0.f); h[j] = a; h_out[(long long)n * CS + j] = a; }
  }
  __syncthreads();
  for (int c = warp; c < C; c += nw) {
    float a = 0.f;
    for (int j = lane; j < CS; j += 32) a += (float)w2[(long long)c * CS + j] * h[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { a = 1.f / (1.f + __expf(-(a + (float)b2[c]))); z[c] = a; z_out[(long long)n * C + c] = a; }
  }
  __syncthreads();
  // 3. scale
  T* on = out + (long long)n * HW * ldo;
  for (long long i = threadIdx.x; i < (long long)HW * C; i += blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    on[(long long)r * ldo + c] = (T)((float)xn[(long long)r * ldx + c] * z[c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) se_bwd_kernel(const T* __restrict__ x, long long ldx, const T* __restrict__ dout, long long ldg,
                                                     T* __restrict__ dx, long long lddx, const T* __restrict__ w1, const T* __restrict__ w2,
                                                     const float* __restrict__ h_in, const float* __restrict__ z_in,
                                                     float* __restrict__ dpre1, float* __restrict__ dpre2, int HW, int C, int CS) {
  dlb_pdl_wait();
  __shared__ float dz[kMaxC], z[kMaxC], ds[kMaxC], dh[kMaxS];
  const int n = blockIdx.x;
  const T* xn = x + (long long)n * HW * ldx;
  const T* gn = dout + (long long)n * HW * ldg;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { dz[c] = 0.f; z[c] = z_in[(long long)n * C + c]; }
  __syncthreads();
  {
    const int lanes = min(C, (int)blockDim.x);
    const int rows = blockDim.x / lanes;
    const int c_l = threadIdx.x % lanes, r_l = threadIdx.x / lanes;
    if (r_l < rows) {
      for (int c = c_l; c < C; c += lanes) {
        float a = 0.f;
        for (int r = r_l; r < HW; r += rows) a += (float)gn[(long long)r * ldg + c] * (float)xn[(long long)r * ldx + c];
        atomicAdd(&dz[c], a);
      }
    }
  }
  __syncthreads();
  // through the sigmoid
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = dz[c] * z[c] * (1.f - z[c]);
    dz[c] = g;
    dpre2[(long long)n * C + c] = g;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // dh[j] = sum_c W2[c][j] * dpre2[c], masked by the ReLU
  for (int j = warp; j < CS; j += nw) {
    float a = 0.f;
    for (int c = lane; c < C; c += 32) a += (float)w2[(long long)c * CS + j] * dz[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) { a = h_in[(long long)n * CS + j] > 0.f ? a : 0.f; dh[j] = a; dpre1[(long long)n * CS + j] = a; }
  }
  __syncthreads();
  // ds[c] = sum_j W1[j][c] * dpre1[j]
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < CS; ++j) a += (float)w1[(long long)j * C + c] * dh[j];
    ds[c] = a * inv;
  }
  __syncthreads();
  T* dn = dx + (long long)n * HW * lddx;
  for (long long i = threadIdx.x; i < (long long)HW * C; i += blockDim.x) {
    const int r = (int)(i / C), c = (int)(i % C);
    dn[(long long)r * lddx + c] = (T)((float)gn[(long long)r * ldg + c] * z[c] + ds[c]);
  }
}

}  // namespace

// x/out: NHWC [N, HW, C] with pixel strides ldx/ldo; w1 [CS][C], b1 [CS], w2 [C][CS], b2 [C] in the activation dtype;
// s_out [N][C], h_out [N][CS], z_out [N][C] fp32 (saved for the backward).
DLB_API int dlb_se_fwd(int dtype, const void* x, long long ldx, void* out, long long ldo, const void* w1, const void* b1, const void* w2,
                       const void* b2, float* s_out, float* h_out, float* z_out, int N, int HW, int C, int CS, void* stream) {
  if (C > kMaxC || CS > kMaxS || N <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16) {
    using T = __nv_bfloat16;
    dlb_launch(se_fwd_kernel<T>, dim3(N), dim3(256), 0, st, (const T*)x, ldx, (T*)out, ldo, (const T*)w1, (const T*)b1, (const T*)w2,
               (const T*)b2, s_out, h_out, z_out, HW, C, CS);
  } else {
    using T = float;
    dlb_launch(se_fwd_kernel<T>, dim3(N), dim3(256), 0, st, (const T*)x, ldx, (T*)out, ldo, (const T*)w1, (const T*)b1, (const T*)w2,
               (const T*)b2, s_out, h_out, z_out, HW, C, CS);
  }
  return dlb_post_launch();
}

// dpre1 [N][CS], dpre2 [N][C] fp32: gradients w.r.t. the two pre-activations (weight grads = dpre^T @ input, bias grads = column sums)
DLB_API int dlb_se_bwd(int dtype, const void* x, long long ldx, const void* dout, long long ldg, void* dx, long long lddx, const void* w1,
                       const void* w2, const float* h_in, const float* z_in, float* dpre1, float* dpre2, int N, int HW, int C, int CS,
                       void* stream) {
  if (C > kMaxC || CS > kMaxS || N <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16) {
    using T = __nv_bfloat16;
    dlb_launch(se_bwd_kernel<T>, dim3(N), dim3(256), 0, st, (const T*)x, ldx, (const T*)dout, ldg, (T*)dx, lddx, (const T*)w1, (const T*)w2,
               h_in, z_in, dpre1, dpre2, HW, C, CS);
  } else {
    using T = float;
    dlb_launch(se_bwd_kernel<T>, dim3(N), dim3(256), 0, st, (const T*)x, ldx, (const T*)dout, ldg, (T*)dx, lddx, (const T*)w1, (const T*)w2,
               h_in, z_in, dpre1, dpre2, HW, C, CS);
  }
  return dlb_post_launch();
}

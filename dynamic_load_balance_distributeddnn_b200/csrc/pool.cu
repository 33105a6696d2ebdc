// Non-overlapping k x k average pooling for NHWC activations (forward / backward).
// ATen's NHWC avg_pool2d kernels ran 6x (fwd) / 16x (bwd) off the bandwidth roofline on the DenseNet
// transition shapes (profiles/r1_01); these are plain 16-byte-vector streaming kernels.
// (reference Net/Densenet.py:32,81; Net/Resnet.py:85 — SURVEY K9)
#include "common.cuh"

namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W,
                                                          int C, int k) {
  dlb_pdl_wait();
  const int Ho = H / k, Wo = W / k, lanes = C / V;
  const int64_t total = (int64_t)N * Ho * Wo * lanes;
  const float inv = 1.f / (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % lanes) * V;
    int64_t r = i / lanes;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) {
        float v[V];
        load_vec<T, V>(x + (((int64_t)n * H + ho * k + dy) * W + wo * k + dx) * C + c, v);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += v[j];
      }
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] *= inv;
    store_vec<T, V>(y + (((int64_t)n * Ho + ho) * Wo + wo) * C + c, acc);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int N, int H, int W,
                                                          int C, int k) {
  dlb_pdl_wait();
  const int Ho = H / k, Wo = W / k, lanes = C / V;
  const int64_t total = (int64_t)N * H * W * lanes;
  const float inv = 1.f / (float)(k * k);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % lanes) * V;
    int64_t r = i / lanes;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float v[V];
    const int ho = h / k, wo = w / k;
    if (ho < Ho && wo < Wo) {
      load_vec<T, V>(dy + (((int64_t)n * Ho + ho) * Wo + wo) * C + c, v);
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] *= inv;
    } else {
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = 0.f;
    }
    store_vec<T, V>(dx + (((int64_t)n * H + h) * W + w) * C + c, v);
  }
}

template <typename T, int V>
int launch(bool fwd, const void* in, void* out, int N, int H, int W, int C, int k, cudaStream_t st) {
  const int64_t total = fwd ? (int64_t)N * (H / k) * (W / k) * (C / V) : (int64_t)N * H * W * (C / V);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) return 0;
  if (fwd) dlb_launch(avgpool_fwd_kernel<T, V>, dim3((int)blocks), dim3(256), 0, st, (const T*)in, (T*)out, N, H, W, C, k);
  else dlb_launch(avgpool_bwd_kernel<T, V>, dim3((int)blocks), dim3(256), 0, st, (const T*)in, (T*)out, N, H, W, C, k);
  return dlb_post_launch();
}

}  // namespace

// x: contiguous NHWC [N,H,W,C]; y: [N,H/k,W/k,C].  direction 0 = forward (in=x,out=y), 1 = backward (in=dy,out=dx).
DLB_API int dlb_avgpool_nhwc(int direction, int dtype, const void* in, void* out, int N, int H, int W, int C, int k, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool fwd = direction == 0;
  const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  if (dtype == DLB_BF16) {
    if (C % 8 == 0 && aligned) return launch<__nv_bfloat16, 8>(fwd, in, out, N, H, W, C, k, st);
    return launch<__nv_bfloat16, 1>(fwd, in, out, N, H, W, C, k, st);
  }
  if (C % 4 == 0 && aligned) return launch<float, 4>(fwd, in, out, N, H, W, C, k, st);
  return launch<float, 1>(fwd, in, out, N, H, W, C, k, st);
}

// ------------------------------------------------------------------------------------------------
// Strided 2-D copy: `rows` rows of `C` contiguous elements, source/destination row strides in elements.
// Moves channel slices in and out of the dense-block buffers (ATen's generic strided copy kernel ran at
// ~1 TB/s on these; profiles/r1_03).
namespace {
template <int VB>
__global__ void __launch_bounds__(256) copy2d_kernel(const unsigned char* __restrict__ src, int64_t lds_b,
                                                     unsigned char* __restrict__ dst, int64_t ldd_b, int64_t rows, int row_bytes) {
  dlb_pdl_wait();
  const int lanes = row_bytes / VB;
  const int64_t total = rows * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / lanes;
    const int c = (int)(i % lanes) * VB;
    if constexpr (VB == 16) *reinterpret_cast<uint4*>(dst + r * ldd_b + c) = *reinterpret_cast<const uint4*>(src + r * lds_b + c);
    else dst[r * ldd_b + c] = src[r * lds_b + c];
  }
}
}  // namespace

DLB_API int dlb_copy2d(const void* src, int64_t lds_bytes, void* dst, int64_t ldd_bytes, int64_t rows, int row_bytes, void* stream) {
  if (rows <= 0 || row_bytes <= 0) return 0;
  const bool v16 = (row_bytes % 16 == 0) && (lds_bytes % 16 == 0) && (ldd_bytes % 16 == 0) &&
                   ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0);
  const int64_t total = rows * (v16 ? row_bytes / 16 : row_bytes);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (v16) dlb_launch(copy2d_kernel<16>, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, (const unsigned char*)src, (int64_t)lds_bytes, (unsigned char*)dst, (int64_t)ldd_bytes, (int64_t)rows, row_bytes);
  else dlb_launch(copy2d_kernel<1>, dim3((int)blocks), dim3(256), 0, (cudaStream_t)stream, (const unsigned char*)src, (int64_t)lds_bytes, (unsigned char*)dst, (int64_t)ldd_bytes, (int64_t)rows, row_bytes);
  return dlb_post_launch();
}

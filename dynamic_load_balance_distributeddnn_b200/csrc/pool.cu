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


// ------------------------------------------------------------------------------------------------
// Max pooling, NHWC, k x k window, stride s, padding p (GoogLeNet's 3x3/s1/p1 and 3x3/s2/p1 pools, MnistNet's 2x2;
// reference Net/GoogleNet.py:42,68,79, Net/MnistNet.py:21-22).  The forward stores the winning tap of every
// output element (one byte); the backward GATHERS: an input element sums the dy of the <= ceil(k/s)^2 windows that chose it,
// so no atomics are needed even for overlapping windows.
namespace {

template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ idx,
                                                          int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p) {
  dlb_pdl_wait();
  const int lanes = C / V;
  const int64_t total = (int64_t)N * Ho * Wo * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % lanes) * V;
    int64_t r = i / lanes;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    float best[V]; unsigned char bi[V];
#pragma unroll
    for (int j = 0; j < V; ++j) { best[j] = -INFINITY; bi[j] = 0; }
    for (int dy = 0; dy < k; ++dy) {
      const int h = ho * s - p + dy;
      if (h < 0 || h >= H) continue;
      for (int dx = 0; dx < k; ++dx) {
        const int w = wo * s - p + dx;
        if (w < 0 || w >= W) continue;
        float v[V];
        load_vec<T, V>(x + (((int64_t)n * H + h) * W + w) * C + c, v);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (v[j] > best[j]) { best[j] = v[j]; bi[j] = (unsigned char)(dy * k + dx); }
      }
    }
    const int64_t o = (((int64_t)n * Ho + ho) * Wo + wo) * C + c;
    store_vec<T, V>(y + o, best);
#pragma unroll
    for (int j = 0; j < V; ++j) idx[o + j] = bi[j];
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx, T* __restrict__ dx,
                                                          int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p) {
  dlb_pdl_wait();
  const int lanes = C / V;
  const int64_t total = (int64_t)N * H * W * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % lanes) * V;
    int64_t r = i / lanes;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    float acc[V];
#pragma unroll
    for (int j = 0; j < V; ++j) acc[j] = 0.f;
    // windows (ho, wo) with ho*s - p <= h < ho*s - p + k
    const int ho0 = max(0, (h + p - k + s) / s), ho1 = min(Ho - 1, (h + p) / s);
    const int wo0 = max(0, (w + p - k + s) / s), wo1 = min(Wo - 1, (w + p) / s);
    for (int ho = ho0; ho <= ho1; ++ho)
      for (int wo = wo0; wo <= wo1; ++wo) {
        const int tap = (h - (ho * s - p)) * k + (w - (wo * s - p));
        const int64_t o = (((int64_t)n * Ho + ho) * Wo + wo) * C + c;
        float g[V];
        load_vec<T, V>(dy + o, g);
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] += (idx[o + j] == tap) ? g[j] : 0.f;
      }
    store_vec<T, V>(dx + (((int64_t)n * H + h) * W + w) * C + c, acc);
  }
}

template <typename T, int V>
int launch_max(bool fwd, const void* a, void* b, unsigned char* idx, int N, int H, int W, int C, int Ho, int Wo, int k, int s, int p,
               cudaStream_t st) {
  const int64_t total = fwd ? (int64_t)N * Ho * Wo * (C / V) : (int64_t)N * H * W * (C / V);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) return 0;
  if (fwd) dlb_launch(maxpool_fwd_kernel<T, V>, dim3((int)blocks), dim3(256), 0, st, (const T*)a, (T*)b, idx, N, H, W, C, Ho, Wo, k, s, p);
  else dlb_launch(maxpool_bwd_kernel<T, V>, dim3((int)blocks), dim3(256), 0, st, (const T*)a, (const unsigned char*)idx, (T*)b, N, H, W, C, Ho, Wo, k, s, p);
  return dlb_post_launch();
}

}  // namespace

// direction 0: in = x [N,H,W,C], out = y [N,Ho,Wo,C], idx written;  direction 1: in = dy [N,Ho,Wo,C], out = dx [N,H,W,C], idx read.
DLB_API int dlb_maxpool_nhwc(int direction, int dtype, const void* in, void* out, void* idx, int N, int H, int W, int C, int k, int s, int p,
                             void* stream) {
  if (k < 1 || k > 15 || s < 1 || p < 0 || 2 * p > k) return -2;
  const int Ho = (H + 2 * p - k) / s + 1, Wo = (W + 2 * p - k) / s + 1;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fwd = direction == 0;
  const bool aligned = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  if (dtype == DLB_BF16) {
    if (C % 8 == 0 && aligned) return launch_max<__nv_bfloat16, 8>(fwd, in, out, (unsigned char*)idx, N, H, W, C, Ho, Wo, k, s, p, st);
    return launch_max<__nv_bfloat16, 1>(fwd, in, out, (unsigned char*)idx, N, H, W, C, Ho, Wo, k, s, p, st);
  }
  if (C % 4 == 0 && aligned) return launch_max<float, 4>(fwd, in, out, (unsigned char*)idx, N, H, W, C, Ho, Wo, k, s, p, st);
  return launch_max<float, 1>(fwd, in, out, (unsigned char*)idx, N, H, W, C, Ho, Wo, k, s, p, st);
}

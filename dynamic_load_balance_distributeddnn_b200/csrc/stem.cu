// Stem convolution: 3x3 / stride 1 / pad 1 with a handful of input channels (the RGB stem `conv1` of every CNN in the zoo:
// reference Net/Densenet.py:42,76  Net/Resnet.py:63  Net/RegNet.py:70  Net/GoogleNet.py:59), NHWC, bf16 or fp32.
//
// With Cin = 3 the implicit-GEMM K dimension is 27: far below a tensor-core tile and not addressable by TMA (6-byte pixels),
// so this is a direct SIMT convolution -- it is purely bandwidth-bound on its Co-channel output (forward) / dY (weight
// gradient), the input and the weights live in shared memory.  No data gradient: the network input needs none.
#include "common.cuh"

namespace {

constexpr int kMaxK = 36;            // 9 taps x Cin <= 4 (a multiple of 4: float4 rows)
constexpr int kMaxCo = 128;

template <typename T> __device__ __forceinline__ float to_f(T v) { return (float)v; }

// forward: thread = (pixel, group of 8 output channels); the 9*Cin inputs of the pixel are read straight from global memory
// (the RGB input is a few MB: L1/L2 resident, the 8 threads of a pixel broadcast), weights from shared memory; the kernel is
// bound by its Co-channel output.
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, long long ldy,
                                                       int N, int H, int W, int Cin, int Co) {
  dlb_pdl_wait();
  __shared__ float sw[kMaxK][kMaxCo];          // weights, [tap*Cin + ci][co]
  const int K = 9 * Cin;
  const int groups = Co / 8;
  for (int i = threadIdx.x; i < K * Co; i += 256) {
    const int co = i / K, k = i - co * K;       // memory order [Co][3][3][Cin]
    sw[k][co] = to_f(w[i]);
  }
  __syncthreads();
  const long long total = (long long)N * H * W;
  const long long items = total * groups;
  for (long long it = (long long)blockIdx.x * 256 + threadIdx.x; it < items; it += (long long)gridDim.x * 256) {
    const long long pix = it / groups;
    const int g = (int)(it - pix * groups);
    const int wq = (int)(pix % W);
    const long long t2 = pix / W;
    const int hq = (int)(t2 % H);
    const long long n = t2 / H;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
      if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const T* px = x + ((n * H + hh) * W + ww) * Cin;
        for (int ci = 0; ci < Cin; ++ci) {
          const float xv = to_f(px[ci]);
          const float* wr = &sw[tap * Cin + ci][g * 8];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, wr[j], acc[j]);
        }
      }
    }
    T* dst = y + pix * ldy + g * 8;
    if constexpr (sizeof(T) == 2) {
      store_vec<T, 8>(dst, acc);
    } else {
      float lo[4] = {acc[0], acc[1], acc[2], acc[3]}, hi[4] = {acc[4], acc[5], acc[6], acc[7]};
      store_vec<T, 4>(dst, lo);
      store_vec<T, 4>(dst + 4, hi);
    }
  }
}

// weight gradient: dW[co][k] += sum_pixels dY[p, co] * patch[p, k].  Per tile of 64 pixels the block stages dY [64][Co] and the
// patches [64][K padded to 28/36] in shared memory; thread (tc, tk) accumulates a 4 x 4 register tile of (co, k) outputs from
// two 128-bit shared loads per pixel (16 FMAs per 2 loads).
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, long long lddy,
                                                         float* __restrict__ dw, int N, int H, int W, int Cin, int Co) {
  dlb_pdl_wait();
  constexpr int TP = 64;
  __shared__ __align__(16) float sdy[TP][kMaxCo];
  __shared__ __align__(16) float sx[TP][kMaxK];
  const int K = 9 * Cin;
  const int kq = (K + 3) / 4;                  // k quads (7 for Cin = 3)
  const int cq = Co / 4;                       // co quads
  const int tk = threadIdx.x % kq, tc = threadIdx.x / kq;
  const bool active = tc < cq;                 // 16 x 7 = 112 threads for Co = 64, Cin = 3; 32 x 7 = 224 for Co = 128
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const long long total = (long long)N * H * W;
  for (long long p0 = (long long)blockIdx.x * TP; p0 < total; p0 += (long long)gridDim.x * TP) {
    __syncthreads();
    for (int i = threadIdx.x; i < TP * Co; i += 256) {
      const int pp = i / Co, co = i - pp * Co;
      const long long pix = p0 + pp;
      sdy[pp][co] = pix < total ? to_f(dy[pix * lddy + co]) : 0.f;
    }
    for (int i = threadIdx.x; i < TP * kMaxK; i += 256) {
      const int pp = i / kMaxK, k = i - pp * kMaxK;
      const long long pix = p0 + pp;
      float v = 0.f;
      if (pix < total && k < K) {
        const int tap = k / Cin, ci = k - tap * Cin;
        const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f(x[((n * H + hh) * W + ww) * Cin + ci]);
      }
      sx[pp][k] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int pp = 0; pp < TP; ++pp) {
        const float4 d = *reinterpret_cast<const float4*>(&sdy[pp][tc * 4]);
        const float4 v = *reinterpret_cast<const float4*>(&sx[pp][tk * 4]);
        const float dd[4] = {d.x, d.y, d.z, d.w}, vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(dd[a], vv[b], acc[a][b]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int co = tc * 4 + a, k = tk * 4 + b;
        if (k < K) atomicAdd(dw + (long long)co * K + k, acc[a][b]);          // [Co][3][3][Cin] == [co][k]
      }
  }
}

}  // namespace

// y[N,H,W,Co] (pixel stride ldy) = conv3x3(x[N,H,W,Cin] dense, w[Co][3][3][Cin]); Cin <= 4, Co % 8 == 0, Co <= 128.
DLB_API int dlb_stem_conv_fwd(int dtype, const void* x, const void* w, void* y, long long ldy, int N, int H, int W, int Cin, int Co,
                              void* stream) {
  if (Cin < 1 || Cin > 4 || (Co % 8) || Co > kMaxCo || Co < 32 || (256 % (Co / 8))) return -2;      // P = 256 / (Co / 8) <= 64
  if (((uintptr_t)y & 15) || (ldy % (dtype == DLB_BF16 ? 8 : 4))) return -3;
  const long long total = (long long)N * H * W;
  long long blocks = (total * (Co / 8) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16)
    dlb_launch(stem_fwd_kernel<__nv_bfloat16>, dim3((int)blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
               (__nv_bfloat16*)y, ldy, N, H, W, Cin, Co);
  else
    dlb_launch(stem_fwd_kernel<float>, dim3((int)blocks), dim3(256), 0, st, (const float*)x, (const float*)w, (float*)y, ldy, N, H, W, Cin, Co);
  return dlb_post_launch();
}

// dw fp32 [Co][3][3][Cin], zero-initialised by the caller (or a gradient sink being accumulated into)
DLB_API int dlb_stem_conv_wgrad(int dtype, const void* x, const void* dy, long long lddy, float* dw, int N, int H, int W, int Cin, int Co,
                                void* stream) {
  if (Cin < 1 || Cin > 4 || Co > kMaxCo || Co < 4 || (Co % 4) || (Co / 4) * ((9 * Cin + 3) / 4) > 256) return -2;
  const long long total = (long long)N * H * W;
  long long blocks = (total + 63) / 64;
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16)
    dlb_launch(stem_wgrad_kernel<__nv_bfloat16>, dim3((int)blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, lddy, dw,
               N, H, W, Cin, Co);
  else
    dlb_launch(stem_wgrad_kernel<float>, dim3((int)blocks), dim3(256), 0, st, (const float*)x, (const float*)dy, lddy, dw, N, H, W, Cin, Co);
  return dlb_post_launch();
}

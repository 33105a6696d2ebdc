// Stem convolution: 3x3 / stride 1 / pad 1 with a handful of input channels (the RGB stem `conv1` of every CNN in the zoo:
// reference Net/Densenet.py:42,76  Net/Resnet.py:63  Net/RegNet.py:70  Net/GoogleNet.py:59), NHWC, bf16 or fp32.
//
// With Cin = 3 the implicit-GEMM K dimension is 27: far below a tensor-core tile and not addressable by TMA (6-byte pixels),
// so this is a direct SIMT convolution -- it is purely bandwidth-bound on its Co-channel output (forward) / dY (weight
// gradient), the input and the weights live in shared memory.  No data gradient: the network input needs none.
#include "common.cuh"

namespace {

constexpr int kMaxK = 36;            // 9 taps x Cin <= 4
constexpr int kMaxCo = 128;

template <typename T> __device__ __forceinline__ float to_f(T v) { return (float)v; }

// forward: block = 256 threads = P pixels x (Co/8) channel groups; every thread produces 8 output channels of one pixel
template <typename T>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, T* __restrict__ y, long long ldy,
                                                       int N, int H, int W, int Cin, int Co) {
  dlb_pdl_wait();
  __shared__ float sw[kMaxK][kMaxCo];          // weights, [tap*Cin + ci][co]
  __shared__ float sx[64][kMaxK + 1];          // input patches of this block's P <= 64 pixels
  const int K = 9 * Cin;
  const int groups = Co / 8;
  const int P = 256 / groups;
  for (int i = threadIdx.x; i < K * Co; i += 256) {
    const int co = i / K, k = i % K;            // memory order [Co][3][3][Cin]
    sw[k][co] = to_f(w[i]);
  }
  const long long total = (long long)N * H * W;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  for (long long p0 = (long long)blockIdx.x * P; p0 < total; p0 += (long long)gridDim.x * P) {
    __syncthreads();
    for (int i = threadIdx.x; i < P * K; i += 256) {
      const int pp = i / K, k = i % K;
      const long long pix = p0 + pp;
      float v = 0.f;
      if (pix < total) {
        const int tap = k / Cin, ci = k % Cin;
        const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f(x[((n * H + hh) * W + ww) * Cin + ci]);
      }
      sx[pp][k] = v;
    }
    __syncthreads();
    const long long pix = p0 + pl;
    if (pl < P && pix < total) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int k = 0; k < K; ++k) {
        const float xv = sx[pl][k];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, sw[k][g * 8 + j], acc[j]);
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
}

// weight gradient: dW[co][k] += sum_pixels dY[p, co] * patch[p, k].  Block = 256 threads; per tile of 64 pixels it stages
// dY [64][Co] and the patches [64][K] in shared memory; thread t owns the outputs (co, k) with (co*K + k) % 256 == t.
template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, long long lddy,
                                                         float* __restrict__ dw, int N, int H, int W, int Cin, int Co) {
  dlb_pdl_wait();
  constexpr int TP = 64;
  __shared__ float sdy[TP][kMaxCo + 1];
  __shared__ float sx[TP][kMaxK + 1];
  const int K = 9 * Cin;
  const int outs = Co * K;
  constexpr int kPer = (kMaxCo * kMaxK + 255) / 256;      // 18 outputs per thread at most
  float acc[kPer];
#pragma unroll
  for (int i = 0; i < kPer; ++i) acc[i] = 0.f;
  const long long total = (long long)N * H * W;
  for (long long p0 = (long long)blockIdx.x * TP; p0 < total; p0 += (long long)gridDim.x * TP) {
    __syncthreads();
    for (int i = threadIdx.x; i < TP * Co; i += 256) {
      const int pp = i / Co, co = i % Co;
      const long long pix = p0 + pp;
      sdy[pp][co] = pix < total ? to_f(dy[pix * lddy + co]) : 0.f;
    }
    for (int i = threadIdx.x; i < TP * K; i += 256) {
      const int pp = i / K, k = i % K;
      const long long pix = p0 + pp;
      float v = 0.f;
      if (pix < total) {
        const int tap = k / Cin, ci = k % Cin;
        const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
        const long long n = pix / ((long long)W * H);
        const int hh = hq + tap / 3 - 1, ww = wq + tap % 3 - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = to_f(x[((n * H + hh) * W + ww) * Cin + ci]);
      }
      sx[pp][k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int o = threadIdx.x + 256 * i;
      if (o < outs) {
        const int co = o / K, k = o % K;
        float a = acc[i];
#pragma unroll 8
        for (int pp = 0; pp < TP; ++pp) a = fmaf(sdy[pp][co], sx[pp][k], a);
        acc[i] = a;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kPer; ++i) {
    const int o = threadIdx.x + 256 * i;
    if (o < outs) atomicAdd(dw + o, acc[i]);               // [Co][3][3][Cin] == [co][k]
  }
}

}  // namespace

// y[N,H,W,Co] (pixel stride ldy) = conv3x3(x[N,H,W,Cin] dense, w[Co][3][3][Cin]); Cin <= 4, Co % 8 == 0, Co <= 128.
DLB_API int dlb_stem_conv_fwd(int dtype, const void* x, const void* w, void* y, long long ldy, int N, int H, int W, int Cin, int Co,
                              void* stream) {
  if (Cin < 1 || Cin > 4 || (Co % 8) || Co > kMaxCo || Co < 32 || (256 % (Co / 8))) return -2;      // P = 256 / (Co / 8) <= 64
  if (((uintptr_t)y & 15) || (ldy % (dtype == DLB_BF16 ? 8 : 4))) return -3;
  const long long total = (long long)N * H * W;
  const int P = 256 / (Co / 8);
  long long blocks = (total + P - 1) / P;
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
  if (Cin < 1 || Cin > 4 || Co > kMaxCo || Co < 1) return -2;
  const long long total = (long long)N * H * W;
  long long blocks = (total + 63) / 64;
  if (blocks > 148 * 2) blocks = 148 * 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DLB_BF16)
    dlb_launch(stem_wgrad_kernel<__nv_bfloat16>, dim3((int)blocks), dim3(256), 0, st, (const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, lddy, dw,
               N, H, W, Cin, Co);
  else
    dlb_launch(stem_wgrad_kernel<float>, dim3((int)blocks), dim3(256), 0, st, (const float*)x, (const float*)dy, lddy, dw, N, H, W, Cin, Co);
  return dlb_post_launch();
}

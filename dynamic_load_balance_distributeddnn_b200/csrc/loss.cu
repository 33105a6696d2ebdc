// Classifier-head loss: softmax + NLL (mean over the batch) forward and d(loss)/d(logits) in ONE launch.
//
// Replaces F.cross_entropy's log_softmax / nll_loss forward + two backward kernels on the CNN path (reference dbs.py:374;
// SURVEY K17).  The gradient seed is additionally multiplied by a device-resident scalar -- the rank's DBS weight
// w_r = local_bs / global_bs (reference dbs.py:293-295 applies it to every gradient tensor after backward): every gradient
// the backward pass produces is then already weighted, so the gradient collective needs no scale pass at all.
#include "common.cuh"

namespace {

// one warp per row; C <= 32 * kMaxPerLane
constexpr int kMaxPerLane = 32;

template <typename T>
__global__ void __launch_bounds__(256) softmax_ce_small_kernel(const T* __restrict__ logits, long long ld, const long long* __restrict__ target,
                                                               float* __restrict__ dlogits, float* __restrict__ loss_out,
                                                               const float* __restrict__ grad_scale, int B, int C) {
  dlb_pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  float row_loss = 0.f;
  if (warp < B) {
    const T* x = logits + (long long)warp * ld;
    float v[kMaxPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < C ? (float)x[c] : -INFINITY;
      mx = fmaxf(mx, v[i]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) { v[i] = __expf(v[i] - mx); s += v[i]; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int t = (int)target[warp];
    const float inv = 1.f / s;
    const float gs = (grad_scale ? *grad_scale : 1.f) / (float)B;
    float xt = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxPerLane; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        const float p = v[i] * inv;
        if (c == t) xt = p;
        dlogits[(long long)warp * C + c] = (p - (c == t ? 1.f : 0.f)) * gs;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) xt += __shfl_xor_sync(0xffffffffu, xt, o);
    row_loss = -__logf(fmaxf(xt, 1e-37f));
  }
  // block reduction of the per-row losses -> one atomic per block
  __shared__ float part[8];
  if (lane == 0) part[threadIdx.x >> 5] = row_loss;
  __syncthreads();
  if (threadIdx.x < 8) {
    float a = part[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) a += __shfl_xor_sync(0xffu, a, o);
    if (threadIdx.x == 0) atomicAdd(loss_out, a / (float)B);
  }
}

}  // namespace

// logits [B, C] (row stride ld, fp32 or bf16), target int64 [B]; dlogits fp32 [B, C] = (softmax - onehot) * grad_scale / B;
// loss_out (fp32 scalar, zeroed here) = mean NLL.
DLB_API int dlb_softmax_ce_small(int dtype, const void* logits, long long ld, const long long* target, float* dlogits, float* loss_out,
                                 const float* grad_scale, int B, int C, void* stream) {
  if (C > 32 * kMaxPerLane || B <= 0) return -2;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(loss_out, 0, sizeof(float), st);
  const int blocks = (B + 7) / 8;
  if (dtype == DLB_BF16)
    dlb_launch(softmax_ce_small_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, st, (const __nv_bfloat16*)logits, ld, target, dlogits, loss_out, grad_scale, B, C);
  else
    dlb_launch(softmax_ce_small_kernel<float>, dim3(blocks), dim3(256), 0, st, (const float*)logits, ld, target, dlogits, loss_out, grad_scale, B, C);
  return dlb_post_launch();
}

// Common device/host helpers for the dlb_b200 native library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DLB_API extern "C" __attribute__((visibility("default")))

// dtype codes shared with Python (ops/_native.py)
enum DlbDtype : int { DLB_F32 = 0, DLB_BF16 = 1 };

extern unsigned long long g_dlb_launches;   // host-side count of kernels launched by this library
extern int g_dlb_pdl;                        // programmatic dependent launch on/off (optim.cu; DLB_PDL env / dlb_set_pdl)

// Programmatic dependent launch: with the attribute set, a kernel may be scheduled while its predecessor in the
// stream is still draining; it must not touch the predecessor's outputs before `dlb_pdl_wait()`.  Every kernel of
// this library calls dlb_pdl_wait() first thing, so only the launch latency / block scheduling is overlapped --
// which is what bounds the step at small per-rank batches (~1 700 graph nodes of a few microseconds each).
__device__ __forceinline__ void dlb_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t dlb_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_dlb_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
static inline int dlb_post_launch(int n = 1) {
  g_dlb_launches += (unsigned long long)n;
  return (int)cudaPeekAtLastError();
}

template <typename T> struct Vec16 {};   // 16-byte vector of T
template <> struct Vec16<float> { static constexpr int N = 4; };
template <> struct Vec16<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T, int V> struct Pack { T v[V]; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, float (&out)[V]) {
  if constexpr (V == 1) {
    out[0] = (float)p[0];
  } else {
    static_assert(sizeof(T) * V == 16, "vector path is 16 bytes");
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    if constexpr (sizeof(T) == 4) {
      out[0] = __uint_as_float(raw.x); out[1] = __uint_as_float(raw.y);
      out[2] = __uint_as_float(raw.z); out[3] = __uint_as_float(raw.w);
    } else {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); out[2 * i] = f.x; out[2 * i + 1] = f.y; }
    }
  }
}

template <typename T, int V>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const float (&in)[V]) {
  if constexpr (V == 1) {
    p[0] = (T)in[0];
  } else {
    uint4 raw;
    if constexpr (sizeof(T) == 4) {
      raw.x = __float_as_uint(in[0]); raw.y = __float_as_uint(in[1]);
      raw.z = __float_as_uint(in[2]); raw.w = __float_as_uint(in[3]);
    } else {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    }
    *reinterpret_cast<uint4*>(p) = raw;
  }
}

__device__ __forceinline__ unsigned long long dlb_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

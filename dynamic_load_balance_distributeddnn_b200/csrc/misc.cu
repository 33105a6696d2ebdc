// Data-path and fault-injection kernels:
//  * augment: uint8 NHWC batch -> normalised NHWC (bf16 / fp32) with random crop (zero padding)
//    and horizontal flip.  Device-side replacement for the reference's per-image PIL transforms
//    (reference dataloader.py:68-75: RandomCrop(32, padding=4), RandomHorizontalFlip, ToTensor,
//    Normalize) — SURVEY K21 / D11.
//  * burn: deterministic on-device straggler (spins for a requested number of microseconds),
//    the GPU analogue of the reference's `time.sleep` injector (dbs.py:100-104).
#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned hash3(unsigned a, unsigned b, unsigned c) {
  unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du;
  h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
  return h;
}

struct NormConst { float mean[4]; float inv_std[4]; };

template <typename T>
__global__ void __launch_bounds__(256) augment_kernel(const unsigned char* __restrict__ src, T* __restrict__ dst,
                                                      int B, int H, int W, int C, int pad, int flip,
                                                      NormConst nc, unsigned seed, const long long* __restrict__ step_ptr) {
  const long long total = (long long)B * H * W * C;
  const unsigned step = step_ptr ? (unsigned)(*step_ptr) : 0u;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int b = (int)(r / H);
    int sh = h, sw = w;
    if (pad > 0 || flip) {
      const unsigned rnd = hash3(seed, step, (unsigned)b);
      if (pad > 0) {
        const int oy = (int)(rnd % (2 * pad + 1)) - pad;
        const int ox = (int)((rnd >> 8) % (2 * pad + 1)) - pad;
        sh = h + oy; sw = w + ox;
      }
      if (flip && ((rnd >> 20) & 1)) sw = W - 1 - sw;   // flip applied on the cropped window
    }
    float v = 0.f;                                      // zero padding BEFORE normalisation == torchvision
    if (sh >= 0 && sh < H && sw >= 0 && sw < W)
      v = (float)src[(((long long)b * H + sh) * W + sw) * C + c] * (1.f / 255.f);
    dst[i] = (T)((v - nc.mean[c]) * nc.inv_std[c]);
  }
}

__global__ void burn_kernel(const float* __restrict__ usec_ptr, float usec_imm, unsigned long long* sink) {
  const float usec = usec_ptr ? *usec_ptr : usec_imm;
  const unsigned long long t0 = dlb_globaltimer();
  const unsigned long long dur = (unsigned long long)(usec * 1000.f);
  unsigned long long x = 0;
  while (dlb_globaltimer() - t0 < dur) { x += 1; }
  if (sink && x == 0xFFFFFFFFFFFFFFFFull) *sink = x;
}

}  // namespace

DLB_API int dlb_augment(const void* src, void* dst, int out_dtype, int B, int H, int W, int C, int pad, int flip,
                        const float* mean, const float* stdv, unsigned seed, const long long* step_ptr, void* stream) {
  if (C > 4) return -2;
  NormConst nc;
  for (int i = 0; i < 4; ++i) { nc.mean[i] = i < C ? mean[i] : 0.f; nc.inv_std[i] = i < C ? 1.f / stdv[i] : 1.f; }
  const long long total = (long long)B * H * W * C;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) return 0;
  if (out_dtype == DLB_BF16)
    augment_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)src, (__nv_bfloat16*)dst, B, H, W, C, pad, flip, nc, seed, step_ptr);
  else
    augment_kernel<float><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)src, (float*)dst, B, H, W, C, pad, flip, nc, seed, step_ptr);
  return dlb_post_launch();
}

DLB_API int dlb_burn(const float* usec_ptr, float usec, void* stream) {
  burn_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(usec_ptr, usec, nullptr);
  return dlb_post_launch();
}

// ------------------------------------------------------------------------------------------------
// Device-side step timers (graph-capturable): `stamp` stores %globaltimer, `stamp_acc` adds (now - *start) to an
// accumulator.  The trainer brackets the compute part of a step (augment -> forward -> backward -> injected straggle)
// with them on the main stream, BEFORE it joins the communication stream, so the DBS feedback signal is the rank's
// pure compute time even though bucket collectives overlap the backward pass (their in-kernel barrier waits overlap
// compute and must not be subtracted from the step time).
namespace {
__global__ void stamp_kernel(unsigned long long* slot) { *slot = dlb_globaltimer(); }
__global__ void stamp_acc_kernel(const unsigned long long* start, unsigned long long* acc) { *acc += dlb_globaltimer() - *start; }
}  // namespace

DLB_API int dlb_stamp(unsigned long long* slot, void* stream) {
  stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot);
  return dlb_post_launch();
}
DLB_API int dlb_stamp_acc(const unsigned long long* start, unsigned long long* acc, void* stream) {
  stamp_acc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(start, acc);
  return dlb_post_launch();
}

// ------------------------------------------------------------------------------------------------
// CUDA-graph introspection: node counts by type of a captured step graph (cudaGraph_t handle from torch's
// CUDAGraph.raw_cuda_graph()).  out[0] = total nodes, out[1] = kernel, out[2] = memcpy, out[3] = memset, out[4] = host,
// out[5] = event record, out[6] = event wait, out[7] = other, out[8] = number of dependency edges.
DLB_API int dlb_graph_node_counts(void* graph, long long* out) {
  cudaGraph_t g = (cudaGraph_t)graph;
  size_t n = 0;
  cudaError_t e = cudaGraphGetNodes(g, nullptr, &n);
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < 9; ++i) out[i] = 0;
  out[0] = (long long)n;
  if (n == 0) return 0;
  cudaGraphNode_t* nodes = new cudaGraphNode_t[n];
  e = cudaGraphGetNodes(g, nodes, &n);
  if (e == cudaSuccess) {
    for (size_t i = 0; i < n; ++i) {
      cudaGraphNodeType t;
      if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess) { out[7]++; continue; }
      switch (t) {
        case cudaGraphNodeTypeKernel: out[1]++; break;
        case cudaGraphNodeTypeMemcpy: out[2]++; break;
        case cudaGraphNodeTypeMemset: out[3]++; break;
        case cudaGraphNodeTypeHost: out[4]++; break;
        case cudaGraphNodeTypeEventRecord: out[5]++; break;
        case cudaGraphNodeTypeWaitEvent: out[6]++; break;
        default: out[7]++; break;
      }
    }
    size_t ne = 0;
    if (cudaGraphGetEdges(g, nullptr, nullptr, &ne) == cudaSuccess) out[8] = (long long)ne;
  }
  delete[] nodes;
  return (int)e;
}

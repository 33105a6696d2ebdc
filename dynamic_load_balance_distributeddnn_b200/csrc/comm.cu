// Gradient collectives over NVLink/NVSwitch peer memory — hand-written for sm_100a.
//
// The reference's hot path is, per parameter tensor: `w * grad` (elementwise kernel) ->
// `dist.all_reduce(async)` -> host `wait()` (reference dbs.py:291-301),
// and once per epoch a host-staged ring all-gather of one float per rank (dbs.py:479-499; C5/K3).
// Here each gradient bucket is ONE kernel that applies every source rank's DBS weight
// w_p = local_bs_p / global_bs while it moves the data:
//
//   one-shot  : out[i] = sum_p w_p * peer_p[i]           every rank pulls the whole bucket from all
//               peers over NVLink, fixed summation order (bit-identical replicas); 1 entry + 1 exit
//               barrier; best for latency-bound buckets.
//   two-shot  : rank r reduces chunk r (pull, weighted, fp32 accumulate) and pushes the result into
//               every peer's output buffer (P2P stores); reduce-scatter + all-gather in one launch.
//   nvls      : rank r pre-scales its own bucket in place (w_r; fused staging), then chunk r is reduced
//               inside the NVSwitch with `multimem.ld_reduce` and broadcast with `multimem.st`.
//
// Cross-GPU synchronisation uses per-(block, source-rank) 32-bit slots in a symmetric flag buffer with a
// self-resetting CAS protocol (sender 0->1 with release.sys, receiver 1->0 with acquire.sys), so the
// kernels are CUDA-graph replay safe (no host-side epoch argument).  The entry barrier's spin time is
// measured with %globaltimer and accumulated on the device: that is the "straggler wait" the DBS
// rebalancer is meant to shrink, measured without any host synchronisation (fixes SURVEY D10).
// Every spin loop has a watchdog so a dead peer raises an error flag instead of hanging the GPU.
#include "common.cuh"

namespace {

constexpr int kMaxWorld = 16;
constexpr int kMaxBlocks = 64;
constexpr int kCommThreads = 512;
constexpr unsigned long long kDefaultTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

struct CommArgs {
  void* in[kMaxWorld];          // peers' symmetric input buffers (element type = wire type)
  void* out[kMaxWorld];         // peers' symmetric output buffers
  uint32_t* flags[kMaxWorld];   // peers' flag buffers: [channel][kMaxBlocks][kMaxWorld]
  void* mc_in;                  // multicast VA of `in` (0 if NVLS unavailable)
  void* mc_out;
  float weights[kMaxWorld];     // used when weights_dev == nullptr and use_weights != 0
  const float* weights_dev;     // per-rank DBS weights on the device (graph-replay friendly)
  int use_weights;
  int rank, world;
  unsigned long long* wait_ns;  // accumulated entry-barrier wait (block 0)
  int* err_flag;
  unsigned long long timeout_ns;
};

// Optimizer step fused behind the collective (momentum SGD on the flat fp32 master / momentum buffers + bf16 shadow refresh
// + clearing of the gradient accumulation buffer): runs as the last phase of the allreduce kernel of each bucket, so
// "allreduce + optimizer" is ONE launch per bucket (reference: per-tensor scale + all_reduce + wait, then torch.optim.SGD's
// per-tensor loop, dbs.py:291-301,238).
struct SgdArgs {
  float* master;            // fp32 parameters (flat, same element indexing as the gradient buffers); nullptr = no fused step
  float* mom;
  __nv_bfloat16* shadow;    // bf16 compute copy of the parameters (may be null)
  const float* lr;          // device scalar
  float mu, wd;
  float* zero_in;           // this rank's gradient accumulation buffer, cleared after use (may be null)
};

__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
}

__device__ __forceinline__ bool put_signal(uint32_t* addr, unsigned long long timeout_ns, int* err) {
  const unsigned long long t0 = dlb_globaltimer();
  unsigned spins = 0;
  while (cas_release_sys(addr, 0u, 1u) != 0u) {
    if ((++spins & 0x3FFu) == 0 && dlb_globaltimer() - t0 > timeout_ns) { if (err) atomicExch(err, 1); return false; }
  }
  return true;
}
__device__ __forceinline__ bool wait_signal(uint32_t* addr, unsigned long long timeout_ns, int* err) {
  const unsigned long long t0 = dlb_globaltimer();
  unsigned spins = 0;
  while (cas_acquire_sys(addr, 1u, 0u) != 1u) {
    if ((++spins & 0x3FFu) == 0 && dlb_globaltimer() - t0 > timeout_ns) { if (err) atomicExch(err, 2); return false; }
  }
  return true;
}

// All-ranks barrier for block `blockIdx.x` on flag channel `ch`.  Must be called by the whole CTA.
__device__ __forceinline__ void block_barrier(const CommArgs& a, int ch) {
  __syncthreads();                                   // CTA's prior writes happen-before the release below
  const int t = threadIdx.x;
  if (t < a.world && t != a.rank) {
    const size_t slot = ((size_t)ch * kMaxBlocks + blockIdx.x) * kMaxWorld;
    put_signal(a.flags[t] + slot + a.rank, a.timeout_ns, a.err_flag);        // remote slot [me]
    wait_signal(a.flags[a.rank] + slot + t, a.timeout_ns, a.err_flag);       // local slot [peer]
  }
  __syncthreads();
}

template <typename T> struct Wire {};
template <> struct Wire<float> {
  static constexpr int V = 4;
  __device__ static __forceinline__ void load(const float* p, float (&o)[4]) {
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p));
  }
  __device__ static __forceinline__ void store(float* p, const float (&v)[4]) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  }
  __device__ static __forceinline__ void mc_reduce(const float* p, float (&o)[4]) {
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "l"(p) : "memory");
  }
  __device__ static __forceinline__ void mc_store(float* p, const float (&v)[4]) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
  }
};
template <> struct Wire<__nv_bfloat16> {
  static constexpr int V = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    uint32_t r[4];
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r[i])); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); r[i] = *reinterpret_cast<uint32_t*>(&h); }
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  }
  __device__ static __forceinline__ void mc_reduce(const __nv_bfloat16* p, float (&o)[8]) {
    uint32_t r[4];
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p) : "memory");
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&r[i])); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void mc_store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]); r[i] = *reinterpret_cast<uint32_t*>(&h); }
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
  }
};

enum Algo : int { ONESHOT = 0, TWOSHOT = 1, NVLS = 2 };

// Weighted pull-reduce of `nvec` vectors starting at element offset `e0`, results delivered by DST.
template <typename T, int WORLD, typename DST>
__device__ __forceinline__ void pull_reduce(const CommArgs& a, const float (&w)[kMaxWorld], long long e0,
                                            long long nvec, DST&& deliver) {
  constexpr int V = Wire<T>::V;
  constexpr int UNROLL = (WORLD <= 4) ? 4 : 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const int world = WORLD > 0 ? WORLD : a.world;
  for (long long v0 = tid; v0 < nvec; v0 += nthreads * UNROLL) {
    float acc[UNROLL][V];
    float tmp[UNROLL][V];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
#pragma unroll
      for (int k = 0; k < V; ++k) acc[u][k] = 0.f;
#pragma unroll
    for (int p = 0; p < (WORLD > 0 ? WORLD : kMaxWorld); ++p) {
      if (p < world) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const long long v = v0 + (long long)u * nthreads;
          if (v < nvec) Wire<T>::load((const T*)a.in[p] + e0 + v * V, tmp[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
          for (int k = 0; k < V; ++k) acc[u][k] = fmaf(w[p], tmp[u][k], acc[u][k]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long v = v0 + (long long)u * nthreads;
      if (v < nvec) deliver(e0 + v * V, acc[u]);
    }
  }
}

template <typename T, int ALGO, int WORLD>
__global__ void __launch_bounds__(kCommThreads, 1)
weighted_allreduce_kernel(const __grid_constant__ CommArgs a, long long offset, long long count, const SgdArgs sgd) {
  constexpr int V = Wire<T>::V;
  const int world = WORLD > 0 ? WORLD : a.world;
  float w[kMaxWorld];
#pragma unroll
  for (int p = 0; p < kMaxWorld; ++p)
    w[p] = !a.use_weights ? 1.f : (a.weights_dev ? (p < world ? a.weights_dev[p] : 0.f) : a.weights[p]);

  if constexpr (ALGO == NVLS) {
    // fused staging: scale my own contribution in place before anyone reduces it in the switch
    if (a.use_weights) {
      // The only cross-rank ordering before the reduce phase is block_barrier(a, 0), which orders block b here against
      // block b on every peer.  Peer r's block b reduces the vectors  c0_r + b*blockDim + t + k*grid*blockDim  of ITS
      // chunk [c0_r, c1_r), so block b must pre-scale exactly those vectors of every chunk (same grid on all ranks):
      // a plain grid-stride loop over the whole bucket would let a peer read elements another block has not scaled yet.
      T* mine = (T*)a.in[a.rank] + offset;
      const float wr = w[a.rank];
      const long long nvec = count / V;
      const long long per = (nvec + world - 1) / world;
      for (int r = 0; r < world; ++r) {
        const long long c0 = min(nvec, per * r), c1 = min(nvec, c0 + per);
        for (long long v = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < c1; v += (long long)gridDim.x * blockDim.x) {
          float x[V];
          Wire<T>::load(mine + v * V, x);
#pragma unroll
          for (int k = 0; k < V; ++k) x[k] *= wr;
          Wire<T>::store(mine + v * V, x);
        }
      }
      __threadfence_system();
    }
  }

  // ---- entry barrier: every peer's bucket is ready; time spent here = straggler wait ----------
  unsigned long long t0 = 0;
  if (blockIdx.x == 0 && threadIdx.x == 0) t0 = dlb_globaltimer();
  block_barrier(a, 0);
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.wait_ns) atomicAdd(a.wait_ns, dlb_globaltimer() - t0);

  if constexpr (ALGO == ONESHOT) {
    T* out = (T*)a.out[a.rank];
    pull_reduce<T, WORLD>(a, w, offset, count / V, [&](long long e, const float (&r)[V]) { Wire<T>::store(out + e, r); });
  } else {
    // chunk owned by this rank: [c0, c1) in vectors
    const long long nvec = count / V;
    const long long per = (nvec + world - 1) / world;
    const long long c0 = min(nvec, per * a.rank), c1 = min(nvec, c0 + per);
    if constexpr (ALGO == TWOSHOT) {
      pull_reduce<T, WORLD>(a, w, offset + c0 * V, c1 - c0, [&](long long e, const float (&r)[V]) {
#pragma unroll
        for (int p = 0; p < (WORLD > 0 ? WORLD : kMaxWorld); ++p)
          if (p < world) Wire<T>::store((T*)a.out[p] + e, r);           // P2P push to every replica
      });
    } else {
      const T* mc_in = (const T*)a.mc_in + offset;
      T* mc_out = (T*)a.mc_out + offset;
      for (long long v = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; v < c1; v += (long long)gridDim.x * blockDim.x) {
        float r[V];
        Wire<T>::mc_reduce(mc_in + v * V, r);
        Wire<T>::mc_store(mc_out + v * V, r);
      }
    }
  }
  // ---- exit barrier: all pushes have landed / all peers are done reading my input --------------
  __threadfence_system();
  block_barrier(a, 1);

  // ---- fused optimizer step ---------------------------------------------------------------------
  // block_barrier orders block b here only against block b of every peer.  Peer r's block b produced / consumed exactly
  // the vectors  c0_r + b*blockDim + t + k*grid*blockDim  of chunk r (two-shot / NVLS), and this rank's block b produced
  // the vectors  b*blockDim + t + k*grid*blockDim  of the whole bucket (one-shot): block b therefore updates precisely
  // those elements -- no grid-wide synchronisation is needed and every element is updated exactly once.
  if (sgd.master != nullptr) {
    const float lr = *sgd.lr;
    const T* red = (const T*)a.out[a.rank] + offset;
    auto update = [&](long long v) {
      float g[V];
      Wire<T>::load(red + v * V, g);
      const long long e = offset + v * V;
#pragma unroll
      for (int q = 0; q < V; q += 4) {
        float4 pv = *reinterpret_cast<float4*>(sgd.master + e + q);
        float4 mv = *reinterpret_cast<float4*>(sgd.mom + e + q);
        mv.x = fmaf(sgd.mu, mv.x, fmaf(sgd.wd, pv.x, g[q]));     mv.y = fmaf(sgd.mu, mv.y, fmaf(sgd.wd, pv.y, g[q + 1]));
        mv.z = fmaf(sgd.mu, mv.z, fmaf(sgd.wd, pv.z, g[q + 2])); mv.w = fmaf(sgd.mu, mv.w, fmaf(sgd.wd, pv.w, g[q + 3]));
        pv.x = fmaf(-lr, mv.x, pv.x); pv.y = fmaf(-lr, mv.y, pv.y); pv.z = fmaf(-lr, mv.z, pv.z); pv.w = fmaf(-lr, mv.w, pv.w);
        *reinterpret_cast<float4*>(sgd.master + e + q) = pv;
        *reinterpret_cast<float4*>(sgd.mom + e + q) = mv;
        if (sgd.shadow) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
          uint2 raw; raw.x = *reinterpret_cast<unsigned*>(&lo); raw.y = *reinterpret_cast<unsigned*>(&hi);
          *reinterpret_cast<uint2*>(sgd.shadow + e + q) = raw;
        }
        if (sgd.zero_in) *reinterpret_cast<float4*>(sgd.zero_in + e + q) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    const long long nvec = count / V;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long first = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (ALGO == ONESHOT) {
      for (long long v = first; v < nvec; v += stride) update(v);
    } else {
      const long long per = (nvec + world - 1) / world;
      for (int r = 0; r < world; ++r) {
        const long long c0 = min(nvec, per * r), c1 = min(nvec, c0 + per);
        for (long long v = c0 + first; v < c1; v += stride) update(v);
      }
    }
  }
}

// One launch: store my scalar into slot [rank] of every peer's table, barrier, done.
__global__ void time_allgather_kernel(const __grid_constant__ CommArgs a, const float* __restrict__ my_value,
                                      float* const* tables_unused, int table_offset_floats) {
  const int t = threadIdx.x;
  if (t < a.world) {
    float* peer_table = (float*)(a.out[t]) + table_offset_floats;   // `out` carries the peers' table buffers here
    const float v = *my_value;
    asm volatile("st.global.relaxed.sys.f32 [%0], %1;" :: "l"(peer_table + a.rank), "f"(v) : "memory");
  }
  __threadfence_system();
  block_barrier(a, 2);
}

__global__ void barrier_kernel(const __grid_constant__ CommArgs a, int channel) {
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) t0 = dlb_globaltimer();
  block_barrier(a, channel);
  if (threadIdx.x == 0 && a.wait_ns) atomicAdd(a.wait_ns, dlb_globaltimer() - t0);
}

struct CommCtx {
  CommArgs args;
};

template <typename T, int ALGO>
void launch_allreduce(const CommArgs& a, long long offset, long long count, int blocks, cudaStream_t st, const SgdArgs& sgd) {
  switch (a.world) {
    case 1: weighted_allreduce_kernel<T, ALGO, 1><<<blocks, kCommThreads, 0, st>>>(a, offset, count, sgd); break;
    case 2: weighted_allreduce_kernel<T, ALGO, 2><<<blocks, kCommThreads, 0, st>>>(a, offset, count, sgd); break;
    case 4: weighted_allreduce_kernel<T, ALGO, 4><<<blocks, kCommThreads, 0, st>>>(a, offset, count, sgd); break;
    case 8: weighted_allreduce_kernel<T, ALGO, 8><<<blocks, kCommThreads, 0, st>>>(a, offset, count, sgd); break;
    default: weighted_allreduce_kernel<T, ALGO, 0><<<blocks, kCommThreads, 0, st>>>(a, offset, count, sgd); break;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI.  A context holds the peer pointer tables; buffers are allocated/exchanged by Python
// (parallel/symm.py: torch symmetric memory or CUDA-IPC fallback).
DLB_API void* dlb_comm_create(int rank, int world, const unsigned long long* in_ptrs, const unsigned long long* out_ptrs,
                              const unsigned long long* flag_ptrs, unsigned long long mc_in, unsigned long long mc_out,
                              unsigned long long wait_ns_ptr, unsigned long long err_flag_ptr) {
  if (world > kMaxWorld || world < 1) return nullptr;
  CommCtx* c = new CommCtx();
  CommArgs& a = c->args;
  for (int i = 0; i < kMaxWorld; ++i) {
    a.in[i] = i < world ? (void*)in_ptrs[i] : nullptr;
    a.out[i] = i < world ? (void*)out_ptrs[i] : nullptr;
    a.flags[i] = i < world ? (uint32_t*)flag_ptrs[i] : nullptr;
    a.weights[i] = i < world ? 1.f / world : 0.f;
  }
  a.mc_in = (void*)mc_in;
  a.mc_out = (void*)mc_out;
  a.weights_dev = nullptr;
  a.use_weights = 0;
  a.rank = rank;
  a.world = world;
  a.wait_ns = (unsigned long long*)wait_ns_ptr;
  a.err_flag = (int*)err_flag_ptr;
  a.timeout_ns = kDefaultTimeoutNs;
  return c;
}

DLB_API void dlb_comm_destroy(void* ctx) { delete (CommCtx*)ctx; }
DLB_API void dlb_comm_set_timeout(void* ctx, double seconds) { ((CommCtx*)ctx)->args.timeout_ns = (unsigned long long)(seconds * 1e9); }
DLB_API int dlb_comm_flag_words() { return 4 * kMaxBlocks * kMaxWorld; }      // channels: 0 entry, 1 exit, 2 barrier / time table, 3 gate
DLB_API int dlb_comm_max_blocks() { return kMaxBlocks; }

// algo: 0 one-shot, 1 two-shot, 2 nvls.  wire: DLB_F32 / DLB_BF16.  offset/count in elements; count must be
// a multiple of the 16-byte vector width.  weights_dev: device float[world] (or null with host weights / none).
static int weighted_allreduce_impl(void* ctx, int algo, int wire, long long offset, long long count, int blocks,
                                   const float* weights_dev, const float* weights_host, const SgdArgs& sgd, void* stream) {
  CommCtx* c = (CommCtx*)ctx;
  CommArgs a = c->args;
  a.weights_dev = weights_dev;
  a.use_weights = (weights_dev != nullptr || weights_host != nullptr) ? 1 : 0;
  if (weights_host) for (int i = 0; i < a.world; ++i) a.weights[i] = weights_host[i];
  const int V = wire == DLB_BF16 ? 8 : 4;
  if (count % V) return -3;
  if (blocks < 1) blocks = 1;
  if (blocks > kMaxBlocks) blocks = kMaxBlocks;
  if (algo == NVLS && (!a.mc_in || !a.mc_out)) return -4;
  cudaStream_t st = (cudaStream_t)stream;
#define GO(T)                                                              \
  do {                                                                     \
    if (algo == ONESHOT) launch_allreduce<T, ONESHOT>(a, offset, count, blocks, st, sgd);      \
    else if (algo == TWOSHOT) launch_allreduce<T, TWOSHOT>(a, offset, count, blocks, st, sgd); \
    else launch_allreduce<T, NVLS>(a, offset, count, blocks, st, sgd);          \
  } while (0)
  if (wire == DLB_BF16) GO(__nv_bfloat16); else GO(float);
#undef GO
  return dlb_post_launch();
}

DLB_API int dlb_weighted_allreduce(void* ctx, int algo, int wire, long long offset, long long count, int blocks,
                                   const float* weights_dev, const float* weights_host, void* stream) {
  SgdArgs none = {};
  return weighted_allreduce_impl(ctx, algo, wire, offset, count, blocks, weights_dev, weights_host, none, stream);
}

// Same collective with the optimizer step fused behind it (see SgdArgs): master/mom fp32 [numel], shadow bf16 or null,
// zero_in = this rank's gradient accumulation buffer to clear (or null).
DLB_API int dlb_weighted_allreduce_sgd(void* ctx, int algo, int wire, long long offset, long long count, int blocks,
                                       const float* weights_dev, float* master, float* mom, void* shadow, const float* lr_ptr,
                                       float* zero_in, float momentum, float weight_decay, void* unused, void* stream) {
  (void)unused;
  SgdArgs sgd;
  sgd.master = master; sgd.mom = mom; sgd.shadow = (__nv_bfloat16*)shadow; sgd.lr = lr_ptr; sgd.mu = momentum; sgd.wd = weight_decay;
  sgd.zero_in = zero_in;
  return weighted_allreduce_impl(ctx, algo, wire, offset, count, blocks, weights_dev, nullptr, sgd, stream);
}

// tables: `out` pointers of a context created over the (small) symmetric time-table buffers.
DLB_API int dlb_time_allgather(void* ctx, const float* my_value, int table_offset_floats, void* stream) {
  CommCtx* c = (CommCtx*)ctx;
  time_allgather_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->args, my_value, nullptr, table_offset_floats);
  return dlb_post_launch();
}

// Bucket gate: a ONE-WARP kernel that waits (flag channel 3) until every peer has reached the same bucket, launched on the
// communication stream right before that bucket's allreduce.  The allreduce kernel itself is 32-64 CTAs x 512 threads; while
// it spins at its entry barrier for a late peer it pins those SMs, and the 1-CTA-per-SM tensor-core kernels of the backward
// pass (50-64 K registers each) cannot co-reside with it: the fast rank's backward stalls bucket by bucket on the straggler
// (measured at 2 GPUs: "compute" 17.7 ms on the fast rank vs 18.4 on the straggler where 15.6 was expected), which also
// hides the imbalance from the DBS time signal.  With the gate the wait costs one warp; the straggler wait is accounted here.
DLB_API int dlb_comm_gate(void* ctx, void* stream) {
  CommCtx* c = (CommCtx*)ctx;
  barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->args, 3);
  return dlb_post_launch();
}

DLB_API int dlb_device_barrier(void* ctx, int channel, void* stream) {
  CommCtx* c = (CommCtx*)ctx;
  barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(c->args, channel);
  return dlb_post_launch();
}

// tcgen05 / TMEM / TMA / mbarrier primitives shared by the tensor-core kernels of this library (sm_100a inline PTX).
// Everything lives in an anonymous namespace: each translation unit gets its own internal copy.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr uint32_t kSpinLimit = 1u << 20;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();      // watchdog: never hang the GPU on a protocol bug
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// explicit shared-space accesses: the 1024-byte re-alignment of the dynamic smem base hides the address space
// from the compiler, which otherwise emits generic LD.E/ST.E (slow path) for the operand-transform loops
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B canonical layout: 8-row groups 1024 B apart (SBO), rows 128 B apart inside a group.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (ignored for swizzled K-major), canonical 1
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void tma_load_3d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// fp32 operands in shared memory consumed as TF32 (10-bit mantissa, the low 13 bits are ignored), fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Element traits of the tensor-core kernels.  In BYTES the two operand types are identical: one SWIZZLE_128B atom row
// is 128 bytes of K (64 bf16 / 32 fp32), one UMMA consumes 32 bytes of K per row (UMMA_K = 16 bf16 / 8 tf32), so the
// shared-memory layouts, descriptor strides and the K-major "+2 per MMA" descriptor advance do not depend on the type.
template <typename E> struct Elt;
template <> struct Elt<__nv_bfloat16> {
  static constexpr int kBytes = 2;
  static constexpr int kAtom = 64;          // elements per 128-byte swizzle row
  static constexpr int kUmmaK = 16;
  static constexpr int kPer16 = 8;          // elements per 16-byte chunk
  static constexpr bool kMn32 = false;      // MN-major operands use the plain SWIZZLE_128B layout
  static constexpr uint32_t kFmt = 1;       // UMMA instruction-descriptor operand format: BF16
  static constexpr CUtensorMapDataType kTmap = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { umma_f16(d, a, b, idesc, acc); }
};
template <> struct Elt<float> {
  static constexpr int kBytes = 4;
  static constexpr int kAtom = 32;
  static constexpr int kUmmaK = 8;
  static constexpr int kPer16 = 4;
  static constexpr bool kMn32 = true;       // MN-major operands need SWIZZLE_128B_BASE32B (see make_smem_desc_mn)
  static constexpr uint32_t kFmt = 2;       // TF32
  static constexpr CUtensorMapDataType kTmap = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  __device__ static __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) { umma_tf32(d, a, b, idesc, acc); }
};

// relu(a*x + b) on one 16-byte chunk (8 bf16 or 4 fp32) with per-channel coefficients held in registers
template <typename E>
__device__ __forceinline__ void affine_relu_chunk(uint4& raw, const float (&ca)[8], const float (&cb)[8]) {
  if constexpr (sizeof(E) == 2) {
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(h[i]);
      h[i] = __floats2bfloat162_rn(fmaxf(fmaf(ca[2 * i], f.x, cb[2 * i]), 0.f), fmaxf(fmaf(ca[2 * i + 1], f.y, cb[2 * i + 1]), 0.f));
    }
  } else {
    raw.x = __float_as_uint(fmaxf(fmaf(ca[0], __uint_as_float(raw.x), cb[0]), 0.f));
    raw.y = __float_as_uint(fmaxf(fmaf(ca[1], __uint_as_float(raw.y), cb[1]), 0.f));
    raw.z = __float_as_uint(fmaxf(fmaf(ca[2], __uint_as_float(raw.z), cb[2]), 0.f));
    raw.w = __float_as_uint(fmaxf(fmaf(ca[3], __uint_as_float(raw.w), cb[3]), 0.f));
  }
}
// load the kPer16 coefficients of one chunk (pointers 16-byte aligned)
template <typename E>
__device__ __forceinline__ void load_coef(const float* pa, const float* pb, float (&ca)[8], float (&cb)[8]) {
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(pa)), b0 = __ldg(reinterpret_cast<const float4*>(pb));
  ca[0] = a0.x; ca[1] = a0.y; ca[2] = a0.z; ca[3] = a0.w; cb[0] = b0.x; cb[1] = b0.y; cb[2] = b0.z; cb[3] = b0.w;
  if constexpr (sizeof(E) == 2) {
    const float4 a1 = __ldg(reinterpret_cast<const float4*>(pa) + 1), b1 = __ldg(reinterpret_cast<const float4*>(pb) + 1);
    ca[4] = a1.x; ca[5] = a1.y; ca[6] = a1.z; ca[7] = a1.w; cb[4] = b1.x; cb[5] = b1.y; cb[6] = b1.z; cb[7] = b1.w;
  }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// MN-major operand descriptor.  16-bit types: SWIZZLE_128B, 8-row K groups 1024 B apart.  32-bit types (TF32): the only legal
// MN-major layout is SWIZZLE_128B_BASE32B (layout type 1): 32-byte chunks of a 128-byte row XOR-ed with (row & 3), 4-row K
// groups 512 B apart -- written by TMA with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (CUTLASS: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout").
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, uint32_t lbo_bytes = 8192, bool base32b = false) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;  // LBO: next 128-byte-wide MN group (one TMA box further)
  d |= (uint64_t)((base32b ? 512 : 1024) >> 4) << 32;       // SBO: next K group (8 rows, or 4 rows for BASE32B)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base32b ? 1 : 2) << 61; // SWIZZLE_128B_BASE32B / SWIZZLE_128B
  return d;
}
// byte offset of logical 16-byte chunk `j` (0..7) of 128-byte row `r` inside a swizzled tile
template <bool BASE32B>
__device__ __forceinline__ uint32_t swz_chunk(int r, int j) {
  if constexpr (BASE32B) return (uint32_t)(r * 128 + (((((j >> 1) ^ (r & 3)) << 1) | (j & 1)) << 4));
  else return (uint32_t)(r * 128 + ((j ^ (r & 7)) << 4));
}

// ---- host side: TMA tensor maps (driver entry point fetched through the runtime, no libcuda link dependency) ----
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    cudaDriverEntryPointQueryResult qres;
    void* ptr = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

// 2-D bf16 tensor map: inner dim `cols` (contiguous), outer dim `rows` with row stride `ld` elements.
inline int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                    int esize = 2, bool atom32 = false) {
  // inner box is always 128 bytes = the swizzle span (64 bf16 or 32 fp32)
  auto enc = get_encode();
  if (!enc) return -10;
  // the driver-API encoder needs a current context on THIS thread (autograd runs backward on its own threads,
  // where a runtime call may not have bound the primary context yet)
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(0); ctx_bound = true; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("DLB_DEBUG_TMAP"))
    fprintf(stderr, "[dlb] cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%lld cols=%lld ld=%lld box_rows=%d\n", (int)r, ptr, rows, cols, ld, box_rows);
  return r == CUDA_SUCCESS ? 0 : -11;
}

// 2-D fp32 tensor map without swizzle (small side tables such as the GroupNorm coefficient rows): OOB rows/cols read as zero
inline int make_map_f32_plain(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  auto enc = get_encode();
  if (!enc) return -10;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(0); ctx_bound = true; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -11;
}

// general rank-n bf16 tensor map (dims/strides innermost first; strides in elements for dims 1..n-1)
inline int make_map_nd(CUtensorMap* map, const void* ptr, int rank, const long long* dims, const long long* strides, const int* box,
                       int esize = 2, bool atom32 = false) {
  auto enc = get_encode();
  if (!enc) return -10;
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(0); ctx_bound = true; }
  cuuint64_t d[5]; cuuint64_t sbytes[4]; cuuint32_t b[5]; cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; b[i] = (cuuint32_t)box[i]; es[i] = 1; }
  for (int i = 1; i < rank; ++i) sbytes[i - 1] = (cuuint64_t)strides[i] * (cuuint64_t)esize;
  CUresult r = enc(map, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), d, sbytes, b, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS && getenv("DLB_DEBUG_TMAP"))
    fprintf(stderr, "[dlb] cuTensorMapEncodeTiled(rank %d) failed (%d)\n", rank, (int)r);
  return r == CUDA_SUCCESS ? 0 : -11;
}

}  // namespace

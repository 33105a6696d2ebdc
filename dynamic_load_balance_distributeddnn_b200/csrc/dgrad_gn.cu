// Fused data-gradient GEMM + GroupNorm(+ReLU) backward for a pre-activation 1x1 convolution
// (GN -> ReLU -> conv1x1, the first half of the DenseNet bottleneck; reference Net/Densenet.py:13-19).
//
// Default backward of the dense-block bottleneck (DLB_FUSED_DGRAD=0 restores the chain dgrad GEMM -> fused GN backward).
// Element type E: bf16 (kind::f16) or fp32 storage with TF32 math (kind::tf32).  A tile is always TWO 128-byte column
// groups wide (128 bf16 / 64 fp32 channels), so every shared-memory box, swizzle pattern and barrier byte count below is
// the same for both types.
//
// Why: with  A = relu(GN(x)),  y = A * W^T  the input gradient is
//     dA = dY * W                                        (GEMM, K = Cmid = 128)
//     dz = dA * [ca*x + cb > 0]                          (ReLU mask recomputed from the forward's affine coefficients)
//     dX += k1[n,c]*dz + k2[n,c]*x + k3[n,c]             (GroupNorm backward; k2,k3 need the per-(n,c) sums of dz, dz*x)
// The three-kernel chain moves 7 x |x| bytes per layer (write dA; read dA,x; read dA,x,dX; write dX), and |x| grows with
// every layer of a dense block: it is 40 % of the DenseNet-121 step (profiles/r1_13).  The GEMM is cheap (K = 128), so this
// kernel runs it TWICE and never writes dA:
//     MODE 1 (statistics): D tile in TMEM -> epilogue reads the x tile (TMA), accumulates sum(dz), sum(dz*x) per (n,c)
//     MODE 2 (apply)     : D tile again   -> epilogue reads x and dX tiles (TMA), writes dX tile (TMA store, in place)
// = 4 x |x| bytes.  Structure, protocols and primitives (tc_common.cuh) are those of gemm_tc.cu (B given as [K][N], MN-major descriptor); new here:
// the per-tile auxiliary TMA loads feeding the epilogue, double buffered with their own full/empty mbarriers, and
// 8 epilogue warps (two per TMEM lane quadrant, 64 columns each).
#include "tc_common.cuh"

namespace {

constexpr int BM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreadsDG = 384;                 // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
constexpr int kABytes = BM * 128;               // 16 KB: [128 rows][128 bytes of k]
constexpr int kHalfBytes = BM * 128;            // one [128 rows][128 bytes of channels] box = 16 KB
constexpr int kTensorBytes = 2 * kHalfBytes;    // x (or dX) tile: 32 KB

template <typename E, int MODE> struct DCfg {
  using EL = Elt<E>;
  static constexpr int AT = EL::kAtom;            // channels per 128-byte column group (64 bf16 / 32 fp32)
  static constexpr int BN = 2 * AT, BK = AT, UMMA_K = EL::kUmmaK;
  static constexpr int kMnBox = BK * 128;         // one MN-major B box: [BK k-rows][128 bytes of n]
  static constexpr int kBBytes = 2 * kMnBox;      // 16 KB bf16 / 8 KB fp32
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = MODE == 2 ? 3 : 4;
  static constexpr int kAuxSlotBytes = (MODE == 2 ? 2 : 1) * kTensorBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2 * kAuxSlotBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// four consecutive channels (index group j4 of the 32-column chunk) out of / into the row's 16-byte chunks
template <typename E>
__device__ __forceinline__ void unpack4(const uint4* ch, int j4, float (&o)[4]) {
  if constexpr (sizeof(E) == 2) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&ch[j4 >> 1]) + (j4 & 1) * 2;
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[0]));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[1]));
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  } else {
    o[0] = __uint_as_float(ch[j4].x); o[1] = __uint_as_float(ch[j4].y);
    o[2] = __uint_as_float(ch[j4].z); o[3] = __uint_as_float(ch[j4].w);
  }
}
template <typename E>
__device__ __forceinline__ void pack4(uint4* ch, int j4, const float (&o)[4]) {
  if constexpr (sizeof(E) == 2) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&ch[j4 >> 1]) + (j4 & 1) * 2;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(o[2], o[3]);
    w[0] = *reinterpret_cast<uint32_t*>(&h0);
    w[1] = *reinterpret_cast<uint32_t*>(&h1);
  } else {
    ch[j4] = make_uint4(__float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
  }
}

struct DgParams {
  int M, N, K;                 // pixels, input channels of the conv (width of x / dX), mid channels (width of dY)
  int rows_per_sample;         // H*W, a multiple of 32 (so a warp's 32 rows belong to one sample)
  const float* ca;             // [samples][cld] forward affine coefficients: relu mask = (ca*x + cb > 0); k1 = ca
  const float* cb;
  const float* k2;             // MODE 2: [samples][cld]
  const float* k3;
  long long cld;               // multiple of 64, >= ceil64(N), zero padded
  float* table;                // MODE 1: [samples][table_ns] (sum dz, sum dz*x) per channel, pre-zeroed
  long long table_ns;
};

template <typename E, int MODE>
__global__ void __launch_bounds__(kThreadsDG, 1)
dgrad_gn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_gl,
                const __grid_constant__ CUtensorMap tmap_gs, const DgParams p) {
  using C = DCfg<E, MODE>;
  using EL = Elt<E>;
  constexpr int BN = C::BN, BK = C::BK, AT = C::AT, UMMA_K = C::UMMA_K;
  constexpr int kStageBytes = C::kStageBytes, kBBytes = C::kBBytes, kMnBox = C::kMnBox;
  constexpr int kCh = 32 / EL::kPer16;                 // 16-byte chunks per 32 channels (4 bf16 / 8 fp32)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* aux_base = smem + C::kStages * kStageBytes;                     // 1024-byte aligned
  uint8_t* bar_base = aux_base + 2 * C::kAuxSlotBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full = empty_bar + C::kStages;                            // [2]
  uint64_t* tmem_empty = tmem_full + 2;                                    // [2]
  uint64_t* aux_full = tmem_empty + 2;                                     // [2] x (+dX) tile landed
  uint64_t* aux_empty = aux_full + 2;                                      // [2] epilogue done with the slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    if constexpr (MODE == 2) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_gl) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_gs) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), kEpiWarps);
      mbar_init(smem_u32(&aux_full[a]), 1);
      mbar_init(smem_u32(&aux_empty[a]), kEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  dlb_pdl_wait();

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int slot = 0; uint32_t aux_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
        const bool h1 = n0 + AT < p.N;                           // second column group intersects the tensor
        // epilogue operands of this tile first: they are the bulk of the bytes and the epilogue is the long pole
        mbar_wait(smem_u32(&aux_empty[slot]), aux_phase ^ 1);
        const uint32_t ab = smem_u32(aux_base + slot * C::kAuxSlotBytes);
        const uint32_t af = smem_u32(&aux_full[slot]);
        mbar_expect_tx(af, (uint32_t)((MODE == 2 ? 2 : 1) * (h1 ? kTensorBytes : kHalfBytes)));
        tma_load_2d(ab, &tmap_x, af, n0, m0);
        if (h1) tma_load_2d(ab + kHalfBytes, &tmap_x, af, n0 + AT, m0);
        if constexpr (MODE == 2) {
          tma_load_2d(ab + kTensorBytes, &tmap_gl, af, n0, m0);
          if (h1) tma_load_2d(ab + kTensorBytes + kHalfBytes, &tmap_gl, af, n0 + AT, m0);
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint32_t sb = sa + kABytes;
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, (uint32_t)(kABytes + (h1 ? kBBytes : kBBytes / 2)));
          tma_load_2d(sa, &tmap_a, fb, kb * BK, m0);
          tma_load_2d(sb, &tmap_b, fb, n0, kb * BK);
          if (h1) tma_load_2d(sb + kMnBox, &tmap_b, fb, n0 + AT, kb * BK);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        slot ^= 1;
        if (slot == 0) aux_phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      // D = f32, A / B format of E, A K-major, B MN-major (bit 16), M = 128, N = both column groups or one (tail tile)
      const uint32_t idesc_base = (1u << 4) | (EL::kFmt << 7) | (EL::kFmt << 10) | (1u << 16) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n0 = (t % num_n) * BN;
        const bool h1 = n0 + AT < p.N;
        const uint32_t idesc = idesc_base | ((uint32_t)((h1 ? BN : AT) >> 3) << 17);
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc_mn(sa + kABytes, kMnBox, EL::kMn32);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            EL::mma(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(8 * UMMA_K * k), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(smem_u32(&empty_bar[stage]));
          if (kb == num_kb - 1) umma_commit(smem_u32(&tmem_full[acc]));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int ew = warp - 4;
    const int q = ew & 3;                        // TMEM lane quadrant (== warp % 4)
    const int hsel = ew >> 2;                    // which 128-byte column group of the tile
    const int sw = lane & 7;                     // 128B-swizzle phase of this thread's row ((q*32 + lane) & 7)
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      mbar_wait(smem_u32(&aux_full[acc]), acc_phase);
      tc_fence_after();
      const int rbase = m0 + q * 32;
      const bool live = (n0 + hsel * AT < p.N) && (rbase < p.M);      // warp-uniform (M % 32 == 0)
      if (live) {
        const int sample = rbase / p.rows_per_sample;
        const uint32_t xrow = smem_u32(aux_base + acc * C::kAuxSlotBytes + hsel * kHalfBytes + (q * 32 + lane) * 128);
        const uint32_t grow = xrow + kTensorBytes;                    // MODE 2: same position in the dX tile
#pragma unroll 1
        for (int c0 = 0; c0 < AT; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + hsel * AT + c0), v);
          const int jb = c0 / EL::kPer16;                             // first 16-byte chunk of this 32-column group
          uint4 xr[kCh];
#pragma unroll
          for (int j = 0; j < kCh; ++j) xr[j] = lds128(xrow + (uint32_t)(((jb + j) ^ sw) << 4));
          const int col = n0 + hsel * AT + c0;                        // global channel of v[0]
          const float4* pa = reinterpret_cast<const float4*>(p.ca + (long long)sample * p.cld + col);
          const float4* pb = reinterpret_cast<const float4*>(p.cb + (long long)sample * p.cld + col);
          if constexpr (MODE == 1) {
            float s[32], ss[32];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 a = __ldg(pa + j4), b = __ldg(pb + j4);
              float xs[4];
              unpack4<E>(xr, j4, xs);
              const float as[4] = {a.x, a.y, a.z, a.w};
              const float bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float da = __uint_as_float(v[4 * j4 + e]);
                const float dz = fmaf(as[e], xs[e], bs[e]) > 0.f ? da : 0.f;
                s[4 * j4 + e] = dz;
                ss[4 * j4 + e] = dz * xs[e];
              }
            }
            // transpose-reduce over the warp's 32 rows: after 5 exchange steps lane j holds column (col + j)
#pragma unroll
            for (int step = 0; step < 5; ++step) {
              const int half = 16 >> step;
              const bool upper = (lane >> (4 - step)) & 1;
#pragma unroll
              for (int j = 0; j < half; ++j) {
                const float send_s = upper ? s[j] : s[j + half];
                const float send_q = upper ? ss[j] : ss[j + half];
                const float keep_s = upper ? s[j + half] : s[j];
                const float keep_q = upper ? ss[j + half] : ss[j];
                s[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, 16 >> step);
                ss[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, 16 >> step);
              }
            }
            if (col + lane < p.N) {
              float* tb = p.table + (long long)sample * p.table_ns + 2 * (col + lane);
              atomicAdd(tb, s[0]);
              atomicAdd(tb + 1, ss[0]);
            }
          } else {
            const float4* p2 = reinterpret_cast<const float4*>(p.k2 + (long long)sample * p.cld + col);
            const float4* p3 = reinterpret_cast<const float4*>(p.k3 + (long long)sample * p.cld + col);
            uint4 gr[kCh];
#pragma unroll
            for (int j = 0; j < kCh; ++j) gr[j] = lds128(grow + (uint32_t)(((jb + j) ^ sw) << 4));
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 a = __ldg(pa + j4), b = __ldg(pb + j4), c2 = __ldg(p2 + j4), c3 = __ldg(p3 + j4);
              float xs[4], gs[4];
              unpack4<E>(xr, j4, xs);
              unpack4<E>(gr, j4, gs);
              const float as[4] = {a.x, a.y, a.z, a.w};
              const float bs[4] = {b.x, b.y, b.z, b.w};
              const float k2s[4] = {c2.x, c2.y, c2.z, c2.w};
              const float k3s[4] = {c3.x, c3.y, c3.z, c3.w};
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float da = __uint_as_float(v[4 * j4 + e]);
                const float dz = fmaf(as[e], xs[e], bs[e]) > 0.f ? da : 0.f;
                o[e] = gs[e] + fmaf(as[e], dz, fmaf(k2s[e], xs[e], k3s[e]));
              }
              pack4<E>(gr, j4, o);
            }
#pragma unroll
            for (int j = 0; j < kCh; ++j) sts128(grow + (uint32_t)(((jb + j) ^ sw) << 4), gr[j]);
          }
        }
        if constexpr (MODE == 2) {
          // the warp's [32 rows][128 bytes] box of the dX tile now holds the updated gradient: TMA-store it in place
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const uint32_t box = smem_u32(aux_base + acc * C::kAuxSlotBytes + kTensorBytes + hsel * kHalfBytes + q * 4096);
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(&tmap_gs), "r"(box), "r"(n0 + hsel * AT), "r"(rbase)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be refilled once it has been read
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&tmem_empty[acc]));
        mbar_arrive(smem_u32(&aux_empty[acc]));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (MODE == 2) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
  }
}

// per-(sample, channel) GroupNorm-backward coefficients from the (sum dz, sum dz*x) table:
//   k2 = -rstd^2 * S2[g] / m,   k3 = -rstd * S1[g] / m - k2 * mean      (S1 = sum_c gamma*A, S2 = sum_c gamma*rstd*(B - mean*A))
// and the affine-parameter gradients  dbeta[c] += A,  dgamma[c] += rstd*(B - mean*A).   One block per sample.
__global__ void __launch_bounds__(256)
gn_bwd_coeff_kernel(const float* __restrict__ table, long long table_ns, const float* __restrict__ gamma,
                    const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ k2,
                    float* __restrict__ k3, long long ld, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    int C, int G, int HW) {
  dlb_pdl_wait();
  extern __shared__ float sm[];                 // s1[G], s2[G]
  const int n = blockIdx.x, cpg = C / G;
  for (int g = threadIdx.x; g < 2 * G; g += blockDim.x) sm[g] = 0.f;
  __syncthreads();
  const float* t = table + (long long)n * table_ns;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g = c / cpg;
    const float mu = mean[n * G + g], r = rstd[n * G + g];
    const float A = t[2 * c], B = t[2 * c + 1];
    const float xh = r * (B - mu * A);          // sum dz * xhat
    atomicAdd(&sm[g], gamma[c] * A);
    atomicAdd(&sm[G + g], gamma[c] * xh);
    if (dgamma != nullptr) { atomicAdd(&dbeta[c], A); atomicAdd(&dgamma[c], xh); }
  }
  __syncthreads();
  const float inv_m = 1.f / ((float)cpg * (float)HW);
  for (int c = threadIdx.x; c < ld; c += blockDim.x) {
    float v2 = 0.f, v3 = 0.f;
    if (c < C) {
      const int g = c / cpg;
      const float mu = mean[n * G + g], r = rstd[n * G + g];
      const float qq = r * r * sm[G + g] * inv_m;
      v2 = -qq;
      v3 = -r * sm[g] * inv_m + qq * mu;
    }
    k2[(long long)n * ld + c] = v2;
    k3[(long long)n * ld + c] = v3;
  }
}

template <typename E, int MODE>
int launch_dg(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tx, const CUtensorMap& tgl, const CUtensorMap& tgs,
              const DgParams& p, int sms, cudaStream_t st) {
  using C = DCfg<E, MODE>;
  constexpr int BN = C::BN;
  auto kern = dgrad_gn_kernel<E, MODE>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = num_tiles < sms ? num_tiles : sms;
  dlb_launch(kern, dim3(grid), dim3(kThreadsDG), (size_t)C::kSmemBytes, st, ta, tb, tx, tgl, tgs, p);
  return dlb_post_launch();
}

template <typename E>
int dgrad_gn_impl(int mode, const void* dy, long long lddy, const void* w, long long ldw, const void* x, long long ldx,
                  void* dx, long long lddx, int M, int N, int K, int rows_per_sample, const float* ca, const float* cb,
                  const float* k2, const float* k3, long long cld, float* table, long long table_ns, int sm_limit, cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes, AT = Elt<E>::kAtom, V = 16 / EB;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (mode != 1 && mode != 2) return -2;
  if ((K % V) || (N % V) || (lddy % V) || (ldw % V) || (ldx % V) || ((uintptr_t)dy & 15) || ((uintptr_t)w & 15) || ((uintptr_t)x & 15)) return -3;
  if (rows_per_sample <= 0 || (rows_per_sample % 32) || (M % rows_per_sample) || M < BM) return -4;
  if (!ca || !cb || (cld % 64) || cld < (N + 63) / 64 * 64 || ((uintptr_t)ca & 15) || ((uintptr_t)cb & 15)) return -5;
  if (mode == 1 && !table) return -6;
  if (mode == 2 && (!dx || !k2 || !k3 || (lddx % V) || ((uintptr_t)dx & 15) || ((uintptr_t)k2 & 15) || ((uintptr_t)k3 & 15))) return -6;
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  int sms = sm_count;
  if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;
  CUtensorMap ta, tb, tx, tgl, tgs;
  int rc = make_map(&ta, dy, M, K, lddy, BM, EB);
  if (rc) return rc - 10;
  rc = make_map(&tb, w, K, N, ldw, AT, EB, Elt<E>::kMn32);      // [BK k-rows][128 bytes of n] boxes, MN-major operand
  if (rc) return rc - 20;
  rc = make_map(&tx, x, M, N, ldx, BM, EB);
  if (rc) return rc - 30;
  if (mode == 2) {
    rc = make_map(&tgl, dx, M, N, lddx, BM, EB);
    if (rc) return rc - 40;
    rc = make_map(&tgs, dx, M, N, lddx, 32, EB);
    if (rc) return rc - 50;
  } else {
    tgl = tx; tgs = tx;                         // unused in mode 1
  }
  DgParams p;
  p.M = M; p.N = N; p.K = K; p.rows_per_sample = rows_per_sample;
  p.ca = ca; p.cb = cb; p.k2 = k2; p.k3 = k3; p.cld = cld; p.table = table; p.table_ns = table_ns;
  return mode == 1 ? launch_dg<E, 1>(ta, tb, tx, tgl, tgs, p, sms, st) : launch_dg<E, 2>(ta, tb, tx, tgl, tgs, p, sms, st);
}

}  // namespace

// mode 1: table[samples][table_ns] += per-(sample, channel) (sum dz, sum dz*x),  dz = (dY*W) * [ca*x + cb > 0]
// mode 2: dX += ca*dz + k2*x + k3                                              (in place, element type of the operands)
//   dy [M][K] (row stride lddy), w [K][N] (row stride ldw; the 1x1 conv weight [Cmid][Cin] as stored), x / dx [M][N]
//   (row strides ldx / lddx: channel slices of the block buffers are addressed in place).  dtype: DLB_BF16 or DLB_F32 (TF32 math).
DLB_API int dlb_dgrad_gn_dt(int dtype, int mode, const void* dy, long long lddy, const void* w, long long ldw, const void* x, long long ldx,
                            void* dx, long long lddx, int M, int N, int K, int rows_per_sample, const float* ca, const float* cb,
                            const float* k2, const float* k3, long long cld, float* table, long long table_ns, int sm_limit,
                            void* stream) {
  if (dtype == DLB_F32)
    return dgrad_gn_impl<float>(mode, dy, lddy, w, ldw, x, ldx, dx, lddx, M, N, K, rows_per_sample, ca, cb, k2, k3, cld, table, table_ns,
                                sm_limit, (cudaStream_t)stream);
  return dgrad_gn_impl<__nv_bfloat16>(mode, dy, lddy, w, ldw, x, ldx, dx, lddx, M, N, K, rows_per_sample, ca, cb, k2, k3, cld, table,
                                      table_ns, sm_limit, (cudaStream_t)stream);
}

DLB_API int dlb_dgrad_gn(int mode, const void* dy, long long lddy, const void* w, long long ldw, const void* x, long long ldx,
                         void* dx, long long lddx, int M, int N, int K, int rows_per_sample, const float* ca, const float* cb,
                         const float* k2, const float* k3, long long cld, float* table, long long table_ns, int sm_limit,
                         void* stream) {
  return dlb_dgrad_gn_dt(DLB_BF16, mode, dy, lddy, w, ldw, x, ldx, dx, lddx, M, N, K, rows_per_sample, ca, cb, k2, k3, cld, table,
                         table_ns, sm_limit, stream);
}

DLB_API int dlb_gn_bwd_coeff(const float* table, long long table_ns, const float* gamma, const float* mean, const float* rstd,
                             float* k2, float* k3, long long ld, float* dgamma, float* dbeta, int N, int C, int G, int HW,
                             void* stream) {
  if (N <= 0 || C <= 0) return 0;
  if (G <= 0 || (C % G) || ld < C) return -3;
  dlb_launch(gn_bwd_coeff_kernel, dim3(N), dim3(256), (size_t)(2 * G) * sizeof(float), (cudaStream_t)stream, table, table_ns, gamma,
             mean, rstd, k2, k3, ld, dgamma, dbeta, C, G, HW);
  return dlb_post_launch();
}

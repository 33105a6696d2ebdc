// Persistent warp-specialised GEMM on the 5th-gen tensor cores (sm_100a), bf16 (kind::f16) or fp32 storage with TF32
// math (kind::tf32) -- template parameter E; in bytes both flavours use the same shared-memory layouts (tc_common.cuh):
//
//     D[M, N] (+)= epilogue( prologue(A)[M, K] * B[N, K]^T )          fp32 accumulation in TMEM
//
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a multi-stage shared-memory ring,
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (UMMA 128 x BN x 16) from shared-memory
//     matrix descriptors, completion tracked with tcgen05.commit -> mbarrier,
//   * accumulators live in TMEM, double buffered (2 x BN columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1; epilogue warps read them back with tcgen05.ld (32x32b.x32),
//   * persistent grid: one CTA per SM walks the tile list.
//
// This is the convolution engine of the CNN zoo for 1x1 convolutions (pure GEMMs in NHWC: A = activations
// [pixels, Cin] with an arbitrary row stride so channel slices of the DenseNet concat buffer are read in
// place; B = weights [Cout][Cin], exactly how the flat parameter store keeps them) and for nn.Linear
// (reference Net/Densenet.py:13,29; Net/Resnet.py:35,39,45; Net/Transformer.py linear layers).
//
// Fused variants (template flags):
//   PRO_GN  : A-operand prologue  a[n,k]*x + b[n,k] -> ReLU  (GroupNorm-apply + ReLU of the *input*, i.e. the
//             DenseNet pre-activation order GN->ReLU->conv) performed in shared memory between the TMA
//             arrival and the MMA issue, so the normalised activation never exists in HBM.
//   EPI_STATS: epilogue additionally accumulates per-(sample, out-channel) sum and sum-of-squares of the
//             bf16-rounded output into the GroupNorm statistics table (feeds the NEXT GroupNorm for free).
#include "tc_common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace {

constexpr int BM = 128;            // UMMA_M (cta_group::1)
// BK = Elt<E>::kAtom elements = 128 bytes = one SWIZZLE_128B atom row; UMMA_K = Elt<E>::kUmmaK (32 bytes of K)
constexpr int kNumEpiWarps = 4;
constexpr int kProWarps = 8;       // prologue-transform warps (PRO_GN only): 16 KB per stage must clear in < ~380 cycles

struct GemmParams {
  int M, N, K;
  void* d;            // output (element type E), row-major with row stride ldd (elements)
  long long ldd;
  int accumulate_out; // (reserved)
  // PRO_GN: per-(sample, k) affine coefficients a,b fp32 [num_samples][K] (row n = pixel_row / rows_per_sample)
  const float* pro_a;
  const float* pro_b;
  long long pro_ld;   // row stride (floats) of pro_a / pro_b, a multiple of 64, zero padded
  int rows_per_sample;
  int pro_tma;        // coefficient tiles are staged by TMA with the operands (<= 8 samples per 128-row tile)
  // EPI_STATS: table fp32 [num_samples][table_ns] holding (sum, sumsq) pairs per output channel
  float* stats;
  long long stats_ns;
  // CONV3 (3x3, stride 1, pad 1, NHWC): A tiles are 4-D TMA boxes {64 ch, W, Hb, Nb} of the activation tensor shifted by
  // the filter tap; out-of-image pixels come back as zeros from the TMA unit (that IS the padding).
  int conv_H, conv_W, conv_C;   // input spatial size and channels (K = 9 * conv_C)
  int conv_kchunks;             // ceil(conv_C / 64)
  int conv_sign;                // +1 forward (x[h+dy, w+dx]), -1 dgrad (dy[h-dy, w-dx])
  int conv_halo_rows;           // HALO: (hb + 2) * W pixels per stage tile
};

// HALO (3x3 on maps with >= 128 pixels per image): one stage holds, for one horizontal tap dx and one channel chunk, the
// input rows h0-1 .. h0+R of the tile (R+2 image rows, loaded ONCE) and the weights of the three vertical taps; the three
// A operands of a stage are the same shared-memory tile at row offsets 0, W, 2W (whole 1024-byte swizzle groups), so the
// activations cross L2->SM three times per tile instead of nine.
template <typename E, int BN, bool HALO = false, bool PRO = false> struct Cfg {
  static constexpr int kEB = Elt<E>::kBytes;
  // PRO: per stage, the GroupNorm coefficients a[.], b[.] of the (up to 8) samples the tile touches for this k-block, fetched
  // by TMA with the operands: 2 arrays x 8 samples x BK floats
  static constexpr int kCoefStage = PRO ? 2 * 8 * Elt<E>::kAtom * 4 : 0;
  static constexpr int kABytes = (HALO ? 192 : BM) * 128;          // 128 rows (or up to 192 halo rows) x one 128-byte swizzle row of K
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + (HALO ? 3 : 1) * kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  // epilogue staging: per epilogue warp, kBoxes boxes of [32 rows][128 bytes] = 4 KB each, 128B-swizzled,
  // drained with cp.async.bulk.tensor stores (fully coalesced, tails clipped by the tensor map)
  static constexpr int kBoxes = (BN * kEB + 127) / 128;
  static constexpr int kEpiBytes = kNumEpiWarps * kBoxes * 4096;
  static constexpr int kStagesWanted = BN >= 256 ? 3 : (BN >= 128 ? 5 : 6);
  static constexpr int kStagesFit = (227 * 1024 - kEpiBytes - 1024 - 256) / (kStageBytes + kCoefStage);
  static constexpr int kStages = kStagesWanted < kStagesFit ? kStagesWanted : kStagesFit;
  static constexpr int kCoefBytes = kStages * kCoefStage;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + kCoefBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kStages >= 2, "tile does not fit");
};

// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator (+ idle), 3 = idle (keeps epilogue warps at
// warp ids 4..7 so that (warp_id % 4) selects their TMEM lane quadrant), 4..7 = epilogue, 8..15 = prologue transform.
// B_MN: the B operand is given as [K][N] row-major (N contiguous) instead of [N][K] -- e.g. the untransposed weight
// matrix in a dgrad GEMM dX = dY * W.  Its tiles are loaded as (BN/64) TMA boxes of [64 k-rows][64 n] and consumed
// through an MN-major shared-memory descriptor, so no transposed copy of the weights is ever made.
// CONV3: implicit-GEMM 3x3 convolution (see GemmParams); with B_MN it is the data-gradient (flipped taps, weights
// consumed untransposed through a 3-D tensor map {Cin, 9, Cout}).
template <typename E, int BN, bool PRO_GN, bool EPI_STATS, bool B_MN = false, bool CONV3 = false, bool HALO = false>
__global__ void __launch_bounds__(PRO_GN ? 512 : 256, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_d, const __grid_constant__ CUtensorMap tmap_ca,
               const __grid_constant__ CUtensorMap tmap_cb, const GemmParams p) {
  static_assert(!HALO || CONV3, "HALO is a flavour of the 3x3 implicit GEMM");
  using C = Cfg<E, BN, HALO, PRO_GN>;
  using EL = Elt<E>;
  constexpr int BK = EL::kAtom;
  constexpr int UMMA_K = EL::kUmmaK;
  constexpr int kMnBox = BK * 128;                 // bytes of one MN-major B box: [BK k-rows][128 bytes of n]
  constexpr int kMnBoxes = (BN * EL::kBytes) / 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* epi_base = smem + C::kStages * C::kStageBytes;                  // 1024-byte aligned (stage sizes are)
  uint8_t* coef_base = epi_base + C::kEpiBytes;                           // PRO_GN: [stage][a | b][8 samples][BK] fp32
  uint8_t* bar_base = coef_base + C::kCoefBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);              // TMA bytes landed
  uint64_t* empty_bar = full_bar + C::kStages;                             // MMA done reading the stage
  uint64_t* ready_bar = empty_bar + C::kStages;                            // PRO_GN: transform done
  uint64_t* tmem_full = ready_bar + C::kStages;                            // [2] accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;                                    // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM, num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = CONV3 ? (HALO ? 3 : 9) * p.conv_kchunks : (p.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_d) : "memory");
    if constexpr (PRO_GN) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_ca) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_cb) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
      mbar_init(smem_u32(&ready_bar[s]), kProWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), kNumEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  dlb_pdl_wait();            // everything above (smem carve-up, mbarrier init, TMEM alloc) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if constexpr (HALO) {
            // stage = (horizontal tap dxi, channel chunk kc): rows h0-1 .. h0+hb of ONE image (box {ch, W, hb+2, 1}; rows
            // outside the image are zero-filled by the TMA unit) + the weights of the vertical taps dy = -1, 0, +1
            const int dxi = kb / p.conv_kchunks, kc = kb - dxi * p.conv_kchunks;
            const int hw = p.conv_H * p.conv_W;
            const int img = m0 / hw, h0 = (m0 - img * hw) / p.conv_W;
            mbar_expect_tx(fb, (uint32_t)(p.conv_halo_rows * 128 + 3 * C::kBBytes));
            tma_load_4d(sa, &tmap_a, fb, kc * BK, (dxi - 1) * p.conv_sign, h0 - 1, img);
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              const int tap = dyi * 3 + dxi;
              if constexpr (B_MN) {
#pragma unroll
                for (int gi = 0; gi < kMnBoxes; ++gi)
                  tma_load_3d(sb + dyi * C::kBBytes + gi * kMnBox, &tmap_b, fb, n0 + gi * BK, tap, kc * BK);
              } else {
                tma_load_2d(sb + dyi * C::kBBytes, &tmap_b, fb, tap * p.conv_C + kc * BK, n0);
              }
            }
          } else if constexpr (CONV3) {
            mbar_expect_tx(fb, C::kStageBytes);
            const int tap = kb / p.conv_kchunks, kc = kb - tap * p.conv_kchunks;
            const int dy = (tap / 3 - 1) * p.conv_sign, dx = (tap % 3 - 1) * p.conv_sign;
            const int hw = p.conv_H * p.conv_W;
            const int img = m0 / hw, h0 = (m0 - img * hw) / p.conv_W;
            tma_load_4d(sa, &tmap_a, fb, kc * BK, dx, h0 + dy, img);
            if constexpr (B_MN) {
#pragma unroll
              for (int gi = 0; gi < kMnBoxes; ++gi) tma_load_3d(sb + gi * kMnBox, &tmap_b, fb, n0 + gi * BK, tap, kc * BK);
            } else {
              tma_load_2d(sb, &tmap_b, fb, tap * p.conv_C + kc * BK, n0);
            }
          } else {
            mbar_expect_tx(fb, C::kStageBytes + ((PRO_GN && p.pro_tma) ? C::kCoefStage : 0));
            if constexpr (PRO_GN) {
              if (p.pro_tma) {
                // coefficient rows of the samples this tile touches, columns of this k-block (rows past the last sample: zeros)
                const uint32_t cd = smem_u32(coef_base + stage * C::kCoefStage);
                const int s_first = m0 / p.rows_per_sample;
                tma_load_2d(cd, &tmap_ca, fb, kb * BK, s_first);
                tma_load_2d(cd + C::kCoefStage / 2, &tmap_cb, fb, kb * BK, s_first);
              }
            }
            tma_load_2d(sa, &tmap_a, fb, kb * BK, m0);
            if constexpr (B_MN) {
#pragma unroll
              for (int gi = 0; gi < kMnBoxes; ++gi) tma_load_2d(sb + gi * kMnBox, &tmap_b, fb, n0 + gi * BK, kb * BK);
            } else {
              tma_load_2d(sb, &tmap_b, fb, kb * BK, n0);
            }
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      // instruction descriptor: D=f32, A/B format (bf16 = 1, tf32 = 2) at bits 7 / 10, B major at bit 16, N=BN, M=128
      const uint32_t idesc = (1u << 4) | (EL::kFmt << 7) | (EL::kFmt << 10) | (B_MN ? (1u << 16) : 0u) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(smem_u32(PRO_GN ? &ready_bar[stage] : &full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          if constexpr (HALO) {
#pragma unroll
            for (int dyi = 0; dyi < 3; ++dyi) {
              // rows (h + dy*sign) of the halo tile: start (1 + dy*sign) image rows into it (W*128 bytes = whole swizzle groups)
              const uint32_t a_off = (uint32_t)((1 + (dyi - 1) * p.conv_sign) * p.conv_W * 128);
              const uint64_t adesc = make_smem_desc(sa + a_off);
              const uint64_t bdesc = B_MN ? make_smem_desc_mn(sb + dyi * C::kBBytes, kMnBox, EL::kMn32) : make_smem_desc(sb + dyi * C::kBBytes);
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; ++k)
                EL::mma(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)((B_MN ? 8 * UMMA_K : 2) * k), idesc, (kb | dyi | k) != 0 ? 1u : 0u);
            }
          } else {
          const uint64_t adesc = make_smem_desc(sa), bdesc = B_MN ? make_smem_desc_mn(sb, kMnBox, EL::kMn32) : make_smem_desc(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: advance 32 bytes (one UMMA_K) inside the 128-byte swizzle atom: +2 in 16-byte units;
            // MN-major: advance UMMA_K k-rows of 128 bytes: +8*UMMA_K
            EL::mma(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)((B_MN ? 8 * UMMA_K : 2) * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          }
          umma_commit(smem_u32(&empty_bar[stage]));            // frees the smem stage when these MMAs retire
          if (kb == num_kb - 1) umma_commit(smem_u32(&tmem_full[acc]));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                                     // TMEM lane quadrant owned by this warp
    int acc = 0; uint32_t acc_phase = 0;
    uint8_t* stage_buf = epi_base + q * (C::kBoxes * 4096);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / num_n) * BM, n0 = (t % num_n) * BN;
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      // the previous tile's bulk stores must have finished READING the staging buffers before we overwrite them
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
        // 32 accumulator columns -> 32*kBytes bytes of the output row = kCh 16-byte chunks of staging box `bx`
        constexpr int kCh = 32 * EL::kBytes / 16;                   // 4 (bf16) or 8 (fp32)
        uint32_t packed[kCh * 4];
        if constexpr (EL::kBytes == 2) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            packed[j] = *reinterpret_cast<uint32_t*>(&h);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) packed[j] = v[j];
        }
        const int col = n0 + c0;
        {
          // box `bx`: [32 rows][128 B], 16-byte chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)
          const int byte0 = c0 * EL::kBytes;
          const int bx = byte0 >> 7;
          const uint32_t box = smem_u32(stage_buf + bx * 4096 + lane * 128);
          const int jb = (byte0 & 127) >> 4;
#pragma unroll
          for (int j = 0; j < kCh; ++j)
            sts128(box + (((jb + j) ^ (lane & 7)) << 4), make_uint4(packed[4 * j], packed[4 * j + 1], packed[4 * j + 2], packed[4 * j + 3]));
          if ((((c0 + 32) * EL::kBytes) & 127) == 0 || c0 + 32 >= BN) {
            // a 128-byte-wide box (or the final partial one) is complete: hand it to the TMA store engine
            fence_proxy_async();
            __syncwarp();
            const int ncol = n0 + bx * BK;
            if (lane == 0 && ncol < p.N && m0 + q * 32 < p.M) {
              asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                           ::"l"(&tmap_d), "r"(smem_u32(stage_buf + bx * 4096)), "r"(ncol), "r"(m0 + q * 32)
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          }
        }
        if constexpr (EPI_STATS) {
          // per-(sample, channel) sum / sumsq of the bf16-rounded outputs; rows of one warp belong to one sample
          // (host guarantees rows_per_sample % 32 == 0).  Transpose-reduce: after 5 exchange steps lane j
          // holds the total of column c0 + j over the warp's 32 rows.
          float s[32], ss[32];
          if constexpr (EL::kBytes == 2) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&packed[j]));
              if (!row_ok) { f.x = 0.f; f.y = 0.f; }
              s[2 * j] = f.x; s[2 * j + 1] = f.y;
              ss[2 * j] = f.x * f.x; ss[2 * j + 1] = f.y * f.y;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float f = row_ok ? __uint_as_float(v[j]) : 0.f;
              s[j] = f; ss[j] = f * f;
            }
          }
#pragma unroll
          for (int step = 0; step < 5; ++step) {
            const int half = 16 >> step;                       // registers kept per lane after this step
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
          // lane -> column mapping produced by the exchange pattern above
          int cidx = 0;
#pragma unroll
          for (int step = 0; step < 5; ++step) cidx |= ((lane >> (4 - step)) & 1) << (4 - step);
          const int sample = (m0 + q * 32) / p.rows_per_sample;
          if (m0 + q * 32 < p.M && col + cidx < p.N) {
            float* tb = p.stats + (long long)sample * p.stats_ns + 2 * (col + cidx);
            atomicAdd(tb, s[0]);
            atomicAdd(tb + 1, ss[0]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all output bytes are globally visible
    __syncwarp();
  } else if (PRO_GN && warp >= 8) {
    // ============================ A-operand prologue: GN-apply + ReLU ========================
    // The TMA wrote a 128-row x 64-col bf16 tile with the 128B swizzle: 16-byte chunk j of row r lives at
    // r*128 + ((j ^ (r & 7)) << 4).  Thread (pw, lane) transforms rows pw*32+lane ... in place.
    // thread -> one 16-byte chunk column (8 channels) and every 16th row: the affine coefficients of its
    // channels stay in registers while it walks down the rows (reloaded only when the sample changes)
    const int tp = threadIdx.x - 256;          // 0..255
    const int cc = tp & 7, r_base = tp >> 3;   // r_base 0..31
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int m0 = (t / num_n) * BM;
      // common case: the whole 128-row tile lies inside one sample and inside M -> one coefficient set per
      // k-block, no per-row integer division in the inner loop
      const int s_first = m0 / p.rows_per_sample;
      const bool uniform = (m0 + BM <= p.M) && ((m0 + BM - 1) / p.rows_per_sample == s_first);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
        // issue all shared loads of this thread first (they are independent), then transform and write back:
        // a load->math->store chain per chunk would expose the shared-memory latency four times per stage
        uint4 raw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r_base + 32 * i;
          raw[i] = lds128(sa + r * 128 + ((cc ^ (r & 7)) << 4));
        }
        int cur = -1;
        float ca[8], cb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ca[i] = 0.f; cb[i] = 0.f; }
        const long long coff = (long long)kb * BK + cc * EL::kPer16;
        // coefficients: staged in shared memory by TMA together with the operands (p.pro_tma) -- a global/L2 load here sits on
        // the critical path of EVERY k-block (ncu r2_03: long-scoreboard stalls dominate, DRAM 10 %, tensor pipe 6 %)
        const uint32_t cs = smem_u32(coef_base + stage * C::kCoefStage) + (uint32_t)(cc * 16 * (EL::kPer16 / 4));
        auto coef_from_smem = [&](int local) {
          const uint32_t pa = cs + (uint32_t)(local * BK * 4), pb = pa + C::kCoefStage / 2;
          const uint4 a0 = lds128(pa), b0 = lds128(pb);
          ca[0] = __uint_as_float(a0.x); ca[1] = __uint_as_float(a0.y); ca[2] = __uint_as_float(a0.z); ca[3] = __uint_as_float(a0.w);
          cb[0] = __uint_as_float(b0.x); cb[1] = __uint_as_float(b0.y); cb[2] = __uint_as_float(b0.z); cb[3] = __uint_as_float(b0.w);
          if constexpr (sizeof(E) == 2) {
            const uint4 a1 = lds128(pa + 16), b1 = lds128(pb + 16);
            ca[4] = __uint_as_float(a1.x); ca[5] = __uint_as_float(a1.y); ca[6] = __uint_as_float(a1.z); ca[7] = __uint_as_float(a1.w);
            cb[4] = __uint_as_float(b1.x); cb[5] = __uint_as_float(b1.y); cb[6] = __uint_as_float(b1.z); cb[7] = __uint_as_float(b1.w);
          }
        };
        if (uniform) {
          if (p.pro_tma) coef_from_smem(0);
          else load_coef<E>(p.pro_a + (long long)s_first * p.pro_ld + coff, p.pro_b + (long long)s_first * p.pro_ld + coff, ca, cb);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r_base + 32 * i;
          const int grow = m0 + r;
          const bool live = uniform || grow < p.M;
          const int sample = uniform ? s_first : min(grow, p.M - 1) / p.rows_per_sample;
          if (!uniform && sample != cur) {
            // coefficient rows are padded to a multiple of 64 and zero-filled by the host (a = b = 0 beyond K)
            if (p.pro_tma) coef_from_smem(sample - s_first);
            else load_coef<E>(p.pro_a + (long long)sample * p.pro_ld + coff, p.pro_b + (long long)sample * p.pro_ld + coff, ca, cb);
            cur = sample;
          }
          affine_relu_chunk<E>(raw[i], ca, cb);
          if (!live) raw[i] = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r_base + 32 * i;
          sts128(sa + r * 128 + ((cc ^ (r & 7)) << 4), raw[i]);
        }
        fence_proxy_async();                                     // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();                                            // every lane's writes + fence precede the warp's single arrival
        if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage])); // 8 arrivals per stage instead of 256 (mbarrier arrivals serialise)
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

// =========================================================================================================
// Weight-gradient GEMM:  dW[Co, Ci] += sum_m dY[m, co] * pro(X)[m, ci]        (reduction over pixels m)
//
// Both operands are read straight from their NHWC activations, i.e. they are "MN-major" for the tensor core
// (the reduction index m is the slow dimension): TMA boxes of 64 channels x 64 pixel-rows land in shared
// memory as 128-byte rows with the 128B swizzle, which is exactly the canonical MN-major SWIZZLE_128B UMMA
// layout (8-row K groups 1024 B apart = SBO, 64-channel MN groups one box = 8192 B apart = LBO).
// The X operand can get the same GroupNorm-apply + ReLU prologue as the forward GEMM, so relu(GN(x)) is
// never materialised for the backward pass either.  The pixel range is split across CTAs (split-K); partial
// results are combined with vector fp32 reductions (red.global.add.v4.f32) into a zero-initialised dW.
struct WgradParams {
  int M, Co, Ci;
  float* dw;            // fp32 [Co][ldw]
  long long ldw;
  int rows_per_split;   // multiple of 64
  int num_splits;
  const float* pro_a;
  const float* pro_b;
  long long pro_ld;
  int rows_per_sample;
};

template <typename E, int BN> struct WCfg {
  static constexpr int kEB = Elt<E>::kBytes;
  static constexpr int kBoxBytes = 64 * 128;                       // 64 pixel rows x 128 bytes of channels (64 bf16 / 32 fp32)
  static constexpr int kABoxes = 128 * kEB / 128;                  // 128 output channels
  static constexpr int kBBoxes = BN * kEB / 128;
  static constexpr int kABytes = kABoxes * kBoxBytes;
  static constexpr int kBBytes = kBBoxes * kBoxBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesWanted = BN >= 256 ? 4 : (BN >= 128 ? 5 : 6);
  static constexpr int kStagesFit = (227 * 1024 - 1024 - 256) / kStageBytes;
  static constexpr int kStages = kStagesWanted < kStagesFit ? kStagesWanted : kStagesFit;
  static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 256;
  static_assert(kStages >= 2, "tile does not fit");
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename E, int BN, bool PRO_GN>
__global__ void __launch_bounds__(PRO_GN ? 512 : 256, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x, const WgradParams p) {
  using C = WCfg<E, BN>;
  using EL = Elt<E>;
  constexpr int UMMA_K = EL::kUmmaK;
  constexpr int kAtom = EL::kAtom;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* bar_base = smem + C::kStages * C::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* ready_bar = empty_bar + C::kStages;
  uint64_t* tmem_full = ready_bar + C::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co_tiles = (p.Co + 127) / 128, ci_tiles = (p.Ci + BN - 1) / BN;
  const int num_units = co_tiles * ci_tiles * p.num_splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
      mbar_init(smem_u32(&ready_bar[s]), kProWarps);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full[a]), 1);
      mbar_init(smem_u32(&tmem_empty[a]), kNumEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  dlb_pdl_wait();            // everything above (smem carve-up, mbarrier init, TMEM alloc) overlapped the previous kernel's tail

  // unit u -> (split, co tile, ci tile); splits vary fastest so neighbouring CTAs stream disjoint pixel ranges
  auto decode = [&](int u, int& sp, int& co0, int& ci0) {
    sp = u % p.num_splits;
    const int t = u / p.num_splits;
    ci0 = (t % ci_tiles) * BN;
    co0 = (t / ci_tiles) * 128;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int sp, co0, ci0; decode(u, sp, co0, ci0);
        const int r0 = sp * p.rows_per_split, r1 = min(p.M, r0 + p.rows_per_split);
        for (int r = r0; r < r1; r += 64) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, C::kStageBytes);
#pragma unroll
          for (int gi = 0; gi < C::kABoxes; ++gi) tma_load_2d(sa + gi * C::kBoxBytes, &tmap_dy, fb, co0 + gi * kAtom, r);
#pragma unroll
          for (int gi = 0; gi < C::kBBoxes; ++gi) tma_load_2d(sb + gi * C::kBoxBytes, &tmap_x, fb, ci0 + gi * kAtom, r);
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D=f32, A/B format, BOTH MN-major (bits 15, 16), N=BN, M=128
      const uint32_t idesc = (1u << 4) | (EL::kFmt << 7) | (EL::kFmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int sp, co0, ci0; decode(u, sp, co0, ci0);
        const int r0 = sp * p.rows_per_split, r1 = min(p.M, r0 + p.rows_per_split);
        mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
        bool first = true;
        for (int r = r0; r < r1; r += 64) {
          mbar_wait(smem_u32(PRO_GN ? &ready_bar[stage] : &full_bar[stage]), phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * C::kStageBytes);
          const uint32_t sb = sa + C::kABytes;
          const uint64_t adesc = make_smem_desc_mn(sa, C::kBoxBytes, EL::kMn32), bdesc = make_smem_desc_mn(sb, C::kBoxBytes, EL::kMn32);
#pragma unroll
          for (int k = 0; k < 64 / UMMA_K; ++k) {
            // advance UMMA_K pixel rows of 128 bytes along K: +8*UMMA_K in 16-byte units
            EL::mma(tmem_d, adesc + (uint64_t)(8 * UMMA_K * k), bdesc + (uint64_t)(8 * UMMA_K * k), idesc, (first && k == 0) ? 0u : 1u);
          }
          first = false;
          umma_commit(smem_u32(&empty_bar[stage]));
          if (r + 64 >= r1) umma_commit(smem_u32(&tmem_full[acc]));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int sp, co0, ci0; decode(u, sp, co0, ci0);
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      const int co = co0 + q * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0), v);
        if (co < p.Co) {
          float* dst = p.dw + (long long)co * p.ldw + ci0 + c0;
          if (ci0 + c0 + 32 <= p.Ci && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              red_add_v4(dst + 4 * j, __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
          } else {
            for (int j = 0; j < 32; ++j)
              if (ci0 + c0 + j < p.Ci) atomicAdd(dst + j, __uint_as_float(v[j]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (PRO_GN && warp >= 8) {
    // X tile: kBBoxes boxes of [64 pixel rows][128 bytes of channels].  thread -> one chunk column (16 bytes) of one
    // box and every (256/cols)-th pixel row; coefficients stay in registers down the rows.
    const int tp = threadIdx.x - 256;          // 0..255
    constexpr int kBoxes = C::kBBoxes;
    constexpr int kCols = kBoxes * 8;          // chunk columns in the tile: 8, 16, 32 (or 64 for fp32 BN = 256)
    constexpr int kRowStep = 256 / kCols > 0 ? 256 / kCols : 1;
    static_assert(kCols <= 256, "tile too wide for the transform warps");
    const int ccg = tp % kCols, r_base = tp / kCols;
    const int box = ccg >> 3, cc = ccg & 7;
    int stage = 0; uint32_t phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int sp, co0, ci0; decode(u, sp, co0, ci0);
      const int r0 = sp * p.rows_per_split, r1 = min(p.M, r0 + p.rows_per_split);
      const int cbase = ci0 + box * kAtom + cc * EL::kPer16;
      const bool in_range = cbase < p.pro_ld;            // coefficient rows are padded to a multiple of 64
      for (int rr = r0; rr < r1; rr += 64) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        const uint32_t sb = smem_u32(smem + stage * C::kStageBytes + C::kABytes + box * C::kBoxBytes);
        int cur = -1;
        float ca[8], cb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ca[i] = 0.f; cb[i] = 0.f; }
        const int s_first = rr / p.rows_per_sample;
        const bool uniform = in_range && (rr + 64 <= p.M) && ((rr + 63) / p.rows_per_sample == s_first);
        if (uniform) load_coef<E>(p.pro_a + (long long)s_first * p.pro_ld + cbase, p.pro_b + (long long)s_first * p.pro_ld + cbase, ca, cb);
        constexpr int kIters = 64 / kRowStep;          // chunks per thread per stage
        constexpr int kBatch = kIters < 4 ? kIters : 4;
#pragma unroll
        for (int i0 = 0; i0 < kIters; i0 += kBatch) {
          uint4 raw[kBatch];
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const int r = r_base + kRowStep * (i0 + j);
            raw[j] = lds128(sb + swz_chunk<EL::kMn32>(r, cc));
          }
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const int r = r_base + kRowStep * (i0 + j);
            const int grow = rr + r;
            const bool live = uniform || (grow < p.M && in_range);
            const int sample = uniform ? s_first : min(grow, p.M - 1) / p.rows_per_sample;
            if (!uniform && sample != cur && in_range) {
              load_coef<E>(p.pro_a + (long long)sample * p.pro_ld + cbase, p.pro_b + (long long)sample * p.pro_ld + cbase, ca, cb);
              cur = sample;
            }
            if (live) affine_relu_chunk<E>(raw[j], ca, cb);
            else raw[j] = make_uint4(0u, 0u, 0u, 0u);
          }
#pragma unroll
          for (int j = 0; j < kBatch; ++j) {
            const int r = r_base + kRowStep * (i0 + j);
            sts128(sb + swz_chunk<EL::kMn32>(r, cc), raw[j]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage]));
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::kTmemCols) : "memory");
  }
}

template <typename E, int BN, bool PRO, bool STATS, bool BMN = false, bool CONV3 = false, bool HALO = false>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, int sms, cudaStream_t st,
           const CUtensorMap* tca = nullptr, const CUtensorMap* tcb = nullptr) {
  using C = Cfg<E, BN, HALO, PRO>;
  auto kern = gemm_tc_kernel<E, BN, PRO, STATS, BMN, CONV3, HALO>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  const int grid = num_tiles < sms ? num_tiles : sms;
  dlb_launch(kern, dim3(grid), dim3(PRO ? 512 : 256), (size_t)C::kSmemBytes, st, ta, tb, td, tca ? *tca : ta, tcb ? *tcb : ta, p);
  return dlb_post_launch();
}

template <typename E, int BN>
int dispatch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p, bool pro, bool stats, int sms, cudaStream_t st,
             bool b_mn = false, const CUtensorMap* tca = nullptr, const CUtensorMap* tcb = nullptr) {
  if constexpr (BN * Elt<E>::kBytes >= 128) {
    if (b_mn) return launch<E, BN, false, false, true>(ta, tb, td, p, sms, st);
  }
  if (pro && stats) return launch<E, BN, true, true>(ta, tb, td, p, sms, st, tca, tcb);
  if (pro) return launch<E, BN, true, false>(ta, tb, td, p, sms, st, tca, tcb);
  if (stats) return launch<E, BN, false, true>(ta, tb, td, p, sms, st);
  return launch<E, BN, false, false>(ta, tb, td, p, sms, st);
}

int sm_count_cached() {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  return sm_count;
}
inline int sms_for(int sm_limit) {
  int sms = sm_count_cached();
  if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;
  return sms;
}
inline GemmParams base_params(int M, int N, int K, void* d, long long ldd) {
  GemmParams p;
  p.M = M; p.N = N; p.K = K; p.d = d; p.ldd = ldd; p.accumulate_out = 0;
  p.pro_a = nullptr; p.pro_b = nullptr; p.pro_ld = 0; p.rows_per_sample = M; p.pro_tma = 0; p.stats = nullptr; p.stats_ns = 0;
  p.conv_H = p.conv_W = p.conv_C = 0; p.conv_kchunks = 1; p.conv_sign = 1; p.conv_halo_rows = 0;
  return p;
}

// ---- element-type generic implementations -------------------------------------------------------------------------
template <typename E>
int gemm_bmn_impl(const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N, int K, int sm_limit,
                  cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes, AT = Elt<E>::kAtom, V = 16 / EB;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % V) || (N % V) || (lda % V) || (ldb % V) || (ldd % V) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15) || ((uintptr_t)d & 15)) return -3;
  const int sms = sms_for(sm_limit);
  CUtensorMap ta, tb, td;
  int rc = make_map(&ta, a, M, K, lda, BM, EB);
  if (rc) return rc - 10;
  rc = make_map(&tb, b, K, N, ldb, AT, EB, Elt<E>::kMn32);  // boxes of [BK k-rows][128 bytes of n], MN-major operand
  if (rc) return rc - 20;
  rc = make_map(&td, d, M, N, ldd, 32, EB);
  if (rc) return rc - 30;
  GemmParams p = base_params(M, N, K, d, ldd);
  if constexpr (EB == 2) {
    const int bn = N <= 64 ? 64 : (N <= 128 ? 128 : 256);
    if (bn == 64) return dispatch<E, 64>(ta, tb, td, p, false, false, sms, st, true);
    if (bn == 128) return dispatch<E, 128>(ta, tb, td, p, false, false, sms, st, true);
    return dispatch<E, 256>(ta, tb, td, p, false, false, sms, st, true);
  } else {
    const int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    if (bn == 32) return dispatch<E, 32>(ta, tb, td, p, false, false, sms, st, true);
    if (bn == 64) return dispatch<E, 64>(ta, tb, td, p, false, false, sms, st, true);
    return dispatch<E, 128>(ta, tb, td, p, false, false, sms, st, true);
  }
}

template <typename E>
int gemm_impl(const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N, int K,
              const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, float* stats, long long stats_ns,
              int sm_limit, cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes, V = 16 / EB;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if ((K % V) || (lda % V) || (ldb % V) || ((uintptr_t)a & 15) || ((uintptr_t)b & 15)) return -3;
  if ((ldd % V) || ((uintptr_t)d & 15) || (N % V)) return -6;
  if (stats && (rows_per_sample % 32)) return -4;
  if (pro_a && ((pro_ld % 64) || pro_ld < K || ((uintptr_t)pro_a & 15) || ((uintptr_t)pro_b & 15))) return -5;
  const int sms = sms_for(sm_limit);
  int bn = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (EB == 4 && bn > 128) bn = 128;                      // fp32 tiles: 128 columns = 512 output bytes per row
  CUtensorMap ta, tb, td;
  int rc = make_map(&ta, a, M, K, lda, BM, EB);
  if (rc) return rc;
  rc = make_map(&tb, b, N, K, ldb, bn, EB);
  if (rc) return rc;
  rc = make_map(&td, d, M, N, ldd, 32, EB);                 // output boxes: 32 rows x 128 bytes, 128B swizzle
  if (rc) return rc;
  GemmParams p = base_params(M, N, K, d, ldd);
  p.pro_a = pro_a; p.pro_b = pro_b; p.pro_ld = pro_ld; p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : M;
  p.stats = stats; p.stats_ns = stats_ns;
  const bool pro = pro_a != nullptr, sts = stats != nullptr;
  CUtensorMap tca, tcb;
  if (pro) {
    // coefficient tiles [<= 8 samples][BK] ride the operand pipeline when a 128-row tile touches at most 8 samples
    static int tma_on = -1;
    if (tma_on < 0) { const char* e = getenv("DLB_PRO_TMA"); tma_on = (e && atoi(e) == 0) ? 0 : 1; }
    const int rps = p.rows_per_sample;
    const long long nsamp = ((long long)M + rps - 1) / rps;
    const int per_tile = (rps >= BM) ? ((BM % rps == 0 || rps % BM == 0) ? 1 : 2) : (BM + rps - 1) / rps + ((BM % rps) ? 1 : 0);
    if (tma_on && per_tile <= 8) {
      int r1 = make_map_f32_plain(&tca, pro_a, nsamp, pro_ld, pro_ld, Elt<E>::kAtom, 8);
      int r2 = make_map_f32_plain(&tcb, pro_b, nsamp, pro_ld, pro_ld, Elt<E>::kAtom, 8);
      if (r1 == 0 && r2 == 0) p.pro_tma = 1;
    }
  }
  const CUtensorMap* pca = p.pro_tma ? &tca : nullptr;
  const CUtensorMap* pcb = p.pro_tma ? &tcb : nullptr;
  switch (bn) {
    case 32: return dispatch<E, 32>(ta, tb, td, p, pro, sts, sms, st, false, pca, pcb);
    case 64: return dispatch<E, 64>(ta, tb, td, p, pro, sts, sms, st, false, pca, pcb);
    case 128: return dispatch<E, 128>(ta, tb, td, p, pro, sts, sms, st, false, pca, pcb);
    default:
      if constexpr (EB == 2) return dispatch<E, 256>(ta, tb, td, p, pro, sts, sms, st, false, pca, pcb);
      return -9;
  }
}

template <typename E, int BN, bool PRO>
int launch_wgrad(const CUtensorMap& tdy, const CUtensorMap& tx, const WgradParams& p, int grid, cudaStream_t st) {
  using C = WCfg<E, BN>;
  auto kern = wgrad_tc_kernel<E, BN, PRO>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  dlb_launch(kern, dim3(grid), dim3(PRO ? 512 : 256), (size_t)C::kSmemBytes, st, tdy, tx, p);
  return dlb_post_launch();
}

template <typename E>
int wgrad_impl(const void* dy, long long lddy, const void* x, long long ldx, float* dw, long long ldw, int M, int Co, int Ci,
               const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, int sm_limit, cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes, V = 16 / EB;
  if (M <= 0 || Co <= 0 || Ci <= 0) return 0;
  if ((lddy % V) || (ldx % V) || ((uintptr_t)dy & 15) || ((uintptr_t)x & 15) || (Co % V) || (Ci % V)) return -3;
  if (pro_a && ((pro_ld % 64) || pro_ld < Ci || ((uintptr_t)pro_a & 15) || ((uintptr_t)pro_b & 15))) return -5;
  const int sms = sms_for(sm_limit);
  int bn = Ci <= 64 ? 64 : (Ci <= 128 ? 128 : 256);
  if (EB == 4) bn = Ci <= 32 ? 32 : (Ci <= 64 ? 64 : 128);
  const int co_tiles = (Co + 127) / 128, ci_tiles = (Ci + bn - 1) / bn;
  int splits = (2 * sms + co_tiles * ci_tiles - 1) / (co_tiles * ci_tiles);
  int max_splits = (M + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  int rows = (M + splits - 1) / splits;
  rows = (rows + 63) / 64 * 64;
  splits = (M + rows - 1) / rows;
  CUtensorMap tdy, tx;
  int rc = make_map(&tdy, dy, M, Co, lddy, 64, EB, Elt<E>::kMn32);      // both operands are MN-major
  if (rc) return rc;
  rc = make_map(&tx, x, M, Ci, ldx, 64, EB, Elt<E>::kMn32);
  if (rc) return rc;
  WgradParams p;
  p.M = M; p.Co = Co; p.Ci = Ci; p.dw = dw; p.ldw = ldw; p.rows_per_split = rows; p.num_splits = splits;
  p.pro_a = pro_a; p.pro_b = pro_b; p.pro_ld = pro_ld; p.rows_per_sample = rows_per_sample > 0 ? rows_per_sample : M;
  const int units = co_tiles * ci_tiles * splits;
  const int grid = units < sms ? units : sms;
  const bool pro = pro_a != nullptr;
  if constexpr (EB == 2) {
    if (bn == 64) return pro ? launch_wgrad<E, 64, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 64, false>(tdy, tx, p, grid, st);
    if (bn == 128) return pro ? launch_wgrad<E, 128, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 128, false>(tdy, tx, p, grid, st);
    return pro ? launch_wgrad<E, 256, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 256, false>(tdy, tx, p, grid, st);
  } else {
    if (bn == 32) return pro ? launch_wgrad<E, 32, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 32, false>(tdy, tx, p, grid, st);
    if (bn == 64) return pro ? launch_wgrad<E, 64, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 64, false>(tdy, tx, p, grid, st);
    return pro ? launch_wgrad<E, 128, true>(tdy, tx, p, grid, st) : launch_wgrad<E, 128, false>(tdy, tx, p, grid, st);
  }
}

// box geometry for 128-pixel tiles made of whole image rows
bool conv_tile_geometry(int N, int H, int W, int& hb, int& nb) {
  if (W <= 0 || H <= 0 || 128 % W) return false;
  const int rows = 128 / W;                   // image rows per tile
  if (rows <= H) { if (H % rows) return false; hb = rows; nb = 1; }
  else { if (rows % H) return false; hb = H; nb = rows / H; }
  return true;
}

template <typename E, bool HALO>
int conv3x3_launch(int dgrad, bool stats, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const GemmParams& p,
                   int sms, cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes;
  if (dgrad) {
    if constexpr (EB == 2) {
      if (bn == 64) return launch<E, 64, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
      if (bn == 128) return launch<E, 128, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
      if constexpr (!HALO) return launch<E, 256, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
      return -9;
    } else {
      if (bn == 32) return launch<E, 32, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
      if (bn == 64) return launch<E, 64, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
      return launch<E, 128, false, false, true, true, HALO>(ta, tb, td, p, sms, st);
    }
  }
  if (stats) {
    if (bn == 32) return launch<E, 32, false, true, false, true, HALO>(ta, tb, td, p, sms, st);
    if (bn == 64) return launch<E, 64, false, true, false, true, HALO>(ta, tb, td, p, sms, st);
    if (bn == 128) return launch<E, 128, false, true, false, true, HALO>(ta, tb, td, p, sms, st);
    if constexpr (EB == 2 && !HALO) return launch<E, 256, false, true, false, true, HALO>(ta, tb, td, p, sms, st);
    return -9;
  }
  if (bn == 32) return launch<E, 32, false, false, false, true, HALO>(ta, tb, td, p, sms, st);
  if (bn == 64) return launch<E, 64, false, false, false, true, HALO>(ta, tb, td, p, sms, st);
  if (bn == 128) return launch<E, 128, false, false, false, true, HALO>(ta, tb, td, p, sms, st);
  if constexpr (EB == 2 && !HALO) return launch<E, 256, false, false, false, true, HALO>(ta, tb, td, p, sms, st);
  return -9;
}

template <typename E>
int conv3x3_impl(int dgrad, const void* x, long long ldx, const void* w, void* y, long long ldy, int N, int H, int W, int Ci,
                 int Co, float* stats, long long stats_ns, int sm_limit, cudaStream_t st) {
  constexpr int EB = Elt<E>::kBytes, AT = Elt<E>::kAtom, V = 16 / EB;
  const int Cin = dgrad ? Co : Ci;            // channels of the tensor being convolved
  const int Cout = dgrad ? Ci : Co;           // channels produced
  int hb, nb;
  if (!conv_tile_geometry(N, H, W, hb, nb)) return -7;
  if ((Cin % V) || (Cout % V) || (ldx % V) || (ldy % V) || ((uintptr_t)x & 15) || ((uintptr_t)w & 15) || ((uintptr_t)y & 15)) return -3;
  if (stats && ((H * W) % 32)) return -4;
  const int M = N * H * W;
  const int sms = sms_for(sm_limit);
  int bn;
  if (!dgrad) {
    bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256));
    if (EB == 4 && bn > 128) bn = 128;
  } else {
    bn = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : 256);
    if (EB == 4) bn = Cout <= 32 ? 32 : (Cout <= 64 ? 64 : 128);
  }
  // halo flavour: tiles made of whole rows of ONE image, (hb + 2) * W <= 192 pixels, row offsets in whole swizzle groups
  static int halo_on = -1;
  if (halo_on < 0) { const char* e = getenv("DLB_CONV3_HALO"); halo_on = (e && atoi(e) == 0) ? 0 : 1; }
  const bool halo = halo_on && nb == 1 && (W % 8) == 0 && (hb + 2) * W <= 192 && bn <= 128;
  CUtensorMap ta, tb, td;
  {
    long long dims[4] = {Cin, W, H, N};
    long long strides[4] = {1, ldx, (long long)W * ldx, (long long)H * W * ldx};
    int box[4] = {AT, W, halo ? hb + 2 : hb, nb};
    int rc = make_map_nd(&ta, x, 4, dims, strides, box, EB);
    if (rc) return rc - 10;
  }
  if (!dgrad) {
    int rc = make_map(&tb, w, Cout, 9LL * Cin, 9LL * Cin, bn, EB);          // [Co][9*Ci], K-major
    if (rc) return rc - 20;
  } else {
    long long dims[3] = {Ci, 9, Co};                                     // w[co][tap][ci]: {ci (n, contiguous), tap, co (k rows)}
    long long strides[3] = {1, Ci, 9LL * Ci};
    int box[3] = {AT, 1, AT};
    int rc = make_map_nd(&tb, w, 3, dims, strides, box, EB, Elt<E>::kMn32);   // MN-major B operand
    if (rc) return rc - 20;
  }
  int rc = make_map(&td, y, M, Cout, ldy, 32, EB);
  if (rc) return rc - 30;
  GemmParams p = base_params(M, Cout, 9 * Cin, y, ldy);
  p.rows_per_sample = H * W; p.stats = stats; p.stats_ns = stats_ns;
  p.conv_H = H; p.conv_W = W; p.conv_C = Cin; p.conv_kchunks = (Cin + AT - 1) / AT; p.conv_sign = dgrad ? -1 : 1;
  p.conv_halo_rows = halo ? (hb + 2) * W : 0;
  if (halo) return conv3x3_launch<E, true>(dgrad, stats != nullptr, bn, ta, tb, td, p, sms, st);
  return conv3x3_launch<E, false>(dgrad, stats != nullptr, bn, ta, tb, td, p, sms, st);
}

}  // namespace

// ---- C ABI.  `dtype`: DLB_BF16 (bf16 operands/outputs, kind::f16) or DLB_F32 (fp32 storage, TF32 tensor-core math, kind::tf32).
// The un-suffixed entry points are the bf16 flavour (kept for callers that predate the fp32/TF32 path). --------------------

// D[M,N] = A[M,K] * B[K,N]   with B given row-major [K][N] (row stride ldb): dgrad of a 1x1 conv / linear
// (dX = dY * W with W = [Cout=K][Cin=N]) without a transposed weight copy.
DLB_API int dlb_gemm_tc_bmn_dt(int dtype, const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N,
                               int K, int sm_limit, void* stream) {
  if (dtype == DLB_F32) return gemm_bmn_impl<float>(a, lda, b, ldb, d, ldd, M, N, K, sm_limit, (cudaStream_t)stream);
  return gemm_bmn_impl<__nv_bfloat16>(a, lda, b, ldb, d, ldd, M, N, K, sm_limit, (cudaStream_t)stream);
}
DLB_API int dlb_gemm_tc_bmn(const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N, int K,
                            int sm_limit, void* stream) {
  return dlb_gemm_tc_bmn_dt(DLB_BF16, a, lda, b, ldb, d, ldd, M, N, K, sm_limit, stream);
}

// D[M,N] (row stride ldd) = pro(A[M,K] (row stride lda)) * B[N,K]^T (row stride ldb).
//   pro_a/pro_b (optional): fp32 [M / rows_per_sample][K] affine coefficients -> A' = relu(a*A + b)
//   stats (optional): fp32 table [M / rows_per_sample][stats_ns] accumulating (sum, sumsq) per output column;
//                     must be zeroed by the caller.
// Requirements: K, N, lda, ldb, ldd multiples of the 16-byte vector (8 bf16 / 4 fp32), 16-byte aligned base pointers; with
// stats rows_per_sample % 32 == 0.
DLB_API int dlb_gemm_tc_dt(int dtype, const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N, int K,
                           const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, float* stats, long long stats_ns,
                           int sm_limit, void* stream) {
  if (dtype == DLB_F32)
    return gemm_impl<float>(a, lda, b, ldb, d, ldd, M, N, K, pro_a, pro_b, pro_ld, rows_per_sample, stats, stats_ns, sm_limit, (cudaStream_t)stream);
  return gemm_impl<__nv_bfloat16>(a, lda, b, ldb, d, ldd, M, N, K, pro_a, pro_b, pro_ld, rows_per_sample, stats, stats_ns, sm_limit, (cudaStream_t)stream);
}
DLB_API int dlb_gemm_tc(const void* a, long long lda, const void* b, long long ldb, void* d, long long ldd, int M, int N, int K,
                        const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, float* stats, long long stats_ns,
                        int sm_limit, void* stream) {
  return dlb_gemm_tc_dt(DLB_BF16, a, lda, b, ldb, d, ldd, M, N, K, pro_a, pro_b, pro_ld, rows_per_sample, stats, stats_ns, sm_limit, stream);
}

// dW[Co][ldw] (fp32, zero-initialised by the caller or accumulated into) += dY[M,Co]^T * pro(X[M,Ci]).
// dY / X: row-major with row strides lddy / ldx (elements, multiples of the 16-byte vector).
DLB_API int dlb_wgrad_tc_dt(int dtype, const void* dy, long long lddy, const void* x, long long ldx, float* dw, long long ldw, int M, int Co,
                            int Ci, const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, int sm_limit, void* stream) {
  if (dtype == DLB_F32)
    return wgrad_impl<float>(dy, lddy, x, ldx, dw, ldw, M, Co, Ci, pro_a, pro_b, pro_ld, rows_per_sample, sm_limit, (cudaStream_t)stream);
  return wgrad_impl<__nv_bfloat16>(dy, lddy, x, ldx, dw, ldw, M, Co, Ci, pro_a, pro_b, pro_ld, rows_per_sample, sm_limit, (cudaStream_t)stream);
}
DLB_API int dlb_wgrad_tc(const void* dy, long long lddy, const void* x, long long ldx, float* dw, long long ldw, int M, int Co, int Ci,
                         const float* pro_a, const float* pro_b, long long pro_ld, int rows_per_sample, int sm_limit, void* stream) {
  return dlb_wgrad_tc_dt(DLB_BF16, dy, lddy, x, ldx, dw, ldw, M, Co, Ci, pro_a, pro_b, pro_ld, rows_per_sample, sm_limit, stream);
}

// 3x3 / stride 1 / pad 1 convolution, NHWC:  y[N,H,W,Co] (row stride ldy) = conv(x[N,H,W,Ci] (pixel stride ldx), w[Co][3][3][Ci]).
// dgrad != 0 computes the data gradient instead: x := dY [N,H,W,Co], output := dX [N,H,W,Ci], same weight tensor.
// stats (optional, forward only): per-(sample, out-channel) (sum, sumsq) table, needs (H*W) % 32 == 0; pre-zeroed.
DLB_API int dlb_conv3x3_tc_dt(int dtype, int dgrad, const void* x, long long ldx, const void* w, void* y, long long ldy, int N, int H, int W,
                              int Ci, int Co, float* stats, long long stats_ns, int sm_limit, void* stream) {
  if (dtype == DLB_F32)
    return conv3x3_impl<float>(dgrad, x, ldx, w, y, ldy, N, H, W, Ci, Co, stats, stats_ns, sm_limit, (cudaStream_t)stream);
  return conv3x3_impl<__nv_bfloat16>(dgrad, x, ldx, w, y, ldy, N, H, W, Ci, Co, stats, stats_ns, sm_limit, (cudaStream_t)stream);
}
DLB_API int dlb_conv3x3_tc(int dgrad, const void* x, long long ldx, const void* w, void* y, long long ldy, int N, int H, int W, int Ci,
                           int Co, float* stats, long long stats_ns, int sm_limit, void* stream) {
  return dlb_conv3x3_tc_dt(DLB_BF16, dgrad, x, ldx, w, y, ldy, N, H, W, Ci, Co, stats, stats_ns, sm_limit, stream);
}

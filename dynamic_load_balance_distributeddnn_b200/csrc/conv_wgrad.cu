// Weight gradient of a 3x3 / stride 1 / pad 1 convolution on the 5th-gen tensor cores (sm_100a), NHWC, bf16 or fp32(TF32):
//
//     dW[co][tap][ci] += sum_{pixels p} dY[p, co] * X[p + tap, ci]              tap = (dy, dx) in {-1,0,1}^2
//
// Reference: every `conv2` of the DenseNet bottleneck (Net/Densenet.py:14,19) and the 3x3 convolutions of ResNet
// (Net/Resnet.py:13,37); in round 1 this was the largest remaining vendor-library kernel of the step.
//
// Both operands are consumed straight from their NHWC activations as MN-major tensor-core operands (the reduction index,
// the pixel, is the slow dimension):
//   A = X shifted by the tap: a 4-D TMA box {128 bytes of channels, W, Hb, Nb} = 128 pixels; out-of-image pixels come back
//       as zeros from the TMA unit (that IS the padding).  M = 128 input channels (2 boxes bf16 / 4 boxes fp32).
//   B = dY: a 2-D TMA box [128 pixels][128 bytes of channels] -> N = 64 (bf16) / 32 (fp32) output channels per unit.
// Accumulators: one TMEM region of N columns per tap; a unit handles a group of taps (<= 512 / N of them) over a range of
// pixel tiles (split-K) and ends with vector-free coalesced fp32 reductions (red.global.add) into dW -- which may be the
// parameter's slice of the flat gradient buffer (gradient sink), so no .grad tensor, cast or pack pass exists.
#include "tc_common.cuh"

namespace {

constexpr int kNumEpiWarpsW3 = 4;

struct W3Params {
  int H, W, Ci, Co;
  int hb, nb;               // image rows / images per 128-pixel tile
  int num_tiles;            // pixel tiles (M / 128)
  int tiles_per_split, num_splits;
  int ci_tiles, co_tiles, tap_groups, taps_per_group;
  float* dw;                // fp32 [Co][9][Ci]
};

template <typename E> struct W3Cfg {
  using EL = Elt<E>;
  static constexpr int kBoxBytes = 128 * 128;                 // 128 pixels x 128 bytes of channels
  static constexpr int kABoxes = 128 * EL::kBytes / 128;      // M = 128 input channels
  static constexpr int kABytes = kABoxes * kBoxBytes;
  static constexpr int kN = EL::kAtom;                        // output channels per unit = one 128-byte MN group
  static constexpr int kAStages = EL::kBytes == 2 ? 5 : 2;
  static constexpr int kBStages = 2;
  static constexpr int kSmemBytes = kAStages * kABytes + kBStages * kBoxBytes + 1024 + 256;
  static constexpr int kMaxTaps = (512 / kN) < 9 ? (512 / kN) : 9;
};

template <typename E>
__global__ void __launch_bounds__(256, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const W3Params p) {
  using C = W3Cfg<E>;
  using EL = Elt<E>;
  constexpr int UMMA_K = EL::kUmmaK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_base = smem + C::kAStages * C::kABytes;
  uint8_t* bar_base = b_base + C::kBStages * C::kBoxBytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* a_empty = a_full + C::kAStages;
  uint64_t* b_full = a_empty + C::kAStages;
  uint64_t* b_empty = b_full + C::kBStages;
  uint64_t* acc_full = b_empty + C::kBStages;
  uint64_t* acc_empty = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_units = p.ci_tiles * p.co_tiles * p.tap_groups * p.num_splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_dy) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kAStages; ++s) { mbar_init(smem_u32(&a_full[s]), 1); mbar_init(smem_u32(&a_empty[s]), 1); }
    for (int s = 0; s < C::kBStages; ++s) { mbar_init(smem_u32(&b_full[s]), 1); mbar_init(smem_u32(&b_empty[s]), 1); }
    mbar_init(smem_u32(acc_full), 1);
    mbar_init(smem_u32(acc_empty), kNumEpiWarpsW3);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  dlb_pdl_wait();

  // unit u -> (split, tap group, co tile, ci tile); splits vary fastest so neighbouring CTAs stream disjoint pixel ranges
  auto decode = [&](int u, int& sp, int& tg, int& co0, int& ci0) {
    sp = u % p.num_splits; u /= p.num_splits;
    tg = u % p.tap_groups; u /= p.tap_groups;
    co0 = (u % p.co_tiles) * C::kN;
    ci0 = (u / p.co_tiles) * 128;
  };
  const int hw = p.H * p.W;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int sp, tg, co0, ci0; decode(u, sp, tg, co0, ci0);
        const int t0 = sp * p.tiles_per_split, t1 = min(p.num_tiles, t0 + p.tiles_per_split);
        const int tap0 = tg * p.taps_per_group, tap1 = min(9, tap0 + p.taps_per_group);
        for (int t = t0; t < t1; ++t) {
          const int m0 = t * 128;
          const int img = m0 / hw, h0 = (m0 - img * hw) / p.W;
          mbar_wait(smem_u32(&b_empty[bs]), bph ^ 1);
          const uint32_t fb = smem_u32(&b_full[bs]);
          mbar_expect_tx(fb, C::kBoxBytes);
          tma_load_2d(smem_u32(b_base + bs * C::kBoxBytes), &tmap_dy, fb, co0, m0);
          if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
          for (int tap = tap0; tap < tap1; ++tap) {
            const int dy = tap / 3 - 1, dx = tap % 3 - 1;
            mbar_wait(smem_u32(&a_empty[as]), aph ^ 1);
            const uint32_t fa = smem_u32(&a_full[as]);
            mbar_expect_tx(fa, C::kABytes);
            const uint32_t sa = smem_u32(smem + as * C::kABytes);
#pragma unroll
            for (int gi = 0; gi < C::kABoxes; ++gi)
              tma_load_4d(sa + gi * C::kBoxBytes, &tmap_x, fa, ci0 + gi * EL::kAtom, dx, h0 + dy, img);
            if (++as == C::kAStages) { as = 0; aph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      // D=f32, A/B format, BOTH MN-major (bits 15, 16), N = kN, M = 128
      const uint32_t idesc = (1u << 4) | (EL::kFmt << 7) | (EL::kFmt << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(C::kN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int as = 0; uint32_t aph = 0; int bs = 0; uint32_t bph = 0; uint32_t acc_phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        int sp, tg, co0, ci0; decode(u, sp, tg, co0, ci0);
        const int t0 = sp * p.tiles_per_split, t1 = min(p.num_tiles, t0 + p.tiles_per_split);
        const int tap0 = tg * p.taps_per_group, tap1 = min(9, tap0 + p.taps_per_group);
        mbar_wait(smem_u32(acc_empty), acc_phase ^ 1);           // the previous unit's accumulators have been drained
        tc_fence_after();
        for (int t = t0; t < t1; ++t) {
          mbar_wait(smem_u32(&b_full[bs]), bph);
          tc_fence_after();
          const uint64_t bdesc = make_smem_desc_mn(smem_u32(b_base + bs * C::kBoxBytes), C::kBoxBytes, EL::kMn32);
          for (int tap = tap0; tap < tap1; ++tap) {
            mbar_wait(smem_u32(&a_full[as]), aph);
            tc_fence_after();
            const uint64_t adesc = make_smem_desc_mn(smem_u32(smem + as * C::kABytes), C::kBoxBytes, EL::kMn32);
            const uint32_t tmem_d = tmem_base + (uint32_t)((tap - tap0) * C::kN);
#pragma unroll
            for (int k = 0; k < 128 / UMMA_K; ++k)
              EL::mma(tmem_d, adesc + (uint64_t)(8 * UMMA_K * k), bdesc + (uint64_t)(8 * UMMA_K * k), idesc, (t == t0 && k == 0) ? 0u : 1u);
            umma_commit(smem_u32(&a_empty[as]));
            if (++as == C::kAStages) { as = 0; aph ^= 1; }
          }
          umma_commit(smem_u32(&b_empty[bs]));
          if (++bs == C::kBStages) { bs = 0; bph ^= 1; }
        }
        umma_commit(smem_u32(acc_full));
        acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
      int sp, tg, co0, ci0; decode(u, sp, tg, co0, ci0);
      const int tap0 = tg * p.taps_per_group, tap1 = min(9, tap0 + p.taps_per_group);
      mbar_wait(smem_u32(acc_full), acc_phase);
      tc_fence_after();
      const int ci = ci0 + q * 32 + lane;                        // this thread's accumulator row = input channel
      for (int tap = tap0; tap < tap1; ++tap) {
#pragma unroll 1
        for (int c0 = 0; c0 < C::kN; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((tap - tap0) * C::kN + c0), v);
          if (ci < p.Ci) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = co0 + c0 + j;
              // lanes of the warp write 32 consecutive floats of dW[co][tap][.]: one coalesced reduction per (co, tap)
              if (co < p.Co) atomicAdd(p.dw + ((long long)co * 9 + tap) * p.Ci + ci, __uint_as_float(v[j]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(acc_empty));
      acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool w3_tile_geometry(int H, int W, int& hb, int& nb) {
  if (W <= 0 || H <= 0 || 128 % W) return false;
  const int rows = 128 / W;
  if (rows <= H) { if (H % rows) return false; hb = rows; nb = 1; }
  else { if (rows % H) return false; hb = H; nb = rows / H; }
  return true;
}

template <typename E>
int wgrad3x3_impl(const void* x, long long ldx, const void* dy, long long lddy, float* dw, int N, int H, int W, int Ci, int Co,
                  int sm_limit, cudaStream_t st) {
  using C = W3Cfg<E>;
  constexpr int EB = Elt<E>::kBytes, AT = Elt<E>::kAtom, V = 16 / EB;
  int hb, nb;
  if (!w3_tile_geometry(H, W, hb, nb)) return -7;
  const long long M = (long long)N * H * W;
  // a last partial tile (images past N inside a multi-image box / rows past M of dY) is zero-filled by the TMA unit on BOTH
  // operands, so it contributes nothing: any N is accepted (DBS hands out odd local batches all the time)
  if ((Ci % V) || (ldx % V) || (lddy % V) || ((uintptr_t)x & 15) || ((uintptr_t)dy & 15) || ((uintptr_t)dw & 3)) return -3;
  static int sm_count = 0;
  if (!sm_count) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev); }
  int sms = sm_count;
  if (sm_limit > 0 && sm_limit < sms) sms = sm_limit;
  CUtensorMap tx, tdy;
  {
    long long dims[4] = {Ci, W, H, N};
    long long strides[4] = {1, ldx, (long long)W * ldx, (long long)H * W * ldx};
    int box[4] = {AT, W, hb, nb};
    int rc = make_map_nd(&tx, x, 4, dims, strides, box, EB, Elt<E>::kMn32);
    if (rc) return rc - 10;
  }
  // dY box: [128 pixels][128 bytes of channels]; the tensor map declares only the Co real channels, so for Co narrower than
  // the box (DenseNet growth 32 in bf16) the TMA unit zero-fills the rest instead of reading neighbouring channels
  W3Params p;
  p.H = H; p.W = W; p.Ci = Ci; p.Co = Co; p.hb = hb; p.nb = nb; p.dw = dw;
  p.num_tiles = (int)((M + 127) / 128);
  p.ci_tiles = (Ci + 127) / 128;
  p.co_tiles = (Co + C::kN - 1) / C::kN;
  p.taps_per_group = C::kMaxTaps;
  p.tap_groups = (9 + p.taps_per_group - 1) / p.taps_per_group;
  if (p.tap_groups > 1) p.taps_per_group = (9 + p.tap_groups - 1) / p.tap_groups;      // balance: 5 + 4
  const int base_units = p.ci_tiles * p.co_tiles * p.tap_groups;
  int splits = (sms + base_units - 1) / base_units;
  if (splits > p.num_tiles) splits = p.num_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.num_tiles + splits - 1) / splits;
  p.num_splits = (p.num_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  {
    int rc = make_map(&tdy, dy, M, Co, lddy, 128, EB, Elt<E>::kMn32);
    if (rc) return rc - 20;
  }
  auto kern = wgrad3x3_tc_kernel<E>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const int units = base_units * p.num_splits;
  const int grid = units < sms ? units : sms;
  dlb_launch(kern, dim3(grid), dim3(256), (size_t)C::kSmemBytes, st, tx, tdy, p);
  return dlb_post_launch();
}

}  // namespace

// dW[Co][3][3][Ci] (fp32, zero-initialised by the caller or accumulated into) += wgrad of the 3x3/s1/p1 convolution.
// x: [N,H,W,Ci] with pixel stride ldx; dy: [N,H,W,Co] with pixel stride lddy (both may be channel slices of wider NHWC buffers).
// Requirements: 128 % W == 0 with whole-row tiles (same geometry as dlb_conv3x3_tc), any N, Ci and the strides
// multiples of the 16-byte vector; Co arbitrary (channels beyond Co inside the 128-byte box are masked in the epilogue).
DLB_API int dlb_wgrad3x3_tc_dt(int dtype, const void* x, long long ldx, const void* dy, long long lddy, float* dw, int N, int H, int W,
                               int Ci, int Co, int sm_limit, void* stream) {
  if (dtype == DLB_F32) return wgrad3x3_impl<float>(x, ldx, dy, lddy, dw, N, H, W, Ci, Co, sm_limit, (cudaStream_t)stream);
  return wgrad3x3_impl<__nv_bfloat16>(x, ldx, dy, lddy, dw, N, H, W, Ci, Co, sm_limit, (cudaStream_t)stream);
}

// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05).
//
//   M = 128 output pixels (an 8 x 16 patch of one image), N = COUT, K = 9 taps x CIN.
//
// Operands are fp16 hi/lo planes of the fp32 tensors (x = hi + lo, each scaled by a power of two),
// and every K step issues three MMAs  D += A_lo*B_hi ; D += A_hi*B_lo ; D += A_hi*B_hi  so the
// fp32 accumulator in TMEM carries ~22 significant bits per product (a single
// fp16/tf32 pass misses the 1e-3 parity bar after 20 DDIM steps, the 3-pass split matches fp32).
//
// Data movement: activations live in HBM as NHWC fp16 planes; a 4-D TMA tensor map {C, W, H, B} with
// box {BK, 16, 8, 1} fetches the (dy,dx)-shifted patch for each tap straight into the K-major
// swizzled layout tcgen05 reads, and TMA's out-of-bounds zero fill *is* the conv's zero padding.
// Weights are pre-packed [tap][COUT][CIN] fp16 hi/lo and fetched by a 3-D map with box {BK, COUT, 1}.
//
// Roles (256 threads, persistent over tiles): warp 0 / 3 = TMA producers, warp 1 = MMA issuer (one elected lane each),
// warp2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> regs -> +bias -> HBM, GroupNorm partial sums
// by warp shuffle).  Two TMEM accumulators so tile i's epilogue overlaps tile i+1's MMAs.
//
// Replaces: the nn.Conv2d calls inside ScheduledCNNRefine (reference
// src/model/head/ddim_depth_estimate_res_swin_addHAHI.py:339-359, UpSample_add :321-333).
#pragma once
#include <cuda_fp8.h>

#include "ptx.cuh"

namespace dd {

// FP8-correction operand planes (conv_halo.cuh, F8): with s = the tensor's power-of-two pre-scale,
//   hi = fp16(s v),  a8 = e4m3(s v / 4),  l8 = e4m3((s v - hi) * 512)        (saturating conversions)
// |s v| must stay below 4 * 448 for a8 not to saturate: reported through the status word like an fp16 overflow.
constexpr float kF8ActDiv = 0.25f, kF8LoMul = 512.f, kF8ActMax = 4.f * 448.f;
__device__ __forceinline__ uint16_t e4m3x2(float a, float b) {  // low byte = a
  return static_cast<uint16_t>(__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3));
}

constexpr int TILE_H = 8;
constexpr int TILE_W = 16;
constexpr int TILE_M = TILE_H * TILE_W;  // 128

enum EpiMode : int {
  EPI_F32_STATS = 0,  // y fp32 NHWC + per-(tile, group) sum / sum-of-squares for the consumer GroupNorm
  EPI_SPLIT = 1,      // y -> scaled fp16 hi/lo planes (input of the next conv; no norm in between)
  EPI_F32 = 2         // y fp32 NHWC only
};

struct ConvArgs {
  int B, H, W;
  int tiles_x, tiles_y, num_tiles;
  const float* bias;       // [COUT]
  float acc_scale;         // 1 / (act_scale * weight_scale): undoes the power-of-two operand scaling
  float* y32;              // [B*H*W][COUT]                       (EPI_F32*)
  float* stats_partial;    // [num_tiles][4][2]                   (EPI_F32_STATS)
  __half* out_hi;          // [B*H*W][COUT]                       (EPI_SPLIT)
  __half* out_lo;
  uint8_t* out_a8;         // non-null (EPI_SPLIT): write the e4m3 planes a8 / l8 instead of the fp16 lo plane
  uint8_t* out_l8;
  float split_scale;       // power-of-two scale applied before the fp16 split of the output
  int* status;             // bit0 set if an fp16 operand would overflow
  int fp8_probe;           // timing probe only: issue the two correction products as FP8 MMAs (results are garbage)
  unsigned long long* clk_probe;  // timing probe only (nullable): CTA 0 adds {SM cycles, nanoseconds} of its lifetime
};

template <int CIN, int COUT, int BK>
struct ConvCfg {
  static_assert((CIN % BK == 0 || CIN < BK) && CIN % 16 == 0 && (BK == 16 || BK == 32 || BK == 64), "bad K chunk");
  static_assert(COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "bad N");
  // CIN < BK (the 16-channel latent with BK = 32): the TMA box is wider than the tensor's channel extent, the missing
  // channels arrive as out-of-bounds zeros and only the first CIN / 16 K-steps are issued.  (Tried for the 16->64 layer:
  // slower, 125 vs 88 us — that mainloop is bound by the bytes TMA pulls from L2, and this doubles them.)
  static constexpr int KC = (CIN + BK - 1) / BK;      // channel chunks per tap
  static constexpr int KSTEPS = (CIN < BK ? CIN : BK) / 16;  // MMA K-steps per chunk
  static constexpr int K_ITERS = 9 * KC;              // pipeline stages consumed per tile
  static constexpr int ROW_BYTES = BK * 2;            // one operand row = swizzle span
  static constexpr int A_BYTES = TILE_M * ROW_BYTES;  // one plane
  static constexpr int B_BYTES = COUT * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int SMEM_BUDGET = 196 * 1024;
  static constexpr int STAGES_RAW = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(STAGES >= 2, "stage too large");
  // Epilogue warps.  ncu (profiles/README.md): for Cout = 64 the mainloop is 27 short MMAs per tile and the epilogue ran at
  // one warp per scheduler, i.e. at single-warp instruction latency (tensor pipe 12.5 % active) -> two sets of four warps,
  // each set drains one 32-channel chunk of the tile.
  static constexpr int EPI_SETS = (COUT == 64) ? 2 : 1;
  static constexpr int EPI_WARPS = 4 * EPI_SETS;
  static constexpr int THREADS = 128 + 32 * EPI_WARPS;
  static constexpr int XPOSE_BYTES = EPI_WARPS * 32 * 32 * 4;  // per-epilogue-warp 32x32 fp32 transpose tile (coalesced y stores)
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 512 /*barriers + scratch*/ + XPOSE_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  // Narrow-N layers: back-to-back MMAs into ONE accumulator serialise on its read-modify-write latency (~105 cycles
  // per MMA measured for N = 64 / 16, vs 32-48 cycles of work).  Give each of the three split passes its own TMEM
  // accumulator (three independent chains, summed in the epilogue; the small terms also add up separately).
  static constexpr int NACC = COUT <= 64 ? 3 : 1;
  static constexpr int ACC_COLS = NACC * COUT;  // TMEM columns per accumulator buffer
  static constexpr int TMEM_COLS_RAW = 2 * ACC_COLS;
  static constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : (TMEM_COLS_RAW <= 64 ? 64 : (TMEM_COLS_RAW <= 128 ? 128 : (TMEM_COLS_RAW <= 256 ? 256 : 512)));
  static constexpr int CH = COUT < 32 ? COUT : 32;    // epilogue column chunk
  static constexpr int GROUP_CH = COUT / 4;           // GroupNorm(4, COUT)
};

template <int CIN, int COUT, int BK, int EPI>
__global__ void __launch_bounds__((ConvCfg<CIN, COUT, BK>::THREADS), 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                    const ConvArgs p) {
  using C = ConvCfg<CIN, COUT, BK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tfull_bar = empty_bar + C::STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* red = reinterpret_cast<float*>(tmem_slot + 2);  // [2][4 warps][4 groups][2]  (64 floats)
  float* xpose = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE_BYTES + 512);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 2);  // two producer threads (activations / weights) arrive per stage
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], C::EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto stage_ptr = [&](int s) { return smem + s * C::STAGE_BYTES; };

  if (warp == 0 || warp == 3) {
    // ------------------------------------------------------------------ TMA producers: warp 0 feeds the activation
    // planes, warp 3 the weight planes.  The whole warp walks the loop and one elected lane issues the copies (a lone
    // `lane == 0` thread got every cp.async.bulk.tensor wrapped in an ELECT / BRA.U.ANY loop; see conv_halo.cuh)
    const bool act = (warp == 0);
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx * TILE_W, y0 = ty * TILE_H;
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        for (int kc = 0; kc < C::KC; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = stage_ptr(stage);
          if (leader) {
            if (act) {
              mbar_arrive_expect_tx(&full_bar[stage], 2 * C::A_BYTES);
              tma_load_4d(s, &tmA_hi, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, img);
              tma_load_4d(s + C::A_BYTES, &tmA_lo, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, img);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], 2 * C::B_BYTES);
              tma_load_3d(s + 2 * C::A_BYTES, &tmB_hi, &full_bar[stage], kc * BK, 0, tap);
              tma_load_3d(s + 2 * C::A_BYTES + C::B_BYTES, &tmB_lo, &full_bar[stage], kc * BK, 0, tap);
            }
          }
          __syncwarp();
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp walks the loop, one
    // elected lane issues (uniform-register operands, back-to-back MMAs; see conv_halo.cuh)
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_f16(TILE_M, COUT);
    int stage = 0;
    uint32_t phase = 0;
    uint32_t acc_phase = 0;  // bit b = phase of accumulator buffer b
    int buf = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[buf], ((acc_phase >> buf) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * C::ACC_COLS);
      for (int it = 0; it < C::K_ITERS; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa_hi = smem_u32(stage_ptr(stage));
        const uint32_t sa_lo = sa_hi + C::A_BYTES;
        const uint32_t sb_hi = sa_hi + 2 * C::A_BYTES;
        const uint32_t sb_lo = sb_hi + C::B_BYTES;
        if (leader) {
#pragma unroll
        for (int k = 0; k < C::KSTEPS; ++k) {
          const uint64_t a_hi = umma_smem_desc(sa_hi + k * 32, C::ROW_BYTES);
          const uint64_t a_lo = umma_smem_desc(sa_lo + k * 32, C::ROW_BYTES);
          const uint64_t b_hi = umma_smem_desc(sb_hi + k * 32, C::ROW_BYTES);
          const uint64_t b_lo = umma_smem_desc(sb_lo + k * 32, C::ROW_BYTES);
          const uint32_t first = (it | k) != 0 ? 1u : 0u;
          // neighbours share an operand (B_hi, then A_hi); see conv_halo.cuh
          if constexpr (C::NACC == 3) {
            umma_f16(d_tmem, a_lo, b_hi, idesc, first);
            umma_f16(d_tmem + 2 * COUT, a_hi, b_hi, idesc, first);
            umma_f16(d_tmem + COUT, a_hi, b_lo, idesc, first);
          } else {
            umma_f16(d_tmem, a_lo, b_hi, idesc, first);
            umma_f16(d_tmem, a_hi, b_hi, idesc, 1u);
            umma_f16(d_tmem, a_hi, b_lo, idesc, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (it == C::K_ITERS - 1) umma_commit(&tfull_bar[buf]);
        }  // leader
        __syncwarp();
        if (++stage == C::STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc_phase ^= (1u << buf);
      buf ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes per set)
    const int q = warp & 3;
    const int es = (warp - 4) >> 2;  // epilogue set: which channel chunks of the tile this warp drains
    constexpr int NCH = COUT / C::CH;
    static_assert(C::EPI_SETS == 1 || NCH == C::EPI_SETS, "one chunk per set");
    const int m = q * 32 + lane;
    const int r = m >> 4, c = m & 15;
    uint32_t full_phase = 0;
    int buf = 0;
    int par = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
      const int x = tx * TILE_W + c, y = ty * TILE_H + r;
      const bool valid = (x < p.W) && (y < p.H);
      const size_t pix = (static_cast<size_t>(img) * p.H + y) * p.W + x;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const uint32_t row_off = static_cast<uint32_t>(pix * COUT);  // < 2^32 elements for every tensor of the path
      float* T = xpose + (es * 4 + q) * 1024;

      mbar_wait(&tfull_bar[buf], (full_phase >> buf) & 1u);
      full_phase ^= (1u << buf);
      tc_fence_after();

      float tsum[4] = {0.f, 0.f, 0.f, 0.f}, tsq[4] = {0.f, 0.f, 0.f, 0.f};
      bool overflow = false;
#ifdef DD_PROBES
      if (p.fp8_probe == 2) {  // timing probe (DD_FP8_PROBE=2): no epilogue work at all -> the mainloop's own tile rate
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        buf ^= 1;
        continue;
      }
#endif
#pragma unroll
      for (int cj = 0; cj < NCH / C::EPI_SETS; ++cj) {
        const int ch0 = (cj * C::EPI_SETS + es) * C::CH;  // compile-time when there is one set (es == 0)
        float v[C::CH];
#pragma unroll
        for (int j = 0; j < C::CH; ++j) v[j] = 0.f;
#pragma unroll
        for (int acc = 0; acc < C::NACC; ++acc) {  // lo*hi + hi*lo first, hi*hi last
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(buf * C::ACC_COLS + acc * COUT + ch0);
          if constexpr (C::CH == 32) {
            uint32_t rr[32];
            tmem_ld_32x32(taddr, rr);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(rr[j]);
          } else {
            uint32_t rr[16];
            tmem_ld_32x16(taddr, rr);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __uint_as_float(rr[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < C::CH; ++j) v[j] = fmaf(v[j], p.acc_scale, __ldg(p.bias + ch0 + j));

        if constexpr (EPI == EPI_F32_STATS) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < C::CH; ++j) {
              // compile-time (both loops are fully unrolled); with two sets: the chunk-local group 0 / 1
              const int g = C::EPI_SETS == 1 ? (cj * C::CH + j) / C::GROUP_CH : j / C::GROUP_CH;
              tsum[g] += v[j];
              tsq[g] = fmaf(v[j], v[j], tsq[g]);
            }
          }
        }

        if constexpr ((EPI == EPI_F32_STATS || EPI == EPI_F32) && C::CH == 32) {
          // thread = row after tcgen05.ld; go through the XOR-swizzled tile so that lane = column and every store
          // instruction writes one full 128-byte row segment (row-per-thread stores made the short-K layers
          // epilogue-bound: 32 partial lines per instruction)
#pragma unroll
          for (int j = 0; j < 32; ++j) T[lane * 32 + ((j ^ lane) & 31)] = v[j];
          __syncwarp();
#pragma unroll 8
          for (int rw = 0; rw < 32; ++rw) {
            const uint32_t o = __shfl_sync(0xffffffffu, row_off, rw) + ch0 + lane;
            if ((vmask >> rw) & 1u) p.y32[o] = T[rw * 32 + ((lane ^ rw) & 31)];
          }
          __syncwarp();
        } else if (valid) {
          if constexpr (EPI == EPI_F32_STATS || EPI == EPI_F32) {
            float4* dst = reinterpret_cast<float4*>(p.y32 + pix * COUT + ch0);
#pragma unroll
            for (int j = 0; j < C::CH / 4; ++j)
              dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
            __align__(16) __half hi[C::CH];
            __align__(16) __half lo[C::CH];
#pragma unroll
            for (int j = 0; j < C::CH; ++j) {
              const float s = v[j] * p.split_scale;
              overflow |= (fabsf(s) > 60000.f);
              hi[j] = __float2half_rn(s);
              lo[j] = __float2half_rn(s - __half2float(hi[j]));
            }
            uint4* dh = reinterpret_cast<uint4*>(p.out_hi + pix * COUT + ch0);
            uint4* dl = reinterpret_cast<uint4*>(p.out_lo + pix * COUT + ch0);
#pragma unroll
            for (int j = 0; j < C::CH / 8; ++j) {
              dh[j] = reinterpret_cast<const uint4*>(hi)[j];
              dl[j] = reinterpret_cast<const uint4*>(lo)[j];
            }
          }
        }
      }
      // accumulator fully drained -> hand the TMEM buffer back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);

      if constexpr (EPI == EPI_F32_STATS) {
        // warp tree -> smem -> 8 threads combine the 4 warps in fixed order (deterministic)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float s = tsum[g], s2 = tsq[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0) {
            if constexpr (C::EPI_SETS == 1) {
              red[((par * 4 + q) * 4 + g) * 2 + 0] = s;
              red[((par * 4 + q) * 4 + g) * 2 + 1] = s2;
            } else if (g < 2) {  // set es holds GroupNorm groups 2 es, 2 es + 1 as its local groups 0, 1
              red[((par * 8 + es * 4 + q) * 2 + g) * 2 + 0] = s;
              red[((par * 8 + es * 4 + q) * 2 + g) * 2 + 1] = s2;
            }
          }
        }
        if constexpr (C::EPI_SETS == 1) asm volatile("bar.sync 1, 128;" ::: "memory");  // the epilogue warps only
        else asm volatile("bar.sync 1, 256;" ::: "memory");
        const int e = threadIdx.x - 128;                // 0..127 (.. 255)
        if (e < 8) {
          const int g = e >> 1, which = e & 1;
          float t = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w)
            t += C::EPI_SETS == 1 ? red[((par * 4 + w) * 4 + g) * 2 + which]
                                  : red[((par * 8 + (g >> 1) * 4 + w) * 2 + (g & 1)) * 2 + which];
          p.stats_partial[(static_cast<size_t>(tile) * 4 + g) * 2 + which] = t;
        }
        par ^= 1;  // double-buffered scratch: one barrier per tile is enough
      }
      if constexpr (EPI == EPI_SPLIT) {
        if (overflow) atomicOr(p.status, 1);
      }
      buf ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace dd

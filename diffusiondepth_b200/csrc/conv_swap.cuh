// 3x3 conv for the NARROW-N layers of the denoiser (Cout = 64 or 16: noise_embedding.0, pred.0, pred.3) with the
// operand roles swapped: the WEIGHTS are the UMMA A operand (M = 128 rows = [W_hi(co) for co<64 ; W_lo(co)]) and a
// 16x16 patch of 256 PIXELS is the B operand (N = 256).
//
// Why (measured, profiles/README.md): with pixels as M = 128 and N = Cout <= 64 every tcgen05.mma costs ~105 cycles
// no matter how little work it does (the 4 KB A-operand fetch from shared memory is the floor), and the 3-pass split
// needs 3 of them per K-step per 128 pixels: 315 cycles.  Swapped, one MMA D[128 x 256] += [W_hi;W_lo] * P^T covers
// 256 pixels and BOTH weight planes at the full N = 256 rate; two of them (P = pixel hi plane, pixel lo plane) per
// K-step give all four partial products (hi*hi, lo*hi in accumulator rows 0..63... and hi*lo, lo*lo in rows 64..127):
// 2 x ~165 cycles per 256 pixels = 165 per 128 pixels, and the result is exact to the full 22-bit operands.
//
// TMEM accumulator: lane = weight row, column = pixel.  The rows are ordered so that every 32-lane quarter (= what one
// epilogue warp can read) is self-contained: lanes 0..15 of quarter q hold W_hi of 16 output channels, lanes 16..31 the
// W_lo rows of the SAME channels, so hi + lo is one warp shuffle (no shared-memory exchange between warps, no block
// barrier per chunk — that exchange made the two tiny layers epilogue-bound).  Cout = 64: quarter q owns channels
// 16q..16q+15 (= GroupNorm group q) for all 256 pixels.  Cout = 16: the 32 rows are replicated in all four quarters
// (M = 128 costs the same as M = 32) and quarter q finishes pixels 64q..64q+63.  Stores go through a warp-private
// [32 px][16 ch] transpose so each lane writes a float4.
#pragma once
#include "conv_umma.cuh"

namespace dd {

constexpr int SWAP_TH = 16;  // pixel tile 16 x 16 = 256 = UMMA N
constexpr int SWAP_TW = 16;
constexpr int SWAP_N = 256;

// HALO = true: ROW-HALO REUSE of the pixel operand, as in conv_halo.cuh.  ncu (round 1 / 2): the plain kernel re-fetches the
// 256-pixel patch for each of the 9 taps and sits on the L2 -> SM feed (lts 54 %, 80 B/clk needed for 512 cycles of MMA
// work per 32-channel stage against the ~40 B/clk an SM gets) with the tensor pipe at 70 %.  Here a ring UNIT holds, for
// one 32-channel chunk and one dx, the [18 rows][16 cols] column-shifted strip of both planes (36 KB); tap (dy, dx) is
// that strip advanced dy rows = 16 pixels x 64 B = two 64-byte-swizzle repeats, so the canonical K-major B descriptor
// still addresses it.  Pixel traffic 9 x 256 -> 3 x 288 rows per chunk; weights get their own ring (8 KB per tap).
template <int CIN, int COUT, int BK, bool HALO = false>
struct SwapCfg {
  static_assert(COUT == 64 || COUT == 16, "swap kernel serves the narrow layers");
  static_assert(CIN % BK == 0 && (BK == 16 || BK == 32), "bad K chunk");
  static_assert(!HALO || BK == 32, "halo variant: 64-byte rows");
  static constexpr int KC = CIN / BK;
  static constexpr int K_ITERS = 9 * KC;
  static constexpr int ROW_BYTES = BK * 2;
  static constexpr int W_BYTES = 128 * ROW_BYTES;        // [W_hi ; W_lo] tile (A operand)
  static constexpr int P_BYTES = SWAP_N * ROW_BYTES;     // one pixel plane tile (B operand)
  static constexpr int STAGE_BYTES = W_BYTES + 2 * P_BYTES;
  static constexpr int XCH_BYTES = 4 * 32 * 16 * 4 + 4 * 4 * 2 * 4;  // 4 warps x [32 px][16 ch] transpose + stats scratch [4 warps][4 groups][2]
  static constexpr int STAGES_RAW = (220 * 1024 - XCH_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static_assert(STAGES >= 2, "stage too large");
  // halo variant: strip units + weight slots
  static constexpr int STRIP_BYTES = (SWAP_TH + 2) * SWAP_TW * ROW_BYTES;  // one plane of one dx: 18 x 16 pixel rows
  static constexpr int UNIT_BYTES = 2 * STRIP_BYTES;                       // hi, lo
  static constexpr int UNITS = 4;
  static constexpr int W_SLOTS = 6;
  static_assert(STRIP_BYTES % 1024 == 0 && W_BYTES % 1024 == 0, "operand tiles stay 1024-byte aligned");
  static constexpr int RING_BYTES = HALO ? UNITS * UNIT_BYTES + W_SLOTS * W_BYTES : STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 1024 + 512 + XCH_BYTES;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
  static constexpr int TMEM_COLS = 512;                  // 2 accumulator buffers x 256 pixel columns
  static constexpr int GROUP_CH = COUT / 4;
};

// HALO: tmP_* are strip maps (box {BK, 16, 18, 1}); barrier arrays: full / empty [0 .. UNITS) for the strip units, then
// [8 .. 8 + W_SLOTS) for the weight slots.
template <int CIN, int COUT, int BK, int EPI, bool HALO = false>
__global__ void __launch_bounds__(256, 1)
conv3x3_swap_kernel(const __grid_constant__ CUtensorMap tmP_hi, const __grid_constant__ CUtensorMap tmP_lo,
                    const __grid_constant__ CUtensorMap tmW, const ConvArgs p) {
  using C = SwapCfg<CIN, COUT, BK, HALO>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ctrl = smem + C::RING_BYTES;
  constexpr int NBAR = HALO ? 16 : C::STAGES;  // halo: 8 slots reserved for units + 8 for weight slots
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);
  uint64_t* empty_bar = full_bar + NBAR;
  uint64_t* tfull_bar = empty_bar + NBAR;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* xch = reinterpret_cast<float*>(ctrl + 512);
  static_assert((2 * NBAR + 4) * 8 + 8 <= 512, "barrier block");
  static_assert(!HALO || (C::UNITS <= 8 && C::W_SLOTS <= 8), "barrier slots");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmP_hi);
    tma_prefetch_desc(&tmP_lo);
    tma_prefetch_desc(&tmW);
    for (int s = 0; s < NBAR; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 4);
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

  if (HALO && (warp == 0 || warp == 3)) {
    // ------------------------------------------------------------------ halo variant: warp 0 streams the strip units
    // (chunk, dx), warp 3 the weight tiles (chunk, dx, dy); whole warp walks, one elected lane issues
    const bool strips = (warp == 0);
    const bool leader = elect_one();
    int slot = 0;
    uint32_t phase = 0;
    uint8_t* w_ring = smem + C::UNITS * C::UNIT_BYTES;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx * SWAP_TW, y0 = ty * SWAP_TH;
      for (int kc = 0; kc < C::KC; ++kc) {
        for (int dx = 0; dx < 3; ++dx) {
          if (strips) {
            mbar_wait(&empty_bar[slot], phase ^ 1);
            uint8_t* d = smem + slot * C::UNIT_BYTES;
            if (leader) {
              mbar_arrive_expect_tx(&full_bar[slot], C::UNIT_BYTES);
              tma_load_4d(d, &tmP_hi, &full_bar[slot], kc * BK, x0 + dx - 1, y0 - 1, img);
              tma_load_4d(d + C::STRIP_BYTES, &tmP_lo, &full_bar[slot], kc * BK, x0 + dx - 1, y0 - 1, img);
            }
            __syncwarp();
            if (++slot == C::UNITS) {
              slot = 0;
              phase ^= 1;
            }
          } else {
            for (int dy = 0; dy < 3; ++dy) {
              mbar_wait(&empty_bar[8 + slot], phase ^ 1);
              if (leader) {
                mbar_arrive_expect_tx(&full_bar[8 + slot], C::W_BYTES);
                tma_load_3d(w_ring + slot * C::W_BYTES, &tmW, &full_bar[8 + slot], kc * BK, 0, dy * 3 + dx);
              }
              __syncwarp();
              if (++slot == C::W_SLOTS) {
                slot = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (!HALO && warp == 0) {
    // ------------------------------------------------------------------ TMA producer: whole warp walks the loop, one
    // elected lane issues the three copies of a stage back to back (see conv_halo.cuh)
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx * SWAP_TW, y0 = ty * SWAP_TH;
      for (int tap = 0; tap < 9; ++tap) {
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        for (int kc = 0; kc < C::KC; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = stage_ptr(stage);
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            tma_load_3d(s, &tmW, &full_bar[stage], kc * BK, 0, tap);
            tma_load_4d(s + C::W_BYTES, &tmP_hi, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, img);
            tma_load_4d(s + C::W_BYTES + C::P_BYTES, &tmP_lo, &full_bar[stage], kc * BK, x0 + dx, y0 + dy, img);
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
    constexpr uint32_t idesc = umma_idesc_f16(128, SWAP_N);
    int stage = 0, buf = 0;
    uint32_t phase = 0, acc_phase = 0;
    if constexpr (HALO) {
      int ws = 0;
      uint32_t wphase = 0;
      const uint32_t w_ring = smem_u32(smem + C::UNITS * C::UNIT_BYTES);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[buf], ((acc_phase >> buf) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * SWAP_N);
        for (int kc = 0; kc < C::KC; ++kc) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(&full_bar[stage], phase);
            const uint32_t unit = smem_u32(smem + stage * C::UNIT_BYTES);
            for (int dy = 0; dy < 3; ++dy) {
              mbar_wait(&full_bar[8 + ws], wphase);
              tc_fence_after();
              const uint32_t sw = w_ring + ws * C::W_BYTES;
              const uint32_t sp_hi = unit + dy * SWAP_TW * C::ROW_BYTES;  // 16 pixels down = two swizzle repeats
              const uint32_t sp_lo = sp_hi + C::STRIP_BYTES;
              if (leader) {
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                  const uint64_t wd = umma_smem_desc(sw + k * 32, C::ROW_BYTES);
                  umma_f16(d_tmem, wd, umma_smem_desc(sp_lo + k * 32, C::ROW_BYTES), idesc, (kc | dx | dy | k) != 0 ? 1u : 0u);
                  umma_f16(d_tmem, wd, umma_smem_desc(sp_hi + k * 32, C::ROW_BYTES), idesc, 1u);
                }
                umma_commit(&empty_bar[8 + ws]);
              }
              __syncwarp();
              if (++ws == C::W_SLOTS) {
                ws = 0;
                wphase ^= 1;
              }
            }
            if (leader) {
              umma_commit(&empty_bar[stage]);
              if (kc == C::KC - 1 && dx == 2) umma_commit(&tfull_bar[buf]);
            }
            __syncwarp();
            if (++stage == C::UNITS) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        acc_phase ^= (1u << buf);
        buf ^= 1;
      }
    } else
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[buf], ((acc_phase >> buf) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * SWAP_N);
      for (int it = 0; it < C::K_ITERS; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sw = smem_u32(stage_ptr(stage));
        const uint32_t sp_hi = sw + C::W_BYTES;
        const uint32_t sp_lo = sp_hi + C::P_BYTES;
        if (leader) {
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t wd = umma_smem_desc(sw + k * 32, C::ROW_BYTES);
          const uint64_t ph = umma_smem_desc(sp_hi + k * 32, C::ROW_BYTES);
          const uint64_t pl = umma_smem_desc(sp_lo + k * 32, C::ROW_BYTES);
          umma_f16(d_tmem, wd, pl, idesc, (it | k) != 0 ? 1u : 0u);  // small terms first
          umma_f16(d_tmem, wd, ph, idesc, 1u);
        }
        umma_commit(&empty_bar[stage]);
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
    // ------------------------------------------------------------------ epilogue (see the header: quarter-local hi/lo rows)
    const int q = warp & 3;
    const int cl = lane & 15;                         // channel within this quarter's 16
    const int cb = COUT == 64 ? 16 * q : 0;           // first channel of this warp
    constexpr int CHUNKS = COUT == 64 ? 8 : 2;        // 32-pixel chunks this warp finishes per tile
    const int col0 = COUT == 64 ? 0 : 64 * q;         // first pixel column of this warp
    float* X = xch + q * (32 * 16);
    float* R = xch + 4 * 32 * 16;                     // [4 warps][4 groups][2]
    uint32_t full_phase = 0;
    int buf = 0;
    const float bias = __ldg(p.bias + cb + cl);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
      const int x0 = tx * SWAP_TW, y0 = ty * SWAP_TH;
      mbar_wait(&tfull_bar[buf], (full_phase >> buf) & 1u);
      full_phase ^= (1u << buf);
      tc_fence_after();
      float s1 = 0.f, s2 = 0.f;
      const uint32_t xmask = (x0 + 16 <= p.W) ? 0xFFFFu : ((1u << (p.W - x0)) - 1u);  // valid columns of this tile
#pragma unroll 1
      for (int cc = 0; cc < CHUNKS; ++cc) {
        const int c0 = col0 + cc * 32;                // chunk = two tile rows of 16 pixels
        uint32_t rr[32];
        tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(buf * SWAP_N + c0), rr);
        tmem_ld_wait();
        const int yr = y0 + (c0 >> 4);
        const bool row_ok0 = yr < p.H, row_ok1 = yr + 1 < p.H;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float a = __uint_as_float(rr[j]);
          const float v = fmaf(a + __shfl_down_sync(0xffffffffu, a, 16), p.acc_scale, bias);  // lanes 0..15: hi + lo
          if (lane < 16) {
            X[j * 16 + cl] = v;
            if ((j < 16 ? row_ok0 : row_ok1) && ((xmask >> (j & 15)) & 1u)) {
              s1 += v;
              s2 = fmaf(v, v, s2);
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int px = (lane >> 2) + 8 * i;         // 0..31: tile row (px >> 4), column (px & 15)
          const int y = yr + (px >> 4), x = px & 15;
          if (y < p.H && ((xmask >> x) & 1u)) {
            const float4 v4 = *reinterpret_cast<const float4*>(X + px * 16 + (lane & 3) * 4);
            *reinterpret_cast<float4*>(p.y32 + ((static_cast<size_t>(img) * p.H + y) * p.W + x0 + x) * COUT + cb +
                                       (lane & 3) * 4) = v4;
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);
      if constexpr (EPI == EPI_F32_STATS) {
        if constexpr (COUT == 64) {  // this warp's 16 channels are exactly GroupNorm group q, over the whole tile
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane == 0) {
            p.stats_partial[(static_cast<size_t>(tile) * 4 + q) * 2 + 0] = s1;
            p.stats_partial[(static_cast<size_t>(tile) * 4 + q) * 2 + 1] = s2;
          }
        } else {                     // groups of 4 channels; each warp saw a quarter of the tile's pixels
#pragma unroll
          for (int o = 2; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
          }
          if (lane < 16 && (lane & 3) == 0) {
            R[(q * 4 + (lane >> 2)) * 2 + 0] = s1;
            R[(q * 4 + (lane >> 2)) * 2 + 1] = s2;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (q == 0 && lane < 8) {
            const int g = lane >> 1, which = lane & 1;
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < 4; ++w) t += R[(w * 4 + g) * 2 + which];
            p.stats_partial[(static_cast<size_t>(tile) * 4 + g) * 2 + which] = t;
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");  // R is reused by the next tile
        }
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

// w [COUT][CIN][3][3] fp32 -> fp16 [tap][128][CIN].  Row r = 32q + i: part = i / 16 (0: hi, 1: lo), channel =
// (cout == 64 ? 16q : 0) + i % 16 — every 32-row quarter carries hi and lo of the same 16 channels (cout == 16: the same
// 32 rows in all four quarters).
__global__ void pack_swap_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int cout, int cin,
                                        float scale) {
  const int n = 9 * 128 * cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ci = i % cin, row = (i / cin) % 128, tap = i / (cin * 128);
    const int q = row >> 5, part = (row >> 4) & 1;
    const int co = (cout == 64 ? 16 * q : 0) + (row & 15);
    const float s = w[(static_cast<size_t>(co) * cin + ci) * 9 + tap] * scale;
    const __half h = __float2half_rn(s);
    out[i] = part == 0 ? h : __float2half_rn(s - __half2float(h));
  }
}

}  // namespace dd

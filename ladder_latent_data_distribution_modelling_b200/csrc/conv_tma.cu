// K1/K2, second generation: stride-1 conv2d / dense as implicit GEMM with BOTH operands fed by TMA.
//
// conv_tc.cu gathers fp32 activations through registers (8 producer warps, ~10 B/clk/SM) -- that gather, not the
// tensor pipe, bounds it.  Here activations and gradients live in HBM as bf16 NHWC and the A operand of every
// k-block (one filter tap x 64 channels of a 128-pixel tile) is ONE 4-D tiled tensor-map copy
//     cp.async.bulk.tensor.4d  box {64 ch, bw, bh, bb}  at  {c0, x0 + dx(tap), y0 + dy(tap), b0}
// whose out-of-bounds zero fill IS the TF zero padding and whose SWIZZLE_128B mode writes exactly the K-major
// (fprop / dgrad) or MN-major (wgrad) canonical UMMA layout.  No im2col matrix, no staging through registers.
//
//   warp 0     TMA producer (one elected lane): tensor-map copy of A, bulk copy of the pre-swizzled weight tile
//              image (fprop/dgrad) or tensor-map copies of the dy tile (wgrad); full/empty mbarrier ring
//   warp 1     MMA issuer: tcgen05.mma.cta_group::1.kind::f16, fp32 accumulators double-buffered in TMEM
//   warps 2-5  epilogue: tcgen05.ld -> smem transpose -> coalesced stores, fused bias+activation (fprop),
//              producer-activation derivative / accumulate / space_to_depth (dgrad), split-K red.add (wgrad);
//              output fp32 or bf16
//   persistent grid = min(#tiles, #SMs)
#include "common.cuh"
#include "ladder_sm100.h"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <cstring>

namespace ladder {
namespace tma {
using namespace ladder::tc;

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int NTHREADS = 192;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int EPI_BYTES = 4 * 32 * 36 * 4;
constexpr int BIAS_BYTES = 256 * 4 + 32 * 4;   // bias tile + depth_to_space column offsets of the tile's 8-column groups
constexpr int STAT_BYTES = 4 * 2 * 256 * 4;     // per epilogue warp: running (sum, sum of squares) of up to 256 output columns
__host__ __device__ constexpr int stages_for(int bn) {
  // 227 KB - alignment slack - barriers - epilogue staging - statistics, divided by the stage size
  return (232448 - 1024 - 256 - EPI_BYTES - BIAS_BYTES - STAT_BYTES) / (A_STAGE_BYTES + bn * BK * 2) > 8
             ? 8
             : (232448 - 1024 - 256 - EPI_BYTES - BIAS_BYTES - STAT_BYTES) / (A_STAGE_BYTES + bn * BK * 2);
}

// ---- halo mode (stride-1 3x3 fprop / dgrad on maps with GH % 16 == 0, GW % 8 == 0)
// The plain kernel fetches one 128-pixel x 64-channel A tile PER FILTER TAP: 9 x 16 KB per 64-channel chunk, and at N <= 128
// the L2 -> SM operand stream (cap ~42 B/clk/SM), not the tensor pipe, bounds it (r1g: 1.77 GB per launch of the N = 64 dgrad;
// r2e: conv2d_7 fprop at 13 TB/s of L2 reads).  In halo mode a tile is 16 rows x 8 columns of ONE image and the A operand of a
// chunk is fetched ONCE: a {64 ch, 16 px, 18 rows} box = the tile plus its 1-pixel halo (zero filled outside the image), 36 KB
// at a 16-pixel row pitch.  Tap (kh, kw) is then the SAME buffer read through a shifted UMMA descriptor: start + ((kh * 16 +
// kw) * 128 B, 8-row groups 2048 B apart (one image row of the halo); the SWIZZLE_128B pattern follows the address bits, so the
// shifted view needs no base offset.
constexpr int HALO_TW = 8, HALO_TH = 16, HALO_PITCH = 16, HALO_ROWS = HALO_TH + 2;
constexpr int HALO_BYTES = HALO_ROWS * HALO_PITCH * 128;      // 36 864 B per 64-channel chunk
// One halo buffer feeds 9 taps x 4 MMAs: 1152 / 2304 / 4608 tensor-pipe cycles at N = 64 / 128 / 256, against ~2-3 k cycles of
// loaded TMA latency for the 36 KB box: the narrower the tile, the more halo buffers must be in flight (r2g: with 2 buffers the
// N <= 128 kernels stalled on the halo ring at every chunk boundary and ran 4-8 % SLOWER than the per-tap kernel).
__host__ __device__ constexpr int halo_stages(int bn) { return bn <= 64 ? 4 : (bn <= 128 ? 3 : 2); }
__host__ __device__ constexpr int halo_b_stages(int bn) {
  return (232448 - 1024 - 256 - EPI_BYTES - BIAS_BYTES - STAT_BYTES - halo_stages(bn) * HALO_BYTES) / (bn * BK * 2) > 9
             ? 9         // (2 * stages + 13) mbarriers must fit the 256-byte barrier block
             : (232448 - 1024 - 256 - EPI_BYTES - BIAS_BYTES - STAT_BYTES - halo_stages(bn) * HALO_BYTES) / (bn * BK * 2);
}

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct Args {
  const __nv_bfloat16* wt;   // FPROP/DGRAD: packed weight tile images [n_tile][kb]
  const float* bias;
  const void* aux;           // DGRAD: saved output of the layer that produced x (activation derivative), fp32 or bf16
  void* out;                 // y / dx (fp32 or bf16), dw (fp32)
  int GW, GH, B;             // pixel grid of the GEMM rows (FPROP: y, DGRAD: dx) or of the reduction (WGRAD: y)
  int C;                     // channels of the TMA-gathered tensor
  int KH, KW;
  int off_y, off_x, sign;    // tap (kh, kw) of grid pixel (y, x) reads source pixel (y*stride + off_y + sign*kh, x*stride + off_x + sign*kw)
  int stride;                // FPROP/WGRAD: conv stride (the tensor map traverses the source with this element stride)
  int Ng;                    // GEMM N: Cout (FPROP, WGRAD) or Cin (DGRAD)
  int act, accumulate, out_bf16, aux_bf16;
  int nkb;                   // FPROP/DGRAD: k-blocks per tile = taps * C/64
  int kb_per_split;          // WGRAD: 64-pixel k-blocks per split
  long long total_kb;        // WGRAD: 64-pixel k-blocks overall
  int m_valid;               // WGRAD: rows of dw that exist (= KH*KW*C)
  int m_tiles, n_tiles, splits;
  int perm_r;                // fused depth_to_space (FPROP) / space_to_depth (DGRAD) store, 0 = off
  int halo_ox, halo_oy;      // halo mode: source offset of the halo box origin relative to the tile origin (min over taps)
  int halo_tiles_x, halo_tiles_per_img;
  int halo_base_mode;        // diagnostic: 1 puts the phase kw into the descriptor's base-offset field (wrong on sm_100a), 0 = none
  float* stat;               // FPROP: per-channel (sum, sum of squares) of the output, accumulated in the epilogue (null = off):
  int stat_groups;           //   [2][groups][Ng]; groups = 1 (batch norm: all rows) or B (instance norm: rows of one sample)
  int stat_group_rows;       //   rows per group (a multiple of 128 when groups > 1, so a tile never straddles two samples)
  int os, opy, opx, OHf, OWf; // DGRAD of a strided conv, one output-parity class: row (b, i, j) of the [B, GH, GW] grid is
                             // dx pixel (b, i*os + opy, j*os + opx) of the full [B, OHf, OWf] map (os = 0/1: off)
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

struct Tile {
  long long t;
  int m_tile, n_tile, nkb;
  long long k_lo;            // WGRAD: first 64-pixel k-block of the slice
  bool valid;
};

template <int MODE>
__device__ __forceinline__ void decode_tile(const Args& a, long long t, long long total, Tile& c) {
  c.t = t;
  c.valid = t < total;
  if (!c.valid) { c.nkb = 0; return; }
  c.n_tile = (int)(t % a.n_tiles);
  const long long r = t / a.n_tiles;
  if (MODE == WGRAD) {
    c.m_tile = (int)(r % a.m_tiles);
    const long long split = r / a.m_tiles;
    c.k_lo = split * a.kb_per_split;
    const long long k_hi = min(a.total_kb, c.k_lo + a.kb_per_split);
    c.nkb = (int)(k_hi - c.k_lo);
  } else {
    c.m_tile = (int)r;
    c.k_lo = 0;
    c.nkb = a.nkb;
  }
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t h) { return __uint_as_float(h << 16); }

// Ragged / scalar tail of the epilogue (N not a multiple of 4, depth_to_space with C' % 4 != 0): kept out of line so the
// hot store loop stays small.
template <int MODE, bool OUT16>
__device__ __noinline__ void slow_store(const Args& a, float4 q, long long m, int col, long long o, long long arow, float slope,
                                        bool is_tanh) {
  float* outf = reinterpret_cast<float*>(a.out);
  __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(a.out);
  const float e4[4] = {q.x, q.y, q.z, q.w};
  const int Ng = a.Ng;
  for (int t = 0; t < 4; ++t) {
    if (col + t >= Ng) break;
    float e = e4[t];
    long long oo = o + t;                      // o: offset of column `col` of this row (row base + column offset)
    if (MODE == FPROP && a.perm_r > 0 && t > 0) {      // depth_to_space: a 4-column group may straddle two (i, j) blocks
      const int r = a.perm_r, Cp = Ng / (r * r), ij0 = col / Cp, ij = (col + t) / Cp;
      if (ij != ij0) {
        const long long off0 = ((long long)(ij0 / r) * a.GW * r + ij0 % r) * Cp + (col - ij0 * Cp);
        oo = o - off0 + ((long long)(ij / r) * a.GW * r + ij % r) * Cp + (col + t - ij * Cp);
      }
    }
    if (MODE == DGRAD) {
      if (a.aux != nullptr) {
        const long long ai = arow + col + t;
        const float ax = a.aux_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.aux)[ai])
                                    : reinterpret_cast<const float*>(a.aux)[ai];
        e *= is_tanh ? 1.f - ax * ax : (ax > 0.f ? 1.f : slope);
      }
      if (a.accumulate) e += OUT16 ? __bfloat162float(outh[oo]) : outf[oo];
    }
    if (MODE == WGRAD) atomicAdd(outf + oo, e);
    else if (OUT16) outh[oo] = __float2bfloat16_rn(e);
    else outf[oo] = e;
  }
}

// add one warp's running column sums to the global statistics and clear them (lane i owns columns i, i + 32, ...)
template <int BN>
__device__ __forceinline__ void stat_flush(const Args& a, float* wstat, int n_tile, int group, int lane) {
  const long long base = (long long)group * a.Ng, sq = (long long)a.stat_groups * a.Ng;
  for (int i = lane; i < BN; i += 32) {
    const int col = n_tile * BN + i;
    if (col < a.Ng) {
      atomicAdd(a.stat + base + col, wstat[i]);
      atomicAdd(a.stat + sq + base + col, wstat[256 + i]);
    }
    wstat[i] = 0.f;
    wstat[256 + i] = 0.f;
  }
}

// K-major SWIZZLE_128B descriptor of a halo-buffer view: 8-row groups 2048 B apart, swizzle phase in the base-offset field
__device__ __forceinline__ uint64_t make_desc_halo(uint32_t saddr, uint32_t base_offset) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((HALO_PITCH * 128) >> 4) << 32) | (1ull << 46) |
         ((uint64_t)(base_offset & 7) << 49) | (2ull << 61);
}

template <int MODE, int BN, bool OUT16, bool HALO = false>
__global__ void __launch_bounds__(NTHREADS, 1)
tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Args a) {
  // plain mode: STAGES x (A tile + B tile).  halo mode: HALO_STAGES halo buffers (ring `h`) + STAGES B tiles (ring `s`)
  constexpr int STAGES = HALO ? halo_b_stages(BN) : stages_for(BN);
  constexpr int HALO_STAGES = halo_stages(BN);
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int A_REGION = HALO ? HALO_STAGES * HALO_BYTES : STAGES * A_STAGE_BYTES;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + A_REGION;
  const uint32_t bars = sB + STAGES * B_STAGE_BYTES;            // full[S], empty[S], tfull[2], tempty[2], slot, hfull[2], hempty[2]
  const uint32_t bar_full = bars, bar_empty = bars + STAGES * 8, bar_tfull = bars + 2 * STAGES * 8,
                 bar_tempty = bars + (2 * STAGES + 2) * 8;
  const uint32_t slot = bars + (2 * STAGES + 4) * 8;
  const uint32_t bar_hfull = bars + (2 * STAGES + 5) * 8, bar_hempty = bars + (2 * STAGES + 9) * 8;   // <= 31 x 8 B <= 256
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (slot - base));
  const uint32_t stage_off = bars + 256 - base;                 // epilogue staging (4 x 32 x 36 floats)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long pixels = (long long)a.B * a.GH * a.GW;
  const long long Mg = MODE == WGRAD ? (long long)a.m_valid : pixels;
  const int Ng = a.Ng;
  const long long total = (long long)a.m_tiles * a.n_tiles * a.splits;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar_hfull + i * 8, 1);
      mbar_init(bar_hempty + i * 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  // Warps 0 and 1 run their loops CONVERGED (all 32 lanes wait on the barriers and keep the counters) and one elected lane
  // issues the TMA / tcgen05 instructions.  r2w: with the loops inside `if (lane == 0)` ptxas cannot prove the operands
  // warp-uniform, moves every descriptor through R2UR and wraps each UTCHMMA / UTMALDG / UBLKCP / UTCBAR in an
  // ELECT .. BRA.U.ANY waterfall loop: ~16 dependent SASS instructions (+ a division in the producer) per MMA, 95-135 clk
  // of issue time against 32 / 64 clk of tensor time at N = 64 / 128 -- the issue warps, not the tensor pipe, set the pace of
  // every N <= 128 layer.  Converged, the operands live in uniform registers and the 4 MMAs of a k-block issue back to back.
  if (warp == 0) {
    // ===================================================== TMA producer
    const bool leader = elect_one();
    if (leader) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
      if (MODE == WGRAD) asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    __syncwarp();
    Tile T;
    decode_tile<MODE>(a, blockIdx.x, total, T);
    // Ring protocol (plain mode): EVERY TILE STARTS AT SLOT 0 and walks the slots in order, so the slot index is a compile-time
    // constant of the unrolled loops below (barrier and tile addresses become base + immediate; the ring position no longer
    // travels through R2UR every k-block); each slot keeps its own barrier parity bit.  Slots skipped at the end of a tile's
    // last pass are simply not used that round.  Halo mode keeps running counters (its weight ring is nested in the halo ring).
    uint32_t empty_par = (1u << STAGES) - 1u;    // bit s: parity of slot s's "empty" barrier to wait for
    uint32_t s = 0, ph = 1;                      // halo mode: weight-ring slot and parity
    uint32_t h = 0, hph = 1;                     // halo ring
    (void)empty_par; (void)s; (void)ph; (void)h; (void)hph;
    while (T.valid) {
      if constexpr (HALO) {
        // tile = 16 rows x 8 columns of image b; per 64-channel chunk ONE halo box, then the 9 weight tiles
        const int taps = a.KH * a.KW;
        const int b0 = T.m_tile / a.halo_tiles_per_img, rem = T.m_tile - b0 * a.halo_tiles_per_img;
        const int ty = rem / a.halo_tiles_x, tx = rem - ty * a.halo_tiles_x;
        const int hx = tx * HALO_TW + a.halo_ox, hy = ty * HALO_TH + a.halo_oy;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wt) + (size_t)T.n_tile * T.nkb * B_STAGE_BYTES;
        for (int c0 = 0; c0 < a.C; c0 += BK) {
          mbar_wait(bar_hempty + h * 8, hph);
          if (leader) {
            mbar_arrive_expect_tx(bar_hfull + h * 8, HALO_BYTES);
            tma_load_4d(sA + h * HALO_BYTES, &mapA, c0, hx, hy, b0, bar_hfull + h * 8);
          }
          __syncwarp();
          if (++h == HALO_STAGES) { h = 0; hph ^= 1; }
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(bar_empty + s * 8, ph);
            if (leader) {
              mbar_arrive_expect_tx(bar_full + s * 8, B_STAGE_BYTES);
              tma_bulk_g2s(sB + s * B_STAGE_BYTES, wsrc, B_STAGE_BYTES, bar_full + s * 8);
            }
            __syncwarp();
            wsrc += B_STAGE_BYTES;
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      } else if (MODE != WGRAD) {
        const unsigned m0 = (unsigned)T.m_tile * BM;
        const int x0 = (int)(m0 % (unsigned)a.GW);
        const unsigned r = m0 / (unsigned)a.GW;
        const int y0 = (int)(r % (unsigned)a.GH), b0 = (int)(r / (unsigned)a.GH);
        const int xb = x0 * a.stride + a.off_x, yb = y0 * a.stride + a.off_y;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wt) + (size_t)T.n_tile * T.nkb * B_STAGE_BYTES;
        int kh = 0, kw = 0, c0 = 0;              // K order: 64-channel chunk major, tap minor
        for (int kb0 = 0; kb0 < T.nkb; kb0 += STAGES) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) {     // slot index is a compile-time constant: barrier / tile addresses are base + immediate
            if (kb0 + s < T.nkb) {
              mbar_wait(bar_empty + s * 8, (empty_par >> s) & 1u);
              empty_par ^= 1u << s;
              if (leader) {
                mbar_arrive_expect_tx(bar_full + s * 8, A_STAGE_BYTES + B_STAGE_BYTES);
                tma_load_4d(sA + s * A_STAGE_BYTES, &mapA, c0, xb + a.sign * kw, yb + a.sign * kh, b0, bar_full + s * 8);
                tma_bulk_g2s(sB + s * B_STAGE_BYTES, wsrc, B_STAGE_BYTES, bar_full + s * 8);
              }
              __syncwarp();
              wsrc += B_STAGE_BYTES;
              if (++kw == a.KW) { kw = 0; if (++kh == a.KH) { kh = 0; c0 += BK; } }
            }
          }
        }
      } else {
        // rows of dw: patch entries (tap, ci); a 128-row tile = two 64-channel blocks (possibly of different taps)
        int oy[2], ox[2], cc[2];
        bool on[2];
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const int kd0 = T.m_tile * BM + mb * 64;
          on[mb] = kd0 < a.m_valid;
          const int tp = kd0 / a.C;
          cc[mb] = kd0 - tp * a.C;
          const int kh = tp / a.KW, kw = tp - kh * a.KW;
          oy[mb] = a.off_y + a.sign * kh;
          ox[mb] = a.off_x + a.sign * kw;
        }
        const uint32_t bytes = (uint32_t)((on[0] ? 8192 : 0) + (on[1] ? 8192 : 0) + B_STAGE_BYTES);
        // pixel position of the slice's first 64-pixel k-block, then stepped by 64 pixels per k-block without divisions
        const unsigned p0 = (unsigned)T.k_lo * BK;
        int x0 = (int)(p0 % (unsigned)a.GW);
        const unsigned r = p0 / (unsigned)a.GW;
        int y0 = (int)(r % (unsigned)a.GH), b0 = (int)(r / (unsigned)a.GH);
        const int step_y = BK / a.GW, step_x = BK - step_y * a.GW;
        for (int kb0 = 0; kb0 < T.nkb; kb0 += STAGES) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) {
            if (kb0 + s < T.nkb) {
              mbar_wait(bar_empty + s * 8, (empty_par >> s) & 1u);
              empty_par ^= 1u << s;
              if (leader) {
                mbar_arrive_expect_tx(bar_full + s * 8, bytes);
#pragma unroll
                for (int mb = 0; mb < 2; ++mb)
                  if (on[mb])
                    tma_load_4d(sA + s * A_STAGE_BYTES + mb * 8192, &mapA, cc[mb], x0 * a.stride + ox[mb], y0 * a.stride + oy[mb], b0, bar_full + s * 8);
#pragma unroll
                for (int nb = 0; nb < BN / 64; ++nb)
                  tma_load_4d(sB + s * B_STAGE_BYTES + nb * 8192, &mapB, T.n_tile * BN + nb * 64, x0, y0, b0, bar_full + s * 8);
              }
              __syncwarp();
              x0 += step_x;
              y0 += step_y;
              if (x0 >= a.GW) { x0 -= a.GW; ++y0; }
              while (y0 >= a.GH) { y0 -= a.GH; ++b0; }
            }
          }
        }
      }
      decode_tile<MODE>(a, T.t + gridDim.x, total, T);
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (one elected lane, converged warp)
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(BM, BN, MODE == WGRAD);
    Tile T;
    decode_tile<MODE>(a, blockIdx.x, total, T);
    uint32_t full_par = 0;                       // plain mode: bit s = parity of slot s's "full" barrier (every tile starts at slot 0)
    uint32_t s = 0, ph = 0;                      // halo mode: weight-ring slot and the parity of its "full" barrier
    uint32_t h = 0, hph = 0;
    (void)full_par; (void)s; (void)ph; (void)h; (void)hph;
    unsigned j = 0;
    while (T.valid) {
      const uint32_t acc = j & 1;
      mbar_wait(bar_tempty + acc * 8, ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      if constexpr (HALO) {
        int kb = 0;
        for (int c0 = 0; c0 < a.C; c0 += BK) {
          mbar_wait(bar_hfull + h * 8, hph);
          tc_fence_after();
          for (int kh = 0; kh < a.KH; ++kh) {
            for (int kw = 0; kw < a.KW; ++kw, ++kb) {
              mbar_wait(bar_full + s * 8, ph);
              tc_fence_after();
              // halo coordinates of this tap's view: (off + sign * k) - origin
              const int vy = a.off_y + a.sign * kh - a.halo_oy, vx = a.off_x + a.sign * kw - a.halo_ox;
              const uint32_t tA = sA + h * HALO_BYTES + (uint32_t)(vy * HALO_PITCH + vx) * 128u, tB = sB + s * B_STAGE_BYTES;
              const uint32_t bo = a.halo_base_mode ? (uint32_t)vx : 0u;
              if (leader) {
                const uint64_t dA = make_desc_halo(tA, bo), dB = make_desc(tB);
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, idesc, (kb | k) != 0);
                umma_commit(bar_empty + s * 8);
              }
              __syncwarp();
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
          }
          if (leader) umma_commit(bar_hempty + h * 8);
          __syncwarp();
          if (++h == HALO_STAGES) { h = 0; hph ^= 1; }
        }
      } else {
        for (int kb0 = 0; kb0 < T.nkb; kb0 += STAGES) {
#pragma unroll
          for (int s = 0; s < STAGES; ++s) {     // compile-time slot (see the producer): descriptors are base + immediate
            if (kb0 + s < T.nkb) {
              mbar_wait(bar_full + s * 8, (full_par >> s) & 1u);
              full_par ^= 1u << s;
              tc_fence_after();
              if (leader) {
                // the 4 K steps of a k-block advance the 14-bit start-address field by 32 B (K-major) / 2048 B (MN-major): the
                // tiles are 1024-byte aligned inside a < 256 KB window, so the field never carries into its neighbours
                const uint32_t first = s == 0 ? (uint32_t)(kb0 != 0) : 1u;
                if (MODE == WGRAD) {
                  const uint64_t dA = make_desc_mn(sA + s * A_STAGE_BYTES, 8192), dB = make_desc_mn(sB + s * B_STAGE_BYTES, 8192);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, dA + 128 * k, dB + 128 * k, idesc, k == 0 ? first : 1u);
                } else {
                  const uint64_t dA = make_desc(sA + s * A_STAGE_BYTES), dB = make_desc(sB + s * B_STAGE_BYTES);
#pragma unroll
                  for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, idesc, k == 0 ? first : 1u);
                }
                umma_commit(bar_empty + s * 8);
              }
              __syncwarp();
            }
          }
        }
      }
      if (leader) umma_commit(bar_tfull + acc * 8);
      __syncwarp();
      ++j;
      decode_tile<MODE>(a, T.t + gridDim.x, total, T);
    }
  } else {
    // ===================================================== epilogue
    // tcgen05.ld hands every thread one accumulator ROW.  FPROP applies bias + activation right there (the bias tile
    // sits in shared memory), then each warp transposes its 32x32 chunk through a private padded smem tile so that 8
    // lanes cover 128 contiguous bytes of one output row.  The store loop is kept small on purpose (slope-form
    // activations, ragged / permuted-scalar cases in a __noinline__ slow path): the fully unrolled first version of
    // this epilogue was instruction-fetch bound (ncu: stall_no_inst on every store-side instruction).
    const int quad = warp & 3;                     // tcgen05.ld: warp w may touch TMEM lanes 32*(w%4)..+31
    float* stage = reinterpret_cast<float*>(smem + stage_off) + quad * (32 * 36);
    float* sbias = reinterpret_cast<float*>(smem + stage_off + EPI_BYTES);
    int* scol = reinterpret_cast<int*>(sbias + 256);   // depth_to_space: element offset of column group (n0 + 8 i) inside a row's block
    float* wstat = reinterpret_cast<float*>(smem + stage_off + EPI_BYTES + BIAS_BYTES) + quad * (2 * 256);   // this warp's sums
    const bool do_stat = MODE == FPROP && a.stat != nullptr;
    if (do_stat)
      for (int i = lane; i < 2 * 256; i += 32) wstat[i] = 0.f;
    int stat_tile = -1, stat_group = 0;            // (n_tile, group) the running sums belong to
    const int etid = tid - 64;                     // 0..127 among the epilogue warps
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    const float slope = a.act == ACT_LEAKY ? 0.2f : (a.act == ACT_RELU ? 0.f : 1.f);
    const bool is_tanh = a.act == ACT_TANH;
    float* outf = reinterpret_cast<float*>(a.out);
    __nv_bfloat16* outh = reinterpret_cast<__nv_bfloat16*>(a.out);
    const float* auxf = reinterpret_cast<const float*>(a.aux);
    const __nv_bfloat16* auxh = reinterpret_cast<const __nv_bfloat16*>(a.aux);
    // Row addressing of the permuted stores (depth_to_space / space_to_depth / parity-class scatter): the (b, y, x)
    // decomposition is done ONCE per tile, one row per lane, with 32-bit divisions; the store loop fetches a row's offset
    // with a warp shuffle.  Calling d2s_dest / s2d_dest per (row, column chunk) -- four 64-bit divisions each -- made the
    // epilogue, not the MMA mainloop, the critical path of every decoder layer (4x on the 16x16 conv).
    const bool need_bhw = HALO || (MODE != WGRAD && (a.perm_r > 0 || (MODE == DGRAD && a.os > 1)));
    const int pr = a.perm_r > 0 ? a.perm_r : 1;
    const int rsh = (pr & (pr - 1)) == 0 ? __ffs(pr) - 1 : -1;
    const int GHr = a.GH / pr, GWr = a.GW / pr, Cp_d2s = Ng / (pr * pr);
    Tile T;
    decode_tile<MODE>(a, blockIdx.x, total, T);
    unsigned j = 0;
    int bias_tile = -1;
    while (T.valid) {
      const uint32_t acc = j & 1;
      const int n0 = T.n_tile * BN;
      if (MODE == FPROP && bias_tile != T.n_tile) {          // (re)load the bias tile; uniform across the 4 warps
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = etid; i < BN; i += 128) sbias[i] = (a.bias != nullptr && n0 + i < Ng) ? __ldg(a.bias + n0 + i) : 0.f;
        if (a.perm_r > 0 && etid < BN / 8) {         // once per n-tile instead of three divisions per (row, 8-column group)
          const int col = n0 + 8 * etid, r = a.perm_r, ij = col / Cp_d2s, c = col - ij * Cp_d2s;
          scol[etid] = ((ij / r) * a.GW * r + ij % r) * Cp_d2s + c;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        bias_tile = T.n_tile;
      }
      if (do_stat) {                                 // running sums follow (n_tile, group): flush when either changes
        const int grp = a.stat_groups > 1 ? (HALO ? T.m_tile / a.halo_tiles_per_img
                                                  : (int)(((long long)T.m_tile * BM) / a.stat_group_rows)) : 0;
        if (stat_tile >= 0 && (stat_tile != T.n_tile || stat_group != grp)) stat_flush<BN>(a, wstat, stat_tile, stat_group, lane);
        stat_tile = T.n_tile;
        stat_group = grp;
      }
      mbar_wait(bar_tfull + acc * 8, (j >> 1) & 1);
      tc_fence_after();
      const long long mrow0 = (long long)T.m_tile * BM + quad * 32;
      // lane L owns the addressing of row L of this warp's 32-row slab; the store loop fetches it with a shuffle
      long long my_o = 0, my_arow = 0;
      if (need_bhw) {                                // rows < 2^31 (checked by ladder_conv2d_tma_supported)
        int pw, ph, pb;
        if constexpr (HALO) {                        // row r of the tile is pixel (ty * 16 + r / 8, tx * 8 + r % 8) of image pb
          pb = T.m_tile / a.halo_tiles_per_img;
          const int rem = T.m_tile - pb * a.halo_tiles_per_img, ty = rem / a.halo_tiles_x, r = quad * 32 + lane;
          ph = ty * HALO_TH + (r >> 3);
          pw = (rem - ty * a.halo_tiles_x) * HALO_TW + (r & 7);
        } else {
          const unsigned mu = (unsigned)(mrow0 + lane);
          pw = (int)(mu % (unsigned)a.GW);
          const unsigned rq = mu / (unsigned)a.GW;
          ph = (int)(rq % (unsigned)a.GH);
          pb = (int)(rq / (unsigned)a.GH);
        }
        my_arow = (((long long)pb * a.GH + ph) * a.GW + pw) * Ng;
        if (MODE == DGRAD && a.os > 1)               // parity class of a strided dgrad: scatter into the full map
          my_arow = (((long long)pb * a.OHf + ph * a.os + a.opy) * a.OWf + pw * a.os + a.opx) * Ng;
        my_o = my_arow;
        if (a.perm_r > 0) {
          if (MODE == FPROP) {                       // depth_to_space: row part of d2s_dest(m, 0, ...)
            my_o = (((long long)pb * a.GH * pr + ph * pr) * ((long long)a.GW * pr) + pw * pr) * Cp_d2s;
          } else {                                   // space_to_depth: s2d_dest(m, 0, ...)
            const int yq = rsh >= 0 ? ph >> rsh : ph / pr, xq = rsh >= 0 ? pw >> rsh : pw / pr;
            my_o = (((long long)pb * GHr + yq) * GWr + xq) * ((long long)Ng * pr * pr) + ((ph - yq * pr) * pr + (pw - xq * pr)) * Ng;
          }
        }
      }
      // tcgen05.ld hands this thread ONE accumulator row (32 columns per chunk): it is finished and stored from registers,
      // 64 (bf16) / 128 (fp32) contiguous bytes per thread and chunk.  (r1 transposed every chunk through shared memory so that
      // 8 lanes shared a row: ~110 SASS instructions per 4-row store iteration, 10.9 k cycles per 128x128 tile against 4.6 k
      // of MMA -- gpurun_out/r2f_conv7.ncu-rep -- i.e. the epilogue, not the tensor pipe, bounded every large conv.)
      const long long m_row = mrow0 + lane;
      const bool row_ok = HALO || m_row < Mg;
      const long long o_row = need_bhw ? my_o : m_row * Ng, a_row = need_bhw ? my_arow : m_row * Ng;
      const bool perm_f = MODE == FPROP && a.perm_r > 0;
      const bool fast_n = (Ng & 7) == 0 && (!perm_f || (Cp_d2s & 7) == 0);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c0, v);
        const int nb = n0 + c0;
        if (nb >= Ng) continue;                    // warp-uniform
        if (MODE == FPROP) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float u = v[i] + sbias[c0 + i];
            v[i] = u > 0.f ? u : u * slope;
          }
          if (is_tanh) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
          }
        }
        if (do_stat) {                               // column sums need the transpose: lane = column over this warp's rows
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(stage + lane * 36 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          __syncwarp();
          const long long left = HALO ? 32 : Mg - mrow0;
          const int rows = left >= 32 ? 32 : (left > 0 ? (int)left : 0);
          float s1 = 0.f, s2 = 0.f;
          for (int r = 0; r < rows; ++r) {
            const float u = stage[r * 36 + lane];      // bank (4 r + lane) % 32: conflict free
            s1 += u;
            s2 = fmaf(u, u, s2);
          }
          wstat[c0 + lane] += s1;
          wstat[256 + c0 + lane] += s2;
          __syncwarp();
        }
        if (!row_ok) continue;
        if (fast_n && nb + 32 <= Ng) {
          // ---- 4 groups of 8 columns; a depth_to_space store re-bases each group (C' % 8 == 0)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = nb + 8 * q;
            const long long coloff = perm_f ? (long long)scol[(c0 >> 3) + q] : (long long)col;
            float e[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) e[t] = v[8 * q + t];
            if (MODE == DGRAD) {
              if (a.aux != nullptr) {
                float ax[8];
                if (a.aux_bf16) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(auxh + a_row + col));
                  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) { ax[2 * t] = bf16_bits_to_float(w4[t] & 0xffffu); ax[2 * t + 1] = bf16_bits_to_float(w4[t] >> 16); }
                } else {
                  const float4 f0 = __ldg(reinterpret_cast<const float4*>(auxf + a_row + col)),
                               f1 = __ldg(reinterpret_cast<const float4*>(auxf + a_row + col + 4));
                  ax[0] = f0.x; ax[1] = f0.y; ax[2] = f0.z; ax[3] = f0.w; ax[4] = f1.x; ax[5] = f1.y; ax[6] = f1.z; ax[7] = f1.w;
                }
#pragma unroll
                for (int t = 0; t < 8; ++t) e[t] *= is_tanh ? 1.f - ax[t] * ax[t] : (ax[t] > 0.f ? 1.f : slope);
              }
              if (a.accumulate) {
                if (OUT16) {
                  const uint4 u = *reinterpret_cast<const uint4*>(outh + o_row + coloff);
                  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int t = 0; t < 4; ++t) { e[2 * t] += bf16_bits_to_float(w4[t] & 0xffffu); e[2 * t + 1] += bf16_bits_to_float(w4[t] >> 16); }
                } else {
                  const float4 f0 = *reinterpret_cast<const float4*>(outf + o_row + coloff),
                               f1 = *reinterpret_cast<const float4*>(outf + o_row + coloff + 4);
                  e[0] += f0.x; e[1] += f0.y; e[2] += f0.z; e[3] += f0.w; e[4] += f1.x; e[5] += f1.y; e[6] += f1.z; e[7] += f1.w;
                }
              }
            }
            if (MODE == WGRAD) {
              atomicAdd(reinterpret_cast<float4*>(outf + o_row + coloff), make_float4(e[0], e[1], e[2], e[3]));   // red.global.add.v4.f32
              atomicAdd(reinterpret_cast<float4*>(outf + o_row + coloff + 4), make_float4(e[4], e[5], e[6], e[7]));
            } else if (OUT16) {
              *reinterpret_cast<uint4*>(outh + o_row + coloff) = pack8(e);
            } else {
              *reinterpret_cast<float4*>(outf + o_row + coloff) = make_float4(e[0], e[1], e[2], e[3]);
              *reinterpret_cast<float4*>(outf + o_row + coloff + 4) = make_float4(e[4], e[5], e[6], e[7]);
            }
          }
        } else {
          // ---- ragged N / narrow depth_to_space groups: 4 columns at a time through the scalar path
          for (int q = 0; q < 8; ++q) {
            const int col = nb + 4 * q;
            if (col >= Ng) break;
            long long coloff = col;
            if (perm_f) {
              const int r = a.perm_r, ij = col / Cp_d2s, c = col - ij * Cp_d2s;
              coloff = ((long long)(ij / r) * a.GW * r + ij % r) * Cp_d2s + c;
            }
            slow_store<MODE, OUT16>(a, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]), m_row, col, o_row + coloff,
                                    a_row, slope, is_tanh);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + acc * 8);
      ++j;
      decode_tile<MODE>(a, T.t + gridDim.x, total, T);
    }
    if (do_stat && stat_tile >= 0) stat_flush<BN>(a, wstat, stat_tile, stat_group, lane);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// Box {64 channels, bw, bh, bb} covering `px` consecutive pixels of the row-major [B, GH, GW] grid; false if the grid
// cannot be cut into such boxes (then the register-gather kernel of conv_tc.cu takes the layer).
static bool pixel_box(int GW, int GH, int B, int px, int& bw, int& bh, int& bb) {
  bw = bh = bb = 1;
  if (GH == 1 && B == 1) {   // one row of pixels (dense layers, see dense_as_row): a ragged last box is zero filled
    bw = px;
    return true;
  }
  if (GW >= px) {
    if (GW % px) return false;
    bw = px;
    return true;
  }
  if (px % GW) return false;
  bw = GW;
  const int rem = px / GW;
  if (GH >= rem) {
    if (GH % rem) return false;
    bh = rem;
    return true;
  }
  if (rem % GH) return false;
  bh = GH;
  bb = rem / GH;
  (void)B;                 // a box taller than the batch is legal: rows past B are zero filled and masked
  return bb <= 256;
}

// Dense layers ([B,1,1,C] "images") are presented to TMA as ONE image row of B pixels: boxes {64, px, 1, 1} walk the
// fastest pixel axis.  (Boxes {64, 1, 1, px} over the batch axis fetch one 128-byte row per outermost index and ran ~5x
// slower: 1.1 TB/s on the 1M-row tap-GEMM of the fashion decoder.)
static void dense_as_row(int& B, int& H, int& W, int KH, int KW, int& OH, int& OW) {
  if (H == 1 && W == 1 && OH == 1 && OW == 1 && KH == 1 && KW == 1) {
    W = OW = B;
    B = 1;
  }
}

// bf16 NHWC tensor [B, SH, SW, C] as a 4-D tensor map {C, SW, SH, B}, SWIZZLE_128B, zero fill outside
static int make_map(CUtensorMap* map, const void* ptr, int B, int SH, int SW, int C, int bw, int bh, int bb, int es = 1);

// halo mode: environment switches (diagnostics): LADDER_HALO=0 disables it, LADDER_HALO_BASE=0 leaves the descriptor's
// base-offset field zero
static int g_halo_on = -1, g_halo_base = -1;       // -1: from the environment on first use; set by ladder_conv2d_tma_set_halo
static bool halo_enabled() {
  if (g_halo_on < 0) { const char* e = getenv("LADDER_HALO"); g_halo_on = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1); }
  return g_halo_on != 0;
}
static int halo_base_mode() {
  // measured (scripts/halo_probe.py, r2f): the hardware applies SWIZZLE_128B to the ADDRESS bits of every row it reads, so a
  // view that starts kw rows into an atom needs NO base offset (base_mode 0 is bit-exact against the per-tap kernel, 1 is wrong)
  if (g_halo_base < 0) { const char* e = getenv("LADDER_HALO_BASE"); g_halo_base = (e && e[0] == '1') ? 1 : 0; }
  return g_halo_base;
}
// stride-1 3x3 layer over a [B, GH, GW] pixel grid whose epilogue never needs the scalar slow path
static bool halo_ok(int mode, int GH, int GW, int C, int KH, int KW, int stride, int Ng, int perm_r) {
  // measured (scripts/halo_probe.py, r2g): +6 % at N = 256 (fashion decoder fprop, 854 -> 903 TFLOP/s), -4 .. -8 % at N <= 128,
  // where the 9 weight tiles per chunk, not the activation tile, are the longer stream: automatic use is limited to N >= 256
  // (LADDER_HALO=2 / ladder_conv2d_tma_set_halo(2, .) forces it everywhere, for tests and probes)
  if (!halo_enabled() || (g_halo_on != 2 && Ng < 256)) return false;
  if (stride != 1 || KH != 3 || KW != 3 || GH % HALO_TH || GW % HALO_TW || C % BK || (Ng & 3)) return false;
  if (perm_r > 0 && mode == FPROP && ((Ng / (perm_r * perm_r)) & 3)) return false;
  return true;
}

static int make_map(CUtensorMap* map, const void* ptr, int B, int SH, int SW, int C, int bw, int bh, int bb, int es) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LADDER_ERR_CUDA, "conv2d_tma: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)SW, (cuuint64_t)SH, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)SW * C * 2, (cuuint64_t)SH * SW * C * 2};
  // element stride es > 1 (strided conv): the box spans bw*es x bh*es source pixels and TMA keeps every es-th one
  const cuuint32_t box[4] = {64, (cuuint32_t)(bw * es), (cuuint32_t)(bh * es), (cuuint32_t)bb};
  const cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LADDER_ERR_CUDA, "conv2d_tma: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return LADDER_OK;
}

template <int MODE, bool HALO = false>
static int launch(const CUtensorMap& mA, const CUtensorMap& mB, Args& a, long long Mg, int bn, cudaStream_t st) {
  a.m_tiles = (int)ceil_div64(Mg, BM);
  a.n_tiles = ceil_div(a.Ng, bn);
  const long long total = (long long)a.m_tiles * a.n_tiles * a.splits;
  const int sms = num_sms();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  auto go = [&](auto kern, int BNv) {
    const size_t ab = HALO ? (size_t)halo_stages(BNv) * HALO_BYTES + (size_t)halo_b_stages(BNv) * BNv * BK * 2
                           : (size_t)stages_for(BNv) * (A_STAGE_BYTES + BNv * BK * 2);
    const size_t smem = ab + 1024 + 256 + EPI_BYTES + BIAS_BYTES + STAT_BYTES;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, NTHREADS, smem, st>>>(mA, mB, a);
  };
  if constexpr (HALO) {
    if (a.out_bf16) {
      switch (bn) {
        case 32: go(tma_kernel<MODE, 32, true, true>, 32); break;
        case 64: go(tma_kernel<MODE, 64, true, true>, 64); break;
        case 128: go(tma_kernel<MODE, 128, true, true>, 128); break;
        default: go(tma_kernel<MODE, 256, true, true>, 256); break;
      }
    } else {
      switch (bn) {
        case 32: go(tma_kernel<MODE, 32, false, true>, 32); break;
        case 64: go(tma_kernel<MODE, 64, false, true>, 64); break;
        case 128: go(tma_kernel<MODE, 128, false, true>, 128); break;
        default: go(tma_kernel<MODE, 256, false, true>, 256); break;
      }
    }
    return check_launch("tcgen05 TMA conv kernel (halo)");
  }
  if (a.out_bf16) {
    if constexpr (MODE != WGRAD) {
      switch (bn) {
        case 32: go(tma_kernel<MODE, 32, true>, 32); break;
        case 64: go(tma_kernel<MODE, 64, true>, 64); break;
        case 128: go(tma_kernel<MODE, 128, true>, 128); break;
        default: go(tma_kernel<MODE, 256, true>, 256); break;
      }
    }
  } else {
    switch (bn) {
      case 32: if constexpr (MODE != WGRAD) { go(tma_kernel<MODE, 32, false>, 32); } break;
      case 64: go(tma_kernel<MODE, 64, false>, 64); break;
      case 128: go(tma_kernel<MODE, 128, false>, 128); break;
      default: go(tma_kernel<MODE, 256, false>, 256); break;
    }
  }
  return check_launch("tcgen05 TMA conv kernel");
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  const long long n8 = n / 8;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i), v1 = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
  for (long long i = n8 * 8 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __float2bfloat16_rn(x[i]);
}

__global__ void bf16_to_f32_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] = __bfloat162float(x[i]);
}

// column sums of a bf16 [rows, cols] matrix (bias gradient of a layer whose dy is stored in bf16)
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ g, long long rows, int cols, float* __restrict__ out) {
  // block = 256 threads = 8 row-lanes x 32 column-lanes; grid.x = column groups of 32, grid.y = row slices
  __shared__ float part[8][33];
  const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < cols)
    for (long long r = (long long)blockIdx.y * 8 + rl; r < rows; r += (long long)gridDim.y * 8) s += __bfloat162float(g[r * cols + c]);
  part[rl][cl] = s;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][cl];
    atomicAdd(out + c, t);
  }
}

}  // namespace tma
}  // namespace ladder

using namespace ladder;
using namespace ladder::tma;

// N tile: as wide as the GEMM allows, narrowed while the launch would fill less than half of the SMs (small-batch
// dense layers: M = batch rows give a single M tile, so the CTA count comes from N tiles alone)
static int choose_bn(int mode, long long Mg, int Ng, long long wgrad_kb = 0) {
  int bn = tc::pick_bn(Ng);
  const int floor_bn = mode == WGRAD ? 64 : 32;
  if (mode == WGRAD && bn < 64) bn = 64;
  if (mode == WGRAD && wgrad_kb > 4) return bn;      // long pixel reductions get their CTAs from split-K instead
  const long long m_tiles = ceil_div64(Mg, BM);
  while (bn > floor_bn && m_tiles * ceil_div(Ng, bn) * 2 <= num_sms()) bn >>= 1;
  return bn;
}

static bool geometry_ok(int B, int GH, int GW, int px) {
  int bw, bh, bb;
  return pixel_box(GW, GH, B, px, bw, bh, bb);
}

extern "C" {

/* mode 0 fprop, 1 dgrad, 2 wgrad: 1 iff the TMA-fed kernel takes this geometry (64-aligned channels of the gathered
 * tensor, pixel grid divisible into 128- (64- for wgrad) pixel boxes; strided fprop / wgrad use the tensor map's element
 * stride, strided dgrad runs one stride-1 GEMM per output-parity class -- no zero-insertion, no wasted MACs) */
int ladder_conv2d_tma_supported(int mode, int B, int H, int W, int Cin, int KH, int KW, int Cout, int stride, int OH, int OW) {
  if (stride < 1 || stride > 8 || B <= 0) return 0;
  dense_as_row(B, H, W, KH, KW, OH, OW);
  if ((long long)B * H * W >= (1LL << 31) / 64 * 64 || (long long)B * OH * OW >= (1LL << 31) / 64 * 64) return 0;
  int bw, bh, bb;
  switch (mode) {
    case 0: return Cin % 64 == 0 && pixel_box(OW, OH, B, 128, bw, bh, bb) && bw * stride <= 256 && bh * stride <= 256;
    case 1:
      if (Cout % 64 != 0) return 0;
      if (stride == 1) return geometry_ok(B, H, W, 128);
      // strided: one stride-1 GEMM per output-parity class over the [B, H/s, W/s] grid (every class needs >= 1 tap)
      return H % stride == 0 && W % stride == 0 && KH >= stride && KW >= stride && geometry_ok(B, H / stride, W / stride, 128);
    case 2: return Cin % 64 == 0 && Cout % 64 == 0 && pixel_box(OW, OH, B, 64, bw, bh, bb) && bw * stride <= 256 && bh * stride <= 256;
    default: return 0;
  }
}

size_t ladder_conv2d_tma_workspace_bytes(int Cin, int KH, int KW, int Cout) {
  size_t m = 0;
  for (int bn = 32; bn <= 256; bn <<= 1) {        // any N tile the launcher may pick
    const size_t f = tc::pack_bytes(Cout, KH * KW * Cin, bn), d = tc::pack_bytes(Cin, KH * KW * Cout, bn);
    if (f > m) m = f;
    if (d > m) m = d;
  }
  return m + 256 + 64 * 1024;    // + alignment slack of the per-class images of a strided dgrad
}

/* diagnostic switch of the halo mode of the stride-1 3x3 fprop / dgrad kernels (see conv_tma.cu): enabled 0 / 1, base_mode 0 / 1
 * (negative: leave unchanged); returns the previous `enabled`.  Defaults: LADDER_HALO / LADDER_HALO_BASE, both 1. */
int ladder_conv2d_tma_set_halo(int enabled, int base_mode) {
  const int prev = halo_enabled() ? 1 : 0;
  halo_base_mode();
  if (enabled >= 0) g_halo_on = enabled > 2 ? 1 : enabled;
  if (base_mode >= 0) g_halo_base = base_mode ? 1 : 0;
  return prev;
}

/* N tile width the TMA launcher uses for GEMM `mode` of this geometry (the packed weight image depends on it) */
int ladder_conv2d_tma_bn(int mode, int B, int H, int W, int Cin, int Cout, int OH, int OW) {
  switch (mode) {
    case 0: return choose_bn(FPROP, (long long)B * OH * OW, Cout);
    case 1: return choose_bn(DGRAD, (long long)B * H * W, Cin);
    default: return 0;
  }
}

/* bf16 tile image of one layer's weights for mode 0 (fprop) / 1 (dgrad) with N tile bn, as the TMA kernels stage it */
size_t ladder_conv2d_tma_pack_bytes(int mode, int KH, int KW, int Cin, int Cout, int bn) {
  return mode == 0 ? tc::pack_bytes(Cout, KH * KW * Cin, bn) : tc::pack_bytes(Cin, KH * KW * Cout, bn);
}
int ladder_conv2d_tma_pack(const float* w, void* image, size_t image_bytes, int mode, int KH, int KW, int Cin, int Cout, int bn,
                           cudaStream_t stream) {
  LADDER_REQUIRE(w && image && (mode == 0 || mode == 1) && bn >= 32, "conv2d_tma_pack: bad arguments");
  return tc::pack(w, image, image_bytes, mode, KH * KW, Cin, Cout, stream, bn);
}

int ladder_f32_to_bf16(const float* x, void* y, long long n, cudaStream_t stream) {
  if (n <= 0) return LADDER_OK;
  LADDER_REQUIRE(x && y, "f32_to_bf16: null pointer");
  LADDER_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "f32_to_bf16: pointers must be 16-byte aligned");
  long long blocks = ceil_div64(ceil_div64(n, 8), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(y), n);
  return check_launch("f32_to_bf16");
}

int ladder_bf16_to_f32(const void* x, float* y, long long n, cudaStream_t stream) {
  if (n <= 0) return LADDER_OK;
  LADDER_REQUIRE(x && y, "bf16_to_f32: null pointer");
  long long blocks = ceil_div64(n, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  bf16_to_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), y, n);
  return check_launch("bf16_to_f32");
}

int ladder_colsum_bf16(const void* g, long long rows, int cols, float* out, cudaStream_t stream) {
  LADDER_REQUIRE(g && out && rows > 0 && cols > 0, "colsum_bf16: bad arguments");
  cudaError_t e = cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "colsum_bf16 memset: %s", cudaGetErrorString(e));
  const int gx = ceil_div(cols, 32);
  long long gy = ceil_div64(rows, 8 * 16);
  const long long cap = (148 * 8 + gx - 1) / gx;
  if (gy > cap) gy = cap;
  if (gy < 1) gy = 1;
  colsum_bf16_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(g), rows, cols, out);
  return check_launch("colsum_bf16");
}

int ladder_conv2d_fprop_tma(const void* x_bf16, const float* w, const float* bias, void* y, int y_bf16, int B, int H, int W,
                            int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act,
                            int out_d2s, void* workspace, size_t workspace_bytes, float* stat_sums, int stat_groups,
                            cudaStream_t stream) {
  LADDER_REQUIRE(x_bf16 && y && Cin > 0 && Cout > 0 && KH > 0 && KW > 0, "conv2d_fprop_tma: bad arguments");
  LADDER_REQUIRE(ladder_conv2d_tma_supported(0, B, H, W, Cin, KH, KW, Cout, stride, OH, OW),
                 "conv2d_fprop_tma: unsupported geometry (see ladder_conv2d_tma_supported)");
  LADDER_REQUIRE(out_d2s == 0 || (out_d2s > 0 && Cout % (out_d2s * out_d2s) == 0),
                 "conv2d_fprop_tma: depth_to_space(%d) output needs Cout %% r^2 == 0", out_d2s);
  LADDER_REQUIRE(((uintptr_t)x_bf16 & 15) == 0, "conv2d_fprop_tma: x must be 16-byte aligned");
  const int stat_group_rows = OH * OW;
  if (stat_sums != nullptr) {
    LADDER_REQUIRE(act == ACT_NONE && out_d2s == 0, "conv2d_fprop_tma: output statistics are those of the linear output (act none, no d2s)");
    LADDER_REQUIRE(stat_groups == 1 || (stat_groups == B && stat_group_rows % BM == 0),
                   "conv2d_fprop_tma: per-sample statistics need OH*OW %% 128 == 0 (got %d) and groups == B", stat_group_rows);
    cudaError_t e = cudaMemsetAsync(stat_sums, 0, (size_t)2 * stat_groups * Cout * sizeof(float), stream);
    if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "conv2d_fprop_tma memset: %s", cudaGetErrorString(e));
  }
  dense_as_row(B, H, W, KH, KW, OH, OW);
  const int bn = choose_bn(FPROP, (long long)B * OH * OW, Cout);
  int rc = LADDER_OK;
  if (w != nullptr) rc = tc::pack(w, workspace, workspace_bytes, FPROP, KH * KW, Cin, Cout, stream, bn);
  else if (workspace == nullptr || workspace_bytes < tc::pack_bytes(Cout, KH * KW * Cin, bn))
    rc = fail(LADDER_ERR_WORKSPACE, "conv2d_fprop_tma: packed weight image too small");
  if (rc) return rc;
  const bool halo = halo_ok(FPROP, OH, OW, Cin, KH, KW, stride, Cout, out_d2s) && H == OH && W == OW;
  int bw, bh, bb;
  pixel_box(OW, OH, B, BM, bw, bh, bb);
  CUtensorMap mA;
  rc = halo ? make_map(&mA, x_bf16, B, H, W, Cin, HALO_PITCH, HALO_ROWS, 1) : make_map(&mA, x_bf16, B, H, W, Cin, bw, bh, bb, stride);
  if (rc) return rc;
  Args a;
  memset(&a, 0, sizeof(a));
  a.stride = stride;
  a.wt = static_cast<const __nv_bfloat16*>(workspace);
  a.bias = bias; a.out = y; a.out_bf16 = y_bf16;
  a.GW = OW; a.GH = OH; a.B = B; a.C = Cin; a.KH = KH; a.KW = KW;
  a.off_y = -pad_t; a.off_x = -pad_l; a.sign = 1;
  a.Ng = Cout; a.act = act; a.nkb = KH * KW * (Cin / BK); a.splits = 1; a.perm_r = out_d2s;
  a.stat = stat_sums; a.stat_groups = stat_sums ? stat_groups : 0; a.stat_group_rows = stat_group_rows;
  if (halo) {
    a.halo_ox = -pad_l; a.halo_oy = -pad_t;                   // taps read offsets off + k, k = 0..2: the origin is tap 0
    a.halo_tiles_x = OW / HALO_TW; a.halo_tiles_per_img = (OH / HALO_TH) * (OW / HALO_TW);
    a.halo_base_mode = halo_base_mode();
    return launch<FPROP, true>(mA, mA, a, (long long)B * OH * OW, bn, stream);
  }
  return launch<FPROP>(mA, mA, a, (long long)B * OH * OW, bn, stream);
}

int ladder_conv2d_dgrad_tma(const void* dy_bf16, const float* w, const void* act_out, int act_out_bf16, void* dx, int dx_bf16,
                            int B, int H, int W, int Cin, int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH,
                            int OW, int act, int accumulate, int out_s2d, void* workspace, size_t workspace_bytes,
                            cudaStream_t stream) {
  LADDER_REQUIRE(dy_bf16 && dx && Cin > 0 && Cout > 0 && KH > 0 && KW > 0, "conv2d_dgrad_tma: bad arguments");
  LADDER_REQUIRE(ladder_conv2d_tma_supported(1, B, H, W, Cin, KH, KW, Cout, stride, OH, OW),
                 "conv2d_dgrad_tma: unsupported geometry (see ladder_conv2d_tma_supported)");
  LADDER_REQUIRE(out_s2d == 0 || (out_s2d > 0 && H % out_s2d == 0 && W % out_s2d == 0),
                 "conv2d_dgrad_tma: space_to_depth(%d) output needs H, W divisible by r", out_s2d);
  LADDER_REQUIRE(((uintptr_t)dy_bf16 & 15) == 0, "conv2d_dgrad_tma: dy must be 16-byte aligned");
  if (stride > 1) {
    // dx(s*i + py, s*j + px) = sum over taps kh = kh0 + s*a, kw = kw0 + s*b (kh0 = (py + pad_t) % s, ...) of
    // dy(i + (py + pad_t - kh0)/s - a, j + (px + pad_l - kw0)/s - b) . w(kh, kw): a stride-1 correlation per class
    LADDER_REQUIRE(w != nullptr, "conv2d_dgrad_tma: strided dgrad packs its per-class weight images itself (w required)");
    LADDER_REQUIRE(out_s2d == 0, "conv2d_dgrad_tma: space_to_depth output is not available for strided layers");
    const int GH = H / stride, GW = W / stride;
    const int bn = choose_bn(DGRAD, (long long)B * GH * GW, Cin);
    int bw, bh, bb;
    pixel_box(GW, GH, B, BM, bw, bh, bb);
    CUtensorMap mA;
    int rc = make_map(&mA, dy_bf16, B, OH, OW, Cout, bw, bh, bb);
    if (rc) return rc;
    size_t ws_off = 0;
    for (int py = 0; py < stride; ++py)
      for (int px = 0; px < stride; ++px) {
        const int kh0 = (py + pad_t) % stride, kw0 = (px + pad_l) % stride;
        const int nkh = (KH - kh0 + stride - 1) / stride, nkw = (KW - kw0 + stride - 1) / stride;
        const size_t bytes = tc::pack_bytes(Cin, nkh * nkw * Cout, bn);
        LADDER_REQUIRE(workspace != nullptr && ws_off + bytes <= workspace_bytes, "conv2d_dgrad_tma: workspace too small");
        uint8_t* img = static_cast<uint8_t*>(workspace) + ws_off;
        rc = tc::pack(w, img, bytes, DGRAD, nkh * nkw, Cin, Cout, stream, bn, tc::TapMap{KW, kh0, kw0, nkw, stride});
        if (rc) return rc;
        ws_off += (bytes + 1023) / 1024 * 1024;
        Args a;
        memset(&a, 0, sizeof(a));
        a.stride = 1;
        a.wt = reinterpret_cast<const __nv_bfloat16*>(img);
        a.aux = act_out; a.aux_bf16 = act_out_bf16; a.out = dx; a.out_bf16 = dx_bf16;
        a.GW = GW; a.GH = GH; a.B = B; a.C = Cout; a.KH = nkh; a.KW = nkw;
        a.off_y = (py + pad_t - kh0) / stride; a.off_x = (px + pad_l - kw0) / stride; a.sign = -1;
        a.Ng = Cin; a.act = act; a.accumulate = accumulate; a.nkb = nkh * nkw * (Cout / BK); a.splits = 1;
        a.os = stride; a.opy = py; a.opx = px; a.OHf = H; a.OWf = W;
        rc = launch<DGRAD>(mA, mA, a, (long long)B * GH * GW, bn, stream);
        if (rc) return rc;
      }
    return LADDER_OK;
  }
  dense_as_row(B, H, W, KH, KW, OH, OW);
  const int bn = choose_bn(DGRAD, (long long)B * H * W, Cin);
  int rc = LADDER_OK;
  if (w != nullptr) rc = tc::pack(w, workspace, workspace_bytes, DGRAD, KH * KW, Cin, Cout, stream, bn);
  else if (workspace == nullptr || workspace_bytes < tc::pack_bytes(Cin, KH * KW * Cout, bn))
    rc = fail(LADDER_ERR_WORKSPACE, "conv2d_dgrad_tma: packed weight image too small");
  if (rc) return rc;
  const bool halo = halo_ok(DGRAD, H, W, Cout, KH, KW, 1, Cin, out_s2d) && H == OH && W == OW;
  int bw, bh, bb;
  pixel_box(W, H, B, BM, bw, bh, bb);
  CUtensorMap mA;
  rc = halo ? make_map(&mA, dy_bf16, B, OH, OW, Cout, HALO_PITCH, HALO_ROWS, 1) : make_map(&mA, dy_bf16, B, OH, OW, Cout, bw, bh, bb);
  if (rc) return rc;
  Args a;
  memset(&a, 0, sizeof(a));
  a.wt = static_cast<const __nv_bfloat16*>(workspace);
  a.stride = 1;
  a.aux = act_out; a.aux_bf16 = act_out_bf16; a.out = dx; a.out_bf16 = dx_bf16;
  a.GW = W; a.GH = H; a.B = B; a.C = Cout; a.KH = KH; a.KW = KW;
  a.off_y = pad_t; a.off_x = pad_l; a.sign = -1;       // dx(y, x) += dy(y + pad_t - kh, x + pad_l - kw) . w(kh, kw)
  a.Ng = Cin; a.act = act; a.accumulate = accumulate; a.nkb = KH * KW * (Cout / BK); a.splits = 1; a.perm_r = out_s2d;
  if (halo) {
    a.halo_ox = pad_l - (KW - 1); a.halo_oy = pad_t - (KH - 1);    // taps read offsets pad - k: the origin is the LAST tap
    a.halo_tiles_x = W / HALO_TW; a.halo_tiles_per_img = (H / HALO_TH) * (W / HALO_TW);
    a.halo_base_mode = halo_base_mode();
    return launch<DGRAD, true>(mA, mA, a, (long long)B * H * W, bn, stream);
  }
  return launch<DGRAD>(mA, mA, a, (long long)B * H * W, bn, stream);
}

/* dw (fp32, HWIO) is overwritten; the bias gradient is ladder_colsum_bf16(dy). */
int ladder_conv2d_wgrad_tma(const void* x_bf16, const void* dy_bf16, float* dw, int B, int H, int W, int Cin, int KH, int KW,
                            int Cout, int stride, int pad_t, int pad_l, int OH, int OW, cudaStream_t stream) {
  LADDER_REQUIRE(x_bf16 && dy_bf16 && dw && Cin > 0 && Cout > 0 && KH > 0 && KW > 0, "conv2d_wgrad_tma: bad arguments");
  LADDER_REQUIRE(ladder_conv2d_tma_supported(2, B, H, W, Cin, KH, KW, Cout, stride, OH, OW),
                 "conv2d_wgrad_tma: unsupported geometry (see ladder_conv2d_tma_supported)");
  LADDER_REQUIRE((((uintptr_t)x_bf16 | (uintptr_t)dy_bf16) & 15) == 0, "conv2d_wgrad_tma: operands must be 16-byte aligned");
  const int patch = KH * KW * Cin;
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)patch * Cout * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "conv2d_wgrad_tma memset: %s", cudaGetErrorString(e));
  dense_as_row(B, H, W, KH, KW, OH, OW);
  int bw, bh, bb;
  pixel_box(OW, OH, B, BK, bw, bh, bb);
  CUtensorMap mA, mB;
  int rc = make_map(&mA, x_bf16, B, H, W, Cin, bw, bh, bb, stride);
  if (rc) return rc;
  rc = make_map(&mB, dy_bf16, B, OH, OW, Cout, bw, bh, bb);
  if (rc) return rc;
  // pixel boxes of 64: the last one may hang over the batch axis (zero filled on both operands)
  const long long box_px = (long long)bw * bh * bb;            // == 64
  const long long total_kb = bb > 1 ? ceil_div64(B, bb) : ceil_div64((long long)B * OH * OW, box_px);
  const int bn = choose_bn(WGRAD, patch, Cout, total_kb);
  const long long tiles = ceil_div64(patch, BM) * ceil_div(Cout, bn);
  long long splits = (2LL * num_sms()) / tiles;                // <= 2 full waves of the persistent grid
  const long long max_splits = ceil_div64(total_kb, 4);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const long long per = ceil_div64(total_kb, splits);
  Args a;
  memset(&a, 0, sizeof(a));
  a.out = dw; a.stride = stride;
  a.GW = OW; a.GH = OH; a.B = B; a.C = Cin; a.KH = KH; a.KW = KW;
  a.off_y = -pad_t; a.off_x = -pad_l; a.sign = 1;
  a.Ng = Cout; a.kb_per_split = (int)per; a.total_kb = total_kb; a.m_valid = patch;
  a.splits = (int)ceil_div64(total_kb, per);
  return launch<WGRAD>(mA, mB, a, patch, bn, stream);
}

}  // extern "C"

// K1/K2 on the 5th-generation tensor cores: conv2d / dense as implicit GEMM with tcgen05.mma.
//
// Same three gather modes and the same C-ABI geometry as conv_igemm.cu (fp32 SIMT), but the
// mainloop is  D[tmem] += A[smem] * B[smem]  in bf16 with fp32 accumulation in TMEM:
//
//   * 4 producer warps gather the A operand (im2col rows for FPROP/DGRAD, transposed patch
//     columns for WGRAD) straight from the fp32 NHWC activations, convert to bf16 in registers
//     and store into the canonical K-major SWIZZLE_128B shared-memory layout (16-byte chunk
//     index XOR row&7) -- TF padding, stride, zero-insertion for strided dgrad all resolved in
//     the gather, so no im2col matrix and no bf16 copy of the activations ever exists in HBM.
//     The B operand is the per-call bf16 K-major repack of the weights (FPROP/DGRAD) or the
//     transposed dy tile (WGRAD).
//   * generic-proxy stores are published to the async proxy with fence.proxy.async, then an
//     mbarrier hand-off (full/empty ring, 4 stages) to
//   * 1 MMA warp: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN,
//     K=16) x4 per 64-wide k-block against shared-memory descriptors; tcgen05.commit releases
//     the stage, a final commit signals
//   * 4 epilogue warps: tcgen05.ld 32x32b.x32 TMEM -> registers, fused bias + activation
//     (FPROP), activation-derivative of the producer layer + accumulate (DGRAD) or split-K
//     red.global.add (WGRAD), vectorised stores.
#include "common.cuh"
#include "ladder_sm100.h"
#include <cuda_bf16.h>
#include <cstdlib>
#include "tc_ptx.cuh"

namespace ladder {
namespace tc {

constexpr int BM = 128;            // UMMA M (TMEM lanes)
constexpr int BK = 64;             // bf16 per k-block = one 128-byte swizzle row
// smem ring depth per N tile width (A 16 KB + B BN*128 B per stage, ~192 KB total)
__host__ __device__ constexpr int stages_for(int bn) { return bn >= 256 ? 4 : 4; }
constexpr int PROD_WARPS = 8;
constexpr int PRODUCERS = PROD_WARPS * 32;   // threads
constexpr int MMA_WARP = PROD_WARPS;         // warp 8
constexpr int NTHREADS = (PROD_WARPS + 1 + 4) * 32;   // 8 producer warps, 1 MMA warp, 4 epilogue warps = 416
constexpr int A_STAGE_BYTES = BM * BK * 2;

enum { FPROP = 0, DGRAD = 1, WGRAD = 2 };

struct TcArgs {
  const float* src;            // gathered activations: x (FPROP/WGRAD) or dy (DGRAD), fp32 NHWC
  const float* src2;           // WGRAD: dy [pixels, Cout]
  const __nv_bfloat16* wt;     // FPROP/DGRAD: packed weights Bt[N][Kpad], K-major bf16
  const float* bias;
  const float* aux;
  float* out;
  int B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW;
  int act, accumulate;
  int k_per_split;             // WGRAD: pixels per split (multiple of BK)
  int Kpad;                    // FPROP/DGRAD: padded reduction length (multiple of BK)
  int m_valid;                 // rows of the output that exist (WGRAD with padded taps)
  int m_tiles, n_tiles, splits;   // persistent tile space
  int perm_r;                  // FPROP: y in depth_to_space(r) layout; DGRAD: dx at the d2s-input position (0 = off)
  int debug;                   // bring-up only (LADDER_TC_DEBUG): 1 skip A loads, 2 skip B copies, 4 skip epilogue stores
};

// PTX wrappers (mbarrier, TMA, tcgen05, descriptors): tc_ptx.cuh

// ------------------------------------------------------------------ gather (same index math as conv_igemm.cu)
struct Geo {
  int B, H, W, Cin, KH, KW, Cout, stride, pad_t, pad_l, OH, OW;
};
template <bool FROM_DY>
__device__ __forceinline__ long long tap_offset(const TcArgs& a, int b, int y, int x, int kh, int kw, int c) {
  if (FROM_DY) {
    int ny = y + a.pad_t - kh, nx = x + a.pad_l - kw;
    if (ny < 0 || nx < 0) return -1;
    if (a.stride > 1) {
      if (ny % a.stride || nx % a.stride) return -1;
      ny /= a.stride; nx /= a.stride;
    }
    if (ny >= a.OH || nx >= a.OW) return -1;
    return (((long long)b * a.OH + ny) * a.OW + nx) * a.Cout + c;
  } else {
    const int iy = y * a.stride - a.pad_t + kh, ix = x * a.stride - a.pad_l + kw;
    if (iy < 0 || ix < 0 || iy >= a.H || ix >= a.W) return -1;
    return (((long long)b * a.H + iy) * a.W + ix) * a.Cin + c;
  }
}
// 8 consecutive patch entries q0..q0+7 of pixel (b, y, x) -> f[8] (zeros outside)
template <bool FROM_DY>
__device__ __forceinline__ void gather8(const TcArgs& a, int b, int y, int x, long long q0, int patch, int C, bool vec,
                                        float (&f)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  if (q0 >= patch) return;
  if (vec) {
    const int tap = (int)(q0 / C), c = (int)(q0 % C);
    const long long off = tap_offset<FROM_DY>(a, b, y, x, tap / a.KW, tap % a.KW, c);
    if (off >= 0) {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(a.src + off));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(a.src + off + 4));
      f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long q = q0 + j;
      if (q < patch) {
        const int tap = (int)(q / C), c = (int)(q % C);
        const long long off = tap_offset<FROM_DY>(a, b, y, x, tap / a.KW, tap % a.KW, c);
        if (off >= 0) f[j] = __ldg(a.src + off);
      }
    }
  }
}

__device__ __forceinline__ void sts64(uint32_t addr, float4 v) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&p0)),
               "r"(*reinterpret_cast<uint32_t*>(&p1)) : "memory");
}

// One unit of persistent work: output tile (m_tile, n_tile) and, for WGRAD, a slice of the pixel reduction.
struct Tile {
  long long t;
  int m_tile, n_tile, nkb;
  long long k_lo;     // WGRAD: first pixel of the slice
  bool valid;
};

template <int MODE>
__device__ __forceinline__ void decode_tile(const TcArgs& a, long long t, long long total, long long pixels, Tile& c) {
  c.t = t;
  c.valid = t < total;
  if (!c.valid) { c.nkb = 0; return; }
  c.n_tile = (int)(t % a.n_tiles);
  const long long r = t / a.n_tiles;
  if (MODE == WGRAD) {
    c.m_tile = (int)(r % a.m_tiles);
    const long long split = r / a.m_tiles;
    c.k_lo = split * a.k_per_split;
    const long long k_hi = min(pixels, c.k_lo + a.k_per_split);
    c.nkb = (int)((k_hi - c.k_lo + BK - 1) / BK);
  } else {
    c.m_tile = (int)r;
    c.k_lo = 0;
    c.nkb = a.Kpad / BK;
  }
}

// Persistent warp-specialised implicit GEMM.  grid = min(#tiles, #SMs); each CTA walks tiles blockIdx.x, +gridDim.x, ...
//   warps 0-7  producers: gather fp32 activations (coalesced: 16 lanes = 256 B = 64 channels of one row), keep TWO
//              k-blocks in flight in registers, convert to bf16, store the SWIZZLE_128B operand tile; one thread
//              also issues the bulk-TMA copy of the pre-packed bf16 B tile (weights, or dy for WGRAD)
//   warp  8    MMA issuer (tcgen05.mma, accumulators double-buffered in TMEM so the epilogue of tile i overlaps
//              the mainloop of tile i+1)
//   warps 9-12 epilogue (tcgen05.ld -> fused bias/activation | act'-multiply/accumulate | split-K red.add)
template <int MODE, int BN>
__global__ void __launch_bounds__(NTHREADS, 1) tc_kernel(TcArgs a) {
  constexpr int STAGES = stages_for(BN);
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                 // SWIZZLE_128B atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * A_STAGE_BYTES;
  const uint32_t bars = sB + STAGES * B_STAGE_BYTES;            // full[S], empty[S], tfull[2], tempty[2]
  const uint32_t bar_full = bars, bar_empty = bars + STAGES * 8, bar_tfull = bars + 2 * STAGES * 8,
                 bar_tempty = bars + (2 * STAGES + 2) * 8;
  const uint32_t slot = bars + (2 * STAGES + 4) * 8;
  volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (slot - base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long pixels = MODE == DGRAD ? (long long)a.B * a.H * a.W : (long long)a.B * a.OH * a.OW;
  const int C = MODE == DGRAD ? a.Cout : a.Cin;                  // channels of the gathered tensor
  const int patch = a.KH * a.KW * C;
  const long long Mg = MODE == WGRAD ? patch : pixels;
  const int Ng = MODE == DGRAD ? a.Cin : a.Cout;
  const long long total = (long long)a.m_tiles * a.n_tiles * a.splits;
  const int gw = MODE == DGRAD ? a.W : a.OW, gh = MODE == DGRAD ? a.H : a.OH;   // pixel grid of the rows

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, PRODUCERS);
      mbar_init(bar_empty + s * 8, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) tmem_alloc(slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *slot_ptr;

  if (warp < PROD_WARPS) {
    // ===================================================== producers
    const int f4 = lane & 15, rsel = lane >> 4;
    const bool fast = (C % BK == 0);
    Tile L, S;                       // load cursor (two k-blocks ahead) and store cursor
    int lkb = 0, skb = 0;
    decode_tile<MODE>(a, blockIdx.x, total, pixels, L);
    S = L;
    // Fast path (64-aligned channels, any stride, FPROP and DGRAD): per tile every thread keeps, for its 8 rows, the
    // batch base offset and the (y, x) origin of the patch; a k-block (= one tap, 64 channels) then costs a handful of
    // integer ops + one predicate per 16-byte load instead of re-deriving (b, y, x) -> offset per element.
    const bool affine = fast;
    long long rbase[8];              // batch offset of the gathered tensor, -1 for rows past the end
    int rvy[8], rvx[8];              // FPROP: y*s - pad_t, x*s - pad_l;  DGRAD: y + pad_t, x + pad_l
    auto pixel_affine = [&](long long p, long long lim, long long& base, int& oy, int& ox) {
      base = -1; oy = 0; ox = 0;
      if (p >= lim) return;
      const unsigned pu = (unsigned)p;
      const int x = (int)(pu % (unsigned)gw);
      const unsigned r = pu / (unsigned)gw;
      const int y = (int)(r % (unsigned)gh), b = (int)(r / (unsigned)gh);
      if (MODE == DGRAD) {
        base = (long long)b * a.OH * a.OW * a.Cout;
        oy = y + a.pad_t; ox = x + a.pad_l;
      } else {
        base = (long long)b * a.H * a.W * a.Cin;
        oy = y * a.stride - a.pad_t; ox = x * a.stride - a.pad_l;
      }
    };
    // element offset of tap (kh, kw) for a row prepared by pixel_affine, or -1 (padding / between strides)
    auto tap_off = [&](long long base, int oy, int ox, int kh, int kw) -> long long {
      if (base < 0) return -1;
      if (MODE == DGRAD) {
        int ny = oy - kh, nx = ox - kw;
        if (ny < 0 || nx < 0) return -1;
        if (a.stride == 2) {
          if ((ny | nx) & 1) return -1;
          ny >>= 1; nx >>= 1;
        } else if (a.stride > 2) {
          if (ny % a.stride || nx % a.stride) return -1;
          ny /= a.stride; nx /= a.stride;
        }
        if (ny >= a.OH || nx >= a.OW) return -1;
        return base + ((long long)ny * a.OW + nx) * a.Cout;
      } else {
        const int iy = oy + kh, ix = ox + kw;
        if (iy < 0 || ix < 0 || iy >= a.H || ix >= a.W) return -1;
        return base + ((long long)iy * a.W + ix) * a.Cin;
      }
    };
    auto row_coords = [&]() {
      if (MODE == WGRAD || !L.valid || !affine) return;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        pixel_affine((long long)L.m_tile * BM + warp * 16 + i * 2 + rsel, pixels, rbase[i], rvy[i], rvx[i]);
    };
    row_coords();
    auto load = [&](float4 (&v)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!L.valid || (a.debug & 1)) return;
      if (MODE != WGRAD) {
        if (affine) {
          // K order on the fast path: 64-channel chunk major, tap minor -- the KH*KW consecutive k-blocks of one
          // chunk re-read the same [pixels x 64 channels] slab shifted by one pixel, so all but the first hit L1
          const int taps = a.KH * a.KW, tap = lkb % taps, c0 = (lkb / taps) * BK, kh = tap / a.KW, kw = tap % a.KW;
          const float* srck = a.src + c0 + 4 * f4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long off = tap_off(rbase[i], rvy[i], rvx[i], kh, kw);
            if (off >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(srck + off));
          }
        } else {
          // general gather (narrow / ragged channel counts, strided dgrad): 4 patch entries per lane, element-wise
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const long long p = (long long)L.m_tile * BM + warp * 16 + i * 2 + rsel;
            if (p >= pixels) continue;
            const unsigned pu = (unsigned)p;
            const int x = (int)(pu % (unsigned)gw);
            const unsigned r = pu / (unsigned)gw;
            const int y = (int)(r % (unsigned)gh), b = (int)(r / (unsigned)gh);
            float e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int q = lkb * BK + 4 * f4 + j;
              if (q < patch) {
                int tap = q / C, c = q % C;
                if (fast) {            // same channel-chunk-major K order as the weight image (strided dgrad lands here)
                  const int taps = a.KH * a.KW;
                  tap = lkb % taps;
                  c = (lkb / taps) * BK + 4 * f4 + j;
                }
                const long long off = tap_offset<MODE == DGRAD>(a, b, y, x, tap / a.KW, tap % a.KW, c);
                if (off >= 0) e[j] = __ldg(a.src + off);
              }
            }
            v[i] = make_float4(e[0], e[1], e[2], e[3]);
          }
        }
      } else {
        // WGRAD A: 2 M-blocks (64 patch entries = 64 channels of one tap) x 64 pixels; i -> (block i/4, pixel (i%4)*16 + ..)
        const long long lim = min(pixels, L.k_lo + a.k_per_split);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = j * 16 + warp * 2 + rsel;
          long long base; int oy, ox;
          pixel_affine(L.k_lo + (long long)lkb * BK + k, lim, base, oy, ox);
#pragma unroll
          for (int mb = 0; mb < 2; ++mb) {
            const long long kd0 = (long long)L.m_tile * BM + mb * 64;
            if (kd0 < patch) {
              const int tap = (int)(kd0 / C), c0 = (int)(kd0 % C);
              const long long off = tap_off(base, oy, ox, tap / a.KW, tap % a.KW);
              if (off >= 0) v[mb * 4 + j] = __ldg(reinterpret_cast<const float4*>(a.src + off + c0 + 4 * f4));
            }
          }
        }
      }
    };
    auto advance_load = [&]() {
      if (!L.valid) return;
      if (++lkb == L.nkb) {
        lkb = 0;
        decode_tile<MODE>(a, L.t + gridDim.x, total, pixels, L);
        row_coords();
      }
    };
    unsigned it = 0;                 // global k-block counter -> stage / phase
    auto store = [&](const float4 (&v)[8]) {
      const int s = it % STAGES;
      mbar_wait(bar_empty + s * 8, ((it / STAGES) & 1) ^ 1);
      const uint32_t tileA = sA + s * A_STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (a.debug & 16) break;
        if (MODE != WGRAD) {
          const int row = warp * 16 + i * 2 + rsel;
          sts64(tileA + row * 128 + (((f4 >> 1) ^ (row & 7)) << 4) + (f4 & 1) * 8, v[i]);
        } else {
          const int mb = i >> 2, k = (i & 3) * 16 + warp * 2 + rsel;
          sts64(tileA + mb * 8192 + k * 128 + (((f4 >> 1) ^ (k & 7)) << 4) + (f4 & 1) * 8, v[i]);
        }
      }
      // NOTE: the generic->async proxy fence is issued by the MMA thread after it has acquired the stage
      // (see below).  A writer-side fence.proxy.async lowers to MEMBAR.ALL.CTA, which would drain this
      // thread's in-flight prefetch loads and serialise the pipeline on HBM latency.
      if (tid == 0 && !(a.debug & 2)) {
        // B tile image: weights [n_tile][kb], or dy [global k-block][n_tile] for WGRAD
        const size_t tile_idx = MODE == WGRAD ? ((size_t)(S.k_lo / BK + skb) * a.n_tiles + S.n_tile)
                                              : ((size_t)S.n_tile * S.nkb + skb);
        mbar_arrive_expect_tx(bar_full + s * 8, B_STAGE_BYTES);
        tma_bulk_g2s(sB + s * B_STAGE_BYTES, reinterpret_cast<const uint8_t*>(a.wt) + tile_idx * B_STAGE_BYTES,
                     B_STAGE_BYTES, bar_full + s * 8);
      } else {
        mbar_arrive(bar_full + s * 8);
      }
      ++it;
      if (++skb == S.nkb) {
        skb = 0;
        decode_tile<MODE>(a, S.t + gridDim.x, total, pixels, S);
      }
    };
    float4 va[8], vb[8];
    load(va); advance_load();
    load(vb); advance_load();
    while (S.valid) {
      store(va);
      load(va); advance_load();
      if (!S.valid) break;
      store(vb);
      load(vb); advance_load();
    }
  } else if (warp == MMA_WARP) {
    // ===================================================== MMA issuer (converged warp, one elected lane issues: see conv_tma.cu)
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(BM, BN, MODE == WGRAD);
    Tile T;
    decode_tile<MODE>(a, blockIdx.x, total, pixels, T);
    uint32_t s = 0, ph = 0;
    unsigned j = 0;
    while (T.valid) {
      const uint32_t acc = j & 1;
      mbar_wait(bar_tempty + acc * 8, ((j >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < T.nkb; ++kb) {
        mbar_wait(bar_full + s * 8, ph);                    // acquire: all producers' st.shared of this stage
        if (!(a.debug & 8)) fence_proxy_async();            // ... made visible to the async proxy (tcgen05.mma reads)
        tc_fence_after();
        const uint32_t tA = sA + s * A_STAGE_BYTES, tB = sB + s * B_STAGE_BYTES;
        if (leader) {
          if (MODE == WGRAD) {     // 16 pixels = two 8-row groups = 2048 B per K step; 64-element MN blocks 8192 B apart
            const uint64_t dA = make_desc_mn(tA, 8192), dB = make_desc_mn(tB, 8192);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, dA + 128 * k, dB + 128 * k, idesc, (kb | k) != 0);
          } else {                 // +32 bytes per 16-element K step inside the swizzle atom
            const uint64_t dA = make_desc(tA), dB = make_desc(tB);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16(tmem_d, dA + 2 * k, dB + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(bar_empty + s * 8);          // stage is free once these MMAs retire
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      if (leader) umma_commit(bar_tfull + acc * 8);          // accumulator complete
      __syncwarp();
      ++j;
      decode_tile<MODE>(a, T.t + gridDim.x, total, pixels, T);
    }
  } else {
    // ===================================================== epilogue (TMEM -> registers -> smem transpose -> global)
    // tcgen05.ld hands every thread one accumulator ROW; stores of that shape are 32 scattered 16-byte pieces per
    // instruction.  Each warp therefore transposes its 32x32 chunk through a private padded smem tile so that 8
    // lanes cover 128 contiguous bytes of one output row (coalesced 16-byte loads/stores, 4 rows per instruction).
    const int quad = warp & 3;                     // tcgen05.ld: warp w may touch lanes 32*(w%4)..+31
    float* stage = reinterpret_cast<float*>(smem + (slot + 16 - base)) + quad * (32 * 36);
    const int sub = lane >> 3, c4 = (lane & 7) * 4;
    Tile T;
    decode_tile<MODE>(a, blockIdx.x, total, pixels, T);
    unsigned j = 0;
    while (T.valid) {
      const uint32_t acc = j & 1;
      mbar_wait(bar_tfull + acc * 8, (j >> 1) & 1);
      tc_fence_after();
      const long long mrow0 = (long long)T.m_tile * BM + quad * 32;
      const int n0 = T.n_tile * BN;
      const long long mlim = MODE == WGRAD ? min(Mg, (long long)a.m_valid) : Mg;
      // destination offset of column 0 of this lane's 8 rows (once per tile; the fused depth_to_space /
      // space_to_depth permutations only change this row term and, for FPROP, a small per-chunk column term)
      long long rowoff[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long m = mrow0 + it * 4 + sub;
        rowoff[it] = m * Ng;
        if (a.perm_r > 0 && m < mlim) {
          if (MODE == FPROP) rowoff[it] = d2s_dest(m, 0, a.OH, a.OW, Ng, a.perm_r);
          if (MODE == DGRAD) rowoff[it] = s2d_dest(m, 0, a.H, a.W, Ng, a.perm_r);
        }
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c0, v);
        const int nb = n0 + c0;
        if (nb >= Ng || (a.debug & 4)) continue;            // warp-uniform
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(stage + lane * 36 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int col = nb + c4;
        bool vec_ok = (Ng & 3) == 0 && col + 4 <= Ng;
        long long coloff = col;
        if (MODE == FPROP && a.perm_r > 0) {      // col = (i*r + j)*C' + c  ->  (i * OW*r + j) * C' + c
          const int r = a.perm_r, Cp = Ng / (r * r), ij = col / Cp, c = col % Cp;
          coloff = ((long long)(ij / r) * a.OW * r + ij % r) * Cp + c;
          vec_ok = vec_ok && (Cp & 3) == 0;
        }
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == FPROP && a.bias != nullptr) {
          if (col < Ng) bias4.x = __ldg(a.bias + col);
          if (col + 1 < Ng) bias4.y = __ldg(a.bias + col + 1);
          if (col + 2 < Ng) bias4.z = __ldg(a.bias + col + 2);
          if (col + 3 < Ng) bias4.w = __ldg(a.bias + col + 3);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + sub;
          const long long m = mrow0 + r;
          float4 q = *reinterpret_cast<const float4*>(stage + r * 36 + c4);
          if (m >= mlim || col >= Ng) continue;
          float* o = a.out + rowoff[it] + coloff;
          float e[4] = {q.x, q.y, q.z, q.w};
          if (MODE == FPROP) {
            e[0] = act_apply(e[0] + bias4.x, a.act); e[1] = act_apply(e[1] + bias4.y, a.act);
            e[2] = act_apply(e[2] + bias4.z, a.act); e[3] = act_apply(e[3] + bias4.w, a.act);
          } else if (MODE == DGRAD) {
            if (vec_ok) {
              if (a.aux != nullptr) {
                const float4 ax = __ldg(reinterpret_cast<const float4*>(a.aux + m * Ng + col));
                e[0] *= act_grad_from_out(ax.x, a.act); e[1] *= act_grad_from_out(ax.y, a.act);
                e[2] *= act_grad_from_out(ax.z, a.act); e[3] *= act_grad_from_out(ax.w, a.act);
              }
              if (a.accumulate) {
                const float4 old = *reinterpret_cast<const float4*>(o);
                e[0] += old.x; e[1] += old.y; e[2] += old.z; e[3] += old.w;
              }
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t)
                if (col + t < Ng) {
                  if (a.aux != nullptr) e[t] *= act_grad_from_out(__ldg(a.aux + m * Ng + col + t), a.act);
                  if (a.accumulate) e[t] += o[t];
                }
            }
          }
          if (MODE == WGRAD) {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (col + t < Ng) atomicAdd(o + t, e[t]);
          } else if (vec_ok) {
            *reinterpret_cast<float4*>(o) = make_float4(e[0], e[1], e[2], e[3]);
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (col + t < Ng) {
                if (MODE == FPROP && a.perm_r > 0) a.out[d2s_dest(m, col + t, a.OH, a.OW, Ng, a.perm_r)] = e[t];
                else o[t] = e[t];
              }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + acc * 8);           // this thread's TMEM reads of the buffer are complete
      ++j;
      decode_tile<MODE>(a, T.t + gridDim.x, total, pixels, T);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tmem_base, TMEM_COLS);
}

int pick_bn(int Ng) { return Ng <= 32 ? 32 : (Ng <= 64 ? 64 : (Ng <= 128 ? 128 : 256)); }
int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------ operand repack (fp32 -> bf16 tile images)
// Weights: logical K-major matrix  FPROP: Bt[n][k] = w[k*Cout + n];  DGRAD: Bt[ci][tap*Cout + co] = w[(tap*Cin + ci)*Cout + co]
// stored as consecutive [BN x 64] tiles (tile index = n_tile * num_kb + kb), each already in the SWIZZLE_128B
// shared-memory image (row rr at rr*128, 16-byte chunk j at ((j ^ (rr & 7)) << 4)), so one bulk TMA copy stages it.
__global__ void pack_weights_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ img, int mode, int taps, int Cin,
                                    int Cout, int bn, int num_kb, int n_tiles, TapMap tm) {
  const int N = mode == FPROP ? Cout : Cin;
  const int Cg = mode == FPROP ? Cin : Cout;        // channels of the gathered (A) tensor
  const int K = taps * Cg;
  const long long total = (long long)n_tiles * num_kb * bn * BK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kk = (int)(i % BK);
    long long r = i / BK;
    const int rr = (int)(r % bn); r /= bn;
    const int kb = (int)(r % num_kb);
    const int nt = (int)(r / num_kb);
    const int n = nt * bn + rr;
    int k = kb * BK + kk;
    if (Cg % BK == 0) {            // channel-chunk-major K order of the affine fast path (see the producer)
      const int tap = kb % taps, c = (kb / taps) * BK + kk;
      k = tap * Cg + c;
    }
    float v = 0.f;
    if (n < N && k < K) {
      if (mode == FPROP) v = w[(long long)k * Cout + n];
      else {
        int tap = k / Cout;
        const int co = k % Cout;
        if (tm.nkw > 0) tap = (tm.kh0 + tm.s * (tap / tm.nkw)) * tm.KW + tm.kw0 + tm.s * (tap % tm.nkw);   // tap subset of one parity class
        v = w[((long long)tap * Cin + n) * Cout + co];
      }
    }
    const long long tile = (long long)nt * num_kb + kb;
    img[tile * bn * BK + rr * BK + ((((kk >> 3) ^ (rr & 7))) << 3) + (kk & 7)] = __float2bfloat16_rn(v);
  }
}

// dy [P, Cout] fp32 -> MN-major bf16 tile images for WGRAD's B operand: image index = kblock * n_tiles + n_tile, each
// [bn/64 blocks][64 pixels][64 channels] with the 16-byte chunk swizzle (chunk ^ (pixel & 7)).  One thread = one chunk.
__global__ void pack_dy_kernel(const float* __restrict__ dy, uint4* __restrict__ img, long long P, int Cout, int bn, int n_tiles,
                               long long n_kblocks) {
  const int chunks_n = n_tiles * bn / 8;
  const long long total = n_kblocks * BK * chunks_n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cn = (int)(i % chunks_n);          // 8-channel chunk along the (padded) Cout axis
    const long long p = i / chunks_n;            // pixel (padded to a multiple of 64)
    const int co = cn * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (p < P && co < Cout) {
      const float* g = dy + p * Cout + co;
      if ((Cout & 3) == 0 && co + 8 <= Cout) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(g)), v1 = __ldg(reinterpret_cast<const float4*>(g + 4));
        f[0] = v0.x; f[1] = v0.y; f[2] = v0.z; f[3] = v0.w; f[4] = v1.x; f[5] = v1.y; f[6] = v1.z; f[7] = v1.w;
      } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (co + q < Cout) f[q] = __ldg(g + q);
      }
    }
    const long long kbg = p / BK;
    const int k = (int)(p % BK);
    const int nt = co / bn, nb = (co % bn) / 64, chunk = (co % 64) / 8;
    const long long dst = (((kbg * n_tiles + nt) * (bn / 64) + nb) * 64 + k) * 8 + (chunk ^ (k & 7));
    img[dst] = pack8(f);
  }
}

template <int MODE>
static int launch(TcArgs& a, long long Mg, int Ng, int splits, cudaStream_t st) {
  int bn = pick_bn(Ng);
  if (MODE == WGRAD && bn < 64) bn = 64;       // MN-major blocks are 64 elements wide
  a.m_tiles = (int)ceil_div64(Mg, BM);
  a.n_tiles = ceil_div(Ng, bn);
  a.splits = splits;
  static const int dbg = getenv("LADDER_TC_DEBUG") ? atoi(getenv("LADDER_TC_DEBUG")) : 0;
  a.debug = dbg;
  const long long total = (long long)a.m_tiles * a.n_tiles * splits;
  const int sms = num_sms();
  const unsigned grid = (unsigned)(total < sms ? total : sms);
  auto go = [&](auto kern, int BNv) {
    const size_t smem = (size_t)stages_for(BNv) * (A_STAGE_BYTES + BNv * BK * 2) + 1024 + 256 + 4 * 32 * 36 * sizeof(float);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, NTHREADS, smem, st>>>(a);
  };
  switch (bn) {
    case 32: if constexpr (MODE != WGRAD) { go(tc_kernel<MODE, 32>, 32); } break;
    case 64: go(tc_kernel<MODE, 64>, 64); break;
    case 128: go(tc_kernel<MODE, 128>, 128); break;
    default: go(tc_kernel<MODE, 256>, 256); break;
  }
  return check_launch("tcgen05 conv kernel");
}

}  // namespace tc
}  // namespace ladder

using namespace ladder;
using namespace ladder::tc;

namespace ladder { namespace tc {
size_t pack_bytes(int N, int K, int bn) {
  if (bn <= 0) bn = pick_bn(N);
  return (size_t)ceil_div(N, bn) * ceil_div(K, BK) * bn * BK * 2;
}
} }
static size_t pack_dy_bytes(long long P, int Cout) {
  int bn = pick_bn(Cout);
  if (bn < 64) bn = 64;
  return (size_t)ceil_div64(P, BK) * BK * ceil_div(Cout, bn) * bn * 2;
}

namespace ladder { namespace tc {
int pack(const float* w, void* ws, size_t ws_bytes, int mode, int taps, int Cin, int Cout, cudaStream_t st, int bn, TapMap tm) {
  const int N = mode == FPROP ? Cout : Cin, K = taps * (mode == FPROP ? Cin : Cout);
  if (bn <= 0) bn = pick_bn(N);
  const size_t need = pack_bytes(N, K, bn);
  if (ws == nullptr || ws_bytes < need) return fail(LADDER_ERR_WORKSPACE, "conv2d_tc: workspace %zu < %zu bytes", ws_bytes, need);
  if ((uintptr_t)ws & 127) return fail(LADDER_ERR_ARG, "conv2d_tc: workspace must be 128-byte aligned");
  const int num_kb = ceil_div(K, BK), n_tiles = ceil_div(N, bn);
  long long blocks = ceil_div64((long long)n_tiles * num_kb * bn * BK, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_weights_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(ws), mode, taps, Cin, Cout, bn, num_kb, n_tiles, tm);
  return check_launch("conv2d_tc weight pack");
}

// Every weight image of one optimiser group in ONE launch (run at the start of each sub-step, so the forward / backward
// GEMMs never repack).  Work unit = one [min(bn,64) rows x 64 k] block of one image, transposed through shared memory
// so that both the fp32 weight reads (HWIO: contiguous along Cout) and the bf16 image writes (16-byte swizzled chunks)
// are coalesced.  desc[i].first = number of units before entry i.
struct PackDesc { long long w_off, img_off, first; int mode, taps, Cin, Cout, bn, pad; };
__global__ void __launch_bounds__(256) pack_multi_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ images,
                                                         const PackDesc* __restrict__ desc, int n, long long total_units) {
  __shared__ float tile[64][65];                   // [kk][row]
  const int tid = threadIdx.x;
  for (long long u = blockIdx.x; u < total_units; u += gridDim.x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {                              // last entry with first <= u
      const int mid = (lo + hi + 1) >> 1;
      if (desc[mid].first <= u) lo = mid; else hi = mid - 1;
    }
    const PackDesc d = desc[lo];
    const long long i = u - d.first;
    const int N = d.mode == FPROP ? d.Cout : d.Cin, Cg = d.mode == FPROP ? d.Cin : d.Cout, K = d.taps * Cg;
    const int num_kb = ceil_div(K, BK), sr = d.bn < 64 ? d.bn : 64, subs = d.bn / sr;
    const int sb = (int)(i % subs);
    const long long r = i / subs;
    const int kb = (int)(r % num_kb), nt = (int)(r / num_kb);
    const int n0 = nt * d.bn + sb * sr;
    const bool chunked = Cg % BK == 0;             // channel-chunk-major K order of the TMA / affine producers
    const int ktap = kb % d.taps, kc0 = (kb / d.taps) * BK;
    const float* w = params + d.w_off;
    if (d.mode == FPROP) {                         // w[k][n]: lanes along n
      const int nn = tid & 63;
      for (int kk = tid >> 6; kk < BK; kk += 4) {
        const int k = chunked ? ktap * Cg + kc0 + kk : kb * BK + kk;
        float v = 0.f;
        if (nn < sr && n0 + nn < N && k < K) v = __ldg(w + (long long)k * d.Cout + n0 + nn);
        tile[kk][nn] = v;
      }
    } else {                                       // w[tap][n][co]: lanes along k = (tap, co)
      const int kk = tid & 63;
      const int k = chunked ? ktap * Cg + kc0 + kk : kb * BK + kk;
      const int tap = k / d.Cout, co = k - tap * d.Cout;
      for (int nn = tid >> 6; nn < sr; nn += 4) {
        float v = 0.f;
        if (n0 + nn < N && k < K) v = __ldg(w + ((long long)tap * d.Cin + n0 + nn) * d.Cout + co);
        tile[kk][nn] = v;
      }
    }
    __syncthreads();
    __nv_bfloat16* img = images + d.img_off + ((long long)nt * num_kb + kb) * d.bn * BK;
    for (int c = tid; c < sr * 8; c += 256) {      // 16-byte chunks: row rr, chunk ch
      const int rr = c >> 3, ch = c & 7, row = sb * sr + rr;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tile[ch * 8 + j][rr];
      *reinterpret_cast<uint4*>(img + row * BK + ((ch ^ (row & 7)) << 3)) = pack8(f);
    }
    __syncthreads();
  }
}
} }

extern "C" {

/* desc_dev: n PackDesc records; total = number of [min(bn,64) x 64] work units over all entries */
int ladder_pack_weights_multi(const float* params, void* images, const void* desc_dev, int n, long long total,
                              cudaStream_t stream) {
  LADDER_REQUIRE(params && images && desc_dev && n > 0 && total > 0, "pack_weights_multi: bad arguments");
  long long blocks = total;                        // units
  if (blocks > 148 * 8) blocks = 148 * 8;
  pack_multi_kernel<<<(unsigned)blocks, 256, 0, stream>>>(params, static_cast<__nv_bfloat16*>(images),
                                                          static_cast<const PackDesc*>(desc_dev), n, total);
  return check_launch("pack_weights_multi");
}

size_t ladder_conv2d_tc_workspace_bytes(int B, int H, int W, int Cin, int KH, int KW, int Cout) {
  const int OHmax = H, OWmax = W;                       // OH*OW <= H*W for every geometry the library accepts
  const size_t f = pack_bytes(Cout, KH * KW * Cin, 0), d = pack_bytes(Cin, KH * KW * Cout, 0);
  const size_t g = Cin % BK == 0 ? pack_dy_bytes((long long)B * OHmax * OWmax, Cout) : 0;
  size_t m = f > d ? f : d;
  if (g > m) m = g;
  return m + 256;
}

/* 1 if wgrad of this geometry runs on the tensor cores (64-channel-aligned input), else 0 (use the fp32 kernel) */
int ladder_conv2d_wgrad_tc_supported(int Cin, int Cout) { (void)Cout; return Cin % BK == 0 ? 1 : 0; }

int ladder_conv2d_fprop_tc(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int Cin,
                           int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act,
                           int out_d2s, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(out_d2s == 0 || (out_d2s > 0 && Cout % (out_d2s * out_d2s) == 0),
                 "conv2d_fprop_tc: depth_to_space(%d) output needs Cout %% r^2 == 0", out_d2s);
  LADDER_REQUIRE(x && w && y && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
                 "conv2d_fprop_tc: bad arguments");
  LADDER_REQUIRE((long long)B * OH * OW < (1LL << 31) && (long long)B * H * W < (1LL << 31), "conv2d_fprop_tc: too many pixels");
  int rc = pack(w, workspace, workspace_bytes, FPROP, KH * KW, Cin, Cout, stream, 0);
  if (rc) return rc;
  TcArgs a{x, nullptr, static_cast<const __nv_bfloat16*>(workspace), bias, nullptr, y, B, H, W, Cin, KH, KW, Cout, stride,
           pad_t, pad_l, OH, OW, act, 0, 0, round_up(KH * KW * Cin, BK), 0, 0, 0, 0, out_d2s};
  return launch<FPROP>(a, (long long)B * OH * OW, Cout, 1, stream);
}

int ladder_conv2d_dgrad_tc(const float* dy, const float* w, const float* act_out, float* dx, int B, int H, int W, int Cin,
                           int KH, int KW, int Cout, int stride, int pad_t, int pad_l, int OH, int OW, int act,
                           int accumulate, int out_s2d, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  LADDER_REQUIRE(out_s2d == 0 || (out_s2d > 0 && H % out_s2d == 0 && W % out_s2d == 0),
                 "conv2d_dgrad_tc: space_to_depth(%d) output needs H, W divisible by r", out_s2d);
  LADDER_REQUIRE(dy && w && dx && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
                 "conv2d_dgrad_tc: bad arguments");
  LADDER_REQUIRE((long long)B * OH * OW < (1LL << 31) && (long long)B * H * W < (1LL << 31), "conv2d_dgrad_tc: too many pixels");
  int rc = pack(w, workspace, workspace_bytes, DGRAD, KH * KW, Cin, Cout, stream, 0);
  if (rc) return rc;
  TcArgs a{dy, nullptr, static_cast<const __nv_bfloat16*>(workspace), nullptr, act_out, dx, B, H, W, Cin, KH, KW, Cout, stride,
           pad_t, pad_l, OH, OW, act, accumulate, 0, round_up(KH * KW * Cout, BK), 0, 0, 0, 0, out_s2d};
  return launch<DGRAD>(a, (long long)B * H * W, Cin, 1, stream);
}

// dw is overwritten; the bias gradient is ladder_colsum(dy).  Requires Cin % 64 == 0 (ladder_conv2d_wgrad_tc_supported).
int ladder_conv2d_wgrad_tc(const float* x, const float* dy, float* dw, int B, int H, int W, int Cin, int KH, int KW, int Cout,
                           int stride, int pad_t, int pad_l, int OH, int OW, void* workspace, size_t workspace_bytes,
                           cudaStream_t stream) {
  LADDER_REQUIRE(x && dy && dw && B > 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && KH > 0 && KW > 0 && stride > 0 && OH > 0 && OW > 0,
                 "conv2d_wgrad_tc: bad arguments");
  LADDER_REQUIRE(Cin % BK == 0, "conv2d_wgrad_tc: Cin must be a multiple of 64 (got %d)", Cin);
  const int patch = KH * KW * Cin;
  const long long pixels = (long long)B * OH * OW;
  LADDER_REQUIRE(pixels < (1LL << 31) && (long long)B * H * W < (1LL << 31), "conv2d_wgrad_tc: too many pixels");
  const size_t need = pack_dy_bytes(pixels, Cout);
  if (workspace == nullptr || workspace_bytes < need)
    return fail(LADDER_ERR_WORKSPACE, "conv2d_wgrad_tc: workspace %zu < %zu bytes", workspace_bytes, need);
  cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)patch * Cout * sizeof(float), stream);
  if (e != cudaSuccess) return fail(LADDER_ERR_CUDA, "conv2d_wgrad_tc memset: %s", cudaGetErrorString(e));
  int bn = pick_bn(Cout);
  if (bn < 64) bn = 64;
  const int n_tiles = ceil_div(Cout, bn);
  const long long n_kblocks = ceil_div64(pixels, BK);
  {
    const long long chunks = n_kblocks * BK * (n_tiles * bn / 8);
    long long blocks = ceil_div64(chunks, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    pack_dy_kernel<<<(unsigned)blocks, 256, 0, stream>>>(dy, static_cast<uint4*>(workspace), pixels, Cout, bn, n_tiles, n_kblocks);
    int rc = check_launch("conv2d_wgrad_tc dy pack");
    if (rc) return rc;
  }
  const long long tiles = ceil_div64(patch, BM) * n_tiles;
  long long splits = (2LL * num_sms()) / tiles;      // <= 2 full waves of the persistent grid
  const long long max_splits = ceil_div64(pixels, 4 * BK);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = ceil_div64(ceil_div64(pixels, splits), BK) * BK;
  TcArgs a{x, nullptr, static_cast<const __nv_bfloat16*>(workspace), nullptr, nullptr, dw, B, H, W, Cin, KH, KW, Cout, stride,
           pad_t, pad_l, OH, OW, 0, 0, (int)per, 0, patch, 0, 0, 0};
  return launch<WGRAD>(a, patch, Cout, (int)ceil_div64(pixels, per), stream);
}

}  // extern "C"

// K3: bf16 implicit-GEMM convolution on tcgen05 tensor cores, fed by TMA, accumulators in TMEM,
// fused per-channel scale/shift (bias / folded BatchNorm) + residual + activation epilogue.
//
// Covers the conv layer classes of FCN_16_standard (reference modules
// medseg/models/ebm/encoder_decoder.py:19-68, :285-348, :351-415, :418-453, :456-503):
//   3x3 stride-1 pad-1 (TAPS = 9), 1x1 (TAPS = 1), and 3x3 stride-2 pad-1 as "compute at full
//   resolution, keep even pixels" (the MMA is far from the bottleneck on these HBM-bound layers),
//   Cin in {16,32,64,128}, Cout a multiple of 16.
// The 16 -> 16 channel 3x3 layers (stride 1 and 2) are dispatched to K3s (conv_small.cuh, warp-level tensor path) by
// the same entry points: with N = 16 this kernel is bound by its shared-memory operand fetch (see there).
//
// Data layout in HBM ("blocked", C8): activations are bf16 [N][C/8][H][W][8] -- channels in groups of eight,
// each group a dense H x W plane of 16-byte pixels.  This is the layout the tensor core consumes: ONE TMA box
// {8*halo_w, halo_h, C/8, 1} per tile has (halo_w*16)-byte contiguous rows and lands in shared memory as
// [C/8][halo_h][halo_w][8ch], which is exactly the UMMA no-swizzle K-major canonical layout
// (LBO = halo_h*halo_w*16, SBO = halo_w*16).  The nine filter taps are nine descriptor START ADDRESSES into
// the same tile: no im2col copy, every input byte crosses HBM once, and the epilogue's 16-byte stores of
// consecutive pixels coalesce into 128-byte lines.  Weights are pre-packed per N tile as
// [n_tile][tap][Cin/8][NT][8] bf16 (K-major core matrices, ops.pack_conv_weight).
//
// One CTA = one SM, persistent over output tiles of 16 rows x (8*MT) columns of one image:
//   warp 0  : TMA producer  -- one box load per tile (zero-filled halo) into a ring of STAGES buffers
//   warp 1  : MMA issuer    -- one elected thread issues TAPS*Cin/16 tcgen05.mma (M=128, N=NT, K=16) per
//             8-column M tile into one of two TMEM accumulator stages, then tcgen05.commit frees the
//             smem stage and publishes the accumulator.
//   warp 2  : TMEM allocator; warp 3 idle.
//   warps 4-11: epilogue    -- two warps per TMEM sub-partition share a tile's (M tile, 16-column chunk) items:
//             tcgen05.ld 16 columns, residual requested before the TMEM wait, y = act(acc*scale + shift + res*rs + rb),
//             bf16 pack, 16-byte stores (optionally scattered 2x2 for ConvTranspose2d k2 s2), optional per-channel
//             sum / sum-of-squares of the stored values (train-mode BatchNorm statistics without a second pass).
//
// VP (vertically packed 3x3, stride 1): with N = Cout <= 64 every tcgen05.mma re-reads a 4 KB pixel tile from shared
// memory for little math, so the tensor pipe idles behind operand fetch (measured ~52 issue cycles per N = 16 MMA for
// 8 cycles of math).  The three VERTICAL taps of a filter column s are therefore packed into the N dimension:
//   B_s = (W[0][s] | W[1][s] | W[2][s])  ->  N = 3*NT, three MMAs per 16-channel K slice instead of nine,
//   D[(row, px), (r, co)] = sum_ci X[y0-1+row][x0+px+s-1][ci] * W[r][s][co][ci]   (accumulated over s and ci)
// and the epilogue finishes the sum across rows: out[j][px][co] = D[j][0,co] + D[j+1][1,co] + D[j+2][2,co].  A tile is
// 16 INPUT rows (= the M rows of the MMA, no vertical halo) and produces 14 output rows; every feature-map height of
// the 224^2 configuration (224, 112, 56, 28, 14) is a multiple of 14.  Rows meet through a small shared-memory exchange
// between the four epilogue warps that share an item.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "ctl_common.cuh"
#include "ctl_tcgen05.cuh"

namespace ctl {
namespace {

using namespace sm100;

constexpr int kTileH = 16;            // output rows per tile (= 8-row groups of the M=128 MMA: one group per row)
constexpr int kConvThreads = 384;      // warps 0-3: TMA, MMA, TMEM allocator, spare; warps 4-11: epilogue

struct ConvParams {
  int N, H, W;                        // input (= full-resolution output) size
  int Cout;                           // total output channels
  int tiles_x, tiles_y;               // tiles per image
  int64_t num_tiles;                  // N * tiles_y * tiles_x
  const __nv_bfloat16* w_packed;      // [Cout/NT][TAPS][Cin/8][NT][8]
  const float* scale;                 // [Cout] or nullptr (=1)
  const float* shift;                 // [Cout] or nullptr (=0)
  const __nv_bfloat16* res;           // blocked [N,Cout/8,Ho,Wo,8] or nullptr
  const float* res_scale;             // [Cout] or nullptr (=1)
  const float* res_shift;             // [Cout] or nullptr (=0)
  __nv_bfloat16* out;                 // blocked [N,Cout/8,Ho,Wo,8]  (up2x: [N,Cout/32,2H,2W,8])
  int act;                            // ctl_act
  int subsample;                      // 1: Ho=H, Wo=W; 2: keep even (y,x) -> Ho=H/2, Wo=W/2 (3x3 stride-2 pad-1)
  int up2x;                           // 1: ConvTranspose2d(k=2,s=2) scatter: GEMM column n = (dy*2+dx)*Cout/4 + co
  double* stats;                      // [2][Cout] sum / sum of squares of the stored outputs, accumulated into (or nullptr)
  double* sal;                        // latent saliency sums: [N][Cout] (channel mode: sum over pixels) or
                                      // [N][H*W] (spatial mode: sum over channels) of the bf16-rounded outputs, accumulated into
  int sal_mode;                       // ctl_mode
  int no_store;                       // 1: the output tensor is not written (only the saliency sums are wanted)
  int bnb_act;                        // != 0: BatchNorm-backward statistics mode (see conv_epilogue<.., BNB>): `res` is the
                                      // BatchNorm INPUT a of the layer this output gradient flows into, res_scale / res_shift
                                      // its forward affine, bnb_act its activation; stats receives sum dv | sum dv*a
  uint64_t magic_img, magic_x;        // K3s: small_div magics of tiles per image / tiles per row
  int diag;                           // profiling only (env CTL_DIAG_SKIP): 1 no MMA, 2 no TMA loads, 4 no epilogue memory traffic, 8 no epilogue
};

#include "ctl_hmma.cuh"
#include "conv_small.cuh"   // K3s: the 16 -> 16 channel 3x3 layers on the warp-level tensor path

template <int CIN, int NT, int TAPS, int MT, int STAGES, bool VP = false>
struct ConvCfg {
  static_assert(!VP || TAPS == 9, "vertical tap packing is for 3x3 filters");
  static constexpr int kPad = TAPS == 9 ? 1 : 0;
  static constexpr int kOutH = VP ? 14 : kTileH;                  // output rows per tile
  static constexpr int kHaloH = VP ? kTileH : kTileH + 2 * kPad;  // VP: the 16 M rows ARE the input rows y0-1 .. y0+14
  static constexpr int kHaloW = 8 * MT + 2 * kPad;
  static constexpr int kChunkBytes = kHaloH * kHaloW * 16;        // one 8-channel plane of the halo tile
  static constexpr int kChunkStride = kChunkBytes;                // the TMA box is written densely
  static constexpr int kStageBytes = (CIN / 8) * kChunkStride;
  static constexpr int kStageTxBytes = kStageBytes;               // bytes the TMA load of one stage delivers
  static constexpr int kWBytes = TAPS * CIN * NT * 2;
  static constexpr int kAccCols = VP ? 3 * NT : NT;               // TMEM columns of one M tile's accumulator
  // accumulator stages: the MMA issuer (and with it the TMA ring) runs up to kAcc tiles ahead of the epilogue
  static constexpr int kAcc = (512 / (MT * kAccCols)) >= 4 ? 4 : 2;
  static constexpr int kTmemCols = kAcc * MT * kAccCols;
  static constexpr int kTmemAlloc = kTmemCols <= 32 ? 32 : kTmemCols <= 64 ? 64 : kTmemCols <= 128 ? 128
                                    : kTmemCols <= 256 ? 256 : 512;
  // VP: per epilogue half, the (r = 1 | r = 2) partial sums of one item: [2][4 float4][128 pixels] fp32 = 16 KB
  static constexpr int kXBytes = VP ? 2 * 2 * 4 * 128 * 16 : 0;
  // smem carve-up (all offsets multiples of 128)
  static constexpr int kOffW = 0;
  static constexpr int kOffA = (kWBytes + 127) / 128 * 128;
  static constexpr int kOffX = kOffA + STAGES * ((kStageBytes + 127) / 128 * 128);
  static constexpr int kOffVec = kOffX + kXBytes;                                       // 4 x NT floats
  static constexpr int kOffBar = kOffVec + 4 * NT * 4;
  static constexpr int kSmemBytes = kOffBar + 256;
  static_assert(kTmemCols <= 512, "accumulators exceed TMEM");
  static_assert(CIN % 16 == 0 && NT % 16 == 0 && kAccCols <= 256, "UMMA shape");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert((kChunkStride >> 4) < 16384 && kHaloW * 16 < 16384 * 16, "descriptor range");
  static_assert(kHaloW * 2 <= 256 && kHaloH <= 256 && CIN / 8 <= 256, "TMA box dimensions (8-byte elements)");
};

// Epilogue of one warp (warps 4-11; q = warp % 4 is the TMEM sub-partition it may read, half = (warp - 4) / 4).  The
// (M tile, 16-column chunk) items of a tile alternate between the two halves; a warp's items of one tile are handled
// as a batch: every residual pixel of the tile is requested BEFORE the accumulator wait (the addresses do not depend on
// it), the TMEM loads of two items are issued back to back behind one tcgen05.wait::ld, and the accumulator stage is
// handed back to the MMA issuer as soon as the last load has landed in registers -- before the arithmetic and stores.
// BNB (with RES and STATS): the output is an activation gradient dy that feeds the backward of h = act(BN(a)); the
// epilogue reads a at the output's own pixels (through the residual path, current tile only) and accumulates
// sum dv and sum dv*a with dv = dy * act'(a*scale + shift) -- the BatchNorm-backward reduction without a pass of its own.
template <int CIN, int NT, int TAPS, int MT, int STAGES, bool VP, bool RES, bool STATS, bool GEN, bool SAL = false,
          bool BNB = false>
__device__ __forceinline__ void conv_epilogue(const ConvParams& p, const uint32_t tmem_base, const float* sVec,
                                              uint64_t* acc_full, uint64_t* acc_empty, const int n0, float* stat_smem,
                                              float4* xbuf) {
  using Cfg = ConvCfg<CIN, NT, TAPS, MT, STAGES, VP>;
  static_assert(!(VP && GEN), "the generic-geometry epilogue (stride 2 / ConvTranspose) runs on unpacked taps");
  constexpr int kAccStages = Cfg::kAcc;
  constexpr int kChunks = NT / 16;
  constexpr int kItems = MT * kChunks;
  constexpr int kPer = (kItems + 1) / 2;          // items of one warp per tile: half, half + 2, ...
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3;
  const int half = (warp - 4) >> 2;
  const int m = q * 32 + lane;                    // accumulator row = pixel within the 16x8 M tile
  const int py = m >> 3, px = m & 7;
  const int Ho = p.H / p.subsample, Wo = p.W / p.subsample;
  const int64_t plane = (int64_t)Ho * Wo * 8;     // elements per 8-channel plane of the output
  const int act = p.act;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = (int)p.num_tiles;         // < 2^31 (checked on the host)
  const int first_item = CTL_DIAGF(p, 8) ? kItems : half;
  // train-mode BatchNorm statistics of the stored (bf16-rounded) outputs: every thread owns one 16-column chunk
  float st_s[STATS ? 16 : 1], st_q[STATS ? 16 : 1];
#pragma unroll
  for (int i = 0; i < (STATS ? 16 : 1); ++i) { st_s[i] = 0.0f; st_q[i] = 0.0f; }

  int acc = 0;
  uint32_t acc_phase = 0;
  // per-tile geometry of this thread's pixel
  struct Geo { int img, y, xb; bool y_ok; int64_t pix; };
  auto geometry = [&](int t) -> Geo {
    Geo g;
    g.img = t / tiles_per_img;
    const int rem = t - g.img * tiles_per_img;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    g.y = ty * Cfg::kOutH + py;
    g.xb = tx * (8 * MT) + px;
    // plain geometry (stride 1, no ConvTranspose scatter): one 64-bit base per tile, items differ by constants
    g.y_ok = py < Cfg::kOutH && g.y < p.H && !CTL_DIAGF(p, 4);
    g.pix = ((int64_t)g.img * (p.Cout >> 3) + (n0 >> 3)) * plane + ((int64_t)g.y * p.W + g.xb) * 8;
    return g;
  };
  // element offset of plane h8 of item u (the residual shares the output's geometry)
  auto offset_of = [&](const Geo& g, int u, int h8, bool& valid) -> int64_t {
    const int item = first_item + 2 * u;
    const int mt = item / kChunks, c0 = (item - mt * kChunks) * 16;
    if (!GEN) {
      valid = g.y_ok && g.xb + mt * 8 < p.W;
      return g.pix + mt * 64 + (int64_t)((c0 >> 3) + h8) * plane;
    }
    const int y = g.y, x = g.xb + mt * 8;
    valid = y < p.H && x < p.W && !CTL_DIAGF(p, 4);
    const int n = n0 + c0 + 8 * h8;              // GEMM column of this plane's first channel
    if (p.up2x) {
      const int cq = p.Cout >> 2, qd = n / cq, co = n - qd * cq;       // n = (dy*2+dx)*Cout/4 + co
      const int y2 = 2 * y + (qd >> 1), x2 = 2 * x + (qd & 1);
      return ((int64_t)g.img * (cq >> 3) + (co >> 3)) * (plane * 4) + ((int64_t)y2 * (2 * Wo) + x2) * 8;
    }
    int yo = y, xo = x;
    if (p.subsample == 2) { valid = valid && !((y | x) & 1); yo = y >> 1; xo = x >> 1; }
    return ((int64_t)g.img * (p.Cout >> 3) + (n >> 3)) * plane + ((int64_t)yo * Wo + xo) * 8;
  };
  const bool has_res = RES && (!GEN || p.res != nullptr);     // the GEN variant serves both
  // the residual pixels of a tile are requested ONE TILE AHEAD (the addresses do not depend on the accumulator): two
  // tiles' worth of 16-byte loads per thread are in flight while the current tile is converted and stored
  auto request_residual = [&](const Geo& g, uint4 (&dst)[RES ? kPer : 1][2]) {
#pragma unroll
    for (int u = 0; u < (RES ? kPer : 1); ++u) {
#pragma unroll
      for (int h8 = 0; h8 < 2; ++h8) {
        dst[u][h8] = make_uint4(0, 0, 0, 0);
        bool valid;
        const int64_t off = offset_of(g, u, h8, valid);
        if (first_item + 2 * u < kItems && valid) dst[u][h8] = __ldg(reinterpret_cast<const uint4*>(p.res + off));
      }
    }
  };

  uint4 rr[RES ? kPer : 1][2], rr_next[(RES && !BNB) ? kPer : 1][2];
  Geo geo = geometry(blockIdx.x);
  if (has_res && !BNB) request_residual(geo, rr);
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    const int t_next = t + gridDim.x;
    Geo geo_next = geo;
    if (t_next < num_tiles) {
      geo_next = geometry(t_next);
      if constexpr (!BNB) {
        if (has_res) request_residual(geo_next, rr_next);
      }
    }
    if constexpr (BNB) request_residual(geo, rr);       // the 32 running sums leave no room for a second tile in flight

    mbar_wait(&acc_full[acc], acc_phase);
    tc_fence_after();
#pragma unroll
    for (int u0 = 0; u0 < kPer; u0 += 2) {
      uint32_t v[2][16];
      if constexpr (!VP) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int item = first_item + 2 * (u0 + b);                       // warp-uniform
          if (u0 + b < kPer && item < kItems) {
            const int mt = item / kChunks, c0 = (item - mt * kChunks) * 16;
            tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + mt) * NT + c0), v[b]);
          }
        }
        tmem_ld_wait();
        if (u0 + 2 >= kPer) {            // the tile's accumulator is in registers: release the TMEM stage now
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
      } else {
        // vertically packed taps: three column groups per item; this thread's D row (py) holds the r = 0 term of
        // output row py, the r = 1 term of output row py - 1 and the r = 2 term of output row py - 2.  The r = 1 / 2
        // groups go through a shared-memory exchange between the four warps (TMEM sub-partitions) of this half.
        float4* xh = xbuf + half * (2 * 4 * 128);
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int item = first_item + 2 * (u0 + b);                       // warp-uniform
          if (u0 + b < kPer && item < kItems) {
            const int mt = item / kChunks, c0 = (item - mt * kChunks) * 16;
            const uint32_t col = (uint32_t)((acc * MT + mt) * Cfg::kAccCols + c0);
            const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16) + col;
            uint32_t v1[16], v2[16];
            tmem_ld_32x32b_x16(tl, v[b]);
            tmem_ld_32x32b_x16(tl + NT, v1);
            tmem_ld_32x32b_x16(tl + 2 * NT, v2);
            tmem_ld_wait();
            const bool last_item = (u0 + b + 1 >= kPer) || (first_item + 2 * (u0 + b + 1) >= kItems);
            if (last_item) {             // everything of this tile's accumulator this warp needs is in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&acc_empty[acc]);
            }
#pragma unroll
            for (int f = 0; f < 4; ++f) {
              xh[(0 * 4 + f) * 128 + m] = make_float4(__uint_as_float(v1[4 * f]), __uint_as_float(v1[4 * f + 1]),
                                                      __uint_as_float(v1[4 * f + 2]), __uint_as_float(v1[4 * f + 3]));
              xh[(1 * 4 + f) * 128 + m] = make_float4(__uint_as_float(v2[4 * f]), __uint_as_float(v2[4 * f + 1]),
                                                      __uint_as_float(v2[4 * f + 2]), __uint_as_float(v2[4 * f + 3]));
            }
            asm volatile("bar.sync %0, 128;" ::"r"(3 + half) : "memory");
            if (py < Cfg::kOutH) {
#pragma unroll
              for (int f = 0; f < 4; ++f) {
                const float4 a1 = xh[(0 * 4 + f) * 128 + m + 8];       // D row py + 1, r = 1
                const float4 a2 = xh[(1 * 4 + f) * 128 + m + 16];      // D row py + 2, r = 2
                v[b][4 * f] = __float_as_uint(__uint_as_float(v[b][4 * f]) + a1.x + a2.x);
                v[b][4 * f + 1] = __float_as_uint(__uint_as_float(v[b][4 * f + 1]) + a1.y + a2.y);
                v[b][4 * f + 2] = __float_as_uint(__uint_as_float(v[b][4 * f + 2]) + a1.z + a2.z);
                v[b][4 * f + 3] = __float_as_uint(__uint_as_float(v[b][4 * f + 3]) + a1.w + a2.w);
              }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(3 + half) : "memory");   // the buffer is rewritten by the next item
          }
        }
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int u = u0 + b;
        const int item = first_item + 2 * u;
        if (u < kPer && item < kItems) {
          const int c0 = (item % kChunks) * 16;
          bool valid;
          int64_t off[2];
          off[0] = offset_of(geo, u, 0, valid);
          off[1] = offset_of(geo, u, 1, valid);
          float sv[SAL ? 16 : 1];               // this pixel's 16 bf16-rounded outputs (zero for pixels outside the image)
#pragma unroll
          for (int i = 0; i < (SAL ? 16 : 1); ++i) sv[i] = 0.0f;
          if (valid) {
            float f[16];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 sc = *reinterpret_cast<const float4*>(sVec + c0 + 4 * i4);
              const float4 sh = *reinterpret_cast<const float4*>(sVec + NT + c0 + 4 * i4);
              f[4 * i4] = fmaf(__uint_as_float(v[b][4 * i4]), sc.x, sh.x);
              f[4 * i4 + 1] = fmaf(__uint_as_float(v[b][4 * i4 + 1]), sc.y, sh.y);
              f[4 * i4 + 2] = fmaf(__uint_as_float(v[b][4 * i4 + 2]), sc.z, sh.z);
              f[4 * i4 + 3] = fmaf(__uint_as_float(v[b][4 * i4 + 3]), sc.w, sh.w);
            }
            if (has_res && !BNB) {
#pragma unroll
              for (int h8 = 0; h8 < 2; ++h8) {
                const uint4 r4 = rr[RES ? u : 0][h8];
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int c = 8 * h8 + 2 * i;
                  f[c] += fmaf(__uint_as_float(rw[i] << 16), sVec[2 * NT + c0 + c], sVec[3 * NT + c0 + c]);
                  f[c + 1] += fmaf(__uint_as_float(rw[i] & 0xffff0000u), sVec[2 * NT + c0 + c + 1], sVec[3 * NT + c0 + c + 1]);
                }
              }
            }
            if (act == CTL_ACT_LRELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.2f * f[i]);
            } else if (act == CTL_ACT_RELU) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = fmaxf(f[i], 0.0f);
            } else if (act == CTL_ACT_SIGMOID) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = 1.0f / (1.0f + __expf(-f[i]));
            }
#pragma unroll
            for (int h8 = 0; h8 < 2; ++h8) {
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 hh = __floats2bfloat162_rn(f[8 * h8 + 2 * i], f[8 * h8 + 2 * i + 1]);
                o[i] = *reinterpret_cast<const uint32_t*>(&hh);
                if (STATS && !BNB) {
                  const float lo = __uint_as_float(o[i] << 16), hi = __uint_as_float(o[i] & 0xffff0000u);
                  const int c = 8 * h8 + 2 * i;
                  st_s[STATS ? c : 0] += lo;      st_q[STATS ? c : 0] = fmaf(lo, lo, st_q[STATS ? c : 0]);
                  st_s[STATS ? c + 1 : 0] += hi;  st_q[STATS ? c + 1 : 0] = fmaf(hi, hi, st_q[STATS ? c + 1 : 0]);
                }
                if (STATS && BNB) {
                  // dv = dy * act'(a*scale + shift) on the STORED (bf16) dy, as the stand-alone reduction reads it
                  const uint4 a4 = rr[RES ? u : 0][h8];
                  const uint32_t aw = i == 0 ? a4.x : i == 1 ? a4.y : i == 2 ? a4.z : a4.w;
                  const int c = 8 * h8 + 2 * i;
                  const float a_lo = __uint_as_float(aw << 16), a_hi = __uint_as_float(aw & 0xffff0000u);
                  const float p_lo = fmaf(a_lo, sVec[2 * NT + c0 + c], sVec[3 * NT + c0 + c]);
                  const float p_hi = fmaf(a_hi, sVec[2 * NT + c0 + c + 1], sVec[3 * NT + c0 + c + 1]);
                  const float neg = p.bnb_act == CTL_ACT_LRELU ? 0.2f : (p.bnb_act == CTL_ACT_RELU ? 0.0f : 1.0f);
                  const float lo = __uint_as_float(o[i] << 16) * (p_lo > 0.0f ? 1.0f : neg);
                  const float hi = __uint_as_float(o[i] & 0xffff0000u) * (p_hi > 0.0f ? 1.0f : neg);
                  st_s[STATS ? c : 0] += lo;      st_q[STATS ? c : 0] = fmaf(lo, a_lo, st_q[STATS ? c : 0]);
                  st_s[STATS ? c + 1 : 0] += hi;  st_q[STATS ? c + 1 : 0] = fmaf(hi, a_hi, st_q[STATS ? c + 1 : 0]);
                }
              }
              if (SAL) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  sv[SAL ? 8 * h8 + 2 * i : 0] = __uint_as_float(o[i] << 16);
                  sv[SAL ? 8 * h8 + 2 * i + 1 : 0] = __uint_as_float(o[i] & 0xffff0000u);
                }
              }
              if (!SAL || !p.no_store) *reinterpret_cast<uint4*>(p.out + off[h8]) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          if (SAL) {
            // fp64 sums of bf16 values: exact (hence order-independent) while the dynamic range of one sample's
            // gradient stays below 2^37 -- the same numbers K1 would sum from the materialised tensor
            const int mt = item / kChunks;
            if (p.sal_mode == CTL_MODE_CHANNEL) {
              // the warp's 32 pixels belong to one image: reduce over them, one atomic per channel
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                double t = (double)sv[SAL ? i : 0];
#pragma unroll
                for (int o2 = 16; o2 > 0; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
                if (lane == i) atomicAdd(p.sal + (int64_t)geo.img * p.Cout + n0 + c0 + i, t);
              }
            } else if (valid) {
              double t = 0.0;
#pragma unroll
              for (int i = 0; i < 16; ++i) t += (double)sv[SAL ? i : 0];
              atomicAdd(p.sal + (int64_t)geo.img * ((int64_t)p.H * p.W) + (int64_t)geo.y * p.W + geo.xb + mt * 8, t);
            }
          }
        }
      }
    }
    if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    geo = geo_next;
    if (RES && !BNB) {
#pragma unroll
      for (int u = 0; u < kPer; ++u) { rr[u][0] = rr_next[u][0]; rr[u][1] = rr_next[u][1]; }
    }
  }
  if (STATS) {
    // kChunks <= 2 (checked on the host): this thread's chunk is fixed -- chunk `half` when there are two, else 0.
    // Warp shuffle reduce -> shared memory (the operand ring is idle: every MMA that read it has completed, or this
    // warp could not have seen the last accumulator) -> ONE fp64 atomic per channel statistic and CTA: the per-warp
    // form issued 8x as many onto the same 2*Cout addresses from all CTAs at once and cost ~30 us per launch.
    const int c0 = (kChunks == 2 ? half : 0) * 16;
    const bool owner = kItems > half;          // MT == 1 && kChunks == 1: the second half never had an item
    float* red = stat_smem + (warp - 4) * 32;  // [8 warps][sum x16 | sumsq x16] of this warp's chunk
#pragma unroll
    for (int i = 0; i < (STATS ? 16 : 1); ++i) {
      float s1 = st_s[i], s2 = st_q[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) { red[i] = owner ? s1 : 0.0f; red[16 + i] = owner ? s2 : 0.0f; }
    }
    asm volatile("bar.sync 2, 256;" ::: "memory");          // the eight epilogue warps
    const int tid = threadIdx.x - 128;
    if (tid < 2 * NT) {
      // column j of statistic `which`: chunk j / 16 lives in the warps of half (kChunks == 2 ? j / 16 : both halves)
      const int which = tid / NT, j = tid - which * NT;
      const int chunk = j >> 4, i = j & 15;
      double total = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const int w_half = w >> 2;
        const bool has = kChunks == 2 ? (w_half == chunk) : true;
        if (has) total += (double)stat_smem[w * 32 + which * 16 + i];
      }
      atomicAdd(p.stats + which * p.Cout + n0 + j, total);
    }
    (void)c0;
  }
}

template <int CIN, int NT, int TAPS, int MT, int STAGES, bool VP>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap, const ConvParams p) {
  pdl_entry();
  using Cfg = ConvCfg<CIN, NT, TAPS, MT, STAGES, VP>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW = smem + Cfg::kOffW;
  uint8_t* sA = smem + Cfg::kOffA;
  float* sVec = reinterpret_cast<float*>(smem + Cfg::kOffVec);     // scale | shift | res_scale | res_shift
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* full = bars;                        // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + STAGES;              // [STAGES]  MMA -> TMA
  constexpr int kAccStages = Cfg::kAcc;
  uint64_t* acc_full = bars + 2 * STAGES;                    // [kAcc]  MMA -> epilogue
  uint64_t* acc_empty = bars + 2 * STAGES + kAccStages;      // [kAcc]  epilogue -> MMA
  uint64_t* w_full = bars + 2 * STAGES + 2 * kAccStages;     // [1]     weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * kAccStages + 1);
  constexpr int kStageStride = (Cfg::kStageBytes + 127) / 128 * 128;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;                // which NT-wide slice of the output channels
  const int n0 = n_tile * NT;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < kAccStages; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    mbar_init(w_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemAlloc>(tmem_slot);
  for (int i = threadIdx.x; i < NT; i += kConvThreads) {
    sVec[i] = p.scale ? p.scale[n0 + i] : 1.0f;
    sVec[NT + i] = p.shift ? p.shift[n0 + i] : 0.0f;
    sVec[2 * NT + i] = p.res_scale ? p.res_scale[n0 + i] : 1.0f;
    sVec[3 * NT + i] = p.res_shift ? p.res_shift[n0 + i] : 0.0f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================================================= TMA producer
    // elect.sync (not `lane == 0`): ptxas then knows ONE thread runs the block, keeps descriptors / addresses in
    // uniform registers and issues UTCHMMA / UTMALDG back to back instead of wrapping each in a divergence loop
    if (elect_one()) {
      mbar_arrive_expect_tx(w_full, Cfg::kWBytes);
      bulk_load_1d(sW, reinterpret_cast<const uint8_t*>(p.w_packed) + (size_t)n_tile * Cfg::kWBytes, Cfg::kWBytes,
                   w_full);
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int img = (int)(t / tiles_per_img);
        const int rem = (int)(t - (int64_t)img * tiles_per_img);
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * Cfg::kOutH - Cfg::kPad, x0 = tx * (8 * MT) - Cfg::kPad;
        mbar_wait(&empty[stage], phase ^ 1);
        if (CTL_DIAGF(p, 2)) {
          mbar_arrive(&full[stage]);
        } else {
          mbar_arrive_expect_tx(&full[stage], Cfg::kStageTxBytes);
          tma_load_4d(sA + stage * kStageStride, &tmap, &full[stage], x0 * 2, y0, 0, img);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(128, Cfg::kAccCols);
      constexpr uint32_t kLboA = Cfg::kChunkStride, kSboA = Cfg::kHaloW * 16;
      constexpr uint32_t kLboB = Cfg::kAccCols * 16, kSboB = 128;
      mbar_wait(w_full, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t sW_addr = smem_u32(sW);
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&acc_empty[acc], acc_phase ^ 1);
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t a_base = smem_u32(sA + stage * kStageStride);
        if (!CTL_DIAGF(p, 1)) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t d_tmem = tmem_base + (uint32_t)((acc * MT + mt) * Cfg::kAccCols);
          // VP: one MMA per filter COLUMN s and K slice, its N dimension spans the three vertical taps; the A tile
          // starts at row 0 (the M rows are the input rows) and is only shifted horizontally
          constexpr int kGroups = VP ? 3 : TAPS;
#pragma unroll
          for (int tap = 0; tap < kGroups; ++tap) {
            const int r = (TAPS == 9 && !VP) ? tap / 3 : 0, s = TAPS == 9 ? (VP ? tap : tap % 3) : 0;
            const uint32_t a_tap = a_base + (uint32_t)((r * Cfg::kHaloW + s + 8 * mt) * 16);
#pragma unroll
            for (int kk = 0; kk < CIN / 16; ++kk) {
              const uint64_t adesc = umma_smem_desc(a_tap + (uint32_t)(2 * kk) * kLboA, kLboA, kSboA);
              const uint64_t bdesc =
                  umma_smem_desc(sW_addr + (uint32_t)((tap * (CIN / 8) + 2 * kk) * (Cfg::kAccCols * 16)), kLboB, kSboB);
              umma_bf16(d_tmem, adesc, bdesc, idesc, (tap | kk) != 0 ? 1u : 0u);
            }
          }
        }
        }
        umma_commit(&empty[stage]);       // smem stage reusable once these MMAs have read it
        umma_commit(&acc_full[acc]);      // accumulator complete
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue: 8 warps, two per TMEM sub-partition
    // separate register budgets: the residual variant keeps a tile's residual pixels in flight while it waits for the
    // accumulator, the statistics variant carries 32 running sums (the host rejects res + stats), the generic-geometry
    // variant (stride 2 / ConvTranspose scatter) pays the per-item 64-bit address arithmetic the others hoist
    float* stat_smem = reinterpret_cast<float*>(sA);
    float4* xbuf = reinterpret_cast<float4*>(smem + Cfg::kOffX);
#define CTL_EPI(RES_, STATS_, GEN_, SAL_) \
    conv_epilogue<CIN, NT, TAPS, MT, STAGES, VP, RES_, STATS_, GEN_, SAL_>(p, tmem_base, sVec, acc_full, acc_empty, n0, stat_smem, xbuf)
#define CTL_EPI_BNB() \
    conv_epilogue<CIN, NT, TAPS, MT, STAGES, VP, true, true, false, false, true>(p, tmem_base, sVec, acc_full, acc_empty, n0, stat_smem, xbuf)
    if constexpr (TAPS == 9 && !VP && CIN == 128) {
      CTL_EPI(true, false, true, false);           // unpacked 3x3 with 128 input channels: only the stride-2 form runs here
    } else {
      if (p.up2x || p.subsample != 1) {
        if constexpr (!VP) CTL_EPI(true, false, true, false);
      } else if (p.bnb_act != 0) {
        if constexpr (TAPS == 9 && !VP && NT <= 32 && CIN <= 64) CTL_EPI_BNB();
      } else if (p.res != nullptr && p.sal != nullptr) {
        if constexpr ((CIN == 64 || CIN == 128) && TAPS == 1)   // the last input-gradient convolution of a decoder (up1: 1x1)
          CTL_EPI(true, false, false, true);
      } else if (p.res != nullptr) {
        CTL_EPI(true, false, false, false);
      } else if (p.stats != nullptr) {
        CTL_EPI(false, true, false, false);
      } else {
        CTL_EPI(false, false, false, false);
      }
    }
#undef CTL_EPI
#undef CTL_EPI_BNB
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemAlloc>(tmem_base);
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Blocked bf16 activation [N][C/8][H][W][8] as a 4-D tensor map of 8-BYTE elements (W*2, H, C/8, N) with box
// (halo_w*2, halo_h, C/8, 1): the whole halo tile of all channel planes in one bulk tensor copy (bit-exact whatever
// the element type; 8-byte elements keep wide halo rows inside the 256-elements-per-box-dimension limit).
int make_act_tmap(CUtensorMap* m, const void* x, int N, int H, int W, int C, int halo_w, int halo_h) {
  EncodeTiledFn enc = encode_tiled_fn();
  CTL_REQUIRE(enc != nullptr, CTL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)(C / 8) * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)halo_w * 2, (cuuint32_t)halo_h, (cuuint32_t)(C / 8), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTL_REQUIRE(r == CUDA_SUCCESS, CTL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CTL_OK;
}

template <int CIN, int NT, int TAPS, int MT, int STAGES, bool VP = false>
int launch_conv(const void* x, const ConvParams& p0, cudaStream_t st) {
  using Cfg = ConvCfg<CIN, NT, TAPS, MT, STAGES, VP>;
  ConvParams p = p0;
  p.diag = diag_flags();
  p.tiles_x = (int)ceil_div(p.W, 8 * MT);
  p.tiles_y = (int)ceil_div(p.H, Cfg::kOutH);
  p.num_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tmap;
  if (int rc = make_act_tmap(&tmap, x, p.N, p.H, p.W, CIN, Cfg::kHaloW, Cfg::kHaloH)) return rc;
  auto kern = conv_tc_kernel<CIN, NT, TAPS, MT, STAGES, VP>;
  CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes),
              "conv smem attribute");
  const int n_tiles = p.Cout / NT;
  const int ctas = (int)std::min<int64_t>(p.num_tiles, std::max(1, sm_count() / n_tiles));
  dim3 grid((unsigned)ctas, (unsigned)n_tiles);
  launch_chained(kern, grid, kConvThreads, Cfg::kSmemBytes, st)(tmap, p);
  CTL_CUDA_OK(cudaGetLastError(), "conv_tc launch");
  return CTL_OK;
}

// 1x1 filters, and 3x3 stride 2 on unpacked taps (full-resolution compute, keep even pixels)
template <int CIN, int TAPS>
int dispatch_nt(const void* x, const ConvParams& p, int nt, cudaStream_t st) {
  // MT = 2 (16x16 pixel tiles, halo overhead 1.27x) while the tile ring fits; STAGES from the smem left
  if constexpr (CIN == 16) {
    if (nt == 16) return launch_conv<16, 16, TAPS, 4, 6>(x, p, st);     // 16 x 32 pixel tiles: fewer, longer TMA rows
    if (nt == 32) return launch_conv<16, 32, TAPS, 4, 6>(x, p, st);
    if (nt == 64) return launch_conv<16, 64, TAPS, 2, 6>(x, p, st);
  } else if constexpr (CIN == 32) {
    if (nt == 16) return launch_conv<32, 16, TAPS, 2, 6>(x, p, st);
    if (nt == 32) return launch_conv<32, 32, TAPS, 2, 6>(x, p, st);
    if (nt == 64) return launch_conv<32, 64, TAPS, 2, 4>(x, p, st);
  } else if constexpr (CIN == 64) {
    if (nt == 16) return launch_conv<64, 16, TAPS, 2, 3>(x, p, st);
    if (nt == 32) return launch_conv<64, 32, TAPS, 2, 3>(x, p, st);
    if (nt == 64) return launch_conv<64, 64, TAPS, 2, 2>(x, p, st);
  } else if constexpr (CIN == 128) {
    if (nt == 16) return launch_conv<128, 16, TAPS, 1, 3>(x, p, st);
    if (nt == 32) return launch_conv<128, 32, TAPS, 1, 2>(x, p, st);
    if (nt == 64) return launch_conv<128, 64, TAPS, 1, TAPS == 9 ? 1 : 2>(x, p, st);
  }
  set_error("ctl_conv2d_c8_bf16: no kernel for Cin=%d, n_tile=%d", CIN, nt);
  return CTL_ERR_UNSUPPORTED;
}

// 3x3 stride 1 with the vertical taps packed into N (N = 3 * nt <= 192): M tiles per tile bounded by the 512 TMEM
// columns (two accumulator stages of MT * 3 * nt columns), ring depth by the shared memory left beside the weights and
// the 32 KB row-exchange buffer
template <int CIN>
int dispatch_vp(const void* x, const ConvParams& p, int nt, cudaStream_t st) {
  if constexpr (CIN == 128) {
    if (nt == 16) return launch_conv<128, 16, 9, 1, 3, true>(x, p, st);
    if (nt == 32) return launch_conv<128, 32, 9, 1, 2, true>(x, p, st);
  }
  set_error("ctl_conv2d_c8_bf16: no packed 3x3 kernel for Cin=%d, n_tile=%d", CIN, nt);
  return CTL_ERR_UNSUPPORTED;
}

#include "stem_dgrad_small.cuh"   // the stem's input gradient on the K3s pattern (called from c8_bwd.cu through the bridge below)

}  // namespace

// CTL_STEM_DGRAD_SMALL=0 (diagnostic): keep the CUDA-core stem_dgrad_kernel
bool stem_dgrad_tensor_path_handles(int64_t H, int64_t W) {
  static const bool on = [] { const char* e = getenv("CTL_STEM_DGRAD_SMALL"); return !(e && e[0] == '0'); }();
  return on && ceil_div(H, (int64_t)kSmTile) * ceil_div(W, (int64_t)kSmTile) < 8192;
}

int stem_dgrad_tensor_path(const void* dy, const float* x, int in_mode, float inv_temp, int N, int Cin, int H, int W,
                           const float* w, float* dx, cudaStream_t st) {
  StemDgradParams p = {};
  p.x = x; p.w = w; p.dx = dx; p.N = N; p.H = H; p.W = W; p.in_mode = in_mode; p.inv_temp = inv_temp;
  return Cin == 1 ? launch_stem_dgrad_small<1>(dy, p, st) : launch_stem_dgrad_small<4>(dy, p, st);
}
}  // namespace ctl

using namespace ctl;

// 1: the kernel of this layer class expects the vertically packed 3x3 weight layout.  Measured on a B200 (batch 64):
// with 128 input channels the packed form cuts the MMA issue time enough to win (128->128 @28^2: 36.3 -> 26.0 us);
// with <= 64 input channels the layers are bound by the epilogue, which the row exchange makes heavier (16->16 @224^2:
// 58.8 -> 92.6 us), so they stay on unpacked taps.
extern "C" int ctl_conv2d_vpacked(int Cin, int Cout, int taps, int subsample) {
  return (taps == 9 && subsample == 1 && Cin == 128 && Cout % 16 == 0) ? 1 : 0;
}

extern "C" int ctl_conv2d_n_tile(int Cin, int Cout, int taps) {
  if (!(Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128) || Cout <= 0 || Cout % 16 || !(taps == 1 || taps == 9))
    return -1;
  // widest N tile whose packed weights + a useful activation ring fit in 227 KB of shared memory
  const int cap = (taps == 9 && Cin == 128) ? 32 : 64;
  for (int nt = cap; nt >= 16; nt >>= 1)
    if (Cout % nt == 0) return nt;
  return -1;
}

static int conv2d_c8_impl(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed,
                          int64_t Cout, int taps, int subsample, int up2x, const float* scale,
                          const float* shift, const void* res, const float* res_scale, const float* res_shift,
                          int act, void* out, double* stats, double* sal, int sal_mode, int store_out, void* stream,
                          int bnb_act = 0) {
  CTL_REQUIRE(x && w_packed && (out || (sal && !store_out)), CTL_ERR_INVALID, "ctl_conv2d_c8_bf16: NULL pointer");
  CTL_REQUIRE(!up2x || (taps == 1 && subsample == 1 && Cout % 32 == 0), CTL_ERR_INVALID,
              "up2x (ConvTranspose2d k2 s2) needs taps == 1, subsample == 1 and Cout = 4 * out_channels, out_channels %% 8 == 0");
  CTL_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535 && H <= 32768 && W <= 32768, CTL_ERR_INVALID,
              "ctl_conv2d_c8_bf16: bad shape N=%lld H=%lld W=%lld", (long long)N, (long long)H, (long long)W);
  CTL_REQUIRE(taps == 1 || taps == 9, CTL_ERR_INVALID, "taps must be 1 (1x1) or 9 (3x3 pad 1), got %d", taps);
  CTL_REQUIRE(subsample == 1 || (subsample == 2 && H % 2 == 0 && W % 2 == 0), CTL_ERR_INVALID,
              "subsample must be 1, or 2 with even H and W");
  CTL_REQUIRE(act >= CTL_ACT_NONE && act <= CTL_ACT_SIGMOID, CTL_ERR_INVALID, "unknown activation %d", act);
  const int nt = ctl_conv2d_n_tile((int)Cin, (int)Cout, taps);
  CTL_REQUIRE(stats == nullptr || (nt > 0 && nt <= 32 && !up2x), CTL_ERR_UNSUPPORTED,
              "fused output statistics need an N tile <= 32 (Cout %% 64 != 0 or 3x3 with Cin 128) and no up2x");
  CTL_REQUIRE(stats == nullptr || res == nullptr || bnb_act != 0, CTL_ERR_UNSUPPORTED,
              "ctl_conv2d_c8_bf16: fused output statistics and a residual input cannot be combined");
  CTL_REQUIRE(bnb_act == 0 || (taps == 9 && subsample == 1 && !up2x && Cin <= 64 && stats && res && res_scale && res_shift &&
                               (bnb_act == CTL_ACT_LRELU || bnb_act == CTL_ACT_RELU)),
              CTL_ERR_UNSUPPORTED, "fused BatchNorm-backward statistics: 3x3 stride-1 convolutions with Cin <= 64 and an N tile <= 32");
  CTL_REQUIRE(N * ((H + 13) / 14) * ((W + 7) / 8) < (int64_t)1 << 31, CTL_ERR_INVALID, "ctl_conv2d_c8_bf16: too many tiles");
  CTL_REQUIRE(nt > 0, CTL_ERR_UNSUPPORTED,
              "ctl_conv2d_c8_bf16 handles Cin in {16,32,64,128} and Cout %% 16 == 0 (got Cin=%lld Cout=%lld)",
              (long long)Cin, (long long)Cout);
  CTL_REQUIRE(aligned16(x) && aligned16(w_packed) && aligned16(out) && (!res || aligned16(res)), CTL_ERR_INVALID,
              "ctl_conv2d_c8_bf16: pointers must be 16-byte aligned");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  ConvParams p = {};
  p.N = (int)N; p.H = (int)H; p.W = (int)W; p.Cout = (int)Cout;
  p.w_packed = (const __nv_bfloat16*)w_packed;
  p.scale = scale; p.shift = shift;
  p.res = (const __nv_bfloat16*)res; p.res_scale = res_scale; p.res_shift = res_shift;
  p.out = (__nv_bfloat16*)out; p.act = act; p.subsample = subsample; p.up2x = up2x; p.stats = stats;
  p.bnb_act = bnb_act;
  if (sal != nullptr) {
    CTL_REQUIRE((Cin == 64 || Cin == 128) && taps == 1 && res != nullptr && stats == nullptr && subsample == 1 && !up2x,
                CTL_ERR_UNSUPPORTED,
                "fused latent saliency: only a 1x1 input-gradient convolution with 64 / 128 input channels and a residual carries it");
    CTL_REQUIRE(sal_mode == CTL_MODE_CHANNEL || sal_mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown saliency mode %d", sal_mode);
    p.sal = sal; p.sal_mode = sal_mode; p.no_store = store_out ? 0 : 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (conv_small_handles((int)Cin, (int)Cout, taps, p) && conv_small_enabled()) return launch_conv_small(x, p, st);
  if (ctl_conv2d_vpacked((int)Cin, (int)Cout, taps, subsample))   // w_packed in the vertically packed layout
    return dispatch_vp<128>(x, p, nt, st);
  if (taps == 9) {                             // tap-major weights
    switch ((int)Cin) {
      case 16: return dispatch_nt<16, 9>(x, p, nt, st);
      case 32: return dispatch_nt<32, 9>(x, p, nt, st);
      case 64: return dispatch_nt<64, 9>(x, p, nt, st);
      default: return dispatch_nt<128, 9>(x, p, nt, st);
    }
  }
  switch ((int)Cin) {
    case 16: return dispatch_nt<16, 1>(x, p, nt, st);
    case 32: return dispatch_nt<32, 1>(x, p, nt, st);
    case 64: return dispatch_nt<64, 1>(x, p, nt, st);
    default: return dispatch_nt<128, 1>(x, p, nt, st);
  }
}

extern "C" int ctl_conv2d_c8_bf16(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed,
                                  int64_t Cout, int taps, int subsample, int up2x, const float* scale,
                                  const float* shift, const void* res, const float* res_scale, const float* res_shift,
                                  int act, void* out, double* stats, void* stream) {
  return conv2d_c8_impl(x, N, H, W, Cin, w_packed, Cout, taps, subsample, up2x, scale, shift, res, res_scale, res_shift, act,
                        out, stats, nullptr, 0, 1, stream);
}

extern "C" int ctl_conv2d_c8_bf16_saliency(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin,
                                           const void* w_packed, int64_t Cout, const void* res, void* out,
                                           double* sal_sums, int sal_mode, int store_out, void* stream) {
  CTL_REQUIRE(sal_sums != nullptr, CTL_ERR_INVALID, "ctl_conv2d_c8_bf16_saliency: NULL saliency buffer");
  return conv2d_c8_impl(x, N, H, W, Cin, w_packed, Cout, 1, 1, 0, nullptr, nullptr, res, nullptr, nullptr, CTL_ACT_NONE, out,
                        nullptr, sal_sums, sal_mode, store_out, stream);
}

extern "C" int ctl_conv2d_c8_bf16_bnbwd(const void* x, int64_t N, int64_t H, int64_t W, int64_t Cin, const void* w_packed,
                                        int64_t Cout, const void* bn_a, const float* bn_scale, const float* bn_shift,
                                        int bn_act, void* out, double* totals, void* stream) {
  CTL_REQUIRE(bn_a && bn_scale && bn_shift && totals, CTL_ERR_INVALID, "ctl_conv2d_c8_bf16_bnbwd: NULL pointer");
  return conv2d_c8_impl(x, N, H, W, Cin, w_packed, Cout, 9, 1, 0, nullptr, nullptr, bn_a, bn_scale, bn_shift, CTL_ACT_NONE, out,
                        totals, nullptr, 0, 1, stream, bn_act);
}

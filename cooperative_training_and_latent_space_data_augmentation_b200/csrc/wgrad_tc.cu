// K3w: convolution weight gradient on tcgen05 tensor cores straight from the blocked C8 activations.
//
//   dW[tap][ci][co] += sum over pixels p of  x[p + tap - pad][ci] * dy[p][co]
//
// (the autograd of nn.Conv2d in medseg/models/ebm/encoder_decoder.py:19-68, :285-348, :351-415, :418-453, :456-503; the
// reference leaves it to torch/cuDNN).  The contraction runs over PIXELS, so both operands are read "MN-major" from
// the very same C8 tiles the forward kernel uses: a core matrix is 8 consecutive pixels (K) x 8
// channels (M or N) = 128 contiguous bytes; LBO = 128 (next 8 pixels of the row), SBO = plane stride (next 8 channels).
// The x halo tile is staged ROW-INTERLEAVED, [row][C/8][pixel][8ch] (a tensor map with the plane and row dimensions
// swapped), so that consecutive 8-row groups of the MMA's M dimension (stride SBO = one row of one plane) walk
// plane 0..C/8-1 of image row y, then of row y+1, ...: ONE tcgen05.mma (K = 16 pixels) covers the three vertical taps
// r = 0..2 of all input channels when Cin <= 32 (M = 64 / 128 = four rows x Cin, the fourth row is discarded), two
// MMAs when Cin = 64, three when Cin = 128 -- 3 / 6 / 9 MMAs per 16-pixel chunk instead of 9 small ones, and the
// horizontal taps s = 0..2 are three start addresses.  Accumulators: [s][row group][NT] TMEM column ranges, fp32,
// kept across ALL tiles a CTA owns; one vectorised red.global.add epilogue per CTA.
//
//   warp 0 : TMA producer -- two box loads per tile (x halo tile, dy tile; zero-filled out of bounds)
//   warp 1 : MMA issuer
//   warp 2 : TMEM allocator
//   warps 4-7 : epilogue (tcgen05.ld -> red.global.add.v4.f32)
//
// M rows that fall on a fourth image row (or past Cin for 1x1) read whatever follows in shared memory; they only feed
// accumulator rows that are never read back.
#include <algorithm>

#include "ctl_common.cuh"
#include "ctl_tcgen05.cuh"

namespace ctl {
namespace {

using namespace sm100;

constexpr int kWgThreads = 256;
constexpr int kWgTH = 8;     // tile rows
constexpr int kWgTW = 32;    // tile columns (two 16-pixel K chunks per row)

struct WgradParams {
  int N, H, W, Cout;
  int tiles_x, tiles_y;
  int64_t num_tiles;
  float* dW;                 // fp32, accumulated into: element (tap, ci, co) at tap*s_tap + ci*s_ci + co*s_co
  int64_t s_co, s_ci, s_tap;
  int diag;                  // profiling only (env CTL_DIAG_SKIP): 1 no MMA, 2 no TMA loads, 4 no epilogue
};

template <int CIN, int NT, int TAPS, int STAGES>
struct WgCfg {
  static constexpr int kPad = TAPS == 9 ? 1 : 0;
  static constexpr int kHaloH = kWgTH + 2 * kPad;
  static constexpr int kHaloW = kWgTW + 2 * kPad;
  static constexpr int kLine = kHaloW * 16;                    // one image row of one 8-channel plane
  static constexpr int kPlanes = CIN / 8;
  static constexpr int kXStage = kHaloH * kPlanes * kLine;     // [row][plane][pixel][8ch]
  static constexpr int kDPlane = kWgTH * kWgTW * 16;
  static constexpr int kDStage = (NT / 8) * kDPlane;           // [plane][row][pixel][8ch]
  static constexpr int kStage = (kXStage + kDStage + 127) / 128 * 128;
  static constexpr int kS = TAPS == 9 ? 3 : 1;                 // horizontal taps = start addresses
  static constexpr int kR = TAPS == 9 ? 3 : 1;                 // vertical taps = image rows folded into M
  // row groups: MMA j covers image rows [j*kRowsPer, ...) with M = kMj(j)
  static constexpr int kRowsPer = CIN <= 32 ? 3 : CIN == 64 ? 2 : 1;
  static constexpr int kG = (kR + kRowsPer - 1) / kRowsPer;
  __host__ __device__ static constexpr int m_of(int j) {
    const int rows = (kR - j * kRowsPer) < kRowsPer ? (kR - j * kRowsPer) : kRowsPer;
    return rows * CIN > 64 ? 128 : 64;
  }
  // Cin <= 32, 3x3: the MMA's M already spans FOUR image rows of x (3 used), so two dy rows ride in one instruction:
  // N = (dy row y | dy row y+1) x NT.  Rows (r', ci) x columns (d, co) hold tap r = r' - d; both useful blocks of the
  // 4 x 2 grid are kept -- twice the useful work per operand fetch of the (fetch-bound) small-N MMAs.
  static constexpr bool kPair = TAPS == 9 && CIN <= 32;
  static constexpr int kNmma = kPair ? 2 * NT : NT;
  static constexpr int kTmemCols = kS * kG * kNmma;
  static constexpr int kTmemAlloc = kTmemCols <= 32 ? 32 : kTmemCols <= 64 ? 64 : kTmemCols <= 128 ? 128
                                    : kTmemCols <= 256 ? 256 : 512;
  // the discarded M rows of the last image rows reach up to 16 lines past the tile
  static constexpr int kReach = kXStage + 16 * kLine;
  static constexpr int kRing = STAGES * kStage;
  static constexpr int kLastReach = (STAGES - 1) * kStage + kReach;
  static constexpr int kOffBar = ((kRing > kLastReach ? kRing : kLastReach) + 127) / 128 * 128;
  static constexpr int kSmemBytes = (kOffBar + 128 + 127) / 128 * 128;
  // the epilogue may stage the CTA's [NT][CIN][TAPS] fp32 result in the operand ring (idle by then)
  static constexpr bool kStageFits = NT * CIN * TAPS * 4 <= kOffBar;
  static_assert(kTmemCols <= 512, "accumulators exceed TMEM");
  static_assert(NT % 16 == 0 && kNmma <= 256 && CIN % 16 == 0, "UMMA shape");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
  static_assert((kLine >> 4) < 16384 && (kDPlane >> 4) < 16384, "descriptor range");
};

// kind::f16 instruction descriptor: D = f32, A = B = bf16, BOTH operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t idesc_bf16_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int CIN, int NT, int TAPS, int STAGES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                const WgradParams p) {
  pdl_entry();
  using Cfg = WgCfg<CIN, NT, TAPS, STAGES>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBar);
  uint64_t* full = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty = bars + STAGES;        // [STAGES] MMA -> TMA
  uint64_t* acc_full = bars + 2 * STAGES; // [1]      MMA -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tile = blockIdx.y;
  const int n0 = n_tile * NT;
  const bool has_work = (int64_t)blockIdx.x < p.num_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(acc_full, 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemAlloc>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // ================================================================= TMA producer
    // elect.sync (not `lane == 0`): one provably-uniform thread, no per-instruction divergence loops around UTCHMMA
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        const int img = (int)(t / tiles_per_img);
        const int rem = (int)(t - (int64_t)img * tiles_per_img);
        const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
        const int y0 = ty * kWgTH, x0 = tx * kWgTW;
        uint8_t* sx = smem + stage * Cfg::kStage;
        mbar_wait(&empty[stage], phase ^ 1);
        if (CTL_DIAGF(p, 2)) {
          mbar_arrive(&full[stage]);
        } else {
          mbar_arrive_expect_tx(&full[stage], Cfg::kXStage + Cfg::kDStage);
          tma_load_4d(sx, &tmap_x, &full[stage], (x0 - Cfg::kPad) * 2, 0, y0 - Cfg::kPad, img);
          if (Cfg::kPair)   // row-interleaved dy tile [row][plane][pixel][8ch]
            tma_load_4d(sx + Cfg::kXStage, &tmap_dy, &full[stage], x0 * 2, n_tile * (NT / 8), y0, img);
          else
            tma_load_4d(sx + Cfg::kXStage, &tmap_dy, &full[stage], x0 * 2, y0, n_tile * (NT / 8), img);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================================================= MMA issuer
    if (has_work && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      uint32_t accumulate = 0;
      for (int64_t t = blockIdx.x; t < p.num_tiles; t += gridDim.x) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t xs = smem_u32(smem + stage * Cfg::kStage);
        const uint32_t ds = xs + Cfg::kXStage;
#pragma unroll 1
        for (int y = CTL_DIAGF(p, 1) ? kWgTH : 0; y < kWgTH; y += (Cfg::kPair ? 2 : 1)) {
#pragma unroll
          for (int xc = 0; xc < kWgTW; xc += 16) {
            // pair: N groups of 8 columns walk (dy row y, planes 0..P-1), (dy row y+1, planes 0..P-1): stride one line
            const uint64_t bdesc =
                Cfg::kPair ? umma_smem_desc(ds + (uint32_t)(y * (NT / 8) * (kWgTW * 16) + xc * 16), 128u, kWgTW * 16)
                           : umma_smem_desc(ds + (uint32_t)((y * kWgTW + xc) * 16), 128u, Cfg::kDPlane);
#pragma unroll
            for (int sx = 0; sx < Cfg::kS; ++sx) {
#pragma unroll
              for (int j = 0; j < Cfg::kG; ++j) {
                const uint32_t a_addr = xs + (uint32_t)((y + j * Cfg::kRowsPer) * Cfg::kPlanes * Cfg::kLine + (xc + sx) * 16);
                const uint64_t adesc = umma_smem_desc(a_addr, 128u, Cfg::kLine);
                umma_bf16(tmem_base + (uint32_t)((sx * Cfg::kG + j) * Cfg::kNmma), adesc, bdesc,
                          idesc_bf16_mn(Cfg::m_of(j), Cfg::kNmma), accumulate);
              }
            }
            accumulate = 1;
          }
        }
        umma_commit(&empty[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ================================================================= epilogue
    if (has_work) {
      const int q = warp - 4;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // nn.Conv2d layout [co][ci][tap]: this CTA's N tile is ONE contiguous block of NT*CIN*TAPS floats.  Transpose the
      // accumulators through the (now idle) operand ring and add the block with coalesced 16-byte reductions -- the
      // direct path below scatters 4-byte atomics with a stride of TAPS (lanes) x CIN*TAPS (registers).
      float* block = p.dW + (int64_t)n0 * CIN * TAPS;
      const bool staged = Cfg::kStageFits && p.s_tap == 1 && p.s_ci == TAPS && p.s_co == (int64_t)CIN * TAPS &&
                          (reinterpret_cast<uintptr_t>(block) & 15u) == 0;
      float* stage = reinterpret_cast<float*>(smem);
      // pair mode: pass d = 0 reads columns [0, NT) (tap r = r'), pass d = 1 columns [NT, 2NT) (tap r = r' - 1); every
      // (tap, ci, co) has exactly one writer per pass, the staged block is stored in pass 0 and added to in pass 1
#pragma unroll 1
      for (int d = 0; d < (Cfg::kPair ? 2 : 1); ++d) {
        if (d == 1 && staged && !CTL_DIAGF(p, 4)) asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
        for (int sx = CTL_DIAGF(p, 4) ? Cfg::kS : 0; sx < Cfg::kS; ++sx) {
#pragma unroll
          for (int j = 0; j < Cfg::kG; ++j) {
            // accumulator row m of this thread: M = 128 -> TMEM lane m; M = 64 -> lane (m % 16) + 32 * (m / 16)
            const bool m128 = Cfg::m_of(j) == 128;
            const int m = m128 ? q * 32 + lane : q * 16 + lane;
            const int line = m >> 3;                                       // 8-row group = (image row, plane)
            const int rp = line / Cfg::kPlanes;                            // image row of x within the MMA's M rows
            const int r = j * Cfg::kRowsPer + rp - d;                      // vertical tap this block contributes to
            const int ci = (line % Cfg::kPlanes) * 8 + (m & 7);
            const bool row_ok = (m128 || lane < 16) && r >= 0 && r < Cfg::kR &&
                                rp < (Cfg::kPair ? Cfg::kRowsPer + 1 : Cfg::kRowsPer);
            const int tap = TAPS == 9 ? r * 3 + sx : 0;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 16) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(tmem_base + ((uint32_t)(q * 32) << 16) +
                                     (uint32_t)((sx * Cfg::kG + j) * Cfg::kNmma + d * NT + c0), v);
              tmem_ld_wait();
              if (row_ok) {
                if (staged) {
                  if (d == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) stage[((c0 + i) * CIN + ci) * TAPS + tap] = __uint_as_float(v[i]);
                  } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) stage[((c0 + i) * CIN + ci) * TAPS + tap] += __uint_as_float(v[i]);
                  }
                  continue;
                }
                float* dst = p.dW + tap * p.s_tap + ci * p.s_ci + (int64_t)(n0 + c0) * p.s_co;
                if (p.s_co == 1 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                  for (int i = 0; i < 16; i += 4)
                    red_add_v4(dst + i, __uint_as_float(v[i]), __uint_as_float(v[i + 1]), __uint_as_float(v[i + 2]),
                               __uint_as_float(v[i + 3]));
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) atomicAdd(dst + i * p.s_co, __uint_as_float(v[i]));
                }
              }
            }
          }
        }
      }
      if (staged && !CTL_DIAGF(p, 4)) {
        asm volatile("bar.sync 1, 128;" ::: "memory");                 // the four epilogue warps only
        const float4* s4 = reinterpret_cast<const float4*>(stage);
        const int tid = threadIdx.x - 128;
        for (int i = tid; i < NT * CIN * TAPS / 4; i += 128) {
          const float4 f = s4[i];
          red_add_v4(block + 4 * i, f.x, f.y, f.z, f.w);
        }
      }
    }
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

// C8 tensor [N][C/8][H][W][8] as a 4-D map of 8-BYTE elements (W*2, H, C/8, N) with box (box_w*2, box_h, box_c8, 1):
// the copy is bit-exact whatever the element type, and 8-byte elements keep a 34-pixel halo row (544 B) inside the
// 256-elements-per-box-dimension limit of the tensor map.
int make_c8_tmap(CUtensorMap* m, const void* x, int N, int H, int W, int C, int box_w, int box_h, int box_c8) {
  EncodeTiledFn enc = encode_tiled_fn();
  CTL_REQUIRE(enc != nullptr, CTL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)(C / 8), (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)(C / 8) * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)box_h, (cuuint32_t)box_c8, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTL_REQUIRE(r == CUDA_SUCCESS, CTL_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return CTL_OK;
}

// x as (W*2, C/8, H, N) -- planes before rows -- so that the box lands row-interleaved: [row][plane][pixel][8ch]
int make_c8_tmap_rows(CUtensorMap* m, const void* x, int N, int H, int W, int C, int box_w, int box_h, int box_c8 = 0) {
  EncodeTiledFn enc = encode_tiled_fn();
  CTL_REQUIRE(enc != nullptr, CTL_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)(C / 8), (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)H * W * 16, (cuuint64_t)W * 16, (cuuint64_t)(C / 8) * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)box_w * 2, (cuuint32_t)(box_c8 > 0 ? box_c8 : C / 8), (cuuint32_t)box_h, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CTL_REQUIRE(r == CUDA_SUCCESS, CTL_ERR_CUDA, "cuTensorMapEncodeTiled (row-interleaved) failed with CUresult %d", (int)r);
  return CTL_OK;
}

template <int CIN, int NT, int TAPS, int STAGES>
int launch_wgrad(const void* x, const void* dy, const WgradParams& p0, cudaStream_t st) {
  using Cfg = WgCfg<CIN, NT, TAPS, STAGES>;
  WgradParams p = p0;
  p.diag = diag_flags();
  p.tiles_x = (int)ceil_div(p.W, kWgTW);
  p.tiles_y = (int)ceil_div(p.H, kWgTH);
  p.num_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  CUtensorMap tx, td;
  if (int rc = make_c8_tmap_rows(&tx, x, p.N, p.H, p.W, CIN, Cfg::kHaloW, Cfg::kHaloH)) return rc;
  if (Cfg::kPair) {
    if (int rc = make_c8_tmap_rows(&td, dy, p.N, p.H, p.W, p.Cout, kWgTW, kWgTH, NT / 8)) return rc;
  } else {
    if (int rc = make_c8_tmap(&td, dy, p.N, p.H, p.W, p.Cout, kWgTW, kWgTH, NT / 8)) return rc;
  }
  auto kern = wgrad_tc_kernel<CIN, NT, TAPS, STAGES>;
  CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes),
              "wgrad smem attribute");
  const int n_tiles = p.Cout / NT;
  // every CTA ends with TAPS*CIN*NT atomics: give it enough tiles to amortise them
  const int64_t min_tiles = std::max<int64_t>(1, (int64_t)TAPS * CIN * NT / 8192);
  int64_t ctas = std::min<int64_t>(ceil_div(p.num_tiles, min_tiles), std::max(1, sm_count() / n_tiles));
  ctas = std::max<int64_t>(1, std::min<int64_t>(ctas, p.num_tiles));
  dim3 grid((unsigned)ctas, (unsigned)n_tiles);
  launch_chained(kern, grid, kWgThreads, Cfg::kSmemBytes, st)(tx, td, p);
  CTL_CUDA_OK(cudaGetLastError(), "wgrad_tc launch");
  return CTL_OK;
}

int wgrad_n_tile(int Cin, int Cout, int taps) {
  // TMEM: kS*kG*NT <= 512 columns -> 3x3: NT <= 32 (Cin 128), 64 (Cin 64); 1x1: 128
  // (pair mode, Cin <= 32 3x3: 3 * 2*NT <= 512 -> NT <= 64)
  const int cap = taps == 9 ? (Cin == 128 ? 32 : 64) : 128;
  for (int nt = cap; nt >= 16; nt >>= 1)
    if (Cout % nt == 0) return nt;
  return -1;
}

template <int CIN, int TAPS>
int dispatch_wgrad(const void* x, const void* dy, const WgradParams& p, int nt, cudaStream_t st) {
  constexpr int S = CIN == 128 ? 2 : CIN == 64 ? 3 : 4;
  if (nt == 16) return launch_wgrad<CIN, 16, TAPS, S>(x, dy, p, st);
  if (nt == 32) return launch_wgrad<CIN, 32, TAPS, S>(x, dy, p, st);
  if constexpr (TAPS == 1 || CIN <= 64) {
    if (nt == 64) return launch_wgrad<CIN, 64, TAPS, (CIN >= 64 ? 2 : S)>(x, dy, p, st);
  }
  if constexpr (TAPS == 1) {
    if (nt == 128) return launch_wgrad<CIN, 128, TAPS, (CIN == 128 ? 1 : 2)>(x, dy, p, st);
  }
  set_error("ctl_conv_wgrad_c8_bf16: no kernel for Cin=%d, n_tile=%d, taps=%d", CIN, nt, TAPS);
  return CTL_ERR_UNSUPPORTED;
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" int ctl_conv_wgrad_c8_bf16(const void* x, const void* dy, int64_t N, int64_t H, int64_t W, int64_t Cin,
                                      int64_t Cout, int taps, float* dW, int64_t stride_co, int64_t stride_ci,
                                      int64_t stride_tap, void* stream) {
  CTL_REQUIRE(x && dy && dW, CTL_ERR_INVALID, "ctl_conv_wgrad_c8_bf16: NULL pointer");
  CTL_REQUIRE(N > 0 && H > 0 && W > 0 && N <= 65535 && H <= 32768 && W <= 32768, CTL_ERR_INVALID,
              "ctl_conv_wgrad_c8_bf16: bad shape N=%lld H=%lld W=%lld", (long long)N, (long long)H, (long long)W);
  CTL_REQUIRE(taps == 1 || taps == 9, CTL_ERR_INVALID, "taps must be 1 (1x1) or 9 (3x3 pad 1), got %d", taps);
  CTL_REQUIRE((Cin == 16 || Cin == 32 || Cin == 64 || Cin == 128) && Cout > 0 && Cout % 16 == 0, CTL_ERR_UNSUPPORTED,
              "ctl_conv_wgrad_c8_bf16 handles Cin in {16,32,64,128} and Cout %% 16 == 0 (got Cin=%lld Cout=%lld)",
              (long long)Cin, (long long)Cout);
  CTL_REQUIRE(aligned16(x) && aligned16(dy) && (reinterpret_cast<uintptr_t>(dW) & 3u) == 0, CTL_ERR_INVALID,
              "ctl_conv_wgrad_c8_bf16: x / dy must be 16-byte aligned");
  CTL_REQUIRE(stride_co > 0 && stride_ci > 0 && stride_tap > 0, CTL_ERR_INVALID, "ctl_conv_wgrad_c8_bf16: strides must be positive");
  const int nt = wgrad_n_tile((int)Cin, (int)Cout, taps);
  CTL_REQUIRE(nt > 0, CTL_ERR_UNSUPPORTED, "ctl_conv_wgrad_c8_bf16: no N tile for Cout=%lld", (long long)Cout);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  WgradParams p = {};
  p.N = (int)N; p.H = (int)H; p.W = (int)W; p.Cout = (int)Cout; p.dW = dW;
  p.s_co = stride_co; p.s_ci = stride_ci; p.s_tap = stride_tap;
  cudaStream_t st = (cudaStream_t)stream;
  if (taps == 9) {
    switch ((int)Cin) {
      case 16: return dispatch_wgrad<16, 9>(x, dy, p, nt, st);
      case 32: return dispatch_wgrad<32, 9>(x, dy, p, nt, st);
      case 64: return dispatch_wgrad<64, 9>(x, dy, p, nt, st);
      default: return dispatch_wgrad<128, 9>(x, dy, p, nt, st);
    }
  }
  switch ((int)Cin) {
    case 16: return dispatch_wgrad<16, 1>(x, dy, p, nt, st);
    case 32: return dispatch_wgrad<32, 1>(x, dy, p, nt, st);
    case 64: return dispatch_wgrad<64, 1>(x, dy, p, nt, st);
    default: return dispatch_wgrad<128, 1>(x, dy, p, nt, st);
  }
}

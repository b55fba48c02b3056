// K1 / K2 / dropout: the latent-masking kernels (HBM-bound streaming work, sm_100a).
//
// Replaces the tail of mask_latent_code_channel_wise / _spatial_wise
// (medseg/models/model_util.py:224-249, :285-312) and the dropout branch of perturb_latent_code
// (medseg/models/advanced_triplet_recon_segmentation_model.py:332-336).
//
// Data layout: dense NCHW exactly as the reference holds the latent codes, viewed as
// [rows = N*C][HW].  A sample's C*HW block is contiguous, a row is HW contiguous elements.
// All global traffic is 128-bit (4 x fp32 / 8 x bf16) when HW is a multiple of the vector
// width and the base pointers are 16-byte aligned; a scalar instantiation covers the rest.
//
// Algorithmic bytes (what roofline.achieved is computed from):
//   saliency reduce : sizeof(g) * N*C*HW                       (+ 4*N*n written)
//   top-p apply     : (sizeof(z) + sizeof(z_out)) * N*C*HW     (+ 4*N*n read/written)
//   dropout         : (sizeof(z) + sizeof(z_out)) * N*C*HW     (+ 4*N*C*HW with the quirk mask)
#include <algorithm>

#include "ctl_common.cuh"
#include "ctl_philox.cuh"

namespace ctl {
namespace {

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------------
// Streaming building block: a group of L lanes owns one row; each lane issues up to U independent
// 128-bit loads before it touches any of them (U*16 bytes in flight per lane).
// ------------------------------------------------------------------------------------------------
constexpr int kU = 8;

template <typename T, int VEC, int L>
__device__ __forceinline__ void load_batch(const T* __restrict__ row, int base, int lane, int nv,
                                           float (&a)[kU][VEC]) {
#pragma unroll
  for (int j = 0; j < kU; ++j) {
    const int v = base + j * L + lane;
    if (v < nv) {
      load_as_float<T, VEC>(row + (int64_t)v * VEC, a[j]);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) a[j][i] = 0.0f;
    }
  }
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// K1, channel mode: one L-lane group per (n,c) row, fp64 accumulation, shuffle reduction.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int L>
__global__ void __launch_bounds__(kThreads)
saliency_channel_kernel(const T* __restrict__ g, float* __restrict__ s, int64_t rows, int HW, int nv) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & (L - 1);
  const unsigned gmask = group_mask<L>();
  const int64_t row = (int64_t)blockIdx.x * (kThreads / L) + threadIdx.x / L;
  if (row >= rows) return;                     // whole group leaves together
  const T* __restrict__ p = g + row * HW;
  double acc0 = 0.0, acc1 = 0.0;
  for (int base = 0; base < nv; base += kU * L) {
    float a[kU][VEC];
    load_batch<T, VEC, L>(p, base, lane, nv, a);
#pragma unroll
    for (int j = 0; j < kU; j += 2) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) { acc0 += (double)a[j][i]; acc1 += (double)a[j + 1][i]; }
    }
  }
  double acc = acc0 + acc1;
#pragma unroll
  for (int o = L / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
  if (lane == 0) s[row] = (float)(acc / (double)HW);
}

// ------------------------------------------------------------------------------------------------
// K1, spatial mode: CTA = X vector columns x Y channel slices of one sample.  A warp reads X*16
// contiguous bytes of one channel row; each thread keeps kU channel rows in flight; the Y partial
// sums meet in shared memory (fp64).
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int X, int Y>
__global__ void __launch_bounds__(X * Y)
saliency_spatial_kernel(const T* __restrict__ g, float* __restrict__ s, int C, int HW, int nv) {
  pdl_launch_dependents();
  __shared__ double red[Y][X * VEC + 1];
  const int x = threadIdx.x % X, y = threadIdx.x / X;
  const int64_t n = blockIdx.y;
  const int v = blockIdx.x * X + x;
  double acc[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = 0.0;
  if (v < nv) {
    const T* __restrict__ p = g + n * (int64_t)C * HW + (int64_t)v * VEC;
    for (int c0 = y; c0 < C; c0 += kU * Y) {
      float a[kU][VEC];
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        const int c = c0 + j * Y;
        if (c < C) {
          load_as_float<T, VEC>(p + (int64_t)c * HW, a[j]);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) a[j][i] = 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < kU; ++j) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[i] += (double)a[j][i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) red[y][x * VEC + i] = acc[i];
  __syncthreads();
  for (int e = threadIdx.x; e < X * VEC; e += X * Y) {
    const int64_t hw = (int64_t)blockIdx.x * X * VEC + e;
    if (hw < HW) {
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < Y; ++j) t += red[j][e];
      s[n * HW + hw] = (float)(t / (double)C);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K2: per-sample k-th largest (radix select on order-preserving keys), mask build, apply.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float f) {
  // larger float -> larger key; NaN sorts first in torch.sort(descending=True) -> largest key
  if (f != f) return 0xffffffffu;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// All threads of the CTA call this; keys[0..n) live in shared memory.  Returns the key of the
// k-th (0-based) largest element.
__device__ uint32_t block_kth_largest_key(const uint32_t* __restrict__ keys, int n, int k,
                                          uint32_t* hist /*[256]*/, uint32_t* sel /*[2]*/) {
  uint32_t prefix = 0, known = 0, krem = (uint32_t)k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const uint32_t key = keys[j];
      if ((key & known) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns bins [255-8l-7, 255-8l], walked from the top
      const int lane = threadIdx.x;
      uint32_t local = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) local += hist[255 - 8 * lane - b];
      uint32_t incl = local;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const uint32_t before = incl - local;          // elements in strictly higher bins (other lanes)
      if (krem >= before && krem < incl) {           // exactly one lane
        uint32_t cum = before;
        for (int b = 0; b < 8; ++b) {
          const uint32_t h = hist[255 - 8 * lane - b];
          if (krem < cum + h) {
            sel[0] = (uint32_t)(255 - 8 * lane - b);
            sel[1] = krem - cum;
            break;
          }
          cum += h;
        }
      }
    }
    __syncthreads();
    prefix |= sel[0] << shift;
    known |= 255u << shift;
    krem = sel[1];
  }
  return prefix;
}

__device__ __forceinline__ float mask_value(float sv, float thr, int soft, const float* __restrict__ rand,
                                            PhiloxKey key, uint64_t gidx, int64_t lidx) {
  if (!(sv > thr)) return 1.0f;
  if (!soft) return 0.0f;
  const float u = rand ? rand[lidx] : philox_uniform(key, gidx);
  return 0.5f * u;
}

template <int VEC, int MODE>
__device__ __forceinline__ void scale_vec(float (&a)[VEC], float mrow, const float* __restrict__ mask_sm, int v) {
#pragma unroll
  for (int i = 0; i < VEC; ++i) a[i] *= (MODE == CTL_MODE_CHANNEL) ? mrow : mask_sm[v * VEC + i];
}

// CTA = rows [c0, c0+rows_per_cta) of one sample; group gi (L lanes) owns rows c0+gi, c0+gi+G, ...
// The first row's data is requested BEFORE the dependency wait / select: z does not depend on K1.
template <typename ZT, typename OT, int VEC, int L, int MODE, bool PDL>
__global__ void __launch_bounds__(kThreads)
topp_mask_apply_kernel(const float* __restrict__ s, const ZT* __restrict__ z, OT* __restrict__ z_out,
                       float* __restrict__ mask_out, float* __restrict__ thr_out,
                       const float* __restrict__ rand, PhiloxKey key, int C, int HW, int nv, int k, int soft,
                       int rows_per_cta, int64_t first_sample) {
  extern __shared__ float mask_sm[];          // n floats: first the keys of s, then (in place) the mask
  __shared__ uint32_t hist[256];
  __shared__ uint32_t sel[2];
  constexpr int G = kThreads / L;
  const int64_t sample = blockIdx.y;
  const int n = (MODE == CTL_MODE_CHANNEL) ? C : HW;
  const int c0 = blockIdx.x * rows_per_cta;
  const int crows = min(rows_per_cta, C - c0);
  const int lane = threadIdx.x & (L - 1);
  const int gi = threadIdx.x / L;
  const int64_t base = (sample * C + c0) * (int64_t)HW;

  float a[kU][VEC];
  if (gi < crows) load_batch<ZT, VEC, L>(z + base + (int64_t)gi * HW, 0, lane, nv, a);

  if (PDL) pdl_wait();                        // s is produced by the preceding K1 launch

  const float* __restrict__ srow = s + sample * n;
  uint32_t* keys = reinterpret_cast<uint32_t*>(mask_sm);
  for (int j = threadIdx.x; j < n; j += kThreads) keys[j] = order_key(srow[j]);
  const float thr = key_to_float(block_kth_largest_key(keys, n, k, hist, sel));   // syncs inside
  if (thr_out && blockIdx.x == 0 && threadIdx.x == 0) thr_out[sample] = thr;

  const uint64_t gbase = (uint64_t)(first_sample + sample) * (uint64_t)n;
  for (int j = threadIdx.x; j < n; j += kThreads) {
    // same thread reads keys[j] and overwrites it with the mask value: no hazard
    const float m = mask_value(key_to_float(keys[j]), thr, soft, rand, key, gbase + j, sample * n + j);
    mask_sm[j] = m;
    const bool mine = (MODE == CTL_MODE_CHANNEL) ? (j >= c0 && j < c0 + crows) : (blockIdx.x == 0);
    if (mine) mask_out[sample * n + j] = m;
  }
  __syncthreads();

  for (int r = gi; r < crows; r += G) {
    const ZT* __restrict__ zi = z + base + (int64_t)r * HW;
    OT* __restrict__ zo = z_out + base + (int64_t)r * HW;
    const float mrow = (MODE == CTL_MODE_CHANNEL) ? mask_sm[c0 + r] : 1.0f;
    for (int vb = 0; vb < nv; vb += kU * L) {
      if (r != gi || vb != 0) load_batch<ZT, VEC, L>(zi, vb, lane, nv, a);
#pragma unroll
      for (int j = 0; j < kU; ++j) {
        const int v = vb + j * L + lane;
        if (v < nv) {
          scale_vec<VEC, MODE>(a[j], mrow, mask_sm, v);
          store_from_float<OT, VEC>(zo + (int64_t)v * VEC, a[j]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Channel dropout: one L-lane group per (n,c) row.
// ------------------------------------------------------------------------------------------------
template <typename ZT, typename OT, int VEC, int L>
__global__ void __launch_bounds__(kThreads)
channel_dropout_kernel(const ZT* __restrict__ z, OT* __restrict__ z_out, float* __restrict__ mask_out,
                       const float* __restrict__ keep, float* __restrict__ keep_out, PhiloxKey key,
                       int64_t rows, int HW, int nv, float p, float scale, uint64_t first_row) {
  const int lane = threadIdx.x & (L - 1);
  const int64_t row = (int64_t)blockIdx.x * (kThreads / L) + threadIdx.x / L;
  if (row >= rows) return;
  const ZT* __restrict__ zi = z + row * HW;
  OT* __restrict__ zo = z_out + row * HW;
  float* __restrict__ mo = mask_out ? mask_out + row * HW : nullptr;
  float a[kU][VEC];
  load_batch<ZT, VEC, L>(zi, 0, lane, nv, a);            // in flight while the Philox draw is computed
  const float kf = keep ? keep[row] : (philox_uniform(key, first_row + (uint64_t)row) >= p ? 1.0f : 0.0f);
  if (keep_out && lane == 0) keep_out[row] = kf;
  const float noise = round_through<ZT>(kf * scale);     // the reference's noise tensor has z's dtype
  for (int vb = 0; vb < nv; vb += kU * L) {
    if (vb != 0) load_batch<ZT, VEC, L>(zi, vb, lane, nv, a);
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int v = vb + j * L + lane;
      if (v < nv) {
        float o[VEC], m[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          o[i] = a[j][i] * noise;
          m[i] = (round_through<OT>(o[i]) == a[j][i]) ? 1.0f : 0.0f;
        }
        store_from_float<OT, VEC>(zo + (int64_t)v * VEC, o);
        if (mo) store_from_float<float, VEC>(mo + (int64_t)v * VEC, m);
      }
    }
  }
}

__global__ void philox_uniform_kernel(PhiloxKey key, uint64_t first, int64_t count, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = philox_uniform(key, first + (uint64_t)i);
}

// ---- launch helpers ------------------------------------------------------------------------------
// lanes per row: enough rows in flight per CTA, yet most of a row covered by one kU-deep batch
inline int pick_lanes(int nv) { return nv >= 128 ? 32 : nv >= 64 ? 16 : nv >= 16 ? 8 : 4; }

template <typename T, int VEC>
int launch_saliency_channel(const T* g, float* s, int64_t rows, int HW, cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int64_t grid64 = ceil_div(rows, kThreads / L);
  CTL_REQUIRE(grid64 <= 0x7fffffff, CTL_ERR_UNSUPPORTED, "too many rows for one launch");
  const unsigned grid = (unsigned)grid64;
  switch (L) {
    case 32: saliency_channel_kernel<T, VEC, 32><<<grid, kThreads, 0, st>>>(g, s, rows, HW, nv); break;
    case 16: saliency_channel_kernel<T, VEC, 16><<<grid, kThreads, 0, st>>>(g, s, rows, HW, nv); break;
    case 8: saliency_channel_kernel<T, VEC, 8><<<grid, kThreads, 0, st>>>(g, s, rows, HW, nv); break;
    default: saliency_channel_kernel<T, VEC, 4><<<grid, kThreads, 0, st>>>(g, s, rows, HW, nv); break;
  }
  CTL_CUDA_OK(cudaGetLastError(), "saliency_channel launch");
  return CTL_OK;
}

template <typename T, int VEC>
int launch_saliency_spatial(const T* g, float* s, int64_t N, int C, int HW, cudaStream_t st) {
  const int nv = HW / VEC;
  constexpr int X = 32, Y = 8;
  dim3 grid((unsigned)ceil_div(nv, X), (unsigned)N);
  saliency_spatial_kernel<T, VEC, X, Y><<<grid, X * Y, 0, st>>>(g, s, C, HW, nv);
  CTL_CUDA_OK(cudaGetLastError(), "saliency_spatial launch");
  return CTL_OK;
}

template <typename ZT, typename OT, int VEC, int L, int MODE, bool PDL>
int launch_topp_kernel(dim3 grid, size_t smem, cudaStream_t st, const float* s, const ZT* z, OT* z_out,
                       float* mask_out, float* thr_out, const float* rand, PhiloxKey key, int C, int HW, int nv, int k,
                       int soft, int rows_per_cta, int64_t first_sample) {
  auto kern = topp_mask_apply_kernel<ZT, OT, VEC, L, MODE, PDL>;
  if (smem > 48 * 1024)
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                "topp smem attribute");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = PDL ? 1 : 0;
  CTL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, s, z, z_out, mask_out, thr_out, rand, key, C, HW, nv, k, soft,
                                 rows_per_cta, first_sample),
              "topp_mask_apply launch");
  return CTL_OK;
}

template <typename ZT, typename OT, int VEC, int MODE, bool PDL>
int launch_topp(const float* s, const ZT* z, OT* z_out, float* mask_out, float* thr_out, const float* rand,
                PhiloxKey key, int64_t N, int C, int HW, int k, int soft, int64_t first_sample, cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int G = kThreads / L;
  // one row per group per CTA keeps every lane busy and the per-CTA select cheap relative to its stream;
  // take more rows per group only when the rows are short (< 4 KB) and the grid stays >= 4 waves
  int rows_per_cta = G;
  while ((int64_t)rows_per_cta * HW * (int64_t)sizeof(ZT) < 16 * 1024 && rows_per_cta < C &&
         N * ceil_div(C, 2 * rows_per_cta) >= 8 * (int64_t)sm_count())
    rows_per_cta *= 2;
  rows_per_cta = std::min(rows_per_cta, C);
  const int chunks = (int)ceil_div(C, rows_per_cta);
  const int n = MODE == CTL_MODE_CHANNEL ? C : HW;
  const size_t smem = sizeof(float) * (size_t)n;
  dim3 grid((unsigned)chunks, (unsigned)N);
#define CTL_LAUNCH_TOPP(LL)                                                                                      \
  return launch_topp_kernel<ZT, OT, VEC, LL, MODE, PDL>(grid, smem, st, s, z, z_out, mask_out, thr_out, rand, key, C, \
                                                        HW, nv, k, soft, rows_per_cta, first_sample)
  switch (L) {
    case 32: CTL_LAUNCH_TOPP(32);
    case 16: CTL_LAUNCH_TOPP(16);
    case 8: CTL_LAUNCH_TOPP(8);
    default: CTL_LAUNCH_TOPP(4);
  }
#undef CTL_LAUNCH_TOPP
}

template <typename ZT, typename OT, int VEC>
int launch_topp_mode(int mode, bool pdl, const float* s, const ZT* z, OT* z_out, float* mask_out, float* thr_out,
                     const float* rand, PhiloxKey key, int64_t N, int C, int HW, int k, int soft,
                     int64_t first_sample, cudaStream_t st) {
#define CTL_ARGS s, z, z_out, mask_out, thr_out, rand, key, N, C, HW, k, soft, first_sample, st
  if (mode == CTL_MODE_CHANNEL)
    return pdl ? launch_topp<ZT, OT, VEC, CTL_MODE_CHANNEL, true>(CTL_ARGS)
               : launch_topp<ZT, OT, VEC, CTL_MODE_CHANNEL, false>(CTL_ARGS);
  return pdl ? launch_topp<ZT, OT, VEC, CTL_MODE_SPATIAL, true>(CTL_ARGS)
             : launch_topp<ZT, OT, VEC, CTL_MODE_SPATIAL, false>(CTL_ARGS);
#undef CTL_ARGS
}

template <typename ZT, typename OT, int VEC>
int launch_dropout(const ZT* z, OT* z_out, float* mask_out, const float* keep, float* keep_out, PhiloxKey key,
                   int64_t rows, int HW, float p, float scale, uint64_t first_row, cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int64_t grid64 = ceil_div(rows, kThreads / L);
  CTL_REQUIRE(grid64 <= 0x7fffffff, CTL_ERR_UNSUPPORTED, "too many rows for one launch");
  const unsigned grid = (unsigned)grid64;
  switch (L) {
    case 32: channel_dropout_kernel<ZT, OT, VEC, 32><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row); break;
    case 16: channel_dropout_kernel<ZT, OT, VEC, 16><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row); break;
    case 8: channel_dropout_kernel<ZT, OT, VEC, 8><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row); break;
    default: channel_dropout_kernel<ZT, OT, VEC, 4><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row); break;
  }
  CTL_CUDA_OK(cudaGetLastError(), "channel_dropout launch");
  return CTL_OK;
}

bool valid_dtype(int d) { return d == CTL_F32 || d == CTL_BF16; }

int check_shape(int64_t N, int64_t C, int64_t HW) {
  CTL_REQUIRE(N > 0 && C > 0 && HW > 0, CTL_ERR_INVALID, "N, C, HW must be positive (got %lld, %lld, %lld)",
              (long long)N, (long long)C, (long long)HW);
  CTL_REQUIRE(N <= 65535 && C <= (1 << 24) && HW <= (1 << 24) && N * C * HW < (1ll << 40), CTL_ERR_UNSUPPORTED,
              "shape [%lld,%lld,%lld] exceeds kernel limits (N<=65535, C,HW<=2^24)", (long long)N, (long long)C,
              (long long)HW);
  return CTL_OK;
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" int ctl_saliency_reduce(const void* g, int g_dtype, int64_t N, int64_t C, int64_t HW, int mode,
                                   float* s_out, void* stream) {
  CTL_REQUIRE(g && s_out, CTL_ERR_INVALID, "ctl_saliency_reduce: NULL pointer");
  CTL_REQUIRE(valid_dtype(g_dtype), CTL_ERR_INVALID, "ctl_saliency_reduce: unknown dtype %d", g_dtype);
  CTL_REQUIRE(mode == CTL_MODE_CHANNEL || mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown mode %d", mode);
  if (int rc = check_shape(N, C, HW)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const bool f32 = g_dtype == CTL_F32;
  const int vw = f32 ? 4 : 8;
  const bool vec = (HW % vw == 0) && aligned16(g);
  if (mode == CTL_MODE_CHANNEL) {
    if (f32) return vec ? launch_saliency_channel<float, 4>((const float*)g, s_out, N * C, (int)HW, st)
                        : launch_saliency_channel<float, 1>((const float*)g, s_out, N * C, (int)HW, st);
    return vec ? launch_saliency_channel<__nv_bfloat16, 8>((const __nv_bfloat16*)g, s_out, N * C, (int)HW, st)
               : launch_saliency_channel<__nv_bfloat16, 1>((const __nv_bfloat16*)g, s_out, N * C, (int)HW, st);
  }
  if (f32) return vec ? launch_saliency_spatial<float, 4>((const float*)g, s_out, N, (int)C, (int)HW, st)
                      : launch_saliency_spatial<float, 1>((const float*)g, s_out, N, (int)C, (int)HW, st);
  return vec ? launch_saliency_spatial<__nv_bfloat16, 8>((const __nv_bfloat16*)g, s_out, N, (int)C, (int)HW, st)
             : launch_saliency_spatial<__nv_bfloat16, 1>((const __nv_bfloat16*)g, s_out, N, (int)C, (int)HW, st);
}

static int topp_impl(bool pdl, const float* s, const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW,
                     int mode, int64_t k, int soft, const float* rand, uint64_t seed, uint64_t offset,
                     int64_t first_sample, float* mask_out, float* thr_out, void* z_out, int out_dtype,
                     void* stream) {
  CTL_REQUIRE(s && z && mask_out && z_out, CTL_ERR_INVALID, "ctl_topp_mask_apply: NULL pointer");
  CTL_REQUIRE(valid_dtype(z_dtype) && valid_dtype(out_dtype), CTL_ERR_INVALID, "ctl_topp_mask_apply: unknown dtype");
  CTL_REQUIRE(mode == CTL_MODE_CHANNEL || mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown mode %d", mode);
  if (int rc = check_shape(N, C, HW)) return rc;
  const int64_t n = mode == CTL_MODE_CHANNEL ? C : HW;
  CTL_REQUIRE(k >= 0, CTL_ERR_INVALID, "k must be >= 0 (got %lld)", (long long)k);
  CTL_REQUIRE(k < n, CTL_ERR_INDEX, "index %lld is out of bounds for dimension 1 with size %lld", (long long)k,
              (long long)n);
  CTL_REQUIRE(first_sample >= 0, CTL_ERR_INVALID, "first_sample must be >= 0");
  CTL_REQUIRE(n <= 51200, CTL_ERR_UNSUPPORTED,
              "one sample's saliency row is kept in shared memory: n=%lld > 51200", (long long)n);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const PhiloxKey key{seed, offset};
  const bool zf = z_dtype == CTL_F32, of = out_dtype == CTL_F32;
  const int vw = zf ? 4 : 8;
  const bool vec = (HW % vw == 0) && aligned16(z) && aligned16(z_out);
  const int Ci = (int)C, HWi = (int)HW, ki = (int)k;
#define CTL_TOPP(ZT, OT, V) \
  return launch_topp_mode<ZT, OT, V>(mode, pdl, s, (const ZT*)z, (OT*)z_out, mask_out, thr_out, rand, key, N, Ci, HWi, \
                                     ki, soft != 0, first_sample, st)
  if (zf && of) { if (vec) CTL_TOPP(float, float, 4); else CTL_TOPP(float, float, 1); }
  if (zf && !of) { if (vec) CTL_TOPP(float, __nv_bfloat16, 4); else CTL_TOPP(float, __nv_bfloat16, 1); }
  if (!zf && of) { if (vec) CTL_TOPP(__nv_bfloat16, float, 8); else CTL_TOPP(__nv_bfloat16, float, 1); }
  if (vec) CTL_TOPP(__nv_bfloat16, __nv_bfloat16, 8); else CTL_TOPP(__nv_bfloat16, __nv_bfloat16, 1);
#undef CTL_TOPP
}

extern "C" int ctl_topp_mask_apply(const float* s, const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW,
                                   int mode, int64_t k, int soft, const float* rand, uint64_t seed, uint64_t offset,
                                   int64_t first_sample, float* mask_out, float* thr_out, void* z_out, int out_dtype,
                                   void* stream) {
  // stand-alone: z may have been written by the kernel just before us in the stream -> ordinary launch
  return topp_impl(false, s, z, z_dtype, N, C, HW, mode, k, soft, rand, seed, offset, first_sample, mask_out, thr_out,
                   z_out, out_dtype, stream);
}

extern "C" int ctl_saliency_mask_apply(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N, int64_t C,
                                       int64_t HW, int mode, int64_t k, int soft, const float* rand, uint64_t seed,
                                       uint64_t offset, int64_t first_sample, float* s_scratch, float* mask_out,
                                       float* thr_out, void* z_out, int out_dtype, void* stream) {
  // validate k first so that an out-of-range k launches nothing (the reference raises before masking)
  const int64_t n = mode == CTL_MODE_CHANNEL ? C : HW;
  CTL_REQUIRE(k >= 0, CTL_ERR_INVALID, "k must be >= 0 (got %lld)", (long long)k);
  CTL_REQUIRE(k < n, CTL_ERR_INDEX, "index %lld is out of bounds for dimension 1 with size %lld", (long long)k,
              (long long)n);
  CTL_REQUIRE(n <= 51200, CTL_ERR_UNSUPPORTED,
              "one sample's saliency row is kept in shared memory: n=%lld > 51200", (long long)n);
  CTL_REQUIRE(s_scratch && z && mask_out && z_out, CTL_ERR_INVALID, "ctl_saliency_mask_apply: NULL pointer");
  if (int rc = ctl_saliency_reduce(g, g_dtype, N, C, HW, mode, s_scratch, stream)) return rc;
  // K2 is launched with programmatic stream serialization: its CTAs become resident while K1 drains, request
  // their first rows of z (K1 never writes z; everything older than K1 has completed because K1 itself was an
  // ordinary launch) and only then wait for K1's s.
  return topp_impl(true, s_scratch, z, z_dtype, N, C, HW, mode, k, soft, rand, seed, offset, first_sample, mask_out,
                   thr_out, z_out, out_dtype, stream);
}

extern "C" int ctl_channel_dropout(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p, float scale,
                                   const float* keep, uint64_t seed, uint64_t offset, int64_t first_sample,
                                   void* z_out, int out_dtype, float* mask_out, float* keep_out, void* stream) {
  CTL_REQUIRE(z && z_out, CTL_ERR_INVALID, "ctl_channel_dropout: NULL pointer");
  CTL_REQUIRE(valid_dtype(z_dtype) && valid_dtype(out_dtype), CTL_ERR_INVALID, "ctl_channel_dropout: unknown dtype");
  CTL_REQUIRE(p >= 0.0f && p <= 1.0f, CTL_ERR_INVALID,
              "dropout probability has to be between 0 and 1, but got %g", (double)p);
  CTL_REQUIRE(first_sample >= 0, CTL_ERR_INVALID, "first_sample must be >= 0");
  if (int rc = check_shape(N, C, HW)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const PhiloxKey key{seed, offset};
  const bool zf = z_dtype == CTL_F32, of = out_dtype == CTL_F32;
  const int vw = zf ? 4 : 8;
  const bool vec = (HW % vw == 0) && aligned16(z) && aligned16(z_out) && (!mask_out || aligned16(mask_out));
  const int64_t rows = N * C;
  const uint64_t first_row = (uint64_t)first_sample * (uint64_t)C;
#define CTL_DROP(ZT, OT, V) \
  return launch_dropout<ZT, OT, V>((const ZT*)z, (OT*)z_out, mask_out, keep, keep_out, key, rows, (int)HW, p, scale, \
                                   first_row, st)
  if (zf && of) { if (vec) CTL_DROP(float, float, 4); else CTL_DROP(float, float, 1); }
  if (zf && !of) { if (vec) CTL_DROP(float, __nv_bfloat16, 4); else CTL_DROP(float, __nv_bfloat16, 1); }
  if (!zf && of) { if (vec) CTL_DROP(__nv_bfloat16, float, 8); else CTL_DROP(__nv_bfloat16, float, 1); }
  if (vec) CTL_DROP(__nv_bfloat16, __nv_bfloat16, 8); else CTL_DROP(__nv_bfloat16, __nv_bfloat16, 1);
#undef CTL_DROP
}

extern "C" int ctl_philox_uniform(uint64_t seed, uint64_t offset, uint64_t first_index, int64_t count, float* out,
                                  void* stream) {
  CTL_REQUIRE(out && count > 0, CTL_ERR_INVALID, "ctl_philox_uniform: bad arguments");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int grid = (int)std::min<int64_t>(ceil_div(count, 256), (int64_t)sm_count() * 8);
  philox_uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(PhiloxKey{seed, offset}, first_index, count, out);
  CTL_CUDA_OK(cudaGetLastError(), "philox_uniform launch");
  return CTL_OK;
}

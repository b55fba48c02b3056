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
#include <stdlib.h>
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
// Per-sample selection: k-th largest of s[0..n) by radix select on order-preserving keys, then the
// mask row.  Called by ONE CTA per sample (the last K1 CTA to arrive, or the stand-alone select kernel).
// ------------------------------------------------------------------------------------------------
constexpr int kSelKeys = 2048;     // rows up to this long are staged in shared memory (8 KB)

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

struct SelectSmem {
  uint32_t keys[kSelKeys];
  uint32_t hist[256];
  uint32_t sel[2];
};

__device__ __forceinline__ float mask_value(float sv, float thr, int soft, const float* __restrict__ rand,
                                            PhiloxKey key, uint64_t gidx, int64_t lidx) {
  if (!(sv > thr)) return 1.0f;
  if (!soft) return 0.0f;
  const float u = rand ? rand[lidx] : philox_uniform(key, gidx);
  return 0.5f * u;
}

// srow is read through L2 (__ldcg): it may have been written by other CTAs of the same launch.
__device__ void select_and_build_mask(const float* srow, int n, int k, int soft, const float* __restrict__ rand,
                                      PhiloxKey key, uint64_t gbase, int64_t lbase, float* __restrict__ mask_row,
                                      float* thr_ptr, SelectSmem& sm) {
  const bool staged = n <= kSelKeys;
  if (staged)
    for (int j = threadIdx.x; j < n; j += blockDim.x) sm.keys[j] = order_key(__ldcg(srow + j));
  uint32_t prefix = 0, known = 0, krem = (uint32_t)k;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm.hist[i] = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const uint32_t kk = staged ? sm.keys[j] : order_key(__ldcg(srow + j));
      if ((kk & known) == prefix) atomicAdd(&sm.hist[(kk >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns bins [255-8l-7, 255-8l], walked from the top
      const int lane = threadIdx.x;
      uint32_t local = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) local += sm.hist[255 - 8 * lane - b];
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
          const uint32_t h = sm.hist[255 - 8 * lane - b];
          if (krem < cum + h) {
            sm.sel[0] = (uint32_t)(255 - 8 * lane - b);
            sm.sel[1] = krem - cum;
            break;
          }
          cum += h;
        }
      }
    }
    __syncthreads();
    prefix |= sm.sel[0] << shift;
    known |= 255u << shift;
    krem = sm.sel[1];
  }
  const float thr = key_to_float(prefix);
  if (thr_ptr && threadIdx.x == 0) *thr_ptr = thr;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float sv = staged ? key_to_float(sm.keys[j]) : __ldcg(srow + j);
    mask_row[j] = mask_value(sv, thr, soft, rand, key, gbase + j, lbase + j);
  }
}

struct SelectArgs {
  float* mask_out;           // [N,n]
  float* thr_out;            // [N] or nullptr
  const float* rand;         // [N,n] or nullptr
  PhiloxKey key;
  int64_t first_sample;
  int k;
  int soft;
  const int64_t* dyn;        // device [3] = {k, offset, first_sample} overriding the by-value fields (CUDA-graph
                             // replay: the launch is recorded once, the per-step draws change), or nullptr
};

// ------------------------------------------------------------------------------------------------
// K1, channel mode: one L-lane group per (n,c) row, all of the row's 128-bit loads issued up front,
// fp64 accumulation in four independent chains, shuffle reduction.  No barriers, no fences.
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int L>
__global__ void __launch_bounds__(kThreads, 2)
saliency_channel_kernel(const T* __restrict__ g, float* __restrict__ s, int64_t rows, int HW, int nv) {
  pdl_launch_dependents();
  const int lane = threadIdx.x & (L - 1);
  const unsigned gmask = group_mask<L>();
  const int64_t ngroups = (int64_t)gridDim.x * (kThreads / L);
  int64_t row = (int64_t)blockIdx.x * (kThreads / L) + threadIdx.x / L;
  auto reduce_store = [&](const float (&a)[kU][VEC], double (&acc)[4], int64_t r, bool last) {
#pragma unroll
    for (int j = 0; j < kU; ++j) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[(j * VEC + i) & 3] += (double)a[j][i];
    }
    if (!last) return;
    double t = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) t += __shfl_xor_sync(gmask, t, o);
    if (lane == 0) s[r] = (float)(t / (double)HW);
  };
  if (nv <= kU * L) {
    // a row is one batch (the latent shapes of the model: 14x14, 16x16, 28x28): persistent groups stride over the rows
    // with the NEXT row's 128-bit loads issued before the current row is reduced -- short-lived CTAs (one row per group)
    // left the kernel at 3.8 TB/s
    float cur[kU][VEC], nxt[kU][VEC];
    if (row < rows) load_batch<T, VEC, L>(g + row * HW, 0, lane, nv, cur);
    for (; row < rows; row += ngroups) {               // group-uniform
      const int64_t rn = row + ngroups;
      const bool more = rn < rows;
      if (more) load_batch<T, VEC, L>(g + rn * HW, 0, lane, nv, nxt);
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      reduce_store(cur, acc, row, true);
      if (more) {
#pragma unroll
        for (int j = 0; j < kU; ++j) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) cur[j][i] = nxt[j][i];
        }
      }
    }
    return;
  }
  for (; row < rows; row += ngroups) {
    const T* __restrict__ p = g + row * HW;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int base = 0; base < nv; base += kU * L) {
      float a[kU][VEC];
      load_batch<T, VEC, L>(p, base, lane, nv, a);
      reduce_store(a, acc, row, base + kU * L >= nv);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1, spatial mode: CTA = X vector columns x Y channel slices of one sample.  A warp reads X*16
// contiguous bytes of one channel row; each thread keeps kU channel rows in flight; the Y partial
// sums meet in shared memory (fp64).
// ------------------------------------------------------------------------------------------------
template <typename T, int VEC, int X, int Y>
__global__ void __launch_bounds__(X * Y, 3)
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

// stand-alone selection for ctl_topp_mask_apply (s supplied by the caller): one CTA per sample
__global__ void __launch_bounds__(kThreads)
select_mask_kernel(const float* s, int n, SelectArgs sa) {
  pdl_launch_dependents();                 // lets K2 become resident and request its rows of z
  pdl_wait();                              // s comes from the preceding launch (K1) when chained
  __shared__ SelectSmem sm;
  const int64_t sample = blockIdx.x;
  if (sa.dyn) {                            // host-validated by the caller; clamped so a stale buffer cannot run wild
    const int64_t kd = sa.dyn[0];
    sa.k = (int)(kd < 0 ? 0 : kd >= n ? n - 1 : kd);
    sa.key.offset = (uint64_t)sa.dyn[1];
    sa.first_sample = sa.dyn[2];
  }
  select_and_build_mask(s + sample * n, n, sa.k, sa.soft, sa.rand, sa.key,
                        (uint64_t)(sa.first_sample + sample) * (uint64_t)n, sample * (int64_t)n,
                        sa.mask_out + sample * (int64_t)n, sa.thr_out ? sa.thr_out + sample : nullptr, sm);
}

// coherent (ld.global, not ld.global.nc) vector read of mask values: under programmatic dependent launch
// this kernel is already resident while K1 still writes the mask, so the read-only path must not be used
template <int VEC>
__device__ __forceinline__ void load_mask(const float* p, float (&m)[VEC]) {
  if constexpr (VEC % 4 == 0) {
#pragma unroll
    for (int q = 0; q < VEC / 4; ++q) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(p) + q);
      m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < VEC; ++i) m[i] = __ldcg(p + i);
  }
}

// ------------------------------------------------------------------------------------------------
// K2: z~ = z * mask.  Pure streaming: one L-lane group per (n,c) row, kU 128-bit loads in flight per
// lane.  The row's data is requested BEFORE the dependency wait -- z does not depend on K1, the mask does.
// ------------------------------------------------------------------------------------------------
template <typename ZT, typename OT, int VEC, int L, int MODE>
__global__ void __launch_bounds__(kThreads, 4)
mask_apply_kernel(const ZT* __restrict__ z, OT* __restrict__ z_out, const float* mask, int64_t rows, int C, int HW,
                  int nv) {
  const int lane = threadIdx.x & (L - 1);
  const int64_t row = (int64_t)blockIdx.x * (kThreads / L) + threadIdx.x / L;
  const bool active = row < rows;
  const ZT* __restrict__ zi = z + row * HW;
  OT* __restrict__ zo = z_out + row * HW;
  float a[kU][VEC];
  if (active) load_batch<ZT, VEC, L>(zi, 0, lane, nv, a);
  pdl_wait();                                            // the mask comes from the preceding launch
  if (!active) return;
  const int64_t sample = row / C;
  const float mrow = (MODE == CTL_MODE_CHANNEL) ? __ldcg(mask + row) : 1.0f;
  const float* mvec = mask + sample * HW;
  for (int vb = 0; vb < nv; vb += kU * L) {
    if (vb != 0) load_batch<ZT, VEC, L>(zi, vb, lane, nv, a);
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int v = vb + j * L + lane;
      if (v < nv) {
        if (MODE == CTL_MODE_CHANNEL) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) a[j][i] *= mrow;
        } else {
          float m[VEC];
          load_mask<VEC>(mvec + v * VEC, m);
#pragma unroll
          for (int i = 0; i < VEC; ++i) a[j][i] *= m[i];
        }
        store_from_float<OT, VEC>(zo + (int64_t)v * VEC, a[j]);
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
                       int64_t rows, int HW, int nv, float p, float scale, uint64_t first_row, const int64_t* dyn,
                       int C) {
  if (dyn) {                                             // {-, offset, first_sample} from device memory (graph replay)
    key.offset = (uint64_t)dyn[1];
    first_row = (uint64_t)dyn[2] * (uint64_t)C;
  }
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

// ------------------------------------------------------------------------------------------------
// Fused tail of the latent masking for the training path: ONE CTA per sample takes the
// per-sample saliency SUMS that the decoder's last input-gradient convolution accumulated in its epilogue
// (conv_tc.cu, fp64, dL/dz itself never reaches HBM), turns them into s exactly as K1 does (fp64 sum / count, rounded
// to fp32 once), selects the k-th largest, builds the mask row and applies it to the sample's code -- written BOTH as
// the NCHW fp32 tensor the reference API returns and as the blocked bf16 C8 tensor the decoder's first convolution
// consumes (no separate layout conversion in front of decoder_inference).
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads)
sums_select_apply_kernel(const double* __restrict__ sums, const float* __restrict__ z, int C, int HW, SelectArgs sa,
                         float* __restrict__ s_out, float* __restrict__ z_out, __nv_bfloat16* __restrict__ z_c8) {
  __shared__ SelectSmem sm;
  const int64_t sample = blockIdx.x;
  const int n = MODE == CTL_MODE_CHANNEL ? C : HW;
  const double count = (double)(MODE == CTL_MODE_CHANNEL ? HW : C);
  if (sa.dyn) {
    const int64_t kd = sa.dyn[0];
    sa.k = (int)(kd < 0 ? 0 : kd >= n ? n - 1 : kd);
    sa.key.offset = (uint64_t)sa.dyn[1];
    sa.first_sample = sa.dyn[2];
  }
  float* srow = s_out + sample * n;
  for (int j = threadIdx.x; j < n; j += blockDim.x) srow[j] = (float)(sums[sample * n + j] / count);
  __syncthreads();
  float* mrow = sa.mask_out + sample * (int64_t)n;
  select_and_build_mask(srow, n, sa.k, sa.soft, sa.rand, sa.key, (uint64_t)(sa.first_sample + sample) * (uint64_t)n,
                        sample * (int64_t)n, mrow, sa.thr_out ? sa.thr_out + sample : nullptr, sm);
  __syncthreads();
  const float* __restrict__ zs = z + sample * (int64_t)C * HW;
  float* __restrict__ zo = z_out + sample * (int64_t)C * HW;
  const int groups = C >> 3;                                  // C % 8 == 0 (checked on the host)
  // gridDim.y CTAs share a sample: each repeats the (tiny, deterministic) selection -- they write identical s / mask / thr
  // values -- and applies its share of the code; one CTA per sample left 84 of 148 SMs idle and took 34 us at [64,128,14,14]
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < groups * HW; idx += gridDim.y * blockDim.x) {
    const int gq = idx / HW, pix = idx - gq * HW;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = __ldcs(zs + (int64_t)(gq * 8 + j) * HW + pix);
    const float mp = MODE == CTL_MODE_SPATIAL ? __ldcg(mrow + pix) : 1.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] *= MODE == CTL_MODE_CHANNEL ? __ldcg(mrow + gq * 8 + j) : mp;
      zo[(int64_t)(gq * 8 + j) * HW + pix] = v[j];
    }
    if (z_c8) store_from_float<__nv_bfloat16, 8>(z_c8 + ((sample * groups + gq) * (int64_t)HW + pix) * 8, v);
  }
}

// ---- launch helpers ------------------------------------------------------------------------------
// lanes per row: enough rows in flight per CTA, yet most of a row covered by one kU-deep batch
inline int pick_lanes(int nv) { return nv >= 128 ? 32 : nv >= 64 ? 16 : nv >= 16 ? 8 : 4; }

template <typename T, int VEC>
int launch_saliency_channel(const T* g, float* s, int64_t N, int C, int HW, cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int64_t rows = N * C;
  // persistent groups: two CTAs per SM stride over the rows (fewer when there are not that many rows)
  const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(rows, kThreads / L), (int64_t)sm_count() * 2));
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

int launch_saliency(const void* g, int g_dtype, int64_t N, int64_t C, int64_t HW, int mode, float* s,
                    cudaStream_t st) {
  const bool f32 = g_dtype == CTL_F32;
  const int vw = f32 ? 4 : 8;
  const bool vec = (HW % vw == 0) && aligned16(g);
  const int Ci = (int)C, HWi = (int)HW;
  if (mode == CTL_MODE_CHANNEL) {
    if (f32) return vec ? launch_saliency_channel<float, 4>((const float*)g, s, N, Ci, HWi, st)
                        : launch_saliency_channel<float, 1>((const float*)g, s, N, Ci, HWi, st);
    return vec ? launch_saliency_channel<__nv_bfloat16, 8>((const __nv_bfloat16*)g, s, N, Ci, HWi, st)
               : launch_saliency_channel<__nv_bfloat16, 1>((const __nv_bfloat16*)g, s, N, Ci, HWi, st);
  }
  if (f32) return vec ? launch_saliency_spatial<float, 4>((const float*)g, s, N, Ci, HWi, st)
                      : launch_saliency_spatial<float, 1>((const float*)g, s, N, Ci, HWi, st);
  return vec ? launch_saliency_spatial<__nv_bfloat16, 8>((const __nv_bfloat16*)g, s, N, Ci, HWi, st)
             : launch_saliency_spatial<__nv_bfloat16, 1>((const __nv_bfloat16*)g, s, N, Ci, HWi, st);
}

// Programmatic dependent launch can be switched off (CTL_PDL=0) for debugging.
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("CTL_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

// one CTA per sample; chained == true: programmatic stream serialization after K1
int launch_select(const float* s, int64_t N, int64_t n, const SelectArgs& sa, bool chained, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)N);
  cfg.blockDim = dim3(kThreads);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (chained && pdl_enabled()) ? 1 : 0;
  CTL_CUDA_OK(cudaLaunchKernelEx(&cfg, select_mask_kernel, s, (int)n, sa), "select_mask launch");
  return CTL_OK;
}

// K2 is always launched with programmatic stream serialization: its CTAs become resident while the
// select kernel drains, request their rows of z and only then wait for the mask.  Safe because the head
// of the chain (K1, or the select kernel when s is supplied) is an ORDINARY launch: everything older,
// including whatever wrote z, completed before the chain started, and no kernel of the chain writes z.
template <typename ZT, typename OT, int VEC, int L, int MODE>
int launch_apply_kernel(const ZT* z, OT* z_out, const float* mask, int64_t rows, int C, int HW, int nv,
                        cudaStream_t st) {
  const int64_t grid64 = ceil_div(rows, kThreads / L);
  CTL_REQUIRE(grid64 <= 0x7fffffff, CTL_ERR_UNSUPPORTED, "too many rows for one launch");
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid64);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  CTL_CUDA_OK(cudaLaunchKernelEx(&cfg, mask_apply_kernel<ZT, OT, VEC, L, MODE>, z, z_out, mask, rows, C, HW, nv),
              "mask_apply launch");
  return CTL_OK;
}

template <typename ZT, typename OT, int VEC>
int launch_apply(int mode, const ZT* z, OT* z_out, const float* mask, int64_t N, int C, int HW, cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int64_t rows = N * C;
#define CTL_K2(LL)                                                                                              \
  return mode == CTL_MODE_CHANNEL                                                                               \
             ? launch_apply_kernel<ZT, OT, VEC, LL, CTL_MODE_CHANNEL>(z, z_out, mask, rows, C, HW, nv, st)       \
             : launch_apply_kernel<ZT, OT, VEC, LL, CTL_MODE_SPATIAL>(z, z_out, mask, rows, C, HW, nv, st)
  switch (L) {
    case 32: CTL_K2(32);
    case 16: CTL_K2(16);
    case 8: CTL_K2(8);
    default: CTL_K2(4);
  }
#undef CTL_K2
}

int launch_apply_any(int mode, const void* z, int z_dtype, void* z_out, int out_dtype, const float* mask, int64_t N,
                     int64_t C, int64_t HW, cudaStream_t st) {
  const bool zf = z_dtype == CTL_F32, of = out_dtype == CTL_F32;
  const int vw = zf ? 4 : 8;
  const bool vec = (HW % vw == 0) && aligned16(z) && aligned16(z_out) && aligned16(mask);
  const int Ci = (int)C, HWi = (int)HW;
#define CTL_APPLY(ZT, OT, V) return launch_apply<ZT, OT, V>(mode, (const ZT*)z, (OT*)z_out, mask, N, Ci, HWi, st)
  if (zf && of) { if (vec) CTL_APPLY(float, float, 4); else CTL_APPLY(float, float, 1); }
  if (zf && !of) { if (vec) CTL_APPLY(float, __nv_bfloat16, 4); else CTL_APPLY(float, __nv_bfloat16, 1); }
  if (!zf && of) { if (vec) CTL_APPLY(__nv_bfloat16, float, 8); else CTL_APPLY(__nv_bfloat16, float, 1); }
  if (vec) CTL_APPLY(__nv_bfloat16, __nv_bfloat16, 8); else CTL_APPLY(__nv_bfloat16, __nv_bfloat16, 1);
#undef CTL_APPLY
}

template <typename ZT, typename OT, int VEC>
int launch_dropout(const ZT* z, OT* z_out, float* mask_out, const float* keep, float* keep_out, PhiloxKey key,
                   int64_t rows, int HW, float p, float scale, uint64_t first_row, const int64_t* dyn, int C,
                   cudaStream_t st) {
  const int nv = HW / VEC;
  const int L = pick_lanes(nv);
  const int64_t grid64 = ceil_div(rows, kThreads / L);
  CTL_REQUIRE(grid64 <= 0x7fffffff, CTL_ERR_UNSUPPORTED, "too many rows for one launch");
  const unsigned grid = (unsigned)grid64;
  switch (L) {
    case 32: channel_dropout_kernel<ZT, OT, VEC, 32><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row, dyn, C); break;
    case 16: channel_dropout_kernel<ZT, OT, VEC, 16><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row, dyn, C); break;
    case 8: channel_dropout_kernel<ZT, OT, VEC, 8><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row, dyn, C); break;
    default: channel_dropout_kernel<ZT, OT, VEC, 4><<<grid, kThreads, 0, st>>>(z, z_out, mask_out, keep, keep_out, key, rows, HW, nv, p, scale, first_row, dyn, C); break;
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

int check_select(int64_t n, int64_t k, int64_t first_sample) {
  CTL_REQUIRE(k >= 0, CTL_ERR_INVALID, "k must be >= 0 (got %lld)", (long long)k);
  CTL_REQUIRE(k < n, CTL_ERR_INDEX, "index %lld is out of bounds for dimension 1 with size %lld", (long long)k,
              (long long)n);
  CTL_REQUIRE(first_sample >= 0, CTL_ERR_INVALID, "first_sample must be >= 0");
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
  return launch_saliency(g, g_dtype, N, C, HW, mode, s_out, (cudaStream_t)stream);
}

extern "C" int ctl_topp_mask_apply(const float* s, const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW,
                                   int mode, int64_t k, int soft, const float* rand, uint64_t seed, uint64_t offset,
                                   int64_t first_sample, float* mask_out, float* thr_out, void* z_out, int out_dtype,
                                   void* stream) {
  CTL_REQUIRE(s && z && mask_out && z_out, CTL_ERR_INVALID, "ctl_topp_mask_apply: NULL pointer");
  CTL_REQUIRE(valid_dtype(z_dtype) && valid_dtype(out_dtype), CTL_ERR_INVALID, "ctl_topp_mask_apply: unknown dtype");
  CTL_REQUIRE(mode == CTL_MODE_CHANNEL || mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown mode %d", mode);
  if (int rc = check_shape(N, C, HW)) return rc;
  const int64_t n = mode == CTL_MODE_CHANNEL ? C : HW;
  if (int rc = check_select(n, k, first_sample)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  SelectArgs sa = {mask_out, thr_out, rand, PhiloxKey{seed, offset}, first_sample, (int)k, soft != 0, nullptr};
  if (int rc = launch_select(s, N, n, sa, /*chained=*/false, st)) return rc;
  return launch_apply_any(mode, z, z_dtype, z_out, out_dtype, mask_out, N, C, HW, st);
}

static int saliency_mask_apply_impl(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N, int64_t C,
                                       int64_t HW, int mode, int64_t k, int soft, const float* rand, uint64_t seed,
                                       uint64_t offset, int64_t first_sample, float* s_scratch, float* mask_out,
                                       float* thr_out, void* z_out, int out_dtype, const int64_t* dyn,
                                       void* stream) {
  CTL_REQUIRE(g && z && s_scratch && mask_out && z_out, CTL_ERR_INVALID,
              "ctl_saliency_mask_apply: NULL pointer");
  CTL_REQUIRE(valid_dtype(g_dtype) && valid_dtype(z_dtype) && valid_dtype(out_dtype), CTL_ERR_INVALID,
              "ctl_saliency_mask_apply: unknown dtype");
  CTL_REQUIRE(mode == CTL_MODE_CHANNEL || mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown mode %d", mode);
  if (int rc = check_shape(N, C, HW)) return rc;
  const int64_t n = mode == CTL_MODE_CHANNEL ? C : HW;
  // validated before anything is launched: the reference raises IndexError before masking
  if (int rc = check_select(n, k, first_sample)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  SelectArgs sa = {mask_out, thr_out, rand, PhiloxKey{seed, offset}, first_sample, (int)k, soft != 0, dyn};
  if (int rc = launch_saliency(g, g_dtype, N, C, HW, mode, s_scratch, st)) return rc;      // ordinary launch
  if (int rc = launch_select(s_scratch, N, n, sa, /*chained=*/true, st)) return rc;         // PDL after K1
  return launch_apply_any(mode, z, z_dtype, z_out, out_dtype, mask_out, N, C, HW, st);
}

extern "C" int ctl_saliency_mask_apply(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N, int64_t C,
                                       int64_t HW, int mode, int64_t k, int soft, const float* rand, uint64_t seed,
                                       uint64_t offset, int64_t first_sample, float* s_scratch, float* mask_out,
                                       float* thr_out, void* z_out, int out_dtype, void* stream) {
  return saliency_mask_apply_impl(g, g_dtype, z, z_dtype, N, C, HW, mode, k, soft, rand, seed, offset, first_sample,
                                  s_scratch, mask_out, thr_out, z_out, out_dtype, nullptr, stream);
}

extern "C" int ctl_saliency_mask_apply_dyn(const void* g, int g_dtype, const void* z, int z_dtype, int64_t N,
                                           int64_t C, int64_t HW, int mode, int soft, const float* rand,
                                           uint64_t seed, const int64_t* step_params, float* s_scratch,
                                           float* mask_out, float* thr_out, void* z_out, int out_dtype, void* stream) {
  CTL_REQUIRE(step_params, CTL_ERR_INVALID, "ctl_saliency_mask_apply_dyn: NULL step_params");
  return saliency_mask_apply_impl(g, g_dtype, z, z_dtype, N, C, HW, mode, /*k=*/0, soft, rand, seed, 0, 0, s_scratch,
                                  mask_out, thr_out, z_out, out_dtype, step_params, stream);
}

extern "C" int ctl_saliency_sums_mask_apply(const double* sums, const float* z, int64_t N, int64_t C, int64_t HW, int mode,
                                            int64_t k, int soft, const float* rand, uint64_t seed, uint64_t offset,
                                            int64_t first_sample, const int64_t* step_params, float* s_out,
                                            float* mask_out, float* thr_out, float* z_out, void* z_c8_out, void* stream) {
  CTL_REQUIRE(sums && z && s_out && mask_out && z_out, CTL_ERR_INVALID, "ctl_saliency_sums_mask_apply: NULL pointer");
  CTL_REQUIRE(mode == CTL_MODE_CHANNEL || mode == CTL_MODE_SPATIAL, CTL_ERR_INVALID, "unknown mode %d", mode);
  if (int rc = check_shape(N, C, HW)) return rc;
  CTL_REQUIRE(C % 8 == 0, CTL_ERR_UNSUPPORTED, "ctl_saliency_sums_mask_apply: C must be a multiple of 8 (got %lld)", (long long)C);
  CTL_REQUIRE(!z_c8_out || aligned16(z_c8_out), CTL_ERR_INVALID, "ctl_saliency_sums_mask_apply: z_c8_out must be 16-byte aligned");
  const int64_t n = mode == CTL_MODE_CHANNEL ? C : HW;
  if (!step_params)
    if (int rc = check_select(n, k, first_sample)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  SelectArgs sa = {mask_out, thr_out, rand, PhiloxKey{seed, offset}, first_sample, (int)k, soft != 0, step_params};
  // CTAs per sample: enough to put ~4 CTAs on every SM, at least ~2 apply iterations per thread
  const int64_t work = (C / 8) * HW;
  const unsigned parts = (unsigned)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(ceil_div((int64_t)sm_count() * 4, N),
                                                                                         ceil_div(work, 2 * kThreads)), 32));
  const dim3 grid((unsigned)N, parts);
  if (mode == CTL_MODE_CHANNEL)
    sums_select_apply_kernel<CTL_MODE_CHANNEL><<<grid, kThreads, 0, st>>>(sums, z, (int)C, (int)HW, sa, s_out, z_out,
                                                                          (__nv_bfloat16*)z_c8_out);
  else
    sums_select_apply_kernel<CTL_MODE_SPATIAL><<<grid, kThreads, 0, st>>>(sums, z, (int)C, (int)HW, sa, s_out, z_out,
                                                                          (__nv_bfloat16*)z_c8_out);
  CTL_CUDA_OK(cudaGetLastError(), "sums_select_apply launch");
  return CTL_OK;
}

static int channel_dropout_impl(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p, float scale,
                                   const float* keep, uint64_t seed, uint64_t offset, int64_t first_sample,
                                   void* z_out, int out_dtype, float* mask_out, float* keep_out, const int64_t* dyn,
                                   void* stream) {
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
                                   first_row, dyn, (int)C, st)
  if (zf && of) { if (vec) CTL_DROP(float, float, 4); else CTL_DROP(float, float, 1); }
  if (zf && !of) { if (vec) CTL_DROP(float, __nv_bfloat16, 4); else CTL_DROP(float, __nv_bfloat16, 1); }
  if (!zf && of) { if (vec) CTL_DROP(__nv_bfloat16, float, 8); else CTL_DROP(__nv_bfloat16, float, 1); }
  if (vec) CTL_DROP(__nv_bfloat16, __nv_bfloat16, 8); else CTL_DROP(__nv_bfloat16, __nv_bfloat16, 1);
#undef CTL_DROP
}

extern "C" int ctl_channel_dropout(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p, float scale,
                                   const float* keep, uint64_t seed, uint64_t offset, int64_t first_sample,
                                   void* z_out, int out_dtype, float* mask_out, float* keep_out, void* stream) {
  return channel_dropout_impl(z, z_dtype, N, C, HW, p, scale, keep, seed, offset, first_sample, z_out, out_dtype,
                              mask_out, keep_out, nullptr, stream);
}

extern "C" int ctl_channel_dropout_dyn(const void* z, int z_dtype, int64_t N, int64_t C, int64_t HW, float p,
                                       float scale, uint64_t seed, const int64_t* step_params, void* z_out,
                                       int out_dtype, float* mask_out, float* keep_out, void* stream) {
  CTL_REQUIRE(step_params, CTL_ERR_INVALID, "ctl_channel_dropout_dyn: NULL step_params");
  return channel_dropout_impl(z, z_dtype, N, C, HW, p, scale, nullptr, seed, 0, 0, z_out, out_dtype, mask_out,
                              keep_out, step_params, stream);
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

// CUDA-core kernels around K3 for the blocked activation layout C8 = bf16 [N][C/8][H][W][8]:
// layout conversion, the stem convolution (Cin = 1 / 4: K = 9 / 36 is below what a UMMA tile can use, SURVEY.md
// 7.3 #4) with the STN input construction fused into its prologue, the 1x1 head (Cout = 1 / 4), nearest x2
// up-sampling and the per-channel scale/shift/activation pass (train-mode BatchNorm statistics: c8_bwd.cu).
// All of them are HBM-bound streaming kernels: 16-byte accesses, consecutive threads on consecutive pixels.
//
// Reference code replaced:
//   stem  : MyEncoder.inc[0] + norm + LeakyReLU           medseg/models/ebm/encoder_decoder.py:370-378
//           construct_input (softmax(logit/T) | one-hot)   medseg/common_utils/basic_operations.py:110-158
//   head  : MyDecoder.final_conv (+ Sigmoid)               medseg/models/ebm/encoder_decoder.py:439-452
//   up2x  : nn.UpsamplingNearest2d(scale_factor=2)         medseg/models/ebm/encoder_decoder.py:294-296
//   BN    : nn.BatchNorm2d training statistics             (norm(out_ch) in every block)
#include <algorithm>

#include "ctl_common.cuh"

namespace ctl {
namespace {

constexpr int kT = 256;

__device__ __forceinline__ float act_fn(float v, int act) {
  switch (act) {
    case CTL_ACT_LRELU: return v > 0.0f ? v : 0.2f * v;
    case CTL_ACT_RELU: return fmaxf(v, 0.0f);
    case CTL_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-v));
    default: return v;
  }
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(w[i] << 16); f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    o[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------ layout
template <typename T>
__global__ void __launch_bounds__(kT) nchw_to_c8_kernel(const T* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                         int64_t total /*N*C/8*HW*/, int C8, int64_t HW) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pix = i % HW, nc = i / HW;             // nc = n*C8 + c8
    const T* src = x + nc * 8 * HW + pix;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = to_f<T>(src[j * HW]);
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(f);
  }
}
template <typename T>
__global__ void __launch_bounds__(kT) c8_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, T* __restrict__ y,
                                                         int64_t total, int C8, int64_t HW) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pix = i % HW, nc = i / HW;
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(x + i * 8), f);
    T* dst = y + nc * 8 * HW + pix;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j * HW] = (T)f[j];
  }
}

// ------------------------------------------------------------------------------------------------ stem
// 3x3 pad-1 conv from a planar (NCHW) input with CIN <= 4 channels to COUT (16) blocked channels, then
// y = act(acc*scale + shift).  in_mode: 0 = fp32 NCHW as is, 1 = softmax(x / temperature) over the CIN channels
// (STN input built from logits), 2 = one-hot of an int64 label map [N,H,W] (STN input built from labels).
// CTA = one 32x8 pixel tile per iteration (persistent): the input tile with its halo is built ONCE in shared memory
// (softmax / one-hot evaluated once per pixel instead of nine times), then every thread owns one output pixel.
template <int CIN, int COUT>
__global__ void __launch_bounds__(kT)
stem_conv_kernel(const float* __restrict__ x, const long long* __restrict__ labels, const float* __restrict__ w /*[COUT][CIN][3][3]*/,
                 const float* __restrict__ scale, const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, int N,
                 int H, int W, int in_mode, float inv_temp, int act) {
  pdl_entry();
  constexpr int TW = 32, TH = 8, HW_T = TW + 2, HH_T = TH + 2;
  __shared__ __align__(16) float sw[COUT * CIN * 9];
  __shared__ float ssc[COUT], ssh[COUT];
  __shared__ float s_in[CIN][HH_T][HW_T + 1];
  // transposed to [ci][tap][COUT] so that one 16-byte broadcast read feeds four output channels
  for (int i = threadIdx.x; i < COUT * CIN * 9; i += kT) {
    const int c = i / (CIN * 9), rest = i - c * (CIN * 9);
    sw[rest * COUT + c] = w[i];
  }
  for (int i = threadIdx.x; i < COUT; i += kT) { ssc[i] = scale ? scale[i] : 1.0f; ssh[i] = shift ? shift[i] : 0.0f; }
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int64_t HW = (int64_t)H * W, num_tiles = (int64_t)N * tiles_x * tiles_y;
  const int py = threadIdx.x / TW, px = threadIdx.x % TW;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t n = tile / (tiles_x * tiles_y);
    const int rem = (int)(tile - n * tiles_x * tiles_y);
    const int y0 = (rem / tiles_x) * TH, x0 = (rem % tiles_x) * TW;
    __syncthreads();                                      // previous tile's readers are done (and sw is visible)
    for (int i = threadIdx.x; i < HH_T * HW_T; i += kT) {
      const int hy = i / HW_T, hx = i - hy * HW_T;
      const int iy = y0 + hy - 1, ix = x0 + hx - 1;
      float v[CIN];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) v[ci] = 0.0f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
        const int64_t q = (int64_t)iy * W + ix;
        if (in_mode == 2) {
          const long long lab = labels[n * HW + q];
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) v[ci] = (lab == ci) ? 1.0f : 0.0f;
        } else {
#pragma unroll
          for (int ci = 0; ci < CIN; ++ci) v[ci] = x[(n * CIN + ci) * HW + q];
          if (in_mode == 1) {
            float mx = v[0];
#pragma unroll
            for (int ci = 1; ci < CIN; ++ci) mx = fmaxf(mx, v[ci]);
            float sum = 0.0f;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) { v[ci] = __expf((v[ci] - mx) * inv_temp); sum += v[ci]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) v[ci] *= inv;
          }
        }
      }
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) s_in[ci][hy][hx] = v[ci];
    }
    __syncthreads();
    const int yy = y0 + py, xx = x0 + px;
    if (yy < H && xx < W) {
      float acc[COUT];
#pragma unroll
      for (int c = 0; c < COUT; ++c) acc[c] = 0.0f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) {
            const float v = s_in[ci][py + r][px + s2];
            const float4* wv = reinterpret_cast<const float4*>(sw + (ci * 9 + r * 3 + s2) * COUT);
#pragma unroll
            for (int c4 = 0; c4 < COUT / 4; ++c4) {
              const float4 ww = wv[c4];
              acc[4 * c4] = fmaf(v, ww.x, acc[4 * c4]);
              acc[4 * c4 + 1] = fmaf(v, ww.y, acc[4 * c4 + 1]);
              acc[4 * c4 + 2] = fmaf(v, ww.z, acc[4 * c4 + 2]);
              acc[4 * c4 + 3] = fmaf(v, ww.w, acc[4 * c4 + 3]);
            }
          }
        }
      }
      const int64_t pix = (int64_t)yy * W + xx;
#pragma unroll
      for (int c8 = 0; c8 < COUT / 8; ++c8) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fn(acc[c8 * 8 + j] * ssc[c8 * 8 + j] + ssh[c8 * 8 + j], act);
        *reinterpret_cast<uint4*>(y + ((n * (COUT / 8) + c8) * HW + pix) * 8) = pack8(f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ head
// 1x1 conv from CIN (16) blocked channels to COUT <= 4 planar fp32 channels (+ bias, optional sigmoid).
template <int CIN>
__global__ void __launch_bounds__(kT)
head_conv_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w /*[COUT][CIN]*/, const float* __restrict__ b,
                 float* __restrict__ y, int N, int64_t HW, int COUT, int act) {
  pdl_entry();
  __shared__ float sw[4 * CIN];
  __shared__ float sb[4];
  for (int i = threadIdx.x; i < COUT * CIN; i += kT) sw[i] = w[i];
  if (threadIdx.x < COUT) sb[threadIdx.x] = b ? b[threadIdx.x] : 0.0f;
  __syncthreads();
  const int64_t total = (int64_t)N * HW;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW, pix = i - n * HW;
    float v[CIN];
#pragma unroll
    for (int c8 = 0; c8 < CIN / 8; ++c8) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + ((n * (CIN / 8) + c8) * HW + pix) * 8), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[c8 * 8 + j] = f[j];
    }
    for (int co = 0; co < COUT; ++co) {
      float a = sb[co];
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) a = fmaf(v[ci], sw[co * CIN + ci], a);
      y[(n * COUT + co) * HW + pix] = act_fn(a, act);
    }
  }
}

// ------------------------------------------------------------------------------------------------ up2x
__global__ void __launch_bounds__(kT)
upsample2x_c8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t planes, int H, int W) {
  pdl_entry();
  const int64_t total = planes * H * W;      // one thread per INPUT pixel: 16 B in, 4 x 16 B out
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pl = i / ((int64_t)H * W);
    const int pix = (int)(i - pl * H * W);
    const int yy = pix / W, xx = pix - yy * W;
    const uint4 v = x[i];
    uint4* o = y + pl * 4 * H * W + (int64_t)(2 * yy) * (2 * W) + 2 * xx;
    o[0] = v; o[1] = v; o[2 * W] = v; o[2 * W + 1] = v;
  }
}

// Train-mode BatchNorm finalisation folded into the kernel that applies it: when `sums` is given, every CTA forms
// scale = gamma * rsqrt(var + eps), shift = beta - mean * scale for all C <= kMaxBnC channels in its prologue from the
// per-channel fp64 sums the convolution epilogue accumulated (sum x | sum x^2) -- fp64 only for mean / variance (no fp64
// division or square root) -- and CTA 0 also publishes scale / shift / mean / var for the backward and updates the
// running statistics exactly like bn_fwd_from_sums_kernel.  No finalisation launch between the convolution and the apply.
constexpr int kMaxBnC = 256;
struct BnFwdSums {
  const double* sums;          // [2][C] or nullptr (then scale / shift are read from global memory)
  double inv_count, unbias;    // 1 / count, count / (count - 1)
  const float* gamma;
  const float* beta;
  float* scale_out;
  float* shift_out;
  float* mean_out;
  float* var_out;
  float* running_mean;         // may be nullptr
  float* running_var;
  float eps, momentum;
};

__device__ __forceinline__ void bn_fwd_prologue(const BnFwdSums& bn, int C, float* s_scale, float* s_shift) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = bn.sums[c] * bn.inv_count;
    double var = bn.sums[C + c] * bn.inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = 1.0f / sqrtf((float)var + bn.eps);
    const float g = bn.gamma ? bn.gamma[c] : 1.0f, b = bn.beta ? bn.beta[c] : 0.0f;
    const float sc = g * inv, sh = b - (float)mean * sc;
    s_scale[c] = sc;
    s_shift[c] = sh;
    if (blockIdx.x == 0) {
      bn.scale_out[c] = sc;
      bn.shift_out[c] = sh;
      if (bn.mean_out) bn.mean_out[c] = (float)mean;
      if (bn.var_out) bn.var_out[c] = (float)var;
      if (bn.running_mean) {
        bn.running_mean[c] = (1.0f - bn.momentum) * bn.running_mean[c] + bn.momentum * (float)mean;
        bn.running_var[c] = (1.0f - bn.momentum) * bn.running_var[c] + bn.momentum * (float)(var * bn.unbias);
      }
    }
  }
  __syncthreads();
}

// y = act(x*scale[c] + shift[c]), blocked -> blocked
__global__ void __launch_bounds__(kT)
scale_shift_act_c8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ scale,
                          const float* __restrict__ shift, int64_t total /*N*C8*HW*/, int C8, int64_t HW, int act,
                          const BnFwdSums bn) {
  pdl_entry();
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC];
  if (bn.sums != nullptr) {
    bn_fwd_prologue(bn, C8 * 8, s_scale, s_shift);
    scale = s_scale;
    shift = s_shift;
  }
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c8 = (int)((i / HW) % C8);
    float f[8];
    unpack8(__ldcs(x + i), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = act_fn(f[j] * scale[c8 * 8 + j] + shift[c8 * 8 + j], act);
    y[i] = pack8(f);
  }
}

// The same pass for large planes (HW >= 2048), structured by plane: a work item is U*kT consecutive positions of one
// 8-channel plane -- scale / shift of its channels in registers, no 64-bit division and no per-position parameter
// loads, all U loads of an item in flight before the first use.
template <int U>
__global__ void __launch_bounds__(kT, 4)
scale_shift_act_planes_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, const float* __restrict__ scale,
                              const float* __restrict__ shift, int64_t planes, int C8, int HW, int items_per_plane, int act,
                              const BnFwdSums bn) {
  pdl_entry();
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC];
  if (bn.sums != nullptr) {
    bn_fwd_prologue(bn, C8 * 8, s_scale, s_shift);
    scale = s_scale;
    shift = s_shift;
  }
  const int64_t items = planes * items_per_plane;
  float sc[8], sh[8];
  int cur_c0 = -1;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t plane = item / items_per_plane;
    const int chunk = (int)(item - plane * items_per_plane);
    const int c0 = (int)(plane % C8) * 8;
    if (c0 != cur_c0) {                                   // block-uniform
      cur_c0 = c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = scale ? scale[c0 + j] : 1.0f; sh[j] = shift ? shift[c0 + j] : 0.0f; }
    }
    const int p0 = chunk * (U * kT) + threadIdx.x;
    const int64_t base = plane * HW;
    uint4 xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = p0 + u * kT < HW ? __ldcs(x + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u * kT < HW) {
        float f[8];
        unpack8(xv[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fn(fmaf(f[j], sc[j], sh[j]), act);
        y[base + p0 + u * kT] = pack8(f);
      }
    }
  }
}

// launches the plane-structured form for planes of >= 2048 positions, else the generic one
int launch_scale_shift_act(const uint4* x, uint4* y, const float* scale, const float* shift, int64_t planes, int C8,
                           int64_t HW, int act, const BnFwdSums& bn, cudaStream_t st) {
  const int64_t total = planes * HW;
  if (HW >= 2048 && HW < ((int64_t)1 << 30)) {
    static int resident[2] = {0, 0};
    const int U = HW >= 8192 ? 4 : 2;
    int& res = resident[U == 4];
    if (res == 0) {
      if (U == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, scale_shift_act_planes_kernel<4>, kT, 0);
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, scale_shift_act_planes_kernel<2>, kT, 0);
      if (res <= 0) res = 4;
    }
    const int ipp = (int)ceil_div(HW, (int64_t)U * kT);
    const unsigned grid = (unsigned)std::min<int64_t>(planes * ipp, (int64_t)sm_count() * res);
    if (U == 4) launch_chained(scale_shift_act_planes_kernel<4>, grid, kT, 0, st)(x, y, scale, shift, planes, C8, (int)HW, ipp, act, bn);
    else launch_chained(scale_shift_act_planes_kernel<2>, grid, kT, 0, st)(x, y, scale, shift, planes, C8, (int)HW, ipp, act, bn);
  } else {
    const unsigned grid = bn.sums != nullptr
                              ? (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, kT), (int64_t)sm_count() * 8))
                              : (unsigned)std::min<int64_t>(ceil_div(total, kT), (int64_t)sm_count() * 16);
    launch_chained(scale_shift_act_c8_kernel, grid, kT, 0, st)(x, y, scale, shift, total, C8, HW, act, bn);
  }
  CTL_CUDA_OK(cudaGetLastError(), "scale_shift_act launch");
  return CTL_OK;
}

// y = act(x*scale[c] + shift[c] + low[.., h/2, w/2]): the tail of a nearest-x2 residual up block whose 1x1 shortcut was
// evaluated BEFORE the up-sampling (conv1x1(up(x)) == up(conv1x1(x))).  One thread per LOW-resolution pixel: 16 B of
// `low`, a 2x2 block of x in, a 2x2 block of y out.  scale / shift may be NULL (1 / 0).
__global__ void __launch_bounds__(kT)
scale_shift_upadd_act_c8_kernel(const uint4* __restrict__ x, const uint4* __restrict__ low, uint4* __restrict__ y,
                                const float* __restrict__ scale, const float* __restrict__ shift, int64_t planes, int C8,
                                int Hl, int Wl, int act, const BnFwdSums bn) {
  pdl_entry();
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC];
  if (bn.sums != nullptr) {
    bn_fwd_prologue(bn, C8 * 8, s_scale, s_shift);
    scale = s_scale;
    shift = s_shift;
  }
  const int64_t total = planes * Hl * Wl;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pl = i / ((int64_t)Hl * Wl);
    const int pix = (int)(i - pl * Hl * Wl);
    const int yy = pix / Wl, xx = pix - yy * Wl;
    const int c0 = (int)(pl % C8) * 8;
    float sc[8], sh[8], lo[8];
    unpack8(__ldg(low + i), lo);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale ? scale[c0 + j] : 1.0f;
      sh[j] = (shift ? shift[c0 + j] : 0.0f) + lo[j];
    }
    const int64_t base = pl * 4 * Hl * Wl + (int64_t)(2 * yy) * (2 * Wl) + 2 * xx;
    const int64_t offs[4] = {base, base + 1, base + 2 * Wl, base + 2 * Wl + 1};
    uint4 in[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) in[q] = __ldcs(x + offs[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float f[8];
      unpack8(in[q], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = act_fn(f[j] * sc[j] + sh[j], act);
      y[offs[q]] = pack8(f);
    }
  }
}

// The same pass for large planes (>= 2048 low-resolution positions), structured by plane like
// scale_shift_act_planes_kernel: per-channel values in registers, one multiply-shift instead of two divisions per
// low-resolution pixel, the 2 * (1 + 4) loads of an item in flight before the first use.
__device__ __forceinline__ int div44(int n, uint64_t magic) { return (int)(((uint64_t)(uint32_t)n * magic) >> 44); }
template <int U>
__global__ void __launch_bounds__(kT, 2)
scale_shift_upadd_act_planes_kernel(const uint4* __restrict__ x, const uint4* __restrict__ low, uint4* __restrict__ y,
                                    const float* __restrict__ scale, const float* __restrict__ shift, int64_t planes, int C8,
                                    int Hl, int Wl, int items_per_plane, int act, const BnFwdSums bn, uint64_t magic_w) {
  pdl_entry();
  __shared__ float s_scale[kMaxBnC], s_shift[kMaxBnC];
  if (bn.sums != nullptr) {
    bn_fwd_prologue(bn, C8 * 8, s_scale, s_shift);
    scale = s_scale;
    shift = s_shift;
  }
  const int HWl = Hl * Wl;
  const int64_t items = planes * items_per_plane;
  float sc[8], sh[8];
  int cur_c0 = -1;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t plane = item / items_per_plane;
    const int chunk = (int)(item - plane * items_per_plane);
    const int c0 = (int)(plane % C8) * 8;
    if (c0 != cur_c0) {                                   // block-uniform
      cur_c0 = c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { sc[j] = scale ? scale[c0 + j] : 1.0f; sh[j] = shift ? shift[c0 + j] : 0.0f; }
    }
    const int p0 = chunk * (U * kT) + threadIdx.x;
    uint4 lo4[U], in[U][4];
    int64_t base[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * kT;
      const bool ok = p < HWl;
      const int yy = div44(ok ? p : 0, magic_w), xx = (ok ? p : 0) - yy * Wl;
      base[u] = plane * 4 * HWl + (int64_t)(2 * yy) * (2 * Wl) + 2 * xx;
      lo4[u] = ok ? __ldg(low + plane * HWl + p) : make_uint4(0, 0, 0, 0);
      in[u][0] = ok ? __ldcs(x + base[u]) : make_uint4(0, 0, 0, 0);
      in[u][1] = ok ? __ldcs(x + base[u] + 1) : make_uint4(0, 0, 0, 0);
      in[u][2] = ok ? __ldcs(x + base[u] + 2 * Wl) : make_uint4(0, 0, 0, 0);
      in[u][3] = ok ? __ldcs(x + base[u] + 2 * Wl + 1) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u * kT < HWl) {
        float lo[8], sl[8];
        unpack8(lo4[u], lo);
#pragma unroll
        for (int j = 0; j < 8; ++j) sl[j] = sh[j] + lo[j];
        const int64_t offs[4] = {base[u], base[u] + 1, base[u] + 2 * Wl, base[u] + 2 * Wl + 1};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
          unpack8(in[u][q], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = act_fn(f[j] * sc[j] + sl[j], act);
          y[offs[q]] = pack8(f);
        }
      }
    }
  }
}

int launch_scale_shift_upadd_act(const uint4* x, const uint4* low, uint4* y, const float* scale, const float* shift,
                                 int64_t planes, int C8, int Hl, int Wl, int act, const BnFwdSums& bn, cudaStream_t st) {
  const int64_t HWl = (int64_t)Hl * Wl, work = planes * HWl;
  if (HWl >= 2048 && HWl < ((int64_t)1 << 28) && Wl < 8192) {
    constexpr int U = 2;
    static int res = 0;
    if (res == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, scale_shift_upadd_act_planes_kernel<U>, kT, 0);
      if (res <= 0) res = 3;
    }
    const int ipp = (int)ceil_div(HWl, (int64_t)U * kT);
    const unsigned grid = (unsigned)std::min<int64_t>(planes * ipp, (int64_t)sm_count() * res);
    const uint64_t magic_w = (((uint64_t)1 << 44) + (uint64_t)Wl - 1) / (uint64_t)Wl;
    launch_chained(scale_shift_upadd_act_planes_kernel<U>, grid, kT, 0, st)(x, low, y, scale, shift, planes, C8, Hl, Wl, ipp, act,
                                                                           bn, magic_w);
  } else {
    const unsigned grid = bn.sums != nullptr
                              ? (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div(work, kT), (int64_t)sm_count() * 8))
                              : (unsigned)std::min<int64_t>(ceil_div(work, kT), (int64_t)sm_count() * 16);
    launch_chained(scale_shift_upadd_act_c8_kernel, grid, kT, 0, st)(x, low, y, scale, shift, planes, C8, Hl, Wl, act, bn);
  }
  CTL_CUDA_OK(cudaGetLastError(), "scale_shift_upadd_act launch");
  return CTL_OK;
}

// ------------------------------------------------------------------------------------------------ weight packing
// fp32 nn.Conv2d weight [Cout][Cin][k][k] -> bf16 [Cout'/NT][taps][Cin'/8][NT][8] (the K-major core-matrix order of
// conv_tc.cu).  transposed == 0: the forward weight (Cout' = Cout, Cin' = Cin).  transposed == 1: the weight of the
// INPUT-gradient convolution, w'[ci][co][r][s] = w[co][ci][k-1-r][k-1-s] (Cout' = Cin, Cin' = Cout).
// Element i of the packed weight -> (packed-view out channel o, in channel c, tap).  tap_major (1x1 filters and the
// 3x3 stride-2 kernel): [n_tile][tap][Cin'/8][NT][8].  Otherwise the vertically packed 3x3 layout of conv_tc.cu:
// [n_tile][s][Cin'/8][3*NT][8] with row n' = r*NT + n of the N dimension (tap = 3*r + s).
__device__ __forceinline__ void packed_index(int i, int ci_p, int taps, int nt, bool tap_major, int& o, int& c, int& tap) {
  int r = i;
  const int j = r & 7; r >>= 3;
  if (tap_major || taps != 9) {
    const int n = r % nt; r /= nt;
    const int q = r % (ci_p >> 3); r /= (ci_p >> 3);
    tap = r % taps;
    o = (r / taps) * nt + n;
    c = q * 8 + j;
  } else {
    const int np = r % (3 * nt); r /= (3 * nt);
    const int q = r % (ci_p >> 3); r /= (ci_p >> 3);
    const int sx = r % 3;
    tap = (np / nt) * 3 + sx;
    o = (r / 3) * nt + np % nt;
    c = q * 8 + j;
  }
}

__global__ void __launch_bounds__(kT)
pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int taps, int nt,
                        int transposed, int tap_major) {
  pdl_entry();
  const int co_p = transposed ? Cin : Cout, ci_p = transposed ? Cout : Cin;     // packed-view channel counts
  const int total = co_p * ci_p * taps;
  for (int i = blockIdx.x * kT + threadIdx.x; i < total; i += gridDim.x * kT) {
    int o, c, tap;
    packed_index(i, ci_p, taps, nt, tap_major != 0, o, c, tap);
    const float v = transposed ? w[((int64_t)c * Cin + o) * taps + (taps - 1 - tap)] : w[((int64_t)o * Cin + c) * taps + tap];
    out[i] = __float2bfloat16_rn(v);
  }
}

// every packing job of a step in ONE launch: blockIdx.y = job (table in device memory, int64 [n][8] =
// {weight ptr, out ptr, Cout, Cin, taps, nt, transposed, tap_major}), blockIdx.x strides over the job's elements
__global__ void __launch_bounds__(kT)
pack_conv_weights_batched_kernel(const int64_t* __restrict__ jobs) {
  pdl_entry();
  const int64_t* job = jobs + (int64_t)blockIdx.y * 8;
  const float* __restrict__ w = reinterpret_cast<const float*>(job[0]);
  __nv_bfloat16* __restrict__ out = reinterpret_cast<__nv_bfloat16*>(job[1]);
  const int Cout = (int)job[2], Cin = (int)job[3], taps = (int)job[4], nt = (int)job[5], transposed = (int)job[6];
  const bool tap_major = job[7] != 0;
  const int co_p = transposed ? Cin : Cout, ci_p = transposed ? Cout : Cin;
  const int total = co_p * ci_p * taps;
  for (int i = blockIdx.x * kT + threadIdx.x; i < total; i += gridDim.x * kT) {
    int o, c, tap;
    packed_index(i, ci_p, taps, nt, tap_major, o, c, tap);
    const float v = transposed ? w[((int64_t)c * Cin + o) * taps + (taps - 1 - tap)] : w[((int64_t)o * Cin + c) * taps + tap];
    out[i] = __float2bfloat16_rn(v);
  }
}

inline unsigned grid_for(int64_t total) {
  return (unsigned)std::min<int64_t>(ceil_div(total, kT), (int64_t)sm_count() * 16);
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" int ctl_pack_conv_weight(const float* weight, int64_t Cout, int64_t Cin, int taps, int transposed, int tap_major,
                                    void* out, void* stream) {
  CTL_REQUIRE(weight && out && (taps == 1 || taps == 9), CTL_ERR_INVALID, "ctl_pack_conv_weight: bad arguments");
  const int co_p = (int)(transposed ? Cin : Cout), ci_p = (int)(transposed ? Cout : Cin);
  const int nt = ctl_conv2d_n_tile(ci_p, co_p, taps);
  CTL_REQUIRE(nt > 0, CTL_ERR_UNSUPPORTED, "ctl_pack_conv_weight: no tcgen05 conv kernel for %d -> %d channels", ci_p, co_p);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t total = Cout * Cin * taps;
  launch_chained(pack_conv_weight_kernel, grid_for(total), kT, 0, (cudaStream_t)stream)(weight, (__nv_bfloat16*)out, (int)Cout, (int)Cin,
                                                                          taps, nt, transposed, tap_major);
  CTL_CUDA_OK(cudaGetLastError(), "pack_conv_weight launch");
  return CTL_OK;
}

extern "C" int ctl_pack_conv_weights_batched(const int64_t* jobs, int64_t n_jobs, int64_t max_elements, void* stream) {
  CTL_REQUIRE(jobs && n_jobs > 0 && n_jobs <= 65535 && max_elements > 0, CTL_ERR_INVALID,
              "ctl_pack_conv_weights_batched: bad arguments");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const unsigned gx = (unsigned)std::min<int64_t>(ceil_div(max_elements, kT * 4), 64);
  launch_chained(pack_conv_weights_batched_kernel, dim3(gx, (unsigned)n_jobs), kT, 0, (cudaStream_t)stream)(jobs);
  CTL_CUDA_OK(cudaGetLastError(), "pack_conv_weights_batched launch");
  return CTL_OK;
}

extern "C" int ctl_nchw_to_c8(const void* x, int x_dtype, int64_t N, int64_t C, int64_t H, int64_t W, void* y,
                              void* stream) {
  CTL_REQUIRE(x && y && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CTL_ERR_INVALID,
              "ctl_nchw_to_c8: bad arguments (C must be a multiple of 8)");
  CTL_REQUIRE(x_dtype == CTL_F32 || x_dtype == CTL_BF16, CTL_ERR_INVALID, "unknown dtype %d", x_dtype);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t HW = H * W, total = N * (C / 8) * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (x_dtype == CTL_F32)
    launch_chained(nchw_to_c8_kernel<float>, grid_for(total), kT, 0, st)((const float*)x, (__nv_bfloat16*)y, total, (int)(C / 8), HW);
  else
    launch_chained(nchw_to_c8_kernel<__nv_bfloat16>, grid_for(total), kT, 0, st)((const __nv_bfloat16*)x, (__nv_bfloat16*)y, total,
                                                                     (int)(C / 8), HW);
  CTL_CUDA_OK(cudaGetLastError(), "nchw_to_c8 launch");
  return CTL_OK;
}

extern "C" int ctl_c8_to_nchw(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, int y_dtype,
                              void* stream) {
  CTL_REQUIRE(x && y && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CTL_ERR_INVALID,
              "ctl_c8_to_nchw: bad arguments (C must be a multiple of 8)");
  CTL_REQUIRE(y_dtype == CTL_F32 || y_dtype == CTL_BF16, CTL_ERR_INVALID, "unknown dtype %d", y_dtype);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t HW = H * W, total = N * (C / 8) * HW;
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == CTL_F32)
    launch_chained(c8_to_nchw_kernel<float>, grid_for(total), kT, 0, st)((const __nv_bfloat16*)x, (float*)y, total, (int)(C / 8), HW);
  else
    launch_chained(c8_to_nchw_kernel<__nv_bfloat16>, grid_for(total), kT, 0, st)((const __nv_bfloat16*)x, (__nv_bfloat16*)y, total,
                                                                     (int)(C / 8), HW);
  CTL_CUDA_OK(cudaGetLastError(), "c8_to_nchw launch");
  return CTL_OK;
}

extern "C" int ctl_stem_conv3x3_c8(const float* x, const int64_t* labels, int in_mode, float temperature, int64_t N,
                                   int64_t Cin, int64_t H, int64_t W, const float* weight, int64_t Cout,
                                   const float* scale, const float* shift, int act, void* y, void* stream) {
  CTL_REQUIRE(y && weight && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_stem_conv3x3_c8: bad arguments");
  CTL_REQUIRE(in_mode >= 0 && in_mode <= 2 && (in_mode == 2 ? labels != nullptr : x != nullptr), CTL_ERR_INVALID,
              "ctl_stem_conv3x3_c8: in_mode %d needs %s", in_mode, in_mode == 2 ? "labels" : "x");
  CTL_REQUIRE((Cin == 1 || Cin == 4) && Cout == 16, CTL_ERR_UNSUPPORTED,
              "ctl_stem_conv3x3_c8 handles Cin in {1,4} -> Cout 16 (got %lld -> %lld)", (long long)Cin, (long long)Cout);
  CTL_REQUIRE(temperature > 0.0f, CTL_ERR_INVALID, "temperature must be positive");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = N * ceil_div(W, 32) * ceil_div(H, 8);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sm_count() * 4);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  if (Cin == 1)
    launch_chained(stem_conv_kernel<1, 16>, grid, kT, 0, st)(x, lab, weight, scale, shift, (__nv_bfloat16*)y, (int)N, (int)H, (int)W,
                                                  in_mode, 1.0f / temperature, act);
  else
    launch_chained(stem_conv_kernel<4, 16>, grid, kT, 0, st)(x, lab, weight, scale, shift, (__nv_bfloat16*)y, (int)N, (int)H, (int)W,
                                                  in_mode, 1.0f / temperature, act);
  CTL_CUDA_OK(cudaGetLastError(), "stem_conv launch");
  return CTL_OK;
}

extern "C" int ctl_head_conv1x1_c8(const void* x, int64_t N, int64_t Cin, int64_t H, int64_t W, const float* weight,
                                   const float* bias, int64_t Cout, int act, float* y, void* stream) {
  CTL_REQUIRE(x && weight && y && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_head_conv1x1_c8: bad arguments");
  CTL_REQUIRE(Cin == 16 && Cout >= 1 && Cout <= 4, CTL_ERR_UNSUPPORTED,
              "ctl_head_conv1x1_c8 handles Cin 16 -> Cout in [1,4] (got %lld -> %lld)", (long long)Cin, (long long)Cout);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  launch_chained(head_conv_kernel<16>, grid_for(N * H * W), kT, 0, (cudaStream_t)stream)((const __nv_bfloat16*)x, weight, bias, y,
                                                                           (int)N, H * W, (int)Cout, act);
  CTL_CUDA_OK(cudaGetLastError(), "head_conv launch");
  return CTL_OK;
}

extern "C" int ctl_upsample2x_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream) {
  CTL_REQUIRE(x && y && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_upsample2x_c8: bad arguments");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t planes = N * (C / 8);
  launch_chained(upsample2x_c8_kernel, grid_for(planes * H * W), kT, 0, (cudaStream_t)stream)((const uint4*)x, (uint4*)y, planes,
                                                                                (int)H, (int)W);
  CTL_CUDA_OK(cudaGetLastError(), "upsample2x launch");
  return CTL_OK;
}

extern "C" int ctl_scale_shift_upadd_act_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* scale,
                                            const float* shift, const void* low, int act, void* y, void* stream) {
  CTL_REQUIRE(x && y && low && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, CTL_ERR_INVALID,
              "ctl_scale_shift_upadd_act_c8: bad arguments (H and W must be even)");
  CTL_REQUIRE(act >= CTL_ACT_NONE && act <= CTL_ACT_SIGMOID, CTL_ERR_INVALID, "unknown activation %d", act);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t planes = N * (C / 8);
  return launch_scale_shift_upadd_act((const uint4*)x, (const uint4*)low, (uint4*)y, scale, shift, planes, (int)(C / 8),
                                      (int)(H / 2), (int)(W / 2), act, BnFwdSums{}, (cudaStream_t)stream);
}

extern "C" int ctl_bn_apply_from_sums_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const double* sums,
                                         const float* gamma, const float* beta, float eps, const void* low, int act,
                                         void* y, float* scale_out, float* shift_out, float* mean_out, float* var_out,
                                         float* running_mean, float* running_var, float momentum, void* stream) {
  CTL_REQUIRE(x && y && sums && scale_out && shift_out && N > 0 && C > 0 && C % 8 == 0 && C <= kMaxBnC && H > 0 && W > 0,
              CTL_ERR_INVALID, "ctl_bn_apply_from_sums_c8: bad arguments (C a multiple of 8, at most %d)", kMaxBnC);
  CTL_REQUIRE(!low || (H % 2 == 0 && W % 2 == 0), CTL_ERR_INVALID, "ctl_bn_apply_from_sums_c8: H and W must be even with `low`");
  CTL_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CTL_ERR_INVALID,
              "running_mean and running_var must be given together");
  CTL_REQUIRE(act >= CTL_ACT_NONE && act <= CTL_ACT_SIGMOID, CTL_ERR_INVALID, "unknown activation %d", act);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const double count = (double)(N * H * W);
  const BnFwdSums bn = {sums, 1.0 / count, count > 1.0 ? count / (count - 1.0) : 1.0, gamma, beta, scale_out, shift_out,
                        mean_out, var_out, running_mean, running_var, eps, momentum};
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t planes = N * (C / 8);
  // every CTA pays the per-channel prologue: a few CTAs per SM striding over the tensor amortise it
  if (low) {
    return launch_scale_shift_upadd_act((const uint4*)x, (const uint4*)low, (uint4*)y, nullptr, nullptr, planes, (int)(C / 8),
                                        (int)(H / 2), (int)(W / 2), act, bn, st);
  } else {
    return launch_scale_shift_act((const uint4*)x, (uint4*)y, nullptr, nullptr, planes, (int)(C / 8), H * W, act, bn, st);
  }
}

extern "C" int ctl_scale_shift_act_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* scale,
                                      const float* shift, int act, void* y, void* stream) {
  CTL_REQUIRE(x && y && scale && shift && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CTL_ERR_INVALID,
              "ctl_scale_shift_act_c8: bad arguments");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  return launch_scale_shift_act((const uint4*)x, (uint4*)y, scale, shift, N * (C / 8), (int)(C / 8), H * W, act, BnFwdSums{},
                                (cudaStream_t)stream);
}

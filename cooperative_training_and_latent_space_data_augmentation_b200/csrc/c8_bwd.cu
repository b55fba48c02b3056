// Backward-pass CUDA-core kernels for the blocked activation layout C8 = bf16 [N][C/8][H][W][8] (HBM-bound streaming
// passes around the tensor-core dgrad / wgrad kernels):
//   train-mode BatchNorm + LeakyReLU/ReLU backward (two-stage reduce + apply), per-channel sums (bias gradients),
//   nearest x2 up-sampling backward (2x2 sum), zero-stuffing / parity split (stride-2 conv and ConvTranspose2d k2 s2
//   backward), the 1x1 head backward and the stem (Cin 1/4) weight / input gradients with the STN-input softmax fused.
//
// Reference code whose autograd these replace (the reference relies on torch autograd for all of them):
//   nn.BatchNorm2d + LeakyReLU(0.2)/ReLU   medseg/models/ebm/encoder_decoder.py:34,43-49,322-328,370-378,394-397,468-476
//   nn.UpsamplingNearest2d                  medseg/models/ebm/encoder_decoder.py:294-296
//   ConvTranspose2d(k2,s2) / Conv2d s2      medseg/models/ebm/encoder_decoder.py:302, :40-41
//   MyDecoder.final_conv (+Sigmoid)         medseg/models/ebm/encoder_decoder.py:439-452
//   MyEncoder.inc[0] + construct_input      medseg/models/ebm/encoder_decoder.py:370-371, medseg/common_utils/basic_operations.py:110-158
#include <algorithm>
#include <cstdlib>
#include <cmath>

#include "ctl_common.cuh"

namespace ctl {
namespace {

constexpr int kT = 256;

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
// derivative of the activation expressed through its OUTPUT h (sign(h) == sign(pre-activation) for LReLU/ReLU)
__device__ __forceinline__ float act_slope(float h, int act) {
  switch (act) {
    case CTL_ACT_LRELU: return h > 0.0f ? 1.0f : 0.2f;
    case CTL_ACT_RELU: return h > 0.0f ? 1.0f : 0.0f;
    case CTL_ACT_SIGMOID: return h * (1.0f - h);
    default: return 1.0f;
  }
}

// plane split: enough CTAs to fill the GPU even when N*C/8 is small (16 channels at 224x224: 128 planes of 800 KB), and
// a CTA count that fills WHOLE waves: with `resident` CTA slots on the device, planes*splits = 2.02 waves runs as three
// (measured: 896 CTAs on 444 slots cost +50 %).  Smallest split count in [1, 64] that gives at least one full wave and
// wastes < 6 % of the last one; else the most efficient candidate.
inline int plane_splits(int64_t planes, int64_t HW, int ctas_per_sm) {
  const int64_t resident = (int64_t)sm_count() * std::max(1, ctas_per_sm);
  const int64_t max_s = std::max<int64_t>(1, std::min<int64_t>(64, HW / (kT * 4)));
  int best = 1;
  double best_eff = 0.0;
  for (int64_t s = 1; s <= max_s; ++s) {
    const double waves = (double)(planes * s) / (double)resident;
    const double eff = waves / std::ceil(waves);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = (int)s; }
    if (waves >= 1.0 && eff >= 0.94) return (int)s;
  }
  return best;
}

// ------------------------------------------------------------------------------------------------ reductions
// Stage 1 of every per-channel reduction.  MODE 0: sum x, sum x^2 (forward statistics / bias gradients).
// MODE 1: dv = dy * act'(h); sum dv, sum dv*a (BatchNorm backward); optionally materialises dv.
// partial: double [(split*planes + plane)*8 + j][2]
// WITH_H = false (h is never read: MODE 0, or MODE 1 with the activation input recomputed from a): the loop is
// double-buffered -- the next four positions' loads are in flight while the current four are reduced.
template <int MODE, bool WITH_H = (MODE == 1)>
__global__ void __launch_bounds__(kT, 2)
plane_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ h, const uint4* __restrict__ a,
                    uint4* __restrict__ dv_out, double* __restrict__ partial, int64_t planes, int64_t HW, int splits,
                    int act, const float* __restrict__ act_scale, const float* __restrict__ act_shift, int C8,
                    double* __restrict__ totals) {
  pdl_entry();
  const int64_t plane = blockIdx.x;
  // h == NULL with an affine: the activation input is recomputed, h = act(a*act_scale[c] + act_shift[c]) has its sign
  float asc[8], ash[8];
  if (MODE == 1 && h == nullptr && act_scale != nullptr) {
    const int c0 = (int)(plane % C8) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) { asc[j] = __ldg(act_scale + c0 + j); ash[j] = __ldg(act_shift + c0 + j); }
  }
  const int split = blockIdx.y;
  const int64_t chunk = (HW + splits - 1) / splits;
  const int64_t lo = split * chunk, hi = min(HW, lo + chunk);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s[j] = 0.0f; q[j] = 0.0f; }
  const bool from_a = MODE == 1 && h == nullptr && act_scale != nullptr;
  auto accumulate = [&](const int64_t idx, const uint4 xr, const uint4 ar, const uint4 hr) {
    float f[8];
    unpack8(xr, f);
    if (MODE == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], f[j], q[j]); }
    } else {
      float av[8];
      unpack8(ar, av);
      if (h != nullptr) {
        float hv[8];
        unpack8(hr, hv);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= act_slope(hv[j], act);
        if (dv_out != nullptr) {
          const uint4 packed = pack8(f);
          dv_out[idx] = packed;
          unpack8(packed, f);                     // reduce what the consumers will read (bf16-rounded)
        }
      } else if (from_a) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= act_slope(av[j] * asc[j] + ash[j], act);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += f[j]; q[j] = fmaf(f[j], av[j], q[j]); }
    }
  };
  // four positions per iteration, every load issued before the first use: 8-12 independent 16-byte loads per thread
  // (~100 KB in flight per SM with three resident CTAs) -- two positions left the HBM pipe at 3.0 TB/s
  const uint4 zero = make_uint4(0, 0, 0, 0);
  constexpr int kP = 4;
  if constexpr (!WITH_H) {
    auto load = [&](const int64_t i, uint4 (&xv)[kP], uint4 (&av)[MODE == 1 ? kP : 1]) {
#pragma unroll
      for (int u = 0; u < kP; ++u) {
        const int64_t idx = plane * HW + i + u * kT;
        const bool ok = i + u * kT < hi;
        xv[u] = ok ? __ldcs(x + idx) : zero;
        if (MODE == 1) av[MODE == 1 ? u : 0] = ok ? __ldg(a + idx) : zero;
      }
    };
    auto consume = [&](const int64_t i, const uint4 (&xv)[kP], const uint4 (&av)[MODE == 1 ? kP : 1]) {
#pragma unroll
      for (int u = 0; u < kP; ++u)
        if (i + u * kT < hi) accumulate(plane * HW + i + u * kT, xv[u], av[MODE == 1 ? u : 0], zero);
    };
    uint4 xa[kP], xb[kP], aa[MODE == 1 ? kP : 1], ab[MODE == 1 ? kP : 1];
    int64_t i = lo + threadIdx.x;
    if (i < hi) load(i, xa, aa);
    for (; i < hi; i += 2 * kP * kT) {
      const int64_t i2 = i + kP * kT, i3 = i + 2 * kP * kT;
      if (i2 < hi) load(i2, xb, ab);
      consume(i, xa, aa);
      if (i3 < hi) load(i3, xa, aa);
      if (i2 < hi) consume(i2, xb, ab);
    }
  } else
  for (int64_t i = lo + threadIdx.x; i < hi; i += kP * kT) {
    uint4 xv[kP], av[kP], hv[kP];
#pragma unroll
    for (int u = 0; u < kP; ++u) {
      const int64_t idx = plane * HW + i + u * kT;
      const bool ok = i + u * kT < hi;
      xv[u] = ok ? __ldcs(x + idx) : zero;
      av[u] = (MODE == 1 && ok) ? __ldg(a + idx) : zero;
      hv[u] = (MODE == 1 && h != nullptr && ok) ? __ldg(h + idx) : zero;
    }
#pragma unroll
    for (int u = 0; u < kP; ++u)
      if (i + u * kT < hi) accumulate(plane * HW + i + u * kT, xv[u], av[u], hv[u]);
  }
  // warp level in fp32 (a thread holds ~20 values, a warp ~640: a few ulp), fp64 from there on
  __shared__ double red[kT / 32][16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[warp][j] = (double)s[j]; red[warp][8 + j] = (double)q[j]; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0.0;
    for (int w = 0; w < kT / 32; ++w) t += red[w][threadIdx.x];
    const int j = threadIdx.x & 7, second = threadIdx.x >> 3;
    if (totals != nullptr)      // per-channel totals [2][C] (zeroed by the caller): ONE fp64 atomic per value and CTA, no
      atomicAdd(totals + second * (C8 * 8) + (int)(plane % C8) * 8 + j, t);      // finalisation launch afterwards
    else
      partial[(((int64_t)split * planes + plane) * 8 + j) * 2 + second] = t;
  }
}

// The BatchNorm-backward reduction for large planes (HW >= 2048, activation input recomputed from a), persistent and
// structured like bn_bwd_apply_planes_kernel: a work item is U*kT consecutive positions of one 8-channel plane, items
// are visited CHANNEL-GROUP-major (all planes of channels 0-7 first, ...), so a CTA's running sums stay valid across
// items and are flushed -- warp shuffles, shared memory, ONE fp64 atomic per statistic -- only when its channel group
// changes (at most C/8 times) instead of once per short-lived CTA; the activation affine sits in registers.
// WITH_H: the activation slope comes from the layer output h (residual tails: h = act(BN(a) + shortcut) cannot be
// recomputed from a), and dv may be materialised (bf16; the sums are then taken over the rounded values the consumers read).
template <int U, bool WITH_H>
__global__ void __launch_bounds__(kT, 3)
bn_bwd_reduce_items_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ h, const uint4* __restrict__ a,
                           uint4* __restrict__ dv_out, int N, int C8, int HW, int items_per_plane, int act,
                           const float* __restrict__ act_scale, const float* __restrict__ act_shift,
                           double* __restrict__ totals) {
  pdl_entry();
  __shared__ double red[kT / 32][16];
  const bool with_act = (WITH_H || act_scale != nullptr) && act != CTL_ACT_NONE;
  const float neg = !with_act ? 1.0f : (act == CTL_ACT_LRELU ? 0.2f : 0.0f);
  const int64_t per_group = (int64_t)N * items_per_plane, items = per_group * C8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s[8], q[8], asc[8], ash[8];
  int cur = -1;
  auto flush = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        q[j] += __shfl_xor_sync(0xffffffffu, q[j], o);
      }
    }
    __syncthreads();                                      // the previous flush's readers are done with `red`
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { red[warp][j] = (double)s[j]; red[warp][8 + j] = (double)q[j]; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
      double t = 0.0;
      for (int w = 0; w < kT / 32; ++w) t += red[w][threadIdx.x];
      const int j = threadIdx.x & 7, second = threadIdx.x >> 3;
      atomicAdd(totals + second * (C8 * 8) + cur * 8 + j, t);
    }
  };
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int c8 = (int)(item / per_group);
    const int64_t r = item - (int64_t)c8 * per_group;
    const int n = (int)(r / items_per_plane), chunk = (int)(r - (int64_t)n * items_per_plane);
    if (c8 != cur) {                                      // block-uniform
      if (cur >= 0) flush();
      cur = c8;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] = 0.0f; q[j] = 0.0f;
        asc[j] = (!WITH_H && with_act) ? __ldg(act_scale + c8 * 8 + j) : 1.0f;
        ash[j] = (!WITH_H && with_act) ? __ldg(act_shift + c8 * 8 + j) : 0.0f;
      }
    }
    const int p0 = chunk * (U * kT) + threadIdx.x;
    const int64_t base = ((int64_t)n * C8 + c8) * HW;
    uint4 xv[U], av4[U], hv4[WITH_H ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = p0 + u * kT < HW;
      xv[u] = ok ? __ldcs(dy + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
      av4[u] = ok ? __ldg(a + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
      if (WITH_H) hv4[WITH_H ? u : 0] = ok ? __ldg(h + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {                          // positions past the plane were loaded as zeros: dv = 0
      float f[8], av[8];
      unpack8(xv[u], f);
      unpack8(av4[u], av);
      if constexpr (WITH_H) {
        float hv[8];
        unpack8(hv4[WITH_H ? u : 0], hv);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= hv[j] > 0.0f ? 1.0f : neg;
        if (dv_out != nullptr) {
          const uint4 packed = pack8(f);
          if (p0 + u * kT < HW) dv_out[base + p0 + u * kT] = packed;
          unpack8(packed, f);                               // reduce what the consumers will read (bf16-rounded)
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= fmaf(av[j], asc[j], ash[j]) > 0.0f ? 1.0f : neg;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s[j] += f[j];
        q[j] = fmaf(f[j], av[j], q[j]);
      }
    }
  }
  if (cur >= 0) flush();
}

// resident CTAs of the reduction per SM (register-limited; queried once per variant)
template <int MODE, bool WITH_H = (MODE == 1)>
int reduce_ctas_per_sm() {
  static int cached = 0;
  if (cached <= 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, plane_reduce_kernel<MODE, WITH_H>, kT, 0) != cudaSuccess || n <= 0) n = 3;
    cached = n;
  }
  return cached;
}

// one warp per channel: sums the partials over (split, n); lane 0 holds the totals
__device__ __forceinline__ void channel_totals(const double* __restrict__ partial, int64_t planes, int splits, int N, int C,
                                               int c, double& S1, double& S2) {
  const int C8 = C >> 3, c8 = c >> 3, j = c & 7;
  const int lane = threadIdx.x & 31;
  double a = 0.0, b = 0.0;
  const int entries = splits * N;
  for (int e = lane; e < entries; e += 32) {
    const int split = e / N, n = e - split * N;
    const int64_t plane = (int64_t)n * C8 + c8;
    const double* p = partial + (((int64_t)split * planes + plane) * 8 + j) * 2;
    a += p[0]; b += p[1];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
  S1 = a; S2 = b;
}

__global__ void channel_sum_finalize_kernel(const double* __restrict__ partial, int64_t planes, int splits, int N, int C,
                                            float* __restrict__ sum_out, float* __restrict__ sumsq_out) {
  pdl_entry();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double S1, S2;
  channel_totals(partial, planes, splits, N, C, c, S1, S2);
  if ((threadIdx.x & 31) == 0) {
    if (sum_out) sum_out[c] = (float)S1;
    if (sumsq_out) sumsq_out[c] = (float)S2;
  }
}


// forward BatchNorm: batch mean / biased variance -> fused affine y = x*scale + shift (scale = gamma*rsqrt(var+eps),
// shift = beta - mean*scale); optional running-stat update (momentum, unbiased variance) like nn.BatchNorm2d.
__global__ void bn_fwd_finalize_kernel(const double* __restrict__ partial, int64_t planes, int splits, int N, int C,
                                       double count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                       float eps, float* __restrict__ scale, float* __restrict__ shift,
                                       float* __restrict__ mean_out, float* __restrict__ var_out, float* running_mean,
                                       float* running_var, float momentum) {
  pdl_entry();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  channel_totals(partial, planes, splits, N, C, c, s, q);
  if ((threadIdx.x & 31) != 0) return;
  const double mean = s / count;
  double var = q / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float inv = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  scale[c] = g * inv;
  shift[c] = b - (float)mean * g * inv;
  if (mean_out) mean_out[c] = (float)mean;
  if (var_out) var_out[c] = (float)var;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// same finalisation from per-channel sums accumulated by the conv epilogue: sums = double [2][C] (sum x | sum x^2)
__global__ void bn_fwd_from_sums_kernel(const double* __restrict__ sums, int C, double count, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, float eps, float* __restrict__ scale,
                                        float* __restrict__ shift, float* __restrict__ mean_out, float* __restrict__ var_out,
                                        float* running_mean, float* running_var, float momentum) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float inv = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
  scale[c] = g * inv;
  shift[c] = b - (float)mean * g * inv;
  if (mean_out) mean_out[c] = (float)mean;
  if (var_out) var_out[c] = (float)var;
  if (running_mean) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// BatchNorm backward coefficients: da = c1*dv + c2*a + c3 with
//   c1 = scale, c2 = -scale*inv*dgamma/M, c3 = scale*(inv*mean*dgamma - dbeta)/M,  scale = gamma*inv, inv = rsqrt(var+eps)
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ partial, int64_t planes, int splits, int N, int C,
                                       double count, const float* __restrict__ mean, const float* __restrict__ var,
                                       float eps, const float* __restrict__ gamma, float* __restrict__ coef,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_entry();
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= C) return;
  double S1, S2;
  channel_totals(partial, planes, splits, N, C, c, S1, S2);
  if ((threadIdx.x & 31) != 0) return;
  const double mu = (double)mean[c];
  const double inv = 1.0 / sqrt((double)var[c] + (double)eps);
  const double g = gamma ? (double)gamma[c] : 1.0;
  const double scale = g * inv;
  const double dg = inv * (S2 - mu * S1);
  coef[c] = (float)scale;
  coef[C + c] = (float)(-scale * inv * dg / count);
  coef[2 * C + c] = (float)(scale * (inv * mu * dg - S1) / count);
  if (dgamma) dgamma[c] = (float)dg;
  if (dbeta) dbeta[c] = (float)S1;
}

// da = c1*dv + c2*a + c3, dv = dy * act'(h) when h is given
struct BnBwdTotals {          // coefficients computed in the kernel's prologue from the reduction's per-channel totals
  const double* totals;       // [2][C]: sum dv | sum dv*a; nullptr -> `coef` is read from global memory
  const float* mean;
  const float* var;
  const float* gamma;
  float* dgamma;              // written by CTA 0 (may be nullptr)
  float* dbeta;
  double count;
  float eps;
};
constexpr int kMaxBnC = 256;

// BatchNorm-backward coefficients from the per-channel totals (same arithmetic as bn_bwd_finalize_kernel), once per CTA
// (C <= 256 values): no finalisation launch.  s_coef: c1 | c2 | c3, [3][C]
__device__ __forceinline__ void bn_bwd_coef_prologue(const BnBwdTotals& bt, int C, float* s_coef) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
      // fp64 only where it matters (the cancellation in S2 - mean*S1); 1/sqrt and the divisions in fp32 -- this
      // runs in every CTA's prologue, and fp64 division / square root are ~100-instruction sequences
      const double S1 = bt.totals[c], S2 = bt.totals[C + c];
      const float mu = bt.mean[c];
      const float inv = 1.0f / sqrtf(bt.var[c] + bt.eps);
      const float g = bt.gamma ? bt.gamma[c] : 1.0f;
      const float scale = g * inv;
      const float dg = inv * (float)(S2 - (double)mu * S1);
      const float rc = (float)(1.0 / bt.count), s1 = (float)S1;
      s_coef[c] = scale;
      s_coef[C + c] = -scale * inv * dg * rc;
      s_coef[2 * C + c] = scale * (inv * mu * dg - s1) * rc;
      if (blockIdx.x == 0) {
        if (bt.dgamma) bt.dgamma[c] = dg;
        if (bt.dbeta) bt.dbeta[c] = s1;
      }
    }
  __syncthreads();
}

__global__ void __launch_bounds__(kT)
bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ h, const uint4* __restrict__ a,
                    const float* __restrict__ coef_in, uint4* __restrict__ da, int64_t total, int C8, int64_t HW, int act,
                    const float* __restrict__ act_scale, const float* __restrict__ act_shift, const BnBwdTotals bt) {
  pdl_entry();
  const int C = C8 * 8;
  __shared__ float s_coef[3 * kMaxBnC];
  const float* coef = coef_in;
  if (bt.totals != nullptr) {
    bn_bwd_coef_prologue(bt, C, s_coef);
    coef = s_coef;
  }
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int c0 = (int)((i / HW) % C8) * 8;
    float f[8], av[8];
    unpack8(__ldcs(dy + i), f);
    unpack8(__ldcs(a + i), av);
    if (h != nullptr) {
      float hv[8];
      unpack8(__ldcs(h + i), hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= act_slope(hv[j], act);
    } else if (act_scale != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] *= act_slope(av[j] * __ldg(act_scale + c0 + j) + __ldg(act_shift + c0 + j), act);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      f[j] = fmaf(coef[c0 + j], f[j], fmaf(coef[C + c0 + j], av[j], coef[2 * C + c0 + j]));
    da[i] = pack8(f);
  }
}

// The same pass for large planes (HW >= 2048), structured by plane: a work item is U*kT consecutive positions of ONE
// 8-channel plane, so the 24 coefficients (+16 activation-affine values) of its channels sit in registers -- the
// generic form pays a 64-bit division and ~40 shared / constant-cache loads per 16-byte position, enough to hold it at
// 0.76 of the HBM peak -- and all 2*U loads of an item are issued before the first use.
template <int U>
__global__ void __launch_bounds__(kT, 3)
bn_bwd_apply_planes_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ a, const float* __restrict__ coef_in,
                           uint4* __restrict__ da, int64_t planes, int C8, int HW, int items_per_plane, int act,
                           const float* __restrict__ act_scale, const float* __restrict__ act_shift, const BnBwdTotals bt) {
  pdl_entry();
  const int C = C8 * 8;
  __shared__ float s_coef[5 * kMaxBnC];                  // c1 | c2 | c3 | act scale | act shift
  if (bt.totals != nullptr) {
    bn_bwd_coef_prologue(bt, C, s_coef);
  } else {
    for (int c = threadIdx.x; c < 3 * C; c += kT) s_coef[c] = coef_in[c];
  }
  const bool with_act = act_scale != nullptr && act != CTL_ACT_NONE;
  for (int c = threadIdx.x; c < C; c += kT) {
    s_coef[3 * C + c] = with_act ? act_scale[c] : 1.0f;
    s_coef[4 * C + c] = with_act ? act_shift[c] : 0.0f;
  }
  __syncthreads();
  const float neg = !with_act ? 1.0f : (act == CTL_ACT_LRELU ? 0.2f : 0.0f);   // act'(p) for p <= 0 (LReLU | ReLU)
  const int64_t items = planes * items_per_plane;
  float c1[8], c2[8], c3[8], asc[8], ash[8];
  int cur_c0 = -1;
  for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
    const int64_t plane = item / items_per_plane;
    const int chunk = (int)(item - plane * items_per_plane);
    const int c0 = (int)(plane % C8) * 8;
    if (c0 != cur_c0) {                                   // block-uniform
      cur_c0 = c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        c1[j] = s_coef[c0 + j]; c2[j] = s_coef[C + c0 + j]; c3[j] = s_coef[2 * C + c0 + j];
        asc[j] = s_coef[3 * C + c0 + j]; ash[j] = s_coef[4 * C + c0 + j];
      }
    }
    const int p0 = chunk * (U * kT) + threadIdx.x;
    const int64_t base = plane * HW;
    uint4 xv[U], av4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool ok = p0 + u * kT < HW;
      xv[u] = ok ? __ldcs(dy + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
      av4[u] = ok ? __ldcs(a + base + p0 + u * kT) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (p0 + u * kT < HW) {
        float f[8], av[8];
        unpack8(xv[u], f);
        unpack8(av4[u], av);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dv = f[j] * (fmaf(av[j], asc[j], ash[j]) > 0.0f ? 1.0f : neg);
          f[j] = fmaf(c1[j], dv, fmaf(c2[j], av[j], c3[j]));
        }
        da[base + p0 + u * kT] = pack8(f);
      }
    }
  }
}

// launches the plane-structured form where it applies (no h tensor, planes of >= 2048 positions), else the generic one
int launch_bn_bwd_apply(const uint4* dy, const uint4* h, const uint4* a, const float* coef, uint4* da, int64_t N, int C8,
                        int64_t HW, int act, const float* act_scale, const float* act_shift, const BnBwdTotals& bt,
                        cudaStream_t st) {
  const int64_t planes = N * C8, total = planes * HW;
  static const bool planes_on = [] { const char* e = getenv("CTL_BN_APPLY_PLANES"); return !(e && e[0] == '0'); }();
  if (planes_on && h == nullptr && HW >= 2048 && HW < ((int64_t)1 << 30) && (act == CTL_ACT_NONE || act == CTL_ACT_LRELU || act == CTL_ACT_RELU)) {
    const int U = HW >= 8192 ? 4 : 2;
    const int ipp = (int)ceil_div(HW, (int64_t)U * kT);
    static int resident[2] = {0, 0};                      // CTAs per SM of the U = 2 / U = 4 forms
    int& res = resident[U == 4];
    if (res == 0) {
      if (U == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, bn_bwd_apply_planes_kernel<4>, kT, 0);
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, bn_bwd_apply_planes_kernel<2>, kT, 0);
      if (res <= 0) res = 3;
    }
    const unsigned grid = (unsigned)std::min<int64_t>(planes * ipp, (int64_t)sm_count() * res);
    if (U == 4)
      launch_chained(bn_bwd_apply_planes_kernel<4>, grid, kT, 0, st)(dy, a, coef, da, planes, C8, (int)HW, ipp, act, act_scale,
                                                                     act_shift, bt);
    else
      launch_chained(bn_bwd_apply_planes_kernel<2>, grid, kT, 0, st)(dy, a, coef, da, planes, C8, (int)HW, ipp, act, act_scale,
                                                                     act_shift, bt);
  } else {
    // every CTA pays the coefficient prologue when the totals are given: a few CTAs per SM striding over the tensor
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, kT), (int64_t)sm_count() * (bt.totals != nullptr ? 6 : 16));
    launch_chained(bn_bwd_apply_kernel, grid, kT, 0, st)(dy, h, a, coef, da, total, C8, HW, act, act_scale, act_shift, bt);
  }
  CTL_CUDA_OK(cudaGetLastError(), "bn_bwd apply launch");
  return CTL_OK;
}


// dv = dy * act'(h) only (activation backward without a BatchNorm in front)
__global__ void __launch_bounds__(kT)
act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ h, uint4* __restrict__ dv, int64_t total, int act) {
  pdl_entry();
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    float f[8], hv[8];
    unpack8(__ldcs(dy + i), f);
    unpack8(__ldcs(h + i), hv);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= act_slope(hv[j], act);
    dv[i] = pack8(f);
  }
}

// ------------------------------------------------------------------------------------------------ resampling
// backward of nearest x2 up-sampling: dx[p] = sum of the 2x2 block of dy; one thread per OUTPUT (low-res) pixel
__global__ void __launch_bounds__(kT)
downsample2x_sum_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int64_t planes, int H, int W) {
  pdl_entry();
  const int64_t total = planes * H * W;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pl = i / ((int64_t)H * W);
    const int pix = (int)(i - pl * H * W);
    const int yy = pix / W, xx = pix - yy * W;
    const uint4* p = dy + pl * 4 * H * W + (int64_t)(2 * yy) * (2 * W) + 2 * xx;
    float acc[8], f[8];
    unpack8(__ldcs(p), acc);
    unpack8(__ldcs(p + 1), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
    unpack8(__ldcs(p + 2 * W), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
    unpack8(__ldcs(p + 2 * W + 1), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
    dx[i] = pack8(acc);
  }
}
// out[2y][2x] = x[y][x], the three other pixels of each 2x2 block = 0 (dy of a stride-2 conv seen at full resolution)
__global__ void __launch_bounds__(kT)
zero_stuff2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t planes, int H, int W) {
  pdl_entry();
  const int64_t total = planes * H * W;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pl = i / ((int64_t)H * W);
    const int pix = (int)(i - pl * H * W);
    const int yy = pix / W, xx = pix - yy * W;
    uint4* o = y + pl * 4 * H * W + (int64_t)(2 * yy) * (2 * W) + 2 * xx;
    o[0] = __ldcs(x + i); o[1] = z; o[2 * W] = z; o[2 * W + 1] = z;
  }
}
// out[d][plane][y][x] = x[plane][2y + d/2][2x + d%2], d = 0..3 (dy of a ConvTranspose2d k2 s2 split by kernel tap)
__global__ void __launch_bounds__(kT)
split_parity2x2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t planes, int H, int W) {
  pdl_entry();
  const int64_t total = planes * H * W;                // H, W: LOW resolution
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t pl = i / ((int64_t)H * W);
    const int pix = (int)(i - pl * H * W);
    const int yy = pix / W, xx = pix - yy * W;
    const uint4* p = x + pl * 4 * H * W + (int64_t)(2 * yy) * (2 * W) + 2 * xx;
    y[i] = __ldcs(p);
    y[total + i] = __ldcs(p + 1);
    y[2 * total + i] = __ldcs(p + 2 * W);
    y[3 * total + i] = __ldcs(p + 2 * W + 1);
  }
}

// ------------------------------------------------------------------------------------------------ head backward
// y = act(W x + b), x: 16 blocked channels, y: COUT planar fp32.  dz = dy * act'(y);
// dx = W^T dz (C8 bf16); dW[co][ci] += sum_p dz[co] x[ci]; db[co] += sum_p dz[co]  (atomic accumulation into fp32).
template <int COUT>
__global__ void __launch_bounds__(kT, 2)
head_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ yout, const __nv_bfloat16* __restrict__ x,
                const float* __restrict__ w, __nv_bfloat16* __restrict__ dx, float* __restrict__ dW,
                float* __restrict__ db, int N, int64_t HW, int act) {
  pdl_entry();
  constexpr int CIN = 16;
  __shared__ float sw[COUT * CIN];
  __shared__ float red[kT / 32][COUT * CIN + COUT];
  for (int i = threadIdx.x; i < COUT * CIN; i += kT) sw[i] = w[i];
  __syncthreads();
  float aw[COUT * CIN], ab[COUT];
#pragma unroll
  for (int i = 0; i < COUT * CIN; ++i) aw[i] = 0.0f;
#pragma unroll
  for (int i = 0; i < COUT; ++i) ab[i] = 0.0f;
  const int64_t total = (int64_t)N * HW;
  // software-pipelined: the NEXT pixel's loads (2 x 16 B of x, COUT x 4 B of dy [and y]) are issued before the current
  // pixel's arithmetic -- with ~64 weight-gradient accumulators per thread only ~512 threads fit an SM, and one pixel's
  // 48 bytes per thread in flight kept the kernel at 2.7 TB/s
  struct Px { uint4 x0, x1; float dy[COUT], y[COUT]; };
  auto fetch = [&](int64_t i, Px& q) {
    const int64_t n = i / HW, pix = i - n * HW;
    q.x0 = __ldcs(reinterpret_cast<const uint4*>(x + ((n * (CIN / 8) + 0) * HW + pix) * 8));
    q.x1 = __ldcs(reinterpret_cast<const uint4*>(x + ((n * (CIN / 8) + 1) * HW + pix) * 8));
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      const int64_t o = (n * COUT + co) * HW + pix;
      q.dy[co] = __ldcs(dy + o);
      q.y[co] = act != CTL_ACT_NONE ? __ldcs(yout + o) : 0.0f;
    }
  };
  const int64_t stride = (int64_t)gridDim.x * kT;
  int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x;
  Px cur, nxt;
  if (i < total) fetch(i, cur);
  for (; i < total; i += stride) {
    const bool more = i + stride < total;
    if (more) fetch(i + stride, nxt);
    const int64_t n = i / HW, pix = i - n * HW;
    float xv[CIN];
    unpack8(cur.x0, *reinterpret_cast<float(*)[8]>(&xv[0]));
    unpack8(cur.x1, *reinterpret_cast<float(*)[8]>(&xv[8]));
    float dz[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float g = cur.dy[co];
      if (act != CTL_ACT_NONE) g *= act_slope(cur.y[co], act);
      dz[co] = g;
      ab[co] += g;
    }
    float dxv[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float t = 0.0f;
#pragma unroll
      for (int co = 0; co < COUT; ++co) {
        t = fmaf(sw[co * CIN + ci], dz[co], t);
        aw[co * CIN + ci] = fmaf(dz[co], xv[ci], aw[co * CIN + ci]);
      }
      dxv[ci] = t;
    }
#pragma unroll
    for (int c8 = 0; c8 < CIN / 8; ++c8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = dxv[c8 * 8 + j];
      *reinterpret_cast<uint4*>(dx + ((n * (CIN / 8) + c8) * HW + pix) * 8) = pack8(f);
    }
    if (more) cur = nxt;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < COUT * CIN; ++i) {
    float v = aw[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][i] = v;
  }
#pragma unroll
  for (int i = 0; i < COUT; ++i) {
    float v = ab[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][COUT * CIN + i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * CIN + COUT; i += kT) {
    float t = 0.0f;
    for (int wq = 0; wq < kT / 32; ++wq) t += red[wq][i];
    if (i < COUT * CIN) atomicAdd(dW + i, t);
    else atomicAdd(db + (i - COUT * CIN), t);
  }
}

// ------------------------------------------------------------------------------------------------ stem backward
// value of the stem's input channel ci at (n, iy, ix) (zero outside the image), honouring in_mode
template <int CIN>
__device__ __forceinline__ void stem_input_at(const float* __restrict__ x, const long long* __restrict__ labels, int in_mode,
                                              float inv_temp, int64_t n, int iy, int ix, int H, int W, float (&v)[CIN]) {
  if (iy < 0 || iy >= H || ix < 0 || ix >= W) {
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) v[ci] = 0.0f;
    return;
  }
  const int64_t HW = (int64_t)H * W, q = (int64_t)iy * W + ix;
  if (in_mode == 2) {
    const long long lab = labels[n * HW + q];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) v[ci] = (lab == ci) ? 1.0f : 0.0f;
    return;
  }
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

// The stem's input as a 16-channel C8 tensor so that the stem convolution and its weight gradient run on the tensor-core
// kernels (K3 / K3w with Cin = 16) instead of the CUDA-core stem kernels: in_mode 0 the image, 1 softmax(x / T), 2 the
// one-hot of the label map (construct_input, basic_operations.py:110-158).  The fp32 value v of input channel c is kept
// to ~16 mantissa bits as a bf16 pair: channel c = hi = bf16(v), channel CIN + c = lo = bf16(v - hi), channel 2*CIN + c
// = hi again (it meets the low half of the split WEIGHT, ops.pad_stem_weight); the rest is zero.  One thread per
// pixel: CIN planar reads, 2 x 16 bytes out.
template <int CIN>
__global__ void __launch_bounds__(kT)
stem_input_c8_kernel(const float* __restrict__ x, const long long* __restrict__ labels, int in_mode, float inv_temp, int N,
                     int H, int W, uint4* __restrict__ out) {
  pdl_entry();
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW, q = i - n * HW;
    const int iy = (int)(q / W), ix = (int)(q - (int64_t)iy * W);
    float v[CIN];
    stem_input_at<CIN>(x, labels, in_mode, inv_temp, n, iy, ix, H, W, v);
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = 0.0f;
#pragma unroll
    for (int c = 0; c < CIN; ++c) {
      const float hi = __bfloat162float(__float2bfloat16_rn(v[c]));
      f[c] = hi;
      f[CIN + c] = v[c] - hi;
      f[2 * CIN + c] = hi;
    }
    out[(n * 2 + 0) * HW + q] = pack8(*reinterpret_cast<const float(*)[8]>(&f[0]));
    out[(n * 2 + 1) * HW + q] = pack8(*reinterpret_cast<const float(*)[8]>(&f[8]));
  }
}

// Weight gradient of the stem: dW[co][ci][r][s] += sum_p dy[p][co] * in[p + (r-1, s-1)][ci].
// CTA = one 32x8 pixel tile per iteration (persistent), input tile with halo and dy tile staged in shared memory.
// Thread (g = tid/16, t = tid%16): pixels g, g+16, ... of the tile; co quad t%4; k slice t/4 (CIN 4: channel t/4, nine
// taps; CIN 1: filter row t/4 < 3, three taps) -> 4 x KT register accumulators, 1 + KT shared loads per 4*KT FMAs.
template <int CIN>
__global__ void __launch_bounds__(kT)
stem_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const long long* __restrict__ labels,
                  int in_mode, float inv_temp, int N, int H, int W, float* __restrict__ dW) {
  pdl_entry();
  constexpr int COUT = 16, TW = 32, TH = 8, HW_T = TW + 2, HH_T = TH + 2;
  constexpr int KT = CIN == 4 ? 9 : 3;
  __shared__ float s_in[CIN][HH_T][HW_T + 1];
  __shared__ __align__(16) float s_dy[TH * TW][COUT];
  __shared__ float s_red[kT / 32][COUT * CIN * 9];
  const int g = threadIdx.x >> 4, t = threadIdx.x & 15;
  const int cq = t & 3, kq = t >> 2;
  const bool k_active = CIN == 4 || kq < 3;
  float acc[4][KT];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < KT; ++b) acc[a][b] = 0.0f;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  const int64_t HW = (int64_t)H * W, num_tiles = (int64_t)N * tiles_x * tiles_y;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t n = tile / (tiles_x * tiles_y);
    const int rem = (int)(tile - n * tiles_x * tiles_y);
    const int y0 = (rem / tiles_x) * TH, x0 = (rem % tiles_x) * TW;
    __syncthreads();                                      // previous iteration's readers are done
    for (int i = threadIdx.x; i < HH_T * HW_T; i += kT) {
      const int hy = i / HW_T, hx = i - hy * HW_T;
      float v[CIN];
      stem_input_at<CIN>(x, labels, in_mode, inv_temp, n, y0 + hy - 1, x0 + hx - 1, H, W, v);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) s_in[ci][hy][hx] = v[ci];
    }
    for (int i = threadIdx.x; i < TH * TW * 2; i += kT) {  // one 16-byte half (8 channels) per thread
      const int p = i >> 1, half = i & 1;
      const int py = p / TW, px = p - py * TW;
      float f[8];
      if (y0 + py < H && x0 + px < W) {
        unpack8(*reinterpret_cast<const uint4*>(dy + ((n * 2 + half) * HW + (int64_t)(y0 + py) * W + x0 + px) * 8), f);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.0f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s_dy[p][half * 8 + j] = f[j];
    }
    __syncthreads();
    if (k_active) {
#pragma unroll 2
      for (int p = g; p < TH * TW; p += 16) {
        const int py = p / TW, px = p - py * TW;
        const float4 d = *reinterpret_cast<const float4*>(&s_dy[p][cq * 4]);
        float in[KT];
        if (CIN == 4) {
#pragma unroll
          for (int k = 0; k < 9; ++k) in[k] = s_in[kq][py + k / 3][px + k % 3];
        } else {
#pragma unroll
          for (int k = 0; k < 3; ++k) in[k] = s_in[0][py + kq][px + k];
        }
#pragma unroll
        for (int k = 0; k < KT; ++k) {
          acc[0][k] = fmaf(d.x, in[k], acc[0][k]);
          acc[1][k] = fmaf(d.y, in[k], acc[1][k]);
          acc[2][k] = fmaf(d.z, in[k], acc[2][k]);
          acc[3][k] = fmaf(d.w, in[k], acc[3][k]);
        }
      }
    }
  }
  // lanes l and l^16 share (cq, kq); then the 8 warps meet in shared memory
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < KT; ++b) {
      float v = acc[a][b] + __shfl_xor_sync(0xffffffffu, acc[a][b], 16);
      if (lane < 16 && k_active) {
        const int co = cq * 4 + a;
        const int kidx = CIN == 4 ? kq * 9 + b : kq * 3 + b;          // ci*9 + tap
        s_red[warp][co * (CIN * 9) + kidx] = v;
      }
    }
  __syncthreads();
  for (int i = threadIdx.x; i < COUT * CIN * 9; i += kT) {
    float tsum = 0.0f;
    for (int wq = 0; wq < kT / 32; ++wq) tsum += s_red[wq][i];
    atomicAdd(dW + i, tsum);
  }
}

// Input gradient of the stem (CIN = 4: the STN input): d_in[p][ci] = sum_{r,s,co} w[co][ci][r][s] * dy[p - (r-1,s-1)][co];
// in_mode 1 chains through softmax(x / T): dx = (s * (d_in - sum_j s_j d_in_j)) / T.  Output planar fp32 [N,4,H,W].
template <int CIN>
__global__ void __launch_bounds__(kT)
stem_dgrad_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                  int in_mode, float inv_temp, int N, int H, int W, float* __restrict__ dx) {
  pdl_entry();
  constexpr int COUT = 16;
  __shared__ __align__(16) float sw[9][COUT][CIN];          // [tap][co][ci]
  for (int i = threadIdx.x; i < COUT * CIN * 9; i += kT) {
    const int co = i / (CIN * 9), rest = i - co * (CIN * 9), ci = rest / 9, tap = rest - ci * 9;
    sw[tap][co][ci] = w[i];
  }
  __syncthreads();
  const int64_t HW = (int64_t)H * W, total = (int64_t)N * HW;
  for (int64_t i = (int64_t)blockIdx.x * kT + threadIdx.x; i < total; i += (int64_t)gridDim.x * kT) {
    const int64_t n = i / HW;
    const int pix = (int)(i - n * HW);
    const int yy = pix / W, xx = pix - yy * W;
    float d[CIN];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) d[ci] = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int qy = yy - (r - 1);
      if (qy < 0 || qy >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int qx = xx - (s - 1);
        if (qx < 0 || qx >= W) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float f[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(dy + ((n * 2 + half) * HW + (int64_t)qy * W + qx) * 8)), f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float* wv = sw[r * 3 + s][half * 8 + j];
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) d[ci] = fmaf(wv[ci], f[j], d[ci]);
          }
        }
      }
    }
    if (in_mode == 1) {
      float v[CIN];
      stem_input_at<CIN>(x, nullptr, 1, inv_temp, n, yy, xx, H, W, v);
      float dot = 0.0f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) dot = fmaf(v[ci], d[ci], dot);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) d[ci] = v[ci] * (d[ci] - dot) * inv_temp;
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) dx[(n * CIN + ci) * HW + pix] = d[ci];
  }
}

inline unsigned grid_for(int64_t total) {
  return (unsigned)std::min<int64_t>(ceil_div(total, kT), (int64_t)sm_count() * 16);
}

int check_c8(const char* who, const void* a, const void* b, int64_t N, int64_t C, int64_t H, int64_t W) {
  CTL_REQUIRE(a && b && N > 0 && C > 0 && C % 8 == 0 && H > 0 && W > 0, CTL_ERR_INVALID, "%s: bad arguments", who);
  CTL_REQUIRE(N * (C / 8) <= 0x7fffffff, CTL_ERR_UNSUPPORTED, "%s: too many planes", who);
  return CTL_OK;
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" size_t ctl_reduce_workspace_bytes(int64_t N, int64_t C) {
  // 64 = upper bound of plane_splits()
  return (N > 0 && C > 0) ? (size_t)64 * (size_t)(N * C) * 2 * sizeof(double) : 0;
}

extern "C" size_t ctl_bn_workspace_bytes(int64_t N, int64_t C) { return ctl_reduce_workspace_bytes(N, C); }

extern "C" int ctl_bn_batch_affine_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, const float* gamma,
                                      const float* beta, float eps, void* workspace, float* scale, float* shift,
                                      float* mean_out, float* var_out, float* running_mean, float* running_var,
                                      float momentum, void* stream) {
  if (int rc = check_c8("ctl_bn_batch_affine_c8", x, workspace, N, C, H, W)) return rc;
  CTL_REQUIRE(scale && shift, CTL_ERR_INVALID, "ctl_bn_batch_affine_c8: NULL pointer");
  CTL_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CTL_ERR_INVALID,
              "running_mean and running_var must be given together");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t planes = N * (C / 8), HW = H * W;
  const int splits = plane_splits(planes, HW, reduce_ctas_per_sm<0>());
  launch_chained(plane_reduce_kernel<0>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
      (const uint4*)x, nullptr, nullptr, nullptr, (double*)workspace, planes, HW, splits, 0, nullptr, nullptr, (int)(C / 8),
      nullptr);
  CTL_CUDA_OK(cudaGetLastError(), "bn_partial_stats launch");
  launch_chained(bn_fwd_finalize_kernel, (unsigned)ceil_div(C, 4), 128, 0, st)((const double*)workspace, planes, splits, (int)N, (int)C,
                                                                   (double)(N * HW), gamma, beta, eps, scale, shift,
                                                                   mean_out, var_out, running_mean, running_var, momentum);
  CTL_CUDA_OK(cudaGetLastError(), "bn_finalize launch");
  return CTL_OK;
}

extern "C" int ctl_bn_affine_from_sums(const double* sums, int64_t C, int64_t count, const float* gamma, const float* beta,
                                       float eps, float* scale, float* shift, float* mean_out, float* var_out,
                                       float* running_mean, float* running_var, float momentum, void* stream) {
  CTL_REQUIRE(sums && scale && shift && C > 0 && count > 0, CTL_ERR_INVALID, "ctl_bn_affine_from_sums: bad arguments");
  CTL_REQUIRE((running_mean == nullptr) == (running_var == nullptr), CTL_ERR_INVALID,
              "running_mean and running_var must be given together");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  launch_chained(bn_fwd_from_sums_kernel, (unsigned)ceil_div(C, 128), 128, 0, (cudaStream_t)stream)(
      sums, (int)C, (double)count, gamma, beta, eps, scale, shift, mean_out, var_out, running_mean, running_var, momentum);
  CTL_CUDA_OK(cudaGetLastError(), "bn_fwd_from_sums launch");
  return CTL_OK;
}

extern "C" int ctl_channel_sums_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* workspace,
                                   float* sum_out, float* sumsq_out, void* stream) {
  if (int rc = check_c8("ctl_channel_sums_c8", x, workspace, N, C, H, W)) return rc;
  CTL_REQUIRE(sum_out || sumsq_out, CTL_ERR_INVALID, "ctl_channel_sums_c8: no output requested");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t planes = N * (C / 8), HW = H * W;
  const int splits = plane_splits(planes, HW, reduce_ctas_per_sm<0>());
  launch_chained(plane_reduce_kernel<0>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
      (const uint4*)x, nullptr, nullptr, nullptr, (double*)workspace, planes, HW, splits, 0, nullptr, nullptr, (int)(C / 8),
      nullptr);
  CTL_CUDA_OK(cudaGetLastError(), "plane_reduce launch");
  launch_chained(channel_sum_finalize_kernel, (unsigned)ceil_div(C, 4), 128, 0, st)((const double*)workspace, planes, splits, (int)N,
                                                                        (int)C, sum_out, sumsq_out);
  CTL_CUDA_OK(cudaGetLastError(), "channel_sum_finalize launch");
  return CTL_OK;
}

extern "C" int ctl_bn_bwd_reduce_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H,
                                    int64_t W, int act, const float* mean, const float* var, float eps,
                                    const float* gamma, void* workspace, void* dv_out, float* coef, float* dgamma,
                                    float* dbeta, const float* act_scale, const float* act_shift, void* stream) {
  if (int rc = check_c8("ctl_bn_bwd_reduce_c8", dy, a, N, C, H, W)) return rc;
  CTL_REQUIRE(mean && var && workspace && coef, CTL_ERR_INVALID, "ctl_bn_bwd_reduce_c8: NULL pointer");
  CTL_REQUIRE(act >= CTL_ACT_NONE && act <= CTL_ACT_RELU, CTL_ERR_INVALID, "activation %d has no BN backward", act);
  CTL_REQUIRE(h != nullptr || dv_out == nullptr, CTL_ERR_INVALID, "dv_out needs h");
  CTL_REQUIRE((act_scale == nullptr) == (act_shift == nullptr) && (h == nullptr || act_scale == nullptr), CTL_ERR_INVALID,
              "ctl_bn_bwd_reduce_c8: give either h or (act_scale, act_shift)");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t planes = N * (C / 8), HW = H * W;
  if (h != nullptr) {
    const int splits = plane_splits(planes, HW, reduce_ctas_per_sm<1, true>());
    launch_chained(plane_reduce_kernel<1, true>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
        (const uint4*)dy, (const uint4*)h, (const uint4*)a, (uint4*)dv_out, (double*)workspace, planes, HW, splits, act,
        act_scale, act_shift, (int)(C / 8), nullptr);
  }
  const int splits = plane_splits(planes, HW, h != nullptr ? reduce_ctas_per_sm<1, true>() : reduce_ctas_per_sm<1, false>());
  if (h == nullptr)
    launch_chained(plane_reduce_kernel<1, false>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
        (const uint4*)dy, nullptr, (const uint4*)a, nullptr, (double*)workspace, planes, HW, splits, act,
        act_scale, act_shift, (int)(C / 8), nullptr);
  CTL_CUDA_OK(cudaGetLastError(), "bn_bwd_reduce launch");
  launch_chained(bn_bwd_finalize_kernel, (unsigned)ceil_div(C, 4), 128, 0, st)((const double*)workspace, planes, splits, (int)N, (int)C,
                                                                   (double)(N * HW), mean, var, eps, gamma, coef, dgamma,
                                                                   dbeta);
  CTL_CUDA_OK(cudaGetLastError(), "bn_bwd_finalize launch");
  return CTL_OK;
}

extern "C" int ctl_bn_bwd_apply_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H,
                                   int64_t W, int act, const float* coef, void* da, const float* act_scale,
                                   const float* act_shift, void* stream) {
  if (int rc = check_c8("ctl_bn_bwd_apply_c8", dy, a, N, C, H, W)) return rc;
  CTL_REQUIRE(coef && da, CTL_ERR_INVALID, "ctl_bn_bwd_apply_c8: NULL pointer");
  CTL_REQUIRE((act_scale == nullptr) == (act_shift == nullptr) && (h == nullptr || act_scale == nullptr), CTL_ERR_INVALID,
              "ctl_bn_bwd_apply_c8: give either h or (act_scale, act_shift)");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t total = N * (C / 8) * H * W;
  (void)total;
  return launch_bn_bwd_apply((const uint4*)dy, (const uint4*)h, (const uint4*)a, coef, (uint4*)da, N, (int)(C / 8), H * W, act,
                             act_scale, act_shift, BnBwdTotals{}, (cudaStream_t)stream);
}

extern "C" int ctl_bn_bwd_c8(const void* dy, const void* h, const void* a, int64_t N, int64_t C, int64_t H, int64_t W,
                             int act, const float* mean, const float* var, float eps, const float* gamma, double* totals,
                             void* dv_out, void* da, float* dgamma, float* dbeta, const float* act_scale,
                             const float* act_shift, void* stream) {
  if (int rc = check_c8("ctl_bn_bwd_c8", dy, a, N, C, H, W)) return rc;
  CTL_REQUIRE(mean && var && totals && da, CTL_ERR_INVALID, "ctl_bn_bwd_c8: NULL pointer");
  CTL_REQUIRE(C <= kMaxBnC, CTL_ERR_UNSUPPORTED, "ctl_bn_bwd_c8: at most %d channels (got %lld)", kMaxBnC, (long long)C);
  CTL_REQUIRE(act >= CTL_ACT_NONE && act <= CTL_ACT_RELU, CTL_ERR_INVALID, "activation %d has no BN backward", act);
  CTL_REQUIRE(h != nullptr || dv_out == nullptr, CTL_ERR_INVALID, "dv_out needs h");
  CTL_REQUIRE((act_scale == nullptr) == (act_shift == nullptr) && (h == nullptr || act_scale == nullptr), CTL_ERR_INVALID,
              "ctl_bn_bwd_c8: give either h or (act_scale, act_shift)");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t planes = N * (C / 8), HW = H * W, total = planes * HW;
  const bool items_form = HW >= 2048 && HW < ((int64_t)1 << 30) && N < ((int64_t)1 << 20) &&
                          (act == CTL_ACT_NONE || act == CTL_ACT_LRELU || act == CTL_ACT_RELU);
  if (items_form) {
    // [with h][U = 4]: CTAs per SM of each form
    static int resident[2][2] = {{0, 0}, {0, 0}};
    const int U = (HW >= 8192 && h == nullptr) ? 4 : 2;     // three tensors per position with h: two positions in flight
    int& res = resident[h != nullptr][U == 4];
    if (res == 0) {
      if (h != nullptr) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, bn_bwd_reduce_items_kernel<2, true>, kT, 0);
      else if (U == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, bn_bwd_reduce_items_kernel<4, false>, kT, 0);
      else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&res, bn_bwd_reduce_items_kernel<2, false>, kT, 0);
      if (res <= 0) res = 3;
    }
    const int ipp = (int)ceil_div(HW, (int64_t)U * kT);
    const unsigned grid = (unsigned)std::min<int64_t>(planes * ipp, (int64_t)sm_count() * res);
    if (h != nullptr)
      launch_chained(bn_bwd_reduce_items_kernel<2, true>, grid, kT, 0, st)(
          (const uint4*)dy, (const uint4*)h, (const uint4*)a, (uint4*)dv_out, (int)N, (int)(C / 8), (int)HW, ipp, act, act_scale,
          act_shift, totals);
    else if (U == 4)
      launch_chained(bn_bwd_reduce_items_kernel<4, false>, grid, kT, 0, st)(
          (const uint4*)dy, nullptr, (const uint4*)a, nullptr, (int)N, (int)(C / 8), (int)HW, ipp, act, act_scale, act_shift, totals);
    else
      launch_chained(bn_bwd_reduce_items_kernel<2, false>, grid, kT, 0, st)(
          (const uint4*)dy, nullptr, (const uint4*)a, nullptr, (int)N, (int)(C / 8), (int)HW, ipp, act, act_scale, act_shift, totals);
  } else if (h != nullptr) {
    const int splits = plane_splits(planes, HW, reduce_ctas_per_sm<1, true>());
    launch_chained(plane_reduce_kernel<1, true>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
        (const uint4*)dy, (const uint4*)h, (const uint4*)a, (uint4*)dv_out, nullptr, planes, HW, splits, act, act_scale,
        act_shift, (int)(C / 8), totals);
  } else {
    const int splits = plane_splits(planes, HW, reduce_ctas_per_sm<1, false>());
    launch_chained(plane_reduce_kernel<1, false>, dim3((unsigned)planes, (unsigned)splits), kT, 0, st)(
        (const uint4*)dy, nullptr, (const uint4*)a, nullptr, nullptr, planes, HW, splits, act, act_scale,
        act_shift, (int)(C / 8), totals);
  }
  CTL_CUDA_OK(cudaGetLastError(), "bn_bwd reduce launch");
  const BnBwdTotals bt = {totals, mean, var, gamma, dgamma, dbeta, (double)(N * HW), eps};
  (void)total;
  // dv materialised (the residual tail feeds it to the shortcut branch too): the apply pass reads it, activation done
  if (dv_out != nullptr)
    return launch_bn_bwd_apply((const uint4*)dv_out, nullptr, (const uint4*)a, nullptr, (uint4*)da, N, (int)(C / 8), HW, act,
                               nullptr, nullptr, bt, st);
  return launch_bn_bwd_apply((const uint4*)dy, (const uint4*)h, (const uint4*)a, nullptr, (uint4*)da, N, (int)(C / 8), HW, act,
                             act_scale, act_shift, bt, st);
}

extern "C" int ctl_act_bwd_c8(const void* dy, const void* h, int64_t N, int64_t C, int64_t H, int64_t W, int act,
                              void* dv, void* stream) {
  if (int rc = check_c8("ctl_act_bwd_c8", dy, h, N, C, H, W)) return rc;
  CTL_REQUIRE(dv, CTL_ERR_INVALID, "ctl_act_bwd_c8: NULL pointer");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t total = N * (C / 8) * H * W;
  launch_chained(act_bwd_kernel, grid_for(total), kT, 0, (cudaStream_t)stream)((const uint4*)dy, (const uint4*)h, (uint4*)dv, total, act);
  CTL_CUDA_OK(cudaGetLastError(), "act_bwd launch");
  return CTL_OK;
}

extern "C" int ctl_downsample2x_sum_c8(const void* dy, int64_t N, int64_t C, int64_t H, int64_t W, void* dx, void* stream) {
  if (int rc = check_c8("ctl_downsample2x_sum_c8", dy, dx, N, C, H, W)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t planes = N * (C / 8);
  launch_chained(downsample2x_sum_kernel, grid_for(planes * H * W), kT, 0, (cudaStream_t)stream)((const uint4*)dy, (uint4*)dx, planes,
                                                                                   (int)H, (int)W);
  CTL_CUDA_OK(cudaGetLastError(), "downsample2x_sum launch");
  return CTL_OK;
}

extern "C" int ctl_zero_stuff2x_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream) {
  if (int rc = check_c8("ctl_zero_stuff2x_c8", x, y, N, C, H, W)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t planes = N * (C / 8);
  launch_chained(zero_stuff2x_kernel, grid_for(planes * H * W), kT, 0, (cudaStream_t)stream)((const uint4*)x, (uint4*)y, planes, (int)H,
                                                                               (int)W);
  CTL_CUDA_OK(cudaGetLastError(), "zero_stuff2x launch");
  return CTL_OK;
}

extern "C" int ctl_split_parity2x2_c8(const void* x, int64_t N, int64_t C, int64_t H, int64_t W, void* y, void* stream) {
  if (int rc = check_c8("ctl_split_parity2x2_c8", x, y, N, C, H, W)) return rc;
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t planes = N * (C / 8);
  launch_chained(split_parity2x2_kernel, grid_for(planes * H * W), kT, 0, (cudaStream_t)stream)((const uint4*)x, (uint4*)y, planes,
                                                                                  (int)H, (int)W);
  CTL_CUDA_OK(cudaGetLastError(), "split_parity2x2 launch");
  return CTL_OK;
}

extern "C" int ctl_head_bwd_c8(const float* dy, const float* y, const void* x, int64_t N, int64_t Cin, int64_t H,
                               int64_t W, const float* weight, int64_t Cout, int act, void* dx, float* dW, float* db,
                               void* stream) {
  CTL_REQUIRE(dy && x && weight && dx && dW && db && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID,
              "ctl_head_bwd_c8: bad arguments");
  CTL_REQUIRE(Cin == 16 && (Cout == 1 || Cout == 4), CTL_ERR_UNSUPPORTED,
              "ctl_head_bwd_c8 handles Cin 16 -> Cout in {1,4} (got %lld -> %lld)", (long long)Cin, (long long)Cout);
  CTL_REQUIRE(act == CTL_ACT_NONE || (act == CTL_ACT_SIGMOID && y), CTL_ERR_INVALID,
              "ctl_head_bwd_c8: act must be none, or sigmoid with the forward output y");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(N * H * W, kT), (int64_t)sm_count() * 4);
  if (Cout == 1)
    launch_chained(head_bwd_kernel<1>, grid, kT, 0, st)(dy, y, (const __nv_bfloat16*)x, weight, (__nv_bfloat16*)dx, dW, db, (int)N, H * W, act);
  else
    launch_chained(head_bwd_kernel<4>, grid, kT, 0, st)(dy, y, (const __nv_bfloat16*)x, weight, (__nv_bfloat16*)dx, dW, db, (int)N, H * W, act);
  CTL_CUDA_OK(cudaGetLastError(), "head_bwd launch");
  return CTL_OK;
}

extern "C" int ctl_stem_input_c8(const float* x, const int64_t* labels, int in_mode, float temperature, int64_t N,
                                 int64_t Cin, int64_t H, int64_t W, void* out, void* stream) {
  CTL_REQUIRE(out && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_stem_input_c8: bad arguments");
  CTL_REQUIRE(in_mode >= 0 && in_mode <= 2 && (in_mode == 2 ? labels != nullptr : x != nullptr), CTL_ERR_INVALID,
              "ctl_stem_input_c8: in_mode %d needs %s", in_mode, in_mode == 2 ? "labels" : "x");
  CTL_REQUIRE(Cin == 1 || Cin == 4, CTL_ERR_UNSUPPORTED, "ctl_stem_input_c8 handles Cin in {1,4} (got %lld)", (long long)Cin);
  CTL_REQUIRE(temperature > 0.0f && aligned16(out), CTL_ERR_INVALID, "temperature must be positive, out 16-byte aligned");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(N * H * W);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  if (Cin == 1)
    launch_chained(stem_input_c8_kernel<1>, grid, kT, 0, st)(x, lab, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, (uint4*)out);
  else
    launch_chained(stem_input_c8_kernel<4>, grid, kT, 0, st)(x, lab, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, (uint4*)out);
  CTL_CUDA_OK(cudaGetLastError(), "stem_input launch");
  return CTL_OK;
}

extern "C" int ctl_stem_wgrad_c8(const void* dy, const float* x, const int64_t* labels, int in_mode, float temperature,
                                 int64_t N, int64_t Cin, int64_t H, int64_t W, float* dW, void* stream) {
  CTL_REQUIRE(dy && dW && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_stem_wgrad_c8: bad arguments");
  CTL_REQUIRE(in_mode >= 0 && in_mode <= 2 && (in_mode == 2 ? labels != nullptr : x != nullptr), CTL_ERR_INVALID,
              "ctl_stem_wgrad_c8: in_mode %d needs %s", in_mode, in_mode == 2 ? "labels" : "x");
  CTL_REQUIRE(Cin == 1 || Cin == 4, CTL_ERR_UNSUPPORTED, "ctl_stem_wgrad_c8 handles Cin in {1,4} (got %lld)", (long long)Cin);
  CTL_REQUIRE(temperature > 0.0f, CTL_ERR_INVALID, "temperature must be positive");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t tiles = N * ceil_div(W, 32) * ceil_div(H, 8);
  const unsigned grid = (unsigned)std::min<int64_t>(tiles, (int64_t)sm_count() * 2);
  const long long* lab = reinterpret_cast<const long long*>(labels);
  if (Cin == 1)
    launch_chained(stem_wgrad_kernel<1>, grid, kT, 0, st)((const __nv_bfloat16*)dy, x, lab, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, dW);
  else
    launch_chained(stem_wgrad_kernel<4>, grid, kT, 0, st)((const __nv_bfloat16*)dy, x, lab, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, dW);
  CTL_CUDA_OK(cudaGetLastError(), "stem_wgrad launch");
  return CTL_OK;
}

extern "C" int ctl_stem_dgrad_c8(const void* dy, const float* x, int in_mode, float temperature, int64_t N, int64_t Cin,
                                 int64_t H, int64_t W, const float* weight, float* dx, void* stream) {
  CTL_REQUIRE(dy && weight && dx && N > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_stem_dgrad_c8: bad arguments");
  CTL_REQUIRE(in_mode == 0 || (in_mode == 1 && x != nullptr), CTL_ERR_INVALID,
              "ctl_stem_dgrad_c8: in_mode must be 0, or 1 with the forward input x (a label map has no gradient)");
  CTL_REQUIRE(Cin == 1 || Cin == 4, CTL_ERR_UNSUPPORTED, "ctl_stem_dgrad_c8 handles Cin in {1,4} (got %lld)", (long long)Cin);
  CTL_REQUIRE(temperature > 0.0f, CTL_ERR_INVALID, "temperature must be positive");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  if (aligned16(dy) && N <= 65535 && stem_dgrad_tensor_path_handles(H, W))
    return stem_dgrad_tensor_path(dy, x, in_mode, 1.0f / temperature, (int)N, (int)Cin, (int)H, (int)W, weight, dx, st);
  const unsigned grid = grid_for(N * H * W);
  if (Cin == 1)
    launch_chained(stem_dgrad_kernel<1>, grid, kT, 0, st)((const __nv_bfloat16*)dy, x, weight, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, dx);
  else
    launch_chained(stem_dgrad_kernel<4>, grid, kT, 0, st)((const __nv_bfloat16*)dy, x, weight, in_mode, 1.0f / temperature, (int)N, (int)H, (int)W, dx);
  CTL_CUDA_OK(cudaGetLastError(), "stem_dgrad launch");
  return CTL_OK;
}

extern "C" int ctl_bn_bwd_apply_totals_c8(const void* dy, const void* a, int64_t N, int64_t C, int64_t H, int64_t W, int act,
                                          const float* mean, const float* var, float eps, const float* gamma,
                                          const double* totals, void* da, float* dgamma, float* dbeta,
                                          const float* act_scale, const float* act_shift, void* stream) {
  if (int rc = check_c8("ctl_bn_bwd_apply_totals_c8", dy, a, N, C, H, W)) return rc;
  CTL_REQUIRE(mean && var && totals && da && act_scale && act_shift, CTL_ERR_INVALID, "ctl_bn_bwd_apply_totals_c8: NULL pointer");
  CTL_REQUIRE(C <= kMaxBnC, CTL_ERR_UNSUPPORTED, "ctl_bn_bwd_apply_totals_c8: at most %d channels", kMaxBnC);
  if (sm_count() < 0) return CTL_ERR_CUDA;
  const int64_t HW = H * W, total = N * (C / 8) * HW;
  const BnBwdTotals bt = {totals, mean, var, gamma, dgamma, dbeta, (double)(N * HW), eps};
  (void)total;
  return launch_bn_bwd_apply((const uint4*)dy, nullptr, (const uint4*)a, nullptr, (uint4*)da, N, (int)(C / 8), HW, act, act_scale,
                             act_shift, bt, (cudaStream_t)stream);
}

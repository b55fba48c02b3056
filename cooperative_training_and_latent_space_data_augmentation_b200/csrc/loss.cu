// Fused 2-D cross entropy on planar fp32 logits with integer label maps.
//
// Replaces, for label-map targets without a mask / class weights,
//   medseg/models/custom_loss.py:706-741   cross_entropy_2D (training CE, via basic_loss_fn :8-19)
//   medseg/models/model_util.py:104-135    cross_entropy_2D (saliency 'ce')
// which run log_softmax -> NHWC transpose copy -> nll_loss(sum) -> divide (and, in custom_loss.py:729, a
// device->host sync for the divisor).  Here: ONE streaming pass over the logits and labels for the loss
// (sum_p [logsumexp(x_p) - x_p[label_p]] * scale) and ONE for its gradient (softmax - onehot) * gout * scale.
// HBM-bound: forward reads 4*C + 8 bytes per pixel, backward reads the same and writes 4*C.
//
// Labels outside [0, C) contribute neither loss nor gradient (nll_loss's ignore_index = -100 behaves like that; any
// other out-of-range label makes torch raise -- the reference never produces one).
#include <algorithm>

#include "ctl_common.cuh"

namespace ctl {
namespace {

constexpr int kT = 256;

template <int C>
__device__ __forceinline__ float pixel_lse(const float (&x)[C]) {
  float m = x[0];
#pragma unroll
  for (int c = 1; c < C; ++c) m = fmaxf(m, x[c]);
  float s = 0.0f;
#pragma unroll
  for (int c = 0; c < C; ++c) s += expf(x[c] - m);
  return m + logf(s);
}

// ws[0]: fp64 running sum, ws[1]: CTA ticket counter -- both zero on entry and zero again on exit (self-cleaning)
template <int C, int VEC>
__global__ void __launch_bounds__(kT)
ce2d_fwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int64_t groups_per_img, int64_t HW,
                int64_t total_groups, double scale, unsigned long long* __restrict__ ws, float* __restrict__ out) {
  pdl_entry();
  float acc = 0.0f;
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total_groups; g += (int64_t)gridDim.x * kT) {
    const int64_t n = g / groups_per_img;
    const int64_t p0 = (g - n * groups_per_img) * VEC;
    float x[C][VEC];
#pragma unroll
    for (int c = 0; c < C; ++c) load_as_float<float, VEC>(logits + (n * C + c) * HW + p0, x[c]);
    long long lab[VEC];
    if constexpr (VEC == 4) {
      const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(labels + n * HW + p0));
      const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(labels + n * HW + p0) + 1);
      lab[0] = a.x; lab[1] = a.y; lab[2] = b.x; lab[3] = b.y;
    } else {
      lab[0] = __ldcs(labels + n * HW + p0);
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float px[C];
#pragma unroll
      for (int c = 0; c < C; ++c) px[c] = x[c][v];
      const float lse = pixel_lse<C>(px);
      float picked = 0.0f;
      bool ok = false;
#pragma unroll
      for (int c = 0; c < C; ++c)
        if (lab[v] == c) { picked = px[c]; ok = true; }
      if (ok) acc += lse - picked;
    }
  }
  __shared__ double red[kT / 32];
  double d = (double)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kT / 32; ++w) t += red[w];
    double* sum = reinterpret_cast<double*>(ws);
    atomicAdd(sum, t);
    __threadfence();
    const unsigned long long ticket = atomicAdd(ws + 1, 1ull);
    if (ticket == gridDim.x - 1) {                 // last CTA: publish and leave the workspace zeroed
      __threadfence();
      const double total = atomicAdd(sum, 0.0);
      out[0] = (float)(total * scale);
      *sum = 0.0;
      ws[1] = 0ull;
      __threadfence();
    }
  }
}

template <int C, int VEC>
__global__ void __launch_bounds__(kT)
ce2d_bwd_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, const float* __restrict__ gout,
                float scale, int64_t groups_per_img, int64_t HW, int64_t total_groups, float* __restrict__ dlogits) {
  pdl_entry();
  const float gs = (gout ? __ldg(gout) : 1.0f) * scale;
  for (int64_t g = (int64_t)blockIdx.x * kT + threadIdx.x; g < total_groups; g += (int64_t)gridDim.x * kT) {
    const int64_t n = g / groups_per_img;
    const int64_t p0 = (g - n * groups_per_img) * VEC;
    float x[C][VEC];
#pragma unroll
    for (int c = 0; c < C; ++c) load_as_float<float, VEC>(logits + (n * C + c) * HW + p0, x[c]);
    long long lab[VEC];
    if constexpr (VEC == 4) {
      const longlong2 a = __ldcs(reinterpret_cast<const longlong2*>(labels + n * HW + p0));
      const longlong2 b = __ldcs(reinterpret_cast<const longlong2*>(labels + n * HW + p0) + 1);
      lab[0] = a.x; lab[1] = a.y; lab[2] = b.x; lab[3] = b.y;
    } else {
      lab[0] = __ldcs(labels + n * HW + p0);
    }
    float d[C][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float px[C];
#pragma unroll
      for (int c = 0; c < C; ++c) px[c] = x[c][v];
      const float lse = pixel_lse<C>(px);
      const bool ok = lab[v] >= 0 && lab[v] < C;
#pragma unroll
      for (int c = 0; c < C; ++c) d[c][v] = ok ? gs * (expf(px[c] - lse) - (lab[v] == c ? 1.0f : 0.0f)) : 0.0f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_from_float<float, VEC>(dlogits + (n * C + c) * HW + p0, d[c]);
  }
}

template <int C, int VEC>
int launch_ce(bool backward, const float* logits, const int64_t* labels, int64_t N, int64_t HW, double scale,
              const float* gout, unsigned long long* ws, float* out, float* dlogits, cudaStream_t st) {
  const int64_t groups_per_img = HW / VEC, total = N * groups_per_img;
  const int sms = sm_count();
  if (sms < 0) return CTL_ERR_CUDA;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total, kT), (int64_t)sms * 8));
  if (!backward) {
    launch_chained(ce2d_fwd_kernel<C, VEC>, grid, kT, 0, st)(logits, labels, groups_per_img, HW, total, scale, ws, out);
  } else {
    launch_chained(ce2d_bwd_kernel<C, VEC>, grid, kT, 0, st)(logits, labels, gout, (float)scale, groups_per_img, HW, total, dlogits);
  }
  CTL_CUDA_OK(cudaGetLastError(), "ce2d launch");
  return CTL_OK;
}

template <int VEC>
int dispatch_ce(int C, bool backward, const float* logits, const int64_t* labels, int64_t N, int64_t HW, double scale,
                const float* gout, unsigned long long* ws, float* out, float* dlogits, cudaStream_t st) {
  switch (C) {
    case 2: return launch_ce<2, VEC>(backward, logits, labels, N, HW, scale, gout, ws, out, dlogits, st);
    case 3: return launch_ce<3, VEC>(backward, logits, labels, N, HW, scale, gout, ws, out, dlogits, st);
    case 4: return launch_ce<4, VEC>(backward, logits, labels, N, HW, scale, gout, ws, out, dlogits, st);
    case 8: return launch_ce<8, VEC>(backward, logits, labels, N, HW, scale, gout, ws, out, dlogits, st);
    default:
      set_error("ctl_ce2d: number of classes must be 2, 3, 4 or 8 (got %d)", C);
      return CTL_ERR_UNSUPPORTED;
  }
}

int ce_common(bool backward, const float* logits, const int64_t* labels, int64_t N, int64_t C, int64_t H, int64_t W,
              double scale, const float* gout, void* ws, float* out, float* dlogits, void* stream) {
  CTL_REQUIRE(logits && labels && (backward ? dlogits != nullptr : (ws != nullptr && out != nullptr)), CTL_ERR_INVALID,
              "ctl_ce2d: NULL pointer");
  CTL_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0, CTL_ERR_INVALID, "ctl_ce2d: bad shape N=%lld C=%lld H=%lld W=%lld",
              (long long)N, (long long)C, (long long)H, (long long)W);
  const int64_t HW = H * W;
  const bool vec = HW % 4 == 0 && aligned16(logits) && aligned16(labels) && (!backward || aligned16(dlogits));
  cudaStream_t st = (cudaStream_t)stream;
  if (vec) return dispatch_ce<4>((int)C, backward, logits, labels, N, HW, scale, gout, (unsigned long long*)ws, out, dlogits, st);
  return dispatch_ce<1>((int)C, backward, logits, labels, N, HW, scale, gout, (unsigned long long*)ws, out, dlogits, st);
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" int ctl_ce2d_fwd(const float* logits, const int64_t* labels, int64_t N, int64_t C, int64_t H, int64_t W,
                            double scale, void* workspace16, float* loss_out, void* stream) {
  return ce_common(false, logits, labels, N, C, H, W, scale, nullptr, workspace16, loss_out, nullptr, stream);
}

extern "C" int ctl_ce2d_bwd(const float* logits, const int64_t* labels, int64_t N, int64_t C, int64_t H, int64_t W,
                            double scale, const float* grad_out, float* dlogits, void* stream) {
  return ce_common(true, logits, labels, N, C, H, W, scale, grad_out, nullptr, nullptr, dlogits, stream);
}

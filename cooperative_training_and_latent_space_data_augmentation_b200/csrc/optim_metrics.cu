// Step-level streaming kernels around the conv blocks:
//
//   ctl_adam_flat          ONE multi-tensor Adam launch over the flat parameter / gradient / moment buffers of all five
//                          sub-networks.  Replaces the five torch.optim.Adam(...).step() calls of
//                          medseg/models/advanced_triplet_recon_segmentation_model.py:774-785 (optimize_all_params) and
//                          medseg/train_adv_supervised_segmentation_triplet.py:230-231; the data-parallel 1/world
//                          average and the next step's gradient clearing are folded into the same pass.
//   ctl_sse_fwd / _bwd     scale * sum (pred - target)^2 and its gradient: 0.5 * MSELoss of advanced...model.py:443-447
//                          (image reconstruction) and torch.mean((decoder(code) - gt)**2) of model_util.py:207-208.
//   ctl_confusion_update   argmax over classes + n x n confusion matrix on the device, replaces
//                          pred.max(1)[1].cpu().numpy() + runningScore._fast_hist / update
//                          (medseg/common_utils/metrics.py:18-28, advanced...model.py:656-659).
//   ctl_confusion_scores   runningScore.get_scores (metrics.py:30-53) from that matrix, fp64, one thread.
//
// All HBM-bound streaming passes with 16-byte accesses; reductions meet in shared memory and leave the CTA as ONE
// atomic per value.
#include <algorithm>

#include "ctl_common.cuh"

namespace ctl {
namespace {

constexpr int kT = 256;
constexpr int kMaxSeg = 8;

struct AdamSegs {
  long long begin[kMaxSeg];     // first element of segment s in the flat buffers (multiple of 4)
  long long end[kMaxSeg];       // one past its last element
  int count;
  unsigned mask;                // bit s set: segment s is stepped by this call
};

__global__ void adam_bump_kernel(float* __restrict__ steps, int count, unsigned mask) {
  pdl_entry();
  const int s = threadIdx.x;
  if (s < count && ((mask >> s) & 1u)) steps[s] += 1.0f;
}

// p, g, m, v: flat fp32 buffers of the same length; steps[s] already holds the step number t >= 1 of segment s.
// torch.optim.Adam (amsgrad=False, maximize=False):  g' = g*grad_scale (+ wd*p);  m = m + (g'-m)(1-b1);
// v = b2*v + (1-b2) g'^2;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(kT)
adam_flat_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 const AdamSegs segs, const float* __restrict__ steps, double lr, double b1d, double b2d, float eps, float wd,
                 float grad_scale, int zero_grad) {
  pdl_entry();
  __shared__ float s_step_size[kMaxSeg], s_bc2_rsqrt[kMaxSeg];
  if ((int)threadIdx.x < segs.count) {
    const double t = (double)steps[threadIdx.x];
    const double bc1 = 1.0 - pow(b1d, t), bc2 = 1.0 - pow(b2d, t);
    s_step_size[threadIdx.x] = (float)(lr / bc1);
    s_bc2_rsqrt[threadIdx.x] = (float)(1.0 / sqrt(bc2));
  }
  __syncthreads();
  const long long n4 = segs.end[segs.count - 1] >> 2;          // the buffers are padded to a multiple of 4
  // the constants as torch.optim.Adam forms them: 1 - beta in DOUBLE, then rounded to fp32 (1.0f - 0.999f is 4.7e-5 off)
  const float b2 = (float)b2d, omb1 = (float)(1.0 - b1d), omb2 = (float)(1.0 - b2d);
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n4; i += (long long)gridDim.x * kT) {
    const long long e = i << 2;
    int s = 0;
#pragma unroll
    for (int k = 1; k < kMaxSeg; ++k)
      if (k < segs.count && e >= segs.begin[k]) s = k;
    if (!((segs.mask >> s) & 1u) || e >= segs.end[s]) continue;   // segment not stepped / alignment padding
    const float4 g4 = *reinterpret_cast<const float4*>(g + e);
    float4 p4 = *reinterpret_cast<const float4*>(p + e);
    float4 m4 = *reinterpret_cast<const float4*>(m + e);
    float4 v4 = *reinterpret_cast<const float4*>(v + e);
    const float step_size = s_step_size[s], rs2 = s_bc2_rsqrt[s];
    float gg[4] = {g4.x, g4.y, g4.z, g4.w}, pp[4] = {p4.x, p4.y, p4.z, p4.w};
    float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gg[k] * grad_scale;
      if (wd != 0.0f) gk = fmaf(wd, pp[k], gk);
      mm[k] = fmaf(gk - mm[k], omb1, mm[k]);
      vv[k] = fmaf(omb2 * gk, gk, b2 * vv[k]);
      const float denom = fmaf(sqrtf(vv[k]), rs2, eps);
      pp[k] -= step_size * (mm[k] / denom);
    }
    *reinterpret_cast<float4*>(p + e) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *reinterpret_cast<float4*>(m + e) = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *reinterpret_cast<float4*>(v + e) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    if (zero_grad) *reinterpret_cast<float4*>(g + e) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- sum of squared errors ---------------------------------------------------------------------------------------------
// ws[0]: fp64 running sum, ws[1]: CTA ticket -- zero on entry, zero again on exit (as ce2d_fwd_kernel)
template <int VEC>
__global__ void __launch_bounds__(kT)
sse_fwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, long long n, double scale,
               unsigned long long* __restrict__ ws, float* __restrict__ out) {
  pdl_entry();
  float acc = 0.0f;
  const long long groups = n / VEC;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < groups; i += (long long)gridDim.x * kT) {
    float a[VEC], b[VEC];
    load_as_float<float, VEC>(pred + i * VEC, a);
    load_as_float<float, VEC>(target + i * VEC, b);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { const float d = a[k] - b[k]; acc = fmaf(d, d, acc); }
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
    if (ticket == gridDim.x - 1) {
      __threadfence();
      const double total = atomicAdd(sum, 0.0);
      out[0] = (float)(total * scale);
      *sum = 0.0;
      ws[1] = 0ull;
      __threadfence();
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(kT)
sse_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target, const float* __restrict__ gout,
               float scale2, long long n, float* __restrict__ dpred) {
  pdl_entry();
  const float gs = (gout ? __ldg(gout) : 1.0f) * scale2;
  const long long groups = n / VEC;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < groups; i += (long long)gridDim.x * kT) {
    float a[VEC], b[VEC], d[VEC];
    load_as_float<float, VEC>(pred + i * VEC, a);
    load_as_float<float, VEC>(target + i * VEC, b);
#pragma unroll
    for (int k = 0; k < VEC; ++k) d[k] = gs * (a[k] - b[k]);
    store_from_float<float, VEC>(dpred + i * VEC, d);
  }
}

// ---- confusion matrix ---------------------------------------------------------------------------------------------------
// One pixel per thread and iteration: prediction = argmax_c logits[n][c][p] (first maximum, as torch.max(1)[1] resolves
// ties on CUDA) or a given label; counts meet in a shared-memory matrix, one 64-bit atomic per cell and CTA.
template <int C>
__global__ void __launch_bounds__(kT)
confusion_kernel(const float* __restrict__ logits, const int64_t* __restrict__ pred_labels,
                 const int64_t* __restrict__ gt, long long N, long long HW, unsigned long long* __restrict__ hist,
                 uint8_t* __restrict__ labels_out) {
  pdl_entry();
  __shared__ unsigned int cell[C * C];
  for (int i = threadIdx.x; i < C * C; i += kT) cell[i] = 0u;
  __syncthreads();
  const long long total = N * HW;
  for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < total; i += (long long)gridDim.x * kT) {
    int pred;
    if (logits) {
      const long long n = i / HW, px = i - n * HW;
      const float* base = logits + n * C * HW + px;
      float best = __ldcs(base);
      pred = 0;
#pragma unroll
      for (int c = 1; c < C; ++c) {
        const float x = __ldcs(base + c * HW);
        if (x > best) { best = x; pred = c; }
      }
    } else {
      pred = (int)__ldcs(pred_labels + i);
    }
    if (labels_out) labels_out[i] = (uint8_t)pred;
    if (gt) {
      const long long t = __ldcs(gt + i);
      if (t >= 0 && t < C && pred >= 0 && pred < C) atomicAdd(&cell[(int)t * C + pred], 1u);
    }
  }
  __syncthreads();
  if (gt)
    for (int i = threadIdx.x; i < C * C; i += kT)
      if (cell[i]) atomicAdd(hist + i, (unsigned long long)cell[i]);
}

// out: [0] overall acc, [1] mean acc (nanmean), [2] freq-weighted acc, [3] mean IoU (nanmean), [4 + c] IoU of class c
__global__ void confusion_scores_kernel(const unsigned long long* __restrict__ hist, int C, double* __restrict__ out) {
  pdl_entry();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  double total = 0.0, diag = 0.0;
  for (int i = 0; i < C; ++i)
    for (int j = 0; j < C; ++j) {
      const double h = (double)hist[i * C + j];
      total += h;
      if (i == j) diag += h;
    }
  double acc_sum = 0.0, iu_sum = 0.0, fw = 0.0;
  int acc_n = 0, iu_n = 0;
  for (int c = 0; c < C; ++c) {
    double row = 0.0, col = 0.0;
    for (int j = 0; j < C; ++j) { row += (double)hist[c * C + j]; col += (double)hist[j * C + c]; }
    const double d = (double)hist[c * C + c];
    const double acc_c = row > 0.0 ? d / row : nan;
    const double den = row + col - d;
    const double iu = den > 0.0 ? d / den : nan;
    if (acc_c == acc_c) { acc_sum += acc_c; ++acc_n; }
    if (iu == iu) { iu_sum += iu; ++iu_n; }
    if (total > 0.0 && row > 0.0 && iu == iu) fw += row / total * iu;
    out[4 + c] = iu;
  }
  out[0] = total > 0.0 ? diag / total : nan;
  out[1] = acc_n ? acc_sum / acc_n : nan;
  out[2] = total > 0.0 ? fw : nan;
  out[3] = iu_n ? iu_sum / iu_n : nan;
}

int stream_grid(long long work_items) {
  const int sms = sm_count();
  if (sms < 0) return -1;
  return (int)std::max<long long>(1, std::min<long long>(ceil_div(work_items, kT), (long long)sms * 8));
}

}  // namespace
}  // namespace ctl

using namespace ctl;

extern "C" int ctl_adam_flat(float* params, float* grads, float* exp_avg, float* exp_avg_sq, const int64_t* seg_bounds_host,
                             int n_segments, unsigned seg_mask, float* steps, double lr, double beta1, double beta2, double eps,
                             double weight_decay, double grad_scale, int zero_grad, void* stream) {
  CTL_REQUIRE(params && grads && exp_avg && exp_avg_sq && seg_bounds_host && steps, CTL_ERR_INVALID,
              "ctl_adam_flat: NULL pointer");
  CTL_REQUIRE(n_segments >= 1 && n_segments <= kMaxSeg, CTL_ERR_INVALID, "ctl_adam_flat: 1..%d segments (got %d)", kMaxSeg,
              n_segments);
  CTL_REQUIRE(aligned16(params) && aligned16(grads) && aligned16(exp_avg) && aligned16(exp_avg_sq), CTL_ERR_INVALID,
              "ctl_adam_flat: buffers must be 16-byte aligned");
  CTL_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0., CTL_ERR_INVALID,
              "ctl_adam_flat: bad hyper-parameters lr=%g betas=(%g,%g) eps=%g", lr, beta1, beta2, eps);
  AdamSegs segs = {};
  segs.count = n_segments;
  segs.mask = seg_mask;
  long long prev_end = 0;
  for (int s = 0; s < n_segments; ++s) {
    segs.begin[s] = seg_bounds_host[2 * s];
    segs.end[s] = seg_bounds_host[2 * s + 1];
    CTL_REQUIRE(segs.begin[s] % 4 == 0 && segs.begin[s] >= prev_end && segs.end[s] >= segs.begin[s], CTL_ERR_INVALID,
                "ctl_adam_flat: segment %d = [%lld, %lld) must start on a multiple of 4 and follow the previous one", s,
                segs.begin[s], segs.end[s]);
    prev_end = segs.end[s];
  }
  // the last float4 of a segment may reach into its alignment padding: the buffers are allocated padded to 4
  for (int s = 0; s < n_segments; ++s) segs.end[s] = (segs.end[s] + 3) / 4 * 4;
  for (int s = 0; s + 1 < n_segments; ++s)
    CTL_REQUIRE(segs.end[s] <= segs.begin[s + 1], CTL_ERR_INVALID, "ctl_adam_flat: segments %d and %d overlap after padding", s, s + 1);
  if ((seg_mask & ((1u << n_segments) - 1u)) == 0 || prev_end == 0) return CTL_OK;
  const int grid = stream_grid(segs.end[n_segments - 1] / 4);
  if (grid < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  launch_chained(adam_bump_kernel, 1, 32, 0, st)(steps, n_segments, seg_mask);
  launch_chained(adam_flat_kernel, grid, kT, 0, st)(params, grads, exp_avg, exp_avg_sq, segs, steps, lr, beta1, beta2, (float)eps,
                                        (float)weight_decay, (float)grad_scale, zero_grad);
  CTL_CUDA_OK(cudaGetLastError(), "adam_flat launch");
  return CTL_OK;
}

extern "C" int ctl_sse_fwd(const float* pred, const float* target, int64_t n, double scale, void* workspace16,
                           float* loss_out, void* stream) {
  CTL_REQUIRE(pred && target && workspace16 && loss_out && n > 0, CTL_ERR_INVALID, "ctl_sse_fwd: NULL pointer or n <= 0");
  const int grid = stream_grid(ceil_div(n, 4));
  if (grid < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  if (n % 4 == 0 && aligned16(pred) && aligned16(target))
    launch_chained(sse_fwd_kernel<4>, grid, kT, 0, st)(pred, target, n, scale, (unsigned long long*)workspace16, loss_out);
  else
    launch_chained(sse_fwd_kernel<1>, grid, kT, 0, st)(pred, target, n, scale, (unsigned long long*)workspace16, loss_out);
  CTL_CUDA_OK(cudaGetLastError(), "sse_fwd launch");
  return CTL_OK;
}

extern "C" int ctl_sse_bwd(const float* pred, const float* target, int64_t n, double scale, const float* grad_out,
                           float* dpred, void* stream) {
  CTL_REQUIRE(pred && target && dpred && n > 0, CTL_ERR_INVALID, "ctl_sse_bwd: NULL pointer or n <= 0");
  const int grid = stream_grid(ceil_div(n, 4));
  if (grid < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  const float s2 = (float)(2.0 * scale);
  if (n % 4 == 0 && aligned16(pred) && aligned16(target) && aligned16(dpred))
    launch_chained(sse_bwd_kernel<4>, grid, kT, 0, st)(pred, target, grad_out, s2, n, dpred);
  else
    launch_chained(sse_bwd_kernel<1>, grid, kT, 0, st)(pred, target, grad_out, s2, n, dpred);
  CTL_CUDA_OK(cudaGetLastError(), "sse_bwd launch");
  return CTL_OK;
}

extern "C" int ctl_confusion_update(const float* logits, const int64_t* pred_labels, const int64_t* gt, int64_t N,
                                    int64_t C, int64_t HW, void* hist, void* labels_out, void* stream) {
  CTL_REQUIRE((logits != nullptr) != (pred_labels != nullptr), CTL_ERR_INVALID,
              "ctl_confusion_update: exactly one of logits / pred_labels must be given");
  CTL_REQUIRE((gt != nullptr) == (hist != nullptr) && (gt || labels_out), CTL_ERR_INVALID,
              "ctl_confusion_update: gt and hist go together; without them labels_out must be given");
  CTL_REQUIRE(N > 0 && HW > 0, CTL_ERR_INVALID, "ctl_confusion_update: bad shape N=%lld HW=%lld", (long long)N, (long long)HW);
  const int grid = stream_grid(N * HW);
  if (grid < 0) return CTL_ERR_CUDA;
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* h = (unsigned long long*)hist;
  uint8_t* lo = (uint8_t*)labels_out;
  switch ((int)C) {
    case 2: launch_chained(confusion_kernel<2>, grid, kT, 0, st)(logits, pred_labels, gt, N, HW, h, lo); break;
    case 3: launch_chained(confusion_kernel<3>, grid, kT, 0, st)(logits, pred_labels, gt, N, HW, h, lo); break;
    case 4: launch_chained(confusion_kernel<4>, grid, kT, 0, st)(logits, pred_labels, gt, N, HW, h, lo); break;
    case 8: launch_chained(confusion_kernel<8>, grid, kT, 0, st)(logits, pred_labels, gt, N, HW, h, lo); break;
    default:
      set_error("ctl_confusion_update: number of classes must be 2, 3, 4 or 8 (got %lld)", (long long)C);
      return CTL_ERR_UNSUPPORTED;
  }
  CTL_CUDA_OK(cudaGetLastError(), "confusion launch");
  return CTL_OK;
}

extern "C" int ctl_confusion_scores(const void* hist, int64_t C, double* scores_out, void* stream) {
  CTL_REQUIRE(hist && scores_out && C >= 1 && C <= 64, CTL_ERR_INVALID, "ctl_confusion_scores: NULL pointer or bad C");
  if (sm_count() < 0) return CTL_ERR_CUDA;
  launch_chained(confusion_scores_kernel, 1, 32, 0, (cudaStream_t)stream)((const unsigned long long*)hist, (int)C, scores_out);
  CTL_CUDA_OK(cudaGetLastError(), "confusion_scores launch");
  return CTL_OK;
}

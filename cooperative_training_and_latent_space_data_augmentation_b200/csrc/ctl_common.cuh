// Shared host/device helpers for the ctl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ctl_b200.h"

namespace ctl {

// ---- host side ---------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);   // records the message, returns CTL_ERR_CUDA
int sm_count();                                   // cached per device; <0 on error
int diag_flags();
// the stem's input gradient on the warp-level tensor path (conv_tc.cu / stem_dgrad_small.cuh)
bool stem_dgrad_tensor_path_handles(int64_t H, int64_t W);
int stem_dgrad_tensor_path(const void* dy, const float* x, int in_mode, float inv_temp, int N, int Cin, int H, int W,
                           const float* w, float* dx, cudaStream_t st);                                 // env CTL_DIAG_SKIP; always 0 unless the library is built with -DCTL_DIAG

// Profiling by elimination (tools/diag_conv.py) exists only in a -DCTL_DIAG build (make DIAG=1): the shipped kernels
// contain no work-skipping switch.
#ifdef CTL_DIAG
#define CTL_DIAGF(p, bit) (((p).diag & (bit)) != 0)
#else
#define CTL_DIAGF(p, bit) (false)
#endif

#define CTL_CUDA_OK(expr, what)                         \
  do {                                                  \
    cudaError_t e__ = (expr);                           \
    if (e__ != cudaSuccess) return ::ctl::cuda_fail(e__, what); \
  } while (0)

#define CTL_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::ctl::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

// ---- chained launches (programmatic dependent launch) -----------------------------------------------------------------
// The cooperative step is ~1900 short, strictly ordered kernels.  Every kernel of this library except the masking
// chain (masking.cu has its own, finer-grained use of the mechanism) is launched with programmatic stream
// serialization and begins with pdl_entry(): it lets the NEXT launch be set up while it runs, then waits until the
// PREVIOUS grid has completed and its memory is visible.  No kernel touches global memory before that wait, so the
// results are those of ordinary in-order launches; what disappears is the launch latency between dependent kernels
// (also inside a captured CUDA graph, where the attribute becomes a programmatic dependency edge).  CTL_PDL_ALL=0
// switches back to ordinary launches.
bool pdl_chain_enabled();

template <typename... KArgs>
struct ChainedLaunch {
  void (*kernel)(KArgs...);
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
  template <typename... Args>
  void operator()(Args&&... args) const {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_chain_enabled() ? 1 : 0;
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // the call sites check cudaGetLastError()
  }
};
template <typename... KArgs>
inline ChainedLaunch<KArgs...> launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st) {
  return ChainedLaunch<KArgs...>{kernel, grid, block, smem, st};
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// first statement of every chained kernel (see launch_chained)
__device__ __forceinline__ void pdl_entry() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- device side: 16-byte streaming access -----------------------------------------------------
template <typename T>
struct Elems16 { static constexpr int value = 16 / sizeof(T); };

// load VEC consecutive elements as floats.  VEC == Elems16<T> -> one 128-bit load; VEC == 1 -> scalar.
template <typename T, int VEC>
__device__ __forceinline__ void load_as_float(const T* __restrict__ p, float (&out)[VEC]);

template <>
__device__ __forceinline__ void load_as_float<float, 4>(const float* __restrict__ p, float (&out)[4]) {
  const float4 v = __ldcs(reinterpret_cast<const float4*>(p));
  out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load_as_float<float, 1>(const float* __restrict__ p, float (&out)[1]) {
  out[0] = __ldcs(p);
}
template <>
__device__ __forceinline__ void load_as_float<__nv_bfloat16, 8>(const __nv_bfloat16* __restrict__ p,
                                                                float (&out)[8]) {
  const uint4 v = __ldcs(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(w[i] << 16);
    out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <>
__device__ __forceinline__ void load_as_float<__nv_bfloat16, 1>(const __nv_bfloat16* __restrict__ p,
                                                                float (&out)[1]) {
  out[0] = __bfloat162float(*p);
}

// store VEC floats as VEC elements of T (round-to-nearest-even for bf16)
template <typename T, int VEC>
__device__ __forceinline__ void store_from_float(T* __restrict__ p, const float (&in)[VEC]);

template <>
__device__ __forceinline__ void store_from_float<float, 4>(float* __restrict__ p, const float (&in)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(in[0], in[1], in[2], in[3]);
}
template <>
__device__ __forceinline__ void store_from_float<float, 8>(float* __restrict__ p, const float (&in)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(in[0], in[1], in[2], in[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(in[4], in[5], in[6], in[7]);
}
template <>
__device__ __forceinline__ void store_from_float<float, 1>(float* __restrict__ p, const float (&in)[1]) {
  *p = in[0];
}
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16, 8>(__nv_bfloat16* __restrict__ p,
                                                                   const float (&in)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16, 4>(__nv_bfloat16* __restrict__ p,
                                                                   const float (&in)[4]) {
  uint32_t w[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(in[2 * i], in[2 * i + 1]);
    w[i] = *reinterpret_cast<const uint32_t*>(&h);
  }
  *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
}
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16, 1>(__nv_bfloat16* __restrict__ p,
                                                                   const float (&in)[1]) {
  *p = __float2bfloat16_rn(in[0]);
}

// value a float takes after a round trip through T (used for the dropout "masked == z" compare)
template <typename T> __device__ __forceinline__ float round_through(float v);
template <> __device__ __forceinline__ float round_through<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_through<__nv_bfloat16>(float v) {
  return __bfloat162float(__float2bfloat16_rn(v));
}

// lanes-of-a-row group mask for sub-warp shuffles
template <int L>
__device__ __forceinline__ unsigned group_mask() {
  if constexpr (L >= 32) {
    return 0xffffffffu;
  } else {
    const unsigned lane = threadIdx.x & 31u;
    return ((1u << L) - 1u) << (lane & ~(unsigned)(L - 1));
  }
}

}  // namespace ctl

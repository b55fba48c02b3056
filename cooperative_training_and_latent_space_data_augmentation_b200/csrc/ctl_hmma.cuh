// Warp-level tensor path helpers shared by K3s (conv_small.cuh) and K3ws (wgrad_small.cuh): ldmatrix, mma.sync
// m16n8k16 bf16 -> fp32, bulk tensor stores, multiply-shift division.  Included inside ctl's anonymous namespace.
#pragma once

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

__device__ __forceinline__ void hmma_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// first product of an accumulator chain: C = the per-channel shift of this thread's two columns (no zeroing of the
// destination registers, no shift arithmetic in the epilogue)
__device__ __forceinline__ void hmma_16816_first(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2], float c0,
                                                 float c1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%10,%11};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "f"(c0), "f"(c1));
}

// 4-D tiled store shared -> global (elements outside the tensor are not written), bulk-group completion
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(m), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// n / d for 0 <= n < 2^31 and 1 <= d < 2^13 with magic = ceil(2^44 / d) (exact: n * (magic*d - 2^44) < 2^44): one wide
// multiply and a shift instead of the ~40-instruction generic division, twice per tile and warp
__device__ __forceinline__ int small_div(int n, uint64_t magic) { return (int)(((uint64_t)(uint32_t)n * magic) >> 44); }
inline uint64_t small_div_magic(int d) { return (((uint64_t)1 << 44) + (uint64_t)d - 1) / (uint64_t)d; }

__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }


// transposed variant: thread (g, t) of matrix i receives M_i[2t][g], M_i[2t+1][g]
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// Micro-benchmark: issue rate of the legacy warp-level tensor path (mma.sync m16n8k16 bf16 -> fp32) on sm_100a.
// One CTA per SM, W warps per CTA, each warp runs ACC independent accumulator chains.  Prints HMMA per clock per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate hmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

// same, but like a real kernel: 18 distinct B fragments (36 registers), 3 A fragments, 6 accumulators
__global__ void hmma_mixed_kernel(float* out, int iters, unsigned a0, unsigned b0, long long* cycles) {
  float d[6][4];
#pragma unroll
  for (int i = 0; i < 6; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  unsigned a[3][4], b[18][2];
#pragma unroll
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) a[i][j] = a0 + i * 4 + j + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 18; ++i) for (int j = 0; j < 2; ++j) b[i][j] = b0 + i * 2 + j + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int jn = 0; jn < 2; ++jn)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(d[r * 2 + jn][0]), "+f"(d[r * 2 + jn][1]), "+f"(d[r * 2 + jn][2]), "+f"(d[r * 2 + jn][3])
                       : "r"(a[s][0]), "r"(a[s][1]), "r"(a[s][2]), "r"(a[s][3]), "r"(b[(r * 3 + s) * 2 + jn][0]), "r"(b[(r * 3 + s) * 2 + jn][1]));
  }
  __syncthreads();
  const long long t1 = clock64();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) sum += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int ACC>
__global__ void hmma_kernel(float* out, int iters, unsigned a0, unsigned b0, long long* cycles) {
  float d[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  unsigned a[4] = {a0, a0 + 1, a0 + 2, a0 + 3}, b[2] = {b0, b0 + 1};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    hmma_kernel<8><<<148, warps * 32>>>(out, 100, 0, 0, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    hmma_kernel<8><<<148, warps * 32>>>(out, iters, 0, 0, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double hmma = (double)iters * 8 * warps;
    printf("warps %2d: %.3f HMMA/clk/SM  (%.1f cyc per HMMA per sub-partition)  %.1f TFLOP/s dense bf16 chip-wide (%.2f ms)\n", warps,
           hmma / c, c / (hmma / 4), hmma * 148 * 4096 / (ms * 1e-3) / 1e12, ms);
  }
  for (int warps : {4, 8, 16}) {
    hmma_mixed_kernel<<<148, warps * 32>>>(out, 100, 0, 0, cyc);
    hmma_mixed_kernel<<<148, warps * 32>>>(out, iters, 0, 0, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double hmma = (double)iters * 18 * warps;
    printf("mixed operands, warps %2d: %.3f HMMA/clk/SM  (%.1f cyc per HMMA per sub-partition)\n", warps, hmma / c, c / (hmma / 4));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

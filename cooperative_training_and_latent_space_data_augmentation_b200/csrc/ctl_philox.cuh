// Philox4x32-10 (Salmon et al., SC'11) -- counter-based RNG of the native random mode.
// Layout (shared with oracle/masking_oracle.py::philox_u32): counter = (lo(i>>2), hi(i>>2),
// lo(offset), hi(offset)), key = (lo(seed), hi(seed)), element i takes output word i & 3.
#pragma once
#include <stdint.h>

namespace ctl {

struct PhiloxKey { uint64_t seed; uint64_t offset; };

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)c[0] * 0xD2511F53ull;
    const uint64_t p1 = (uint64_t)c[2] * 0xCD9E8D57ull;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

__host__ __device__ __forceinline__ uint32_t philox_u32(PhiloxKey key, uint64_t index) {
  const uint64_t blk = index >> 2;
  uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)key.offset, (uint32_t)(key.offset >> 32)};
  philox4x32_10(c, (uint32_t)key.seed, (uint32_t)(key.seed >> 32));
  const uint32_t w = (uint32_t)index & 3u;   // no dynamic indexing: keeps c[] in registers
  return w == 0 ? c[0] : w == 1 ? c[1] : w == 2 ? c[2] : c[3];
}

// U[0,1) with 24 random bits
__host__ __device__ __forceinline__ float philox_uniform(PhiloxKey key, uint64_t index) {
  return (float)(philox_u32(key, index) >> 8) * (1.0f / 16777216.0f);
}

}  // namespace ctl

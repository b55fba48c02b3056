// Input gradient of the stem (MyEncoder.inc[0], 3x3 pad 1, Cin = 1 / 4 -> 16) on the K3s pattern: the gradient w.r.t. the
// STN / FTN input that flows back through the 16-channel dy tensor of the stem convolution.
//
//   d_in[p][ci] = sum_{r,s,co} w[co][ci][r][s] * dy[p - (r-1, s-1)][co]           (a 3x3 convolution of dy, K = co = 16)
//   in_mode 1 chains through softmax(x / T):  dx = s * (d_in - sum_j s_j d_in_j) / T
//
// The CUDA-core kernel it replaces (stem_dgrad_kernel, c8_bwd.cu) re-read every dy pixel nine times and spent 576 FMAs
// per pixel: 178 us at batch 64 @224^2.  Here dy tiles arrive by TMA exactly like K3s's input, the A fragments are
// ldmatrix loads of the tile, and the fp32 weights enter as a bf16 (hi | lo) pair of B fragments (N = ci padded to 8, two
// mma.sync per tap: ~16 weight mantissa bits, as in the forward stem), 18 mma.sync per 16 pixels.  The finished row is
// chained through the softmax in registers (the four channels of a pixel sit in two neighbouring lanes) and stored planar
// fp32.  Included by conv_tc.cu inside ctl's anonymous namespace (shares the activation tensor map and K3s's constants).

struct StemDgradParams {
  const float* x;        // forward input, planar fp32 [N, CIN, H, W] (in_mode 1), else unused
  const float* w;        // fp32 [16][CIN][3][3]
  float* dx;             // planar fp32 [N, CIN, H, W]
  int N, H, W, in_mode;
  float inv_temp;
  int tiles_x, tiles_y;
  int64_t num_tiles;
  uint64_t magic_img, magic_x;
};

template <int CIN>
__global__ void __launch_bounds__(kSmThreads, 2)
stem_dgrad_small_kernel(const __grid_constant__ CUtensorMap tmap, const StemDgradParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmOffBar);
  uint64_t* empty = full + kSmStages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    for (int i = 0; i < kSmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSmWarps); }
    mbar_fence_init();
  }
  __syncthreads();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = (int)p.num_tiles;
  auto issue_load = [&](int t, int stage) {
    const int img = small_div(t, p.magic_img);
    const int rem = t - img * tiles_per_img;
    const int ty = small_div(rem, p.magic_x), tx = rem - ty * p.tiles_x;
    mbar_arrive_expect_tx(&full[stage], kSmStageBytes);
    tma_load_4d(smem + stage * kSmStageBytes, &tmap, &full[stage], (tx * kSmTile - 1) * 2, ty * kSmTile - 1, 0, img);
  };
  if (warp == 0 && elect_one()) {
#pragma unroll
    for (int k = 0; k < kSmStages - 1; ++k)
      if (blockIdx.x + k * gridDim.x < num_tiles) issue_load(blockIdx.x + k * gridDim.x, k);
  }
  const int g = lane >> 2, tq = lane & 3;
  const int sx = warp & 1, band = warp >> 1;
  // B fragments: B[k = co][n = ci] of halo tap (r', s') is w[co][ci][2 - r'][2 - s'] (the correlation of dy with the
  // flipped filter); b0 = co 2tq, 2tq+1, b1 = co 2tq+8, 2tq+9, column n = g (zero for g >= CIN); hi and lo halves
  uint32_t wb[9][2][2];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int wt = 8 - tap;                                  // (2 - r') * 3 + (2 - s')
    float v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int co = 2 * tq + (q & 1) + 8 * (q >> 1);
      v[q] = g < CIN ? __ldg(p.w + (co * CIN + g) * 9 + wt) : 0.0f;
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {                         // q pairs (0,1) -> b0, (2,3) -> b1
      const float h0 = __bfloat162float(__float2bfloat16_rn(v[2 * hf])), h1 = __bfloat162float(__float2bfloat16_rn(v[2 * hf + 1]));
      const __nv_bfloat162 hi = __floats2bfloat162_rn(h0, h1), lo = __floats2bfloat162_rn(v[2 * hf] - h0, v[2 * hf + 1] - h1);
      wb[tap][0][hf] = *reinterpret_cast<const uint32_t*>(&hi);
      wb[tap][1][hf] = *reinterpret_cast<const uint32_t*>(&lo);
    }
  }
  const int mat = lane >> 3;
  const uint32_t lane_off = (uint32_t)((mat >> 1) * kSmChunk +
                                       ((band * kSmBandRows) * kSmHalo + sx * 16 + (lane & 7) + 8 * (mat & 1)) * 16);
  const uint32_t smem_base = smem_u32(smem);
  const int64_t HW = (int64_t)p.H * p.W;
  const bool owner = 2 * tq < CIN;                           // this lane's channel pair (2tq, 2tq+1) exists

  int stage = 0, prev_stage = kSmStages - 1;
  uint32_t phase = 0, prev_phase = 1;
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    if (warp == 0) {
      const int t_ahead = t + (kSmStages - 1) * gridDim.x;
      if (t_ahead < num_tiles && elect_one()) {
        if (t != (int)blockIdx.x) mbar_wait(&empty[prev_stage], prev_phase);
        issue_load(t_ahead, prev_stage);
      }
      __syncwarp();
    }
    const int img = small_div(t, p.magic_img);
    const int rem = t - img * tiles_per_img;
    const int ty = small_div(rem, p.magic_x), tx = rem - ty * p.tiles_x;
    const int y0 = ty * kSmTile + band * kSmBandRows;
    const int x0 = tx * kSmTile + sx * 16 + g;               // this thread's pixels: x0 and x0 + 8
    mbar_wait(&full[stage], phase);
    const uint32_t a_base = smem_base + (uint32_t)(stage * kSmStageBytes) + lane_off;
    float acc[3][4];
    float xin[3][4];                                         // forward input of output row j (in_mode 1), two rows ahead of use
#pragma unroll
    for (int i = 0; i < kSmBandRows + 2; ++i) {
      uint32_t a[3][4];
#pragma unroll
      for (int s = 0; s < 3; ++s) ldmatrix_x4(a[s], a_base + (uint32_t)((i * kSmHalo + s) * 16));
      if constexpr (CIN == 4) {
        if (p.in_mode == 1 && i < kSmBandRows) {
          // this lane's channel pair (2tq, 2tq+1) of pixels x0 / x0 + 8 in row y0 + i: consumed when that row completes
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int xx = x0 + 8 * hp;
            const bool ok = y0 + i < p.H && xx < p.W && owner;
            const int64_t q = (int64_t)(y0 + i) * p.W + xx;
            xin[i % 3][2 * hp] = ok ? __ldg(p.x + ((int64_t)img * CIN + 2 * tq) * HW + q) : 0.0f;
            xin[i % 3][2 * hp + 1] = ok ? __ldg(p.x + ((int64_t)img * CIN + 2 * tq + 1) * HW + q) : 0.0f;
          }
        }
      }
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const int j = i - r;
          if (j >= 0 && j < kSmBandRows) {
            if (r == 0 && s == 0) hmma_16816_first(acc[j % 3], a[s], wb[r * 3 + s][0], 0.0f, 0.0f);
            else hmma_16816(acc[j % 3], a[s], wb[r * 3 + s][0]);
            hmma_16816(acc[j % 3], a[s], wb[r * 3 + s][1]);
          }
        }
      if (i == kSmBandRows + 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      if (i >= 2) {
        const int j = i - 2, y = y0 + j;
        // d[hp][e]: d_in of pixel x0 + 8*hp, channel 2tq + e
        float d[2][2] = {{acc[j % 3][0], acc[j % 3][1]}, {acc[j % 3][2], acc[j % 3][3]}};
        if (p.in_mode == 1) {
          if constexpr (CIN == 4) {
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
              float v0 = xin[j % 3][2 * hp], v1 = xin[j % 3][2 * hp + 1];
              // the other channel pair of this pixel lives in the neighbouring lane (tq ^ 1); lanes tq >= 2 idle along
              const float o0 = __shfl_xor_sync(0xffffffffu, v0, 1), o1 = __shfl_xor_sync(0xffffffffu, v1, 1);
              const float mx = fmaxf(fmaxf(v0, v1), fmaxf(o0, o1));
              v0 = __expf((v0 - mx) * p.inv_temp); v1 = __expf((v1 - mx) * p.inv_temp);
              const float e0 = __expf((o0 - mx) * p.inv_temp), e1 = __expf((o1 - mx) * p.inv_temp);
              const float inv = 1.0f / (v0 + v1 + e0 + e1);
              v0 *= inv; v1 *= inv;
              float dot = fmaf(v0, d[hp][0], v1 * d[hp][1]);
              dot += __shfl_xor_sync(0xffffffffu, dot, 1);
              d[hp][0] = v0 * (d[hp][0] - dot) * p.inv_temp;
              d[hp][1] = v1 * (d[hp][1] - dot) * p.inv_temp;
            }
          } else {
            // softmax over a single channel is the constant 1: its input gradient is zero
            d[0][0] = 0.0f; d[1][0] = 0.0f;
          }
        }
        if (y < p.H && owner) {
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            const int xx = x0 + 8 * hp;
            if (xx < p.W) {
              const int64_t q = (int64_t)y * p.W + xx;
              p.dx[((int64_t)img * CIN + 2 * tq) * HW + q] = d[hp][0];
              if (CIN > 1) p.dx[((int64_t)img * CIN + 2 * tq + 1) * HW + q] = d[hp][1];
            }
          }
        }
      }
    }
    prev_stage = stage;
    prev_phase = phase;
    if (++stage == kSmStages) { stage = 0; phase ^= 1; }
  }
}

template <int CIN>
int launch_stem_dgrad_small(const void* dy, const StemDgradParams& p0, cudaStream_t st) {
  StemDgradParams p = p0;
  p.tiles_x = (int)ceil_div(p.W, kSmTile);
  p.tiles_y = (int)ceil_div(p.H, kSmTile);
  p.num_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  p.magic_img = small_div_magic(p.tiles_x * p.tiles_y);
  p.magic_x = small_div_magic(p.tiles_x);
  CUtensorMap tmap;
  if (int rc = make_act_tmap(&tmap, dy, p.N, p.H, p.W, 16, kSmHalo, kSmHalo)) return rc;
  auto kern = stem_dgrad_small_kernel<CIN>;
  static int resident = 0;
  if (resident == 0) {
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes), "stem_dgrad_small smem attribute");
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                "stem_dgrad_small carve-out attribute");
    CTL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kSmThreads, kSmSmemBytes), "stem_dgrad_small occupancy");
    if (resident < 1) resident = 1;
  }
  const int ctas = (int)std::min<int64_t>(p.num_tiles, (int64_t)std::min(resident, 2) * sm_count());
  launch_chained(kern, (unsigned)ctas, kSmThreads, kSmSmemBytes, st)(tmap, p);
  CTL_CUDA_OK(cudaGetLastError(), "stem_dgrad_small launch");
  return CTL_OK;
}

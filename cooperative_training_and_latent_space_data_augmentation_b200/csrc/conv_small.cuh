// K3s: 3x3 stride-1 pad-1 convolution, 16 -> 16 channels, on the warp-level tensor path (mma.sync m16n8k16 bf16 -> fp32).
// Included by conv_tc.cu (shares ConvParams, the activation tensor map and the epilogue semantics of K3).
//
// Why this layer class leaves tcgen05 (measured on a B200, batch 64 @224^2): with N = Cout = 16 a
// tcgen05.mma (M=128, N=16, K=16) does 8 cycles of math but takes ~52-57 cycles to issue in shared-memory-operand mode --
// the 4 KB pixel operand is re-read from shared memory for each of the nine taps -- so the layer is MMA-issue bound at
// ~470 cycles per 128 pixels (54 us) against an HBM floor of 360 (32 us).  Packing taps into N trades that for 3x the
// TMEM read traffic in the epilogue (measured: slower).  The warp-level path keeps the pixel operand in REGISTERS:
// ldmatrix once per input row and horizontal tap, reused for the three output rows that row contributes to; its rate is
// N-independent (measured tools/micro/hmma_rate.cu: 8 cycles per m16n8k16 per SM sub-partition = 1024 MAC/clk/SM), which
// puts this layer at 288 tensor cycles per 128 pixels -- below the HBM floor.  For Cout >= 32 the arithmetic turns
// around (cost ~ Cin*Cout here, ~ Cin on tcgen05) and K3 stays the faster kernel.
//
// One CTA = 8 warps, two CTAs per SM, persistent over 32 x 32 pixel tiles of one image.
// One elected lane of warp 0 keeps a ring of kSmStages halo tiles [2 planes][34 rows][34 px][8 ch] (the C8 layout as it lies in HBM, one
// 4-D box per tile, zero-filled outside the image).  Compute warp (strip sx = warp & 1, band = warp >> 1) owns 16 pixel
// columns x 8 output rows: it walks the band's 10 input rows, loads the three horizontally shifted 16 px x 16 ch
// A fragments of a row with ldmatrix.x4 (conflict-free: 8 consecutive pixels = 128 contiguous bytes) and issues
// 3 (r) x 3 (s) x 2 (n8) mma.sync into the rolling accumulators of output rows i, i-1, i-2; the row that completes is
// finished in registers -- y = act(acc*scale + shift [+ res*rs + rb]), bf16 pack, 4-byte stores (four lanes = one 16-byte
// C8 pixel, a warp instruction = 128 contiguous bytes) -- with the optional per-channel statistics of K3's epilogue
// (STATS: sum / sum of squares of the stored outputs; BNB: sum dv | sum dv*a for the BatchNorm backward).  The 36
// B-fragment registers (nine taps x two n8 tiles) are loaded once per CTA.
// (no namespace / include guard of its own: conv_tc.cu includes it once inside ctl's anonymous namespace)

constexpr int kSmTile = 32;                          // output tile: 32 x 32 pixels
constexpr int kSmHalo = kSmTile + 2;
constexpr int kSmChunk = kSmHalo * kSmHalo * 16;     // one 8-channel plane of the halo tile
constexpr int kSmStageBytes = 2 * kSmChunk;          // 36,992 (a multiple of 128)
constexpr int kSmStages = 2;
constexpr int kSmWarps = 8;                          // compute warps: 2 strips of 16 px x 4 bands of 8 rows
constexpr int kSmBandRows = 8;
constexpr int kSmThreads = kSmWarps * 32;
constexpr int kSmOutWarpBytes = 2 * kSmBandRows * 16 * 16;   // a warp's output block [2 planes][8 rows][16 px][8 ch] = 4 KB
constexpr int kSmOffOut = kSmStages * kSmStageBytes;
constexpr int kSmOffBar = kSmOffOut + kSmWarps * kSmOutWarpBytes;
constexpr int kSmOffStat = kSmOffBar + 128;          // bars: full[2] | empty[2] | res_full[8 warps]; then [8 warps][2][16] floats
constexpr int kSmSmemBytes = kSmOffStat + kSmWarps * 32 * 4;
static_assert(kSmStageBytes % 128 == 0, "TMA destination alignment");
static_assert(2 * (kSmSmemBytes + 1024) <= 228 * 1024, "two CTAs per SM");

template <bool RES, bool STATS, bool BNB>
__global__ void __launch_bounds__(kSmThreads, 2)
conv_small_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out,
                  const __grid_constant__ CUtensorMap tmap_res, const ConvParams p) {
  pdl_entry();
  static_assert(!BNB || (RES && STATS), "the BatchNorm-backward form reads a through the residual path");
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmOffBar);      // [kSmStages] TMA -> compute
  uint64_t* empty = full + kSmStages;                                   // [kSmStages] compute -> TMA
  uint64_t* res_full = empty + kSmStages;                               // [kSmWarps] a warp's residual block landed
  float* stat_smem = reinterpret_cast<float*>(smem + kSmOffStat);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap_out);
    if (RES) tma_prefetch_desc(&tmap_res);
    for (int i = 0; i < kSmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSmWarps); }
    for (int i = 0; i < kSmWarps; ++i) mbar_init(&res_full[i], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = (int)p.num_tiles;

  // TMA producer duty (one elected lane of warp 0): tile k + kSmStages - 1 is requested at the top of iteration k into the
  // stage iteration k - 1 consumed, so two tiles are always in flight behind the one being computed
  auto issue_load = [&](int t, int stage) {
    const int img = small_div(t, p.magic_img);
    const int rem = t - img * tiles_per_img;
    const int ty = small_div(rem, p.magic_x), tx = rem - ty * p.tiles_x;
    if (CTL_DIAGF(p, 2)) { mbar_arrive(&full[stage]); return; }
    mbar_arrive_expect_tx(&full[stage], kSmStageBytes);
    tma_load_4d(smem + stage * kSmStageBytes, &tmap, &full[stage], (tx * kSmTile - 1) * 2, ty * kSmTile - 1, 0, img);
  };
  if (warp == 0 && elect_one()) {
#pragma unroll
    for (int k = 0; k < kSmStages - 1; ++k)
      if (blockIdx.x + k * gridDim.x < num_tiles) issue_load(blockIdx.x + k * gridDim.x, k);
  }

  // =================================================================== compute warps
  const int g = lane >> 2, tq = lane & 3;
  const int sx = warp & 1, band = warp >> 1;
  // B fragments of the nine taps and two n8 tiles: packed weights [tap][Cin/8 = 2][16 co][8 ci]
  uint32_t wb[9][2][2];
  {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(p.w_packed);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int jn = 0; jn < 2; ++jn) {
        wb[tap][jn][0] = __ldg(w32 + ((tap * 2 + 0) * 16 + 8 * jn + g) * 4 + tq);
        wb[tap][jn][1] = __ldg(w32 + ((tap * 2 + 1) * 16 + 8 * jn + g) * 4 + tq);
      }
  }
  // this thread's four channels: 8*jn + 2*tq + e
  // without a per-channel scale the shift is the accumulators' initial value (no epilogue arithmetic); with one
  // (inference: folded BatchNorm) the accumulators start at zero and the epilogue applies acc*scale + shift
  const bool has_scale = p.scale != nullptr;
  float sc[2][2], sh[2][2], c_init[2][2], rs[RES ? 2 : 1][2], rb[RES ? 2 : 1][2];
#pragma unroll
  for (int jn = 0; jn < 2; ++jn)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 8 * jn + 2 * tq + e;
      sc[jn][e] = has_scale ? __ldg(p.scale + c) : 1.0f;
      sh[jn][e] = p.shift ? __ldg(p.shift + c) : 0.0f;
      c_init[jn][e] = has_scale ? 0.0f : sh[jn][e];
      if (RES) {
        rs[RES ? jn : 0][e] = p.res_scale ? __ldg(p.res_scale + c) : 1.0f;
        rb[RES ? jn : 0][e] = p.res_shift ? __ldg(p.res_shift + c) : 0.0f;
      }
    }
  const bool has_res = RES && p.res != nullptr;
  const bool res_identity = p.res_scale == nullptr && p.res_shift == nullptr;
  const int act = p.act;
  const float slope = act == CTL_ACT_LRELU ? 0.2f : 1.0f;
  const float act_floor = act == CTL_ACT_RELU ? 0.0f : -INFINITY;
  const float bnb_neg = p.bnb_act == CTL_ACT_LRELU ? 0.2f : (p.bnb_act == CTL_ACT_RELU ? 0.0f : 1.0f);
  float st_s[STATS ? 2 : 1][2], st_q[STATS ? 2 : 1][2];
#pragma unroll
  for (int jn = 0; jn < (STATS ? 2 : 1); ++jn) { st_s[jn][0] = st_s[jn][1] = 0.0f; st_q[jn][0] = st_q[jn][1] = 0.0f; }

  // ldmatrix row address of this lane: matrix = lane / 8 -> (pixels 0-7 | 8-15) x (plane 0 | plane 1)
  const int mat = lane >> 3;
  const uint32_t lane_off = (uint32_t)((mat >> 1) * kSmChunk +
                                       ((band * kSmBandRows) * kSmHalo + sx * 16 + (lane & 7) + 8 * (mat & 1)) * 16);
  const uint32_t smem_base = smem_u32(smem);
  // this warp's output block in shared memory: [plane][row][16 px][4 words]; lane (g, tq) owns word g*4 + tq of a row
  uint32_t* const o_warp = reinterpret_cast<uint32_t*>(smem + kSmOffOut + warp * kSmOutWarpBytes) + lane;

  uint32_t res_phase = 0;
  int stage = 0, prev_stage = kSmStages - 1;
  uint32_t phase = 0, prev_phase = 1;                       // parity of the previous iteration's `empty` completion
  for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
    if (warp == 0) {
      const int t_ahead = t + (kSmStages - 1) * gridDim.x;
      if (t_ahead < num_tiles && elect_one()) {
        if (t != (int)blockIdx.x) mbar_wait(&empty[prev_stage], prev_phase);   // every warp is done reading that stage
        issue_load(t_ahead, prev_stage);
      }
      __syncwarp();
    }
    const int img = small_div(t, p.magic_img);
    const int rem = t - img * tiles_per_img;
    const int ty = small_div(rem, p.magic_x), tx = rem - ty * p.tiles_x;
    const int y0 = ty * kSmTile + band * kSmBandRows;
    const int x = tx * kSmTile + sx * 16 + g;                // this thread's pixels: x and x + 8
    const bool x_ok0 = x < p.W, x_ok1 = x + 8 < p.W;
    const int rows_ok = p.H - y0;                            // output rows of this band inside the image
    const bool inside = rows_ok >= kSmBandRows && tx * kSmTile + sx * 16 + 16 <= p.W;   // warp-uniform

    mbar_wait(&full[stage], phase);
    if (CTL_DIAGF(p, 8)) {                                   // profiling: no compute at all
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      prev_stage = stage; prev_phase = phase;
      if (++stage == kSmStages) { stage = 0; phase ^= 1; }
      continue;
    }
    const uint32_t a_base = smem_base + (uint32_t)(stage * kSmStageBytes) + lane_off;
    // the bulk store of the previous tile's block must have read it before this tile's first row overwrites it
    if (lane == 0) {
      if (t != (int)blockIdx.x) tma_store_wait_read();
      if (has_res) {
        // the residual pixels of the block (for BNB: the BatchNorm input a) land IN the output block: every lane
        // reads its own words and overwrites them with the result
        mbar_arrive_expect_tx(&res_full[warp], kSmOutWarpBytes);
        tma_load_4d(o_warp - lane, &tmap_res, &res_full[warp], (tx * kSmTile + sx * 16) * 2, y0, 0, img);
      }
    }
    __syncwarp();
    float acc[3][2][4];
#pragma unroll
    for (int i = 0; i < kSmBandRows + 2; ++i) {
      uint32_t a[3][4];
#pragma unroll
      for (int s = 0; s < 3; ++s) ldmatrix_x4(a[s], a_base + (uint32_t)((i * kSmHalo + s) * 16));
      // issue order: the taps that COMPLETE output row i - 2 (r = 2) and continue row i - 1 (r = 1) first, the taps that
      // open row i (r = 0) last -- the finished row's epilogue has no dependency on that last group and overlaps it
      auto taps = [&](int s, int r) {
        const int j = i - r;                                 // output row (of the band) this input row feeds through tap row r
        if (j >= 0 && j < kSmBandRows) {
#pragma unroll
          for (int jn = 0; jn < 2; ++jn) {
            if (CTL_DIAGF(p, 1)) continue;
            if (r == 0 && s == 0) hmma_16816_first(acc[j % 3][jn], a[s], wb[r * 3 + s][jn], c_init[jn][0], c_init[jn][1]);
            else hmma_16816(acc[j % 3][jn], a[s], wb[r * 3 + s][jn]);
          }
        }
      };
#pragma unroll
      for (int s = 0; s < 3; ++s) { taps(s, 0); taps(s, 1); taps(s, 2); }
      if (i == kSmBandRows + 1) {                            // the stage has been read: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      if constexpr (RES) {
        if (i == 2 && has_res) mbar_wait(&res_full[warp], res_phase);
      }
      if (i >= 2) {
        // ---- output row j = i - 2 is complete: f[jn][2*hp + e] = channel 8*jn + 2*tq + e of pixel x + 8*hp
        const int j = i - 2;
        // `masked`: the block crosses the image border -- pixels outside must not enter the statistics (the bulk store
        // clips them by itself); interior blocks skip the per-value selects
        auto finish_row = [&](auto masked) {
          constexpr bool kMasked = decltype(masked)::value;
          float f[2][4];
          uint32_t rw4[RES ? 4 : 1];
          if constexpr (RES) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rw4[q] = o_warp[((q >> 1) * kSmBandRows + j) * 64 + (q & 1) * 32];
          }
#pragma unroll
          for (int jn = 0; jn < 2; ++jn)
#pragma unroll
            for (int v = 0; v < 4; ++v) f[jn][v] = acc[j % 3][jn][v];
          if (has_scale) {                                   // warp-uniform
#pragma unroll
            for (int jn = 0; jn < 2; ++jn)
#pragma unroll
              for (int v = 0; v < 4; ++v) f[jn][v] = fmaf(f[jn][v], sc[jn][v & 1], sh[jn][v & 1]);
          }
          if constexpr (RES && !BNB) {
            if (res_identity) {
#pragma unroll
              for (int q = 0; q < 4; ++q) { f[q >> 1][2 * (q & 1)] += bf_lo(rw4[RES ? q : 0]); f[q >> 1][2 * (q & 1) + 1] += bf_hi(rw4[RES ? q : 0]); }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                f[q >> 1][2 * (q & 1)] += fmaf(bf_lo(rw4[RES ? q : 0]), rs[RES ? q >> 1 : 0][0], rb[RES ? q >> 1 : 0][0]);
                f[q >> 1][2 * (q & 1) + 1] += fmaf(bf_hi(rw4[RES ? q : 0]), rs[RES ? q >> 1 : 0][1], rb[RES ? q >> 1 : 0][1]);
              }
            }
          }
          if (act != CTL_ACT_NONE) {                         // warp-uniform; leaky ReLU (slope 0.2) | ReLU (slope 1, floor 0)
#pragma unroll
            for (int jn = 0; jn < 2; ++jn)
#pragma unroll
              for (int v = 0; v < 4; ++v) f[jn][v] = fmaxf(fmaxf(f[jn][v], slope * f[jn][v]), act_floor);
          }
          const bool y_ok = j < rows_ok;
#pragma unroll
          for (int jn = 0; jn < 2; ++jn)
#pragma unroll
            for (int hp = 0; hp < 2; ++hp) {
              const __nv_bfloat162 hh = __floats2bfloat162_rn(f[jn][2 * hp], f[jn][2 * hp + 1]);
              const uint32_t ow = *reinterpret_cast<const uint32_t*>(&hh);
              const bool ok = !kMasked || (y_ok && (hp == 0 ? x_ok0 : x_ok1));
              if constexpr (STATS && !BNB) {
                const float lo = ok ? bf_lo(ow) : 0.0f, hi = ok ? bf_hi(ow) : 0.0f;
                st_s[STATS ? jn : 0][0] += lo; st_q[STATS ? jn : 0][0] = fmaf(lo, lo, st_q[STATS ? jn : 0][0]);
                st_s[STATS ? jn : 0][1] += hi; st_q[STATS ? jn : 0][1] = fmaf(hi, hi, st_q[STATS ? jn : 0][1]);
              }
              if constexpr (BNB) {
                // dv = dy * act'(a*scale + shift) on the STORED (bf16) dy, as the stand-alone reduction reads it
                const uint32_t rw = rw4[RES ? 2 * jn + hp : 0];
                const float a_lo = bf_lo(rw), a_hi = bf_hi(rw);
                const float p_lo = fmaf(a_lo, rs[RES ? jn : 0][0], rb[RES ? jn : 0][0]);
                const float p_hi = fmaf(a_hi, rs[RES ? jn : 0][1], rb[RES ? jn : 0][1]);
                const float lo = ok ? bf_lo(ow) * (p_lo > 0.0f ? 1.0f : bnb_neg) : 0.0f;
                const float hi = ok ? bf_hi(ow) * (p_hi > 0.0f ? 1.0f : bnb_neg) : 0.0f;
                st_s[STATS ? jn : 0][0] += lo; st_q[STATS ? jn : 0][0] = fmaf(lo, a_lo, st_q[STATS ? jn : 0][0]);
                st_s[STATS ? jn : 0][1] += hi; st_q[STATS ? jn : 0][1] = fmaf(hi, a_hi, st_q[STATS ? jn : 0][1]);
              }
              o_warp[(jn * kSmBandRows + j) * 64 + hp * 32] = ow;
            }
        };
        if (STATS && !inside) finish_row(std::true_type{});
        else finish_row(std::false_type{});
      }
    }
    // the warp's 16 px x 8 rows x 2 planes leave as ONE bulk tensor store (clipped at the image border by the TMA unit):
    // no global store instruction, no store back-pressure in the compute warps
    fence_async_smem();
    __syncwarp();
    if (lane == 0 && !CTL_DIAGF(p, 4))
      tma_store_4d(&tmap_out, o_warp, (tx * kSmTile + sx * 16) * 2, y0, 0, img);
    prev_stage = stage;
    prev_phase = phase;
    if (++stage == kSmStages) { stage = 0; phase ^= 1; }
    res_phase ^= 1;
  }

  if (lane == 0) tma_store_wait_all();
  if constexpr (STATS) {
    // lanes with the same tq hold the same four channels: reduce over g, then over the eight warps through shared
    // memory, ONE fp64 atomic per channel statistic and CTA (as in K3's epilogue)
#pragma unroll
    for (int jn = 0; jn < 2; ++jn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s1 = st_s[jn][e], s2 = st_q[jn][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (g == 0) {
          stat_smem[warp * 32 + 8 * jn + 2 * tq + e] = s1;
          stat_smem[warp * 32 + 16 + 8 * jn + 2 * tq + e] = s2;
        }
      }
    asm volatile("bar.sync 1, 256;" ::: "memory");            // the eight compute warps
    if (threadIdx.x < 32) {
      const int which = threadIdx.x >> 4, c = threadIdx.x & 15;
      double total = 0.0;
#pragma unroll
      for (int w = 0; w < kSmWarps; ++w) total += (double)stat_smem[w * 32 + which * 16 + c];
      atomicAdd(p.stats + which * p.Cout + c, total);
    }
  }
}

// ---- stride 2 (3x3 pad 1, 16 -> 16 channels): the encoder's first down-sampling convolution.  K3 computes it at full
// resolution and keeps the even pixels (4x the MMA work, 71 us at 224^2); here the A fragment's eight row addresses are
// simply every second pixel of the halo row (ldmatrix takes one address per row).  Output tile 16 x 16 from the same
// 34 x 34 input box; a warp owns 16 output columns x 2 output rows (9 ldmatrix.x4 + 18 mma.sync per row, no row reuse
// between output rows at this stride) and its 1 KB output block leaves as one bulk tensor store.
constexpr int kSm2Tile = 16;
constexpr int kSm2Rows = 2;                                  // output rows per warp
constexpr int kSm2OutWarpBytes = 2 * kSm2Rows * 16 * 16;     // [2 planes][2 rows][16 px][8 ch]

template <bool STATS>
__global__ void __launch_bounds__(kSmThreads, 2)
conv_small_s2_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const ConvParams p) {
  pdl_entry();
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + kSmOffBar);
  uint64_t* empty = full + kSmStages;
  float* stat_smem = reinterpret_cast<float*>(smem + kSmOffStat);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap);
    tma_prefetch_desc(&tmap_out);
    for (int i = 0; i < kSmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kSmWarps); }
    mbar_fence_init();
  }
  __syncthreads();
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int num_tiles = (int)p.num_tiles;
  const int Ho = p.H >> 1, Wo = p.W >> 1;
  auto issue_load = [&](int t, int stage) {
    const int img = small_div(t, p.magic_img);
    const int rem = t - img * tiles_per_img;
    const int ty = small_div(rem, p.magic_x), tx = rem - ty * p.tiles_x;
    mbar_arrive_expect_tx(&full[stage], kSmStageBytes);
    tma_load_4d(smem + stage * kSmStageBytes, &tmap, &full[stage], (2 * tx * kSm2Tile - 1) * 2, 2 * ty * kSm2Tile - 1, 0, img);
  };
  if (warp == 0 && elect_one()) {
#pragma unroll
    for (int k = 0; k < kSmStages - 1; ++k)
      if (blockIdx.x + k * gridDim.x < num_tiles) issue_load(blockIdx.x + k * gridDim.x, k);
  }
  const int g = lane >> 2, tq = lane & 3;
  uint32_t wb[9][2][2];
  {
    const uint32_t* w32 = reinterpret_cast<const uint32_t*>(p.w_packed);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int jn = 0; jn < 2; ++jn) {
        wb[tap][jn][0] = __ldg(w32 + ((tap * 2 + 0) * 16 + 8 * jn + g) * 4 + tq);
        wb[tap][jn][1] = __ldg(w32 + ((tap * 2 + 1) * 16 + 8 * jn + g) * 4 + tq);
      }
  }
  const bool has_scale = p.scale != nullptr;
  float sc[2][2], sh[2][2], c_init[2][2];
#pragma unroll
  for (int jn = 0; jn < 2; ++jn)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 8 * jn + 2 * tq + e;
      sc[jn][e] = has_scale ? __ldg(p.scale + c) : 1.0f;
      sh[jn][e] = p.shift ? __ldg(p.shift + c) : 0.0f;
      c_init[jn][e] = has_scale ? 0.0f : sh[jn][e];
    }
  const int act = p.act;
  const float slope = act == CTL_ACT_LRELU ? 0.2f : 1.0f;
  const float act_floor = act == CTL_ACT_RELU ? 0.0f : -INFINITY;
  float st_s[STATS ? 2 : 1][2], st_q[STATS ? 2 : 1][2];
#pragma unroll
  for (int jn = 0; jn < (STATS ? 2 : 1); ++jn) { st_s[jn][0] = st_s[jn][1] = 0.0f; st_q[jn][0] = st_q[jn][1] = 0.0f; }
  // ldmatrix row address: output pixel px of the strip reads halo column 2*px + s of halo row 2*j + r
  const int mat = lane >> 3;
  const uint32_t lane_off = (uint32_t)((mat >> 1) * kSmChunk +
                                       ((2 * kSm2Rows * warp) * kSmHalo + 2 * ((lane & 7) + 8 * (mat & 1))) * 16);
  const uint32_t smem_base = smem_u32(smem);
  uint32_t* const o_warp = reinterpret_cast<uint32_t*>(smem + kSmOffOut + warp * kSm2OutWarpBytes) + lane;

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
    const int y0 = ty * kSm2Tile + kSm2Rows * warp;           // first output row of this warp
    const int x = tx * kSm2Tile + g;
    const bool x_ok0 = x < Wo, x_ok1 = x + 8 < Wo;
    const int rows_ok = Ho - y0;
    mbar_wait(&full[stage], phase);
    const uint32_t a_base = smem_base + (uint32_t)(stage * kSmStageBytes) + lane_off;
    if (t != (int)blockIdx.x) {
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
    }
#pragma unroll
    for (int j = 0; j < kSm2Rows; ++j) {
      float acc[2][4];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        uint32_t a[3][4];
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3) ldmatrix_x4(a[s3], a_base + (uint32_t)(((2 * j + r) * kSmHalo + s3) * 16));
#pragma unroll
        for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
          for (int jn = 0; jn < 2; ++jn) {
            if (r == 0 && s3 == 0) hmma_16816_first(acc[jn], a[s3], wb[r * 3 + s3][jn], c_init[jn][0], c_init[jn][1]);
            else hmma_16816(acc[jn], a[s3], wb[r * 3 + s3][jn]);
          }
      }
      if (j == kSm2Rows - 1) {                               // the stage has been read
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stage]);
      }
      float f[2][4];
#pragma unroll
      for (int jn = 0; jn < 2; ++jn)
#pragma unroll
        for (int v = 0; v < 4; ++v) f[jn][v] = has_scale ? fmaf(acc[jn][v], sc[jn][v & 1], sh[jn][v & 1]) : acc[jn][v];
      if (act != CTL_ACT_NONE) {
#pragma unroll
        for (int jn = 0; jn < 2; ++jn)
#pragma unroll
          for (int v = 0; v < 4; ++v) f[jn][v] = fmaxf(fmaxf(f[jn][v], slope * f[jn][v]), act_floor);
      }
      const bool y_ok = j < rows_ok;
#pragma unroll
      for (int jn = 0; jn < 2; ++jn)
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const __nv_bfloat162 hh = __floats2bfloat162_rn(f[jn][2 * hp], f[jn][2 * hp + 1]);
          const uint32_t ow = *reinterpret_cast<const uint32_t*>(&hh);
          if constexpr (STATS) {
            const bool ok = y_ok && (hp == 0 ? x_ok0 : x_ok1);
            const float lo = ok ? bf_lo(ow) : 0.0f, hi = ok ? bf_hi(ow) : 0.0f;
            st_s[STATS ? jn : 0][0] += lo; st_q[STATS ? jn : 0][0] = fmaf(lo, lo, st_q[STATS ? jn : 0][0]);
            st_s[STATS ? jn : 0][1] += hi; st_q[STATS ? jn : 0][1] = fmaf(hi, hi, st_q[STATS ? jn : 0][1]);
          }
          o_warp[(jn * kSm2Rows + j) * 64 + hp * 32] = ow;
        }
    }
    fence_async_smem();
    __syncwarp();
    if (lane == 0) tma_store_4d(&tmap_out, o_warp, (tx * kSm2Tile) * 2, y0, 0, img);
    prev_stage = stage;
    prev_phase = phase;
    if (++stage == kSmStages) { stage = 0; phase ^= 1; }
  }
  if (lane == 0) tma_store_wait_all();
  if constexpr (STATS) {
#pragma unroll
    for (int jn = 0; jn < 2; ++jn)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float s1 = st_s[jn][e], s2 = st_q[jn][e];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (g == 0) {
          stat_smem[warp * 32 + 8 * jn + 2 * tq + e] = s1;
          stat_smem[warp * 32 + 16 + 8 * jn + 2 * tq + e] = s2;
        }
      }
    __syncthreads();
    if (threadIdx.x < 32) {
      const int which = threadIdx.x >> 4, c = threadIdx.x & 15;
      double total = 0.0;
#pragma unroll
      for (int w = 0; w < kSmWarps; ++w) total += (double)stat_smem[w * 32 + which * 16 + c];
      atomicAdd(p.stats + which * p.Cout + c, total);
    }
  }
}

// true when this layer / epilogue combination runs on K3s
inline bool conv_small_handles(int Cin, int Cout, int taps, const ConvParams& p) {
  if (!(Cin == 16 && Cout == 16 && taps == 9 && !p.up2x && p.sal == nullptr && p.act != CTL_ACT_SIGMOID)) return false;
  if (p.subsample == 2)                                      // forward only: no residual / BatchNorm-backward form
    return p.res == nullptr && p.bnb_act == 0 && ceil_div(p.H / 2, kSm2Tile) * ceil_div(p.W / 2, kSm2Tile) < 8192;
  return p.subsample == 1 && ceil_div(p.H, kSmTile) * ceil_div(p.W, kSmTile) < 8192 &&
         !(p.stats != nullptr && p.res != nullptr && p.bnb_act == 0);
}

int make_act_tmap(CUtensorMap* m, const void* x, int N, int H, int W, int C, int halo_w, int halo_h);

template <bool RES, bool STATS, bool BNB>
int launch_conv_small_variant(const CUtensorMap& tmap, const CUtensorMap& tmap_out, const CUtensorMap& tmap_res,
                              const ConvParams& p, cudaStream_t st) {
  auto kern = conv_small_kernel<RES, STATS, BNB>;
  static int resident = 0;                                   // CTAs of this variant that fit on one SM
  if (resident == 0) {
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes), "conv_small smem attribute");
    // two CTAs per SM need (almost) the whole 228 KB as shared memory: ask for the largest carve-out
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                "conv_small carve-out attribute");
    CTL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kSmThreads, kSmSmemBytes), "conv_small occupancy");
    if (getenv("CTL_VERBOSE")) fprintf(stderr, "[ctl] conv_small<%d,%d,%d>: %d CTAs per SM\n", (int)RES, (int)STATS, (int)BNB, resident);
    if (resident < 1) resident = 1;
  }
  const int ctas = (int)std::min<int64_t>(p.num_tiles, (int64_t)std::min(resident, 2) * sm_count());
  launch_chained(kern, (unsigned)ctas, kSmThreads, kSmSmemBytes, st)(tmap, tmap_out, tmap_res, p);
  CTL_CUDA_OK(cudaGetLastError(), "conv_small launch");
  return CTL_OK;
}

// CTL_CONV_SMALL=0 (diagnostic): keep this layer class on the tcgen05 kernel
inline bool conv_small_enabled() {
  static const bool on = [] { const char* e = getenv("CTL_CONV_SMALL"); return !(e && e[0] == '0'); }();
  return on;
}

template <bool STATS>
int launch_conv_small_s2_variant(const CUtensorMap& tmap, const CUtensorMap& tmap_out, const ConvParams& p, cudaStream_t st) {
  auto kern = conv_small_s2_kernel<STATS>;
  static int resident = 0;
  if (resident == 0) {
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmSmemBytes), "conv_small_s2 smem attribute");
    CTL_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared),
                "conv_small_s2 carve-out attribute");
    CTL_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, kSmThreads, kSmSmemBytes), "conv_small_s2 occupancy");
    if (resident < 1) resident = 1;
  }
  const int ctas = (int)std::min<int64_t>(p.num_tiles, (int64_t)std::min(resident, 2) * sm_count());
  launch_chained(kern, (unsigned)ctas, kSmThreads, kSmSmemBytes, st)(tmap, tmap_out, p);
  CTL_CUDA_OK(cudaGetLastError(), "conv_small_s2 launch");
  return CTL_OK;
}

int launch_conv_small(const void* x, const ConvParams& p0, cudaStream_t st) {
  ConvParams p = p0;
  p.diag = diag_flags();
  if (p.subsample == 2) {
    p.tiles_x = (int)ceil_div(p.W / 2, kSm2Tile);
    p.tiles_y = (int)ceil_div(p.H / 2, kSm2Tile);
    p.num_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
    p.magic_img = small_div_magic(p.tiles_x * p.tiles_y);
    p.magic_x = small_div_magic(p.tiles_x);
    CUtensorMap tmap, tmap_out;
    if (int rc = make_act_tmap(&tmap, x, p.N, p.H, p.W, 16, kSmHalo, kSmHalo)) return rc;
    if (int rc = make_act_tmap(&tmap_out, p.out, p.N, p.H / 2, p.W / 2, 16, 16, kSm2Rows)) return rc;
    return p.stats != nullptr ? launch_conv_small_s2_variant<true>(tmap, tmap_out, p, st)
                              : launch_conv_small_s2_variant<false>(tmap, tmap_out, p, st);
  }
  p.tiles_x = (int)ceil_div(p.W, kSmTile);
  p.tiles_y = (int)ceil_div(p.H, kSmTile);
  p.num_tiles = (int64_t)p.N * p.tiles_x * p.tiles_y;
  p.magic_img = small_div_magic(p.tiles_x * p.tiles_y);
  p.magic_x = small_div_magic(p.tiles_x);
  CUtensorMap tmap;
  CUtensorMap tmap_out;                                      // a warp's output block: 16 px x 8 rows x 2 planes
  if (int rc = make_act_tmap(&tmap, x, p.N, p.H, p.W, 16, kSmHalo, kSmHalo)) return rc;
  if (int rc = make_act_tmap(&tmap_out, p.out, p.N, p.H, p.W, 16, 16, kSmBandRows)) return rc;
  CUtensorMap tmap_res = tmap_out;                           // the residual shares the output's geometry
  if (p.res != nullptr)
    if (int rc = make_act_tmap(&tmap_res, p.res, p.N, p.H, p.W, 16, 16, kSmBandRows)) return rc;
  if (p.bnb_act != 0) return launch_conv_small_variant<true, true, true>(tmap, tmap_out, tmap_res, p, st);
  if (p.res != nullptr) return launch_conv_small_variant<true, false, false>(tmap, tmap_out, tmap_res, p, st);
  if (p.stats != nullptr) return launch_conv_small_variant<false, true, false>(tmap, tmap_out, tmap_res, p, st);
  return launch_conv_small_variant<false, false, false>(tmap, tmap_out, tmap_res, p, st);
}

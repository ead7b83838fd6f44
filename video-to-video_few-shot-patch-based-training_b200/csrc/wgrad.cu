// wgrad.cu — convolution weight gradient on tcgen05 (sm_100a).
//
//   dW[tap][ci][co] += sum_{n,y,x} X[n, ci, y+dy-pad_t, x+dx-pad_l] * dY[n, co, y, x]
//
// (autograd of nn.Conv2d in GeneratorJ; fires inside manual_backward, reference lightning_model.py:241.)
//
// GEMM view per tap: M = ci (128-row blocks), N = co, K = pixels.  Both operands come straight from the
// P8 activation layout [c/8][y][x][8]: 8 channels are contiguous, so the channel dimension is the
// "MN-major" dimension of a SWIZZLE_NONE UMMA operand (core matrix = 8 pixels x 8 channels = 128 B),
// and the pixel shift of a tap is again a 16-byte shift of the descriptor start address.
//
// Work decomposition (v2): one CTA owns a *unit* = (tap row dy, a run of <= GW taps in that row, a slice of
// NCg output channels, a 128-channel ci block) over a strided subset of 8x8-pixel tiles.  Units are sized so
// that a CTA needs <= 256 TMEM columns and ~100 KB of shared memory: TWO CTAs are co-resident per SM, which
// matters because one CTA's MMA stream cannot hide its own operand-fetch latency (profiles/r1_issue_experiments.md).
// Per tile: one TMA load of the haloed X rows (zero fill = conv padding) + one of the dY tile, 4 x ntaps MMAs
// (K = 16 pixels each) with the taps innermost so that consecutive MMAs hit different accumulators.
// The accumulators stay in TMEM over the CTA's whole pixel range and are finally added to dW with 128-bit
// fp32 vector reductions.
#include "internal.h"
#include "ptx.cuh"

namespace pbt {

constexpr int kMaxUnits = 64;
constexpr int kTileRows = 8;

struct WgradKParams {
  int n_img, H, W;
  int Cm, NC, NCg;
  int KH, KW, pad_t, pad_l;
  int splits, tiles_x, tiles_y, n_tiles;
  int BW;  // haloed X width in pixels: 8 + GW - 1
  uint32_t idesc;
  int acc_stride, tmem_cols, stages;
  uint32_t x_stage_bytes, dy_stage_bytes, x_plane_bytes;
  float* dw;
  const float* inv_scale;
  int n_units;
  signed char u_dy[kMaxUnits], u_dx0[kMaxUnits], u_ntap[kMaxUnits];
  short u_noff[kMaxUnits];
};

constexpr int kWgThreads = 192;

__global__ void __launch_bounds__(kWgThreads, 2)
wgrad_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDY, const WgradKParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sX = smem;
  uint8_t* sDY = sX + (size_t)p.stages * p.x_stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sDY + (size_t)p.stages * p.dy_stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x, unit = blockIdx.y, mb = blockIdx.z;
  const int dy = p.u_dy[unit], dx0 = p.u_dx0[unit], ntap = p.u_ntap[unit], noff = p.u_noff[unit];
  const int my_tiles = split < p.n_tiles ? (p.n_tiles - split + p.splits - 1) / p.splits : 0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tmapX);
    prefetch_tmap(&tmapDY);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // programmatic dependent launch: set-up above overlapped the previous launch's drain
  const uint32_t tmem_base = *tmem_slot;

  if (my_tiles > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t tx_bytes = (uint32_t)(16 * kTileRows * p.BW * 16) + (uint32_t)((p.NCg / 8) * kTileRows * 128);
        for (int it = 0; it < my_tiles; ++it) {
          const int tile = split + it * p.splits;
          const int n = tile / tiles_per_img;
          const int rem = tile - n * tiles_per_img;
          const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
          const int x0 = txi * 8, y0 = tyi * kTileRows;
          const int st = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], tx_bytes);
          tma_load_4d(sX + (size_t)st * p.x_stage_bytes, &tmapX, &full[st], (x0 - p.pad_l + dx0) * 8, y0 + dy - p.pad_t,
                      mb * 16, n);
          tma_load_4d(sDY + (size_t)st * p.dy_stage_bytes, &tmapDY, &full[st], x0 * 8, y0, noff / 8, n);
        }
      }
    } else if (warp == 1) {
      // whole warp runs the uniform loops; only the tcgen05 instructions are predicated on one elected lane
      const bool leader = elect_one();
      // descriptor words (MN-major, SWIZZLE_NONE): lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version<<14.
      // Only the low word changes per MMA: +1 (one pixel = 16 B) per tap, +2 tile rows per K=16 step.
      const uint32_t row16 = (uint32_t)p.BW;                                   // haloed row pitch in 16-byte units
      const uint32_t a_lo_const = (row16 & 0x3FFF) << 16;                      // LBO = next 8-pixel K group = next row
      const uint32_t a_hi = ((p.x_plane_bytes >> 4) & 0x3FFF) | (1u << 14);    // SBO = next 8-channel group = next plane
      const uint32_t b_lo_const = (128u >> 4) << 16;
      const uint32_t b_hi = ((uint32_t)(kTileRows * 128) >> 4) | (1u << 14);
      const uint32_t a_kstep = 2u * row16, b_kstep = 2u * 8u;
      const uint32_t idesc = p.idesc;
      const uint32_t acc_stride = (uint32_t)p.acc_stride;
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t x_lo = (smem_u32(sX + (size_t)st * p.x_stage_bytes) >> 4) | a_lo_const;
        const uint32_t dy_lo = (smem_u32(sDY + (size_t)st * p.dy_stage_bytes) >> 4) | b_lo_const;
        const uint32_t first = it == 0 ? 0u : 1u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < kTileRows / 2; ++k) {  // K = 16 pixels = two 8-pixel tile rows per MMA
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (dy_lo + (uint32_t)k * b_kstep);
            for (int j = 0; j < ntap; ++j) {
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (x_lo + (uint32_t)j + (uint32_t)k * a_kstep);
              umma_f16(tmem_base + (uint32_t)j * acc_stride, adesc, bdesc, idesc, k == 0 ? first : 1u);
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(acc_full);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int ci = mb * 128 + q * 32 + lane;
      const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      for (int j = 0; j < ntap; ++j) {
        float* dst = p.dw + ((long long)(dy * p.KW + dx0 + j) * p.Cm + ci) * p.NC + noff;
        for (int c0 = 0; c0 < p.NCg; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.acc_stride + c0), raw);
          tmem_ld_wait();
          if (ci < p.Cm) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {  // 128-bit vector reduction (sm_90+): 4x fewer L2 atomic requests
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i),
                           "f"(__uint_as_float(raw[i]) * inv), "f"(__uint_as_float(raw[i + 1]) * inv),
                           "f"(__uint_as_float(raw[i + 2]) * inv), "f"(__uint_as_float(raw[i + 3]) * inv)
                           : "memory");
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// Taps-in-M variant for convs with few input channels (the first conv: cin padded to 16).  With M = channels a 128-row
// MMA carries 16 useful rows.  Here M = 64 = (8 horizontal taps) x (8 channels of ONE plane): the 8-row groups of the
// MN-major A operand are 16 bytes apart (SBO = one pixel), i.e. group dx is the same channel plane shifted by dx
// pixels - overlapping operand windows are legal.  One MMA per (k-step, channel plane) replaces KW MMAs, and only the
// real channel planes are loaded.  Accumulator row m = 8*dx + ci lives in TMEM lane (m % 16) + 32*(m / 16) (the M = 64
// layout uses the first 16 lanes of every 32-lane quarter).  Unit = (tap row dy, output-channel slice).
__global__ void __launch_bounds__(kWgThreads, 2)
wgrad_taps_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapDY, const WgradKParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sX = smem;
  uint8_t* sDY = sX + (size_t)p.stages * p.x_stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(sDY + (size_t)p.stages * p.dy_stage_bytes);
  uint64_t* empty = full + p.stages;
  uint64_t* acc_full = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int split = blockIdx.x, unit = blockIdx.y;
  const int dy = p.u_dy[unit], noff = p.u_noff[unit];
  const int planes = p.Cm / 8;
  const int my_tiles = split < p.n_tiles ? (p.n_tiles - split + p.splits - 1) / p.splits : 0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
    prefetch_tmap(&tmapX);
    prefetch_tmap(&tmapDY);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();   // programmatic dependent launch: set-up above overlapped the previous launch's drain
  const uint32_t tmem_base = *tmem_slot;

  if (my_tiles > 0) {
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t tx_bytes = (uint32_t)(planes * kTileRows * p.BW * 16) + (uint32_t)((p.NCg / 8) * kTileRows * 128);
        for (int it = 0; it < my_tiles; ++it) {
          const int tile = split + it * p.splits;
          const int n = tile / tiles_per_img;
          const int rem = tile - n * tiles_per_img;
          const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
          const int x0 = txi * 8, y0 = tyi * kTileRows;
          const int st = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], tx_bytes);
          tma_load_4d(sX + (size_t)st * p.x_stage_bytes, &tmapX, &full[st], (x0 - p.pad_l) * 8, y0 + dy - p.pad_t, 0, n);
          tma_load_4d(sDY + (size_t)st * p.dy_stage_bytes, &tmapDY, &full[st], x0 * 8, y0, noff / 8, n);
        }
      }
    } else if (warp == 1) {
      const bool leader = elect_one();
      const uint32_t row16 = (uint32_t)p.BW;                                   // haloed row pitch in 16-byte units
      const uint32_t a_lo_const = (row16 & 0x3FFF) << 16;                      // LBO = next 8-pixel K group = next row
      const uint32_t a_hi = 1u | (1u << 14);                                   // SBO = next M group = next tap = +1 pixel (16 B)
      const uint32_t b_lo_const = (128u >> 4) << 16;
      const uint32_t b_hi = ((uint32_t)(kTileRows * 128) >> 4) | (1u << 14);
      const uint32_t a_kstep = 2u * row16, b_kstep = 2u * 8u;
      const uint32_t plane16 = p.x_plane_bytes >> 4;
      const uint32_t idesc = p.idesc;
      const uint32_t acc_stride = (uint32_t)p.acc_stride;
      for (int it = 0; it < my_tiles; ++it) {
        const int st = it % p.stages;
        const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t x_lo = (smem_u32(sX + (size_t)st * p.x_stage_bytes) >> 4) | a_lo_const;
        const uint32_t dy_lo = (smem_u32(sDY + (size_t)st * p.dy_stage_bytes) >> 4) | b_lo_const;
        const uint32_t first = it == 0 ? 0u : 1u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < kTileRows / 2; ++k) {
            const uint64_t bdesc = ((uint64_t)b_hi << 32) | (dy_lo + (uint32_t)k * b_kstep);
            for (int pl = 0; pl < planes; ++pl) {
              const uint64_t adesc = ((uint64_t)a_hi << 32) | (x_lo + (uint32_t)pl * plane16 + (uint32_t)k * a_kstep);
              umma_f16(tmem_base + (uint32_t)pl * acc_stride, adesc, bdesc, idesc, k == 0 ? first : 1u);
            }
          }
          umma_commit(&empty[st]);
        }
        __syncwarp();
      }
      if (leader) umma_commit(acc_full);
      __syncwarp();
    } else {
      const int q = warp & 3;
      const int dx = 2 * q + (lane >> 3), ci8 = lane & 7;   // accumulator row 8*dx + ci8 sits in lane (lane < 16) of quarter q
      const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      for (int pl = 0; pl < planes; ++pl) {
        const int ci = pl * 8 + ci8;
        for (int c0 = 0; c0 < p.NCg; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(pl * p.acc_stride + c0), raw);
          tmem_ld_wait();
          if (lane < 16 && dx < p.KW && ci < p.Cm) {
            float* dst = p.dw + ((long long)(dy * p.KW + dx) * p.Cm + ci) * p.NC + noff;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c0 + i),
                           "f"(__uint_as_float(raw[i]) * inv), "f"(__uint_as_float(raw[i + 1]) * inv),
                           "f"(__uint_as_float(raw[i + 2]) * inv), "f"(__uint_as_float(raw[i + 3]) * inv)
                           : "memory");
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_conv_wgrad(const pbt_wgrad_desc_t* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(d != nullptr, "wgrad: null descriptor");
  const pbt_act_t& x = d->x;
  const pbt_act_t& g = d->dy;
  PBT_REQUIRE(x.ptr && g.ptr && aligned16(x.ptr) && aligned16(g.ptr) && d->dw, "wgrad: null or misaligned tensors");
  PBT_REQUIRE(x.n == g.n && x.h == g.h && x.w == g.w && x.n > 0, "wgrad: x / dy shape mismatch");
  PBT_REQUIRE(x.c % 8 == 0 && x.c > 0, "wgrad: cin must be a multiple of 8");
  PBT_REQUIRE(g.c % 16 == 0 && g.c >= 16 && g.c <= 256, "wgrad: cout must be a multiple of 16 in [16,256]");
  PBT_REQUIRE(d->kh >= 1 && d->kh <= 7 && d->kw >= 1 && d->kw <= 7, "wgrad: kernel size must be in [1,7]");
  PBT_REQUIRE(d->dtype == PBT_BF16 || d->dtype == PBT_FP16, "wgrad: bad dtype");
  PBT_REQUIRE(x.img_stride % 8 == 0 && g.img_stride % 8 == 0, "wgrad: img_stride must be a multiple of 8");

  WgradKParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = x.n; p.H = x.h; p.W = x.w;
  p.Cm = x.c; p.NC = g.c;
  p.KH = d->kh; p.KW = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  if (x.c <= 32 && d->kw >= 4 && d->kw <= 8 && !(d->debug_flags & 2)) {
    // ---- taps-in-M variant (few input channels, wide kernel): see wgrad_taps_kernel
    int n_split = 1;
    while ((x.c / 8) * (int)round_up((uint32_t)(p.NC / n_split), 32) > 256 && (p.NC / n_split) % 32 == 0) n_split *= 2;
    p.NCg = p.NC / n_split;
    p.acc_stride = (int)round_up((uint32_t)p.NCg, 32);
    PBT_REQUIRE((x.c / 8) * p.acc_stride <= 512, "wgrad: unit exceeds tensor memory (512 columns)");
    p.n_units = 0;
    for (int dyi = 0; dyi < p.KH; ++dyi)
      for (int ns = 0; ns < n_split; ++ns) {
        PBT_REQUIRE(p.n_units < kMaxUnits, "wgrad: too many work units");
        p.u_dy[p.n_units] = (signed char)dyi;
        p.u_noff[p.n_units] = (short)(ns * p.NCg);
        ++p.n_units;
      }
    p.BW = 16;                                   // 8 pixels + 7 tap shifts (+1: the unused 8th tap group stays inside the row)
    p.tiles_x = ceil_div(p.W, 8);
    p.tiles_y = ceil_div(p.H, kTileRows);
    p.n_tiles = p.n_img * p.tiles_x * p.tiles_y;
    int cols = 32;
    while (cols < (x.c / 8) * p.acc_stride) cols <<= 1;
    p.tmem_cols = cols;
    p.idesc = make_idesc_f16(64, p.NCg, d->dtype == PBT_BF16 ? 1 : 0, 1, 1);
    p.x_plane_bytes = (uint32_t)(kTileRows * p.BW * 16);
    p.x_stage_bytes = round_up((uint32_t)(x.c / 8) * p.x_plane_bytes + 128, 128);   // +128: the dummy tap group reads past the last row
    p.dy_stage_bytes = (uint32_t)((p.NCg / 8) * kTileRows * 128);
    p.stages = 4;
    p.dw = d->dw;
    p.inv_scale = d->inv_scale;
    const int ovr = (d->debug_flags >> 8) & 0xff;
    int splits = ((ovr ? ovr : 4) * num_sms()) / p.n_units;   // short MMA chains, small flush: latency-bound, more CTAs help
    if (!ovr && splits > p.n_tiles / 24) splits = p.n_tiles / 24;
    if (splits < 1) splits = 1;
    if (splits > p.n_tiles) splits = p.n_tiles;
    p.splits = splits;
    const uint32_t smem_bytes = (uint32_t)p.stages * (p.x_stage_bytes + p.dy_stage_bytes) + 8u * (2 * p.stages + 1) + 16 + 128;
    CUtensorMap tx, tg;
    int rc = make_p8_tmap(&tx, x, p.BW, kTileRows, x.c / 8);
    if (rc != PBT_OK) return rc;
    rc = make_p8_tmap(&tg, g, 8, kTileRows, p.NCg / 8);
    if (rc != PBT_OK) return rc;
    PBT_CUDA_CHECK(cudaFuncSetAttribute(wgrad_taps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    dim3 grid(p.splits, p.n_units, 1);
    pbt::launch(wgrad_taps_kernel, grid, kWgThreads, smem_bytes, stream, tx, tg, p);
    PBT_CUDA_CHECK(cudaGetLastError());
    return PBT_OK;
  }
  // unit shape: <= 256 TMEM columns per CTA.  Narrow kernels split the output channels, wide ones the tap row.
  int n_split = 1;
  while (p.NC / n_split > 128 && (p.NC / n_split) % 32 == 0) n_split *= 2;
  if (p.KW <= 3 && p.KW * (int)round_up((uint32_t)(p.NC / n_split), 32) > 256 && (p.NC / n_split) % 32 == 0) n_split *= 2;
  p.NCg = p.NC / n_split;
  p.acc_stride = (int)round_up((uint32_t)p.NCg, 32);
  int gw = 256 / p.acc_stride;
  if (gw > p.KW) gw = p.KW;
  if (gw < 1) gw = 1;
  const int groups_per_row = ceil_div(p.KW, gw);
  gw = ceil_div(p.KW, groups_per_row);  // balance the runs (7 taps, limit 4 -> 4 + 3)
  p.n_units = 0;
  for (int dyi = 0; dyi < p.KH; ++dyi)
    for (int gi = 0; gi < groups_per_row; ++gi)
      for (int ns = 0; ns < n_split; ++ns) {
        PBT_REQUIRE(p.n_units < kMaxUnits, "wgrad: too many work units");
        const int dx0 = gi * gw;
        p.u_dy[p.n_units] = (signed char)dyi;
        p.u_dx0[p.n_units] = (signed char)dx0;
        p.u_ntap[p.n_units] = (signed char)((dx0 + gw <= p.KW) ? gw : p.KW - dx0);
        p.u_noff[p.n_units] = (short)(ns * p.NCg);
        ++p.n_units;
      }
  p.BW = 8 + gw - 1;
  p.tiles_x = ceil_div(p.W, 8);
  p.tiles_y = ceil_div(p.H, kTileRows);
  p.n_tiles = p.n_img * p.tiles_x * p.tiles_y;
  int cols = 32;
  while (cols < gw * p.acc_stride) cols <<= 1;
  PBT_REQUIRE(cols <= 512, "wgrad: unit exceeds tensor memory (512 columns)");
  p.tmem_cols = cols;
  p.idesc = make_idesc_f16(128, p.NCg, d->dtype == PBT_BF16 ? 1 : 0, 1, 1);
  p.x_plane_bytes = (uint32_t)(kTileRows * p.BW * 16);
  p.x_stage_bytes = round_up(16 * p.x_plane_bytes, 128);
  p.dy_stage_bytes = (uint32_t)((p.NCg / 8) * kTileRows * 128);
  p.stages = 3;
  p.dw = d->dw;
  p.inv_scale = d->inv_scale;
  const int m_blocks = ceil_div(p.Cm, 128);
  // Pixel split: one wave of CTAs (two per SM), but at least ~24 tiles per CTA - every CTA ends with a flush of its
  // accumulators into dW by fp32 reductions, and on small maps those reductions, not the MMAs, set the time
  // (tools/wgrad_occ.py: 128->128 3x3 on 80 x 20x20 maps: 46 us with 4 CTAs/SM worth of splits, 27 us with one wave).
  const int ovr = (d->debug_flags >> 8) & 0xff;  // (bring-up: bits 8-15 override the CTA budget multiplier)
  int splits = ((ovr ? ovr : 2) * num_sms()) / (m_blocks * p.n_units);
  if (!ovr && splits > p.n_tiles / 24) splits = p.n_tiles / 24;
  if (splits < 1) splits = 1;
  if (splits > p.n_tiles) splits = p.n_tiles;
  p.splits = splits;
  const uint32_t smem_bytes = (uint32_t)p.stages * (p.x_stage_bytes + p.dy_stage_bytes) + 8u * (2 * p.stages + 1) + 16 + 128;
  PBT_REQUIRE(smem_bytes <= 227 * 1024, "wgrad: configuration does not fit shared memory");

  CUtensorMap tx, tg;
  int rc = make_p8_tmap(&tx, x, p.BW, kTileRows, 16);
  if (rc != PBT_OK) return rc;
  rc = make_p8_tmap(&tg, g, 8, kTileRows, p.NCg / 8);
  if (rc != PBT_OK) return rc;

  PBT_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  dim3 grid(p.splits, p.n_units, m_blocks);
  pbt::launch(wgrad_kernel, grid, kWgThreads, smem_bytes, stream, tx, tg, p);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

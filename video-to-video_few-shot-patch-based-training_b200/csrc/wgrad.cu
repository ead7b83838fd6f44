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
// One CTA owns (pixel split, tap row dy, ci block): it streams 8x16-pixel tiles (TMA, zero fill outside
// the image = the conv padding), keeps kw accumulators [128 x co] in TMEM (kw*co <= 512 columns) over its
// whole pixel range, and finally adds them to dW with fp32 reductions.
#include "internal.h"
#include "ptx.cuh"

namespace pbt {

struct WgradKParams {
  int n_img, H, W;
  int Cm, NC;
  int KH, KW, pad_t, pad_l;
  int splits, tiles_x, tiles_y, n_tiles;
  int BW;
  uint32_t idesc;
  int acc_stride, tmem_cols, stages;
  uint32_t x_stage_bytes, dy_stage_bytes, x_plane_bytes;
  float* dw;
  const float* inv_scale;
  int debug_flags;
};

constexpr int kWgThreads = 192;

__global__ void __launch_bounds__(kWgThreads, 1)
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
  const int split = blockIdx.x, dy = blockIdx.y, mb = blockIdx.z;
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
  const uint32_t tmem_base = *tmem_slot;

  if (my_tiles > 0) {
    if (warp == 0) {
      if (elect_one()) {
        for (int it = 0; it < my_tiles; ++it) {
          const int tile = split + it * p.splits;
          const int n = tile / tiles_per_img;
          const int rem = tile - n * tiles_per_img;
          const int tyi = rem / p.tiles_x, txi = rem - tyi * p.tiles_x;
          const int x0 = txi * 8, y0 = tyi * 16;
          const int st = it % p.stages;
          const uint32_t ph = (uint32_t)(it / p.stages) & 1u;
          mbar_wait(&empty[st], ph ^ 1u);
          mbar_arrive_expect_tx(&full[st], (uint32_t)(16 * 16 * p.BW * 16) + (uint32_t)((p.NC / 8) * 2048));
          tma_load_4d(sX + (size_t)st * p.x_stage_bytes, &tmapX, &full[st], (x0 - p.pad_l) * 8, y0 + dy - p.pad_t, mb * 16, n);
          tma_load_4d(sDY + (size_t)st * p.dy_stage_bytes, &tmapDY, &full[st], x0 * 8, y0, 0, n);
        }
      }
    } else if (warp == 1) {
      // whole warp runs the uniform loops; only the tcgen05 instructions are predicated on one elected lane
      const bool leader = elect_one();
      {
        // descriptor words (MN-major, SWIZZLE_NONE): lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version<<14.
        // Only the low word changes per MMA: +dx (one pixel = 16 B) per tap, +2 tile rows per K=16 step.
        const uint32_t row16 = (uint32_t)p.BW;                                   // haloed row pitch in 16-byte units
        const uint32_t a_lo_const = (row16 & 0x3FFF) << 16;                      // LBO = next 8-pixel K group = next row
        const uint32_t a_hi = ((p.x_plane_bytes >> 4) & 0x3FFF) | (1u << 14);    // SBO = next 8-channel group = next plane
        const uint32_t b_lo_const = (128u >> 4) << 16;
        const uint32_t b_hi = (2048u >> 4) | (1u << 14);
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
            for (int k = 0; k < 8; ++k) {  // 8 x K=16 pixels (two 8-pixel tile rows each)
              const uint64_t bdesc = ((uint64_t)b_hi << 32) | (dy_lo + (uint32_t)k * b_kstep);
              // taps inner: consecutive MMAs go to different accumulators (same-accumulator chains serialise)
              for (int dx = 0; dx < p.KW; ++dx) {
                const uint64_t adesc = ((uint64_t)a_hi << 32) | (x_lo + (uint32_t)dx + (uint32_t)k * a_kstep);
                umma_f16(tmem_base + (uint32_t)dx * acc_stride, adesc, bdesc, idesc, k == 0 ? first : 1u);
              }
            }
            umma_commit(&empty[st]);
          }
          __syncwarp();
        }
        if (leader) umma_commit(acc_full);
        __syncwarp();
      }
    } else {
      const int q = warp & 3;
      const int ci = mb * 128 + q * 32 + lane;
      const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      for (int dx = 0; dx < p.KW; ++dx) {
        float* dst = p.dw + ((long long)(dy * p.KW + dx) * p.Cm + ci) * p.NC;
        for (int c0 = 0; c0 < p.NC; c0 += 16) {
          uint32_t raw[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(dx * p.acc_stride + c0), raw);
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
  p.BW = 8 + p.KW - 1;
  p.tiles_x = ceil_div(p.W, 8);
  p.tiles_y = ceil_div(p.H, 16);
  p.n_tiles = p.n_img * p.tiles_x * p.tiles_y;
  p.acc_stride = (int)round_up((uint32_t)p.NC, 32);
  int cols = 32;
  while (cols < p.KW * p.acc_stride) cols <<= 1;
  PBT_REQUIRE(cols <= 512, "wgrad: kw*cout exceeds tensor memory (512 columns)");
  p.tmem_cols = cols;
  p.idesc = make_idesc_f16(128, p.NC, d->dtype == PBT_BF16 ? 1 : 0, 1, 1);
  p.x_plane_bytes = (uint32_t)(16 * p.BW * 16);
  p.x_stage_bytes = round_up(16 * p.x_plane_bytes, 128);
  p.dy_stage_bytes = (uint32_t)((p.NC / 8) * 2048);
  p.stages = 2;
  p.dw = d->dw;
  p.inv_scale = d->inv_scale;
  p.debug_flags = d->debug_flags;
  const int m_blocks = ceil_div(p.Cm, 128);
  int splits = (2 * num_sms()) / (m_blocks * p.KH);
  if (splits < 1) splits = 1;
  if (splits > p.n_tiles) splits = p.n_tiles;
  p.splits = splits;
  const uint32_t smem_bytes = (uint32_t)p.stages * (p.x_stage_bytes + p.dy_stage_bytes) + 8u * (2 * p.stages + 1) + 16 + 128;
  PBT_REQUIRE(smem_bytes <= 227 * 1024, "wgrad: configuration does not fit shared memory");

  CUtensorMap tx, tg;
  int rc = make_p8_tmap(&tx, x, p.BW, 16, 16);
  if (rc != PBT_OK) return rc;
  rc = make_p8_tmap(&tg, g, 8, 16, p.NC / 8);
  if (rc != PBT_OK) return rc;

  PBT_CUDA_CHECK(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  dim3 grid(p.splits, p.KH, m_blocks);
  wgrad_kernel<<<grid, kWgThreads, smem_bytes, stream>>>(tx, tg, p);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

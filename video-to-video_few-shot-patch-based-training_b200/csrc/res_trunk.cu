// res_trunk.cu — the residual trunk of GeneratorJ (reference src/models/generator.py:18-58,107-110,223-224) on patch-sized
// maps as ONE launch: r_{b+1} = r_b + IN(convB(relu(IN(convA(relu(r_b)))))), 128 -> 128 channels, 3x3, for all blocks.
//
// Why: at the training shape (80 patches, 20x20 maps) the general conv kernel spends a launch per conv on 400 pixels per image:
// 8x16-pixel tiles cover 52 % of their MMA rows, every CTA re-streams the packed weights, and each conv is followed by two
// normalisation launches — 42 dependent launches for the 7 blocks.  An image's map is 100 KB in 16 bit, so here ONE CTA owns ONE
// image and keeps the running activation in shared memory for the whole trunk:
//
//   * the map lives in shared memory in the P8 plane layout with a row pitch of w+2 pixels: the two pad pixels are the zero halo
//     of the row above / below, and the GEMM's M index is simply the flattened pitched pixel index, so a 3x3 tap is a constant
//     byte offset of the A descriptor (no im2col, no per-tile halo) and 20x20 pixels occupy 440 of 512 MMA rows (86 %);
//   * tcgen05.mma (M 128, N 128, K 16; fp32 accumulators for up to four 128-row tiles = all 512 TMEM columns) reads A straight
//     from that map and B from a ring of packed-weight stages filled by cp.async.bulk; the weights of conv i+1 stream in while
//     the epilogue of conv i runs;
//   * InstanceNorm needs statistics over ONE image = one CTA: the epilogue writes the raw conv output (16 bit) back into the map
//     and to global memory (the backward pass reads it), reduces sum / sum of squares per channel in shared memory, then
//     normalises in place (+ ReLU, or + residual) — the result IS the next conv's A operand.  No grid-wide reduction, no
//     statistics launch, no normalise launch.
//
// Everything the unfused path saves for the backward sweep is written exactly as before (raw outputs, scale / shift tables,
// normalised activations), so generator_bwd.py is unchanged.
#include <stdlib.h>

#include "internal.h"
#include "ptx.cuh"

namespace pbt {

constexpr int kTEpiWarps = 16;              // = planes: pass 2 gives every warp one plane
constexpr int kTEpiThreads = 32 * kTEpiWarps;
static_assert(kTEpiWarps == 16, "pass 2 maps one warp to one of the 16 channel planes");
constexpr int kTrunkThreads = 64 + kTEpiThreads;   // warp 0: weight producer, warp 1: MMA issuer, warps 2-17: epilogue
constexpr int kTC = 128;                       // channels (N and K of every conv)
constexpr int kTPlanes = kTC / 8;
constexpr int kTStageTaps = 3;                 // taps per weight stage
constexpr int kTStages = 3;
constexpr uint32_t kTChunk = 32u * kTC * 2u;   // bytes of one (channel block of 32, tap) of packed weights
constexpr uint32_t kTStageBytes = kTStageTaps * kTChunk;

struct TrunkParams {
  const uint8_t* a0;            // relu(r_0), P8 16-bit
  float* r;                     // residual stream fp32 [n][16][h][w][8], in place
  uint8_t* raw_a[PBT_TRUNK_MAX_BLOCKS];
  uint8_t* hmid[PBT_TRUNK_MAX_BLOCKS];
  uint8_t* raw_b[PBT_TRUNK_MAX_BLOCKS];
  uint8_t* a_next[PBT_TRUNK_MAX_BLOCKS];   // relu(r_{b+1}); null for the last block
  const uint8_t* w_a[PBT_TRUNK_MAX_BLOCKS];
  const uint8_t* w_b[PBT_TRUNK_MAX_BLOCKS];
  float* scale_a[PBT_TRUNK_MAX_BLOCKS];
  float* shift_a[PBT_TRUNK_MAX_BLOCKS];
  float* scale_b[PBT_TRUNK_MAX_BLOCKS];
  float* shift_b[PBT_TRUNK_MAX_BLOCKS];
  uint8_t* last16;              // r_nb in 16 bit (may be a channel view of a wider tensor)
  long long last16_img_stride;  // elements
  long long img_stride;         // elements, of every other 16-bit tensor
  int n, h, w, nb, ntiles, pitch, rows_alloc;
  uint32_t plane_bytes, tmem_cols, idesc;
  int ws;
  float eps;
};

__device__ long long g_trunk_dbg[8];   // PBT_TRUNK_DBG build only: cycles per phase, block 0
#ifndef PBT_TRUNK_DBG
#define PBT_TRUNK_DBG 0
#endif
#define TDBG(slot, expr)                                                              \
  do {                                                                                \
    if (PBT_TRUNK_DBG && blockIdx.x == 0) {                                           \
      const long long _t0 = clock64();                                                \
      expr;                                                                           \
      if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&g_trunk_dbg[slot], (unsigned long long)(clock64() - _t0)); \
    } else {                                                                          \
      expr;                                                                           \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint64_t tdesc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <int DT>
__global__ void __launch_bounds__(kTrunkThreads, 1) res_trunk_fwd_kernel(const __grid_constant__ TrunkParams p) {
  extern __shared__ __align__(128) uint8_t smem[];   // used directly: pointer arithmetic on the array keeps LDS / STS (no generic accesses)
  uint8_t* sX = smem;                                                   // [16 planes][rows_alloc (+1)][16 B]
  uint8_t* sB = sX + (size_t)kTPlanes * p.plane_bytes;                  // weight ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)kTStages * kTStageBytes);
  uint64_t* b_full = bars;
  uint64_t* b_empty = bars + kTStages;
  uint64_t* acc_full = bars + 2 * kTStages;
  uint64_t* x_ready = acc_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_ready + 1);
  float* s_scale = reinterpret_cast<float*>(tmem_slot + 4);   // 16-byte aligned (the barrier block is a multiple of 16 bytes)
  float* s_shift = s_scale + kTC;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x;
  const int hw = p.h * p.w;
  const int M = p.h * p.pitch;              // flattened pitched pixels (pad pixels included)
  const int OFF = p.pitch + 1;              // map index of pixel (0, 0): one halo row and one halo pixel in front
  const int nconv = 2 * p.nb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTStages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(acc_full, 1);
    mbar_init(x_ready, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  // zero the whole map (halo rows, pad pixels and the overrun rows behind the last tile stay zero for good)
  {
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* x4 = reinterpret_cast<uint4*>(sX);
    const int total = (int)((size_t)kTPlanes * p.plane_bytes / 16);
    for (int i = threadIdx.x; i < total; i += kTrunkThreads) x4[i] = z;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_sync();
  const uint32_t tmem_base = *tmem_slot;
  // relu(r_0) -> map
  {
    const uint8_t* src = p.a0 + 2 * (long long)img * p.img_stride;
    const int total = kTPlanes * hw;
    for (int i = threadIdx.x; i < total; i += kTrunkThreads) {
      const int pl = i / hw, pix = i - pl * hw;
      const int y = pix / p.w, x = pix - y * p.w;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + ((long long)pl * hw + pix) * 16));
      *reinterpret_cast<uint4*>(sX + (size_t)pl * p.plane_bytes + (size_t)(OFF + y * p.pitch + x) * 16) = v;
    }
  }
  fence_proxy_async();
  __syncthreads();

  if (warp == 0) {
    // ------------------------------------------------------------ weight producer
    if (elect_one()) {
      int bi = 0;
      for (int ci = 0; ci < nconv; ++ci) {
        const uint8_t* wsrc = (ci & 1) ? p.w_b[ci >> 1] : p.w_a[ci >> 1];
        for (int cb = 0; cb < kTC / 32; ++cb)
          for (int g = 0; g < 9 / kTStageTaps; ++g, ++bi) {
            const int sb = bi % kTStages;
            mbar_wait_parked(&b_empty[sb], ((uint32_t)(bi / kTStages) & 1u) ^ 1u);
            mbar_arrive_expect_tx(&b_full[sb], kTStageBytes);
            bulk_load_1d(sB + (size_t)sb * kTStageBytes, wsrc + ((size_t)cb * 9 + (size_t)g * kTStageTaps) * kTChunk, kTStageBytes,
                         &b_full[sb]);
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    const bool leader = elect_one();
    const uint32_t a_lo_const = ((p.plane_bytes >> 4) & 0x3FFF) << 16;     // LBO: next 8 channels = next plane
    const uint32_t a_hi = (128u >> 4) | (1u << 14);                        // SBO: next 8 pixels = 128 contiguous bytes
    const uint32_t b_kstride = (uint32_t)kTC * 16u;
    const uint32_t b_lo_const = ((b_kstride >> 4) & 0x3FFF) << 16;
    const uint32_t b_hi = (128u >> 4) | (1u << 14);
    const uint32_t x_base = smem_u32(sX) >> 4;
    int bi = 0;
    for (int ci = 0; ci < nconv; ++ci) {
      if (ci > 0) {
        TDBG(1, mbar_wait_parked(x_ready, (uint32_t)(ci - 1) & 1u));      // the map holds this conv's input, the accumulators are drained
        tc_fence_after();
      }
      for (int cb = 0; cb < kTC / 32; ++cb)
        for (int g = 0; g < 9 / kTStageTaps; ++g, ++bi) {
          const int sb = bi % kTStages;
          TDBG(0, mbar_wait(&b_full[sb], (uint32_t)(bi / kTStages) & 1u));
          tc_fence_after();
          if (leader) {
            const uint32_t b_stage = (smem_u32(sB + (size_t)sb * kTStageBytes) >> 4) | b_lo_const;
#pragma unroll
            for (int j = 0; j < kTStageTaps; ++j) {
              const int ky = g, kx = j;                       // kTStageTaps == 3: one stage = one kernel row
              const uint32_t a_tap = x_base + (uint32_t)(OFF + (ky - 1) * p.pitch + (kx - 1)) + (uint32_t)(cb * 4) * (p.plane_bytes >> 4);
              const uint32_t b_tap = b_stage + (uint32_t)j * (kTChunk >> 4);
#pragma unroll
              for (int k = 0; k < 2; ++k)
                for (int t = 0; t < p.ntiles; ++t) {
                  const uint32_t d_t = tmem_base + (uint32_t)t * kTC;
                  const uint64_t ad = tdesc(((a_tap + (uint32_t)t * 128u + (uint32_t)k * 2u * (p.plane_bytes >> 4)) & 0x3FFF) | a_lo_const, a_hi);
                  const uint64_t bd = tdesc(b_tap + (uint32_t)k * 2u * (b_kstride >> 4), b_hi);
                  const uint32_t en = (cb == 0 && g == 0 && j == 0 && k == 0) ? 0u : 1u;
                  // weight-stationary run over the tiles of the image: B is fetched from shared memory once per (tap, K step)
                  if (!p.ws || p.ntiles == 1) umma_f16(d_t, ad, bd, p.idesc, en);
                  else if (t == 0) umma_f16_ws<0>(d_t, ad, bd, p.idesc, en);
                  else if (t == p.ntiles - 1) umma_f16_ws<2>(d_t, ad, bd, p.idesc, en);
                  else umma_f16_ws<1>(d_t, ad, bd, p.idesc, en);
                }
            }
            umma_commit(&b_empty[sb]);
          }
          __syncwarp();
        }
      if (leader) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (16 warps)
    // TMEM lanes are reachable per warp quadrant (warp id % 4): the four warps of a quadrant take 32 accumulator columns each in
    // pass 1; passes 2 and 3 only touch shared / global memory and are split over all 512 threads.
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;               // 0..3: column group (pass 1), row quarter (pass 2), plane group (pass 3)
    const int et = (warp - 2) * 32 + lane;         // 0..511
    const int r = q * 32 + lane;                   // accumulator row inside a tile
    const float inv_cnt = 1.f / (float)hw;
    for (int ci = 0; ci < nconv; ++ci) {
      const int b = ci >> 1;
      const bool second = ci & 1;
      const bool last = second && b == p.nb - 1;
      TDBG(2, mbar_wait_parked(acc_full, (uint32_t)ci & 1u));
      tc_fence_after();
      const long long t_p1 = PBT_TRUNK_DBG ? clock64() : 0;
      // pass 1: accumulators -> 16 bit -> map (pad / overrun rows: zero) and global raw output
      uint8_t* raw = (second ? p.raw_b[b] : p.raw_a[b]) + 2 * (long long)img * p.img_stride;
      for (int t = 0; t < p.ntiles; ++t) {
        const int m = t * 128 + r;
        if (t * 128 + q * 32 >= M) continue;       // warp-uniform: all 32 rows lie behind the map (they stay zero)
        const int y = m / p.pitch, x = m - y * p.pitch;
        const bool valid = m < M && x < p.w;
        const int pix = y * p.w + x;
        uint32_t v[32];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kTC + grp * 32), v);
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kTC + grp * 32 + 16), v + 16);
        tmem_ld_wait();
        uint8_t* xrow = sX + (size_t)(OFF + m) * 16 + (size_t)(4 * grp) * p.plane_bytes;
        uint8_t* grow = raw + ((long long)(4 * grp) * hw + pix) * 16;
#pragma unroll
        for (int pl = 0; pl < 4; ++pl) {
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = valid ? __uint_as_float(v[pl * 8 + k]) : 0.f;
          const uint4 o = pack8<DT>(f);
          *reinterpret_cast<uint4*>(xrow + (size_t)pl * p.plane_bytes) = o;
          if (valid) *reinterpret_cast<uint4*>(grow + (long long)pl * hw * 16) = o;
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(kTEpiThreads) : "memory");
      const long long t_p2 = PBT_TRUNK_DBG ? clock64() : 0;
      // pass 2: per-channel statistics of the 16-bit values (pad pixels are zero and add nothing).  warp = plane: a lane reads the
      // 8-channel chunk of every 32nd pixel (one 128-bit load per 8 values), the 32 lanes are folded by shuffles, lanes 0-7 finish
      // the plane's 8 channels
      {
        const int pl = warp - 2;                     // kTEpiWarps == kTPlanes
        const uint8_t* col = sX + (size_t)pl * p.plane_bytes + (size_t)OFF * 16;
        float sm[8], sq[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[k] = sq[k] = 0.f;
        for (int m = lane; m < M; m += 32) {
          float f[8];
          unpack8<DT>(*reinterpret_cast<const uint4*>(col + (size_t)m * 16), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            sm[k] += f[k];
            sq[k] = fmaf(f[k], f[k], sq[k]);
          }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sm[k] += __shfl_xor_sync(0xffffffffu, sm[k], o);
            sq[k] += __shfl_xor_sync(0xffffffffu, sq[k], o);
          }
        }
        float s = sm[0], ss = sq[0];
#pragma unroll
        for (int k = 1; k < 8; ++k)
          if (lane == k) { s = sm[k]; ss = sq[k]; }
        if (lane < 8) {
          const int c = pl * 8 + lane;
          const float mean = s * inv_cnt;
          const float var = fmaxf(ss * inv_cnt - mean * mean, 0.f);
          const float rstd = rsqrtf(var + p.eps);
          s_scale[c] = rstd;
          s_shift[c] = -mean * rstd;
          ((second ? p.scale_b[b] : p.scale_a[b]) + (long long)img * kTC)[c] = rstd;
          ((second ? p.shift_b[b] : p.shift_a[b]) + (long long)img * kTC)[c] = -mean * rstd;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kTEpiThreads) : "memory");
      const long long t_p3 = PBT_TRUNK_DBG ? clock64() : 0;
      // pass 3: normalise in place.  conv A: + ReLU -> hmid (map + global).  conv B: + residual -> r (fp32, global), and
      // relu(r) -> the next block's input (map + global), or r in 16 bit behind the last block.  thread = (row, group of 4 planes)
      uint8_t* out16 = second ? (last ? p.last16 + 2 * (long long)img * p.last16_img_stride
                                      : p.a_next[b] + 2 * (long long)img * p.img_stride)
                              : p.hmid[b] + 2 * (long long)img * p.img_stride;
      float* rimg = p.r + (long long)img * kTC * hw;
      const int p4 = grp * 4;
      for (int t = 0; t < p.ntiles; ++t) {
        const int m = t * 128 + r;
        const int y = m / p.pitch, x = m - y * p.pitch;
        if (!(m < M && x < p.w)) continue;
        const int pix = y * p.w + x;
        uint8_t* xrow = sX + (size_t)(OFF + m) * 16;
        float4 rr[8];
        if (second) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4* rp = reinterpret_cast<const float4*>(rimg + ((long long)(p4 + i) * hw + pix) * 8);
            rr[2 * i] = rp[0];
            rr[2 * i + 1] = rp[1];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int pl = p4 + i;
          float f[8];
          unpack8<DT>(*reinterpret_cast<const uint4*>(xrow + (size_t)pl * p.plane_bytes), f);
          const float4 sc0 = *reinterpret_cast<const float4*>(s_scale + pl * 8), sc1 = *reinterpret_cast<const float4*>(s_scale + pl * 8 + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_shift + pl * 8), sh1 = *reinterpret_cast<const float4*>(s_shift + pl * 8 + 4);
          f[0] = fmaf(f[0], sc0.x, sh0.x); f[1] = fmaf(f[1], sc0.y, sh0.y); f[2] = fmaf(f[2], sc0.z, sh0.z); f[3] = fmaf(f[3], sc0.w, sh0.w);
          f[4] = fmaf(f[4], sc1.x, sh1.x); f[5] = fmaf(f[5], sc1.y, sh1.y); f[6] = fmaf(f[6], sc1.z, sh1.z); f[7] = fmaf(f[7], sc1.w, sh1.w);
          if (second) {
            f[0] += rr[2 * i].x; f[1] += rr[2 * i].y; f[2] += rr[2 * i].z; f[3] += rr[2 * i].w;
            f[4] += rr[2 * i + 1].x; f[5] += rr[2 * i + 1].y; f[6] += rr[2 * i + 1].z; f[7] += rr[2 * i + 1].w;
            if (!last) {
              float4* wp = reinterpret_cast<float4*>(rimg + ((long long)pl * hw + pix) * 8);
              wp[0] = make_float4(f[0], f[1], f[2], f[3]);
              wp[1] = make_float4(f[4], f[5], f[6], f[7]);
            }
          }
          if (!last) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
          }
          const uint4 o = pack8<DT>(f);
          if (!last) *reinterpret_cast<uint4*>(xrow + (size_t)pl * p.plane_bytes) = o;
          *reinterpret_cast<uint4*>(out16 + ((long long)pl * hw + pix) * 16) = o;
        }
      }
      fence_proxy_async();       // the map was written through the generic proxy, the next conv's MMAs read it through the async one
      tc_fence_before();
      asm volatile("bar.sync 1, %0;" ::"n"(kTEpiThreads) : "memory");
      if (et == 0) mbar_arrive(x_ready);
      if (PBT_TRUNK_DBG && blockIdx.x == 0 && et == 0) {
        const long long t_e = clock64();
        atomicAdd((unsigned long long*)&g_trunk_dbg[3], (unsigned long long)(t_p2 - t_p1));
        atomicAdd((unsigned long long*)&g_trunk_dbg[4], (unsigned long long)(t_p3 - t_p2));
        atomicAdd((unsigned long long*)&g_trunk_dbg[5], (unsigned long long)(t_e - t_p3));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace pbt

using namespace pbt;

static bool tact_ok(const pbt_act_t& t, int n, int h, int w) {
  return t.ptr && aligned16(t.ptr) && t.n == n && t.c == kTC && t.h == h && t.w == w && t.img_stride % 8 == 0;
}

// debug builds (-DPBT_TRUNK_DBG=1): cycles block 0 spent {waiting for weights, waiting for the map, waiting for the accumulators
// (sum over the 4 epilogue warps' lane 0), pass 1, pass 2, pass 3}; reading resets the counters
extern "C" int pbt_debug_trunk_cycles(long long* out8) {
  long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out8, g_trunk_dbg, sizeof(z)) != cudaSuccess) return PBT_ERR_ARG;
  if (cudaMemcpyToSymbol(g_trunk_dbg, z, sizeof(z)) != cudaSuccess) return PBT_ERR_ARG;
  return PBT_OK;
}

static uint32_t trunk_plane_bytes(int h, int w) {
  const int pitch = w + 2, ntiles = (h * pitch + 127) / 128;
  const int rows = (ntiles * 128 + 2 * pitch + 2 + 7) / 8 * 8;
  return (uint32_t)rows * 16u + 16u;     // + 16: planes start 4 banks apart
}
static size_t trunk_smem_bytes(int h, int w) {
  return 128 + (size_t)kTPlanes * trunk_plane_bytes(h, w) + (size_t)kTStages * kTStageBytes + 8 * (2 * kTStages + 2) + 16 + 2 * kTC * 4;
}

extern "C" int pbt_res_trunk_supported(int32_t channels, int32_t h, int32_t w) {
  return channels == kTC && h >= 1 && w >= 1 && h * (w + 2) <= 512 && trunk_smem_bytes(h, w) <= 227 * 1024;
}

extern "C" int pbt_res_trunk_fwd(const pbt_res_trunk_desc_t* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(d && d->n_blocks >= 1 && d->n_blocks <= PBT_TRUNK_MAX_BLOCKS, "res_trunk: 1..16 blocks");
  PBT_REQUIRE(d->dtype == PBT_BF16 || d->dtype == PBT_FP16, "res_trunk: bad dtype");
  const pbt_act_t& a0 = d->a[0];
  PBT_REQUIRE(a0.ptr && a0.n > 0 && pbt_res_trunk_supported(a0.c, a0.h, a0.w), "res_trunk: needs 128 channels and a map of h*(w+2) <= 512 pixels that fits shared memory");
  const int n = a0.n, h = a0.h, w = a0.w;
  const long long dense = (long long)kTC * h * w;
  PBT_REQUIRE(d->residual32 && aligned16(d->residual32) && d->eps > 0.f, "res_trunk: residual stream / eps");
  PBT_REQUIRE(tact_ok(d->last16, n, h, w), "res_trunk: last16");
  TrunkParams p;
  memset(&p, 0, sizeof(p));
  for (int b = 0; b < d->n_blocks; ++b) {
    PBT_REQUIRE(tact_ok(d->a[b], n, h, w) && tact_ok(d->raw_a[b], n, h, w) && tact_ok(d->hmid[b], n, h, w) && tact_ok(d->raw_b[b], n, h, w),
                "res_trunk: activation tensors must be [n, 128, h, w] P8");
    PBT_REQUIRE(d->a[b].img_stride == dense && d->raw_a[b].img_stride == dense && d->hmid[b].img_stride == dense &&
                    d->raw_b[b].img_stride == dense, "res_trunk: saved activations must be dense tensors");
    PBT_REQUIRE(d->w_a[b] && d->w_b[b] && aligned16(d->w_a[b]) && aligned16(d->w_b[b]), "res_trunk: packed weights");
    PBT_REQUIRE(d->scale_a[b] && d->shift_a[b] && d->scale_b[b] && d->shift_b[b], "res_trunk: scale / shift tables");
    p.raw_a[b] = static_cast<uint8_t*>(d->raw_a[b].ptr);
    p.hmid[b] = static_cast<uint8_t*>(d->hmid[b].ptr);
    p.raw_b[b] = static_cast<uint8_t*>(d->raw_b[b].ptr);
    p.a_next[b] = b + 1 < d->n_blocks ? static_cast<uint8_t*>(d->a[b + 1].ptr) : nullptr;
    p.w_a[b] = static_cast<const uint8_t*>(d->w_a[b]);
    p.w_b[b] = static_cast<const uint8_t*>(d->w_b[b]);
    p.scale_a[b] = d->scale_a[b]; p.shift_a[b] = d->shift_a[b];
    p.scale_b[b] = d->scale_b[b]; p.shift_b[b] = d->shift_b[b];
  }
  p.a0 = static_cast<const uint8_t*>(a0.ptr);
  p.r = d->residual32;
  p.last16 = static_cast<uint8_t*>(d->last16.ptr);
  p.last16_img_stride = d->last16.img_stride;
  p.img_stride = dense;
  p.n = n; p.h = h; p.w = w; p.nb = d->n_blocks;
  p.pitch = w + 2;
  p.ntiles = (h * p.pitch + 127) / 128;
  p.plane_bytes = trunk_plane_bytes(h, w);
  p.rows_alloc = (int)((p.plane_bytes - 16u) / 16u);
  p.tmem_cols = p.ntiles <= 1 ? 128u : (p.ntiles == 2 ? 256u : 512u);
  p.idesc = make_idesc_f16(128, kTC, d->dtype == PBT_BF16 ? 1 : 0, 0, 0);
  p.eps = d->eps;
  {
    const char* e = getenv("PBT_TRUNK_WS");
    p.ws = e ? atoi(e) : 0;
  }
  const size_t smem = trunk_smem_bytes(h, w);
  if (d->dtype == PBT_BF16) {
    PBT_CUDA_CHECK(cudaFuncSetAttribute(res_trunk_fwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbt::launch(res_trunk_fwd_kernel<0>, dim3(n), dim3(kTrunkThreads), smem, st, p);
  } else {
    PBT_CUDA_CHECK(cudaFuncSetAttribute(res_trunk_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pbt::launch(res_trunk_fwd_kernel<1>, dim3(n), dim3(kTrunkThreads), smem, st, p);
  }
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

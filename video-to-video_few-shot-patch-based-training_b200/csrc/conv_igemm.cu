// conv_igemm.cu — stride-1 convolution as implicit GEMM on tcgen05 (sm_100a).
//
// Replaces nn.Conv2d in GeneratorJ (reference src/models/generator.py:41,49,126,133,136,141,171,200)
// for forward and dgrad.
//
// Design (B200-first, not a cuDNN-style im2col):
//  * activations live in the P8 layout [n][c/8][y][x][8] (16-bit).  One CTA owns T x-adjacent
//    output tiles of 8x16 pixels.  For every block of input channels ONE TMA load brings the haloed
//    input region ((8T+kw-1) x (16+kh-1) pixels x blk_c channels, zero-filled outside the image)
//    into shared memory; all kh*kw taps are then served from that single copy by shifting the
//    start address of a SWIZZLE_NONE UMMA descriptor by whole pixels (16 B).  The 49x (7x7) / 9x (3x3)
//    re-read of the input from L2 that a per-tap im2col load does never happens.
//  * UMMA A operand (K-major, no swizzle): row r of the 128-row tile = pixel (r>>3, r&7); the 8 rows
//    of a core matrix are 8 x-adjacent pixels (16 B apart), core matrices along M are image rows
//    (SBO = haloed row pitch), core matrices along K are channel planes (LBO = plane pitch).
//  * B operand: pre-packed weights [blk][tap][k/8][cout][8] streamed by 1-D bulk copies through a ring whose
//    stages hold a GROUP of taps (one full/empty barrier round trip per group, not per tap).
//  * fp32 accumulators in TMEM (T accumulators of `cout` columns); M=128, N=cout, K=16 per instruction.
//  * warp roles: warp 0 = TMA/bulk producer, warp 1 = TMEM owner + single-thread MMA issuer,
//    warps 2..5 = epilogue (TMEM -> registers -> fused bias/act/affine/mask/residual/stats/1x1-head -> global).
//  * the issuing thread is the critical resource (one tcgen05.mma per 16..64 tensor cycles): the kernel is
//    templated on T and on the K steps per channel block so that the issue loop is fully unrolled and each MMA
//    costs two 32-bit adds on the low descriptor words.
#include "internal.h"
#include "ptx.cuh"

namespace pbt {

struct ConvKParams {
  int n_img, H, W;
  int blk_p, n_blk, Cp;
  int KH, KW, pad_t, pad_l;
  int NC, acc_stride, tmem_cols;
  int tiles_x, tiles_y;
  int BW, BH;
  int dt;
  uint32_t idesc;
  int a_stages, b_stages, b_group;
  uint32_t a_stage_bytes, b_stage_bytes;
  const uint8_t* wpack;
  const float* bias;
  int act;
  const float* post_scale;
  const float* post_shift;
  const uint8_t* mask;
  long long mask_img_stride;
  const float* addend32;
  float* out32;
  uint8_t* out;
  long long out_img_stride;
  float* stats_partial;
  const float* head_w;
  const float* head_b;
  float* head_out;
  int head_tanh;
  int debug_flags;
  long long* debug_buf;  // bring-up: per-CTA phase timestamps (clock64), 8 slots per CTA
};

// warps: 0 = producer, 1..NI = MMA issuers (one per accumulator when PBT_MULTI_ISSUE), then 4 epilogue warps
// Measured on B200 (profiles/r1_issue_experiments.md): one CTA's MMAs execute back to back with ~40-70 cycles of
// exposed operand-fetch latency each, regardless of how many warps issue them or how accumulators are
// interleaved; streams of DIFFERENT co-resident CTAs overlap.  One issuer warp is therefore enough.
#ifndef PBT_MULTI_ISSUE
#define PBT_MULTI_ISSUE 0
#endif
constexpr int kMaxThreads = 256;
__host__ __device__ constexpr int num_issuers(int T) { return PBT_MULTI_ISSUE ? T : 1; }
__host__ __device__ constexpr int conv_threads(int T) { return 32 * (5 + num_issuers(T)); }

// Sum over the 32 lanes of a warp of 16 per-lane values; afterwards every lane holds the total of
// column  col = 8*b4 + 4*b3 + 2*b2 + b1  (b_i = bit i of the lane id); lanes differing only in bit 0 agree.
__device__ __forceinline__ float warp_colsum16(const float* v, int lane) {
  float a[8], b[4], c[2];
  bool hi = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float send = hi ? v[i] : v[i + 8];
    float keep = hi ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float send = hi ? a[i] : a[i + 4];
    float keep = hi ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  hi = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float send = hi ? b[i] : b[i + 2];
    float keep = hi ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  hi = lane & 2;
  float send = hi ? c[0] : c[1];
  float keep = hi ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  return d;
}

__device__ __forceinline__ void unpack8_rt(int dt, const uint4& u, float* f) {
  if (dt == 0) unpack8<0>(u, f);
  else unpack8<1>(u, f);
}
__device__ __forceinline__ uint4 pack8_rt(int dt, const float* f) { return dt == 0 ? pack8<0>(f) : pack8<1>(f); }

__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

template <int T, int KB>
__global__ void __launch_bounds__(kMaxThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmapA, const ConvKParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_stages * p.a_stage_bytes;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + (size_t)p.b_stages * p.b_stage_bytes);
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 2);  // [4 warps][2][NC]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int NI = num_issuers(T);
  constexpr int kEpi0 = 32 * (1 + NI);  // first epilogue thread

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  const int n = blockIdx.x / tiles_per_img;
  const int rem = blockIdx.x - n * tiles_per_img;
  const int tyi = rem / p.tiles_x;
  const int txi = rem - tyi * p.tiles_x;
  const int x0 = txi * 8 * T;
  const int y0 = tyi * 16;
  const int ntaps = p.KH * p.KW;
  const int ngroups = (ntaps + p.b_group - 1) / p.b_group;

#define PBT_STAMP(slot)                                                         \
  do {                                                                           \
    if (p.debug_buf) p.debug_buf[(long long)blockIdx.x * 8 + (slot)] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) {
    PBT_STAMP(0);
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], NI);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], NI);
    }
    mbar_init(acc_full, NI);
    fence_barrier_init();
    prefetch_tmap(&tmapA);
  }
  if (warp == 1) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    PBT_STAMP(1);
    if (p.debug_buf) {
      uint32_t smid;
      asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
      p.debug_buf[(long long)blockIdx.x * 8 + 7] = smid;
    }
  }

  if (warp == 0) {
    // ------------------------------------------------------------ producer
    if (elect_one()) {
      int bi = 0;  // running B-group counter
      for (int cb = 0; cb < p.n_blk; ++cb) {
        const int sa = cb % p.a_stages;
        const uint32_t pa = (uint32_t)(cb / p.a_stages) & 1u;
        mbar_wait(&a_empty[sa], pa ^ 1u);
        mbar_arrive_expect_tx(&a_full[sa], (uint32_t)(p.blk_p * p.BH * p.BW * 16));
        tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, &tmapA, &a_full[sa], (x0 - p.pad_l) * 8, y0 - p.pad_t,
                    cb * p.blk_p, n);
        const int pib = min(p.blk_p, p.Cp - cb * p.blk_p);
        const uint32_t chunk = (uint32_t)(pib * p.NC * 16);
        const uint8_t* wsrc = p.wpack + (size_t)cb * ntaps * ((size_t)p.blk_p * p.NC * 16);
        for (int g = 0; g < ngroups; ++g, ++bi) {
          const int sb = bi % p.b_stages;
          const uint32_t pb = (uint32_t)(bi / p.b_stages) & 1u;
          const int tap0 = g * p.b_group;
          const int nt = min(p.b_group, ntaps - tap0);
          mbar_wait(&b_empty[sb], pb ^ 1u);
          mbar_arrive_expect_tx(&b_full[sb], chunk * (uint32_t)nt);
          // the taps of a group are contiguous in the packed weights: one bulk copy per group
          bulk_load_1d(sB + (size_t)sb * p.b_stage_bytes, wsrc + (size_t)tap0 * chunk, chunk * (uint32_t)nt, &b_full[sb]);
        }
      }
    }
  } else if (warp <= NI) {
    // ------------------------------------------------------------ MMA issuer(s)
    // With NI == T every accumulator has its own issuing warp: the MMA streams of different warps overlap their
    // operand-fetch latency (a single stream executes MMAs back to back with the fetch latency exposed).
    const int ti = warp - 1;
    // The WHOLE warp runs the (warp-uniform) loops so that descriptor arithmetic and barrier waits live in the
    // uniform datapath; only the tcgen05 instructions are predicated on one elected lane.  (With the loop inside
    // `if (elect_one())` the compiler built every descriptor in vector registers and paid ~10 R2UR moves per
    // group of MMAs, which made the issuing thread — not the tensor pipe — the bottleneck.)
    const bool leader = elect_one();
    {
      const uint32_t plane_bytes = (uint32_t)(p.BH * p.BW * 16);
      const uint32_t row_bytes = (uint32_t)(p.BW * 16);
      const uint32_t b_kstride = (uint32_t)(p.NC * 16);
      // descriptor words: lo = addr>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version(1)<<14
      const uint32_t a_lo_const = ((plane_bytes >> 4) & 0x3FFF) << 16;
      const uint32_t a_hi = ((row_bytes >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t b_lo_const = ((b_kstride >> 4) & 0x3FFF) << 16;
      const uint32_t b_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_kstep = (2u * plane_bytes) >> 4;  // two channel planes per K=16
      const uint32_t b_kstep = (2u * b_kstride) >> 4;
      const uint32_t acc_stride = (uint32_t)p.acc_stride;
      const uint32_t idesc = p.idesc;
      int bi = 0;
      for (int cb = 0; cb < p.n_blk; ++cb) {
        const int sa = cb % p.a_stages;
        const uint32_t pa = (uint32_t)(cb / p.a_stages) & 1u;
        mbar_wait(&a_full[sa], pa);
        tc_fence_after();
        if (cb == 0 && leader) PBT_STAMP(2);
        const uint32_t a_base = (smem_u32(sA + (size_t)sa * p.a_stage_bytes) >> 4) | a_lo_const;
        const int pib = min(p.blk_p, p.Cp - cb * p.blk_p);
        const int k16n = pib >> 1;
        const uint32_t chunk16 = (uint32_t)(pib * p.NC);  // bytes/16 of one tap's weights
        int dy = 0, dx = 0;
        for (int g = 0; g < ngroups; ++g, ++bi) {
          const int sb = bi % p.b_stages;
          const uint32_t pb = (uint32_t)(bi / p.b_stages) & 1u;
          const int nt = min(p.b_group, ntaps - g * p.b_group);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          uint32_t b_lo = (smem_u32(sB + (size_t)sb * p.b_stage_bytes) >> 4) | b_lo_const;
          for (int j = 0; j < nt; ++j) {
            const uint32_t a_tap = a_base + (uint32_t)(dy * p.BW + dx);
            const uint32_t first = (cb == 0 && g == 0 && j == 0) ? 0u : 1u;
            // k outer / t inner: consecutive MMAs target DIFFERENT accumulators, so the read-modify-write
            // dependency on one TMEM tile is T instructions apart (measured: same-accumulator chains serialise)
            if (leader) {
#pragma unroll
              for (int k = 0; k < KB; ++k) {
                if (k < k16n) {
#pragma unroll
                  for (int t = 0; t < T; ++t)
                    if (NI == 1 || t == ti)
                      umma_f16(tmem_base + (uint32_t)t * acc_stride,
                               desc64(a_tap + (uint32_t)(t * 8) + (uint32_t)k * a_kstep, a_hi),
                               desc64(b_lo + (uint32_t)k * b_kstep, b_hi), idesc, (k == 0) ? first : 1u);
                }
              }
            }
            b_lo += chunk16;
            if (++dx == p.KW) { dx = 0; ++dy; }
          }
          if (leader) umma_commit(&b_empty[sb]);
          __syncwarp();
        }
        if (leader) umma_commit(&a_empty[sa]);
        __syncwarp();
      }
      if (leader) {
        umma_commit(acc_full);
        if (ti == 0) PBT_STAMP(3);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..5)
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int NC = p.NC;
    const int dt = p.dt;
    const bool do_stats = p.stats_partial != nullptr;
    float* my_stats = s_stats + (size_t)q * 2 * NC;
    if (do_stats) {
      for (int i = lane; i < 2 * NC; i += 32) my_stats[i] = 0.f;
      __syncwarp();
    }
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    const int y = y0 + ty;
    const long long plane_px = (long long)p.H * p.W;
    const int col_of_lane = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (threadIdx.x == kEpi0) PBT_STAMP(4);

    for (int t = 0; t < T; ++t) {
      const int x = x0 + 8 * t + tx;
      const bool valid = (y < p.H) && (x < p.W);
      const long long pix = (long long)y * p.W + x;
      float h0 = 0.f, h1 = 0.f, h2 = 0.f;
      for (int c0 = 0; c0 < NC; c0 += 16) {
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.acc_stride + c0), raw);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(&p.bias[c0 + i]);
        }
        if (p.act == PBT_ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (p.act == PBT_ACT_LEAKY02) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
        }
        if (p.post_scale) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(&p.post_scale[c0 + i]), __ldg(&p.post_shift[c0 + i]));
        }
        if (valid) {
          if (p.mask) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint4 m = *reinterpret_cast<const uint4*>(
                  p.mask + 2 * ((long long)n * p.mask_img_stride + ((long long)(c0 / 8 + hh) * plane_px + pix) * 8));
              float mf[8];
              unpack8_rt(dt, m, mf);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[hh * 8 + i] = mf[i] > 0.f ? v[hh * 8 + i] : 0.f;
            }
          }
          if (p.addend32) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float4* ap = reinterpret_cast<const float4*>(
                  p.addend32 + (((long long)n * (NC / 8) + (c0 / 8 + hh)) * plane_px + pix) * 8);
              const float4 a0 = ap[0], a1 = ap[1];
              v[hh * 8 + 0] += a0.x; v[hh * 8 + 1] += a0.y; v[hh * 8 + 2] += a0.z; v[hh * 8 + 3] += a0.w;
              v[hh * 8 + 4] += a1.x; v[hh * 8 + 5] += a1.y; v[hh * 8 + 6] += a1.z; v[hh * 8 + 7] += a1.w;
            }
          }
          if (p.out32) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float4* op =
                  reinterpret_cast<float4*>(p.out32 + (((long long)n * (NC / 8) + (c0 / 8 + hh)) * plane_px + pix) * 8);
              op[0] = make_float4(v[hh * 8 + 0], v[hh * 8 + 1], v[hh * 8 + 2], v[hh * 8 + 3]);
              op[1] = make_float4(v[hh * 8 + 4], v[hh * 8 + 5], v[hh * 8 + 6], v[hh * 8 + 7]);
            }
          }
        }
        if (p.out) {
          // round to the storage type; statistics and the head see the rounded values
          const uint4 u0 = pack8_rt(dt, v), u1 = pack8_rt(dt, v + 8);
          if (valid) {
            uint8_t* ob = p.out + 2 * ((long long)n * p.out_img_stride + ((long long)(c0 / 8) * plane_px + pix) * 8);
            *reinterpret_cast<uint4*>(ob) = u0;
            *reinterpret_cast<uint4*>(ob + 2 * plane_px * 8) = u1;
          }
          unpack8_rt(dt, u0, v);
          unpack8_rt(dt, u1, v + 8);
        }
        if (p.head_w) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            h0 = fmaf(v[i], __ldg(&p.head_w[c0 + i]), h0);
            h1 = fmaf(v[i], __ldg(&p.head_w[NC + c0 + i]), h1);
            h2 = fmaf(v[i], __ldg(&p.head_w[2 * NC + c0 + i]), h2);
          }
        }
        if (do_stats) {
          float s[16], s2[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s[i] = valid ? v[i] : 0.f;
            s2[i] = s[i] * s[i];
          }
          const float cs = warp_colsum16(s, lane);
          const float cs2 = warp_colsum16(s2, lane);
          if ((lane & 1) == 0) {
            my_stats[c0 + col_of_lane] += cs;
            my_stats[NC + c0 + col_of_lane] += cs2;
          }
        }
      }
      if (p.head_w && valid) {
        h0 += __ldg(&p.head_b[0]);
        h1 += __ldg(&p.head_b[1]);
        h2 += __ldg(&p.head_b[2]);
        if (p.head_tanh) {
          h0 = tanhf(h0);
          h1 = tanhf(h1);
          h2 = tanhf(h2);
        }
        float* ho = p.head_out + (long long)n * 3 * plane_px + pix;
        ho[0] = h0;
        ho[plane_px] = h1;
        ho[2 * plane_px] = h2;
      }
    }
    if (do_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int e = threadIdx.x - kEpi0;  // 0..127
      float* dst = p.stats_partial + ((long long)n * tiles_per_img + rem) * 2 * NC;
      for (int i = e; i < 2 * NC; i += 128)
        dst[i] = s_stats[i] + s_stats[2 * NC + i] + s_stats[4 * NC + i] + s_stats[6 * NC + i];
    }
  }

  if (threadIdx.x == kEpi0) PBT_STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0) PBT_STAMP(6);
}

static int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

template <int T, int KB>
static int launch_conv(const CUtensorMap& tmap, const ConvKParams& p, int grid, uint32_t smem_bytes, cudaStream_t stream) {
  PBT_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<T, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  conv_igemm_kernel<T, KB><<<grid, conv_threads(T), smem_bytes, stream>>>(tmap, p);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

template <int T>
static int launch_conv_kb(int kb, const CUtensorMap& tmap, const ConvKParams& p, int grid, uint32_t smem, cudaStream_t s) {
  switch (kb) {
    case 1: return launch_conv<T, 1>(tmap, p, grid, smem, s);
    case 2: return launch_conv<T, 2>(tmap, p, grid, smem, s);
    default: return launch_conv<T, 4>(tmap, p, grid, smem, s);
  }
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_conv_num_tiles(int32_t h, int32_t w, int32_t tiles_per_cta) {
  if (tiles_per_cta < 1) return PBT_ERR_ARG;
  return ceil_div(w, 8 * tiles_per_cta) * ceil_div(h, 16);
}

extern "C" int pbt_conv_fwd(const pbt_conv_desc_t* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(d != nullptr, "conv: null descriptor");
  const pbt_act_t& in = d->in;
  PBT_REQUIRE(in.ptr && aligned16(in.ptr), "conv: input pointer null or not 16-byte aligned");
  PBT_REQUIRE(in.n > 0 && in.h > 0 && in.w > 0, "conv: empty input");
  PBT_REQUIRE(in.c > 0 && in.c % 16 == 0, "conv: cin must be a multiple of 16");
  PBT_REQUIRE(d->cout >= 16 && d->cout <= 256 && d->cout % 16 == 0, "conv: cout must be a multiple of 16 in [16,256]");
  PBT_REQUIRE(d->blk_c == 16 || d->blk_c == 32 || d->blk_c == 64, "conv: blk_c must be 16, 32 or 64");
  PBT_REQUIRE(d->kh >= 1 && d->kh <= 7 && d->kw >= 1 && d->kw <= 7, "conv: kernel size must be in [1,7]");
  PBT_REQUIRE(d->pad_t >= 0 && d->pad_t < d->kh && d->pad_l >= 0 && d->pad_l < d->kw, "conv: bad padding");
  PBT_REQUIRE(d->tiles_per_cta >= 1 && d->tiles_per_cta <= 3, "conv: tiles_per_cta must be 1..3");
  PBT_REQUIRE(d->dtype == PBT_BF16 || d->dtype == PBT_FP16, "conv: bad dtype");
  PBT_REQUIRE(d->wpack && aligned16(d->wpack), "conv: packed weights null or misaligned");
  PBT_REQUIRE(in.img_stride % 8 == 0, "conv: img_stride must be a multiple of 8 elements");
  PBT_REQUIRE((long long)in.w * 8 < (1ll << 31), "conv: image too wide");

  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.n_img = in.n; p.H = in.h; p.W = in.w;
  p.Cp = in.c / 8;
  p.blk_p = d->blk_c / 8;
  p.n_blk = ceil_div(p.Cp, p.blk_p);
  p.KH = d->kh; p.KW = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  const int T = d->tiles_per_cta;
  p.NC = d->cout;
  p.dt = d->dtype;
  p.acc_stride = (int)round_up((uint32_t)p.NC, 32);
  p.tmem_cols = pow2_cols(T * p.acc_stride);
  PBT_REQUIRE(p.tmem_cols <= 512, "conv: tiles_per_cta*cout exceeds tensor memory (512 columns)");
  p.BW = 8 * T + p.KW - 1;
  p.BH = 16 + p.KH - 1;
  PBT_REQUIRE(p.BW <= 32, "conv: haloed tile wider than 32 pixels (reduce tiles_per_cta)");
  p.tiles_x = ceil_div(p.W, 8 * T);
  p.tiles_y = ceil_div(p.H, 16);
  p.idesc = make_idesc_f16(128, p.NC, d->dtype == PBT_BF16 ? 1 : 0, 0, 0);
  p.a_stage_bytes = round_up((uint32_t)(p.blk_p * p.BH * p.BW * 16), 128);
  p.a_stages = p.n_blk > 1 ? 2 : 1;
  p.wpack = static_cast<const uint8_t*>(d->wpack);
  p.bias = d->bias; p.act = d->act; p.post_scale = d->post_scale; p.post_shift = d->post_shift;
  PBT_REQUIRE((d->post_scale == nullptr) == (d->post_shift == nullptr), "conv: post_scale/post_shift must come together");
  if (d->mask.ptr) {
    PBT_REQUIRE(d->mask.c >= d->cout && d->mask.h == in.h && d->mask.w == in.w && d->mask.n == in.n && aligned16(d->mask.ptr),
                "conv: mask shape mismatch");
    p.mask = static_cast<const uint8_t*>(d->mask.ptr);
    p.mask_img_stride = d->mask.img_stride;
  }
  p.addend32 = d->addend32; p.out32 = d->out32;
  if (d->out.ptr) {
    PBT_REQUIRE(d->out.c >= d->cout && d->out.h == in.h && d->out.w == in.w && d->out.n == in.n && aligned16(d->out.ptr),
                "conv: output shape mismatch");
    p.out = static_cast<uint8_t*>(d->out.ptr);
    p.out_img_stride = d->out.img_stride;
  }
  p.stats_partial = d->stats_partial;
  p.head_w = d->head_w; p.head_b = d->head_b; p.head_out = d->head_out; p.head_tanh = d->head_tanh;
  PBT_REQUIRE(!d->head_w || (d->head_b && d->head_out), "conv: head needs head_b and head_out");
  p.debug_flags = d->debug_flags;
  p.debug_buf = static_cast<long long*>(d->debug_buf);

  // shared memory budget: A ring + B ring (groups of taps) + barriers + tmem slot + stats scratch.
  // Aim at two co-resident CTAs per SM (one CTA's epilogue/prologue overlaps the other's main loop).
  const uint32_t tail = 8u * (2 * 2 + 2 * 8 + 1) + 16 + (uint32_t)(4 * 2 * p.NC * 4) + 128;
  const uint32_t a_total = (uint32_t)p.a_stages * p.a_stage_bytes;
  const uint32_t chunk = (uint32_t)(p.blk_p * p.NC * 16);  // one tap of one channel block
  const int ntaps = p.KH * p.KW;
  const uint32_t budget = 112 * 1024;
  int group = (int)(16384u / chunk);
  if (group < 1) group = 1;
  if (group > ntaps) group = ntaps;
  int stages = 4;
  auto fits = [&](int g, int s) { return a_total + (uint32_t)s * round_up((uint32_t)g * chunk, 128) + tail <= budget; };
  while (!fits(group, stages) && (stages > 2 || group > 1)) {
    if (stages > 2) --stages;
    else --group;
  }
  if (d->debug_flags & 4) group = 1;  // bring-up: one tap per stage
  p.b_group = group;
  p.b_stages = stages;
  p.b_stage_bytes = round_up((uint32_t)group * chunk, 128);
  uint32_t smem_bytes = a_total + (uint32_t)p.b_stages * p.b_stage_bytes + tail;
  if ((d->debug_flags & 8) && smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // bring-up: force one CTA per SM
  PBT_REQUIRE(smem_bytes <= 227 * 1024, "conv: configuration does not fit shared memory");

  CUtensorMap tmap;
  int rc = make_p8_tmap(&tmap, in, p.BW, p.BH, p.blk_p);
  if (rc != PBT_OK) return rc;

  const int grid = p.n_img * p.tiles_x * p.tiles_y;
  const int kb = d->blk_c / 16;
  switch (T) {
    case 1: return launch_conv_kb<1>(kb, tmap, p, grid, smem_bytes, stream);
    case 2: return launch_conv_kb<2>(kb, tmap, p, grid, smem_bytes, stream);
    default: return launch_conv_kb<3>(kb, tmap, p, grid, smem_bytes, stream);
  }
}

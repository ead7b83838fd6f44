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
//    EW epilogue warps (8: two per TMEM lane quarter; 4 in the small-footprint configuration):
//    TMEM -> registers -> fused bias/act/affine/mask/residual/stats/1x1-head -> global.  In the on-load modes the same
//    warps first build the A tiles: bilinear x2 interpolation of a low-res staging tile (upsample-on-load) or
//    InstanceNorm + activation of the landed raw tile in place (normalise-on-load).
//  * the kernel is templated on T and on the K steps per channel block so that the issue loop is fully unrolled and
//    each MMA costs two 32-bit adds on the low descriptor words; on EW; and on PAIR (cta_group::2: a cluster of two
//    CTAs runs one M=256 MMA stream and each CTA stages half of the weight columns).
//  * CTAs are persistent over the units of long launches (barrier phases continue across units).
//  Measurements behind these choices: profiles/r1_issue_experiments.md, tools/conv_timeline.py, tools/conv_occ.py.
#include <cstdlib>
#include <type_traits>

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
  int up, LH, LW, LBH, LBW;   // upsample-on-load: low-res dims and staging box (pixels)
  int n_units;                // work units (CTA tiles of T sub-tiles) of the launch; CTAs are persistent over them
  int nrm, nblk0, pre_c, pre_act;  // normalise-on-load: the first nblk0 channel blocks come from `pre` and get act(x*s+t)
  const float* pre_scale;
  const float* pre_shift;
  uint32_t l_stage_bytes;
  int up_nrm;            // upsample-on-load whose first nblk0 channel blocks are raw conv outputs (normalised in the staged low-res tile)
  int tp;                // tap pairs (first layer, <= 8 input channels): K = 16 = one 8-channel plane at pixels x and x+1
  int a_planes;          // channel planes TMA-loaded per block (= blk_p; 1 in tap-pair mode)
  int KWs, dx_step;      // tap walk along x: dx += dx_step while dx < KWs (KW, 1; tap pairs: KW + 1, 2)
  int ws_b;              // issue the T MMAs of a (tap, K step) as one weight-stationary run (B fetched once)
  int VH, VW;            // valid output window (<= H, W): outputs outside it are stored as zero and left out of the statistics
  int bt;                // batch tiles: the T tiles of a CTA are the SAME spatial tile of T consecutive images (small maps)
  uint32_t a_tile16;     // A-descriptor step from tile t to t+1 in 16-byte units: 8 pixels (x-adjacent) or one staged tile (bt)
  uint32_t a_tile_bytes; // bt: shared-memory pitch of the T per-image tiles inside an A stage
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
constexpr int kMaxThreads = PBT_MULTI_ISSUE ? 384 : 320;
constexpr int kEpiWarps = 8;  // default: two warps per TMEM lane quadrant, they split the accumulator columns.
// EW = 4 (one warp per quadrant, 192 threads) is the small-footprint configuration: four co-resident CTAs per SM for
// the layers whose CTAs are short (cout <= 64 or tiny maps) - more MMA streams in flight hide the per-CTA phases.
__host__ __device__ constexpr int num_issuers(int T) { return PBT_MULTI_ISSUE ? T : 1; }
__host__ __device__ constexpr int conv_threads(int T, int EW) { return 32 * (1 + num_issuers(T) + EW); }

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

// bilinear x2, align_corners=True: source index / weight of destination coordinate `dst` (same arithmetic as
// upsample2x_kernel in elementwise.cu and as torch's area_pixel_compute_source_index)
__device__ __forceinline__ void up_src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}
// first low-res row/column touched by the haloed tile that starts at high-res coordinate `hr0` (may be negative)
__device__ __forceinline__ int up_origin(int hr0, int hr_size, int lr_size) {
  const float scale = hr_size > 1 ? (float)(lr_size - 1) / (float)(hr_size - 1) : 0.f;
  const int h = hr0 < 0 ? 0 : hr0;
  int i0 = (int)(scale * (float)h);
  if (i0 > lr_size - 1) i0 = lr_size - 1;
  return i0;
}

template <int T, int KB, int EW, bool PAIR>
// two (EW = 8) or four (EW = 4) co-resident CTAs per SM are essential (their MMA streams overlap): cap registers accordingly
__global__ void __launch_bounds__(EW == 8 ? kMaxThreads : 192, PBT_MULTI_ISSUE ? 1 : (EW == 8 ? 2 : 4))
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapP, const ConvKParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;
  uint8_t* sB = sA + (size_t)p.a_stages * p.a_stage_bytes;
  uint8_t* sL = sB + (size_t)p.b_stages * p.b_stage_bytes;  // low-res staging (upsample-on-load only), 2 stages
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sL + 2 * (size_t)p.l_stage_bytes);
  uint64_t* a_empty = a_full + p.a_stages;
  uint64_t* b_full = a_empty + p.a_stages;
  uint64_t* b_empty = b_full + p.b_stages;
  uint64_t* acc_full = b_empty + p.b_stages;
  uint64_t* l_full = acc_full + 1;
  uint64_t* l_empty = l_full + 2;
  uint64_t* a_land = l_empty + 2;  // normalise-on-load: raw tile landed (TMA), not yet normalised
  uint64_t* a_peer = a_land + 2;   // CTA-pair mode, leader only: the peer CTA's A stage / weight stage is ready
  uint64_t* b_peer = a_peer + 2;
  uint64_t* acc_empty = b_peer + 4;  // persistent CTAs: the epilogue has drained the accumulators of the previous unit
  uint64_t* acc_peer = acc_empty + 1;  // CTA-pair mode, leader only: the peer's accumulators are drained too
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_peer + 1);
  float* s_stats = reinterpret_cast<float*>(tmem_slot + 2);  // [8 warps][2][NC]
  const int TS = p.bt ? T : 1;                         // statistics slots per CTA: one per image in batch-tile mode
  float* s_head = s_stats + EW * 2 * p.NC * TS;        // [T][4 quadrants][32 lanes][3]
  float* s_norm = s_head + (EW == 8 ? 3 * 4 * 32 * 3 : 0);                    // [2][pre_c] scale / shift of this image

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int NI = num_issuers(T);
  constexpr int kEpi0 = 32 * (1 + NI);  // first epilogue thread

  // CTA-pair mode (PAIR): this CTA and its cluster peer (x-adjacent unit) run ONE M=256 cta_group::2 MMA stream issued
  // by the leader (rank 0).  Each CTA loads its own activation tile and HALF of the weight columns, so an SM reads
  // 4096 + 16*N bytes of operands per MMA instead of 4096 + 32*N (the shared-memory bound of the N = 64 layers).
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int NCb = PAIR ? p.NC / 2 : p.NC;  // weight columns staged by this CTA
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  // Persistent CTAs: the grid is one wave (SMs x co-resident CTAs) and every CTA walks the units
  // blockIdx.x, blockIdx.x + gridDim.x, ...  Barrier stages / parities run on counters that continue across units, so the
  // producer prefetches the next unit's first tiles while the current unit's last MMAs and epilogue run, and the
  // per-CTA launch gap + setup (~10-25 % of a short CTA's life, tools/conv_timeline.py) is paid once.
#define PBT_UNIT_GEOM(unit)                                  \
  const int ngrp = (unit) / tiles_per_img;                   \
  const int n = p.bt ? ngrp * T : ngrp;                      \
  const int rem = (unit) - ngrp * tiles_per_img;             \
  const int tyi = rem / p.tiles_x;                           \
  const int txi = rem - tyi * p.tiles_x;                     \
  const int x0 = txi * (p.bt ? 8 : 8 * T);                   \
  const int y0 = tyi * 16;
  const int ntaps = p.tp ? p.KH * (p.KWs / 2) : p.KH * p.KW;   // MMA groups per channel block (tap pairs: 4 per 7-tap row)
  const int ngroups = (ntaps + p.b_group - 1) / p.b_group;

#define PBT_STAMP(slot)                                                         \
  do {                                                                           \
    if (p.debug_buf) p.debug_buf[(long long)blockIdx.x * 8 + (slot)] = clock64(); \
  } while (0)
  if (threadIdx.x == 0) {
    PBT_STAMP(0);
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], (p.up || p.nrm) ? 32 * EW : 1);  // transform modes: every transform thread arrives
      mbar_init(&a_empty[i], NI);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], NI);
    }
    mbar_init(acc_full, NI);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&l_full[i], 1);
      mbar_init(&l_empty[i], 32 * EW);
      mbar_init(&a_land[i], 1);
      mbar_init(&a_peer[i], 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&b_peer[i], 1);
    mbar_init(acc_empty, 32 * EW);
    mbar_init(acc_peer, 1);
    fence_barrier_init();
    prefetch_tmap(&tmapA);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols);
    else tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  // programmatic dependent launch: barriers, tensor memory and descriptors were set up while the previous launch of the
  // stream was still draining; from here on global memory written by earlier launches is read
  pdl_sync();
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
     int bi = 0;   // running B-group counter (continues across units)
     int cbt = 0;  // running channel-block counter at the start of the unit
     for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, cbt += p.n_blk) {
      PBT_UNIT_GEOM(unit)
      (void)rem;
      // activation tile (or, with upsample-on-load, its low-res footprint) of channel block `cb`
      auto issue_act = [&](int cb) {
        const int c = cbt + cb;
        if (p.up) {
          const int sl = c & 1;
          mbar_wait(&l_empty[sl], ((uint32_t)(c >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&l_full[sl], (uint32_t)(p.blk_p * p.LBH * p.LBW * 16));
          tma_load_4d(sL + (size_t)sl * p.l_stage_bytes, &tmapA, &l_full[sl], up_origin(x0 - p.pad_l, p.W, p.LW) * 8,
                      up_origin(y0 - p.pad_t, p.H, p.LH), cb * p.blk_p, n);
        } else {
          const int sa = c % p.a_stages;
          mbar_wait(&a_empty[sa], ((uint32_t)(c / p.a_stages) & 1u) ^ 1u);
          uint64_t* bar = p.nrm ? &a_land[sa] : &a_full[sa];
          if (p.bt) {   // the same haloed tile of T consecutive images (images past the batch are zero-filled by TMA)
            mbar_arrive_expect_tx(bar, (uint32_t)(p.a_planes * p.BH * p.BW * 16 * T));
            for (int tt = 0; tt < T; ++tt)
              tma_load_4d(sA + (size_t)sa * p.a_stage_bytes + (size_t)tt * p.a_tile_bytes, &tmapA, bar, (x0 - p.pad_l) * 8,
                          y0 - p.pad_t, cb * p.blk_p, n + tt);
            return;
          }
          mbar_arrive_expect_tx(bar, (uint32_t)(p.a_planes * p.BH * p.BW * 16));
          if (cb < p.nblk0)
            tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, &tmapP, bar, (x0 - p.pad_l) * 8, y0 - p.pad_t, cb * p.blk_p, n);
          else
            tma_load_4d(sA + (size_t)sa * p.a_stage_bytes, &tmapA, bar, (x0 - p.pad_l) * 8, y0 - p.pad_t,
                        (cb - p.nblk0) * p.blk_p, n);
        }
      };
      issue_act(0);
      for (int cb = 0; cb < p.n_blk; ++cb) {
        // Block cb+1 is requested once the weight ring of block cb is primed (b_stages groups in flight): by then
        // the MMAs of block cb-1 have retired (its stage is free, no blocking wait here), and the load — plus the
        // upsample transform — overlaps almost a whole block of MMAs instead of starting when the ring drains.
        const int gpre = min(p.b_stages, ngroups - 1);
        const int pib = min(p.blk_p, p.Cp - cb * p.blk_p);
        const uint32_t chunk = (uint32_t)(pib * NCb * 16);
        const uint8_t* wsrc = p.wpack + (size_t)cb * ntaps * ((size_t)p.blk_p * p.NC * 16) + (size_t)crank * ntaps * chunk;
        for (int g = 0; g < ngroups; ++g, ++bi) {
          const int sb = bi % p.b_stages;
          const uint32_t pb = (uint32_t)(bi / p.b_stages) & 1u;
          const int tap0 = g * p.b_group;
          const int nt = min(p.b_group, ntaps - tap0);
          mbar_wait(&b_empty[sb], pb ^ 1u);
          mbar_arrive_expect_tx(&b_full[sb], chunk * (uint32_t)nt);
          // the taps of a group are contiguous in the packed weights: one bulk copy per group
          bulk_load_1d(sB + (size_t)sb * p.b_stage_bytes, wsrc + (size_t)tap0 * chunk, chunk * (uint32_t)nt, &b_full[sb]);
          if (g == gpre && cb + 1 < p.n_blk && p.a_stages > 1) issue_act(cb + 1);
        }
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
    if (PAIR && crank != 0) {
      // peer CTA of a pair: no MMAs here - forward "my stage is ready" to the leader's barriers, in the order the
      // leader consumes them
      int bi = 0, cbt = 0, it = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, cbt += p.n_blk, ++it) {
        if (it > 0) {  // this CTA's accumulators of the previous unit are drained -> tell the leader
          mbar_wait(acc_empty, (uint32_t)(it - 1) & 1u);
          if (leader) mbar_arrive_remote(acc_peer, 0);
        }
        for (int cb = 0; cb < p.n_blk; ++cb) {
          const int c = cbt + cb;
          const int sa = c % p.a_stages;
          mbar_wait(&a_full[sa], (uint32_t)(c / p.a_stages) & 1u);
          if (leader) mbar_arrive_remote(&a_peer[sa], 0);
          for (int g = 0; g < ngroups; ++g, ++bi) {
            const int sb = bi % p.b_stages;
            mbar_wait(&b_full[sb], (uint32_t)(bi / p.b_stages) & 1u);
            if (leader) mbar_arrive_remote(&b_peer[sb], 0);
          }
          __syncwarp();
        }
      }
    } else {
      const uint32_t plane_bytes = (uint32_t)(p.BH * p.BW * 16);
      const uint32_t row_bytes = (uint32_t)(p.BW * 16);
      const uint32_t b_kstride = (uint32_t)(NCb * 16);
      // descriptor words: lo = addr>>4 | (LBO>>4)<<16 ; hi = SBO>>4 | version(1)<<14
      // tap pairs: the second 8-element K group is the SAME plane one pixel (16 bytes) to the right -> LBO = 16 B
      const uint32_t a_lo_const = p.tp ? (1u << 16) : (((plane_bytes >> 4) & 0x3FFF) << 16);
      const uint32_t a_hi = ((row_bytes >> 4) & 0x3FFF) | (1u << 14);
      const uint32_t b_lo_const = ((b_kstride >> 4) & 0x3FFF) << 16;
      const uint32_t b_hi = (128u >> 4) | (1u << 14);
      const uint32_t a_kstep = (2u * plane_bytes) >> 4;  // two channel planes per K=16
      const uint32_t b_kstep = (2u * b_kstride) >> 4;
      const uint32_t acc_stride = (uint32_t)p.acc_stride;
      const uint32_t a_tstep = p.a_tile16;
      const bool ws_b = p.ws_b != 0;
      const uint32_t idesc = p.idesc;
      int bi = 0, cbt = 0, it = 0;
      for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, cbt += p.n_blk, ++it) {
      if (it > 0) {  // the accumulators are overwritten: wait until the epilogue has read the previous unit's
        mbar_wait(acc_empty, (uint32_t)(it - 1) & 1u);
        if (PAIR) mbar_wait_cluster(acc_peer, (uint32_t)(it - 1) & 1u);
        tc_fence_after();
      }
      for (int cb = 0; cb < p.n_blk; ++cb) {
        const int sa = (cbt + cb) % p.a_stages;
        const uint32_t pa = (uint32_t)((cbt + cb) / p.a_stages) & 1u;
        mbar_wait(&a_full[sa], pa);
        if (PAIR) mbar_wait_cluster(&a_peer[sa], pa);
        tc_fence_after();
        if (cb == 0 && leader) PBT_STAMP(2);
        const uint32_t a_base = (smem_u32(sA + (size_t)sa * p.a_stage_bytes) >> 4) | a_lo_const;
        const int pib = min(p.blk_p, p.Cp - cb * p.blk_p);
        const int k16n = pib >> 1;
        const uint32_t chunk16 = (uint32_t)(pib * NCb);  // bytes/16 of one tap's weights (this CTA's columns)
        int dy = 0, dx = 0;
        for (int g = 0; g < ngroups; ++g, ++bi) {
          const int sb = bi % p.b_stages;
          const uint32_t pb = (uint32_t)(bi / p.b_stages) & 1u;
          const int nt = min(p.b_group, ntaps - g * p.b_group);
          mbar_wait(&b_full[sb], pb);
          if (PAIR) mbar_wait_cluster(&b_peer[sb], pb);
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
                    if (NI == 1 || t == ti) {
                      if (PAIR)
                        umma_f16_pair(tmem_base + (uint32_t)t * acc_stride,
                                      desc64(a_tap + (uint32_t)t * a_tstep + (uint32_t)k * a_kstep, a_hi),
                                      desc64(b_lo + (uint32_t)k * b_kstep, b_hi), idesc, (k == 0) ? first : 1u);
                      else if (T > 1 && NI == 1 && ws_b) {
                        // weight-stationary run over the T tiles: B is fetched from shared memory once per (tap, K step)
                        const uint32_t dt_ = tmem_base + (uint32_t)t * acc_stride;
                        const uint64_t ad = desc64(a_tap + (uint32_t)t * a_tstep + (uint32_t)k * a_kstep, a_hi);
                        const uint64_t bd = desc64(b_lo + (uint32_t)k * b_kstep, b_hi);
                        const uint32_t en = (k == 0) ? first : 1u;
                        if (t == 0) umma_f16_ws<0>(dt_, ad, bd, idesc, en);
                        else if (t == T - 1) umma_f16_ws<2>(dt_, ad, bd, idesc, en);
                        else umma_f16_ws<1>(dt_, ad, bd, idesc, en);
                      } else
                        umma_f16(tmem_base + (uint32_t)t * acc_stride,
                                 desc64(a_tap + (uint32_t)t * a_tstep + (uint32_t)k * a_kstep, a_hi),
                                 desc64(b_lo + (uint32_t)k * b_kstep, b_hi), idesc, (k == 0) ? first : 1u);
                    }
                }
              }
            }
            b_lo += chunk16;
            dx += p.dx_step;
            if (dx >= p.KWs) { dx = 0; ++dy; }
          }
          if (leader) {
            if (PAIR) umma_commit_pair(&b_empty[sb]);
            else umma_commit(&b_empty[sb]);
          }
          __syncwarp();
        }
        if (leader) {
          if (PAIR) umma_commit_pair(&a_empty[sa]);
          else umma_commit(&a_empty[sa]);
        }
        __syncwarp();
      }
      if (leader) {
        if (PAIR) umma_commit_pair(acc_full);
        else umma_commit(acc_full);
        if (ti == 0) PBT_STAMP(3);
      }
      __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (8 warps)
    // TMEM lanes are reachable per warp quadrant (warp id % 4); the two warps of a quadrant take alternate
    // 16-column chunks of the accumulators, which halves the epilogue latency of a tile.
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int ew = warp - (1 + NI);
    const int half = ew >> 2;
    const int NC = p.NC;
    const int dt = p.dt;
    const bool do_stats = p.stats_partial != nullptr;
    float* my_stats = s_stats + (size_t)ew * 2 * NC * TS;   // [TS][2][NC]
    int cbt = 0, it = 0;
    for (int unit = blockIdx.x; unit < p.n_units; unit += gridDim.x, cbt += p.n_blk, ++it) {
    PBT_UNIT_GEOM(unit)
    if (do_stats) {
      for (int i = lane; i < 2 * NC * TS; i += 32) my_stats[i] = 0.f;
      __syncwarp();
    }
    const int r = q * 32 + lane;
    const int ty = r >> 3, tx = r & 7;
    const int y = y0 + ty;
    const long long plane_px = (long long)p.H * p.W;
    const int col_of_lane = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

    if (p.up) {
      // ---- upsample-on-load: while the MMAs run, these warps build the haloed A tiles of the 2x bilinear
      // (align_corners=True, reference src/models/generator.py:13) upsampled input from the low-res staging copy.
      const int et = threadIdx.x - kEpi0;  // 0..255
      const float sy = p.H > 1 ? (float)(p.LH - 1) / (float)(p.H - 1) : 0.f;
      const float sx = p.W > 1 ? (float)(p.LW - 1) / (float)(p.W - 1) : 0.f;
      const int ly0 = up_origin(y0 - p.pad_t, p.H, p.LH), lx0 = up_origin(x0 - p.pad_l, p.W, p.LW);
      const int npos = p.BH * p.BW;
      // each thread owns up to two haloed-tile positions; their tap offsets / weights are fixed for the whole CTA
      int tap00[2], tap01[2], tap10[2], tap11[2], dsto[2];
      float w00[2], w01[2], w10[2], w11[2];
      bool has[2], ins[2];
#pragma unroll
      for (int qq = 0; qq < 2; ++qq) {
        const int pos = et + qq * 32 * EW;
        has[qq] = pos < npos;
        const int r = pos / p.BW, c = pos - r * p.BW;
        const int Y = y0 - p.pad_t + r, X = x0 - p.pad_l + c;
        ins[qq] = has[qq] && Y >= 0 && Y < p.H && X >= 0 && X < p.W;
        int i0 = 0, i1 = 0, j0 = 0, j1 = 0;
        float wy = 0.f, wx = 0.f;
        if (ins[qq]) {
          up_src_index(Y, sy, p.LH, i0, i1, wy);
          up_src_index(X, sx, p.LW, j0, j1, wx);
          i0 -= ly0; i1 -= ly0; j0 -= lx0; j1 -= lx0;
        }
        tap00[qq] = (i0 * p.LBW + j0) * 16; tap01[qq] = (i0 * p.LBW + j1) * 16;
        tap10[qq] = (i1 * p.LBW + j0) * 16; tap11[qq] = (i1 * p.LBW + j1) * 16;
        w00[qq] = (1.f - wy) * (1.f - wx); w01[qq] = (1.f - wy) * wx;
        w10[qq] = wy * (1.f - wx);         w11[qq] = wy * wx;
        dsto[qq] = pos * 16;
      }
      const int l_plane = p.LBH * p.LBW * 16, a_plane = npos * 16;
      if (p.up_nrm) {   // the first nblk0 channel blocks of the low-res input are RAW conv outputs: their norm + activation table
        for (int i = et; i < p.pre_c; i += 32 * EW) {
          const int nn = min(n, p.n_img - 1);
          s_norm[i] = __ldg(&p.pre_scale[(long long)nn * p.pre_c + i]);
          s_norm[p.pre_c + i] = __ldg(&p.pre_shift[(long long)nn * p.pre_c + i]);
        }
        asm volatile("bar.sync 3, %0;" ::"r"(32 * EW) : "memory");
      }
      for (int cb = 0; cb < p.n_blk; ++cb) {
        const int c = cbt + cb;
        const int sa = c % p.a_stages, sl = c & 1;
        mbar_wait(&l_full[sl], (uint32_t)(c >> 1) & 1u);
        mbar_wait(&a_empty[sa], ((uint32_t)(c / p.a_stages) & 1u) ^ 1u);
        const uint8_t* src = sL + (size_t)sl * p.l_stage_bytes;
        uint8_t* dstA = sA + (size_t)sa * p.a_stage_bytes;
        if (p.up_nrm && cb < p.nblk0) {
          // normalise + activate the staged LOW-RES tile in place (each low-res pixel once, ~4x fewer than the interpolated
          // positions), then interpolate: bilinear(act(x*s+t)) exactly as the reference upsamples the activated tensor
          // (src/models/generator.py:204-206 feeding :13).  All transform threads must see the finished tile: named barrier.
          uint8_t* stg = sL + (size_t)sl * p.l_stage_bytes;
          const int lpos = p.LBH * p.LBW;
          for (int i = et; i < lpos * p.blk_p; i += 32 * EW) {
            const int pln = i / lpos, pos = i - pln * lpos;
            const int ch = (cb * p.blk_p + pln) * 8;
            uint4* ptr = reinterpret_cast<uint4*>(stg + (size_t)pln * l_plane + (size_t)pos * 16);
            float v[8];
            unpack8_rt(dt, *ptr, v);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float t = fmaf(v[k], s_norm[ch + k], s_norm[p.pre_c + ch + k]);
              if (p.pre_act == PBT_ACT_RELU) t = fmaxf(t, 0.f);
              else if (p.pre_act == PBT_ACT_LEAKY02) t = t > 0.f ? t : 0.2f * t;
              v[k] = t;
            }
            *ptr = pack8_rt(dt, v);
          }
          asm volatile("bar.sync 3, %0;" ::"r"(32 * EW) : "memory");
        }
#pragma unroll
        for (int qq = 0; qq < 2; ++qq) {
          if (!has[qq]) continue;
          if (!ins[qq]) {  // conv zero padding outside the (upsampled) image
            for (int pln = 0; pln < p.blk_p; ++pln)
              *reinterpret_cast<uint4*>(dstA + (size_t)pln * a_plane + dsto[qq]) = make_uint4(0u, 0u, 0u, 0u);
            continue;
          }
#pragma unroll 2
          for (int pln = 0; pln < p.blk_p; ++pln) {
            const uint8_t* pb = src + (size_t)pln * l_plane;
            const uint4 ua = *reinterpret_cast<const uint4*>(pb + tap00[qq]);
            const uint4 ub = *reinterpret_cast<const uint4*>(pb + tap01[qq]);
            const uint4 uc = *reinterpret_cast<const uint4*>(pb + tap10[qq]);
            const uint4 ud = *reinterpret_cast<const uint4*>(pb + tap11[qq]);
            if (dt == 1) {
              // fp16 operands: blend with packed half2 FMAs (16 HFMA2 per chunk instead of ~75 fp32-path instructions;
              // the extra half-precision roundings are ~2^-11 relative, far inside the forward tolerance)
              const __half2 h00 = __float2half2_rn(w00[qq]), h01 = __float2half2_rn(w01[qq]);
              const __half2 h10 = __float2half2_rn(w10[qq]), h11 = __float2half2_rn(w11[qq]);
              const __half2* pa = reinterpret_cast<const __half2*>(&ua);
              const __half2* pb2 = reinterpret_cast<const __half2*>(&ub);
              const __half2* pc = reinterpret_cast<const __half2*>(&uc);
              const __half2* pd = reinterpret_cast<const __half2*>(&ud);
              uint4 o;
              __half2* po = reinterpret_cast<__half2*>(&o);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                po[k] = __hfma2(h11, pd[k], __hfma2(h10, pc[k], __hfma2(h01, pb2[k], __hmul2(h00, pa[k]))));
              *reinterpret_cast<uint4*>(dstA + (size_t)pln * a_plane + dsto[qq]) = o;
            } else {
              float a[8], b[8], cc[8], d[8], rr[8];
              unpack8_rt(dt, ua, a);
              unpack8_rt(dt, ub, b);
              unpack8_rt(dt, uc, cc);
              unpack8_rt(dt, ud, d);
#pragma unroll
              for (int k = 0; k < 8; ++k) rr[k] = w00[qq] * a[k] + w01[qq] * b[k] + w10[qq] * cc[k] + w11[qq] * d[k];
              *reinterpret_cast<uint4*>(dstA + (size_t)pln * a_plane + dsto[qq]) = pack8_rt(dt, rr);
            }
          }
        }
        fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        mbar_arrive(&a_full[sa]);
        mbar_arrive(&l_empty[sl]);
      }
    }
    if (p.nrm) {
      // ---- normalise-on-load: the raw output of the previous conv lands in the A stage by TMA; these warps apply
      // its InstanceNorm + activation in place (y = act(x*scale + shift), reference src/models/generator.py:36-40,
      // 204-206) before the tensor core reads it, so the normalised tensor never round-trips HBM.  Pixels outside
      // the image stay zero (the conv pads the NORMALISED tensor with zeros).
      const int et = threadIdx.x - kEpi0;
      for (int i = et; i < p.pre_c; i += 32 * EW) {
        const int nn = min(n, p.n_img - 1);  // (a pair's padding CTA past the last unit loads zeros and stores nothing)
        s_norm[i] = __ldg(&p.pre_scale[(long long)nn * p.pre_c + i]);
        s_norm[p.pre_c + i] = __ldg(&p.pre_shift[(long long)nn * p.pre_c + i]);
      }
      asm volatile("bar.sync 3, %0;" ::"r"(32 * EW) : "memory");
      const int npos = p.BH * p.BW;
      constexpr int NQ = EW == 8 ? 3 : 4;  // tile positions per thread: BH*BW <= NQ * 32 * EW (host-checked)
      int dsto[NQ];
      bool ins[NQ];
#pragma unroll
      for (int qq = 0; qq < NQ; ++qq) {
        const int pos = et + qq * 32 * EW;
        const int r = pos / p.BW, c = pos - r * p.BW;
        const int Y = y0 - p.pad_t + r, X = x0 - p.pad_l + c;
        ins[qq] = pos < npos && Y >= 0 && Y < p.H && X >= 0 && X < p.W;
        dsto[qq] = pos * 16;
      }
      const int a_plane = npos * 16;
      for (int cb = 0; cb < p.n_blk; ++cb) {
        const int sa = (cbt + cb) % p.a_stages;
        mbar_wait(&a_land[sa], (uint32_t)((cbt + cb) / p.a_stages) & 1u);
        if (cb < p.nblk0) {
          uint8_t* tile = sA + (size_t)sa * p.a_stage_bytes;
          for (int pln = 0; pln < p.blk_p; ++pln) {
            const int ch = (cb * p.blk_p + pln) * 8;
            float sc[8], sh[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              sc[k] = s_norm[ch + k];
              sh[k] = s_norm[p.pre_c + ch + k];
            }
            if (dt == PBT_FP16) {
              // fp16 operands: packed half2 math (4 HFMA2 + 4 HMNMX2 per 8 channels instead of ~30 fp32-path
              // instructions); rounding scale/shift to fp16 adds ~2^-11 relative error, far inside the forward tolerance
              __half2 sc2[4], sh2[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                sc2[j] = __floats2half2_rn(sc[2 * j], sc[2 * j + 1]);
                sh2[j] = __floats2half2_rn(sh[2 * j], sh[2 * j + 1]);
              }
              const __half2 zero2 = __float2half2_rn(0.f), leak2 = __float2half2_rn(0.2f);
#pragma unroll
              for (int qq = 0; qq < NQ; ++qq) {
                if (!ins[qq]) continue;
                uint4* ptr = reinterpret_cast<uint4*>(tile + (size_t)pln * a_plane + dsto[qq]);
                uint4 u = *ptr;
                __half2* hp = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  __half2 t = __hfma2(hp[j], sc2[j], sh2[j]);
                  if (p.pre_act == PBT_ACT_RELU) t = __hmax2(t, zero2);
                  else if (p.pre_act == PBT_ACT_LEAKY02) t = __hmax2(t, __hmul2(t, leak2));
                  hp[j] = t;
                }
                *ptr = u;
              }
              continue;
            }
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq) {
              if (!ins[qq]) continue;
              uint4* ptr = reinterpret_cast<uint4*>(tile + (size_t)pln * a_plane + dsto[qq]);
              float v[8];
              unpack8_rt(dt, *ptr, v);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                float t = fmaf(v[k], sc[k], sh[k]);
                if (p.pre_act == PBT_ACT_RELU) t = fmaxf(t, 0.f);
                else if (p.pre_act == PBT_ACT_LEAKY02) t = t > 0.f ? t : 0.2f * t;
                v[k] = t;
              }
              *ptr = pack8_rt(dt, v);
            }
          }
        }
        fence_proxy_async();
        mbar_arrive(&a_full[sa]);
      }
    }
    mbar_wait_backoff(acc_full, (uint32_t)it & 1u, 256);
    tc_fence_after();
    if (threadIdx.x == kEpi0) PBT_STAMP(4);

    // Feature bits live in a register: testing `p.<field>` inside the column loop costs one dependent constant-bank load
    // + uniform branch per feature per iteration (measured ~1000 cycles per 16-column chunk, 360 without them).
    enum : uint32_t { kFBias = 1, kFRelu = 2, kFLeaky = 4, kFAffine = 8, kFMask = 16, kFAddend = 32, kFOut32 = 64,
                      kFOut16 = 128, kFHead = 256, kFStats = 512, kFWindow = 1024, kFGeneric = 0x80000000u };
    uint32_t feat = (p.bias ? kFBias : 0u) | (p.act == PBT_ACT_RELU ? kFRelu : 0u) | (p.act == PBT_ACT_LEAKY02 ? kFLeaky : 0u) |
                    (p.post_scale ? kFAffine : 0u) | (p.mask ? kFMask : 0u) | (p.addend32 ? kFAddend : 0u) |
                    (p.out32 ? kFOut32 : 0u) | (p.out ? kFOut16 : 0u) | (p.head_w ? kFHead : 0u) | (do_stats ? kFStats : 0u) |
                    ((p.VH < p.H || p.VW < p.W) ? kFWindow : 0u);
    asm volatile("" : "+r"(feat));  // keep it a register value, not re-derived constant loads
    // The column loop is instantiated once per common feature set (compile-time mask) plus a generic copy (run-time
    // mask): the fully generic body is ~13 KB of mostly skipped code and ran at ~1000 cycles per 16-column chunk
    // (instruction fetch across the skipped regions), a specialised body at ~400.
    auto epi_loop = [&](auto cf) {
      constexpr uint32_t CF = decltype(cf)::value;
      const uint32_t fm = CF == kFGeneric ? feat : CF;
      // statistics variants run the tile loop innermost: one cross-lane column reduction per 16-column chunk, not per tile
      constexpr bool kColOuter = CF != kFGeneric && (CF & kFStats) != 0 && (CF & kFHead) == 0;
      // one 16-column chunk of tile t: accumulator -> registers -> fused pointwise work -> stores
      auto chunk = [&](int t, int c0, float (&hd)[3], float* sacc) {
        const int x = p.bt ? x0 + tx : x0 + 8 * t + tx;
        const int nt = p.bt ? n + t : n;                      // image of tile t
        float* st_t = my_stats + (p.bt ? t * 2 * NC : 0);
        const bool valid = (y < p.H) && (x < p.W) && (nt < p.n_img);
        const long long pix = (long long)y * p.W + x;
        uint32_t raw[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * p.acc_stride + c0), raw);
        tmem_ld_wait();
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
        if (fm & kFBias) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += __ldg(&p.bias[c0 + i]);
        }
        if (fm & kFRelu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        } else if (fm & kFLeaky) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
        }
        if (fm & kFAffine) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], __ldg(&p.post_scale[c0 + i]), __ldg(&p.post_shift[c0 + i]));
        }
        if (fm & kFWindow) {   // (generic variant only) outputs outside the valid window are zero for every consumer
          if (!(y < p.VH && x < p.VW)) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
        }
        if (valid) {
          if (fm & kFMask) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const uint4 m = *reinterpret_cast<const uint4*>(
                  p.mask + 2 * ((long long)nt * p.mask_img_stride + ((long long)(c0 / 8 + hh) * plane_px + pix) * 8));
              float mf[8];
              unpack8_rt(dt, m, mf);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[hh * 8 + i] = mf[i] > 0.f ? v[hh * 8 + i] : 0.f;
            }
          }
          if (fm & kFAddend) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const float4* ap = reinterpret_cast<const float4*>(
                  p.addend32 + (((long long)nt * (NC / 8) + (c0 / 8 + hh)) * plane_px + pix) * 8);
              const float4 a0 = ap[0], a1 = ap[1];
              v[hh * 8 + 0] += a0.x; v[hh * 8 + 1] += a0.y; v[hh * 8 + 2] += a0.z; v[hh * 8 + 3] += a0.w;
              v[hh * 8 + 4] += a1.x; v[hh * 8 + 5] += a1.y; v[hh * 8 + 6] += a1.z; v[hh * 8 + 7] += a1.w;
            }
          }
          if (fm & kFOut32) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              float4* op = reinterpret_cast<float4*>(p.out32 + (((long long)nt * (NC / 8) + (c0 / 8 + hh)) * plane_px + pix) * 8);
              op[0] = make_float4(v[hh * 8 + 0], v[hh * 8 + 1], v[hh * 8 + 2], v[hh * 8 + 3]);
              op[1] = make_float4(v[hh * 8 + 4], v[hh * 8 + 5], v[hh * 8 + 6], v[hh * 8 + 7]);
            }
          }
        }
        if (fm & kFOut16) {
          // round to the storage type; statistics and the head see the rounded values
          const uint4 u0 = pack8_rt(dt, v), u1 = pack8_rt(dt, v + 8);
          if (valid) {
            uint8_t* ob = p.out + 2 * ((long long)nt * p.out_img_stride + ((long long)(c0 / 8) * plane_px + pix) * 8);
            *reinterpret_cast<uint4*>(ob) = u0;
            *reinterpret_cast<uint4*>(ob + 2 * plane_px * 8) = u1;
          }
          unpack8_rt(dt, u0, v);
          unpack8_rt(dt, u1, v + 8);
        }
        if (fm & kFHead) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            hd[0] = fmaf(v[i], __ldg(&p.head_w[c0 + i]), hd[0]);
            hd[1] = fmaf(v[i], __ldg(&p.head_w[NC + c0 + i]), hd[1]);
            hd[2] = fmaf(v[i], __ldg(&p.head_w[2 * NC + c0 + i]), hd[2]);
          }
        }
        if (fm & kFStats) {
          if (kColOuter) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float sv = valid ? v[i] : 0.f;
              sacc[i] += sv;
              sacc[16 + i] = fmaf(sv, sv, sacc[16 + i]);
            }
          } else {
            float s[16], s2[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              s[i] = valid ? v[i] : 0.f;
              s2[i] = s[i] * s[i];
            }
            const float cs = warp_colsum16(s, lane);
            const float cs2 = warp_colsum16(s2, lane);
            if ((lane & 1) == 0) {
              st_t[c0 + col_of_lane] += cs;
              st_t[NC + c0 + col_of_lane] += cs2;
            }
          }
        }
      };
      if constexpr (kColOuter) {
        for (int c0 = half * 16; c0 < NC; c0 += 4 * EW) {
          float sacc[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) sacc[i] = 0.f;
          float hd[3] = {0.f, 0.f, 0.f};
#pragma unroll
          for (int t = 0; t < T; ++t) {
            chunk(t, c0, hd, sacc);
            if (p.bt || t == T - 1) {   // batch tiles: every tile is another image -> its own statistics slot
              const float cs = warp_colsum16(sacc, lane);
              const float cs2 = warp_colsum16(sacc + 16, lane);
              float* st_t = my_stats + (p.bt ? t * 2 * NC : 0);
              if ((lane & 1) == 0) {
                st_t[c0 + col_of_lane] += cs;
                st_t[NC + c0 + col_of_lane] += cs2;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) sacc[i] = 0.f;
            }
          }
        }
      } else {
        for (int t = 0; t < T; ++t) {
          float hd[3] = {0.f, 0.f, 0.f};
          for (int c0 = half * 16; c0 < NC; c0 += 4 * EW) chunk(t, c0, hd, nullptr);
          if (!(fm & kFHead)) continue;
          if (EW == 8) {  // combine the two column halves of the fused 1x1 head
            float* hs = s_head + ((size_t)(t * 4 + q) * 32 + lane) * 3;
            if (half == 1) {
              hs[0] = hd[0];
              hs[1] = hd[1];
              hs[2] = hd[2];
            }
            asm volatile("bar.sync 2, %0;" ::"r"(32 * EW) : "memory");
            if (half == 0) {
              hd[0] += hs[0];
              hd[1] += hs[1];
              hd[2] += hs[2];
            }
          }
          const int x = p.bt ? x0 + tx : x0 + 8 * t + tx;
          const int nt = p.bt ? n + t : n;
          if (y < p.H && x < p.W && half == 0 && nt < p.n_img) {
            float h0 = hd[0] + __ldg(&p.head_b[0]), h1 = hd[1] + __ldg(&p.head_b[1]), h2 = hd[2] + __ldg(&p.head_b[2]);
            if (p.head_tanh) {
              h0 = tanhf(h0);
              h1 = tanhf(h1);
              h2 = tanhf(h2);
            }
            float* ho = p.head_out + (long long)nt * 3 * plane_px + (long long)y * p.W + x;
            ho[0] = h0;
            ho[plane_px] = h1;
            ho[2 * plane_px] = h2;
          }
        }
      }
    };
    switch (feat) {
      case kFOut16 | kFStats: epi_loop(std::integral_constant<uint32_t, kFOut16 | kFStats>{}); break;
      case kFBias | kFRelu | kFOut16: epi_loop(std::integral_constant<uint32_t, kFBias | kFRelu | kFOut16>{}); break;
      case kFBias | kFRelu | kFAffine | kFOut16: epi_loop(std::integral_constant<uint32_t, kFBias | kFRelu | kFAffine | kFOut16>{}); break;
      case kFBias | kFRelu | kFHead: epi_loop(std::integral_constant<uint32_t, kFBias | kFRelu | kFHead>{}); break;
      case kFBias | kFRelu | kFHead | kFOut16: epi_loop(std::integral_constant<uint32_t, kFBias | kFRelu | kFHead | kFOut16>{}); break;
      case kFBias | kFRelu | kFOut16 | kFStats: epi_loop(std::integral_constant<uint32_t, kFBias | kFRelu | kFOut16 | kFStats>{}); break;
      case kFOut16: epi_loop(std::integral_constant<uint32_t, kFOut16>{}); break;
      case kFMask | kFOut16: epi_loop(std::integral_constant<uint32_t, kFMask | kFOut16>{}); break;
      case kFMask | kFAddend | kFOut32: epi_loop(std::integral_constant<uint32_t, kFMask | kFAddend | kFOut32>{}); break;
      default: epi_loop(std::integral_constant<uint32_t, kFGeneric>{}); break;
    }
    // this thread's accumulator reads are complete: the MMA warp may start the next unit
    tc_fence_before();
    mbar_arrive(acc_empty);
    if (do_stats) {
      asm volatile("bar.sync 1, %0;" ::"r"(32 * EW) : "memory");
      const int e = threadIdx.x - kEpi0;  // 0..255
      for (int ts = 0; ts < TS; ++ts) {
        float* dst = p.stats_partial + ((long long)(n + ts) * tiles_per_img + rem) * 2 * NC;
        for (int i = e; i < 2 * NC && n + ts < p.n_img; i += 32 * EW) {
          float acc = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < EW; ++w8) acc += s_stats[(w8 * TS + ts) * 2 * NC + i];
          dst[i] = acc;
        }
      }
      if (unit + (int)gridDim.x < p.n_units)  // the scratch is zeroed again for the next unit: everyone has read it
        asm volatile("bar.sync 1, %0;" ::"r"(32 * EW) : "memory");
    }
    }  // unit loop
  }

  if (threadIdx.x == kEpi0) PBT_STAMP(5);
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs are done with each other's shared memory and barriers
  if (warp == 1) {
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
  if (threadIdx.x == 0) PBT_STAMP(6);
}

static int pow2_cols(int c) {
  int v = 32;
  while (v < c) v <<= 1;
  return v;
}

template <int T, int KB, int EW>
static int launch_conv(const CUtensorMap& tmap, const CUtensorMap& tmapP, const ConvKParams& p, int grid, uint32_t smem_bytes,
                       cudaStream_t stream) {
  PBT_CUDA_CHECK(cudaFuncSetAttribute(conv_igemm_kernel<T, KB, EW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  pbt::launch(conv_igemm_kernel<T, KB, EW, false>, grid, conv_threads(T, EW), smem_bytes, stream, tmap, tmapP, p);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

// CTA-pair configuration: clusters of two x-adjacent units (grid rounded up to even; a padding CTA loads zeros).
template <int T, int KB, int EW>
static int launch_conv_pair(const CUtensorMap& tmap, const CUtensorMap& tmapP, const ConvKParams& p, int grid, uint32_t smem_bytes,
                            cudaStream_t stream) {
  auto kfn = conv_igemm_kernel<T, KB, EW, true>;
  PBT_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);   // even: the unit count and the SM count are
  cfg.blockDim = dim3((unsigned)conv_threads(T, EW), 1, 1);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  PBT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kfn, tmap, tmapP, p));
  return PBT_OK;
}

template <int T>
static int launch_conv_kb(int kb, int ew, const CUtensorMap& tmap, const CUtensorMap& tmapP, const ConvKParams& p, int grid, uint32_t smem,
                          cudaStream_t s) {
  switch (kb) {
    case 1: return ew == 4 ? launch_conv<T, 1, 4>(tmap, tmapP, p, grid, smem, s) : launch_conv<T, 1, 8>(tmap, tmapP, p, grid, smem, s);
    case 2: return ew == 4 ? launch_conv<T, 2, 4>(tmap, tmapP, p, grid, smem, s) : launch_conv<T, 2, 8>(tmap, tmapP, p, grid, smem, s);
    default: return ew == 4 ? launch_conv<T, 4, 4>(tmap, tmapP, p, grid, smem, s) : launch_conv<T, 4, 8>(tmap, tmapP, p, grid, smem, s);
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
  const int up = d->upsample2x ? 1 : 0;  // input is the low-res tensor; the conv sees its bilinear x2 upsample
  const int has_pre = d->pre.ptr != nullptr;
  PBT_REQUIRE(d->in.ptr || has_pre, "conv: no input tensor");
  // `pre` (optional) supplies the first pre.c input channels, normalised + activated on load; `in` the rest
  pbt_act_t in = d->in;
  if (!in.ptr) {  // all channels come from `pre`: keep the geometry, zero channels
    in = d->pre;
    in.c = 0;
  }
  PBT_REQUIRE(in.ptr && aligned16(in.ptr), "conv: input pointer null or not 16-byte aligned");
  PBT_REQUIRE(in.n > 0 && in.h > 0 && in.w > 0, "conv: empty input");
  PBT_REQUIRE(in.c % 16 == 0 && (in.c > 0 || has_pre), "conv: cin must be a multiple of 16");
  if (has_pre) {
    PBT_REQUIRE(!up, "conv: normalise-on-load cannot be combined with upsample-on-load");
    PBT_REQUIRE(aligned16(d->pre.ptr) && d->pre.n == in.n && d->pre.h == in.h && d->pre.w == in.w && d->pre.img_stride % 8 == 0,
                "conv: `pre` tensor shape mismatch");
    PBT_REQUIRE(d->pre.c > 0 && d->pre.c % d->blk_c == 0 && d->pre.c <= 256, "conv: pre.c must be a multiple of blk_c, <= 256");
    PBT_REQUIRE(d->pre_scale && d->pre_shift, "conv: `pre` needs pre_scale / pre_shift");
  }
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
  p.n_img = in.n; p.H = in.h << up; p.W = in.w << up;
  p.up = up; p.LH = in.h; p.LW = in.w;
  p.Cp = in.c / 8 + (has_pre ? d->pre.c / 8 : 0);
  p.blk_p = d->blk_c / 8;
  p.nrm = has_pre; p.nblk0 = has_pre ? d->pre.c / d->blk_c : 0; p.pre_c = has_pre ? d->pre.c : 0;
  p.pre_act = d->pre_act; p.pre_scale = d->pre_scale; p.pre_shift = d->pre_shift;
  if (up && d->up_raw_channels > 0) {
    // upsample-on-load over an input whose FIRST up_raw_channels channels are the raw output of the previous conv
    PBT_REQUIRE(!has_pre && d->pre_scale && d->pre_shift && d->up_raw_channels % d->blk_c == 0 && d->up_raw_channels <= in.c &&
                    d->up_raw_channels <= 256, "conv: up_raw_channels needs pre_scale / pre_shift and a multiple of blk_c <= cin, 256");
    p.up_nrm = 1; p.nblk0 = d->up_raw_channels / d->blk_c; p.pre_c = d->up_raw_channels;
  }
  p.n_blk = ceil_div(p.Cp, p.blk_p);
  p.KH = d->kh; p.KW = d->kw; p.pad_t = d->pad_t; p.pad_l = d->pad_l;
  const int T = d->tiles_per_cta;
  const int tp = d->tap_pairs ? 1 : 0;
  PBT_REQUIRE(!tp || (!up && !has_pre && !d->cta_pair && in.c == 16 && d->blk_c == 16 && !d->batch_tiles),
              "conv: tap_pairs needs a 16-channel input (<= 8 real channels), blk_c 16, and excludes upsample2x / pre / cta_pair / batch_tiles");
  p.tp = tp;
  const int bt = d->batch_tiles ? 1 : 0;
  PBT_REQUIRE(!bt || (!up && !has_pre && !d->cta_pair && T >= 2), "conv: batch_tiles needs tiles_per_cta >= 2 and excludes upsample2x / pre / cta_pair");
  p.bt = bt;
  p.NC = d->cout;
  p.dt = d->dtype;
  p.acc_stride = (int)round_up((uint32_t)p.NC, 32);
  p.tmem_cols = pow2_cols(T * p.acc_stride);
  PBT_REQUIRE(p.tmem_cols <= 512, "conv: tiles_per_cta*cout exceeds tensor memory (512 columns)");
  p.KWs = tp ? (p.KW + 1) / 2 * 2 : p.KW;
  p.dx_step = tp ? 2 : 1;
  p.a_planes = tp ? 1 : p.blk_p;
  p.BW = (bt ? 8 : 8 * T) + p.KWs - 1;
  p.BH = 16 + p.KH - 1;
  PBT_REQUIRE(p.BW <= 32, "conv: haloed tile wider than 32 pixels (reduce tiles_per_cta)");
  p.tiles_x = ceil_div(p.W, bt ? 8 : 8 * T);
  p.tiles_y = ceil_div(p.H, 16);
  const int pair = d->cta_pair ? 1 : 0;
  if (pair) {
    // instantiated pair configurations: default footprint (blk_c 32, T 1|2|3) and small footprint (blk_c 16, T 2)
    PBT_REQUIRE(p.NC % 32 == 0, "conv: cta_pair needs cout % 32 == 0");
    if (d->ctas_per_sm == 4)
      PBT_REQUIRE(d->blk_c == 16 && T == 2 && p.NC <= 64, "conv: cta_pair + ctas_per_sm=4 needs blk_c 16, tiles_per_cta 2, cout <= 64");
    else
      PBT_REQUIRE(d->blk_c == 32 && T >= 1 && T <= 3, "conv: cta_pair needs blk_c 32");
  }
  p.idesc = make_idesc_f16(pair ? 256 : 128, p.NC, d->dtype == PBT_BF16 ? 1 : 0, 0, 0);
  p.a_tile_bytes = round_up((uint32_t)(p.a_planes * p.BH * p.BW * 16), 128);
  p.a_stage_bytes = bt ? (uint32_t)T * p.a_tile_bytes : p.a_tile_bytes;
  p.a_tile16 = bt ? p.a_tile_bytes / 16 : 8;
  p.LBH = (p.BH + 1) / 2 + 2;
  p.LBW = (p.BW + 1) / 2 + 2;
  p.l_stage_bytes = up ? round_up((uint32_t)(p.blk_p * p.LBH * p.LBW * 16), 128) : 0;
  PBT_REQUIRE(!up || (d->kh == 3 && d->kw == 3 && d->pad_t == 1 && d->pad_l == 1), "conv: upsample-on-load supports 3x3 pad 1");
  p.a_stages = p.n_blk > 1 ? 2 : 1;
  p.wpack = static_cast<const uint8_t*>(d->wpack);
  p.bias = d->bias; p.act = d->act; p.post_scale = d->post_scale; p.post_shift = d->post_shift;
  PBT_REQUIRE((d->post_scale == nullptr) == (d->post_shift == nullptr), "conv: post_scale/post_shift must come together");
  if (d->mask.ptr) {
    PBT_REQUIRE(d->mask.c >= d->cout && d->mask.h == p.H && d->mask.w == p.W && d->mask.n == in.n && aligned16(d->mask.ptr),
                "conv: mask shape mismatch");
    p.mask = static_cast<const uint8_t*>(d->mask.ptr);
    p.mask_img_stride = d->mask.img_stride;
  }
  p.addend32 = d->addend32; p.out32 = d->out32;
  if (d->out.ptr) {
    PBT_REQUIRE(d->out.c >= d->cout && d->out.h == p.H && d->out.w == p.W && d->out.n == in.n && aligned16(d->out.ptr),
                "conv: output shape mismatch");
    p.out = static_cast<uint8_t*>(d->out.ptr);
    p.out_img_stride = d->out.img_stride;
  }
  PBT_REQUIRE(d->valid_h >= 0 && d->valid_h <= p.H && d->valid_w >= 0 && d->valid_w <= p.W, "conv: valid window exceeds the map");
  p.VH = d->valid_h ? d->valid_h : p.H;
  p.VW = d->valid_w ? d->valid_w : p.W;
  p.stats_partial = d->stats_partial;
  p.head_w = d->head_w; p.head_b = d->head_b; p.head_out = d->head_out; p.head_tanh = d->head_tanh;
  PBT_REQUIRE(!d->head_w || (d->head_b && d->head_out), "conv: head needs head_b and head_out");
  p.debug_flags = d->debug_flags;
  p.debug_buf = static_cast<long long*>(d->debug_buf);

  // shared memory budget: A ring + B ring (groups of taps) + barriers + tmem slot + stats scratch.
  // Aim at two co-resident CTAs per SM (one CTA's epilogue/prologue overlaps the other's main loop).
  // ctas_per_sm = 4 requests the small-footprint configuration (4 epilogue warps, <= 128 TMEM columns, <= 56 KB of
  // shared memory -> four co-resident CTAs); shapes that do not fit it run the default configuration.
  PBT_REQUIRE(d->ctas_per_sm == 0 || d->ctas_per_sm == 2 || d->ctas_per_sm == 4, "conv: ctas_per_sm must be 0, 2 or 4");
  const uint32_t a_total = (uint32_t)p.a_stages * p.a_stage_bytes + 2 * p.l_stage_bytes;
  const uint32_t chunk = (uint32_t)(p.blk_p * (pair ? p.NC / 2 : p.NC) * 16);  // one tap of one channel block (this CTA's columns)
  const int ntaps = tp ? p.KH * (p.KWs / 2) : p.KH * p.KW;
  int ew = (d->ctas_per_sm == 4 && !up && (!has_pre || p.BH * p.BW <= 4 * 128) && p.tmem_cols <= 128) ? 4 : 8;  // epilogue warps
  uint32_t smem_bytes = 0;
  for (;;) {
    const uint32_t tail = 8u * (2 * 2 + 2 * 8 + 1) + 16 + (uint32_t)(ew * 2 * p.NC * 4 * (bt ? T : 1)) + (ew == 8 ? 3 * 4 * 32 * 3 * 4 : 0) +
                          8 * 14 + (uint32_t)(2 * p.pre_c * 4) + 128;
    const uint32_t budget = ew == 4 ? 55 * 1024 : 112 * 1024;
    static const unsigned stage_big = getenv("PBT_B_STAGE_BYTES") ? (unsigned)atoi(getenv("PBT_B_STAGE_BYTES")) : 16384u;  // (tuning knob)
    static const unsigned stage_small = getenv("PBT_B_STAGE_BYTES_SMALL") ? (unsigned)atoi(getenv("PBT_B_STAGE_BYTES_SMALL")) : 8192u;
    int group = (int)((ew == 4 ? stage_small : stage_big) / chunk);
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
    smem_bytes = a_total + (uint32_t)p.b_stages * p.b_stage_bytes + tail;
    if (ew == 4 && smem_bytes > 56 * 1024) {
      PBT_REQUIRE(!pair, "conv: cta_pair with ctas_per_sm=4 does not fit 56 KB of shared memory");
      ew = 8;
      continue;
    }
    break;
  }
  if ((d->debug_flags & 8) && smem_bytes < 120 * 1024) smem_bytes = 120 * 1024;  // bring-up: force one CTA per SM
  PBT_REQUIRE(smem_bytes <= 227 * 1024, "conv: configuration does not fit shared memory");

  CUtensorMap tmap, tmapP;
  int rc = PBT_OK;
  if (in.c > 0) rc = up ? make_p8_tmap(&tmap, in, p.LBW, p.LBH, p.blk_p) : make_p8_tmap(&tmap, in, p.BW, p.BH, p.a_planes);
  if (rc != PBT_OK) return rc;
  if (has_pre) {
    rc = make_p8_tmap(&tmapP, d->pre, p.BW, p.BH, p.blk_p);
    if (rc != PBT_OK) return rc;
    if (in.c == 0) tmap = tmapP;
  } else {
    tmapP = tmap;
  }

  // one wave of persistent CTAs: SMs x co-resident CTAs (registers: 2 or 4; tensor memory; shared memory)
  int grid = (bt ? ceil_div(p.n_img, T) : p.n_img) * p.tiles_x * p.tiles_y;
  if (pair) grid = (grid + 1) & ~1;   // a pair's padding unit loads zeros and stores nothing
  p.n_units = grid;
  int occ = ew == 4 ? 4 : 2;
  if (512 / p.tmem_cols < occ) occ = 512 / p.tmem_cols;
  if ((int)(233472u / (smem_bytes + 1024u)) < occ) occ = (int)(233472u / (smem_bytes + 1024u));
  if (occ < 1) occ = 1;
  // Persistent only for queues of >= 3 units per CTA slot (measured: 2..8 are within 1 %, 3 is best at 960x540): the
  // static round-robin then balances well enough and
  // the kernel owns the GPU long enough.  Shorter launches, and launches flagged `concurrent` (the backward sweep, where
  // side-stream wgrad kernels must be able to take SM slots and tensor memory in between), keep one unit per CTA and the
  // hardware's dynamic block scheduling.
  static const int persist_min = getenv("PBT_PERSIST_MIN") ? atoi(getenv("PBT_PERSIST_MIN")) : 3;  // (tuning knob)
  if (!(d->debug_flags & 128) && !d->concurrent && grid >= persist_min * occ * num_sms()) grid = occ * num_sms();  // (bring-up: bit 7 = one unit per CTA)
  // Weight-stationary MMA runs (tcgen05.mma.ws, B kept in a collector buffer across the T tiles of a (tap, K step)): measured
  // on B200 (tools/ws_experiment.py, bit-identical results) +5 % on the N = 64 small-footprint layers (smoothers: 941 -> 893 us
  // per 4 x 1080p frames; their MMA streams are bound by the shared-memory operand fetch, A 4 KB + B 2 KB per MMA), neutral to
  // -3 % on N = 128 and on the 7x7 layers, so it is on exactly there.  Debug bit 5 forces it on (N in {64, 128, 256}), bit 6 off.
  p.ws_b = (!pair && T > 1 && ((p.NC == 64 && ew == 4 && ntaps <= 9) ||
                               ((d->debug_flags & 32) && (p.NC == 64 || p.NC == 128 || p.NC == 256)))) ? 1 : 0;
  if (d->debug_flags & 64) p.ws_b = 0;
  const int kb = d->blk_c / 16;
  if (pair) {
    PBT_REQUIRE((ew == 4) == (d->ctas_per_sm == 4), "conv: cta_pair + ctas_per_sm=4 shape does not fit the small footprint");
    if (ew == 4) return launch_conv_pair<2, 1, 4>(tmap, tmapP, p, grid, smem_bytes, stream);
    if (T == 1) return launch_conv_pair<1, 2, 8>(tmap, tmapP, p, grid, smem_bytes, stream);
    return T == 2 ? launch_conv_pair<2, 2, 8>(tmap, tmapP, p, grid, smem_bytes, stream)
                  : launch_conv_pair<3, 2, 8>(tmap, tmapP, p, grid, smem_bytes, stream);
  }
  switch (T) {
    case 1: return launch_conv_kb<1>(kb, ew, tmap, tmapP, p, grid, smem_bytes, stream);
    case 2: return launch_conv_kb<2>(kb, ew, tmap, tmapP, p, grid, smem_bytes, stream);
    default: return launch_conv_kb<3>(kb, ew, tmap, tmapP, p, grid, smem_bytes, stream);
  }
}

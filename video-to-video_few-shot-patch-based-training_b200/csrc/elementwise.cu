// elementwise.cu — HBM-bound kernels around the tensor-core convs: layout/dtype conversion,
// normalisation finalize/apply (+activation, residual, space-to-depth), bilinear x2 upsample and its
// transpose, normalisation backward (reduce + apply), fused head backward, small reductions.
// All kernels move one P8 pixel chunk (8 channels = 16 B) per thread access: coalesced 128-bit
// loads/stores, grid-stride loops sized to a multiple of the SM count.
#include <cooperative_groups.h>

#include "internal.h"
#include "ptx.cuh"

namespace pbt {

constexpr int kEwThreads = 256;

static inline int ew_grid(long long items) {
  long long blocks = (items + kEwThreads - 1) / kEwThreads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// grid (pixel chunks, planes, images) for kernels that loop over the pixels of one (image, plane)
static inline dim3 ew_grid3(int hw, int planes, int n) {
  long long per = (long long)planes * n;
  long long cap = ((long long)num_sms() * 16 + per - 1) / per;
  long long bx = ((long long)hw + kEwThreads - 1) / kEwThreads;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  return dim3((unsigned)bx, (unsigned)planes, (unsigned)n);
}

struct ActView {
  uint8_t* ptr;
  long long img_stride;  // elements
  int n, c, h, w;
};
static inline ActView view(const pbt_act_t& t) {
  ActView v;
  v.ptr = static_cast<uint8_t*>(t.ptr);
  v.img_stride = t.img_stride;
  v.n = t.n; v.c = t.c; v.h = t.h; v.w = t.w;
  return v;
}
__device__ __forceinline__ uint4* chunk_ptr(const ActView& v, int n, int plane, long long pix) {
  return reinterpret_cast<uint4*>(v.ptr + 2 * ((long long)n * v.img_stride + ((long long)plane * v.h * v.w + pix) * 8));
}

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == PBT_ACT_RELU) return fmaxf(v, 0.f);
  if (act == PBT_ACT_LEAKY02) return v > 0.f ? v : 0.2f * v;
  return v;
}
__device__ __forceinline__ float act_grad(float xhat, int act) {
  if (act == PBT_ACT_RELU) return xhat > 0.f ? 1.f : 0.f;
  if (act == PBT_ACT_LEAKY02) return xhat > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// ------------------------------------------------------------------ NCHW -> P8
template <int DT, bool SRC_HALF>
__global__ void nchw_to_p8_kernel(const void* __restrict__ x, int n, int c, int h, int w, ActView out) {
  pdl_sync();
  const long long hw = (long long)h * w;
  const int planes = out.c / 8;
  const long long total = (long long)n * planes * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int pl = (int)((i / hw) % planes);
    const int ni = (int)(i / (hw * planes));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = pl * 8 + k;
      float v = 0.f;
      if (ch < c) {
        const long long src = ((long long)ni * c + ch) * hw + pix;
        v = SRC_HALF ? __half2float(static_cast<const __half*>(x)[src]) : static_cast<const float*>(x)[src];
      }
      f[k] = v;
    }
    *chunk_ptr(out, ni, pl, pix) = pack8<DT>(f);
  }
}

template <int DT>
__global__ void p8_to_nchw_kernel(ActView in, int c, float* __restrict__ out, float mul) {
  pdl_sync();
  const long long hw = (long long)in.h * in.w;
  const int planes = (c + 7) / 8;
  const long long total = (long long)in.n * planes * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int pl = (int)((i / hw) % planes);
    const int ni = (int)(i / (hw * planes));
    float f[8];
    unpack8<DT>(*chunk_ptr(in, ni, pl, pix), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = pl * 8 + k;
      if (ch < c) out[((long long)ni * c + ch) * hw + pix] = f[k] * mul;
    }
  }
}

__global__ void p8f_to_nchw_kernel(const float* __restrict__ in, int n, int c_total, int c, int h, int w,
                                   float* __restrict__ out) {
  pdl_sync();
  const long long hw = (long long)h * w;
  const int planes = (c + 7) / 8;
  const long long total = (long long)n * planes * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int pl = (int)((i / hw) % planes);
    const int ni = (int)(i / (hw * planes));
    const float* src = in + (((long long)ni * (c_total / 8) + pl) * hw + pix) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = pl * 8 + k;
      if (ch < c) out[((long long)ni * c + ch) * hw + pix] = src[k];
    }
  }
}

// ------------------------------------------------------------------ uint8 frames
__device__ __forceinline__ float u8_to_norm(uint8_t u) {
  // torchvision ToTensor (x/255) then Normalize(0.5, 0.5): (v - 0.5) / 0.5, each step rounded to fp32
  return __fdiv_rn(__fsub_rn(__fdiv_rn((float)u, 255.f), 0.5f), 0.5f);
}

template <int DT>
__global__ void u8hwc_to_p8_kernel(const uint8_t* __restrict__ img, int n, int h, int w, int c, ActView out) {
  pdl_sync();
  const long long hw = (long long)h * w;
  const int planes = out.c / 8;
  const long long total = (long long)n * planes * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int pl = (int)((i / hw) % planes);
    const int ni = (int)(i / (hw * planes));
    float f[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int ch = pl * 8 + k;
      f[k] = ch < c ? u8_to_norm(img[((long long)ni * hw + pix) * c + ch]) : 0.f;
    }
    *chunk_ptr(out, ni, pl, pix) = pack8<DT>(f);
  }
}

__global__ void nchw_to_u8hwc_kernel(const float* __restrict__ y, int n, int c, int h, int w, uint8_t* __restrict__ out) {
  pdl_sync();
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int ni = (int)(i / hw);
    for (int ch = 0; ch < c; ++ch) {
      float v = y[((long long)ni * c + ch) * hw + pix];
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = fminf(fmaxf(__fmul_rn(__fadd_rn(v, 1.f), 127.5f), 0.f), 255.f);
      out[i * c + ch] = (uint8_t)rintf(v);
    }
  }
}

__global__ void u8hwc_to_norm_chw_kernel(const uint8_t* __restrict__ img, int h, int w, int c, float* __restrict__ out) {
  pdl_sync();
  const long long hw = (long long)h * w;
  const long long total = hw * c;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int ch = (int)(i / hw);
    out[i] = u8_to_norm(img[pix * c + ch]);
  }
}

__global__ void mask_dilate7_kernel(const uint8_t* __restrict__ m, int h, int w, uint8_t* __restrict__ out) {
  pdl_sync();
  const long long total = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i % w);
    int any = 0;
    for (int dy = -3; dy <= 3; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      for (int dx = -3; dx <= 3; ++dx) {
        const int xx = x + dx;
        if (xx < 0 || xx >= w) continue;
        any |= m[(long long)yy * w + xx];
      }
    }
    out[i] = any ? 1 : 0;
  }
}

// 7x7 erosion of a thresholded mask (reference generator.py:327-351 `_process_mask`: box sum of the 0/1 mask, zero padded,
// kept only where all 49 pixels are set): out = 1.0 where the whole window is inside the mask, else 0.0
__global__ void mask_erode7_kernel(const uint8_t* __restrict__ m, int n, int h, int w, float* __restrict__ out) {
  pdl_sync();
  const long long hw = (long long)h * w, total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const uint8_t* mi = m + (i - pix);
    const int y = (int)(pix / w), x = (int)(pix % w);
    int all = (y >= 3 && y < h - 3 && x >= 3 && x < w - 3) ? 1 : 0;   // zero padding: border windows are never full
    for (int dy = -3; dy <= 3 && all; ++dy)
      for (int dx = -3; dx <= 3; ++dx) all &= (mi[(long long)(y + dy) * w + x + dx] != 0);
    out[i] = all ? 1.f : 0.f;
  }
}

// mask composite + uint8 conversion of the frame loop (reference generator.py:562-563,643-647):
//   out = round(clamp((clamp(rgb*(1-m) + y*m, -1, 1) + 1) * 127.5, 0, 255)),  rgb = the frame's first three channels
// normalised like the generator input; every step rounded to fp32 like the tensor-library expression (no FMA contraction)
__global__ void composite_to_u8_kernel(const float* __restrict__ y, const uint8_t* __restrict__ frame, int c,
                                       const float* __restrict__ mask, int n, int h, int w, uint8_t* __restrict__ out) {
  pdl_sync();
  const long long hw = (long long)h * w, total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int ni = (int)(i / hw);
    const float mk = mask ? mask[i] : 1.f;
    const float one_m = __fsub_rn(1.f, mk);
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      float v = y[((long long)ni * 3 + ch) * hw + pix];
      if (mask) v = __fadd_rn(__fmul_rn(u8_to_norm(frame[i * c + ch]), one_m), __fmul_rn(v, mk));
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = fminf(fmaxf(__fmul_rn(__fadd_rn(v, 1.f), 127.5f), 0.f), 255.f);
      out[i * 3 + ch] = (uint8_t)rintf(v);
    }
  }
}

// zero the 16-byte pixel chunks outside [0,vh) x [0,vw) of every (image, plane)
__global__ void zero_border_kernel(ActView t, int vh, int vw) {
  pdl_sync();
  const int hw = t.h * t.w;
  const int pl = blockIdx.y, ni = blockIdx.z;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
    const int y = pix / t.w, x = pix - y * t.w;
    if (y >= vh || x >= vw) *chunk_ptr(t, ni, pl, pix) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// space-to-depth P8 -> NCHW fp32: out[n][ch][2y+py][2x+px] = in[n][(py*2+px)*cpp + ch][y][x] * mul
template <int DT>
__global__ void p8s2d_to_nchw_kernel(ActView in, int cpp, int c, float* __restrict__ out, const float* __restrict__ mul_dev) {
  pdl_sync();
  const float mul = mul_dev ? __ldg(mul_dev) : 1.f;
  const int oh = 2 * in.h, ow = 2 * in.w;
  const long long total = (long long)in.n * c * oh * ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % ow);
    const int Y = (int)((i / ow) % oh);
    const int ch = (int)((i / ((long long)ow * oh)) % c);
    const int ni = (int)(i / ((long long)ow * oh * c));
    const int sc = ((Y & 1) * 2 + (X & 1)) * cpp + ch;
    float f[8];
    unpack8<DT>(*chunk_ptr(in, ni, sc >> 3, (long long)(Y >> 1) * in.w + (X >> 1)), f);
    out[i] = f[sc & 7] * mul;
  }
}

// ------------------------------------------------------------------ norm finalize
// Stage 1 of the deterministic statistics reduction (large images have thousands of tiles): block
// (32-channel group, chunk, image) sums the tiles of its chunk in a fixed order (double accumulation) and
// overwrites the FIRST tile slot of its own chunk with the result, so stage 2 only visits tiles/chunk slots.
__global__ void norm_reduce_partials_kernel(float* __restrict__ partial, int tiles, int c, int chunk) {
  pdl_sync();
  __shared__ double s_sum[8][32], s_sq[8][32];
  const int cg = blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = threadIdx.x >> 5;
  const int t0 = blockIdx.y * chunk;
  const int t1 = min(tiles, t0 + chunk);
  float* base = partial + (long long)blockIdx.z * tiles * 2 * c;
  double a = 0.0, b = 0.0;
  if (cg < c) {
    for (int t = t0 + row; t < t1; t += 8) {
      const float* pp = base + (long long)t * 2 * c;
      a += (double)pp[cg];
      b += (double)pp[c + cg];
    }
  }
  s_sum[row][threadIdx.x & 31] = a;
  s_sq[row][threadIdx.x & 31] = b;
  __syncthreads();  // every read of this chunk is done before its first slot is overwritten
  if (row == 0 && cg < c) {
    for (int r2 = 1; r2 < 8; ++r2) {
      a += s_sum[r2][threadIdx.x];
      b += s_sq[r2][threadIdx.x];
    }
    float* dst = base + (long long)t0 * 2 * c;
    dst[cg] = (float)a;
    dst[c + cg] = (float)b;
  }
}

// one block per (image, 32-channel group) in instance mode; per 32-channel group in batch mode.
// `tiles` slots are visited with stride `tile_stride` (= the stage-1 chunk, or 1).
__global__ void norm_finalize_kernel(const float* __restrict__ partial, int n, int tiles, int tile_stride, int img_tiles, int c,
                                     long long count_per_image,
                                     float eps, int batch_mode, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                                     float* __restrict__ scale, float* __restrict__ shift, float* mean_out, float* rstd_out) {
  pdl_sync();
  __shared__ double s_sum[8][32], s_sq[8][32];
  const int cg = blockIdx.x * 32 + (threadIdx.x & 31);
  const int row = threadIdx.x >> 5;  // 0..7
  const int img = batch_mode ? 0 : blockIdx.y;
  const int n_lo = batch_mode ? 0 : img, n_hi = batch_mode ? n : img + 1;
  double a = 0.0, b = 0.0;
  if (cg < c) {
    // (image, slot) pairs are walked flat so that all 8 rows stay busy in batch mode (n images x few slots)
    const int total = (n_hi - n_lo) * tiles;
    for (int idx = row; idx < total; idx += 8) {
      const int ni = n_lo + idx / tiles, t = idx % tiles;
      const float* pp = partial + ((long long)ni * img_tiles + (long long)t * tile_stride) * 2 * c;
      a += (double)pp[cg];
      b += (double)pp[c + cg];
    }
  }
  s_sum[row][threadIdx.x & 31] = a;
  s_sq[row][threadIdx.x & 31] = b;
  __syncthreads();
  if (row == 0 && cg < c) {
    for (int r2 = 1; r2 < 8; ++r2) {
      a += s_sum[r2][threadIdx.x];
      b += s_sq[r2][threadIdx.x];
    }
    const double cnt = (double)count_per_image * (batch_mode ? n : 1);
    const double mean = a / cnt;
    double var = b / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float meanf = (float)mean;
    if (batch_mode) {
      const float g = gamma ? gamma[cg] : 1.f, be = beta ? beta[cg] : 0.f;
      const float sc = g * rstd, sh = be - meanf * sc;
      for (int ni = 0; ni < n; ++ni) {
        scale[(long long)ni * c + cg] = sc;
        shift[(long long)ni * c + cg] = sh;
      }
      if (mean_out) mean_out[cg] = meanf;
      if (rstd_out) rstd_out[cg] = rstd;
      if (running_mean) {
        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
        running_mean[cg] = (1.f - momentum) * running_mean[cg] + momentum * meanf;
        running_var[cg] = (1.f - momentum) * running_var[cg] + momentum * (float)unbiased;
      }
    } else {
      scale[(long long)img * c + cg] = rstd;
      shift[(long long)img * c + cg] = -meanf * rstd;
      if (mean_out) mean_out[(long long)img * c + cg] = meanf;
      if (rstd_out) rstd_out[(long long)img * c + cg] = rstd;
    }
  }
}

// ------------------------------------------------------------------ norm apply
struct NormApplyK {
  ActView x, out, out_relu, out_s2d, res16;
  const float* scale;
  const float* shift;
  int per_channel, act;
  const float* residual32;
  float* out32;
  // fused InstanceNorm finalize: scale / shift are computed here from the conv's per-tile (sum, sumsq) partials and also
  // written out (block x == 0) for the backward pass, instead of a separate pbt_norm_finalize launch
  const float* partial;
  float* scale_out;
  float* shift_out;
  int tiles;
  double inv_count;
  float eps;
};

template <int DT>
__global__ void norm_apply_kernel(NormApplyK p) {
  pdl_sync();
  const int hw = p.x.h * p.x.w;
  const int planes = p.x.c / 8;
  const int pl = blockIdx.y, ni = blockIdx.z;  // grid = (pixel chunks, planes, images): no 64-bit div/mod per element
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = 1.f;
    sh[k] = 0.f;
  }
  if (p.partial) {
    // per-image statistics of this block's 8 channels: the first 16 threads add up the tiles (double, fixed order - the same
    // arithmetic as norm_finalize_kernel), everyone reads the result from shared memory
    __shared__ double s_ss[16];
    if (threadIdx.x < 16) {
      const int which = threadIdx.x >> 3, k = threadIdx.x & 7;
      const float* pp = p.partial + (long long)ni * p.tiles * 2 * p.x.c + (long long)which * p.x.c + pl * 8 + k;
      double a = 0.0;
      for (int t = 0; t < p.tiles; ++t) a += (double)pp[(long long)t * 2 * p.x.c];
      s_ss[threadIdx.x] = a * p.inv_count;     // mean, mean of squares
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double mean = s_ss[k];
      double var = s_ss[8 + k] - mean * mean;
      if (var < 0.0) var = 0.0;
      const float rstd = (float)(1.0 / sqrt(var + (double)p.eps));
      sc[k] = rstd;
      sh[k] = -(float)mean * rstd;
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) {
      p.scale_out[(long long)ni * p.x.c + pl * 8 + threadIdx.x] = sc[threadIdx.x];
      p.shift_out[(long long)ni * p.x.c + pl * 8 + threadIdx.x] = sh[threadIdx.x];
    }
  } else if (p.scale) {
    const long long so = (p.per_channel ? 0 : (long long)ni * p.x.c) + pl * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = __ldg(&p.scale[so + k]);
      sh[k] = __ldg(&p.shift[so + k]);
    }
  }
  const bool affine = p.scale || p.partial;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += gridDim.x * blockDim.x) {
    float f[8];
    unpack8<DT>(*chunk_ptr(p.x, ni, pl, pix), f);
    if (affine) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc[k], sh[k]);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = apply_act(f[k], p.act);
    const long long dense = (((long long)ni * planes + pl) * hw + pix) * 8;
    if (p.residual32) {
      const float4 a0 = *reinterpret_cast<const float4*>(p.residual32 + dense);
      const float4 a1 = *reinterpret_cast<const float4*>(p.residual32 + dense + 4);
      f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w;
      f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
    }
    if (p.res16.ptr) {
      float r16[8];
      unpack8<DT>(*chunk_ptr(p.res16, ni, pl, pix), r16);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r16[k];
    }
    if (p.out32) {
      *reinterpret_cast<float4*>(p.out32 + dense) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(p.out32 + dense + 4) = make_float4(f[4], f[5], f[6], f[7]);
    }
    if (p.out.ptr || p.out_s2d.ptr) {
      const uint4 u = pack8<DT>(f);
      if (p.out.ptr) *chunk_ptr(p.out, ni, pl, pix) = u;
      if (p.out_s2d.ptr) {
        const int y = (int)(pix / p.x.w), x = (int)(pix % p.x.w);
        const int phase = (y & 1) * 2 + (x & 1);
        *chunk_ptr(p.out_s2d, ni, phase * planes + pl, (long long)(y >> 1) * p.out_s2d.w + (x >> 1)) = u;
      }
    }
    if (p.out_relu.ptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.f);
      *chunk_ptr(p.out_relu, ni, pl, pix) = pack8<DT>(f);
    }
  }
}

// ------------------------------------------------------------------ bilinear x2 (align_corners=True)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  const float s = scale * (float)dst;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = s - (float)i0;
}

// One thread = a 2x2 block of output pixels of one (image, plane).
// The 2x2 outputs read at most 3x3 input pixels (align_corners ratio ~0.5), so every input chunk is loaded,
// normalised and activated once per thread instead of four times.
// Optional fused producer: taps are normalised + activated on load (x*scale+shift, act), so the
// InstanceNorm/ReLU of the previous conv never round-trips HBM at the low resolution.
template <int DT>
__global__ void upsample2x_kernel(ActView in, ActView out, const float* __restrict__ scale, const float* __restrict__ shift,
                                  int act) {
  pdl_sync();
  const int oh = out.h, ow = out.w;
  // flat index over the 2x2 output blocks of one (image, plane): no idle threads on patch-sized maps (a 64x16-pixel
  // thread block covered an 80-pixel-wide map at 62 %)
  const int bw = (ow + 1) >> 1, bh = (oh + 1) >> 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= bw * bh) return;
  const int by = t / bw;
  const int X0 = (t - by * bw) * 2, Y0 = by * 2;
  const int pl = blockIdx.y, ni = blockIdx.z;
  const float sy = oh > 1 ? (float)(in.h - 1) / (float)(oh - 1) : 0.f;
  const float sx = ow > 1 ? (float)(in.w - 1) / (float)(ow - 1) : 0.f;
  int ya[2], yb[2], xa[2], xb[2];
  float ly[2], lx[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    src_index(min(Y0 + j, oh - 1), sy, in.h, ya[j], yb[j], ly[j]);
    src_index(min(X0 + j, ow - 1), sx, in.w, xa[j], xb[j], lx[j]);
  }
  const int ylo = ya[0], xlo = xa[0];
  if constexpr (DT == 1) {
    // fp16 tensors: packed half2 arithmetic end to end (the fp32 path below needs 128 registers - two 256-thread blocks
    // per SM - and ran at a quarter of the HBM roofline on patch-sized maps).  Weights rounded to fp16 add ~2^-11
    // relative error per blend, the same order as the fp16 rounding of the result.
    __half2 sc2[4], sh2[4];
    if (scale) {
      const long long so = (long long)ni * in.c + pl * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        sc2[k] = __floats2half2_rn(__ldg(&scale[so + 2 * k]), __ldg(&scale[so + 2 * k + 1]));
        sh2[k] = __floats2half2_rn(__ldg(&shift[so + 2 * k]), __ldg(&shift[so + 2 * k + 1]));
      }
    }
    const __half2 zero2 = __float2half2_rn(0.f), leak2 = __float2half2_rn(0.2f);
    uint4 hv[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int yy = min(ylo + r, in.h - 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int xx = min(xlo + c, in.w - 1);
        hv[r][c] = *chunk_ptr(in, ni, pl, (long long)yy * in.w + xx);
        if (scale) {
          __half2* h = reinterpret_cast<__half2*>(&hv[r][c]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            __half2 t = __hfma2(h[k], sc2[k], sh2[k]);
            if (act == PBT_ACT_RELU) t = __hmax2(t, zero2);
            else if (act == PBT_ACT_LEAKY02) t = __hmax2(t, __hmul2(t, leak2));
            h[k] = t;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (Y0 + j >= oh) break;
      const int ra = ya[j] - ylo, rb = yb[j] - ylo;   // 0..1 and 0..2
      const __half2 hy = __float2half2_rn(1.f - ly[j]), wy = __float2half2_rn(ly[j]);
      uint4 row[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint4 top = ra == 0 ? hv[0][c] : hv[1][c];
        const uint4 bot = rb == 0 ? hv[0][c] : (rb == 1 ? hv[1][c] : hv[2][c]);
        const __half2* tp = reinterpret_cast<const __half2*>(&top);
        const __half2* bp = reinterpret_cast<const __half2*>(&bot);
        __half2* rp = reinterpret_cast<__half2*>(&row[c]);
#pragma unroll
        for (int k = 0; k < 4; ++k) rp[k] = __hfma2(wy, bp[k], __hmul2(hy, tp[k]));
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (X0 + i >= ow) break;
        const int ca = xa[i] - xlo, cb = xb[i] - xlo;
        const __half2 hx = __float2half2_rn(1.f - lx[i]), wx = __float2half2_rn(lx[i]);
        const uint4 l = ca == 0 ? row[0] : row[1];
        const uint4 rr = cb == 0 ? row[0] : (cb == 1 ? row[1] : row[2]);
        const __half2* lp = reinterpret_cast<const __half2*>(&l);
        const __half2* rp = reinterpret_cast<const __half2*>(&rr);
        uint4 o;
        __half2* op = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) op[k] = __hfma2(wx, rp[k], __hmul2(hx, lp[k]));
        *chunk_ptr(out, ni, pl, (long long)(Y0 + j) * ow + X0 + i) = o;
      }
    }
    return;
  }
  float sc[8], sh[8];
  if (scale) {
    const long long so = (long long)ni * in.c + pl * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      sc[k] = __ldg(&scale[so + k]);
      sh[k] = __ldg(&shift[so + k]);
    }
  }
  float v[3][3][8];  // [row][col][channel] of the 3x3 input neighbourhood (clamped at the border)
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int yy = min(ylo + r, in.h - 1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int xx = min(xlo + c, in.w - 1);
      unpack8<DT>(*chunk_ptr(in, ni, pl, (long long)yy * in.w + xx), v[r][c]);
      if (scale) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[r][c][k] = apply_act(fmaf(v[r][c][k], sc[k], sh[k]), act);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (Y0 + j >= oh) break;
    const int ra = ya[j] - ylo, rb = yb[j] - ylo;   // 0..1 and 0..2
    const float hy = 1.f - ly[j], wy = ly[j];
    float row[3][8];                                // the three columns lerped along y
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float top = ra == 0 ? v[0][c][k] : v[1][c][k];
        const float bot = rb == 0 ? v[0][c][k] : (rb == 1 ? v[1][c][k] : v[2][c][k]);
        row[c][k] = hy * top + wy * bot;
      }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (X0 + i >= ow) break;
      const int ca = xa[i] - xlo, cb = xb[i] - xlo;
      const float hx = 1.f - lx[i], wx = lx[i];
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float l = ca == 0 ? row[0][k] : row[1][k];
        const float rr = cb == 0 ? row[0][k] : (cb == 1 ? row[1][k] : row[2][k]);
        o[k] = hx * l + wx * rr;
      }
      *chunk_ptr(out, ni, pl, (long long)(Y0 + j) * ow + X0 + i) = pack8<DT>(o);
    }
  }
}

// transpose of the above: each low-res pixel gathers the high-res gradients that read it
template <int DT>
__global__ void upsample2x_bwd_kernel(ActView gout, ActView gin16, float* gin32, int ih, int iw) {
  pdl_sync();
  const int oh = gout.h, ow = gout.w;
  const long long ihw = (long long)ih * iw;
  const int planes = gout.c / 8;
  const float sy = oh > 1 ? (float)(ih - 1) / (float)(oh - 1) : 0.f;
  const float sx = ow > 1 ? (float)(iw - 1) / (float)(ow - 1) : 0.f;
  const int pl = blockIdx.y, ni = blockIdx.z;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < (int)ihw; pix += gridDim.x * blockDim.x) {
    const int y = pix / iw, x = pix - y * iw;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if constexpr (DT == 1) {
      // fp16 gradients: the pixels that read low-res (y, x) lie in rows 2y-2..2y+3 and columns 2x-2..2x+3
      // ((x-1)/r >= 2x-2 and (x+1)/r <= 2x+3 for r = (iw-1)/(2iw-1)).  Columns are walked branch-free with packed half2
      // FMAs (a zero weight for the pixels that do not contribute): no divergence between the lanes of a warp, a
      // quarter of the fp32 path's instructions.  Rows are skipped warp-uniformly.
      const int Yb = 2 * y - 2, Xb = 2 * x - 2;
      __half2 wx2[6];
      float wyf[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        int a0, a1;
        float l, wv = 0.f;
        wyf[j] = 0.f;
        if (Yb + j >= 0 && Yb + j < oh) {
          src_index(Yb + j, sy, ih, a0, a1, l);
          if (a0 == y) wyf[j] += 1.f - l;
          if (a1 == y) wyf[j] += l;
        }
        if (Xb + j >= 0 && Xb + j < ow) {
          src_index(Xb + j, sx, iw, a0, a1, l);
          if (a0 == x) wv += 1.f - l;
          if (a1 == x) wv += l;
        }
        wx2[j] = __float2half2_rn(wv);
      }
      const __half2 zero2 = __float2half2_rn(0.f);
#pragma unroll
      for (int jy = 0; jy < 6; ++jy) {
        if (wyf[jy] == 0.f) continue;
        __half2 r2[4] = {zero2, zero2, zero2, zero2};
        const long long rowbase = (long long)(Yb + jy) * ow;
#pragma unroll
        for (int jx = 0; jx < 6; ++jx) {
          const int X = min(max(Xb + jx, 0), ow - 1);       // clamped address, zero weight outside the image
          const uint4 u = *chunk_ptr(gout, ni, pl, rowbase + X);
          const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
          for (int k = 0; k < 4; ++k) r2[k] = __hfma2(wx2[jx], h[k], r2[k]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(r2[k]);
          acc[2 * k] = fmaf(wyf[jy], f.x, acc[2 * k]);
          acc[2 * k + 1] = fmaf(wyf[jy], f.y, acc[2 * k + 1]);
        }
      }
      if (gin16.ptr) *chunk_ptr(gin16, ni, pl, pix) = pack8<DT>(acc);
      if (gin32) {
        float* dst = gin32 + (((long long)ni * planes + pl) * ihw + pix) * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
      continue;
    }
    const int Ylo = max(0, 2 * y - 3), Yhi = min(oh - 1, 2 * y + 4);
    const int Xlo = max(0, 2 * x - 3), Xhi = min(ow - 1, 2 * x + 4);
    // separable weights of the (at most 8 x 8) high-res window: 16 index computations instead of 8 + 64
    float wys[8], wxs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int a0, a1;
      float l;
      wys[j] = wxs[j] = 0.f;
      if (Ylo + j <= Yhi) {
        src_index(Ylo + j, sy, ih, a0, a1, l);
        if (a0 == y) wys[j] += 1.f - l;
        if (a1 == y) wys[j] += l;
      }
      if (Xlo + j <= Xhi) {
        src_index(Xlo + j, sx, iw, a0, a1, l);
        if (a0 == x) wxs[j] += 1.f - l;
        if (a1 == x) wxs[j] += l;
      }
    }
#pragma unroll
    for (int jy = 0; jy < 8; ++jy) {
      if (wys[jy] == 0.f) continue;
#pragma unroll
      for (int jx = 0; jx < 8; ++jx) {
        if (wxs[jx] == 0.f) continue;
        float g[8];
        unpack8<DT>(*chunk_ptr(gout, ni, pl, (long long)(Ylo + jy) * ow + Xlo + jx), g);
        const float wgt = wys[jy] * wxs[jx];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, g[k], acc[k]);
      }
    }
    if (gin16.ptr) *chunk_ptr(gin16, ni, pl, pix) = pack8<DT>(acc);
    if (gin32) {
      float* dst = gin32 + (((long long)ni * planes + pl) * ihw + pix) * 8;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

// ------------------------------------------------------------------ norm backward
struct NormBwdK {
  ActView x, ga, gb16, dx;
  const float* mean;   // "scale"/"shift" of the ABI carry rstd and -mean*rstd (xhat = x*scale + shift)
  const float* shift;
  int per_channel, act, ga_is_s2d, batch_mode, relu_mask_x;
  const float* gb32;
  float* sums;
  const float* kmul;
  float inv_count;
};

template <int DT>
__device__ __forceinline__ void load_gact(const NormBwdK& p, int ni, int pl, long long pix, int planes, long long hw,
                                          float* gact, float* xhat, float* xr) {
  unpack8<DT>(*chunk_ptr(p.x, ni, pl, pix), xr);
  const long long so = (p.per_channel ? 0 : (long long)ni * p.x.c) + pl * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) xhat[k] = fmaf(xr[k], __ldg(&p.mean[so + k]), __ldg(&p.shift[so + k]));
  float g[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = 0.f;
  if (p.ga.ptr) {
    float t[8];
    if (p.ga_is_s2d) {
      const int y = (int)(pix / p.x.w), x = (int)(pix % p.x.w);
      const int phase = (y & 1) * 2 + (x & 1);
      unpack8<DT>(*chunk_ptr(p.ga, ni, phase * planes + pl, (long long)(y >> 1) * p.ga.w + (x >> 1)), t);
    } else {
      unpack8<DT>(*chunk_ptr(p.ga, ni, pl, pix), t);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += t[k];
  }
  if (p.gb16.ptr) {
    float t[8];
    unpack8<DT>(*chunk_ptr(p.gb16, ni, pl, pix), t);
#pragma unroll
    for (int k = 0; k < 8; ++k) g[k] += t[k];
  }
  if (p.gb32) {
    const float* s = p.gb32 + (((long long)ni * planes + pl) * hw + pix) * 8;
    const float4 a0 = *reinterpret_cast<const float4*>(s), a1 = *reinterpret_cast<const float4*>(s + 4);
    g[0] += a0.x; g[1] += a0.y; g[2] += a0.z; g[3] += a0.w;
    g[4] += a1.x; g[5] += a1.y; g[6] += a1.z; g[7] += a1.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) gact[k] = g[k] * act_grad(xhat[k], p.act);
}

// grid: (chunks, planes, n); each block reduces a pixel range of one (image, plane)
template <int DT>
__global__ void norm_bwd_reduce_kernel(NormBwdK p) {
  pdl_sync();
  const long long hw = (long long)p.x.h * p.x.w;
  const int planes = p.x.c / 8;
  const int pl = blockIdx.y, ni = blockIdx.z;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < hw; pix += (long long)gridDim.x * blockDim.x) {
    float gact[8], xhat[8], xr[8];
    load_gact<DT>(p, ni, pl, pix, planes, hw, gact, xhat, xr);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += gact[k];
      s2[k] = fmaf(gact[k], xhat[k], s2[k]);
    }
  }
  __shared__ float red[kEwThreads / 32][16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[wid][k] = s1[k];
      red[wid][8 + k] = s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) t += red[w2][threadIdx.x];
    const int which = threadIdx.x >> 3, k = threadIdx.x & 7;
    float* dst = p.sums + (p.batch_mode ? 0 : (long long)ni * 2 * p.x.c) + (long long)which * p.x.c + pl * 8 + k;
    atomicAdd(dst, t);
  }
}

template <int DT>
__global__ void norm_bwd_apply_kernel(NormBwdK p) {
  pdl_sync();
  const long long hw = (long long)p.x.h * p.x.w;
  const int planes = p.x.c / 8;
  const int pl = blockIdx.y, ni = blockIdx.z;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < (int)hw; pix += gridDim.x * blockDim.x) {
    float gact[8], xhat[8], r[8], xr[8];
    load_gact<DT>(p, ni, pl, pix, planes, hw, gact, xhat, xr);
    const float* sp = p.sums + (p.batch_mode ? 0 : (long long)ni * 2 * p.x.c) + pl * 8;
    const long long ko = (p.per_channel ? 0 : (long long)ni * p.x.c) + pl * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float m1 = __ldg(&sp[k]) * p.inv_count, m2 = __ldg(&sp[p.x.c + k]) * p.inv_count;
      r[k] = __ldg(&p.kmul[ko + k]) * (gact[k] - m1 - xhat[k] * m2);
      if (p.relu_mask_x && !(xr[k] > 0.f)) r[k] = 0.f;
    }
    *chunk_ptr(p.dx, ni, pl, pix) = pack8<DT>(r);
  }
}

// Fused reduce + apply for per-image statistics on small maps (patch training): one CTA owns one (image, plane) slice,
// sums it, then applies.  No atomics, one launch.  STAGE: the slice (x and the activated gradient, 32 B per pixel) is
// kept in shared memory between the two passes, so DRAM is read once; without it the second pass re-reads global memory
// (L2 hits only while the resident CTAs' slices fit L2 - at 80 x 128 x 80x80 they do not).  grid: (planes, n).
template <int DT, bool STAGE>
__global__ void __launch_bounds__(512) norm_bwd_fused_kernel(NormBwdK p) {
  pdl_sync();
  extern __shared__ __align__(16) uint4 s_slice[];  // STAGE: [hw][2] = (x chunk, gact chunk)
  const int hw = p.x.h * p.x.w;
  const int planes = p.x.c / 8;
  const int pl = blockIdx.x, ni = blockIdx.y;
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  for (int pix = threadIdx.x; pix < hw; pix += blockDim.x) {
    float gact[8], xhat[8], xr[8];
    load_gact<DT>(p, ni, pl, pix, planes, hw, gact, xhat, xr);
    if (STAGE) {
      const uint4 gq = pack8<DT>(gact);
      s_slice[2 * pix] = pack8<DT>(xr);   // exact: xr came out of the same 16-bit chunk
      s_slice[2 * pix + 1] = gq;
      unpack8<DT>(gq, gact);              // the sums see what the second pass will see
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += gact[k];
      s2[k] = fmaf(gact[k], xhat[k], s2[k]);
    }
  }
  __shared__ float red[32][16];
  __shared__ float tot[16];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[wid][k] = s1[k];
      red[wid][8 + k] = s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) t += red[w2][threadIdx.x];
    tot[threadIdx.x] = t;
    const int which = threadIdx.x >> 3, k = threadIdx.x & 7;
    p.sums[(long long)ni * 2 * p.x.c + (long long)which * p.x.c + pl * 8 + k] = t;
  }
  __syncthreads();
  float m1[8], m2[8], km[8], mu[8], sh[8];
  const long long ko = (p.per_channel ? 0 : (long long)ni * p.x.c) + pl * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m1[k] = tot[k] * p.inv_count;
    m2[k] = tot[8 + k] * p.inv_count;
    km[k] = __ldg(&p.kmul[ko + k]);
    mu[k] = __ldg(&p.mean[ko + k]);
    sh[k] = __ldg(&p.shift[ko + k]);
  }
  for (int pix = threadIdx.x; pix < hw; pix += blockDim.x) {
    float gact[8], xhat[8], r[8], xr[8];
    if (STAGE) {
      unpack8<DT>(s_slice[2 * pix], xr);
      unpack8<DT>(s_slice[2 * pix + 1], gact);
#pragma unroll
      for (int k = 0; k < 8; ++k) xhat[k] = fmaf(xr[k], mu[k], sh[k]);
    } else {
      load_gact<DT>(p, ni, pl, pix, planes, hw, gact, xhat, xr);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      r[k] = km[k] * (gact[k] - m1[k] - xhat[k] * m2[k]);
      if (p.relu_mask_x && !(xr[k] > 0.f)) r[k] = 0.f;
    }
    *chunk_ptr(p.dx, ni, pl, pix) = pack8<DT>(r);
  }
}

// Cluster variant for large slices (hw >= 4096): the (image, plane) slice is split over a cluster of kNbCluster CTAs, each
// staging its quarter in shared memory (51 KB at 80x80 -> four CTAs per SM instead of one 205 KB CTA, whose 16 warps could
// not keep enough loads in flight: 190 us for 393 MB).  The 16 partial sums are exchanged through distributed shared
// memory and added in rank order (deterministic).  grid: (planes * kNbCluster, n), cluster (kNbCluster, 1, 1).
constexpr int kNbCluster = 4;
template <int DT>
__global__ void __launch_bounds__(kEwThreads) norm_bwd_fused_cluster_kernel(NormBwdK p) {
  pdl_sync();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) uint4 s_slice[];  // [pixels of this CTA][2] = (x chunk, gact chunk)
  __shared__ float red[kEwThreads / 32][16];
  __shared__ float part[16];
  __shared__ float tot[16];
  const int hw = p.x.h * p.x.w;
  const int planes = p.x.c / 8;
  const int rank = (int)cluster.block_rank();
  const int pl = blockIdx.x / kNbCluster, ni = blockIdx.y;
  const int per = (hw + kNbCluster - 1) / kNbCluster;
  const int p0 = rank * per, p1 = min(hw, p0 + per);
  float s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s1[k] = s2[k] = 0.f;
  for (int pix = p0 + threadIdx.x; pix < p1; pix += blockDim.x) {
    float gact[8], xhat[8], xr[8];
    load_gact<DT>(p, ni, pl, pix, planes, hw, gact, xhat, xr);
    const uint4 gq = pack8<DT>(gact);
    s_slice[2 * (pix - p0)] = pack8<DT>(xr);
    s_slice[2 * (pix - p0) + 1] = gq;
    unpack8<DT>(gq, gact);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s1[k] += gact[k];
      s2[k] = fmaf(gact[k], xhat[k], s2[k]);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], o);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[wid][k] = s1[k];
      red[wid][8 + k] = s2[k];
    }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) t += red[w2][threadIdx.x];
    part[threadIdx.x] = t;
  }
  cluster.sync();                                   // every CTA's partial sums are visible cluster-wide
  if (threadIdx.x < 16) {
    float t = 0.f;
    for (int r = 0; r < kNbCluster; ++r) t += *cluster.map_shared_rank(&part[threadIdx.x], r);
    tot[threadIdx.x] = t;
    if (rank == 0) {
      const int which = threadIdx.x >> 3, k = threadIdx.x & 7;
      p.sums[(long long)ni * 2 * p.x.c + (long long)which * p.x.c + pl * 8 + k] = t;
    }
  }
  cluster.sync();                                   // nobody exits while a peer may still read its `part`
  float m1[8], m2[8], km[8], mu[8], sh[8];
  const long long ko = (p.per_channel ? 0 : (long long)ni * p.x.c) + pl * 8;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    m1[k] = tot[k] * p.inv_count;
    m2[k] = tot[8 + k] * p.inv_count;
    km[k] = __ldg(&p.kmul[ko + k]);
    mu[k] = __ldg(&p.mean[ko + k]);
    sh[k] = __ldg(&p.shift[ko + k]);
  }
  for (int pix = p0 + threadIdx.x; pix < p1; pix += blockDim.x) {
    float gact[8], xhat[8], r[8], xr[8];
    unpack8<DT>(s_slice[2 * (pix - p0)], xr);
    unpack8<DT>(s_slice[2 * (pix - p0) + 1], gact);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      xhat[k] = fmaf(xr[k], mu[k], sh[k]);
      r[k] = km[k] * (gact[k] - m1[k] - xhat[k] * m2[k]);
      if (p.relu_mask_x && !(xr[k] > 0.f)) r[k] = 0.f;
    }
    *chunk_ptr(p.dx, ni, pl, pix) = pack8<DT>(r);
  }
}

// ------------------------------------------------------------------ head backward
// grid: (chunks, planes of s); block reduces dW[3][8], dbias_prev[8] (+ db[3] on plane 0)
template <int DT>
__global__ void head_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, ActView s,
                                const float* __restrict__ head_w, const float* __restrict__ gscale, int head_tanh,
                                float* dw, float* db, ActView gs, float* dbias_prev) {
  pdl_sync();
  const long long hw = (long long)s.h * s.w;
  const long long total = (long long)s.n * hw;
  const int pl = blockIdx.y;
  const int C = s.c;
  const float sc = gscale ? __ldg(gscale) : 1.f;
  float w0[8], w1[8], w2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    w0[k] = __ldg(&head_w[pl * 8 + k]);
    w1[k] = __ldg(&head_w[C + pl * 8 + k]);
    w2[k] = __ldg(&head_w[2 * C + pl * 8 + k]);
  }
  float acc[35];
#pragma unroll
  for (int k = 0; k < 35; ++k) acc[k] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int ni = (int)(i / hw);
    float gz[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const long long o = ((long long)ni * 3 + j) * hw + pix;
      const float yy = y[o];
      gz[j] = gy[o] * (head_tanh ? (1.f - yy * yy) : 1.f) * sc;
    }
    float sv[8], g[8];
    unpack8<DT>(*chunk_ptr(s, ni, pl, pix), sv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] = fmaf(gz[0], sv[k], acc[k]);
      acc[8 + k] = fmaf(gz[1], sv[k], acc[8 + k]);
      acc[16 + k] = fmaf(gz[2], sv[k], acc[16 + k]);
      const float gk = sv[k] > 0.f ? (w0[k] * gz[0] + w1[k] * gz[1] + w2[k] * gz[2]) : 0.f;
      g[k] = gk;
    }
    const uint4 u = pack8<DT>(g);
    *chunk_ptr(gs, ni, pl, pix) = u;
    unpack8<DT>(u, g);  // bias grad sums what the dgrad/wgrad kernels will actually see
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[24 + k] += g[k];
    acc[32] += gz[0];
    acc[33] += gz[1];
    acc[34] += gz[2];
  }
  __shared__ float red[kEwThreads / 32][35];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 35; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[wid][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 35) {
    float t = 0.f;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) t += red[w2][threadIdx.x];
    const int k = threadIdx.x;
    if (k < 24) atomicAdd(&dw[(k >> 3) * C + pl * 8 + (k & 7)], t);
    else if (k < 32) { if (dbias_prev) atomicAdd(&dbias_prev[pl * 8 + (k - 24)], t); }
    else if (pl == 0) atomicAdd(&db[k - 32], t);
  }
}

// grid: (chunks, planes)
template <int DT>
__global__ void channel_sum_kernel(ActView g, float* out, const float* __restrict__ inv_scale) {
  pdl_sync();
  const long long hw = (long long)g.h * g.w;
  const long long total = (long long)g.n * hw;
  const int pl = blockIdx.y;
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8<DT>(*chunk_ptr(g, (int)(i / hw), pl, i % hw), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
  }
  __shared__ float red[kEwThreads / 32][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) red[wid][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) t += red[w2][threadIdx.x];
    atomicAdd(&out[pl * 8 + threadIdx.x], t * (inv_scale ? __ldg(inv_scale) : 1.f));
  }
}

__global__ void absmax_kernel(const float* __restrict__ g, long long count, float* out) {
  pdl_sync();
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(g[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));  // m >= 0: int order == float order
}

__global__ void make_grad_scale_kernel(const float* amax, float target, float* scale2, const float* adjust) {
  pdl_sync();
  const float a = *amax;
  if (adjust) target *= *adjust;
  float s = 1.f;
  if (a > 0.f && isfinite(a)) {
    int e;
    frexpf(target / a, &e);  // target/a = f * 2^e, f in [0.5,1)
    s = ldexpf(1.f, e - 1);  // largest power of two <= target/a
  }
  scale2[0] = s;
  scale2[1] = 1.f / s;
}

// Overflow feedback for the fp16 gradient scale (the AMP recipe, device-side so that it lives inside a CUDA graph):
// `probe` is the most downstream gradient of the sweep (the first conv's weight gradient) - an fp16 overflow anywhere
// upstream reaches it as inf/nan.  adjust[0] is the multiplier applied to the scale target of the NEXT sweep:
// /16 on overflow (floor 2^-20), x2 after 256 clean sweeps (cap 1).  adjust[1] counts clean sweeps, adjust[2] overflows.
__global__ void grad_scale_feedback_kernel(const float* __restrict__ probe, long long count, float* adjust) {
  pdl_sync();
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  int b = 0;
  for (long long i = threadIdx.x; i < count; i += blockDim.x) b |= !isfinite(probe[i]);
  if (b) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (bad) {
      adjust[0] = fmaxf(adjust[0] * 0.0625f, 9.5367431640625e-07f);
      adjust[1] = 0.f;
      adjust[2] += 1.f;
    } else {
      adjust[1] += 1.f;
      if (adjust[1] >= 256.f) {
        adjust[1] = 0.f;
        adjust[0] = fminf(adjust[0] * 2.f, 1.f);
      }
    }
  }
}

// ------------------------------------------------------------------ perceptual (feature) loss
// The feature tensor holds 2*nb images: [0, nb) are the features of the generated patches, [nb, 2*nb) of the targets.
// flags: 1 = `g` already holds the gradient that arrives from the layers behind this tensor (add to it), 2 = the tensor is a
// ReLU output (mask the gradient with f > 0), 4 = the tensor is a tap (contributes taps * sum (f - t)^2 to the loss and
// gmul * (f - t) to the gradient).  The squared-difference sum is reduced without float atomics: per-block partial, and the
// block that finishes last adds all partials in index order (double) to *loss, so the value is the same on every run.
template <int DT>
__global__ void feature_mse_kernel(ActView f, int nb, float gmul, int flags, ActView g, float* __restrict__ partial,
                                   unsigned* counter, float* loss, float loss_mul) {
  pdl_sync();
  const long long hw = (long long)f.h * f.w;
  const int planes = f.c / 8;
  const long long total = (long long)nb * planes * hw;
  const bool acc_in = flags & 1, relu = flags & 2, tap = flags & 4;
  float sq = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const int pl = (int)((i / hw) % planes);
    const int ni = (int)(i / (hw * planes));
    float a[8], b[8], gv[8];
    unpack8<DT>(*chunk_ptr(f, ni, pl, pix), a);
    if (tap) unpack8<DT>(*chunk_ptr(f, ni + nb, pl, pix), b);
    if (g.ptr && acc_in) unpack8<DT>(*chunk_ptr(g, ni, pl, pix), gv);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = (g.ptr && acc_in) ? gv[k] : 0.f;
      if (tap) {
        const float d = a[k] - b[k];
        sq = fmaf(d, d, sq);
        v = fmaf(gmul, d, v);
      }
      gv[k] = (relu && !(a[k] > 0.f)) ? 0.f : v;
    }
    if (g.ptr) *chunk_ptr(g, ni, pl, pix) = pack8<DT>(gv);
  }
  if (!tap) return;
  __shared__ double red[kEwThreads / 32];
  __shared__ bool last;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)sq;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) t += red[w2];
    partial[blockIdx.x] = (float)t;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t = 0.0;
  for (int b2 = threadIdx.x; b2 < (int)gridDim.x; b2 += blockDim.x) t += (double)__ldcg(&partial[b2]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w2 = 0; w2 < kEwThreads / 32; ++w2) s += red[w2];
    *loss += (float)(s * (double)loss_mul);
    *counter = 0u;
  }
}

// 2x2 stride-2 max pooling (floor mode) and its transpose; the gradient goes to the FIRST maximum of the window in
// row-major order, as the tensor library's max_pool2d does
template <int DT>
__global__ void maxpool2_kernel(ActView x, ActView y) {
  pdl_sync();
  const long long ohw = (long long)y.h * y.w;
  const int planes = y.c / 8;
  const long long total = (long long)y.n * planes * ohw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long opix = i % ohw;
    const int pl = (int)((i / ohw) % planes);
    const int ni = (int)(i / (ohw * planes));
    const int oy = (int)(opix / y.w), ox = (int)(opix - (long long)oy * y.w);
    const long long p00 = (long long)(2 * oy) * x.w + 2 * ox;
    float m[8], v[8];
    unpack8<DT>(*chunk_ptr(x, ni, pl, p00), m);
    const long long off[3] = {p00 + 1, p00 + x.w, p00 + x.w + 1};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      unpack8<DT>(*chunk_ptr(x, ni, pl, off[q]), v);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = (v[k] > m[k]) ? v[k] : m[k];
    }
    *chunk_ptr(y, ni, pl, opix) = pack8<DT>(m);
  }
}

template <int DT>
__global__ void maxpool2_bwd_kernel(ActView x, ActView dy, ActView dx) {
  pdl_sync();
  const long long ohw = (long long)dy.h * dy.w;
  const int planes = dy.c / 8;
  const long long total = (long long)dx.n * planes * ohw;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long opix = i % ohw;
    const int pl = (int)((i / ohw) % planes);
    const int ni = (int)(i / (ohw * planes));
    const int oy = (int)(opix / dy.w), ox = (int)(opix - (long long)oy * dy.w);
    const long long p00 = (long long)(2 * oy) * x.w + 2 * ox;
    const long long off[4] = {p00, p00 + 1, p00 + x.w, p00 + x.w + 1};
    float m[8], v[8], gy[8], o[4][8];
    int arg[8];
    unpack8<DT>(*chunk_ptr(x, ni, pl, off[0]), m);
#pragma unroll
    for (int k = 0; k < 8; ++k) arg[k] = 0;
#pragma unroll
    for (int q = 1; q < 4; ++q) {
      unpack8<DT>(*chunk_ptr(x, ni, pl, off[q]), v);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (v[k] > m[k]) { m[k] = v[k]; arg[k] = q; }
    }
    unpack8<DT>(*chunk_ptr(dy, ni, pl, opix), gy);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[q][k] = (arg[k] == q) ? gy[k] : 0.f;
      *chunk_ptr(dx, ni, pl, off[q]) = pack8<DT>(o[q]);
    }
    // odd sizes: the last row / column is outside every window
    if ((x.w & 1) && ox == dy.w - 1) {
      *chunk_ptr(dx, ni, pl, p00 + 2) = zero4;
      *chunk_ptr(dx, ni, pl, p00 + x.w + 2) = zero4;
    }
    if ((x.h & 1) && oy == dy.h - 1) {
      *chunk_ptr(dx, ni, pl, p00 + 2 * (long long)x.w) = zero4;
      *chunk_ptr(dx, ni, pl, p00 + 2 * (long long)x.w + 1) = zero4;
      if ((x.w & 1) && ox == dy.w - 1) *chunk_ptr(dx, ni, pl, p00 + 2 * (long long)x.w + 2) = zero4;
    }
  }
}

}  // namespace pbt

using namespace pbt;

#define DISPATCH_DT(dtype, ...)                              \
  do {                                                       \
    if ((dtype) == PBT_BF16) { constexpr int DT = 0; __VA_ARGS__; } \
    else if ((dtype) == PBT_FP16) { constexpr int DT = 1; __VA_ARGS__; } \
    else { pbt::set_last_error("bad dtype"); return PBT_ERR_ARG; } \
  } while (0)

static bool act_ok(const pbt_act_t& t) { return t.ptr && aligned16(t.ptr) && t.c % 8 == 0 && t.img_stride % 8 == 0 && t.n > 0 && t.h > 0 && t.w > 0; }

extern "C" int pbt_nchw_to_p8(const void* x, int32_t src_is_half, int32_t n, int32_t c, int32_t h, int32_t w,
                              const pbt_act_t* out, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(x && out && act_ok(*out), "nchw_to_p8: bad tensors");
  PBT_REQUIRE(out->n == n && out->h == h && out->w == w && out->c >= c, "nchw_to_p8: shape mismatch");
  const long long items = (long long)n * (out->c / 8) * h * w;
  DISPATCH_DT(dtype, {
    if (src_is_half) pbt::launch(nchw_to_p8_kernel<DT, true>, ew_grid(items), kEwThreads, 0, st, x, n, c, h, w, view(*out));
    else pbt::launch(nchw_to_p8_kernel<DT, false>, ew_grid(items), kEwThreads, 0, st, x, n, c, h, w, view(*out));
  });
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_p8_to_nchw_f32(const pbt_act_t* in, int32_t c, float* out, float mul, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(in && out && act_ok(*in) && c > 0 && c <= in->c, "p8_to_nchw: bad tensors");
  const long long items = (long long)in->n * ((c + 7) / 8) * in->h * in->w;
  DISPATCH_DT(dtype, pbt::launch(p8_to_nchw_kernel<DT>, ew_grid(items), kEwThreads, 0, st, view(*in), c, out, mul));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_p8f_to_nchw_f32(const float* in, int32_t n, int32_t c_total, int32_t c, int32_t h, int32_t w, float* out,
                                   void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(in && out && c_total % 8 == 0 && c <= c_total && n > 0, "p8f_to_nchw: bad tensors");
  const long long items = (long long)n * ((c + 7) / 8) * h * w;
  pbt::launch(p8f_to_nchw_kernel, ew_grid(items), kEwThreads, 0, st, in, n, c_total, c, h, w, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_u8hwc_to_p8(const uint8_t* img, int32_t n, int32_t h, int32_t w, int32_t c, const pbt_act_t* out,
                               int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(img && out && act_ok(*out), "u8hwc_to_p8: bad tensors");
  PBT_REQUIRE(out->n == n && out->h == h && out->w == w && out->c >= c, "u8hwc_to_p8: shape mismatch");
  const long long items = (long long)n * (out->c / 8) * h * w;
  DISPATCH_DT(dtype, pbt::launch(u8hwc_to_p8_kernel<DT>, ew_grid(items), kEwThreads, 0, st, img, n, h, w, c, view(*out)));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_nchw_to_u8hwc(const float* y, int32_t n, int32_t c, int32_t h, int32_t w, uint8_t* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(y && out && n > 0 && c > 0, "nchw_to_u8hwc: bad tensors");
  pbt::launch(nchw_to_u8hwc_kernel, ew_grid((long long)n * h * w), kEwThreads, 0, st, y, n, c, h, w, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_u8hwc_to_norm_chw(const uint8_t* img, int32_t h, int32_t w, int32_t c, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(img && out && h > 0 && w > 0 && c > 0, "u8hwc_to_norm_chw: bad tensors");
  pbt::launch(u8hwc_to_norm_chw_kernel, ew_grid((long long)h * w * c), kEwThreads, 0, st, img, h, w, c, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_mask_dilate7(const uint8_t* mask, int32_t h, int32_t w, uint8_t* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(mask && out && h > 0 && w > 0, "mask_dilate7: bad tensors");
  pbt::launch(mask_dilate7_kernel, ew_grid((long long)h * w), kEwThreads, 0, st, mask, h, w, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_zero_border(const pbt_act_t* t, int32_t valid_h, int32_t valid_w, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(t && act_ok(*t) && valid_h >= 0 && valid_h <= t->h && valid_w >= 0 && valid_w <= t->w, "zero_border: bad arguments");
  if (valid_h == t->h && valid_w == t->w) return PBT_OK;
  pbt::launch(zero_border_kernel, ew_grid3(t->h * t->w, t->c / 8, t->n), kEwThreads, 0, st, view(*t), valid_h, valid_w);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_p8s2d_to_nchw_f32(const pbt_act_t* in, int32_t cpp, int32_t c, float* out, const float* mul_dev, int32_t dtype,
                                     void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(in && act_ok(*in) && out && cpp > 0 && c > 0 && c <= cpp && 4 * cpp <= in->c, "p8s2d_to_nchw: bad arguments");
  const long long items = (long long)in->n * c * 4 * in->h * in->w;
  DISPATCH_DT(dtype, pbt::launch(p8s2d_to_nchw_kernel<DT>, ew_grid(items), kEwThreads, 0, st, view(*in), cpp, c, out, mul_dev));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_mask_erode7(const uint8_t* mask, int32_t n, int32_t h, int32_t w, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(mask && out && n > 0 && h > 0 && w > 0, "mask_erode7: bad tensors");
  pbt::launch(mask_erode7_kernel, ew_grid((long long)n * h * w), kEwThreads, 0, st, mask, n, h, w, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_composite_to_u8(const float* y, const uint8_t* frame, int32_t c, const float* mask, int32_t n, int32_t h,
                                   int32_t w, uint8_t* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(y && out && n > 0 && h > 0 && w > 0, "composite_to_u8: bad tensors");
  PBT_REQUIRE(!mask || (frame && c >= 3), "composite_to_u8: a mask needs the uint8 frame with >= 3 channels");
  pbt::launch(composite_to_u8_kernel, ew_grid((long long)n * h * w), kEwThreads, 0, st, y, frame, c, mask, n, h, w, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_norm_finalize(const float* partial, int32_t n, int32_t tiles, int32_t c, int64_t count_per_image, float eps,
                                 int32_t batch_mode, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float momentum, float* scale, float* shift, float* mean_out,
                                 float* rstd_out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(partial && scale && shift && n > 0 && tiles > 0 && c > 0 && count_per_image > 0, "norm_finalize: bad arguments");
  // NOTE: `partial` is consumed (stage 1 folds chunks of tiles in place when there are many tiles)
  int slots = tiles, stride = 1;
  // measured: one block walking >64 tiles is latency bound (25 us at 510 tiles vs 3+3.5 us in two stages); batch
  // statistics walk n * tiles slots in ONE block per channel group (101 us at 80 x 25), so fold every image first
  const bool fold_batch = batch_mode && tiles > 1 && (long long)n * tiles > 64;
  if (tiles > 64 || fold_batch) {
    int chunk = ceil_div(tiles, 64);
    if (chunk < 16) chunk = 16;
    if (fold_batch && tiles <= 64) chunk = tiles;
    slots = ceil_div(tiles, chunk);
    stride = chunk;
    dim3 g1(ceil_div(c, 32), slots, n);
    pbt::launch(norm_reduce_partials_kernel, g1, 256, 0, st, const_cast<float*>(partial), tiles, c, chunk);
    PBT_CUDA_CHECK(cudaGetLastError());
  }
  dim3 grid(ceil_div(c, 32), batch_mode ? 1 : n);
  pbt::launch(norm_finalize_kernel, grid, 256, 0, st, partial, n, slots, stride, tiles, c, count_per_image, eps, batch_mode, gamma, beta,
                                            running_mean, running_var, momentum, scale, shift, mean_out, rstd_out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_norm_apply(const pbt_norm_apply_desc_t* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(d && act_ok(d->x), "norm_apply: bad input");
  PBT_REQUIRE((d->scale == nullptr) == (d->shift == nullptr), "norm_apply: scale/shift must come together");
  NormApplyK p;
  memset(&p, 0, sizeof(p));
  p.x = view(d->x);
  auto same = [&](const pbt_act_t& t) { return t.n == d->x.n && t.c >= d->x.c && t.h == d->x.h && t.w == d->x.w && act_ok(t); };
  if (d->out.ptr) { PBT_REQUIRE(same(d->out), "norm_apply: out shape mismatch"); p.out = view(d->out); }
  if (d->out_relu.ptr) { PBT_REQUIRE(same(d->out_relu), "norm_apply: out_relu shape mismatch"); p.out_relu = view(d->out_relu); }
  if (d->out_s2d.ptr) {
    PBT_REQUIRE(d->x.h % 2 == 0 && d->x.w % 2 == 0 && d->out_s2d.h == d->x.h / 2 && d->out_s2d.w == d->x.w / 2 &&
                    d->out_s2d.c >= 4 * d->x.c && d->out_s2d.n == d->x.n && act_ok(d->out_s2d),
                "norm_apply: out_s2d shape mismatch");
    p.out_s2d = view(d->out_s2d);
  }
  p.scale = d->scale; p.shift = d->shift; p.per_channel = d->per_channel; p.act = d->act;
  p.residual32 = d->residual32; p.out32 = d->out32;
  if (d->partial) {
    PBT_REQUIRE(d->scale && d->shift && !d->per_channel && d->tiles > 0 && d->tiles <= 64 && d->count > 0,
                "norm_apply: fused finalize needs scale / shift outputs, per-image statistics, 1..64 tiles and a pixel count");
    p.partial = d->partial; p.scale_out = const_cast<float*>(d->scale); p.shift_out = const_cast<float*>(d->shift);
    p.tiles = d->tiles; p.inv_count = 1.0 / (double)d->count; p.eps = d->eps;
    p.scale = nullptr; p.shift = nullptr;
  }
  if (d->residual16.ptr) {
    PBT_REQUIRE(!d->residual32 && same(d->residual16), "norm_apply: residual16 shape mismatch (or both residuals given)");
    p.res16 = view(d->residual16);
  }
  PBT_REQUIRE(d->x.n <= 65535 && (long long)d->x.h * d->x.w < (1ll << 31), "norm_apply: tensor too large for one launch");
  DISPATCH_DT(d->dtype, pbt::launch(norm_apply_kernel<DT>, ew_grid3(d->x.h * d->x.w, d->x.c / 8, d->x.n), kEwThreads, 0, st, p));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_upsample2x(const pbt_act_t* in, const pbt_act_t* out, const float* scale, const float* shift, int32_t act,
                              int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(in && out && act_ok(*in) && act_ok(*out), "upsample2x: bad tensors");
  PBT_REQUIRE(out->h == 2 * in->h && out->w == 2 * in->w && out->n == in->n && out->c >= in->c, "upsample2x: shape mismatch");
  PBT_REQUIRE((scale == nullptr) == (shift == nullptr), "upsample2x: scale/shift must come together");
  PBT_REQUIRE(in->n <= 65535 && in->c / 8 <= 65535, "upsample2x: too many images / planes for one launch");
  dim3 grid(ceil_div(((out->w + 1) / 2) * ((out->h + 1) / 2), kEwThreads), in->c / 8, in->n);
  DISPATCH_DT(dtype, pbt::launch(upsample2x_kernel<DT>, grid, kEwThreads, 0, st, view(*in), view(*out), scale, shift, act));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_upsample2x_bwd(const pbt_act_t* gout, const pbt_act_t* gin16, float* gin32, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(gout && act_ok(*gout) && gout->h % 2 == 0 && gout->w % 2 == 0, "upsample2x_bwd: bad gout");
  const int ih = gout->h / 2, iw = gout->w / 2;
  ActView g16;
  memset(&g16, 0, sizeof(g16));
  if (gin16 && gin16->ptr) {
    PBT_REQUIRE(act_ok(*gin16) && gin16->h == ih && gin16->w == iw && gin16->n == gout->n && gin16->c >= gout->c,
                "upsample2x_bwd: gin16 shape mismatch");
    g16 = view(*gin16);
  }
  PBT_REQUIRE(g16.ptr || gin32, "upsample2x_bwd: no output");
  DISPATCH_DT(dtype, pbt::launch(upsample2x_bwd_kernel<DT>, ew_grid3(ih * iw, gout->c / 8, gout->n), kEwThreads, 0, st, view(*gout), g16, gin32, ih, iw));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

static int fill_norm_bwd(const pbt_norm_bwd_desc_t* d, NormBwdK& p) {
  PBT_REQUIRE(d && act_ok(d->x) && d->scale && d->shift && d->sums, "norm_bwd: bad arguments");
  memset(&p, 0, sizeof(p));
  p.x = view(d->x);
  if (d->ga.ptr) {
    if (d->ga_is_s2d)
      PBT_REQUIRE(act_ok(d->ga) && d->ga.h * 2 == d->x.h && d->ga.w * 2 == d->x.w && d->ga.c >= 4 * d->x.c && d->ga.n == d->x.n,
                  "norm_bwd: s2d grad shape mismatch");
    else
      PBT_REQUIRE(act_ok(d->ga) && d->ga.h == d->x.h && d->ga.w == d->x.w && d->ga.c >= d->x.c && d->ga.n == d->x.n,
                  "norm_bwd: grad shape mismatch");
    p.ga = view(d->ga);
  }
  if (d->gb16.ptr) {
    PBT_REQUIRE(act_ok(d->gb16) && d->gb16.h == d->x.h && d->gb16.w == d->x.w && d->gb16.c >= d->x.c && d->gb16.n == d->x.n,
                "norm_bwd: gb16 shape mismatch");
    p.gb16 = view(d->gb16);
  }
  p.mean = d->scale; p.shift = d->shift; p.per_channel = d->per_channel; p.act = d->act;
  p.ga_is_s2d = d->ga_is_s2d; p.batch_mode = d->batch_mode; p.relu_mask_x = d->relu_mask_x; p.gb32 = d->gb32; p.sums = d->sums; p.kmul = d->kmul;
  PBT_REQUIRE(d->count > 0, "norm_bwd: count must be positive");
  p.inv_count = (float)(1.0 / (double)d->count);
  return PBT_OK;
}

extern "C" int pbt_norm_bwd_reduce(const pbt_norm_bwd_desc_t* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  NormBwdK p;
  int rc = fill_norm_bwd(d, p);
  if (rc) return rc;
  const long long hw = (long long)d->x.h * d->x.w;
  int chunks = (int)((hw + kEwThreads * 4 - 1) / (kEwThreads * 4));
  if (chunks < 1) chunks = 1;
  if (chunks > 64) chunks = 64;
  dim3 grid(chunks, d->x.c / 8, d->x.n);
  DISPATCH_DT(d->dtype, pbt::launch(norm_bwd_reduce_kernel<DT>, grid, kEwThreads, 0, st, p));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_norm_bwd_apply(const pbt_norm_bwd_desc_t* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  NormBwdK p;
  int rc = fill_norm_bwd(d, p);
  if (rc) return rc;
  PBT_REQUIRE(d->kmul && act_ok(d->dx) && d->dx.h == d->x.h && d->dx.w == d->x.w && d->dx.c >= d->x.c && d->dx.n == d->x.n,
              "norm_bwd_apply: dx shape mismatch");
  p.dx = view(d->dx);
  DISPATCH_DT(d->dtype, pbt::launch(norm_bwd_apply_kernel<DT>, ew_grid3(d->x.h * d->x.w, d->x.c / 8, d->x.n), kEwThreads, 0, st, p));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_norm_bwd_fused(const pbt_norm_bwd_desc_t* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  NormBwdK p;
  int rc = fill_norm_bwd(d, p);
  if (rc) return rc;
  PBT_REQUIRE(!d->batch_mode, "norm_bwd_fused: per-image statistics only (use reduce + apply for batch statistics)");
  PBT_REQUIRE((long long)d->x.h * d->x.w <= (1 << 20), "norm_bwd_fused: map too large");
  PBT_REQUIRE(d->kmul && act_ok(d->dx) && d->dx.h == d->x.h && d->dx.w == d->x.w && d->dx.c >= d->x.c && d->dx.n == d->x.n,
              "norm_bwd_fused: dx shape mismatch");
  p.dx = view(d->dx);
  dim3 grid(d->x.c / 8, d->x.n);
  const long long hw = (long long)d->x.h * d->x.w;
  const size_t slice = (size_t)hw * 32;            // x chunk + activated-gradient chunk per pixel
  const int threads = hw >= 4096 ? 512 : kEwThreads;   // one big CTA per SM when the slice fills shared memory
  if (hw >= 4096 && slice / kNbCluster + 64 <= 56 * 1024) {
    // large slices: a cluster of CTAs per slice (see norm_bwd_fused_cluster_kernel)
    const size_t part_bytes = (size_t)((hw + kNbCluster - 1) / kNbCluster) * 32;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(d->x.c / 8 * kNbCluster), (unsigned)d->x.n, 1);
    cfg.blockDim = dim3(kEwThreads, 1, 1);
    cfg.dynamicSmemBytes = part_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kNbCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (d->dtype == PBT_BF16) {
      PBT_CUDA_CHECK(cudaFuncSetAttribute(norm_bwd_fused_cluster_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_bytes));
      PBT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, norm_bwd_fused_cluster_kernel<0>, p));
    } else if (d->dtype == PBT_FP16) {
      PBT_CUDA_CHECK(cudaFuncSetAttribute(norm_bwd_fused_cluster_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_bytes));
      PBT_CUDA_CHECK(cudaLaunchKernelEx(&cfg, norm_bwd_fused_cluster_kernel<1>, p));
    } else {
      pbt::set_last_error("bad dtype");
      return PBT_ERR_ARG;
    }
  } else if (slice <= 200 * 1024) {
    if (d->dtype == PBT_BF16) {
      PBT_CUDA_CHECK(cudaFuncSetAttribute(norm_bwd_fused_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice));
      pbt::launch(norm_bwd_fused_kernel<0, true>, grid, threads, slice, st, p);
    } else if (d->dtype == PBT_FP16) {
      PBT_CUDA_CHECK(cudaFuncSetAttribute(norm_bwd_fused_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)slice));
      pbt::launch(norm_bwd_fused_kernel<1, true>, grid, threads, slice, st, p);
    } else {
      pbt::set_last_error("bad dtype");
      return PBT_ERR_ARG;
    }
  } else {
    DISPATCH_DT(d->dtype, pbt::launch(norm_bwd_fused_kernel<DT, false>, grid, kEwThreads, 0, st, p));
  }
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_head_bwd(const float* gy, const float* y, const pbt_act_t* s, const float* head_w, const float* gscale,
                            int32_t head_tanh, float* dw, float* db, const pbt_act_t* gs, float* dbias_prev, int32_t dtype,
                            void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(gy && y && s && gs && head_w && dw && db && act_ok(*s) && act_ok(*gs), "head_bwd: bad tensors");
  PBT_REQUIRE(gs->n == s->n && gs->h == s->h && gs->w == s->w && gs->c >= s->c, "head_bwd: gs shape mismatch");
  const long long items = (long long)s->n * s->h * s->w;
  int chunks = (int)((items + kEwThreads * 8 - 1) / (kEwThreads * 8));
  if (chunks < 1) chunks = 1;
  if (chunks > 2 * num_sms()) chunks = 2 * num_sms();
  dim3 grid(chunks, s->c / 8);
  DISPATCH_DT(dtype, pbt::launch(head_bwd_kernel<DT>, grid, kEwThreads, 0, st, gy, y, view(*s), head_w, gscale, head_tanh, dw, db,
                                                                      view(*gs), dbias_prev));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_channel_sum(const pbt_act_t* g, float* out, const float* inv_scale, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(g && out && act_ok(*g), "channel_sum: bad tensors");
  const long long items = (long long)g->n * g->h * g->w;
  int chunks = (int)((items + kEwThreads * 8 - 1) / (kEwThreads * 8));
  if (chunks < 1) chunks = 1;
  if (chunks > 2 * num_sms()) chunks = 2 * num_sms();
  dim3 grid(chunks, g->c / 8);
  DISPATCH_DT(dtype, pbt::launch(channel_sum_kernel<DT>, grid, kEwThreads, 0, st, view(*g), out, inv_scale));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_absmax_f32(const float* g, int64_t count, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(g && out && count > 0, "absmax: bad arguments");
  PBT_CUDA_CHECK(cudaMemsetAsync(out, 0, sizeof(float), st));
  pbt::launch(absmax_kernel, ew_grid(count), kEwThreads, 0, st, g, count, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_grad_scale_feedback(const float* probe, int64_t count, float* adjust, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(probe && adjust && count > 0, "grad_scale_feedback: bad arguments");
  pbt::launch(grad_scale_feedback_kernel, 1, 1024, 0, st, probe, count, adjust);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_make_grad_scale(const float* amax, float target, float* scale2, const float* adjust, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(amax && scale2 && target > 0.f, "make_grad_scale: bad arguments");
  pbt::launch(make_grad_scale_kernel, 1, 1, 0, st, amax, target, scale2, adjust);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_feature_mse(const pbt_act_t* f, int32_t n_pairs, float grad_mul, int32_t flags, const pbt_act_t* g,
                               float* partial, uint32_t* counter, float* loss, float loss_mul, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(f && act_ok(*f) && n_pairs > 0 && f->n >= 2 * n_pairs && (flags & ~7) == 0, "feature_mse: bad feature tensor");
  PBT_REQUIRE(!(flags & 4) || (partial && counter && loss), "feature_mse: a tap needs partial / counter / loss");
  pbt_act_t none;
  memset(&none, 0, sizeof(none));
  if (g && g->ptr) {
    PBT_REQUIRE(act_ok(*g) && g->n >= n_pairs && g->c == f->c && g->h == f->h && g->w == f->w, "feature_mse: gradient shape");
  } else {
    PBT_REQUIRE(flags & 4, "feature_mse: nothing to do");
    g = &none;
  }
  const long long items = (long long)n_pairs * (f->c / 8) * f->h * f->w;
  int grid = ew_grid(items);
  if (grid > 4096) grid = 4096;          // `partial` holds one float per block
  DISPATCH_DT(dtype, pbt::launch(feature_mse_kernel<DT>, grid, kEwThreads, 0, st, view(*f), n_pairs, grad_mul, flags, view(*g), partial,
                                 counter, loss, loss_mul));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_maxpool2(const pbt_act_t* x, const pbt_act_t* y, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(x && y && act_ok(*x) && act_ok(*y) && y->n == x->n && y->c == x->c && y->h == x->h / 2 && y->w == x->w / 2,
              "maxpool2: bad tensors");
  const long long items = (long long)y->n * (y->c / 8) * y->h * y->w;
  DISPATCH_DT(dtype, pbt::launch(maxpool2_kernel<DT>, ew_grid(items), kEwThreads, 0, st, view(*x), view(*y)));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_maxpool2_bwd(const pbt_act_t* x, const pbt_act_t* dy, const pbt_act_t* dx, int32_t dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(x && dy && dx && act_ok(*x) && act_ok(*dy) && act_ok(*dx), "maxpool2_bwd: bad tensors");
  PBT_REQUIRE(dx->n <= x->n && dy->n >= dx->n && dx->c == x->c && dy->c == x->c && dx->h == x->h && dx->w == x->w &&
                  dy->h == x->h / 2 && dy->w == x->w / 2, "maxpool2_bwd: shapes");
  const long long items = (long long)dx->n * (dy->c / 8) * dy->h * dy->w;
  DISPATCH_DT(dtype, pbt::launch(maxpool2_bwd_kernel<DT>, ew_grid(items), kEwThreads, 0, st, view(*x), view(*dy), view(*dx)));
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

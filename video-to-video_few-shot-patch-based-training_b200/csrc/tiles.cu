// tiles.cu — tiled inference mode of the reference driver (generator.py:427-565 `process_large_image`):
// the frame is cut into patch_size windows around sampled mask pixels, every window goes through the generator,
// and the outputs are blended with Gaussian weights, normalised and composited with the input through the mask.
//   tile_gather : window b of the frame -> batch slot b, windows smaller than the patch (frame border) are CENTRED in a
//                 zero patch (`ensure_valid_patch_size`, :470-497)
//   tile_blend  : output[.., y0+i, x0+j] += G(patch)[.., i, j] * w[i, j] ; weights[.., y0+i, x0+j] += w[i, j]   (:536-541;
//                 note the reference adds the centred patch at the window's top-left corner, reproduced as is)
//   tile_finish : output / weights (where weights > 1e-8), then rgb*(1-mask) + output*mask                      (:553-560)
#include "internal.h"

namespace pbt {

// grid: (pixel chunks of P*P, channels, tiles)
__global__ void tile_gather_kernel(const float* __restrict__ src, int h, int w, const int* __restrict__ boxes, int patch,
                                   float* __restrict__ out, int channels) {
  pdl_sync();
  const int b = blockIdx.z, ch = blockIdx.y;
  const int y0 = boxes[4 * b + 0], y1 = boxes[4 * b + 1], x0 = boxes[4 * b + 2], x1 = boxes[4 * b + 3];
  const int hc = min(y1 - y0, patch), wc = min(x1 - x0, patch);
  const int oy = (patch - hc) / 2, ox = (patch - wc) / 2;
  const float* s = src + (long long)ch * h * w;
  float* o = out + ((long long)b * channels + ch) * patch * patch;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < patch * patch; i += gridDim.x * blockDim.x) {
    const int py = i / patch, px = i - py * patch;
    const int sy = py - oy, sx = px - ox;
    float v = 0.f;
    if (sy >= 0 && sy < hc && sx >= 0 && sx < wc) v = s[(long long)(y0 + sy) * w + x0 + sx];
    o[i] = v;
  }
}

// grid: (pixel chunks of P*P, tiles)
__global__ void tile_blend_kernel(const float* __restrict__ proc, const int* __restrict__ boxes, const int* __restrict__ widx,
                                  const float* __restrict__ wtab, int patch, int h, int w, float* __restrict__ acc,
                                  float* __restrict__ wsum) {
  pdl_sync();
  const int b = blockIdx.y;
  const int y0 = boxes[4 * b + 0], x0 = boxes[4 * b + 2];
  const float* wt = wtab + (long long)widx[b] * patch * patch;
  const float* pp = proc + (long long)b * 3 * patch * patch;
  const long long hw = (long long)h * w;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < patch * patch; i += gridDim.x * blockDim.x) {
    const int py = i / patch, px = i - py * patch;
    const int Y = y0 + py, X = x0 + px;
    if (Y >= h || X >= w) continue;
    const float wv = wt[i];
    const long long o = (long long)Y * w + X;
    atomicAdd(&acc[o], pp[i] * wv);
    atomicAdd(&acc[hw + o], pp[patch * patch + i] * wv);
    atomicAdd(&acc[2 * hw + o], pp[2 * patch * patch + i] * wv);
    atomicAdd(&wsum[o], wv);
  }
}

__global__ void tile_finish_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, const float* __restrict__ rgb,
                                   const float* __restrict__ mask, long long hw, float* __restrict__ out) {
  pdl_sync();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw; i += (long long)gridDim.x * blockDim.x) {
    const float ws = wsum[i];
    const float d = ws > 1e-8f ? ws : 1.f;
    const float m = mask[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float o = acc[c * hw + i] / d;
      out[c * hw + i] = rgb[c * hw + i] * (1.f - m) + o * m;
    }
  }
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_tile_gather(const float* src, int32_t channels, int32_t h, int32_t w, const int32_t* boxes_dev,
                               int32_t n_tiles, int32_t patch, float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n_tiles == 0) return PBT_OK;
  PBT_REQUIRE(src && boxes_dev && out && channels > 0 && channels <= 65535 && h > 0 && w > 0 && patch > 0 && n_tiles > 0 &&
                  n_tiles <= 65535, "tile_gather: bad arguments");
  dim3 grid(ceil_div(patch * patch, 256), channels, n_tiles);
  pbt::launch(tile_gather_kernel, grid, 256, 0, st, src, h, w, boxes_dev, patch, out, channels);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_tile_blend(const float* proc, const int32_t* boxes_dev, const int32_t* weight_index_dev,
                              const float* weight_table, int32_t n_tiles, int32_t patch, int32_t h, int32_t w, float* acc,
                              float* wsum, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n_tiles == 0) return PBT_OK;
  PBT_REQUIRE(proc && boxes_dev && weight_index_dev && weight_table && acc && wsum && patch > 0 && h > 0 && w > 0 &&
                  n_tiles > 0 && n_tiles <= 65535, "tile_blend: bad arguments");
  dim3 grid(ceil_div(patch * patch, 256), n_tiles);
  pbt::launch(tile_blend_kernel, grid, 256, 0, st, proc, boxes_dev, weight_index_dev, weight_table, patch, h, w, acc, wsum);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_tile_finish(const float* acc, const float* wsum, const float* rgb, const float* mask, int32_t h, int32_t w,
                               float* out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(acc && wsum && rgb && mask && out && h > 0 && w > 0, "tile_finish: bad arguments");
  const long long hw = (long long)h * w;
  long long blocks = (hw + 255) / 256;
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pbt::launch(tile_finish_kernel, (unsigned)blocks, 256, 0, st, acc, wsum, rgb, mask, hw, out);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

// gather.cu — masked random-patch gather: one launch cuts every patch of a batch out of the
// device-resident keyframes (pre, post, guide channels) and writes them already concatenated.
// Bit-exact restatement of StyleTransferDataset._cut_patch (reference src/data/dataset.py:209-232)
// + default_collate + torch.cat (lightning_model.py:211-221): pure fp32 copies with the reference's
// clamping quirks (last row/column never included, short patches zero-padded from the top-left).
#include "internal.h"

namespace pbt {

constexpr int kMaxSrc = 8;
struct GatherOut {
  float* out[kMaxSrc];
  int ch_off[kMaxSrc];
  int ch_total[kMaxSrc];
};

// one thread = 4 consecutive x of one row; block = (float4 groups of a row) x (rows); grid = (1, plane groups, patch): no per-element index arithmetic (the first version spent more instructions on 64-bit div/mod than on
// the copy and reached 0.53 of the HBM peak on the 10 GB stress case), patch geometry computed once per thread from
// block-uniform values.  float4 stores, scalar loads (window columns are not 16-byte aligned in general).
__global__ void patch_gather_kernel(const float* const* __restrict__ src_ptrs, int n_src, int n_images, int ch,
                                    const int* __restrict__ img_hw, const int* __restrict__ pos, int n_patches, int P,
                                    GatherOut o) {
  pdl_sync();
  const int gx = threadIdx.x;
  const int b = blockIdx.z;
  const int half = P / 2;
  const int img = pos[b * 3 + 0], y = pos[b * 3 + 1], x = pos[b * 3 + 2];
  const int H = img_hw[img * 2 + 0], W = img_hw[img * 2 + 1];
  const int hn = max(0, y - half), hx = min(y + half, H - 1);
  const int xn = max(0, x - half), xx = min(x + half, W - 1);
  const int rows = hx - hn, cols = xx - xn;
  const int planes = n_src * ch;
  // blockIdx.y selects a group of (source, channel) planes; the block walks them and the rows of the patch
  for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
    const int s = pl / ch, c = pl - s * ch;
    const float* src = src_ptrs[(long long)s * n_images + img] + (long long)c * H * W + (long long)hn * W + xn;
    float* dst = o.out[s] + ((long long)b * o.ch_total[s] + o.ch_off[s] + c) * P * P + gx * 4;
    for (int py = threadIdx.y; py < P; py += blockDim.y) {
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int px = gx * 4 + k;
        v[k] = (py < rows && px < cols) ? __ldg(src + (long long)py * W + px) : 0.f;
      }
      if ((P & 3) == 0) {
        *reinterpret_cast<float4*>(dst + (long long)py * P) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (gx * 4 + k < P) dst[(long long)py * P + k] = v[k];
      }
    }
  }
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_patch_gather(const float* const* src_ptrs, int32_t n_src, int32_t n_images, int32_t ch,
                                const int32_t* img_hw, const int32_t* pos, int32_t n_patches, int32_t patch,
                                float* const* outs, const int32_t* out_ch_off, const int32_t* out_ch_total, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (n_patches == 0) return PBT_OK;  // empty batch: nothing to do (pointers may legitimately be null)
  PBT_REQUIRE(n_patches > 0, "patch_gather: negative patch count");
  PBT_REQUIRE(src_ptrs && img_hw && pos && outs && out_ch_off && out_ch_total, "patch_gather: null argument");
  PBT_REQUIRE(n_src >= 1 && n_src <= kMaxSrc, "patch_gather: n_src must be in [1,8]");
  PBT_REQUIRE(n_images > 0 && ch > 0 && patch > 0, "patch_gather: bad sizes");
  GatherOut o;
  for (int s = 0; s < n_src; ++s) {
    PBT_REQUIRE(outs[s] != nullptr, "patch_gather: null output");
    PBT_REQUIRE(patch % 4 != 0 || aligned16(outs[s]), "patch_gather: output not 16-byte aligned");
    PBT_REQUIRE(out_ch_off[s] >= 0 && out_ch_off[s] + ch <= out_ch_total[s], "patch_gather: channel slot out of range");
    o.out[s] = outs[s];
    o.ch_off[s] = out_ch_off[s];
    o.ch_total[s] = out_ch_total[s];
  }
  const int qx = (patch + 3) / 4;                  // float4 groups per row = threads in x
  PBT_REQUIRE(qx <= 1024 && n_patches <= 65535 && n_src * ch <= 65535, "patch_gather: patch or batch too large for one launch");
  int rpb = 256 / qx;                              // rows per block
  if (rpb < 1) rpb = 1;
  if (rpb > patch) rpb = patch;
  // enough blocks to fill the GPU, as much work per block as that allows (a block per plane only for small batches)
  int ygroups = ceil_div(8 * num_sms(), n_patches);
  if (ygroups > n_src * ch) ygroups = n_src * ch;
  if (ygroups < 1) ygroups = 1;
  dim3 block((unsigned)qx, (unsigned)rpb), grid(1u, (unsigned)ygroups, (unsigned)n_patches);
  pbt::launch(patch_gather_kernel, grid, block, 0, st, src_ptrs, n_src, n_images, ch, img_hw, pos, n_patches, patch, o);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

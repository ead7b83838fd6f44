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

// one thread = 4 consecutive x of one (patch, source, channel, row); float4 store, scalar coalesced loads
__global__ void patch_gather_kernel(const float* const* __restrict__ src_ptrs, int n_src, int n_images, int ch,
                                    const int* __restrict__ img_hw, const int* __restrict__ pos, int n_patches, int P,
                                    GatherOut o) {
  const int qx = (P + 3) / 4;  // float4 groups per row
  const long long per_patch = (long long)n_src * ch * P * qx;
  const long long total = per_patch * n_patches;
  const int half = P / 2;
  const bool vec_ok = (P % 4) == 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int gx = (int)(r % qx); r /= qx;
    const int py = (int)(r % P); r /= P;
    const int c = (int)(r % ch); r /= ch;
    const int s = (int)(r % n_src); r /= n_src;
    const int b = (int)r;
    const int img = pos[b * 3 + 0], y = pos[b * 3 + 1], x = pos[b * 3 + 2];
    const int H = img_hw[img * 2 + 0], W = img_hw[img * 2 + 1];
    const int hn = max(0, y - half), hx = min(y + half, H - 1);
    const int xn = max(0, x - half), xx = min(x + half, W - 1);
    const int rows = hx - hn, cols = xx - xn;
    const float* src = src_ptrs[(long long)s * n_images + img] + (long long)c * H * W;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int px = gx * 4 + k;
      v[k] = (py < rows && px < cols) ? src[(long long)(hn + py) * W + (xn + px)] : 0.f;
    }
    float* dst = o.out[s] + (((long long)b * o.ch_total[s] + o.ch_off[s] + c) * P + py) * P + gx * 4;
    if (vec_ok) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (gx * 4 + k < P) dst[k] = v[k];
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
  const long long total = (long long)n_patches * n_src * ch * patch * ((patch + 3) / 4);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  patch_gather_kernel<<<(int)blocks, 256, 0, st>>>(src_ptrs, n_src, n_images, ch, img_hw, pos, n_patches, patch, o);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

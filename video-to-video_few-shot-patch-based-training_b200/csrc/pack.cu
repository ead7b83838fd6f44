// pack.cu — one launch that converts every conv parameter of the generator (fp32 or fp16 master copy,
// [co][ci][kh][kw] as nn.Conv2d stores it, reference src/models/generator.py:41,49,126,133,136,171,200) into the
// 16-bit packed operand layout of conv_igemm.cu:   w[cin_block][tap][k/8][n][k%8]
// including the two re-indexings the native path needs:
//   * space-to-depth form of the stride-2 3x3 convs (2x2 stride-1 kernel over 4*ci channels)
//   * dgrad form (input/output channels transposed, taps flipped) of the (possibly space-to-depth) kernel
//   * CTA-pair form (mode bit 2): w[cin_block][half][tap][k/8][n/2][k%8] for the cta_group::2 conv configuration
// Training re-packs after every optimiser step, so this replaces ~250 tiny tensor-library launches per step.
#include "internal.h"
#include "ptx.cuh"

namespace pbt {

__device__ __forceinline__ float load_src(const pbt_pack_job_t& j, int o, int i, int y, int x) {
  const long long idx = (((long long)o * j.ci + i) * j.kh + y) * j.kw + x;
  return j.src_is_half ? __half2float(static_cast<const __half*>(j.w)[idx]) : static_cast<const float*>(j.w)[idx];
}

// forward kernel V[o][i][ty][tx] (plain or space-to-depth view of the source)
__device__ __forceinline__ float virt(const pbt_pack_job_t& j, int o, int i, int ty, int tx) {
  if (!(j.mode & 1)) return load_src(j, o, i, ty, tx);
  if (j.mode & 8) {
    // 4x4 stride-2 pad-1 kernel (critic, reference src/models/discriminator.py:44-76) as a 3x3 stride-1 pad-1 kernel over
    // the space-to-depth input with `cpp` channels per phase: output row r reads input rows 2r-1..2r+2 = block r-1 phase 1
    // (ky 0), block r phases 0/1 (ky 1/2), block r+1 phase 0 (ky 3)
    const int cpp = j.reserved ? j.reserved : j.ci;
    const int ph = i / cpp, c = i - ph * cpp;
    if (ph >= 4 || c >= j.ci) return 0.f;
    const int py = ph >> 1, px = ph & 1;
    const int ky = ty == 0 ? (py == 1 ? 0 : -1) : (ty == 1 ? 1 + py : (py == 0 ? 3 : -1));
    const int kx = tx == 0 ? (px == 1 ? 0 : -1) : (tx == 1 ? 1 + px : (px == 0 ? 3 : -1));
    if (ky < 0 || kx < 0) return 0.f;
    return load_src(j, o, c, ky, kx);
  }
  const int ph = i / j.ci, c = i - ph * j.ci;
  const int py = ph >> 1, px = ph & 1;
  // (s2d tap, phase) -> original 3x3 tap: (0,1)->0, (1,0)->1, (1,1)->2, (0,0)-> no contribution
  const int dy = ty == 0 ? (py == 1 ? 0 : -1) : (py == 0 ? 1 : 2);
  const int dx = tx == 0 ? (px == 1 ? 0 : -1) : (px == 0 ? 1 : 2);
  if (dy < 0 || dx < 0) return 0.f;
  return load_src(j, o, c, dy, dx);
}

// tap-pair layout (mode bit 4): first-layer kernels with ci <= 8.  [tap = ky * ceil(kw/2) + j][k/8 = g][n][k%8 = c] holds
// w[n][c][ky][2j + g] (zero where c >= ci or 2j + g >= kw): one K = 16 MMA covers two horizontally adjacent taps.
__device__ __forceinline__ void pack_tap_pairs(const pbt_pack_job_t& j) {
  const int kwp = (j.kw + 1) / 2;
  const int taps = j.kh * kwp;
  const long long total = (long long)taps * 16 * j.n_out;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int per_tap = 16 * j.n_out;
    const int tap = (int)(e / per_tap);
    const int r2 = (int)(e - (long long)tap * per_tap);
    const int g = r2 / (j.n_out * 8);
    const int r3 = r2 - g * j.n_out * 8;
    const int n = r3 >> 3, c = r3 & 7;
    const int ky = tap / kwp, dx = 2 * (tap - ky * kwp) + g;
    float v = 0.f;
    if (n < min(j.n_keep, j.co) && c < j.ci && dx < j.kw) v = load_src(j, n, c, ky, dx);
    if (j.dtype == PBT_BF16) static_cast<__nv_bfloat16*>(j.dst)[e] = __float2bfloat16_rn(v);
    else static_cast<__half*>(j.dst)[e] = __float2half_rn(v);
  }
}

__global__ void pack_weights_kernel(const pbt_pack_job_t* __restrict__ jobs) {
  pdl_sync();
  const pbt_pack_job_t j = jobs[blockIdx.y];
  if (j.mode & 16) {
    pack_tap_pairs(j);
    return;
  }
  const bool s2d = j.mode & 1, dgrad = j.mode & 2;
  const bool s2d4 = s2d && (j.mode & 8);
  const int vkh = s2d ? (s2d4 ? 3 : 2) : j.kh, vkw = s2d ? (s2d4 ? 3 : 2) : j.kw;
  const int vo = j.co, vi = s2d ? 4 * (s2d4 && j.reserved ? j.reserved : j.ci) : j.ci;
  const int po = dgrad ? vi : vo, pi = dgrad ? vo : vi;  // packed conv: N = po rows, K = pi channels
  const int taps = vkh * vkw;
  const int n_lim = min(j.n_keep, po);
  const long long total = (long long)taps * j.k_pad * j.n_out;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    // destination order: [cb][tap][k8][n][8]
    const int full = j.blk_c * j.n_out * taps;          // elements of one full channel block
    const int cb = (int)(e / full);
    const int r = (int)(e - (long long)cb * full);
    const int kc = min(j.blk_c, j.k_pad - cb * j.blk_c);  // channels in this block
    const int per_tap = kc * j.n_out;
    const int tap = r / per_tap;
    if (tap >= taps) continue;                           // tail of a short last block
    const int r2 = r - tap * per_tap;
    const int k8 = r2 / (j.n_out * 8);
    const int r3 = r2 - k8 * j.n_out * 8;
    const int n = r3 >> 3, kk = r3 & 7;
    const int k = cb * j.blk_c + k8 * 8 + kk;
    const int ty = tap / vkw, tx = tap - ty * vkw;
    float v = 0.f;
    if (n < n_lim && k < pi) v = dgrad ? virt(j, k, n, vkh - 1 - ty, vkw - 1 - tx) : virt(j, n, k, ty, tx);
    long long off;
    if (j.mode & 4) {
      // CTA-pair layout [cb][half][tap][k8][n/2][8]: each CTA of a pair streams the taps of ITS half of the output columns
      const int nh = j.n_out >> 1, half = n / nh, nn = n - half * nh;
      off = (long long)cb * full + (long long)half * taps * kc * nh + (long long)tap * kc * nh + (long long)k8 * nh * 8 + nn * 8 + kk;
    } else {
      off = (long long)cb * full + (long long)tap * per_tap + (long long)k8 * j.n_out * 8 + n * 8 + kk;
    }
    if (j.dtype == PBT_BF16) static_cast<__nv_bfloat16*>(j.dst)[off] = __float2bfloat16_rn(v);
    else static_cast<__half*>(j.dst)[off] = __float2half_rn(v);
  }
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_pack_weights(const pbt_pack_job_t* jobs_dev, int32_t n_jobs, int64_t max_elems, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(jobs_dev && n_jobs > 0 && n_jobs <= 65535 && max_elems > 0, "pack_weights: bad arguments");
  long long bx = (max_elems + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, (unsigned)n_jobs);
  pbt::launch(pack_weights_kernel, grid, 256, 0, st, jobs_dev);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

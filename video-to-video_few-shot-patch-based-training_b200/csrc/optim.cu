// optim.cu — fused optimiser tail of the G-only training step:
//   total_norm = || all gradients ||_2 ; g *= min(1, max_norm / (total_norm + 1e-6))      (clip_grad_norm_,
//                                                                          reference lightning_model.py:245-248)
//   Adam with L2 weight decay (torch.optim.Adam semantics, reference lightning_model.py:326-329,
//   config/optimizer/default.yaml:2-10):  g += wd*p ; m = lerp(m, g, 1-b1) ; v = b2*v + (1-b2)*g*g ;
//   p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// Two launches over a device table of (param, grad, m, v, count) instead of ~50 foreach / reduction launches over the
// 48 parameter tensors.  The step counter lives on the device so the whole tail is CUDA-graph capturable.
#include "internal.h"

namespace pbt {

__device__ __forceinline__ float block_sum(float v) {
  __shared__ float red[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
  if (wid == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

// grid: (chunks, jobs).  state[3 + job * chunks + chunk] = this block's sum g^2 ; state[1] (step count) += 1 once per
// launch.  No atomics: the total is later summed in a fixed order, so identical gradients give a bit-identical norm -
// data-parallel replicas, which all hold the same all-reduced gradients, therefore apply bit-identical updates.
__global__ void grad_sqnorm_kernel(const pbt_optim_job_t* __restrict__ jobs, float* __restrict__ state) {
  pdl_sync();
  const pbt_optim_job_t j = jobs[blockIdx.y];
  const float* g = static_cast<const float*>(j.grad);
  float acc = 0.f;
  const long long n4 = j.count >> 2;
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (vec) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
      const float4 t = g4[i];
      acc += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
    for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.count; i += (long long)gridDim.x * blockDim.x)
      acc += g[i] * g[i];
  } else {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.count; i += (long long)gridDim.x * blockDim.x)
      acc += g[i] * g[i];
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) {
    state[3 + blockIdx.y * gridDim.x + blockIdx.x] = acc;
    if (blockIdx.x == 0 && blockIdx.y == 0) state[1] += 1.f;
  }
}

struct AdamK {
  float max_norm, lr, b1, b2, omb1, omb2, eps, wd;  // omb = 1 - beta rounded from double, as torch passes it
};

__global__ void clip_adam_kernel(const pbt_optim_job_t* __restrict__ jobs, float* state, AdamK k,
                                 float* __restrict__ norm_out) {
  pdl_sync();
  const pbt_optim_job_t j = jobs[blockIdx.y];
  // every block sums the per-block partials in the same fixed order (strided per thread, then a fixed tree)
  __shared__ float s_total;
  {
    float acc = 0.f;
    const int n_part = (int)(gridDim.x * gridDim.y);
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) acc += state[3 + i];
    acc = block_sum(acc);
    if (threadIdx.x == 0) s_total = acc;
    __syncthreads();
  }
  const float total_sq = s_total;
  const float total = sqrtf(total_sq);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
    if (norm_out) *norm_out = total;
    state[0] = total_sq;   // read by clip_adam_skip_kernel
  }
  if (!isfinite(total)) return;  // inf/nan gradients (fp16 overflow): the step is skipped, AMP-style (see clip_adam_skip_kernel)
  const float t = state[1];
  float coef = 1.f;
  if (k.max_norm > 0.f) coef = fminf(1.f, k.max_norm / (total + 1e-6f));
  const float bc1 = 1.f - powf(k.b1, t);
  const float bc2 = 1.f - powf(k.b2, t);
  const float step_size = k.lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  float* p = static_cast<float*>(j.param);
  const float* g = static_cast<const float*>(j.grad);
  float* m = j.exp_avg;
  float* v = j.exp_avg_sq;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < j.count; i += (long long)gridDim.x * blockDim.x) {
    const float pv = p[i];
    const float gv = fmaf(k.wd, pv, g[i] * coef);
    const float mv = fmaf(k.omb1, gv - m[i], m[i]);
    const float vv = fmaf(k.omb2, gv * gv, k.b2 * v[i]);
    m[i] = mv;
    v[i] = vv;
    const float denom = fmaf(sqrtf(vv), inv_sqrt_bc2, k.eps);
    p[i] = pv - step_size * (mv / denom);
  }
}

// L1 reconstruction loss of the G-only step, value and gradient in one pass (reference lightning_model.py:267-268:
// L1Loss(G(x), post) * reconstruction_weight):  loss += weight/count * sum|y - t| ;  gy = weight/count * sign(y - t).
__global__ void l1_loss_kernel(const float* __restrict__ y, const float* __restrict__ t, long long count, float scale,
                               float* __restrict__ loss, float* __restrict__ gy) {
  pdl_sync();
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const float d = y[i] - t[i];
    acc += fabsf(d);
    gy[i] = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
  }
  acc = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(loss, acc * scale);
}

// a skipped step does not count: take back the increment of grad_sqnorm_kernel, count the skip in state[2]
__global__ void clip_adam_skip_kernel(float* state) {
  pdl_sync();
  if (!isfinite(state[0])) {
    state[1] -= 1.f;
    state[2] += 1.f;
  }
}

}  // namespace pbt

using namespace pbt;

extern "C" int pbt_clip_adam_step(const pbt_optim_job_t* jobs_dev, int32_t n_jobs, int64_t max_elems, float* state,
                                  double max_norm, double lr, double beta1, double beta2, double eps, double weight_decay,
                                  float* norm_out, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(jobs_dev && state && n_jobs > 0 && n_jobs <= 65535 && max_elems > 0, "clip_adam: bad arguments");
  PBT_REQUIRE(lr >= 0.0 && beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, "clip_adam: bad hyper-parameters");
  long long bx = (max_elems + 256 * 8 - 1) / (256 * 8);
  if (bx > 32) bx = 32;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)n_jobs);
  pbt::launch(grad_sqnorm_kernel, grid, 256, 0, st, jobs_dev, state);
  PBT_CUDA_CHECK(cudaGetLastError());
  // hyper-parameters arrive as doubles (Python floats) and are rounded once, the way torch's kernels receive them
  AdamK k{(float)max_norm, (float)lr, (float)beta1, (float)beta2, (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps,
          (float)weight_decay};
  pbt::launch(clip_adam_kernel, grid, 256, 0, st, jobs_dev, state, k, norm_out);
  PBT_CUDA_CHECK(cudaGetLastError());
  pbt::launch(clip_adam_skip_kernel, 1, 1, 0, st, state);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

extern "C" int pbt_l1_loss_fwd_bwd(const float* y, const float* target, int64_t count, float weight, float* loss, float* gy,
                                   void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  PBT_REQUIRE(y && target && loss && gy && count > 0, "l1_loss: bad arguments");
  PBT_CUDA_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), st));
  long long blocks = (count + 256 * 8 - 1) / (256 * 8);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  pbt::launch(l1_loss_kernel, (unsigned)blocks, 256, 0, st, y, target, count, (float)((double)weight / (double)count), loss, gy);
  PBT_CUDA_CHECK(cudaGetLastError());
  return PBT_OK;
}

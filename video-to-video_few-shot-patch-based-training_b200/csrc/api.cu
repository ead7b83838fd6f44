// api.cu — C-ABI plumbing of libpbt: status strings, error capture, TMA tensor-map construction,
// and the host-only order-statistics tree used by the patch sampler.
#include <stdlib.h>

#include "internal.h"

namespace pbt {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* msg) {
  strncpy(g_last_error, msg ? msg : "", sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

int cuda_fail(cudaError_t e, const char* where) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
  set_last_error(buf);
  return PBT_ERR_CUDA;
}

encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<encode_tiled_fn>(sym);
  return fn;
}

int make_p8_tmap(CUtensorMap* map, const pbt_act_t& t, int box_w_px, int box_h, int box_planes) {
  encode_tiled_fn enc = get_encode_tiled();
  if (!enc) {
    set_last_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return PBT_ERR_CUDA;
  }
  if (box_w_px * 8 > 256 || box_h > 256 || box_planes > 256) {
    set_last_error("tensor map box too large");
    return PBT_ERR_ARG;
  }
  const cuuint64_t dims[4] = {(cuuint64_t)t.w * 8, (cuuint64_t)t.h, (cuuint64_t)(t.c / 8), (cuuint64_t)t.n};
  const cuuint64_t strides[3] = {(cuuint64_t)t.w * 16, (cuuint64_t)t.h * t.w * 16, (cuuint64_t)t.img_stride * 2};
  const cuuint32_t box[4] = {(cuuint32_t)(box_w_px * 8), (cuuint32_t)box_h, (cuuint32_t)box_planes, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, t.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d (w=%d h=%d c=%d n=%d box %dx%dx%d)", (int)r,
             t.w, t.h, t.c, t.n, box_w_px, box_h, box_planes);
    set_last_error(buf);
    return PBT_ERR_CUDA;
  }
  return PBT_OK;
}

int num_sms() {
  static int sms = 0;
  if (sms) return sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  return sms;
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PBT_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

}  // namespace pbt

extern "C" int pbt_abi_version(void) { return PBT_ABI_VERSION; }

extern "C" const char* pbt_error_string(int status) {
  switch (status) {
    case PBT_OK: return "ok";
    case PBT_ERR_ARG: return "invalid argument (shape, alignment or enum)";
    case PBT_ERR_CUDA: return "CUDA error";
    case PBT_ERR_UNSUPPORTED: return "unsupported configuration";
    case PBT_ERR_SMEM: return "configuration does not fit shared/tensor memory";
    default: return "unknown status";
  }
}

extern "C" const char* pbt_last_cuda_error(void) { return pbt::g_last_error; }

// ---------------------------------------------------------------------------------------------
// Order-statistics (Fenwick) tree over n slots, 1 = still unused.  Replaces the O(n) list.pop of
// reference src/data/dataset.py:254-256 with O(log n); the selection rule is identical: the list
// of unused indices stays sorted, so popping position k takes the k-th smallest unused index.
// ---------------------------------------------------------------------------------------------
extern "C" void pbt_ostree_reset(int32_t* tree, int32_t n) {
  // linear-time Fenwick construction of an all-ones array
  for (int32_t i = 1; i <= n; ++i) tree[i] = 1;
  tree[0] = n;  // slot 0 holds the number of remaining indices
  for (int32_t i = 1; i <= n; ++i) {
    int32_t j = i + (i & -i);
    if (j <= n) tree[j] += tree[i];
  }
}

extern "C" int32_t pbt_ostree_take(int32_t* tree, int32_t n, int32_t k) {
  if (k < 0 || k >= tree[0]) return -1;
  int32_t pos = 0, remaining = k + 1;
  int32_t step = 1;
  while ((step << 1) <= n) step <<= 1;
  for (; step > 0; step >>= 1) {
    int32_t nxt = pos + step;
    if (nxt <= n && tree[nxt] < remaining) {
      pos = nxt;
      remaining -= tree[nxt];
    }
  }
  // pos+1 is the 1-based slot of the k-th remaining element; clear it
  for (int32_t i = pos + 1; i <= n; i += i & -i) tree[i] -= 1;
  tree[0] -= 1;
  return pos;
}

#include "host_common.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace snapb200 {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return SNAPB200_OK;
  return set_error(SNAPB200_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

void count_launch() { ++g_launches; }

int check_launch(const char* what) {
  ++g_launches;
  return check_cuda(cudaPeekAtLastError(), what);
}

int num_sms() {
  static int cached[SNAP_MAX_DEVICES];  // per device ordinal; a racing first call writes the same value twice
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < SNAP_MAX_DEVICES) {
    n = __atomic_load_n(&cached[dev], __ATOMIC_RELAXED);
    if (n > 0) return n;
  }
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  if (dev >= 0 && dev < SNAP_MAX_DEVICES) __atomic_store_n(&cached[dev], n, __ATOMIC_RELAXED);
  return n;
}

int ensure_dyn_smem(const void* func, size_t bytes, DynSmemState* st, const char* what) {
  int dev = 0;
  if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
  const bool tracked = dev >= 0 && dev < SNAP_MAX_DEVICES;
  if (tracked && __atomic_load_n(&st->bytes[dev], __ATOMIC_ACQUIRE) >= (unsigned long long)bytes) return SNAPB200_OK;
  if (int rc = check_cuda(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), what))
    return rc;
  if (tracked) {
    unsigned long long cur = __atomic_load_n(&st->bytes[dev], __ATOMIC_RELAXED);
    while (cur < (unsigned long long)bytes &&
           !__atomic_compare_exchange_n(&st->bytes[dev], &cur, (unsigned long long)bytes, true, __ATOMIC_RELEASE,
                                        __ATOMIC_RELAXED)) {
    }
  }
  return SNAPB200_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, long long rows, long long cols,
                      long long ld, int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SNAP_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes (ld=%lld)", ld);
  SNAP_REQUIRE(box_cols == 64 || box_cols == 32, "box_cols must be 32 or 64");
  SNAP_REQUIRE(box_rows >= 1 && box_rows <= 256, "box_rows out of range");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SNAPB200_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d",
                     (int)r, rows, cols, ld, box_rows, box_cols);
  return SNAPB200_OK;
}

int make_tmap_2d_bf16_plain(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld,
                            int box_rows, int box_cols) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SNAP_REQUIRE((ld * 2) % 16 == 0 && box_cols % 8 == 0, "TMA pitch / box width must be multiples of 16 bytes");
  SNAP_REQUIRE(box_rows >= 1 && box_rows <= 256 && box_cols <= 256, "box out of range");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled(plain) failed (%d) rows=%lld cols=%lld box=%dx%d",
                     (int)r, rows, cols, box_rows, box_cols);
  return SNAPB200_OK;
}

int make_tmap_window4d_bf16(CUtensorMap* out, const void* base, int Wo, long long step_bytes, int Hq,
                            long long row_bytes, int N) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SNAP_REQUIRE(step_bytes % 16 == 0 && row_bytes % 16 == 0, "window step / row pitch must be multiples of 16 bytes");
  cuuint64_t dims[4] = {32, (cuuint64_t)Wo, (cuuint64_t)Hq, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)step_bytes, (cuuint64_t)row_bytes, (cuuint64_t)row_bytes * (cuuint64_t)Hq};
  cuuint32_t box[4] = {32, 128, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled(window4d) failed (%d) Wo=%d step=%lld Hq=%d row=%lld N=%d",
                     (int)r, Wo, step_bytes, Hq, row_bytes, N);
  return SNAPB200_OK;
}

int make_tmap_rows3d_bf16(CUtensorMap* out, const void* base, int C, long long ld, int W, long long rows) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  SNAP_REQUIRE((ld * 2) % 16 == 0 && C % 8 == 0, "TMA pitch must be a multiple of 16 bytes");
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)W};
  cuuint32_t box[3] = {64, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled(rows3d) failed (%d) C=%d ld=%lld W=%d rows=%lld", (int)r, C, ld, W, rows);
  return SNAPB200_OK;
}

int make_tmap_nd_bf16_plain(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr)
    return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver)");
  SNAP_REQUIRE(rank >= 2 && rank <= 5, "tensor-map rank must be 2..5");
  SNAP_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16B aligned");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    SNAP_REQUIRE(box[i] >= 1 && box[i] <= 256, "box extent out of range");
  }
  for (int i = 0; i + 1 < rank; ++i) {
    st[i] = strides_bytes[i];
    SNAP_REQUIRE(st[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
  }
  SNAP_REQUIRE((box[0] * 2) % 16 == 0, "inner box extent must be a multiple of 16 bytes");
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SNAPB200_ERR_CUDA, "cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
  return SNAPB200_OK;
}

}  // namespace snapb200

extern "C" {
const char* snapb200_last_error(void) { return snapb200::g_err; }
int snapb200_version(void) { return 100; }
long long snapb200_launch_count(void) { return snapb200::g_launches; }
void snapb200_launch_count_reset(void) { snapb200::g_launches = 0; }
}

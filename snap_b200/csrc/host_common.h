// Host-side helpers shared by the C-ABI translation units: thread-local error string, launch
// counter, device properties and TMA tensor-map construction (driver entry point resolved at run
// time so the library does not link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/snapb200.h"

namespace snapb200 {

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int check_launch(const char* what);  // cudaPeekAtLastError + launch counter
int num_sms();  // SM count of the CURRENT device (cached per device ordinal)
void count_launch();

// Per-device "largest dynamic shared memory size opted in so far" for ONE kernel function.  The attribute
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): a host that drives several devices from one
// process (the reference's pmap host, snap/trainer.py:452-464) has to opt every device in.  Zero-initialised static
// storage; thread-safe (relaxed atomics: setting the attribute twice is harmless).
constexpr int SNAP_MAX_DEVICES = 64;
struct DynSmemState {
  unsigned long long bytes[SNAP_MAX_DEVICES];
};
// Opt `func` in to `bytes` of dynamic shared memory on the current device unless a value >= bytes was already set there.
int ensure_dyn_smem(const void* func, size_t bytes, DynSmemState* st, const char* what);

// 2D bf16 tensor map: dims {cols (inner), rows}, row pitch ld elements, box {box_cols, box_rows},
// swizzle = box_cols * 2 bytes (64 or 128), zero fill out of bounds.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, long long rows, long long cols,
                      long long ld, int box_rows, int box_cols);

// same, but un-swizzled (INTERLEAVE_NONE / SWIZZLE_NONE) with an arbitrary 16-byte-multiple box width:
// used to build the "8 rows x 16 B core matrix" K-major operand layout of the sliding-window correlation
int make_tmap_2d_bf16_plain(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld,
                            int box_rows, int box_cols);

// 4D bf16 "sliding window" map over a packed NHWC image (implicit root conv): dims {32 (window elements, contiguous),
// Wo (stride step_bytes: windows OVERLAP), Hq (row pitch), N}; box {32, 128, 1, 1}; SWIZZLE_64B.
int make_tmap_window4d_bf16(CUtensorMap* out, const void* base, int Wo, long long step_bytes, int Hq,
                            long long row_bytes, int N);

// 3D bf16 map over a dense NHWC output viewed as [rows = N*H][W][C] (pixel pitch ld elements): box {64 channels, 32 pixels,
// 1 row}, SWIZZLE_128B.  A store of 32 consecutive pixels of one image row is clipped at W by the hardware, so a GEMM whose
// M tiles run over W rounded up to 128 (the implicit root conv) needs no per-row remap in its epilogue.
int make_tmap_rows3d_bf16(CUtensorMap* out, const void* base, int C, long long ld, int W, long long rows);

// rank-N (<= 5) un-swizzled bf16 map: dims / box innermost first, strides_bytes[i] = pitch of dimension i + 1
int make_tmap_nd_bf16_plain(CUtensorMap* out, const void* base, int rank, const unsigned long long* dims,
                            const unsigned long long* strides_bytes, const unsigned* box);

#define SNAP_REQUIRE(cond, ...)                                        \
  do {                                                                 \
    if (!(cond)) return set_error(SNAPB200_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

}  // namespace snapb200

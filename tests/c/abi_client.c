/* A plain-C host of libsnapb200.so: what a cgo / JNI / ctypes binding does (include/snapb200.h only, no torch, no
   Python).  Loads the library, resolves entry points, checks that argument validation fails loudly without a GPU.
   Built and run by tests/test_abi.py::test_plain_c_client. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "snapb200.h"

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) {
    printf("dlopen failed: %s\n", dlerror());
    return 2;
  }
  int (*version)(void) = (int (*)(void))dlsym(h, "snapb200_version");
  const char* (*last_error)(void) = (const char* (*)(void))dlsym(h, "snapb200_last_error");
  int (*gemm)(const SnapGemmParams*, void*) = (int (*)(const SnapGemmParams*, void*))dlsym(h, "snapb200_gemm_bf16");
  int (*score)(const SnapLocScoreParams*, const void*, const float*, const float*, const uint8_t*, const float*, void*,
               size_t, float*, void*) =
      (int (*)(const SnapLocScoreParams*, const void*, const float*, const float*, const uint8_t*, const float*, void*,
               size_t, float*, void*))dlsym(h, "snapb200_loc_pose_scoring");
  int (*count)(const uint8_t*, const uint8_t*, int, int, int, float*, float*, void*) =
      (int (*)(const uint8_t*, const uint8_t*, int, int, int, float*, float*, void*))dlsym(h, "snapb200_xcorr_count");
  if (!version || !last_error || !gemm || !score || !count) return 3;
  SnapGemmParams g;
  memset(&g, 0, sizeof g);
  int rc = gemm(&g, 0);
  printf("version %d; gemm rc %d (%s)\n", version(), rc, last_error());
  if (rc != SNAPB200_ERR_INVALID) return 4;
  SnapLocScoreParams p;
  memset(&p, 0, sizeof p);
  p.B = 1; p.N = 8; p.H = 16; p.W = 12; p.P = 4; p.cell_size = 0.2f;   /* W not a multiple of 8 */
  rc = score(&p, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  printf("pose scoring rc %d (%s)\n", rc, last_error());
  if (rc != SNAPB200_ERR_INVALID) return 5;
  rc = count(0, 0, 1, 36, 128, 0, 0, 0);
  if (rc != SNAPB200_ERR_INVALID) return 6;
  printf("ok\n");
  return 0;
}

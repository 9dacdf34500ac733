// Internal host interface of the GEMM engine (used by gemm.cu and xcorr.cu).
#pragma once
#include "host_common.h"

namespace snapb200 {
struct GemmParams;
int launch_gemm(const void* A, long long a_rows, int a_cols, long long a_ld, const void* B,
                long long b_rows, int b_cols, long long b_ld, int bn, int bk, const GemmParams& p,
                cudaStream_t s, int ctas_per_sm = 1);
int pick_bn(int n, int bk);
}  // namespace snapb200

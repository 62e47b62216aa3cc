// tcgen05 / TMEM / TMA bf16 GEMM entry points (gemm_tcgen05.cu).
#pragma once
#include <cuda_runtime.h>
#include "common.h"

namespace b200 {

// D[M,N] = epi( sum_k A(m,k) * B(n,k) ).  a_mn=false: A stored [M,K]; a_mn=true: A stored [K,M].
// b_mn=false: B stored [N,K]; b_mn=true: B stored [K,N].  All operands bf16 row-major, D bf16 or fp32.
// epilogue: 0 none, 1 +bias[N], 2 +bias then ReLU, 3 +bias then GELU(erf).
void launch_gemm_bf16(const void* a, const void* b, void* d, const void* bias, int M, int N, int K, bool a_mn, bool b_mn,
                      int epilogue, DType out_dtype, bool accumulate, cudaStream_t stream);
void launch_gemm_nt_bf16(const void* a, const void* b, void* d, const void* bias, int M, int N, int K, int epilogue,
                         DType out_dtype, cudaStream_t stream);
bool gemm_shape_supported(int M, int N, int K, bool a_mn, bool b_mn);
// 0 = auto (CTA pairs for large problems), 1 = always 1-CTA kernel, 2 = always the cta_group::2 kernel
void set_gemm_cta_mode(int mode);

}  // namespace b200

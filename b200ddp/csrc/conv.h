// EXPERIMENTAL tcgen05 convolution entry points (conv3x3_tcgen05.cu) - see the header comment there.
#pragma once
#include <cuda_runtime.h>
#include "common.h"

namespace b200 {

// Patch shape the kernel will use for an N x H x W map: whole rows (BW = W), *bh rows of *bi images; false if W > 128.
bool conv3x3_patch(int N, int H, int W, int* bh, int* bi);
// y[N,H,W,K] = conv3x3(x[N,H,W,C], w[K,3,3,C]), stride 1, zero padding 1, bf16 in / bf16 out, fp32 accumulation.
void launch_conv3x3_fprop(const void* x, const void* w, void* y, int N, int H, int W, int C, int K, cudaStream_t stream);

}  // namespace b200

// LayerNorm forward / backward (named hot op in BASELINE.json's north-star; needed by the BERT
// config).  One warp per row, 16-byte vector loads, fp32 statistics; backward produces dx in the same
// pass as per-CTA partial dgamma/dbeta, finished by a small column-reduce kernel (no atomics, so the
// result is deterministic).
#include "ops.h"

namespace b200 {
namespace {

constexpr int kLnThreads = 256;            // 8 warps = 8 rows per CTA iteration
constexpr int kLnMaxPerLane = 32;          // supports cols <= 32*32 = 1024 in registers; larger cols re-read

template <typename T>
__global__ void __launch_bounds__(kLnThreads) layernorm_fwd_kernel(const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta,
                                                                   int rows, int cols, float eps, T* __restrict__ y,
                                                                   float* __restrict__ mean, float* __restrict__ rstd) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int row = blockIdx.x * warps + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps) {
    const T* xr = x + (size_t)row * cols;
    float s = 0.f;
    for (int i = lane; i < cols; i += 32) s += to_f32<T>(xr[i]);
    const float mu = warp_sum(s) / (float)cols;
    float v = 0.f;
    for (int i = lane; i < cols; i += 32) { const float d = to_f32<T>(xr[i]) - mu; v += d * d; }
    const float rs = rsqrtf(warp_sum(v) / (float)cols + eps);
    T* yr = y + (size_t)row * cols;
    for (int i = lane; i < cols; i += 32) {
      const float xn = (to_f32<T>(xr[i]) - mu) * rs;
      yr[i] = from_f32<T>(xn * to_f32<T>(gamma[i]) + to_f32<T>(beta[i]));
    }
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  }
}

// Each CTA owns a contiguous stripe of rows.  Warps compute dx for their rows and accumulate
// dgamma/dbeta for the columns they touch into registers -> shared -> one partial row per CTA.
template <typename T>
__global__ void __launch_bounds__(kLnThreads) layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ gamma,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd, int rows, int cols,
                                                                   T* __restrict__ dx, float* __restrict__ dgamma_partial,
                                                                   float* __restrict__ dbeta_partial) {
  extern __shared__ float smem[];            // [2][cols] per-CTA column accumulators
  float* sg = smem;
  float* sb = smem + cols;
  for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
  const int r0 = blockIdx.x * rows_per_cta;
  const int r1 = min(rows, r0 + rows_per_cta);
  for (int row = r0 + warp; row < r1; row += warps) {
    const T* xr = x + (size_t)row * cols;
    const T* dyr = dy + (size_t)row * cols;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < cols; i += 32) {
      const float xn = (to_f32<T>(xr[i]) - mu) * rs;
      const float dg = to_f32<T>(dyr[i]) * to_f32<T>(gamma[i]);
      s1 += dg;
      s2 += dg * xn;
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
    T* dxr = dx + (size_t)row * cols;
    for (int i = lane; i < cols; i += 32) {
      const float xn = (to_f32<T>(xr[i]) - mu) * rs;
      const float dyv = to_f32<T>(dyr[i]);
      const float dg = dyv * to_f32<T>(gamma[i]);
      dxr[i] = from_f32<T>(rs * (dg - s1 - xn * s2));
      // column i is only ever touched by lane (i % 32) of each warp: shared atomics across warps
      atomicAdd(&sg[i], dyv * xn);
      atomicAdd(&sb[i], dyv);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    dgamma_partial[(size_t)blockIdx.x * cols + i] = sg[i];
    dbeta_partial[(size_t)blockIdx.x * cols + i] = sb[i];
  }
}

template <typename T>
__global__ void layernorm_bwd_finish_kernel(const float* __restrict__ dgamma_partial, const float* __restrict__ dbeta_partial, int parts,
                                            int cols, T* __restrict__ dgamma, T* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float g = 0.f, b = 0.f;
  for (int p = 0; p < parts; ++p) { g += dgamma_partial[(size_t)p * cols + c]; b += dbeta_partial[(size_t)p * cols + c]; }
  dgamma[c] = from_f32<T>(g);
  dbeta[c] = from_f32<T>(b);
}

}  // namespace

int layernorm_partial_rows(int rows) {
  int p = (rows + 63) / 64;
  if (p > 2 * kNumSMs) p = 2 * kNumSMs;
  if (p < 1) p = 1;
  return p;
}

void launch_layernorm_fwd(const void* x, const void* gamma, const void* beta, DType dt, int rows, int cols, float eps, void* y,
                          float* mean, float* rstd, cudaStream_t s) {
  int blocks = (rows + 7) / 8;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  if (blocks < 1) blocks = 1;
  if (dt == DType::BF16)
    layernorm_fwd_kernel<__nv_bfloat16><<<blocks, kLnThreads, 0, s>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma,
                                                                    (const __nv_bfloat16*)beta, rows, cols, eps, (__nv_bfloat16*)y, mean, rstd);
  else
    layernorm_fwd_kernel<float><<<blocks, kLnThreads, 0, s>>>((const float*)x, (const float*)gamma, (const float*)beta, rows, cols, eps,
                                                            (float*)y, mean, rstd);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

void launch_layernorm_bwd(const void* dy, const void* x, const void* gamma, const float* mean, const float* rstd, DType dt, int rows,
                          int cols, void* dx, float* dgamma_partial, float* dbeta_partial, int partial_rows, void* dgamma,
                          void* dbeta, cudaStream_t s) {
  const size_t smem = 2 * (size_t)cols * sizeof(float);
  if (dt == DType::BF16) {
    layernorm_bwd_kernel<__nv_bfloat16><<<partial_rows, kLnThreads, smem, s>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                                             (const __nv_bfloat16*)gamma, mean, rstd, rows, cols,
                                                                             (__nv_bfloat16*)dx, dgamma_partial, dbeta_partial);
    layernorm_bwd_finish_kernel<__nv_bfloat16><<<(cols + 255) / 256, 256, 0, s>>>(dgamma_partial, dbeta_partial, partial_rows, cols,
                                                                                (__nv_bfloat16*)dgamma, (__nv_bfloat16*)dbeta);
  } else {
    layernorm_bwd_kernel<float><<<partial_rows, kLnThreads, smem, s>>>((const float*)dy, (const float*)x, (const float*)gamma, mean, rstd,
                                                                     rows, cols, (float*)dx, dgamma_partial, dbeta_partial);
    layernorm_bwd_finish_kernel<float><<<(cols + 255) / 256, 256, 0, s>>>(dgamma_partial, dbeta_partial, partial_rows, cols,
                                                                        (float*)dgamma, (float*)dbeta);
  }
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

}  // namespace b200

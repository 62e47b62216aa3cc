// Multi-tensor fused optimizer path: the stock stack runs
// `_foreach_norm` -> stack -> `vector_norm` -> clamp -> `_foreach_mul_` -> `_foreach_add_` (>= 6
// launches + a fresh grad allocation per step).  Here: [sum of squares partials come from the
// allreduce epilogue, or from multi_sqnorm on one GPU] -> clip_coef (1 tiny block) -> multi_sgd,
// which applies clip * lr * g (+ weight decay, momentum, nesterov), maintains fp32 master weights
// for bf16 parameters, and can zero the gradient in the same pass.
#include "ops.h"

namespace b200 {

namespace {

constexpr int kOptThreads = 256;

__device__ __forceinline__ int find_tensor(const uint32_t* blk0, int count, uint32_t b) {
  int lo = 0, hi = count - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (blk0[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename GT>
__global__ void __launch_bounds__(kOptThreads) multi_sqnorm_kernel(const __grid_constant__ OptTable tab, float* __restrict__ partials) {
  __shared__ uint32_t blk0[kMaxOptTensors];
  __shared__ float red[33];
  for (int i = threadIdx.x; i < tab.count; i += blockDim.x) blk0[i] = tab.t[i].blk0;
  __syncthreads();
  const int k = find_tensor(blk0, tab.count, blockIdx.x);
  const OptSlot& s = tab.t[k];
  const uint32_t begin = (blockIdx.x - s.blk0) * kOptChunk;
  const uint32_t end = min(begin + (uint32_t)kOptChunk, s.numel);
  const GT* g = reinterpret_cast<const GT*>(s.g);
  float acc = 0.f;
  if (g != nullptr) {
    for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
      const float v = to_f32<GT>(g[i]);
      acc += v * v;
    }
  }
  const float total = block_sum(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) clip_coef_kernel(const float* __restrict__ partials, int n, float max_norm,
                                                         float grad_scale, float* __restrict__ coef_out,
                                                         float* __restrict__ norm_out) {
  __shared__ float red[33];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];   // fixed order -> deterministic
  const float total = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(total) * grad_scale;
    if (norm_out) *norm_out = norm;
    // torch.nn.utils.clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max=1)
    float coef = max_norm / (norm + 1e-6f);
    *coef_out = (max_norm > 0.f) ? fminf(coef, 1.f) : 1.f;
  }
}

template <typename PT, typename GT>
__global__ void __launch_bounds__(kOptThreads) multi_sgd_kernel(const __grid_constant__ OptTable tab, const __grid_constant__ SgdHyper h) {
  __shared__ uint32_t blk0[kMaxOptTensors];
  for (int i = threadIdx.x; i < tab.count; i += blockDim.x) blk0[i] = tab.t[i].blk0;
  __syncthreads();
  const int k = find_tensor(blk0, tab.count, blockIdx.x);
  const OptSlot& s = tab.t[k];
  if (s.g == nullptr) return;   // parameter without a gradient: untouched, like torch
  const uint32_t begin = (blockIdx.x - s.blk0) * kOptChunk;
  const uint32_t end = min(begin + (uint32_t)kOptChunk, s.numel);
  PT* p = reinterpret_cast<PT*>(s.p);
  GT* g = reinterpret_cast<GT*>(const_cast<void*>(s.g));
  float* master = h.master ? h.master + s.flat_off : nullptr;
  float* mom = h.momentum_buf ? h.momentum_buf + s.flat_off : nullptr;
  const float lr = *h.lr;
  const float coef = (h.clip_coef ? *h.clip_coef : 1.f) * h.grad_scale;
  const bool first = h.step_count ? (*h.step_count == 0) : false;
  for (uint32_t i = begin + threadIdx.x; i < end; i += blockDim.x) {
    float w = master ? master[i] : to_f32<PT>(p[i]);
    float d = to_f32<GT>(g[i]) * coef;
    if (h.weight_decay != 0.f) d += h.weight_decay * w;
    if (mom) {
      float m = first ? d : h.momentum * mom[i] + (1.f - h.dampening) * d;
      mom[i] = m;
      d = h.nesterov ? d + h.momentum * m : m;
    }
    w -= lr * d;
    if (master) master[i] = w;
    p[i] = from_f32<PT>(w);
    if (h.zero_grad) g[i] = from_f32<GT>(0.f);
  }
}

__global__ void scale_inplace_kernel(float* x, size_t n, const float* scalar) {
  const float s = *scalar;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}

}  // namespace

void launch_multi_sqnorm(const OptTable& tab, DType g_dtype, float* partials, cudaStream_t s) {
  if (tab.total_blocks <= 0) return;
  if (g_dtype == DType::BF16) multi_sqnorm_kernel<__nv_bfloat16><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, partials);
  else multi_sqnorm_kernel<float><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, partials);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

void launch_clip_coef(const float* partials, int n, float max_norm, float grad_scale, float* coef_out, float* norm_out,
                      cudaStream_t s) {
  clip_coef_kernel<<<1, 1024, 0, s>>>(partials, n, max_norm, grad_scale, coef_out, norm_out);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

void launch_multi_sgd(const OptTable& tab, DType p_dtype, DType g_dtype, const SgdHyper& h, cudaStream_t s) {
  if (tab.total_blocks <= 0) return;
  const bool pb = p_dtype == DType::BF16, gb = g_dtype == DType::BF16;
  if (pb && gb) multi_sgd_kernel<__nv_bfloat16, __nv_bfloat16><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, h);
  else if (pb) multi_sgd_kernel<__nv_bfloat16, float><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, h);
  else if (gb) multi_sgd_kernel<float, __nv_bfloat16><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, h);
  else multi_sgd_kernel<float, float><<<tab.total_blocks, kOptThreads, 0, s>>>(tab, h);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

void launch_scale_inplace(float* x, size_t n, const float* scalar, cudaStream_t s) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  if (blocks < 1) blocks = 1;
  scale_inplace_kernel<<<blocks, 256, 0, s>>>(x, n, scalar);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

}  // namespace b200

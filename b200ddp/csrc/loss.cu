// Fused loss forward+backward kernels: the stock path runs sub / pow / mean-reduce and a
// separate mse_loss_backward; cross-entropy is log_softmax + nll_loss + their two backwards.
// Each loss here is one launch that emits the scalar loss AND the gradient w.r.t. its input, so the
// autograd Function's backward is just "return the saved tensor (x upstream scale)".
#include "ops.h"

namespace b200 {
namespace {

constexpr int kLossThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kLossThreads) mse_fwd_bwd_kernel(const T* __restrict__ out, const T* __restrict__ tgt, size_t n,
                                                                   float gscale, float* __restrict__ loss, T* __restrict__ dout,
                                                                   float* __restrict__ scratch) {
  __shared__ float red[33];
  __shared__ bool last;
  const float inv_n = 1.f / (float)n;
  const float k = 2.f * inv_n * gscale;
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float d = to_f32<T>(out[i]) - to_f32<T>(tgt[i]);
    acc += d * d;
    dout[i] = from_f32<T>(k * d);
  }
  const float total = block_sum(acc, red);
  // deterministic two-stage reduction: the last block to arrive sums the partials in index order
  unsigned int* counter = reinterpret_cast<unsigned int*>(scratch + gridDim.x);
  if (threadIdx.x == 0) {
    scratch[blockIdx.x] = total;
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    last = (ticket == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    float s = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) s += __ldcg(scratch + i);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *loss = s * inv_n; *counter = 0u; }
  }
}

// One block per row; the row is read once into registers/smem-free streaming form twice (second read
// hits L1/L2: a 30522-wide bf16 row is 61 KB).
template <typename T>
__global__ void __launch_bounds__(kLossThreads) xent_fwd_bwd_kernel(const T* __restrict__ logits, const long long* __restrict__ targets,
                                                                    int rows, int cols, long long ignore_index, float gscale,
                                                                    float* __restrict__ row_loss, T* __restrict__ dlogits) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const T* x = logits + (size_t)row * cols;
  T* dx = dlogits + (size_t)row * cols;
  const long long t = targets[row];
  const bool ignored = (t == ignore_index) || t < 0 || t >= cols;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) m = fmaxf(m, to_f32<T>(x[i]));
  m = block_max(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) s += __expf(to_f32<T>(x[i]) - m);
  s = block_sum(s, red);
  const float lse = m + __logf(s);
  const float inv_s = 1.f / s;
  if (threadIdx.x == 0) row_loss[row] = ignored ? 0.f : lse - to_f32<T>(x[t]);
  const float valid = row_loss[rows];   // written by count_valid_kernel (mean is over non-ignored rows)
  const float g = (ignored || valid <= 0.f) ? 0.f : gscale / valid;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    float p = __expf(to_f32<T>(x[i]) - m) * inv_s;
    if (i == t) p -= 1.f;
    dx[i] = from_f32<T>(p * g);
  }
}

__global__ void __launch_bounds__(1024) xent_finish_kernel(const float* __restrict__ row_loss, const long long* __restrict__ targets,
                                                           int rows, int cols, long long ignore_index, float* __restrict__ loss) {
  __shared__ float red[33];
  float s = 0.f, cnt = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const long long t = targets[i];
    if (t != ignore_index && t >= 0 && t < cols) { s += row_loss[i]; cnt += 1.f; }
  }
  s = block_sum(s, red);
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) *loss = cnt > 0.f ? s / cnt : 0.f;
}

__global__ void __launch_bounds__(1024) count_valid_kernel(const long long* __restrict__ targets, int rows, int cols,
                                                           long long ignore_index, float* __restrict__ out) {
  __shared__ float red[33];
  float cnt = 0.f;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const long long t = targets[i];
    if (t != ignore_index && t >= 0 && t < cols) cnt += 1.f;
  }
  cnt = block_sum(cnt, red);
  if (threadIdx.x == 0) *out = cnt;
}

}  // namespace

int mse_blocks(size_t n) {
  size_t b = (n + kLossThreads * 8 - 1) / (kLossThreads * 8);
  if (b < 1) b = 1;
  if (b > 2 * kNumSMs) b = 2 * kNumSMs;
  return (int)b;
}

void launch_mse_fwd_bwd(const void* out, const void* target, DType dt, size_t n, float gscale, float* loss, void* dout,
                        float* scratch, int blocks, cudaStream_t s) {
  if (dt == DType::BF16)
    mse_fwd_bwd_kernel<__nv_bfloat16><<<blocks, kLossThreads, 0, s>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)target, n,
                                                                     gscale, loss, (__nv_bfloat16*)dout, scratch);
  else
    mse_fwd_bwd_kernel<float><<<blocks, kLossThreads, 0, s>>>((const float*)out, (const float*)target, n, gscale, loss,
                                                             (float*)dout, scratch);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

void launch_xent_fwd_bwd(const void* logits, const long long* targets, DType dt, int rows, int cols, long long ignore_index,
                         float gscale, float* row_loss, float* loss, void* dlogits, cudaStream_t s) {
  // mean over non-ignored rows: count them on device first (row_loss has rows+1 floats; the last
  // one carries the count) so no host round trip is needed.
  const float g = gscale;
  count_valid_kernel<<<1, 1024, 0, s>>>(targets, rows, cols, ignore_index, row_loss + rows);
  if (dt == DType::BF16)
    xent_fwd_bwd_kernel<__nv_bfloat16><<<rows, kLossThreads, 0, s>>>((const __nv_bfloat16*)logits, targets, rows, cols, ignore_index, g,
                                                                   row_loss, (__nv_bfloat16*)dlogits);
  else
    xent_fwd_bwd_kernel<float><<<rows, kLossThreads, 0, s>>>((const float*)logits, targets, rows, cols, ignore_index, g, row_loss,
                                                           (float*)dlogits);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
  xent_finish_kernel<<<1, 1024, 0, s>>>(row_loss, targets, rows, cols, ignore_index, loss);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

}  // namespace b200

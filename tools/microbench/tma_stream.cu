// Microbenchmark: how fast can ONE elected thread per CTA stream 128-byte-row TMA boxes into a shared-memory ring, as a function
// of ring depth, box rows, CTA count and footprint (L2-resident or not)?  No MMA: the consumer only recycles the slot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tools/microbench/tma_stream.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_spin(uint64_t* b, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* s, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(s)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// variant: 0 = producer + consumer threads, try_wait; 1 = same with test_wait spin; 2 = ONE thread per pair does both (wait full -> reissue), try_wait
__global__ void __launch_bounds__(256, 1) stream_kernel(const __grid_constant__ CUtensorMap map, int iters, int stages, int box_rows, int total_rows, int col_blocks,
                                                        int pairs, int variant, unsigned long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int slot = box_rows * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pairs * stages * slot);
  if (threadIdx.x == 0) { for (int i = 0; i < pairs * 32; ++i) mbar_init(&bars[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, role = warp & 1;
  const long long t0 = clock64();
  if (pair < pairs && lane == 0) {
    uint64_t* full = bars + pair * 32;
    uint64_t* empty = full + 16;
    uint8_t* ring = smem + pair * stages * slot;
    // division-free walk: each (CTA, pair) owns a band of rows and sweeps the 64-column blocks of it
    const int band = (int)(((long long)(blockIdx.x * pairs + pair) * 4 * box_rows) % (total_rows - 4 * box_rows));
    auto coords = [&](int i, int* col, int* row) {
      *col = (i & (col_blocks - 1)) * 64;                      // col_blocks is a power of two
      *row = band + ((i >> 5) & 3) * box_rows;
    };
    if (variant == 2) {
      if (role == 0) {
        int col, row;
        for (int i = 0; i < stages && i < iters; ++i) { coords(i, &col, &row); mbar_expect_tx(&full[i], slot); tma_load_2d(&map, &full[i], ring + i * slot, col, row); }
        int s = 0; uint32_t ph = 0;
        for (int i = 0; i < iters; ++i) {
          mbar_wait(&full[s], ph);
          if (i + stages < iters) { coords(i + stages, &col, &row); mbar_expect_tx(&full[s], slot); tma_load_2d(&map, &full[s], ring + s * slot, col, row); }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    } else if (role == 0) {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        if (variant == 1) mbar_spin(&empty[s], ph ^ 1); else mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], slot);
        int col, row; coords(i, &col, &row);
        tma_load_2d(&map, &full[s], ring + s * slot, col, row);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    } else {
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < iters; ++i) {
        if (variant == 1) mbar_spin(&full[s], ph); else mbar_wait(&full[s], ph);
        mbar_arrive(&empty[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  CK(cudaFree(0));
  const int cols = 2048;                        // bf16 elements per row (4 KB pitch)
  const size_t big_rows = 131072;               // 512 MB
  void* buf; CK(cudaMalloc(&buf, big_rows * cols * 2)); CK(cudaMemset(buf, 1, big_rows * cols * 2));
  unsigned long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  printf("%-6s %-5s %-7s %-6s %-5s %-7s | %10s %12s %12s\n", "footpr", "ctas", "boxrows", "stages", "pairs", "variant", "cyc/box", "B/clk/SM", "TB/s total");
  for (int fp = 0; fp < 2; ++fp) {
    const size_t rows = fp == 0 ? 2048 : big_rows;       // 8 MB (L2 resident) or 512 MB
    for (int ctas : {148})
    for (int box_rows : {64, 128})
    for (int pairs : {1, 2, 4})
    for (int variant : {0, 1, 2})
    for (int stages : {2, 4, 8}) {
      if (pairs * stages * box_rows * 128 > 200 * 1024) continue;
      CUtensorMap map;
      cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
      cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
      cuuint32_t es[2] = {1, 1};
      CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
      const int iters = 512;
      const int smem = pairs * stages * box_rows * 128 + 4096;
      for (int rep = 0; rep < 2; ++rep)
        stream_kernel<<<ctas, 256, smem>>>(map, iters, stages, box_rows, (int)rows, cols / 64, pairs, variant, cyc);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      stream_kernel<<<ctas, 256, smem>>>(map, iters, stages, box_rows, (int)rows, cols / 64, pairs, variant, cyc);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<unsigned long long> h(ctas);
      CK(cudaMemcpy(h.data(), cyc, ctas * 8, cudaMemcpyDeviceToHost));
      double avg = 0; for (auto v : h) avg += (double)v; avg /= ctas;
      const double per_box = avg / (iters * pairs);
      const double bytes = (double)box_rows * 128;
      printf("%-6s %-5d %-7d %-6d %-5d %-7d | %10.0f %12.1f %12.2f\n", fp == 0 ? "8MB" : "512MB", ctas, box_rows, stages, pairs, variant, per_box, bytes / per_box,
             (double)ctas * pairs * iters * bytes / (ms * 1e-3) / 1e12);
    }
  }
  return 0;
}

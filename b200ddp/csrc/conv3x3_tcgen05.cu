// EXPERIMENTAL (written without a GPU at hand; not wired into any model; test gated behind B200DDP_TEST_OPTIN=1).
//
// 3x3 / stride-1 / pad-1 convolution forward for channels_last bf16 activations as NINE SHIFTED GEMMs accumulated
// in one TMEM tile - no im2col buffer and no im2col-mode descriptors:
//
//   y[n,h,w,:] = sum_{r,s} x[n, h+r-1, w+s-1, :] . W[:, r, s, :]^T
//
//   * A operand: a rank-4 tiled TMA map over the NHWC input, dims {C, W, H, N}, box {64, BW, BH, BI}.  The box lands a
//     patch of BW x BH x BI pixels as that many consecutive 128-byte rows (128B swizzle) - exactly a K-major A tile of
//     the GEMM - and tap (r, s) is the same box at coordinates (c0, s-1, h0+r-1, n0).  Coordinates that fall outside the
//     tensor are zero-filled by the TMA unit: that is the padding.
//   * B operand: the filter of a channels_last nn.Conv2d is stored [K][r][s][C], i.e. the K-major matrix [K, 9C]; k-block
//     (tap, c0) is the plain 2-D box at column tap*C + c0.
//   * K loop = 9 taps x C/64 blocks through the same warp-specialised pipeline as gemm_tcgen05.cu (TMA producer lane,
//     single-thread tcgen05.mma issuer, double-buffered TMEM accumulator, 4 epilogue warps).
//
// Patch shape: BW = W (whole rows), BH = largest divisor of H with W*BH <= 128, BI = images per tile when a whole image
// fits.  ResNet-50 maps: 56x2, 28x4 (112 of the 128 accumulator rows are real), 14x7, 7x7x2 (98 rows).  Rows of the
// accumulator beyond the patch multiply stale shared memory and are never stored.
//
// Known limit (docs/ROADMAP.md section 2): every tap re-fetches its A patch from L2, so at C = 64..128 the kernel is
// L2-bandwidth bound at roughly a third of tensor peak; loading the halo once and addressing the nine taps through the
// descriptor base offset is the next step.
#include "conv.h"

#include <cuda.h>
#include <stdexcept>
#include <string>

#include "drv.h"
#include "tc_primitives.cuh"

namespace b200 {
namespace {
using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 128 + kEpiWarps * 32;
constexpr int kAccumStages = 2;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2;     // 16 KB reserved per stage (only rows < patch are written)

struct ConvParams {
  int N, H, W, C, K;        // batch, height, width, input channels, output channels
  int BH, BI;               // patch: W x BH pixels of BI images
  int tiles_h, tiles_n;     // H / BH, N / BI
  __nv_bfloat16* y;         // [N, H, W, K]
};

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* smem, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

template <int BLOCK_N>
struct ConvSmem {
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = BLOCK_N == 256 ? 4 : 6;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
};

template <int BLOCK_N>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_fprop_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
  using L = ConvSmem<BLOCK_N>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = kAccumStages * BLOCK_N < 32 ? 32 : kAccumStages * BLOCK_N;   // 128 / 256 / 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * L::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + kAccumStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccumStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = p.tiles_n * p.tiles_h;
  const int num_n = (p.K + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int cblocks = p.C / BLOCK_K;
  const int num_k_blocks = 9 * cblocks;
  const int patch_rows = p.W * p.BH * p.BI;                   // real rows of the 128-row tile
  const uint32_t stage_tx = (uint32_t)(patch_rows * BLOCK_K * 2 + L::kBBytes);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_w); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kAccumStages; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / num_n, nt = tile - mt * num_n;      // n-tiles of one patch run on neighbouring CTAs: the patch stays in L2
        const int img0 = (mt / p.tiles_h) * p.BI;
        const int h0 = (mt % p.tiles_h) * p.BH;
        const int n0 = nt * BLOCK_N;
        for (int tap = 0; tap < 9; ++tap) {
          const int r = tap / 3, s = tap - 3 * r;
          for (int cb = 0; cb < cblocks; ++cb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * L::kStageBytes;
            uint8_t* sb = sa + kABytes;
            mbar_expect_tx(&full_bar[stage], stage_tx);
            tma_load_4d(&map_x, &full_bar[stage], sa, cb * BLOCK_K, s - 1, h0 + r - 1, img0);     // out-of-range rows / columns arrive as zeros
            tma_load_2d(&map_w, &full_bar[stage], sb, tap * p.C + cb * BLOCK_K, n0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int accum = 0;
      uint32_t accum_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[accum], accum_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(accum * BLOCK_N);
        for (int kb = 0; kb < num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * UMMA_K * 2, 16, 1024);
            const uint64_t db = make_smem_desc(b_addr + k * UMMA_K * 2, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[accum]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++accum == kAccumStages) { accum = 0; accum_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> NHWC global =====================
    const int ew = warp - 4;
    int accum = 0;
    uint32_t accum_phase = 0;
    const int per_img = p.W * p.BH;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / num_n, nt = tile - mt * num_n;
      const int img0 = (mt / p.tiles_h) * p.BI;
      const int h0 = (mt % p.tiles_h) * p.BH;
      const int n0 = nt * BLOCK_N;
      mbar_wait(&tmem_full_bar[accum], accum_phase);
      tc_fence_after();
      const int i = ew * 32 + lane;                              // accumulator row == pixel index inside the patch
      const bool live = i < patch_rows;
      const int bi = i / per_img, rem = i - bi * per_img;
      const int hh = rem / p.W, ww = rem - hh * p.W;
      __nv_bfloat16* out = p.y + ((((size_t)(img0 + bi) * p.H + (h0 + hh)) * p.W + ww) * p.K + n0);
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(accum * BLOCK_N);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N; c += 32) {
        if (n0 + c >= p.K) break;                                // warp-uniform
        uint32_t rr[32];
        tmem_ld32(taddr + (uint32_t)c, rr);
        tmem_ld_wait();
        if (live) {
          if (n0 + c + 32 <= p.K) {
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
              __nv_bfloat162 h[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(__uint_as_float(rr[q + 2 * j]), __uint_as_float(rr[q + 2 * j + 1]));
              *reinterpret_cast<uint4*>(out + c + q) = *reinterpret_cast<uint4*>(h);      // K % 8 == 0 -> 16-byte aligned
            }
          } else {
            for (int q = 0; q < 32; ++q)
              if (n0 + c + q < p.K) out[c + q] = __float2bfloat16_rn(__uint_as_float(rr[q]));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[accum]);
      if (++accum == kAccumStages) { accum = 0; accum_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

CUtensorMap encode(const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
  auto& drv = Driver::get();
  if (!drv.TensorMapEncodeTiled) throw std::runtime_error("conv3x3: cuTensorMapEncodeTiled unavailable");
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { B200_CUDA_CHECK(cudaFree(nullptr)); ctx_bound = true; }
  CUtensorMap map;
  cuuint32_t elem_strides[4] = {1, 1, 1, 1};
  B200_DRV_CHECK(drv.TensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(ptr), dims, strides, box, elem_strides,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE));
  return map;
}

template <int BLOCK_N>
void launch(const CUtensorMap& mx, const CUtensorMap& mw, const ConvParams& p, cudaStream_t stream) {
  using L = ConvSmem<BLOCK_N>;
  static_assert(L::kTotal <= 227 * 1024, "shared memory budget");
  auto kernel = conv3x3_fprop_kernel<BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int tiles = p.tiles_n * p.tiles_h * ((p.K + BLOCK_N - 1) / BLOCK_N);
  const int grid = tiles < kNumSMs ? tiles : kNumSMs;
  kernel<<<grid, kThreads, L::kTotal, stream>>>(mx, mw, p);
  B200_CUDA_CHECK(cudaGetLastError()); B200_COUNT_LAUNCH(1);
}

}  // namespace

bool conv3x3_patch(int N, int H, int W, int* bh, int* bi) {
  if (W < 1 || W > BLOCK_M) return false;
  int h = 0;
  for (int d = 1; d <= H; ++d)
    if (H % d == 0 && W * d <= BLOCK_M) h = d;
  if (h == 0) return false;
  int i = 1;
  if (h == H)
    for (int d = 1; d <= N; ++d)
      if (N % d == 0 && W * H * d <= BLOCK_M) i = d;
  *bh = h;
  *bi = i;
  return true;
}

void launch_conv3x3_fprop(const void* x, const void* w, void* y, int N, int H, int W, int C, int K, cudaStream_t stream) {
  int bh = 0, bi = 0;
  if (C % BLOCK_K != 0 || K % 8 != 0) throw std::runtime_error("conv3x3: needs C_in % 64 == 0 and C_out % 8 == 0");
  if (!conv3x3_patch(N, H, W, &bh, &bi)) throw std::runtime_error("conv3x3: image rows wider than 128 pixels are not tiled yet");
  ConvParams p;
  p.N = N; p.H = H; p.W = W; p.C = C; p.K = K;
  p.BH = bh; p.BI = bi;
  p.tiles_h = H / bh;
  p.tiles_n = N / bi;
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  const int block_n = K <= 64 ? 64 : (K <= 128 ? 128 : 256);
  // input: NHWC, innermost first {C, W, H, N}
  cuuint64_t xd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t xs[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t xb[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bi};
  const CUtensorMap mx = encode(x, 4, xd, xs, xb);
  // filter: [K][3][3][C] == row-major [K, 9C]
  cuuint64_t wd[2] = {(cuuint64_t)9 * C, (cuuint64_t)K};
  cuuint64_t ws[1] = {(cuuint64_t)9 * C * 2};
  cuuint32_t wb[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)block_n};
  const CUtensorMap mw = encode(w, 2, wd, ws, wb);
  if (block_n == 64) launch<64>(mx, mw, p, stream);
  else if (block_n == 128) launch<128>(mx, mw, p, stream);
  else launch<256>(mx, mw, p, stream);
}

}  // namespace b200

// tcgen05 convolution entry points (conv_tcgen05.cu, conv_wgrad_tcgen05.cu) for channels_last bf16 activations.
// They replace the cuDNN calls behind the model's forward / backward (reference hot path: /root/reference/ddp.py:221,231).
#pragma once
#include <cuda_runtime.h>
#include "common.h"

namespace b200 {

// How an output tile of <= 128 accumulator rows is cut out of the N x H x W output grid.
//   mode 0 "flat" : 1x1 convolutions - the NHWC tensor is the row-major [N*H*W, C] matrix, a tile is 128 consecutive pixels.
//   mode 1 "patch": filter taps re-fetch their (shifted) input patch: tile = BI images x BH rows x W columns, dense.
//   mode 2 "halo" : the (BH+2) x (W+2) input halo is fetched ONCE per 64-channel block and the nine taps are nine
//                   shared-memory descriptors into it (row offsets of the K-major, 128B-swizzled tile); accumulator rows
//                   enumerate BH x (W+2) positions, the two extra columns per row are discarded by the epilogue.
struct ConvTilePlan {
  int mode = 0;
  int BH = 0, BI = 0, tiles_h = 0;
  int Wp = 0;            // halo row pitch (W + 2)
  int num_m_tiles = 0;
  int dense_rows = 0;    // output rows stored per tile
  int acc_rows = 0;      // accumulator rows in use (<= 128)
  int a_rows = 0;        // shared-memory rows one A load delivers
};
// want_mode: -1 auto, else force (returns false if the geometry does not fit the mode)
bool conv_tile_plan(int N, int H, int W, int R, int S, int want_mode, ConvTilePlan* plan);

struct ConvLaunchCfg {
  int mode = -1;             // -1 auto
  int block_n = 0;           // 0 auto; 64 / 128 / 256
  int kc = 0;                // 0 auto; 64-deep k chunks per pipeline stage (flat / patch tilings)
  void* debug_counters = nullptr;   // device int64[16]: cycle breakdown of CTA 0's producer / issuer / epilogue (diagnostics)
  int set_base_offset = 0;   // halo mode experiment knob: 1 = fill the descriptor base-offset field (measured WRONG on B200: the swizzle phase follows the absolute address)
};

// Forward (dgrad = false):  y[N,H,W,Cout] = conv(x[N,H,W,Cin], w[Cout,R,S,Cin]),  stride 1, padding (R-1)/2.
// Data gradient (dgrad = true): dx[N,H,W,Cin] = conv^T(dy[N,H,W,Cout], w[Cout,R,S,Cin]) - same kernel, mirrored taps, filter read MN-major.
// `a` is x (or dy), `d` the output; R = S in {1, 3}.  col_stats (forward only, optional): [2][conv_stat_groups(...)][Cout] fp32 partial
// column sums / sums of squares of the stored output (BatchNorm statistics from the epilogue).
// Data-gradient epilogue fusion (optional): `addend` [N,H,W,Cin] is added to dx (the shortcut gradient of a residual block - no
// separate add kernel); and if `bn_x` is given, dx is the output gradient of the BatchNorm(+ReLU) whose input was bn_x: the
// epilogue then also writes per-CTA partial sums S1 = sum dx*m, S2 = sum dx*m*xhat into col_stats ([2][conv_stat_groups][Cin]),
// which replaces that BatchNorm's backward reduction pass over dx and bn_x.
struct ConvBwdFusion {
  const void* addend = nullptr;
  const void* bn_x = nullptr;
  const void* bn_mask = nullptr;     // [N*H*W, Cin / 8] bit mask (ReLU) or nullptr
  const float* bn_mean = nullptr;
  const float* bn_rstd = nullptr;
};
void launch_conv_tap_gemm(const void* a, const void* w, void* d, int N, int H, int W, int Cin, int Cout, int R, int S, bool dgrad,
                          const ConvLaunchCfg& cfg, float* col_stats, cudaStream_t stream, const ConvBwdFusion* fuse = nullptr);


// rows G of the column-statistics workspace [2][G][Cout] the forward launch with this configuration will write
int conv_stat_groups(int N, int H, int W, int Cin, int Cout, int R, int S, const ConvLaunchCfg& cfg);

// Weight gradient: dw[Cout,R,S,Cin] = sum over pixels of dy[n,h,w,co] * x[n,h+r-pad,w+s-pad,ci]  (stride 1).
// Split over the pixel dimension across all SMs; fp32 partials in `workspace`, reduced by a second (fully parallel,
// fixed-order -> deterministic) kernel that writes bf16.  workspace_floats(...) sizes the scratch.
struct WgradCfg {
  int split = 0;             // 0 auto
  int tile_m = 0, tile_n = 0;   // 0 auto: output tile (Cout x Cin-per-tap) handled by one CTA
};
size_t conv_wgrad_workspace_floats(int N, int H, int W, int Cin, int Cout, int R, int S, const WgradCfg& cfg);
void launch_conv_wgrad(const void* dy, const void* x, void* dw, int N, int H, int W, int Cin, int Cout, int R, int S,
                       const WgradCfg& cfg, float* workspace, cudaStream_t stream);

}  // namespace b200

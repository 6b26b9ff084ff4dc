// Host-visible part of the fused front (csrc/front_tc.cuh): parameter block and launch entry points.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace hgsfa {

constexpr int FR_HEAD = 768;             // bytes of chunk head (bias of the children | means | output bias)
enum { FR_IN_ROWMAJOR = 0, FR_IN_TILED = 1 };

struct FrontDev {
  int n_sub, img_w, in_dim, out_dim;
  int nn0, nn1, nn2;          // MMA N per level
  int nv_out;                 // valid output columns of a level-2 node
  int nch1, nch2;             // chunks per level-1 / level-2 node
  int sub_bytes;              // weight bytes per subtree
  int sub_per_cta;            // subtrees per CTA (even)
  float s0, s1, s2;           // accumulator -> value (2^-t)
  float clo0, chi0, clo1, chi1, clo2, chi2, p0, p1, p2;
  const int2* pair_xy;        // [n_sub / 2]
  const int* l0_off;          // [n_sub][4]  dy | dx << 8
  const int* out_col;         // [n_sub]
  const uint8_t* wimg;        // [n_sub][sub_bytes]
};


// select the instantiation for the child widths (np1, np2) and reserve its dynamic shared memory
int front_set_attributes(int np1, int np2);
// x: `n` windows, row-major u8 (mode FR_IN_ROWMAJOR: leading dimension ld, 16-byte aligned rows, read in place through a
// 3-D tensor map) or window-minor tiles (FR_IN_TILED); out: third-layer activations, window-minor f32 tiles
int front_launch(const FrontDev& fd, int np1, int np2, int img_h, int sm_count, int mode, const uint8_t* x,
                 int64_t ld, int64_t n, float* out, cudaStream_t st);
bool front_tensor_maps_available();

// ---- single-layer FP16-split kernel (csrc/back_tc.cuh) ----
struct Run;
constexpr int BK_NW = 6;                  // weight-ring stages
constexpr int BK_NA = 4;                  // A-ring stages per group
constexpr int BK_WSTAGE = FR_HEAD + 32 * 64 * 4;
constexpr int BK_EXP_WARPS = 16;          // two tile groups x four lane quarters x two halves
constexpr int BK_THREADS = (BK_EXP_WARPS + 3) * 32;
constexpr int BK_COL_ACC = 0, BK_COL_A = 128, BK_GCOLS = 256;
constexpr int BK_TRI_N = 10;              // triangular products over 10 centred inputs (cuicuilco's s10 selectors)
constexpr int BK_SM_HEAD = 1024 + 4096, BK_HEAD_WARP = 128 + 192, BK_SM_W = BK_SM_HEAD + BK_EXP_WARPS * BK_HEAD_WARP * 4;
constexpr int BK_SM_X = BK_SM_W + BK_NW * BK_WSTAGE;
enum { BKB_WFULL = 0, BKB_WFREE = BK_NW, BKB_G = 2 * BK_NW, BKB_AFULL = 0, BKB_AFREE = BK_NA, BKB_DFULL = 2 * BK_NA,
       BKB_XFULL = 2 * BK_NA + 2, BKB_XFREE = 2 * BK_NA + 4, BKB_GSTRIDE = 2 * BK_NA + 6, BKB_COUNT = 2 * BK_NW + 2 * BKB_GSTRIDE };
enum { BK_ID_RAW = 0, BK_ID = 1, BK_POW = 2, BK_TRI = 3 };

struct BkGroup { int32_t kind, row0, cnt, tri; };     // one 8-term group of the A layout

struct BackDev {
  int n_nodes, d_in, in_dim, out_dim, shared, n_runs, npc;
  int nn;                       // MMA N (16 .. 64)
  int n_chunks;                 // 32-term chunks per node
  int nstx;                     // receptive-field stages per group (1 or 2)
  int x_stage_bytes;            // d_in * 128 * 4 rounded to 128
  int tri_row0;                 // first input row of the triangular products (0 when the op has none)
  int mean_floats;              // d_in rounded up to 8, + 8 (the head of a node: bias[nn] | mean[mean_floats])
  int chunk_bytes;              // FR_HEAD + 32 * nn * 4
  float scale, clo, chi, p, tri_scale;
  const Run* runs;              // [n_nodes][n_runs]
  const int* out_col;           // [n_nodes]  (first output column of the node, col_off included)
  const int* n_valid;           // [n_nodes]
  const BkGroup* groups;        // [n_chunks * 4]
  const uint8_t* wimg;          // [n_w][n_chunks][chunk_bytes]
};

// reserve the kernel's dynamic shared memory (once per device)
int back_set_attributes();
// shared-memory bytes of a launch (0 if the op does not fit); sets nstx and x_stage_bytes
size_t back_layout(BackDev& bd);
// xin / xout: window-minor f32 tiles
int back_launch(const BackDev& bd, int sm_count, const float* xin, float* xout, int64_t ntiles, cudaStream_t st);

}  // namespace hgsfa

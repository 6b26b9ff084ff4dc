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

}  // namespace hgsfa

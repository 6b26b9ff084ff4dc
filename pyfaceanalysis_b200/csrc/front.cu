// Host side of the fused front: tensor-map encoding and launch of csrc/front_tc.cuh (its own translation unit: the
// kernel is instantiated per child-width pair and dominates the compile time of the library).
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "front_tc.cuh"
#include "back_tc.cuh"

namespace hgsfa {

// ---- fused front (csrc/front_tc.cuh) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {      // driver entry point through the runtime: libhgsfa.so does not link libcuda
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// (child width of level 1, child width of level 2), padded to 8 -- keep in step with front.SUPPORTED_NP
#ifdef HGSFA_FRONT_DEV
#define HG_FRONT_CASES(X) X(16, 24)
#else
#define HG_FRONT_CASES(X) X(8, 8) X(8, 16) X(16, 16) X(16, 24) X(16, 32)
#endif

int front_set_attributes(int np1, int np2) {
#define HG_ATTR(A, B)                                                                                                         \
  if (np1 == A && np2 == B) {                                                                                                 \
    HG_CUDA(cudaFuncSetAttribute(front_kernel<A, B, FR_IN_ROWMAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM));  \
    HG_CUDA(cudaFuncSetAttribute(front_kernel<A, B, FR_IN_TILED>, cudaFuncAttributeMaxDynamicSharedMemorySize, FR_SMEM));     \
    return 0;                                                                                                                 \
  }
  HG_FRONT_CASES(HG_ATTR)
#undef HG_ATTR
  return fail("fused front: child widths (%d, %d) are not instantiated", np1, np2);
}

// x: `n` windows, either row-major u8 (leading dimension ld, 16-byte aligned) or window-minor tiles; out: tiled f32
int front_launch(const FrontDev& fd_in, int np1, int np2, int img_h, int sm_count, int mode, const uint8_t* x,
                 int64_t ld, int64_t n, float* out, cudaStream_t st) {
  const int64_t ntiles = ceil_div(n, TILE);
  if (ntiles <= 0) return 0;
  FrontDev fd = fd_in;
  CUtensorMap tm;
  std::memset(&tm, 0, sizeof(tm));
  if (mode == FR_IN_ROWMAJOR) {
    EncodeTiledFn enc = encode_tiled_fn();
    HG_CHECK(enc, "fused front: cuTensorMapEncodeTiled is not available from this driver");
    // dimensions ordered (x, window, y): the box lands in shared memory as [8 rows][128 windows][16 bytes]
    // (a SWIZZLE_128B box of [window][row][16 bytes] faults on this driver with a 16-byte inner extent)
    const cuuint64_t gdim[3] = {cuuint64_t(fd.img_w), cuuint64_t(n), cuuint64_t(img_h)};
    const cuuint64_t gstr[2] = {cuuint64_t(ld), cuuint64_t(fd.img_w)};
    const cuuint32_t box[3] = {16, cuuint32_t(TILE), 8};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(x), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HG_CHECK(r == CUDA_SUCCESS, "fused front: cuTensorMapEncodeTiled failed (%d)", int(r));
  }
  // subtrees per CTA: all of them when the tile pairs alone fill the GPU, else split (pairs stay together)
  const int64_t pairs = ceil_div(ntiles, 2);
  int parts = int(ceil_div(int64_t(sm_count) * 2, pairs));
  parts = std::max(1, std::min(parts, fd.n_sub / 2));
  fd.sub_per_cta = int(ceil_div(fd.n_sub / 2, parts)) * 2;
  dim3 grid((unsigned)pairs, (unsigned)ceil_div(fd.n_sub, fd.sub_per_cta));
#define HG_LAUNCH(A, B)                                                                                             \
  if (np1 == A && np2 == B) {                                                                   \
    if (mode == FR_IN_ROWMAJOR) front_kernel<A, B, FR_IN_ROWMAJOR><<<grid, FR_THREADS, FR_SMEM, st>>>(fd, tm, x, out, ntiles); \
    else front_kernel<A, B, FR_IN_TILED><<<grid, FR_THREADS, FR_SMEM, st>>>(fd, tm, x, out, ntiles);                 \
  }
  HG_FRONT_CASES(HG_LAUNCH)
#undef HG_LAUNCH
  HG_CUDA(cudaGetLastError());
  return 0;
}


bool front_tensor_maps_available() { return encode_tiled_fn() != nullptr; }

// ---- single-layer FP16-split kernel ----
size_t back_layout(BackDev& bd) {
  bd.x_stage_bytes = int((size_t(bd.d_in) * TILE * 4 + 127) & ~size_t(127));
  const size_t limit = size_t(227) * 1024;
  for (int nstx = 2; nstx >= 1; --nstx) {
    const size_t total = size_t(BK_SM_X) + size_t(2) * nstx * bd.x_stage_bytes;
    if (total <= limit) {
      bd.nstx = nstx;
      return total;
    }
  }
  return 0;
}

int back_set_attributes() {
  HG_CUDA(cudaFuncSetAttribute(back_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  return 0;
}

int back_launch(const BackDev& bd_in, int sm_count, const float* xin, float* xout, int64_t ntiles, cudaStream_t st) {
  if (ntiles <= 0) return 0;
  BackDev bd = bd_in;
  const size_t smem = back_layout(bd);
  HG_CHECK(smem > 0, "single-layer FP16 kernel: a receptive field of %d inputs does not fit in shared memory", bd.d_in);
  // nodes per CTA: as many as possible (weights of node i+1 stream behind node i) while the grid still covers the SMs
  const int64_t pairs = ceil_div(ntiles, 2);
  int64_t npc = (pairs * bd.n_nodes) / (int64_t(sm_count) * 3);
  npc = std::max<int64_t>(1, std::min<int64_t>(npc, std::min(bd.n_nodes, 32)));
  bd.npc = int(npc);
  dim3 grid((unsigned)pairs, (unsigned)ceil_div(bd.n_nodes, bd.npc));
  back_kernel<<<grid, BK_THREADS, smem, st>>>(bd, xin, xout, ntiles, BK_SM_X);
  HG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hgsfa

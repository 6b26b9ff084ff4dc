// Cascade controller on the device (sm_100a): coordinate update, discard mask, order-preserving
// compaction.  Replaces, for a whole batch of windows at once,
//   update_current_subimage_coordinates   (reference face_analysis.py:803-840)
//   identify_patches_to_discard           (reference face_analysis.py:842-887)
//   the boolean-index compaction block    (reference FaceDetectUpdated.py:739-759)
// (section 8f-1).
//
// All arithmetic is float64 in the reference's operation order; every multiply / add / divide is an
// explicit round-to-nearest intrinsic so that the compiler cannot contract them into FMAs -- the
// results are bit-identical to numpy's for identical inputs, and comparisons keep their strictness.
// Because windows of several scales (and images) travel together, the per-scale scalars of the
// reference (patch_width, patch_height -> max_Dx_diff, max_Dy_diff, base_side) are per-window here.
#include "common.cuh"

namespace hgsfa {

struct CascadeParams {
  double net_Dx, net_Dy, net_Dang;
  double regression_width, regression_height;
  double min_scale_radio, max_scale_radio;
  double tolerance_posxy, tolerance_scale, tolerance_angle;
  double desired_sampling;
  double cut_off_face;
};

enum { ST_DISC = 0, ST_POSX = 1, ST_POSY = 2, ST_PANG = 3, ST_SCALE = 4 };

__global__ void cascade_update_kernel(int type, double* __restrict__ coords, double* __restrict__ angles,
                                      const double* __restrict__ reg_out, const double* __restrict__ orig_coords,
                                      const double* __restrict__ orig_angles, const int* __restrict__ orig_index,
                                      const double* __restrict__ patch_wh,  // [n_orig][2] patch_width, patch_height
                                      int64_t n, CascadeParams p, uint8_t* __restrict__ keep,
                                      double* __restrict__ conf) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x0 = coords[4 * i], y0 = coords[4 * i + 1], x1 = coords[4 * i + 2], y1 = coords[4 * i + 3];
  double ang = angles[i];
  const double reg = reg_out[i];
  const int oi = orig_index[i];
  const double pw = patch_wh[2 * oi], ph = patch_wh[2 * oi + 1];
  bool wrong = false;
  switch (type) {
    case ST_DISC:
      // new_wrong_images = curr_disc >= cut_off_face   (NaN >= c is false: the window is kept)
      wrong = reg >= p.cut_off_face;
      if (conf) conf[i] = reg;
      break;
    case ST_POSX: {
      const double width = __dsub_rn(x1, x0);
      const double r = __ddiv_rn(__dmul_rn(reg, width), p.regression_width);
      x0 = __dsub_rn(x0, r);
      x1 = __dsub_rn(x1, r);
      const double ox0 = orig_coords[4 * oi], ox1 = orig_coords[4 * oi + 2];
      const double delta = __dsub_rn(__ddiv_rn(__dadd_rn(x1, x0), 2.0), __ddiv_rn(__dadd_rn(ox1, ox0), 2.0));
      const double max_diff = __ddiv_rn(__dmul_rn(p.net_Dx, pw), p.regression_width);
      wrong = fabs(delta) > __dmul_rn(max_diff, p.tolerance_posxy);
      break;
    }
    case ST_POSY: {
      const double height = __dsub_rn(y1, y0);
      const double r = __ddiv_rn(__dmul_rn(reg, height), p.regression_height);
      y0 = __dsub_rn(y0, r);
      y1 = __dsub_rn(y1, r);
      const double oy0 = orig_coords[4 * oi + 1], oy1 = orig_coords[4 * oi + 3];
      const double delta = __dsub_rn(__ddiv_rn(__dadd_rn(y1, y0), 2.0), __ddiv_rn(__dadd_rn(oy1, oy0), 2.0));
      const double max_diff = __ddiv_rn(__dmul_rn(p.net_Dy, ph), p.regression_height);
      wrong = fabs(delta) > __dmul_rn(max_diff, p.tolerance_posxy);
      break;
    }
    case ST_PANG: {
      ang = __dadd_rn(ang, reg);
      const double oa = orig_angles[oi];
      const double lim = __dmul_rn(p.net_Dang, p.tolerance_angle);
      wrong = (ang > __dadd_rn(oa, lim)) || (ang < __dsub_rn(oa, lim));
      break;
    }
    case ST_SCALE: {
      const double old_w = __dsub_rn(x1, x0), old_h = __dsub_rn(y1, y0);
      const double xc = __ddiv_rn(__dadd_rn(x1, x0), 2.0), yc = __ddiv_rn(__dadd_rn(y1, y0), 2.0);
      const double w = __dmul_rn(__ddiv_rn(old_w, reg), p.desired_sampling);
      const double h = __dmul_rn(__ddiv_rn(old_h, reg), p.desired_sampling);
      x0 = __dsub_rn(xc, __ddiv_rn(w, 2.0));
      x1 = __dadd_rn(xc, __ddiv_rn(w, 2.0));
      y0 = __dsub_rn(yc, __ddiv_rn(h, 2.0));
      y1 = __dadd_rn(yc, __ddiv_rn(h, 2.0));
      const double dx = __dsub_rn(x0, x1), dy = __dsub_rn(y0, y1);
      const double side = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      const double base = __dsqrt_rn(__dadd_rn(__dmul_rn(pw, pw), __dmul_rn(ph, ph)));
      const double ratio = __ddiv_rn(side, base);
      wrong = (ratio > __dmul_rn(p.max_scale_radio, p.tolerance_scale)) ||
              (ratio < __ddiv_rn(p.min_scale_radio, p.tolerance_scale));
      break;
    }
    default:
      break;
  }
  coords[4 * i] = x0; coords[4 * i + 1] = y0; coords[4 * i + 2] = x1; coords[4 * i + 3] = y1;
  angles[i] = ang;
  keep[i] = wrong ? 0 : 1;
}

// ---- order-preserving compaction: block counts -> scan of counts -> per-block scatter of source indices
constexpr int CB = 1024;

__global__ void __launch_bounds__(CB) compact_count_kernel(const uint8_t* __restrict__ keep, int64_t n,
                                                           int* __restrict__ block_counts) {
  const int64_t i = int64_t(blockIdx.x) * CB + threadIdx.x;
  const int k = (i < n && keep[i]) ? 1 : 0;
  const int c = __syncthreads_count(k);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(int* __restrict__ block_counts, int n_blocks,
                                                            int64_t* __restrict__ total) {
  // single block: exclusive scan of the block counts in place
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += 1024) {
    const int idx = base + threadIdx.x;
    const int v = idx < n_blocks ? block_counts[idx] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if ((threadIdx.x & 31) >= o) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      int w = warp_sums[threadIdx.x];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, w, o);
        if (threadIdx.x >= o) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const int warp_prefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const int incl = x + warp_prefix + carry;
    if (idx < n_blocks) block_counts[idx] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(CB) compact_scatter_kernel(const uint8_t* __restrict__ keep, int64_t n,
                                                             const int* __restrict__ block_offsets,
                                                             int* __restrict__ src_index) {
  __shared__ int warp_counts[CB / 32];
  const int64_t i = int64_t(blockIdx.x) * CB + threadIdx.x;
  const int k = (i < n && keep[i]) ? 1 : 0;
  const unsigned ballot = __ballot_sync(0xffffffffu, k);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_counts[warp] = __popc(ballot);
  __syncthreads();
  int prefix = 0;
  for (int w = 0; w < warp; ++w) prefix += warp_counts[w];
  if (k) src_index[block_offsets[blockIdx.x] + prefix + __popc(ballot & ((1u << lane) - 1u))] = (int)i;
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                   const int* __restrict__ index, int64_t n_out, int64_t row_bytes) {
  // one warp per output row, 4-byte granules (every array the cascade compacts is 4-byte aligned)
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_out) return;
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src + size_t(index[row]) * row_bytes);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + size_t(row) * row_bytes);
  for (int64_t w = lane; w < row_bytes / 4; w += 32) d[w] = s[w];
}

}  // namespace hgsfa

using namespace hgsfa;

extern "C" int hgsfa_cascade_update_device(int type, double* d_coords, double* d_angles, const double* d_reg_out,
                                           const double* d_orig_coords, const double* d_orig_angles,
                                           const int32_t* d_orig_index, const double* d_patch_wh, int64_t n,
                                           const double* params12, uint8_t* d_keep, double* d_conf, void* stream) {
  HG_CHECK(type >= ST_DISC && type <= ST_SCALE, "Network type unknown!!!: %d", type);
  HG_CHECK(n >= 0, "hgsfa_cascade_update: negative count");
  HG_CHECK(params12, "hgsfa_cascade_update: null parameters");
  if (n == 0) return 0;
  HG_CHECK(d_coords && d_angles && d_reg_out && d_orig_coords && d_orig_angles && d_orig_index && d_patch_wh && d_keep,
           "hgsfa_cascade_update: null buffer");
  CascadeParams p;
  p.net_Dx = params12[0]; p.net_Dy = params12[1]; p.net_Dang = params12[2];
  p.regression_width = params12[3]; p.regression_height = params12[4];
  p.min_scale_radio = params12[5]; p.max_scale_radio = params12[6];
  p.tolerance_posxy = params12[7]; p.tolerance_scale = params12[8]; p.tolerance_angle = params12[9];
  p.desired_sampling = params12[10]; p.cut_off_face = params12[11];
  PtrDeviceGuard guard(d_coords);
  HG_CHECK(guard.ok, "hgsfa_cascade_update: cannot select device %d", guard.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cascade_update_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(type, d_coords, d_angles, d_reg_out, d_orig_coords,
                                                                   d_orig_angles, d_orig_index, d_patch_wh, n, p, d_keep,
                                                                   d_conf);
  HG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hgsfa_compact_index_device(const uint8_t* d_keep, int64_t n, int32_t* d_src_index, int64_t* d_count,
                                          int32_t* d_scratch, int64_t scratch_ints, void* stream) {
  HG_CHECK(n >= 0 && n < (int64_t(1) << 31), "hgsfa_compact_index: count %lld out of range", (long long)n);
  HG_CHECK(d_count, "hgsfa_compact_index: null count");
  PtrDeviceGuard guard(d_count);
  HG_CHECK(guard.ok, "hgsfa_compact_index: cannot select device %d", guard.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    HG_CUDA(cudaMemsetAsync(d_count, 0, sizeof(int64_t), st));
    return 0;
  }
  const int n_blocks = (int)ceil_div(n, CB);
  HG_CHECK(d_keep && d_src_index && d_scratch && scratch_ints >= n_blocks,
           "hgsfa_compact_index: scratch of %lld ints needed", (long long)n_blocks);
  compact_count_kernel<<<n_blocks, CB, 0, st>>>(d_keep, n, d_scratch);
  compact_scan_kernel<<<1, 1024, 0, st>>>(d_scratch, n_blocks, d_count);
  compact_scatter_kernel<<<n_blocks, CB, 0, st>>>(d_keep, n, d_scratch, d_src_index);
  HG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hgsfa_gather_rows_device(const void* d_src, void* d_dst, const int32_t* d_index, int64_t n_out,
                                        int64_t row_bytes, void* stream) {
  HG_CHECK(n_out >= 0 && row_bytes > 0 && row_bytes % 4 == 0, "hgsfa_gather_rows: bad shape (%lld rows of %lld bytes)",
           (long long)n_out, (long long)row_bytes);
  if (n_out == 0) return 0;
  HG_CHECK(d_src && d_dst && d_index, "hgsfa_gather_rows: null buffer");
  PtrDeviceGuard guard(d_dst);
  HG_CHECK(guard.ok, "hgsfa_gather_rows: cannot select device %d", guard.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gather_rows_kernel<<<(unsigned)ceil_div(n_out * 32, 256), 256, 0, st>>>(static_cast<const uint8_t*>(d_src),
                                                                         static_cast<uint8_t*>(d_dst), d_index, n_out,
                                                                         row_bytes);
  HG_CUDA(cudaGetLastError());
  return 0;
}

// Window extraction: sliding-window box -> fixed-size patch (sm_100a).
//
// Replaces load_network_subimages (reference face_analysis.py:775-800) ->
// cuicuilco extract_subimages_rotate + images_asarray -> Pillow Image.transform(size, EXTENT, box,
// NEAREST | BILINEAR | BICUBIC), one Python iteration and one C call per window.
//
// Tiled (window-minor) output -- what the detector uses -- is produced by ONE kernel, crop_tiled_kernel (below):
// index tables in shared memory, batched gathers.  Row-major output (the drop-in host API) and patches too large
// for shared-memory tables use the two-kernel path:
//  1. crop_index_kernel: per window and axis, the 64-entry source-index table of Pillow's NEAREST
//     resampler, reproduced bit-exactly: a = (hi - lo) / n in double, xo = lo + a * 0.5, then n
//     *sequential* double additions (ImagingScaleAffine accumulates; the multiply form differs in the
//     last bit on some boxes -- SURVEY.md Appendix B.3).  -1 marks out-of-image.
//  2. crop_gather_kernel: CTA = (tile of 128 windows, group of output rows).  Lanes run along the
//     output columns of one window, so a warp's 32 byte-gathers fall into 1-4 cache lines of one
//     image row (the image is L2-resident); patches are transposed through shared memory and
//     written window-minor ("TILED") so that the flow kernels read them with 512-byte warp accesses,
//     or written row-major for the drop-in host API.
// Rotated windows (angle != 0), BILINEAR and BICUBIC take the generic per-pixel affine path in the same kernel.
#include "common.cuh"

namespace hgsfa {

constexpr int TILE_W = HGSFA_TILE;

// optional image table: window w reads image img_index[w] (pointer + height / width per image), so that the
// windows of a whole batch of images are extracted by one launch
struct ImageTable {
  const uint8_t* const* ptrs;   // [n_images] device pointers
  const int* hw;                // [n_images][2] = H, W
  const int* index;             // [n_windows]
};

__global__ void crop_index_kernel(const double* __restrict__ boxes, int64_t n, int ow, int oh, int W, int H,
                                  ImageTable tab_img, int* __restrict__ xtab, int* __restrict__ ytab) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= 2 * n) return;
  const int64_t w = t >> 1;
  const int axis = int(t & 1);
  if (tab_img.index) {
    const int im = tab_img.index[w];
    H = tab_img.hw[2 * im];
    W = tab_img.hw[2 * im + 1];
  }
  const double lo = boxes[w * 4 + axis], hi = boxes[w * 4 + 2 + axis];
  const int cnt = axis ? oh : ow;
  const int size = axis ? H : W;
  int* tab = (axis ? ytab + w * oh : xtab + w * ow);
  const double a = __ddiv_rn(__dsub_rn(hi, lo), double(cnt));
  double xo = __dadd_rn(lo, __dmul_rn(a, 0.5));
  for (int c = 0; c < cnt; ++c) {
    int idx = -1;
    if (!(xo < 0.0)) {
      // C (int) cast of a non-negative double; anything at or beyond `size` is out of the image
      if (xo < double(size)) idx = int(xo);
    }
    tab[c] = idx;
    xo = __dadd_rn(xo, a);
  }
}

__device__ __forceinline__ double bilinear_at(const uint8_t* __restrict__ img, int W, int H, double xin, double yin,
                                             bool* valid) {
  // Pillow bilinear_filter8: reject outside [0, size); shift by -0.5; floor; lerp with clamped neighbours
  if (!(xin >= 0.0 && xin < double(W) && yin >= 0.0 && yin < double(H))) { *valid = false; return 0.0; }
  *valid = true;
  const double xs = __dsub_rn(xin, 0.5), ys = __dsub_rn(yin, 0.5);
  const double xf = floor(xs), yf = floor(ys);
  const int x = int(xf), y = int(yf);
  const double dx = __dsub_rn(xs, xf), dy = __dsub_rn(ys, yf);
  const int x0 = min(max(x, 0), W - 1), x1 = min(max(x + 1, 0), W - 1);
  const int y0 = min(max(y, 0), H - 1), y1 = min(max(y + 1, 0), H - 1);
  const double p00 = img[size_t(y0) * W + x0], p01 = img[size_t(y0) * W + x1];
  const double p10 = img[size_t(y1) * W + x0], p11 = img[size_t(y1) * W + x1];
  const double v1 = __dadd_rn(p00, __dmul_rn(__dsub_rn(p01, p00), dx));
  const double v2 = __dadd_rn(p10, __dmul_rn(__dsub_rn(p11, p10), dx));
  return __dadd_rn(v1, __dmul_rn(__dsub_rn(v2, v1), dy));
}

// Pillow bicubic_filter8 (Geometry.c), pinned against Pillow 12.2 (tests/test_oracle_crop.py): reject outside
// [0, size); shift by -0.5; floor; 4 x 4 neighbourhood starting one pixel up-left with clamped columns; rows
// outside the image repeat the previous row's value; clip to [0, 255]; mode 'L' truncates.
__device__ __forceinline__ double bicubic_poly(double v1, double v2, double v3, double v4, double d) {
  const double p1 = v2;
  const double p2 = __dadd_rn(-v1, v3);
  const double p3 = __dsub_rn(__dadd_rn(__dmul_rn(2.0, __dsub_rn(v1, v2)), v3), v4);
  const double p4 = __dadd_rn(__dsub_rn(__dadd_rn(-v1, v2), v3), v4);
  return __dadd_rn(p1, __dmul_rn(d, __dadd_rn(p2, __dmul_rn(d, __dadd_rn(p3, __dmul_rn(d, p4))))));
}
__device__ __forceinline__ uint8_t bicubic_at(const uint8_t* __restrict__ img, int W, int H, double xin, double yin) {
  if (!(xin >= 0.0 && xin < double(W) && yin >= 0.0 && yin < double(H))) return 0;
  const double xs = __dsub_rn(xin, 0.5), ys = __dsub_rn(yin, 0.5);
  const double xf = floor(xs), yf = floor(ys);
  const double dx = __dsub_rn(xs, xf), dy = __dsub_rn(ys, yf);
  const int x = int(xf) - 1, y = int(yf) - 1;
  int xc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xc[k] = min(max(x + k, 0), W - 1);
  double v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = (k == 0) ? min(max(y, 0), H - 1) : y + k;
    if (k == 0 || (yy >= 0 && yy < H)) {
      const uint8_t* row = img + size_t(yy) * W;
      v[k] = bicubic_poly(double(row[xc[0]]), double(row[xc[1]]), double(row[xc[2]]), double(row[xc[3]]), dx);
    } else {
      v[k] = v[k - 1];
    }
  }
  const double r = bicubic_poly(v[0], v[1], v[2], v[3], dy);
  if (r <= 0.0) return 0;
  if (r >= 255.0) return 255;
  return uint8_t(r);
}

// generic sample of output pixel (c, r) of window `box` rotated by delta_ang = -angle about its centre
__device__ __forceinline__ uint8_t sample_generic(const uint8_t* __restrict__ img, int W, int H, const double* box,
                                                  double cs, double sn, bool rotated, int ow, int oh, int c, int r,
                                                  int filter) {
  const double x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
  const double ax = __ddiv_rn(__dsub_rn(x1, x0), double(ow)), ay = __ddiv_rn(__dsub_rn(y1, y0), double(oh));
  double X, Y;
  if (rotated) {
    const double cx = __dmul_rn(__dadd_rn(x0, x1), 0.5), cy = __dmul_rn(__dadd_rn(y0, y1), 0.5);
    const double u = __dsub_rn(__dadd_rn(__dmul_rn(ax, double(c) + 0.5), x0), cx);
    const double v = __dsub_rn(__dadd_rn(__dmul_rn(ay, double(r) + 0.5), y0), cy);
    X = __dadd_rn(cx, __dsub_rn(__dmul_rn(u, cs), __dmul_rn(v, sn)));
    Y = __dadd_rn(cy, __dadd_rn(__dmul_rn(u, sn), __dmul_rn(v, cs)));
  } else {
    X = __dadd_rn(__dmul_rn(ax, double(c) + 0.5), x0);
    Y = __dadd_rn(__dmul_rn(ay, double(r) + 0.5), y0);
  }
  if (filter == HGSFA_BICUBIC) return bicubic_at(img, W, H, X, Y);
  if (filter == HGSFA_BILINEAR) {
    bool valid;
    const double v = bilinear_at(img, W, H, X, Y, &valid);
    return valid ? uint8_t(v) : uint8_t(0);  // mode 'L': truncation
  }
  if (!(X >= 0.0 && Y >= 0.0)) return 0;
  if (!(X < double(W) && Y < double(H))) return 0;
  return img[size_t(int(Y)) * W + int(X)];
}

template <typename T>
__device__ __forceinline__ T cvt_px(uint8_t v) { return T(v); }

// grid (n_tiles, ceil(oh / ROWS)); 256 threads.  OUT_TILED: dst[tile][pixel][128]; else dst[window][pixel].
// FUSED: the NEAREST index tables of the tile's 128 windows are computed by this CTA into shared memory
// (thread = (window, axis), the same sequential double accumulation as crop_index_kernel) instead of being
// read from global tables: no 2 x 256-byte table per window through HBM, no strided table stores.
template <typename OUT_T, bool OUT_TILED, bool FUSED>
__global__ void __launch_bounds__(256) crop_gather_kernel(const uint8_t* img, int H, int W,
                                                          const double* __restrict__ boxes,
                                                          const double* __restrict__ angles, int64_t n, int ow, int oh,
                                                          int filter, ImageTable tab_img,
                                                          const int* __restrict__ xtab,
                                                          const int* __restrict__ ytab, OUT_T* __restrict__ dst,
                                                          int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // staging of one output row for the whole tile: [ow][TILE_W + 4] bytes (padding breaks bank conflicts)
  uint8_t* stage = smem_raw;
  const int stage_ld = TILE_W + 4;
  const int64_t tile = blockIdx.x;
  const int r_begin = blockIdx.y * rows_per_cta;
  const int r_end = min(oh, r_begin + rows_per_cta);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npix = ow * oh;
  // shared tables, row stride + 1 so that the 256 table-building threads hit different banks
  const int xld = ow + 1, yld = oh + 1;
  int* xs = reinterpret_cast<int*>(smem_raw + ((size_t(ow) * stage_ld + 15) & ~size_t(15)));
  int* ys = xs + size_t(TILE_W) * xld;
  if (FUSED) {
    const int wl = tid >> 1, axis = tid & 1;
    const int64_t w = tile * TILE_W + wl;
    const int cnt = axis ? oh : ow;
    int* tab = axis ? ys + wl * yld : xs + wl * xld;
    if (w < n) {
      int Hh = H, Ww = W;
      if (tab_img.index) {
        const int im = tab_img.index[w];
        Hh = tab_img.hw[2 * im];
        Ww = tab_img.hw[2 * im + 1];
      }
      const double lo = boxes[w * 4 + axis], hi = boxes[w * 4 + 2 + axis];
      const int size = axis ? Hh : Ww;
      const double a = __ddiv_rn(__dsub_rn(hi, lo), double(cnt));
      double xo = __dadd_rn(lo, __dmul_rn(a, 0.5));
      for (int c = 0; c < cnt; ++c) {
        int idx = -1;
        if (!(xo < 0.0) && xo < double(size)) idx = int(xo);
        tab[c] = idx;
        xo = __dadd_rn(xo, a);
      }
    } else {
      for (int c = 0; c < cnt; ++c) tab[c] = -1;
    }
    __syncthreads();
  }

  for (int r = r_begin; r < r_end; ++r) {
    // each warp extracts row r of windows warp, warp + 8, ...
    for (int wl = warp; wl < TILE_W; wl += 8) {
      const int64_t w = tile * TILE_W + wl;
      const bool live = w < n;
      if (live && tab_img.index) {
        const int im = tab_img.index[w];
        img = tab_img.ptrs[im];
        H = tab_img.hw[2 * im];
        W = tab_img.hw[2 * im + 1];
      }
      double ang = 0.0;
      if (live && angles) ang = angles[w];
      const bool generic = live && (ang != 0.0 || filter != HGSFA_NEAREST);
      double cs = 1.0, sn = 0.0;
      double box[4] = {0, 0, 0, 0};
      if (generic) {
        box[0] = boxes[w * 4 + 0]; box[1] = boxes[w * 4 + 1]; box[2] = boxes[w * 4 + 2]; box[3] = boxes[w * 4 + 3];
        if (ang != 0.0) {
          // the reference extracts with delta_ang = -angle (face_analysis.py:781)
          const double th = __ddiv_rn(__dmul_rn(-ang, 3.141592653589793), 180.0);
          sincos(th, &sn, &cs);
        }
      }
      const int y = (live && !generic) ? (FUSED ? ys[wl * yld + r] : ytab[w * oh + r]) : -1;
      const uint8_t* row = img + size_t(max(y, 0)) * W;
      for (int c = lane; c < ow; c += 32) {
        uint8_t v = 0;
        if (generic) {
          v = sample_generic(img, W, H, box, cs, sn, ang != 0.0, ow, oh, c, r, filter);
        } else if (y >= 0) {
          const int x = FUSED ? xs[wl * xld + c] : xtab[w * ow + c];
          if (x >= 0) v = __ldg(row + x);
        }
        if (OUT_TILED) {
          stage[c * stage_ld + wl] = v;
        } else if (live) {
          dst[size_t(w) * npix + size_t(r) * ow + c] = cvt_px<OUT_T>(v);
        }
      }
    }
    if (OUT_TILED) {
      __syncthreads();
      // 128 consecutive windows of one pixel are contiguous in the tiled layout
      for (int idx = tid; idx < ow * (TILE_W / 4); idx += 256) {
        const int c = idx / (TILE_W / 4), q = idx % (TILE_W / 4);
        const uchar4 v = *reinterpret_cast<const uchar4*>(stage + c * stage_ld + q * 4);
        OUT_T* o = dst + (size_t(tile) * npix + size_t(r) * ow + c) * TILE_W + q * 4;
        if (sizeof(OUT_T) == 1) {
          *reinterpret_cast<uchar4*>(o) = v;
        } else {
          o[0] = cvt_px<OUT_T>(v.x); o[1] = cvt_px<OUT_T>(v.y); o[2] = cvt_px<OUT_T>(v.z); o[3] = cvt_px<OUT_T>(v.w);
        }
      }
      __syncthreads();
    }
  }
}

// Tiled output, tables in shared memory, memory-level parallelism: the fast path of the detector.
// CTA = tile of 128 windows x a group of output rows, 256 threads.
//   1. thread (window, axis) builds the NEAREST index table of its window in shared memory (the same sequential
//      double accumulation as crop_index_kernel); thread w < 128 also resolves the window's image pointer / size /
//      angle once, so the row loop never chases the image table.
//   2. per output row, a warp handles its 16 windows in batches of CROP_U: all CROP_U byte gathers of a batch are
//      issued before the first is consumed (the single-window loop was bound by one L2 round trip per window
//      and row: 11.8 ms for 476 928 windows).  Rotated / BILINEAR windows take the generic per-pixel path.
//   3. the row is transposed through shared memory and written window-minor, 128 bytes per pixel and tile.
constexpr int CROP_U = 8;
template <typename OUT_T>
__global__ void __launch_bounds__(256) crop_tiled_kernel(const uint8_t* img0, int H0, int W0,
                                                         const double* __restrict__ boxes,
                                                         const double* __restrict__ angles, int64_t n, int ow, int oh,
                                                         int filter, ImageTable tab_img, OUT_T* __restrict__ dst,
                                                         int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* stage = smem_raw;
  const int stage_ld = TILE_W + 4;
  const int xld = ow + 1, yld = oh + 1;
  int* xs = reinterpret_cast<int*>(smem_raw + ((size_t(ow) * stage_ld + 15) & ~size_t(15)));
  int* ys = xs + size_t(TILE_W) * xld;
  __shared__ const uint8_t* w_img[TILE_W];
  __shared__ int w_W[TILE_W], w_H[TILE_W];
  __shared__ double w_ang[TILE_W];
  const int64_t tile = blockIdx.x;
  const int r_begin = blockIdx.y * rows_per_cta;
  const int r_end = min(oh, r_begin + rows_per_cta);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int npix = ow * oh;
  {
    const int wl = tid >> 1, axis = tid & 1;
    const int64_t w = tile * TILE_W + wl;
    const int cnt = axis ? oh : ow;
    int* tab = axis ? ys + wl * yld : xs + wl * xld;
    int Hh = H0, Ww = W0;
    const uint8_t* ip = img0;
    if (w < n && tab_img.index) {
      const int im = tab_img.index[w];
      ip = tab_img.ptrs[im];
      Hh = tab_img.hw[2 * im];
      Ww = tab_img.hw[2 * im + 1];
    }
    if (axis == 0) {
      w_img[wl] = ip;
      w_W[wl] = Ww;
      w_H[wl] = Hh;
      w_ang[wl] = (w < n && angles) ? angles[w] : 0.0;
    }
    if (w < n) {
      const double lo = boxes[w * 4 + axis], hi = boxes[w * 4 + 2 + axis];
      const int size = axis ? Hh : Ww;
      const double a = __ddiv_rn(__dsub_rn(hi, lo), double(cnt));
      double xo = __dadd_rn(lo, __dmul_rn(a, 0.5));
      for (int c = 0; c < cnt; ++c) {
        int idx = -1;
        if (!(xo < 0.0) && xo < double(size)) idx = int(xo);
        tab[c] = idx;
        xo = __dadd_rn(xo, a);
      }
    } else {
      for (int c = 0; c < cnt; ++c) tab[c] = -1;
    }
  }
  __syncthreads();

  for (int r = r_begin; r < r_end; ++r) {
    for (int j0 = 0; j0 < TILE_W / 8; j0 += CROP_U) {
      const uint8_t* rowp[CROP_U];
      bool any_generic = false;
#pragma unroll
      for (int u = 0; u < CROP_U; ++u) {
        const int wl = warp + 8 * (j0 + u);
        const int y = ys[wl * yld + r];
        rowp[u] = (y >= 0) ? w_img[wl] + size_t(y) * w_W[wl] : nullptr;
        any_generic |= (w_ang[wl] != 0.0) || (filter != HGSFA_NEAREST);
      }
      if (!any_generic) {
        for (int c = lane; c < ow; c += 32) {
          uint8_t v[CROP_U];
#pragma unroll
          for (int u = 0; u < CROP_U; ++u) {
            const int x = xs[(warp + 8 * (j0 + u)) * xld + c];
            v[u] = (rowp[u] && x >= 0) ? __ldg(rowp[u] + x) : uint8_t(0);
          }
#pragma unroll
          for (int u = 0; u < CROP_U; ++u) stage[c * stage_ld + warp + 8 * (j0 + u)] = v[u];
        }
      } else {
        for (int u = 0; u < CROP_U; ++u) {
          const int wl = warp + 8 * (j0 + u);
          const int64_t w = tile * TILE_W + wl;
          const double ang = w_ang[wl];
          const int yy = ys[wl * yld + r];
          const bool generic = w < n && (ang != 0.0 || filter != HGSFA_NEAREST);
          double cs = 1.0, sn = 0.0;
          double box[4] = {0, 0, 0, 0};
          if (generic) {
            box[0] = boxes[w * 4 + 0]; box[1] = boxes[w * 4 + 1]; box[2] = boxes[w * 4 + 2]; box[3] = boxes[w * 4 + 3];
            if (ang != 0.0) {
              const double th = __ddiv_rn(__dmul_rn(-ang, 3.141592653589793), 180.0);   // delta_ang = -angle
              sincos(th, &sn, &cs);
            }
          }
          for (int c = lane; c < ow; c += 32) {
            uint8_t v = 0;
            if (generic) {
              v = sample_generic(w_img[wl], w_W[wl], w_H[wl], box, cs, sn, ang != 0.0, ow, oh, c, r, filter);
            } else if (yy >= 0) {
              const int x = xs[wl * xld + c];
              if (x >= 0) v = __ldg(w_img[wl] + size_t(yy) * w_W[wl] + x);
            }
            stage[c * stage_ld + wl] = v;
          }
        }
      }
    }
    __syncthreads();
    for (int idx = tid; idx < ow * (TILE_W / 4); idx += 256) {
      const int c = idx / (TILE_W / 4), q = idx % (TILE_W / 4);
      const uchar4 v = *reinterpret_cast<const uchar4*>(stage + c * stage_ld + q * 4);
      OUT_T* o = dst + (size_t(tile) * npix + size_t(r) * ow + c) * TILE_W + q * 4;
      if (sizeof(OUT_T) == 1) {
        *reinterpret_cast<uchar4*>(o) = v;
      } else {
        o[0] = cvt_px<OUT_T>(v.x); o[1] = cvt_px<OUT_T>(v.y); o[2] = cvt_px<OUT_T>(v.z); o[3] = cvt_px<OUT_T>(v.w);
      }
    }
    __syncthreads();
  }
}

// Row-major uint8 patches (what the fused flow front reads in place through its tensor map): the fast path of the detector.
// CTA = CROP_RW windows, 256 threads, warps independent (4 windows each).
//   1. lane (window, axis) builds the NEAREST index table of its window in shared memory (Pillow's sequential double
//      accumulation, as in crop_index_kernel) and resolves the window's image pointer / size / angle.
//   2. a warp takes one window at a time; a lane produces 4 consecutive pixels of a row (4 byte gathers from the L1 / L2
//      resident image, packed into one 4-byte store; a warp's stores cover 128 contiguous bytes = 2 patch rows).
//      Rotated / BILINEAR / BICUBIC windows take the generic per-pixel path.
constexpr int CROP_RW = 32;
__global__ void __launch_bounds__(256) crop_rows_u8_kernel(const uint8_t* img0, int H0, int W0, const double* __restrict__ boxes,
                                                           const double* __restrict__ angles, int64_t n, int ow, int oh,
                                                           int filter, ImageTable tab_img, uint8_t* __restrict__ dst) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  int* xs = reinterpret_cast<int*>(smem_raw);                 // [CROP_RW][ow]
  int* ys = xs + CROP_RW * ow;                                // [CROP_RW][oh]
  __shared__ const uint8_t* w_img[CROP_RW];
  __shared__ int w_W[CROP_RW], w_H[CROP_RW];
  __shared__ double w_ang[CROP_RW];
  const int64_t w0 = int64_t(blockIdx.x) * CROP_RW;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // every warp owns 4 consecutive windows and builds their tables itself (lanes 0-7 = 4 windows x 2 axes), so warps only
  // synchronise with themselves: while one warp walks its 64-step accumulations the others gather
  if (lane < 8) {
    const int wl = warp * 4 + (lane >> 1), axis = lane & 1;
    const int64_t w = w0 + wl;
    const int cnt = axis ? oh : ow;
    int* tab = axis ? ys + wl * oh : xs + wl * ow;
    int Hh = H0, Ww = W0;
    const uint8_t* ip = img0;
    if (w < n && tab_img.index) {
      const int im = tab_img.index[w];
      ip = tab_img.ptrs[im];
      Hh = tab_img.hw[2 * im];
      Ww = tab_img.hw[2 * im + 1];
    }
    if (axis == 0) {
      w_img[wl] = ip;
      w_W[wl] = Ww;
      w_H[wl] = Hh;
      w_ang[wl] = (w < n && angles) ? angles[w] : 0.0;
    }
    if (w < n) {
      const double lo = boxes[w * 4 + axis], hi = boxes[w * 4 + 2 + axis];
      const int size = axis ? Hh : Ww;
      const double a = __ddiv_rn(__dsub_rn(hi, lo), double(cnt));
      double xo = __dadd_rn(lo, __dmul_rn(a, 0.5));
      for (int c = 0; c < cnt; ++c) {
        int idx = -1;
        if (!(xo < 0.0) && xo < double(size)) idx = int(xo);
        tab[c] = idx;
        xo = __dadd_rn(xo, a);
      }
    }
  }
  __syncwarp();
  const int chunks_per_row = ow / 4, n_chunks = chunks_per_row * oh;
  for (int wl = warp * 4; wl < warp * 4 + 4; ++wl) {
    const int64_t w = w0 + wl;
    if (w >= n) break;
    const uint8_t* img = w_img[wl];
    const int W = w_W[wl], H = w_H[wl];
    const double ang = w_ang[wl];
    uint32_t* out = reinterpret_cast<uint32_t*>(dst + size_t(w) * ow * oh);
    if (ang == 0.0 && filter == HGSFA_NEAREST) {
      const int* xt = xs + wl * ow;
      const int* yt = ys + wl * oh;
      if (chunks_per_row <= 32 && (32 % chunks_per_row) == 0) {
        // the usual patch widths (64: 16 chunks per row): a lane keeps ONE column chunk for the whole window -- its four
        // source columns and their validity live in registers -- and walks down the rows, 32 / chunks_per_row rows per
        // step: per step one table read (the row), four byte gathers, one packed 4-byte store
        const int cc = lane % chunks_per_row, rstep = 32 / chunks_per_row;
        const int4 x = *reinterpret_cast<const int4*>(xt + cc * 4);
        const uint32_t mask = (x.x >= 0 ? 0xffu : 0u) | (x.y >= 0 ? 0xff00u : 0u) | (x.z >= 0 ? 0xff0000u : 0u) | (x.w >= 0 ? 0xff000000u : 0u);
        const int x0 = max(x.x, 0), x1 = max(x.y, 0), x2 = max(x.z, 0), x3 = max(x.w, 0);
        uint32_t* o = out + lane;
#pragma unroll 4
        for (int r = lane / chunks_per_row; r < oh; r += rstep, o += 32) {
          const int y = yt[r];
          uint32_t v = 0u;
          if (y >= 0) {
            const uint8_t* row = img + size_t(y) * W;
            v = (uint32_t(__ldg(row + x0)) | (uint32_t(__ldg(row + x1)) << 8) | (uint32_t(__ldg(row + x2)) << 16) |
                 (uint32_t(__ldg(row + x3)) << 24)) & mask;
          }
          *o = v;
        }
      } else {
#pragma unroll 4
        for (int ch = lane; ch < n_chunks; ch += 32) {
          const int r = ch / chunks_per_row, c0 = (ch - r * chunks_per_row) * 4;
          const int y = yt[r];
          uint32_t v = 0u;
          if (y >= 0) {
            const uint8_t* row = img + size_t(y) * W;
            const int4 x = *reinterpret_cast<const int4*>(xt + c0);
            const uint32_t p0 = x.x >= 0 ? uint32_t(__ldg(row + x.x)) : 0u;
            const uint32_t p1 = x.y >= 0 ? uint32_t(__ldg(row + x.y)) : 0u;
            const uint32_t p2 = x.z >= 0 ? uint32_t(__ldg(row + x.z)) : 0u;
            const uint32_t p3 = x.w >= 0 ? uint32_t(__ldg(row + x.w)) : 0u;
            v = p0 | (p1 << 8) | (p2 << 16) | (p3 << 24);
          }
          out[ch] = v;
        }
      }
    } else {
      double cs = 1.0, sn = 0.0;
      const double box[4] = {boxes[w * 4 + 0], boxes[w * 4 + 1], boxes[w * 4 + 2], boxes[w * 4 + 3]};
      if (ang != 0.0) {
        const double th = __ddiv_rn(__dmul_rn(-ang, 3.141592653589793), 180.0);   // delta_ang = -angle (face_analysis.py:781)
        sincos(th, &sn, &cs);
      }
      uint8_t* o = dst + size_t(w) * ow * oh;
      for (int p = lane; p < ow * oh; p += 32) {
        const int r = p / ow, c = p - r * ow;
        o[p] = sample_generic(img, W, H, box, cs, sn, ang != 0.0, ow, oh, c, r, filter);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// Age-stage crop (reference normalize_image, face_normalization_tools.py:274-324, + the 96 x 96 sub-sampling of
// face_analysis.py:1230-1246): integer crop around the rotation centre -> BICUBIC rotation -> BICUBIC EXTENT to
// 256 x 260 -> NEAREST sub-sampling, each step a uint8 image in the reference.  Here one thread evaluates one output
// sample through the whole chain: the pixel of the 256 x 260 image it needs is a bicubic over 4 x 4 pixels of the
// rotated image, each of which is a bicubic over 4 x 4 pixels of the (virtual) crop -- same double arithmetic, same
// clamping, same truncation to uint8 at every level as Pillow (pinned: tests/test_oracle_normalize.py), no
// intermediate image in memory.  params per face (pyfaceanalysis_b200/normalize.py): crop origin x, y, crop width,
// height, rotate flag (0 copy, 1 affine, 2 rotate-180), Pillow's rotation matrix a0..a5, EXTENT affine xs, x0, ys, y0.
// ------------------------------------------------------------------------------------------------------------
constexpr int AGE_PARAMS = 16;

struct CropImg {      // virtual uint8 image: the source image shifted by the integer crop origin, zero outside
  const uint8_t* img;
  int W, H, ox, oy;
  __device__ __forceinline__ double operator()(int x, int y) const {
    const int sx = x + ox, sy = y + oy;
    return (sx >= 0 && sx < W && sy >= 0 && sy < H) ? double(img[size_t(sy) * W + sx]) : 0.0;
  }
};

// Pillow bicubic_filter8 over a virtual image `px` of size (W, H); returns -1 where Pillow's filter rejects the point
template <typename PX>
__device__ __forceinline__ int bicubic_virtual(const PX& px, int W, int H, double xin, double yin) {
  if (!(xin >= 0.0 && xin < double(W) && yin >= 0.0 && yin < double(H))) return -1;
  const double xs = __dsub_rn(xin, 0.5), ys = __dsub_rn(yin, 0.5);
  const double xf = floor(xs), yf = floor(ys);
  const double dx = __dsub_rn(xs, xf), dy = __dsub_rn(ys, yf);
  const int x = int(xf) - 1, y = int(yf) - 1;
  int xc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) xc[k] = min(max(x + k, 0), W - 1);
  double v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int yy = (k == 0) ? min(max(y, 0), H - 1) : y + k;
    if (k == 0 || (yy >= 0 && yy < H)) v[k] = bicubic_poly(px(xc[0], yy), px(xc[1], yy), px(xc[2], yy), px(xc[3], yy), dx);
    else v[k] = v[k - 1];
  }
  const double r = bicubic_poly(v[0], v[1], v[2], v[3], dy);
  if (r <= 0.0) return 0;
  if (r >= 255.0) return 255;
  return int(uint8_t(r));
}

struct RotImg {       // virtual uint8 image: Pillow's rotate(angle, BICUBIC) of the crop (same size, zero fill)
  CropImg crop;
  int W, H, mode;
  double a0, a1, a2, a3, a4, a5;
  __device__ __forceinline__ double operator()(int x, int y) const {
    if (mode == 0) return crop(x, y);
    if (mode == 2) return crop(W - 1 - x, H - 1 - y);
    const double xin = double(x) + 0.5, yin = double(y) + 0.5;
    const double X = __dadd_rn(__dadd_rn(__dmul_rn(a0, xin), __dmul_rn(a1, yin)), a2);
    const double Y = __dadd_rn(__dadd_rn(__dmul_rn(a3, xin), __dmul_rn(a4, yin)), a5);
    const int v = bicubic_virtual(crop, W, H, X, Y);
    return v < 0 ? 0.0 : double(v);
  }
};

__global__ void __launch_bounds__(256) age_crop_kernel(ImageTable tab_img, const double* __restrict__ params, int64_t n,
                                                       const int* __restrict__ xtab, const int* __restrict__ ytab, int ow, int oh,
                                                       int mid_w, int mid_h, float* __restrict__ dst) {
  const int64_t face = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (face >= n || p >= ow * oh) return;
  const double* q = params + face * AGE_PARAMS;
  const int im = tab_img.index[face];
  RotImg rot;
  rot.crop.img = tab_img.ptrs[im];
  rot.crop.H = tab_img.hw[2 * im];
  rot.crop.W = tab_img.hw[2 * im + 1];
  rot.crop.ox = int(q[0]);
  rot.crop.oy = int(q[1]);
  rot.W = int(q[2]);
  rot.H = int(q[3]);
  rot.mode = int(q[4]);
  rot.a0 = q[5]; rot.a1 = q[6]; rot.a2 = q[7]; rot.a3 = q[8]; rot.a4 = q[9]; rot.a5 = q[10];
  const int r = p / ow, c = p - r * ow;
  const int xi = xtab[c], yi = ytab[r];
  float v = 0.f;
  if (xi >= 0 && yi >= 0) {
    // pixel (xi, yi) of the normalised image: EXTENT as the affine (xs, 0, x0, 0, ys, y0) evaluated like affine_transform
    const double xin = double(xi) + 0.5, yin = double(yi) + 0.5;
    const double X = __dadd_rn(__dadd_rn(__dmul_rn(q[11], xin), __dmul_rn(0.0, yin)), q[12]);
    const double Y = __dadd_rn(__dadd_rn(__dmul_rn(0.0, xin), __dmul_rn(q[13], yin)), q[14]);
    const int s = bicubic_virtual(rot, rot.W, rot.H, X, Y);
    v = s < 0 ? 0.f : float(s);
  }
  (void)mid_w; (void)mid_h;
  dst[(size_t(face / TILE_W) * ow * oh + p) * TILE_W + (face % TILE_W)] = v;       // window-minor tiles, like the eye patches
}

// Per-patch contrast normalisation of TILED float patches, in place (cuicuilco's
// "AgeContrastEnhancement_Avg_Std" as defined by oracle/crop.py: v = x / 255;
// y = (v - mean(v)) / (std(v) + 1e-8) * obj_std + obj_avg).  A CTA owns 32 windows of a tile: lane = window, warp = one
// of CONTRAST_PARTS interleaved pixel subsets, so every load of a warp is 128 contiguous bytes of one pixel row of the
// tile.  Statistics in double; the partial sums are combined in a fixed order (deterministic).  Round 1 ran one thread per
// window over all pixels: 2.7 ms for the ~1 100 eye patches of a batch, most of the eye stage; this form takes ~0.1 ms.
constexpr int CONTRAST_PARTS = 32;
__global__ void __launch_bounds__(32 * CONTRAST_PARTS) contrast_avg_std_kernel(float* __restrict__ x, int64_t n, int64_t dim,
                                                                              double obj_avg, double obj_std) {
  __shared__ double part[CONTRAST_PARTS][33];
  __shared__ double stat[32];
  const int64_t tile = blockIdx.x;
  const int lane = threadIdx.x & 31, sub = threadIdx.x >> 5;
  const int w = blockIdx.y * 32 + lane;
  const bool live = tile * TILE_W + w < n;          // dead lanes only take part in the barriers
  float* p = x + size_t(tile) * dim * TILE_W + w;
  auto reduce = [&](double v) {                      // sum over the parts of this lane's window, same order for every window
    part[sub][lane] = v;
    __syncthreads();
    if (sub == 0) {
      double t = 0.0;
      for (int k = 0; k < CONTRAST_PARTS; ++k) t += part[k][lane];
      stat[lane] = t;
    }
    __syncthreads();
    return stat[lane];
  };
  double s1 = 0.0;
  if (live)
    for (int64_t f = sub; f < dim; f += CONTRAST_PARTS) s1 += double(p[f * TILE_W]) / 255.0;
  const double mean = reduce(s1) / double(dim);
  // population variance like numpy.std, second pass over the centred values
  double var = 0.0;
  if (live)
    for (int64_t f = sub; f < dim; f += CONTRAST_PARTS) {
      const double d = double(p[f * TILE_W]) / 255.0 - mean;
      var = fma(d, d, var);
    }
  const double scale = obj_std / (sqrt(reduce(var) / double(dim)) + 1e-8);
  if (live)
    for (int64_t f = sub; f < dim; f += CONTRAST_PARTS) {
      const double v = double(p[f * TILE_W]) / 255.0;
      p[f * TILE_W] = float((v - mean) * scale + obj_avg);
    }
}

struct CropScratch {
  DevBuf xtab, ytab;
};

// per-device scratch for the index tables (grow-only)
static CropScratch& scratch_for(int device) {
  static thread_local CropScratch s[16];
  return s[device & 15];
}

}  // namespace hgsfa

using namespace hgsfa;

namespace {

int crop_launch(const uint8_t* d_img, int H, int W, ImageTable tab, const double* d_boxes, const double* d_angles,
                int64_t n, int ow, int oh, int filter, void* d_out, int out_dtype, int out_layout, void* stream) {
  HG_CHECK(ow > 0 && oh > 0, "hgsfa_crop_extent: bad patch size ow=%d oh=%d", ow, oh);
  HG_CHECK(ow <= 1024 && oh <= 1024, "hgsfa_crop_extent: patch size %dx%d too large", ow, oh);
  HG_CHECK(filter == HGSFA_NEAREST || filter == HGSFA_BILINEAR || filter == HGSFA_BICUBIC,
           "hgsfa_crop_extent: unsupported interpolation %d (NEAREST=0, BILINEAR=2, BICUBIC=3)", filter);
  HG_CHECK(n >= 0, "hgsfa_crop_extent: negative window count");
  HG_CHECK(out_layout == HGSFA_ROWMAJOR || (out_layout == HGSFA_TILED && out_dtype != HGSFA_F64),
           "hgsfa_crop_extent: tiled output must be u8 or f32");
  if (n == 0) return 0;
  HG_CHECK(d_boxes && d_out, "hgsfa_crop_extent: null buffer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  PtrDeviceGuard guard(d_out);                     // the device that owns the output, not the thread's current one
  HG_CHECK(guard.ok, "hgsfa_crop_extent: cannot select device %d", guard.device);
  const int device = guard.device;
  CropScratch& sc = scratch_for(device);
  // tiled output with tables that fit in shared memory: one fused kernel; otherwise tables through global memory
  const size_t stage_bytes = (size_t(ow) * (TILE_W + 4) + 15) & ~size_t(15);
  const size_t fused_smem = stage_bytes + size_t(TILE_W) * (ow + 1 + oh + 1) * sizeof(int);
  const bool fused = out_layout == HGSFA_TILED && fused_smem <= size_t(160) * 1024;
  // row-major uint8 patches with tables that fit in shared memory: the detector's path (crop_rows_u8_kernel)
  const size_t rows_smem = size_t(CROP_RW) * (ow + oh) * sizeof(int);
  if (out_layout == HGSFA_ROWMAJOR && out_dtype == HGSFA_U8 && ow % 16 == 0 && rows_smem <= size_t(96) * 1024 &&
      (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
    static thread_local bool rows_attr[16] = {};
    if (!rows_attr[device & 15]) {
      HG_CUDA(cudaFuncSetAttribute(crop_rows_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      rows_attr[device & 15] = true;
    }
    crop_rows_u8_kernel<<<(unsigned)ceil_div(n, CROP_RW), 256, rows_smem, st>>>(d_img, H, W, d_boxes, d_angles, n, ow, oh, filter, tab,
                                                                              static_cast<uint8_t*>(d_out));
    HG_CUDA(cudaGetLastError());
    return 0;
  }
  int* xtab = nullptr;
  int* ytab = nullptr;
  if (!fused) {
    if (sc.xtab.reserve(size_t(n) * ow * sizeof(int))) return 1;
    if (sc.ytab.reserve(size_t(n) * oh * sizeof(int))) return 1;
    xtab = static_cast<int*>(sc.xtab.p);
    ytab = static_cast<int*>(sc.ytab.p);
    crop_index_kernel<<<(unsigned)ceil_div(2 * n, 128), 128, 0, st>>>(d_boxes, n, ow, oh, W, H, tab, xtab, ytab);
    HG_CUDA(cudaGetLastError());
  }

  // all rows of a tile in one CTA when there are enough tiles to fill the GPU (the fused tables are built once
  // per CTA), groups of 8 rows otherwise
  const int64_t n_tiles = ceil_div(n, TILE_W);
  const int rows_per_cta = (fused && n_tiles >= 592) ? oh : 8;
  dim3 grid((unsigned)n_tiles, (unsigned)ceil_div(oh, rows_per_cta));
  const size_t smem = size_t(ow) * (TILE_W + 4);
  if (fused) {
    static thread_local bool attr_done[16] = {};     // the attribute is per device
    bool& attr_set = attr_done[device & 15];
    if (!attr_set) {
      HG_CUDA(cudaFuncSetAttribute(crop_tiled_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      HG_CUDA(cudaFuncSetAttribute(crop_tiled_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set = true;
    }
  }
#define HG_LAUNCH_CROP(T, TILED)                                                                                   \
  crop_gather_kernel<T, TILED, false><<<grid, 256, (TILED) ? smem : 0, st>>>(d_img, H, W, d_boxes, d_angles, n, ow, oh, filter, \
                                                                      tab, xtab, ytab, static_cast<T*>(d_out), rows_per_cta)
#define HG_LAUNCH_CROP_FUSED(T)                                                                                    \
  crop_tiled_kernel<T><<<grid, 256, fused_smem, st>>>(d_img, H, W, d_boxes, d_angles, n, ow, oh, filter, tab,         \
                                                      static_cast<T*>(d_out), rows_per_cta)
  if (fused) {
    if (out_dtype == HGSFA_U8) HG_LAUNCH_CROP_FUSED(uint8_t);
    else HG_LAUNCH_CROP_FUSED(float);
  } else if (out_layout == HGSFA_TILED) {
    if (out_dtype == HGSFA_U8) HG_LAUNCH_CROP(uint8_t, true);
    else HG_LAUNCH_CROP(float, true);
  } else {
    if (out_dtype == HGSFA_U8) HG_LAUNCH_CROP(uint8_t, false);
    else if (out_dtype == HGSFA_F32) HG_LAUNCH_CROP(float, false);
    else HG_LAUNCH_CROP(double, false);
  }
#undef HG_LAUNCH_CROP_FUSED
#undef HG_LAUNCH_CROP
  HG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int hgsfa_crop_extent_device(const uint8_t* d_img, int H, int W, const double* d_boxes,
                                        const double* d_angles, int64_t n, int ow, int oh, int filter, void* d_out,
                                        int out_dtype, int out_layout, void* stream) {
  HG_CHECK(H > 0 && W > 0, "hgsfa_crop_extent: bad image size H=%d W=%d", H, W);
  HG_CHECK(d_img || n == 0, "hgsfa_crop_extent: null image");
  return crop_launch(d_img, H, W, ImageTable{nullptr, nullptr, nullptr}, d_boxes, d_angles, n, ow, oh, filter, d_out,
                     out_dtype, out_layout, stream);
}

extern "C" int hgsfa_crop_extent_batch_device(const uint8_t* const* d_img_ptrs, const int32_t* d_img_hw,
                                              const int32_t* d_img_index, const double* d_boxes, const double* d_angles,
                                              int64_t n, int ow, int oh, int filter, void* d_out, int out_dtype,
                                              int out_layout, void* stream) {
  HG_CHECK((d_img_ptrs && d_img_hw && d_img_index) || n == 0, "hgsfa_crop_extent_batch: null image table");
  return crop_launch(nullptr, 1, 1, ImageTable{d_img_ptrs, d_img_hw, d_img_index}, d_boxes, d_angles, n, ow, oh, filter,
                     d_out, out_dtype, out_layout, stream);
}

extern "C" int hgsfa_contrast_avg_std_device(float* d_patches_tiled, int64_t n, int64_t dim, double obj_avg, double obj_std,
                                             void* stream) {
  HG_CHECK(n >= 0 && dim > 0, "hgsfa_contrast_avg_std: bad shape n=%lld dim=%lld", (long long)n, (long long)dim);
  if (n == 0) return 0;
  HG_CHECK(d_patches_tiled, "hgsfa_contrast_avg_std: null buffer");
  PtrDeviceGuard guard(d_patches_tiled);
  HG_CHECK(guard.ok, "hgsfa_contrast_avg_std: cannot select device %d", guard.device);
  contrast_avg_std_kernel<<<dim3((unsigned)ceil_div(n, TILE_W), TILE_W / 32), 32 * CONTRAST_PARTS, 0,
                            static_cast<cudaStream_t>(stream)>>>(d_patches_tiled, n, dim, obj_avg, obj_std);
  HG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hgsfa_age_crop_device(const uint8_t* const* d_img_ptrs, const int32_t* d_img_hw, const int32_t* d_img_index,
                                     const double* d_params, int64_t n, const int32_t* d_xtab, const int32_t* d_ytab, int ow,
                                     int oh, float* d_out_tiled, void* stream) {
  HG_CHECK(n >= 0 && ow > 0 && oh > 0, "hgsfa_age_crop: bad shape n=%lld ow=%d oh=%d", (long long)n, ow, oh);
  if (n == 0) return 0;
  HG_CHECK(d_img_ptrs && d_img_hw && d_img_index && d_params && d_xtab && d_ytab && d_out_tiled, "hgsfa_age_crop: null buffer");
  HG_CHECK(n <= 65535, "hgsfa_age_crop: %lld faces in one call (max 65535)", (long long)n);
  PtrDeviceGuard guard(d_out_tiled);
  HG_CHECK(guard.ok, "hgsfa_age_crop: cannot select device %d", guard.device);
  dim3 grid((unsigned)ceil_div(int64_t(ow) * oh, 256), (unsigned)n);
  age_crop_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(ImageTable{d_img_ptrs, d_img_hw, d_img_index}, d_params, n,
                                                                      d_xtab, d_ytab, ow, oh, 256, 260, d_out_tiled);
  HG_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int hgsfa_crop_extent(const uint8_t* img, int H, int W, const double* boxes, const double* angles,
                                 int64_t n, int ow, int oh, int filter, void* out, int out_dtype, int device,
                                 void* stream) {
  HG_CHECK(n >= 0, "hgsfa_crop_extent: negative window count");
  if (n == 0) return 0;
  HG_CHECK(img && boxes && out, "hgsfa_crop_extent: null buffer");
  HG_CHECK(H > 0 && W > 0 && ow > 0 && oh > 0, "hgsfa_crop_extent: bad sizes H=%d W=%d ow=%d oh=%d", H, W, ow, oh);
  DeviceGuard guard(device);
  HG_CHECK(guard.ok, "hgsfa_crop_extent: cannot select device %d", device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t out_bytes = size_t(n) * ow * oh * dtype_size(out_dtype);
  uint8_t* d_img = nullptr;
  double *d_boxes = nullptr, *d_angles = nullptr;
  void* d_out = nullptr;
  int rc = 1;
  do {
    if (cudaMalloc(&d_img, size_t(H) * W) != cudaSuccess || cudaMalloc(&d_boxes, size_t(n) * 4 * sizeof(double)) != cudaSuccess ||
        (angles && cudaMalloc(&d_angles, size_t(n) * sizeof(double)) != cudaSuccess) ||
        cudaMalloc(&d_out, out_bytes) != cudaSuccess) {
      fail("hgsfa_crop_extent: device allocation failed (%zu output bytes)", out_bytes);
      break;
    }
    if (cudaMemcpyAsync(d_img, img, size_t(H) * W, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_boxes, boxes, size_t(n) * 4 * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess ||
        (angles && cudaMemcpyAsync(d_angles, angles, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess)) {
      fail("hgsfa_crop_extent: host-to-device copy failed");
      break;
    }
    if (hgsfa_crop_extent_device(d_img, H, W, d_boxes, d_angles, n, ow, oh, filter, d_out, out_dtype, HGSFA_ROWMAJOR, st))
      break;
    if (cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      fail("hgsfa_crop_extent: device-to-host copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      break;
    }
    rc = 0;
  } while (0);
  cudaFree(d_img); cudaFree(d_boxes); cudaFree(d_angles); cudaFree(d_out);
  return rc;
}

// Fused HiGSFA layer kernels and the flow plan (sm_100a).
//
// Replaces mdp.Flow.execute over hinet.Switchboard / Layer / CloneLayer nodes whose children are
// SFANode / PCANode / WhiteningNode / GeneralExpansionNode / iGSFANode (reference call sites
// FaceDetectUpdated.py:699, face_analysis.py:1064,1257).
//
// Data layout in HBM ("TILED", hgsfa.h): activations of every layer are kept window-minor,
//   X[tile][feature][128 windows], so that one warp reading one feature of one tile moves 512
// contiguous bytes and the receptive-field gather of a Switchboard is a *feature-index* indirection
// that is uniform across the warp -- the gather costs no extra memory traffic and is fused into the
// operand fetch of the projection.
//
// One layer operation = for every (node, window tile):
//     x0 = X[gather[node]] - in_offset[node]                      (Switchboard + mean subtraction)
//     for each pass p:   A_p = terms_p(x0, rows of earlier passes) (GeneralExpansion term table)
//                        Y_p = A_p @ W_p[node] + b_p[node]         (SFA / PCA / iGSFA projection)
//                        Y_p -> global output columns and/or shared-memory rows
// Expanded features (A_p) only ever exist as 16-row chunks in shared memory; the slow-feature part of
// an iGSFA node stays in shared memory between its two passes.
//
// Thread mapping: a CTA of 4 warps owns (node, TWC window tiles).  Lanes own 4 consecutive windows
// (one float4) so all weight reads are warp-uniform shared-memory broadcasts; a thread accumulates a
// 4 windows x NT outputs register tile with packed FFMA2 (fma.rn.f32x2), the only way to leave issue
// slots free next to the FP32 pipe on sm_100 (tools/microbench.cu: 72 TFLOP/s either way, but FFMA
// alone saturates the issue port).
#include <cstring>
#include <vector>

#include "common.cuh"

namespace hgsfa {

constexpr int KC = 16;         // expansion rows per shared-memory chunk
constexpr int TILE = HGSFA_TILE;
constexpr int MAX_PASSES = 4;
constexpr int THREADS = 128;

enum TermOp { OP_ID = 0, OP_MUL = 1, OP_ABSPOW = 2, OP_SGNPOW = 3, OP_MUL3 = 4, OP_ABS = 5, OP_CLIP = 6 };
enum { DST_GLOBAL = 1, DST_ROWS = 2 };

struct Term { int32_t op, i, j; float p; };

struct PassDev {
  const Term* terms;
  const float* W;      // [n_w][K][Npad]
  const float* b;      // [n_w][Npad]
  const int* n_valid;  // [n_nodes] columns written to global
  const int* col_off;  // [n_nodes] first column (relative to the node's out_col)
  int K, Npad, dst, row0, cfg;
};

struct OpDev {
  int n_nodes, d_in, in_dim, out_dim, n_passes, shared, n_rows, twc;
  float clip_lo, clip_hi;  // saturation applied to values stored to the output buffer
  const int* gather;       // [n_nodes][d_in] feature index in the input buffer
  const float* in_offset;  // [n_w][d_in]
  const int* out_col;      // [n_nodes]
  PassDev pass[MAX_PASSES];
};

// ------------------------------------------------------------------------------------------------
// operand fetch
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_tiled(const float* x, size_t idx4) {
  return __ldg(reinterpret_cast<const float4*>(x) + idx4);
}
__device__ __forceinline__ float4 ld_tiled(const uint8_t* x, size_t idx4) {
  uchar4 v = __ldg(reinterpret_cast<const uchar4*>(x) + idx4);
  return make_float4(float(v.x), float(v.y), float(v.z), float(v.w));
}

__device__ __forceinline__ float abspow(float x, float p) {
  // |x|^p, p > 0; lg2(0) = -inf -> ex2(-inf) = 0
  return exp2f(p * __log2f(fabsf(x)));
}

template <typename IN_T>
struct Fetch {
  const OpDev& op;
  const IN_T* xin;
  const float* sR;  // [n_rows][twc][128]
  const int* gather;
  const float* offs;
  int lane;

  // value of source row i for window tile `tile` (global) / tile slot `slot` (shared rows)
  __device__ __forceinline__ float4 operator()(int i, int64_t tile, int slot) const {
    if (i < op.d_in) {
      const int f = __ldg(gather + i);
      const float o = __ldg(offs + i);
      float4 v = ld_tiled(xin, (size_t(tile) * op.in_dim + f) * (TILE / 4) + lane);
      v.x -= o; v.y -= o; v.z -= o; v.w -= o;
      return v;
    }
    return reinterpret_cast<const float4*>(sR)[(size_t(i - op.d_in) * op.twc + slot) * (TILE / 4) + lane];
  }
};

template <typename IN_T>
__device__ __forceinline__ float4 eval_term(const Term t, const Fetch<IN_T>& src, int64_t tile, int slot) {
  float4 a = src(t.i, tile, slot);
  switch (t.op) {
    case OP_ID:
      return a;
    case OP_MUL: {
      float4 b = src(t.j, tile, slot);
      return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
    }
    case OP_ABSPOW:
      return make_float4(abspow(a.x, t.p), abspow(a.y, t.p), abspow(a.z, t.p), abspow(a.w, t.p));
    case OP_SGNPOW:
      return make_float4(copysignf(abspow(a.x, t.p), a.x), copysignf(abspow(a.y, t.p), a.y),
                         copysignf(abspow(a.z, t.p), a.z), copysignf(abspow(a.w, t.p), a.w));
    case OP_MUL3: {
      float4 b = src(t.j, tile, slot);
      float4 c = src(int(t.p), tile, slot);
      return make_float4(a.x * b.x * c.x, a.y * b.y * c.y, a.z * b.z * c.z, a.w * b.w * c.w);
    }
    case OP_ABS:
      return make_float4(fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w));
    case OP_CLIP:
      return make_float4(fminf(fmaxf(a.x, -t.p), t.p), fminf(fmaxf(a.y, -t.p), t.p),
                         fminf(fmaxf(a.z, -t.p), t.p), fminf(fmaxf(a.w, -t.p), t.p));
    default:
      return make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ------------------------------------------------------------------------------------------------
// packed FP32 FMA:  (d.lo, d.hi) += (a, a) * (w.lo, w.hi)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a2, unsigned long long w2) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a2), "l"(w2));
}
__device__ __forceinline__ unsigned long long dup2(float a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

__device__ __forceinline__ float4 clamp4(float4 v, float lo, float hi) {
  return make_float4(fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi), fminf(fmaxf(v.z, lo), hi),
                     fminf(fmaxf(v.w, lo), hi));
}

// ------------------------------------------------------------------------------------------------
// one pass: A = terms(x0, rows);  Y = A @ W + b
//   WM window tiles x WN column tiles of NT outputs are processed concurrently by the 4 warps.
// ------------------------------------------------------------------------------------------------
template <typename IN_T, int WM, int WN, int NT>
__device__ __forceinline__ void run_pass(const OpDev& op, const PassDev& ps, const IN_T* __restrict__ xin,
                                         float* __restrict__ xout, int64_t tile0, int64_t ntiles, int node,
                                         float* sA, float* sW, float* sR) {
  static_assert(WM * WN == THREADS / 32, "warp grid must cover the CTA");
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int wm = warp / WN, wn = warp % WN;
  const int nw = op.shared ? 0 : node;
  const float* Wg = ps.W + size_t(nw) * ps.K * ps.Npad;
  const float* bg = ps.b + size_t(nw) * ps.Npad;
  Fetch<IN_T> src{op, xin, sR, op.gather + size_t(node) * op.d_in, op.in_offset + size_t(nw) * op.d_in, lane};
  const int n0 = wn * NT;                 // first output column of this warp
  const bool col_active = n0 < ps.Npad;   // Npad is a multiple of NT
  const int w_chunk4 = KC * ps.Npad / 4;

  for (int tg = 0; tg < op.twc; tg += WM) {
    unsigned long long acc[4][NT / 2];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < NT / 2; ++q) acc[r][q] = 0ull;
    const int64_t my_tile = tile0 + tg + wm;
    const bool tile_active = (tg + wm) < op.twc && my_tile < ntiles;

    for (int k0 = 0; k0 < ps.K; k0 += KC) {
      __syncthreads();  // previous chunk fully consumed (and rows of the previous pass visible)
      // ---- build the expansion chunk A[KC][WM][128] ----
      for (int u = warp; u < KC * WM; u += THREADS / 32) {
        const int row = u / WM, slot = u % WM;
        const int64_t tile = tile0 + tg + slot;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((tg + slot) < op.twc && tile < ntiles) {
          const Term t = ps.terms[k0 + row];
          v = eval_term(t, src, tile, tg + slot);
        }
        reinterpret_cast<float4*>(sA)[(row * WM + slot) * (TILE / 4) + lane] = v;
      }
      // ---- stage the weight chunk W[k0:k0+KC][Npad] ----
      {
        const float4* Wg4 = reinterpret_cast<const float4*>(Wg + size_t(k0) * ps.Npad);
        float4* sW4 = reinterpret_cast<float4*>(sW);
        for (int idx = tid; idx < w_chunk4; idx += THREADS) sW4[idx] = __ldg(Wg4 + idx);
      }
      __syncthreads();
      // ---- register-tile GEMM: 4 windows x NT outputs per thread ----
      if (tile_active && col_active) {
        const float4* a4 = reinterpret_cast<const float4*>(sA) + wm * (TILE / 4) + lane;
        const float* wrow = sW + n0;
#pragma unroll 4
        for (int kc = 0; kc < KC; ++kc) {
          const float4 a = a4[kc * WM * (TILE / 4)];
          const unsigned long long ax = dup2(a.x), ay = dup2(a.y), az = dup2(a.z), aw = dup2(a.w);
          const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(wrow + kc * ps.Npad);
#pragma unroll
          for (int q = 0; q < NT / 4; ++q) {
            const ulonglong2 w = w2[q];
            ffma2(acc[0][2 * q], ax, w.x); ffma2(acc[0][2 * q + 1], ax, w.y);
            ffma2(acc[1][2 * q], ay, w.x); ffma2(acc[1][2 * q + 1], ay, w.y);
            ffma2(acc[2][2 * q], az, w.x); ffma2(acc[2][2 * q + 1], az, w.y);
            ffma2(acc[3][2 * q], aw, w.x); ffma2(acc[3][2 * q + 1], aw, w.y);
          }
        }
      }
    }

    // ---- epilogue: bias, then global columns and/or shared rows ----
    if (tile_active && col_active) {
      const int nvalid = (ps.dst & DST_GLOBAL) ? __ldg(ps.n_valid + node) : 0;
      const int col0 = (ps.dst & DST_GLOBAL) ? (__ldg(op.out_col + node) + __ldg(ps.col_off + node)) : 0;
#pragma unroll
      for (int q = 0; q < NT / 2; ++q) {
        const float2 y0 = unpack2(acc[0][q]), y1 = unpack2(acc[1][q]);
        const float2 y2 = unpack2(acc[2][q]), y3 = unpack2(acc[3][q]);
        const int n = n0 + 2 * q;
        const float b0 = __ldg(bg + n), b1 = __ldg(bg + n + 1);
        const float4 v0 = make_float4(y0.x + b0, y1.x + b0, y2.x + b0, y3.x + b0);
        const float4 v1 = make_float4(y0.y + b1, y1.y + b1, y2.y + b1, y3.y + b1);
        if (ps.dst & DST_GLOBAL) {
          float4* o = reinterpret_cast<float4*>(xout) + (size_t(my_tile) * op.out_dim + col0) * (TILE / 4) + lane;
          if (n < nvalid) o[size_t(n) * (TILE / 4)] = clamp4(v0, op.clip_lo, op.clip_hi);
          if (n + 1 < nvalid) o[size_t(n + 1) * (TILE / 4)] = clamp4(v1, op.clip_lo, op.clip_hi);
        }
        if (ps.dst & DST_ROWS) {
          float4* r4 = reinterpret_cast<float4*>(sR);
          r4[(size_t(ps.row0 + n) * op.twc + (tg + wm)) * (TILE / 4) + lane] = v0;
          r4[(size_t(ps.row0 + n + 1) * op.twc + (tg + wm)) * (TILE / 4) + lane] = v1;
        }
      }
    }
  }
}

// NTMAX = 16 omits the 4 x 32 register tile (cfg 3) so that the common instantiation keeps a small
// register footprint; ops whose widest pass has more than 64 output columns use NTMAX = 32.
template <typename IN_T, int NTMAX>
__global__ void __launch_bounds__(THREADS) layer_kernel(const OpDev op, const IN_T* __restrict__ xin,
                                                        float* __restrict__ xout, int64_t ntiles, int npad_max) {
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                              // [KC][twc][128]
  float* sW = sA + KC * op.twc * TILE;           // [KC][npad_max]
  float* sR = sW + KC * npad_max;                // [n_rows][twc][128]
  const int node = blockIdx.y;
  const int64_t tile0 = int64_t(blockIdx.x) * op.twc;
#pragma unroll 1
  for (int p = 0; p < op.n_passes; ++p) {
    const PassDev& ps = op.pass[p];
    switch (ps.cfg) {
      case 0: run_pass<IN_T, 4, 1, 16>(op, ps, xin, xout, tile0, ntiles, node, sA, sW, sR); break;
      case 1: run_pass<IN_T, 2, 2, 16>(op, ps, xin, xout, tile0, ntiles, node, sA, sW, sR); break;
      case 2: run_pass<IN_T, 1, 4, 16>(op, ps, xin, xout, tile0, ntiles, node, sA, sW, sR); break;
      default:
        if constexpr (NTMAX >= 32) run_pass<IN_T, 1, 4, 32>(op, ps, xin, xout, tile0, ntiles, node, sA, sW, sR);
        break;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// layout kernels: row-major <-> tiled
// ------------------------------------------------------------------------------------------------
// grid (ceil(dim/64), n_tiles), 256 threads.  Row-major reads are contiguous along features, tiled
// writes contiguous along windows; the 128 x 64 block is transposed through shared memory.
template <typename SRC, typename DST>
__global__ void __launch_bounds__(256) tile_windows_kernel(const SRC* __restrict__ src, int64_t n, int64_t dim,
                                                           int64_t ld, DST* __restrict__ dst) {
  __shared__ DST s[64][TILE + (sizeof(DST) == 1 ? 4 : 1)];
  const int64_t tile = blockIdx.y;
  const int f0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TILE * 64; idx += 256) {
    const int w = idx >> 6, f = idx & 63;
    const int64_t gw = tile * TILE + w;
    DST v = DST(0);
    if (gw < n && f0 + f < dim) v = DST(src[gw * ld + f0 + f]);
    s[f][w] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < TILE * 64; idx += 256) {
    const int f = idx >> 7, w = idx & 127;
    if (f0 + f < dim) dst[(tile * dim + f0 + f) * TILE + w] = s[f][w];
  }
}

// tiled f32 [tile][dim][128] -> row-major (n x cols) f32 / f64, keeping the first `cols` features
template <typename DST>
__global__ void __launch_bounds__(256) untile_kernel(const float* __restrict__ src, int64_t n, int64_t dim,
                                                     int64_t cols, DST* __restrict__ dst) {
  __shared__ float s[TILE][33];
  const int64_t tile = blockIdx.y;
  const int f0 = blockIdx.x * 32;
  const int tid = threadIdx.x;
  for (int idx = tid; idx < TILE * 32; idx += 256) {
    const int f = idx >> 7, w = idx & 127;
    s[w][f] = (f0 + f < cols) ? src[(tile * dim + f0 + f) * TILE + w] : 0.f;
  }
  __syncthreads();
  for (int idx = tid; idx < TILE * 32; idx += 256) {
    const int w = idx >> 5, f = idx & 31;
    const int64_t gw = tile * TILE + w;
    if (gw < n && f0 + f < cols) dst[gw * cols + f0 + f] = DST(s[w][f]);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: plan
// ------------------------------------------------------------------------------------------------
struct PassHost {
  int K, Npad, dst, row0, cfg, K_real, N_real;
};
struct OpHost {
  OpDev dev;
  PassHost pass[MAX_PASSES];
  int64_t alg_flops, exe_flops;
  int npad_max;
  bool wide;   // some pass uses the 4 x 32 register tile (cfg 3)
  size_t smem_bytes;
};

}  // namespace hgsfa

using namespace hgsfa;

struct hgsfa_plan_s {
  int device = 0;
  cudaStream_t stream = nullptr;     // compute stream owned by the plan
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  int64_t input_dim = 0, output_dim = 0;
  std::vector<OpHost> ops;
  DevBuf params;                     // all plan arrays in one allocation
  DevBuf tin, front[2], mid, back[2], stage_x[2], stage_y[2];
  int64_t front_chunk = 16384, back_chunk = 262144;
  int split = 0;                     // ops [0, split) run per front chunk, [split, n) per back chunk
  int64_t launches = 0;
  double last_ms = 0.0;
};

namespace {

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  template <typename T>
  const T* take(size_t count) {
    size_t bytes = (count * sizeof(T) + 15) & ~size_t(15);
    if (size_t(end - p) < bytes) { ok = false; return nullptr; }
    const T* r = reinterpret_cast<const T*>(p);
    p += bytes;
    return r;
  }
};

const int CFG_WM[4] = {4, 2, 1, 1};
const int CFG_NT[4] = {16, 16, 16, 32};
const int CFG_WN[4] = {1, 2, 4, 4};

template <typename IN_T>
int launch_layer(hgsfa_plan_s* pl, const OpHost& op, const void* xin, float* xout, int64_t ntiles,
                 cudaStream_t st) {
  if (ntiles <= 0) return 0;
  dim3 grid((unsigned)ceil_div(ntiles, op.dev.twc), (unsigned)op.dev.n_nodes);
  if (op.wide)
    layer_kernel<IN_T, 32><<<grid, THREADS, op.smem_bytes, st>>>(op.dev, static_cast<const IN_T*>(xin), xout, ntiles,
                                                                op.npad_max);
  else
    layer_kernel<IN_T, 16><<<grid, THREADS, op.smem_bytes, st>>>(op.dev, static_cast<const IN_T*>(xin), xout, ntiles,
                                                                op.npad_max);
  pl->launches++;
  HG_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int hgsfa_plan_create(const void* blob, size_t nbytes, int device, hgsfa_plan_t* out) {
  HG_CHECK(blob && out, "hgsfa_plan_create: null argument");
  HG_CHECK(nbytes >= 64, "hgsfa_plan_create: blob too small (%zu bytes)", nbytes);
  int ndev = 0;
  HG_CUDA(cudaGetDeviceCount(&ndev));
  HG_CHECK(device >= 0 && device < ndev, "hgsfa_plan_create: device %d out of range (%d devices)", device, ndev);
  DeviceGuard guard(device);
  HG_CHECK(guard.ok, "hgsfa_plan_create: cannot select device %d", device);

  const uint8_t* base = static_cast<const uint8_t*>(blob);
  HG_CHECK(std::memcmp(base, "HGSFAPL1", 8) == 0, "hgsfa_plan_create: bad magic (not a plan blob)");
  const int64_t* hdr = reinterpret_cast<const int64_t*>(base + 8);
  auto pl = new hgsfa_plan_s();
  pl->device = device;
  pl->input_dim = hdr[0];
  pl->output_dim = hdr[1];
  const int64_t n_ops = hdr[2];
  if (n_ops <= 0 || n_ops > 4096 || pl->input_dim <= 0 || pl->output_dim <= 0) {
    delete pl;
    return fail("hgsfa_plan_create: implausible header (n_ops=%lld in=%lld out=%lld)", (long long)n_ops,
                (long long)hdr[0], (long long)hdr[1]);
  }
  // the device copy of the blob: array pointers below are offsets into it
  if (pl->params.reserve(nbytes)) { delete pl; return 1; }
  if (cudaMemcpy(pl->params.p, blob, nbytes, cudaMemcpyHostToDevice) != cudaSuccess) {
    pl->params.release(); delete pl;
    return fail("hgsfa_plan_create: parameter upload failed");
  }
  const uint8_t* dbase = static_cast<const uint8_t*>(pl->params.p);
  auto dev_ptr = [&](const void* host) { return dbase + (static_cast<const uint8_t*>(host) - base); };

  Cursor cur{base + 64, base + nbytes};
  int64_t cur_dim = pl->input_dim;
  for (int64_t o = 0; o < n_ops && cur.ok; ++o) {
    const int64_t* oh = cur.take<int64_t>(16);
    if (!oh) break;
    OpHost op{};
    OpDev& d = op.dev;
    d.n_nodes = (int)oh[0]; d.d_in = (int)oh[1]; d.in_dim = (int)oh[2]; d.out_dim = (int)oh[3];
    d.n_passes = (int)oh[4]; d.shared = (int)oh[5]; d.n_rows = (int)oh[6]; d.twc = (int)oh[7];
    op.alg_flops = oh[8]; op.exe_flops = oh[9];
    {
      double clip[2];
      std::memcpy(clip, oh + 12, sizeof(clip));
      d.clip_lo = float(clip[0]); d.clip_hi = float(clip[1]);
    }
    bool sane = d.n_nodes > 0 && d.n_nodes <= 65535 && d.d_in > 0 && d.in_dim == cur_dim && d.out_dim > 0 &&
                d.n_passes >= 1 && d.n_passes <= MAX_PASSES && (d.twc == 1 || d.twc == 2 || d.twc == 4) &&
                d.n_rows >= 0;
    if (!sane) {
      pl->params.release(); delete pl;
      return fail("hgsfa_plan_create: op %lld has an inconsistent header (nodes=%d d_in=%d in_dim=%d expected %lld)",
                  (long long)o, d.n_nodes, d.d_in, d.in_dim, (long long)cur_dim);
    }
    const int n_w = d.shared ? 1 : d.n_nodes;
    const int32_t* gather = cur.take<int32_t>(size_t(d.n_nodes) * d.d_in);
    const float* in_off = cur.take<float>(size_t(n_w) * d.d_in);
    const int32_t* out_col = cur.take<int32_t>(d.n_nodes);
    if (!cur.ok) break;
    for (size_t g = 0; g < size_t(d.n_nodes) * d.d_in; ++g)
      if (gather[g] < 0 || gather[g] >= d.in_dim) {
        pl->params.release(); delete pl;
        return fail("hgsfa_plan_create: op %lld gather index %d outside [0,%d)", (long long)o, gather[g], d.in_dim);
      }
    d.gather = reinterpret_cast<const int*>(dev_ptr(gather));
    d.in_offset = reinterpret_cast<const float*>(dev_ptr(in_off));
    d.out_col = reinterpret_cast<const int*>(dev_ptr(out_col));
    op.npad_max = 0;
    int max_wm = 1;
    for (int p = 0; p < d.n_passes && cur.ok; ++p) {
      const int64_t* ph = cur.take<int64_t>(8);
      if (!ph) break;
      PassHost& hp = op.pass[p];
      hp.K = (int)ph[0]; hp.Npad = (int)ph[1]; hp.dst = (int)ph[2]; hp.row0 = (int)ph[3]; hp.cfg = (int)ph[4];
      hp.K_real = (int)ph[5]; hp.N_real = (int)ph[6];
      bool psane = hp.K > 0 && hp.K % KC == 0 && hp.cfg >= 0 && hp.cfg <= 3 && hp.Npad > 0 &&
                   hp.Npad % CFG_NT[hp.cfg] == 0 && hp.Npad <= CFG_NT[hp.cfg] * CFG_WN[hp.cfg] &&
                   (hp.dst & (DST_GLOBAL | DST_ROWS)) && hp.row0 >= 0 &&
                   (!(hp.dst & DST_ROWS) || hp.row0 + hp.Npad <= d.n_rows);
      if (!psane) {
        pl->params.release(); delete pl;
        return fail("hgsfa_plan_create: op %lld pass %d inconsistent (K=%d Npad=%d cfg=%d dst=%d row0=%d rows=%d)",
                    (long long)o, p, hp.K, hp.Npad, hp.cfg, hp.dst, hp.row0, d.n_rows);
      }
      const Term* terms = cur.take<Term>(hp.K);
      const float* W = cur.take<float>(size_t(n_w) * hp.K * hp.Npad);
      const float* b = cur.take<float>(size_t(n_w) * hp.Npad);
      const int32_t* n_valid = cur.take<int32_t>(d.n_nodes);
      const int32_t* col_off = cur.take<int32_t>(d.n_nodes);
      if (!cur.ok) break;
      const int n_src = d.d_in + d.n_rows;
      for (int k = 0; k < hp.K; ++k) {
        const Term& t = terms[k];
        bool tok = t.op >= 0 && t.op <= OP_CLIP && t.i >= 0 && t.i < n_src;
        if (t.op == OP_MUL || t.op == OP_MUL3) tok = tok && t.j >= 0 && t.j < n_src;
        if (t.op == OP_MUL3) tok = tok && int(t.p) >= 0 && int(t.p) < n_src;
        if (!tok) {
          pl->params.release(); delete pl;
          return fail("hgsfa_plan_create: op %lld pass %d term %d invalid (op=%d i=%d j=%d)", (long long)o, p, k,
                      t.op, t.i, t.j);
        }
      }
      for (int nd = 0; nd < d.n_nodes; ++nd)
        if ((hp.dst & DST_GLOBAL) &&
            (n_valid[nd] < 0 || n_valid[nd] > hp.Npad || out_col[nd] + col_off[nd] < 0 ||
             out_col[nd] + col_off[nd] + n_valid[nd] > d.out_dim)) {
          pl->params.release(); delete pl;
          return fail("hgsfa_plan_create: op %lld pass %d node %d writes outside the output buffer", (long long)o, p, nd);
        }
      PassDev& dp = d.pass[p];
      dp.terms = reinterpret_cast<const Term*>(dev_ptr(terms));
      dp.W = reinterpret_cast<const float*>(dev_ptr(W));
      dp.b = reinterpret_cast<const float*>(dev_ptr(b));
      dp.n_valid = reinterpret_cast<const int*>(dev_ptr(n_valid));
      dp.col_off = reinterpret_cast<const int*>(dev_ptr(col_off));
      dp.K = hp.K; dp.Npad = hp.Npad; dp.dst = hp.dst; dp.row0 = hp.row0; dp.cfg = hp.cfg;
      if (hp.Npad > op.npad_max) op.npad_max = hp.Npad;
      if (CFG_WM[hp.cfg] > max_wm) max_wm = CFG_WM[hp.cfg];
      if (hp.cfg == 3) op.wide = true;
    }
    if (!cur.ok) break;
    if (d.twc < max_wm) {
      pl->params.release(); delete pl;
      return fail("hgsfa_plan_create: op %lld twc=%d smaller than a pass's tile group %d", (long long)o, d.twc, max_wm);
    }
    op.smem_bytes = sizeof(float) * (size_t(KC) * d.twc * TILE + size_t(KC) * op.npad_max +
                                     size_t(d.n_rows) * d.twc * TILE);
    if (op.smem_bytes > 227 * 1024) {
      pl->params.release(); delete pl;
      return fail("hgsfa_plan_create: op %lld needs %zu bytes of shared memory (> 227 KB)", (long long)o, op.smem_bytes);
    }
    cur_dim = d.out_dim;
    pl->ops.push_back(op);
  }
  if (!cur.ok || (int64_t)pl->ops.size() != n_ops || cur_dim != pl->output_dim) {
    pl->params.release(); delete pl;
    return fail("hgsfa_plan_create: truncated or inconsistent blob (%zu of %lld ops parsed, final dim %lld vs %lld)",
                pl->ops.size(), (long long)n_ops, (long long)cur_dim, (long long)hdr[1]);
  }
  size_t max_smem = 0;
  for (auto& op : pl->ops) max_smem = op.smem_bytes > max_smem ? op.smem_bytes : max_smem;
  cudaError_t es[4] = {
      cudaFuncSetAttribute(layer_kernel<uint8_t, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
      cudaFuncSetAttribute(layer_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
      cudaFuncSetAttribute(layer_kernel<uint8_t, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
      cudaFuncSetAttribute(layer_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem)};
  for (cudaError_t e : es)
    if (e != cudaSuccess) {
      pl->params.release(); delete pl;
      return fail("hgsfa_plan_create: cannot reserve %zu bytes of dynamic shared memory: %s", max_smem,
                  cudaGetErrorString(e));
    }
  // back segment = trailing ops with few nodes: they need many windows per launch to fill 148 SMs
  pl->split = (int)pl->ops.size();
  while (pl->split > 0 && pl->ops[pl->split - 1].dev.n_nodes <= 8) pl->split--;
  bool ok = cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i)
    ok = cudaEventCreateWithFlags(&pl->ev_h2d[i], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&pl->ev_done[i], cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreate(&pl->ev_t0) == cudaSuccess && cudaEventCreate(&pl->ev_t1) == cudaSuccess;
  if (!ok) {
    hgsfa_plan_destroy(pl);
    return fail("hgsfa_plan_create: stream / event creation failed");
  }
  *out = pl;
  return 0;
}

extern "C" int hgsfa_plan_destroy(hgsfa_plan_t pl) {
  if (!pl) return 0;
  DeviceGuard guard(pl->device);
  if (pl->stream) cudaStreamSynchronize(pl->stream);
  if (pl->copy_stream) cudaStreamSynchronize(pl->copy_stream);
  pl->params.release(); pl->tin.release(); pl->mid.release();
  for (int i = 0; i < 2; ++i) {
    pl->front[i].release(); pl->back[i].release(); pl->stage_x[i].release(); pl->stage_y[i].release();
    if (pl->ev_h2d[i]) cudaEventDestroy(pl->ev_h2d[i]);
    if (pl->ev_done[i]) cudaEventDestroy(pl->ev_done[i]);
  }
  if (pl->ev_t0) cudaEventDestroy(pl->ev_t0);
  if (pl->ev_t1) cudaEventDestroy(pl->ev_t1);
  if (pl->stream) cudaStreamDestroy(pl->stream);
  if (pl->copy_stream) cudaStreamDestroy(pl->copy_stream);
  delete pl;
  return 0;
}

extern "C" int hgsfa_plan_info(hgsfa_plan_t pl, int64_t* input_dim, int64_t* output_dim, int64_t* n_ops) {
  HG_CHECK(pl, "hgsfa_plan_info: null plan");
  if (input_dim) *input_dim = pl->input_dim;
  if (output_dim) *output_dim = pl->output_dim;
  if (n_ops) *n_ops = (int64_t)pl->ops.size();
  return 0;
}

extern "C" int hgsfa_plan_flops(hgsfa_plan_t pl, int64_t n, int x_dtype, double* alg, double* exe, double* bytes) {
  HG_CHECK(pl, "hgsfa_plan_flops: null plan");
  double a = 0, e = 0;
  for (auto& op : pl->ops) { a += double(op.alg_flops); e += double(op.exe_flops); }
  if (alg) *alg = a * double(n);
  if (exe) *exe = e * double(n);
  if (bytes) *bytes = double(n) * (double(pl->input_dim) * dtype_size(x_dtype) + double(pl->output_dim) * 4.0);
  return 0;
}

extern "C" int hgsfa_plan_stats(hgsfa_plan_t pl, int64_t* launches, double* last_ms) {
  HG_CHECK(pl, "hgsfa_plan_stats: null plan");
  if (launches) *launches = pl->launches;
  if (last_ms) {
    float ms = 0.f;
    DeviceGuard guard(pl->device);
    if (cudaEventQuery(pl->ev_t1) == cudaSuccess && cudaEventElapsedTime(&ms, pl->ev_t0, pl->ev_t1) == cudaSuccess)
      pl->last_ms = ms;
    *last_ms = pl->last_ms;
  }
  return 0;
}

extern "C" int hgsfa_plan_set_chunks(hgsfa_plan_t pl, int64_t front_chunk, int64_t back_chunk) {
  HG_CHECK(pl, "hgsfa_plan_set_chunks: null plan");
  if (front_chunk > 0) pl->front_chunk = ceil_div(front_chunk, 4 * TILE) * 4 * TILE;
  if (back_chunk > 0) pl->back_chunk = ceil_div(back_chunk, 4 * TILE) * 4 * TILE;
  if (pl->back_chunk < pl->front_chunk) pl->back_chunk = pl->front_chunk;
  pl->back_chunk = ceil_div(pl->back_chunk, pl->front_chunk) * pl->front_chunk;
  return 0;
}

extern "C" int hgsfa_tile_windows_device(const void* d_src, int dtype, int64_t n, int64_t dim, int64_t ld,
                                         void* d_dst, int dst_dtype, void* stream) {
  HG_CHECK(d_src && d_dst, "hgsfa_tile_windows_device: null pointer");
  HG_CHECK(n >= 0 && dim > 0 && ld >= dim, "hgsfa_tile_windows_device: bad shape n=%lld dim=%lld ld=%lld",
           (long long)n, (long long)dim, (long long)ld);
  if (n == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)ceil_div(dim, 64), (unsigned)ceil_div(n, TILE));
  if (dtype == HGSFA_U8 && dst_dtype == HGSFA_U8)
    tile_windows_kernel<uint8_t, uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)d_src, n, dim, ld, (uint8_t*)d_dst);
  else if (dtype == HGSFA_U8 && dst_dtype == HGSFA_F32)
    tile_windows_kernel<uint8_t, float><<<grid, 256, 0, st>>>((const uint8_t*)d_src, n, dim, ld, (float*)d_dst);
  else if (dtype == HGSFA_F32 && dst_dtype == HGSFA_F32)
    tile_windows_kernel<float, float><<<grid, 256, 0, st>>>((const float*)d_src, n, dim, ld, (float*)d_dst);
  else if (dtype == HGSFA_F64 && dst_dtype == HGSFA_F32)
    tile_windows_kernel<double, float><<<grid, 256, 0, st>>>((const double*)d_src, n, dim, ld, (float*)d_dst);
  else
    return fail("hgsfa_tile_windows_device: unsupported dtype pair %d -> %d", dtype, dst_dtype);
  HG_CUDA(cudaGetLastError());
  return 0;
}

namespace {

// run ops [o0, o1) over `ntiles` tiles.  Input of op o0 is `xin` (u8 or f32 tiled); the output of
// op o1-1 goes to `final_out`; intermediate buffers ping-pong between pp[0] / pp[1].
int run_ops(hgsfa_plan_s* pl, int o0, int o1, const void* xin, bool xin_u8, float* final_out, DevBuf* pp,
            int64_t ntiles, cudaStream_t st) {
  const void* cur = xin;
  bool cur_u8 = xin_u8;
  for (int o = o0; o < o1; ++o) {
    const OpHost& op = pl->ops[o];
    float* dst = (o == o1 - 1) ? final_out : static_cast<float*>(pp[(o - o0) & 1].p);
    int rc = cur_u8 ? launch_layer<uint8_t>(pl, op, cur, dst, ntiles, st) : launch_layer<float>(pl, op, cur, dst, ntiles, st);
    if (rc) return rc;
    cur = dst;
    cur_u8 = false;
  }
  return 0;
}

}  // namespace

extern "C" int hgsfa_plan_execute_device(hgsfa_plan_t pl, const void* d_x, int x_dtype, int x_layout, int64_t n,
                                         int64_t ld, void* d_y, int y_dtype, int64_t y_cols, void* stream) {
  HG_CHECK(pl, "hgsfa_plan_execute_device: null plan");
  HG_CHECK(n >= 0, "hgsfa_plan_execute_device: negative window count");
  HG_CHECK(y_cols > 0 && y_cols <= pl->output_dim, "hgsfa_plan_execute_device: y_cols=%lld outside (0, %lld]",
           (long long)y_cols, (long long)pl->output_dim);
  HG_CHECK(y_dtype == HGSFA_F32 || y_dtype == HGSFA_F64, "hgsfa_plan_execute_device: y dtype must be f32 or f64");
  HG_CHECK(x_dtype == HGSFA_U8 || x_dtype == HGSFA_F32 || x_dtype == HGSFA_F64, "hgsfa_plan_execute_device: bad x dtype %d", x_dtype);
  if (x_layout == HGSFA_ROWMAJOR)
    HG_CHECK(ld >= pl->input_dim, "flow input has dimension %lld, should be %lld", (long long)ld, (long long)pl->input_dim);
  else
    HG_CHECK(x_layout == HGSFA_TILED && x_dtype != HGSFA_F64, "hgsfa_plan_execute_device: tiled input must be u8 or f32");
  if (n == 0) return 0;
  HG_CHECK(d_x && d_y, "hgsfa_plan_execute_device: null buffer");
  DeviceGuard guard(pl->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->stream;

  const int n_ops = (int)pl->ops.size();
  const int split = pl->split;
  int64_t fc = pl->front_chunk, bc = pl->back_chunk;
  const int64_t n_pad = ceil_div(n, TILE) * TILE;
  if (bc > n_pad) bc = ceil_div(n_pad, 4 * TILE) * 4 * TILE;
  if (fc > bc) fc = bc;
  if (split == 0) fc = bc;        // every op is a back op: one launch sequence per back chunk
  if (split == n_ops) bc = fc;    // no back segment: nothing to accumulate across front chunks
  bc = ceil_div(bc, fc) * fc;

  int64_t maxf_front = 0, maxf_back = 0;
  for (int o = 0; o < split; ++o) maxf_front = std::max<int64_t>(maxf_front, pl->ops[o].dev.out_dim);
  for (int o = split; o < n_ops; ++o) maxf_back = std::max<int64_t>(maxf_back, pl->ops[o].dev.out_dim);
  const bool in_u8 = (x_dtype == HGSFA_U8);
  const size_t in_el = in_u8 ? 1 : 4;
  if (x_layout == HGSFA_ROWMAJOR && pl->tin.reserve(size_t(fc) * pl->input_dim * in_el)) return 1;
  if (split > 0)
    for (int i = 0; i < 2; ++i)
      if (pl->front[i].reserve(size_t(fc) * maxf_front * 4)) return 1;
  const int64_t mid_dim = split > 0 ? pl->ops[split - 1].dev.out_dim : pl->input_dim;
  if (split > 0 && split < n_ops && pl->mid.reserve(size_t(bc) * mid_dim * 4)) return 1;
  if (split < n_ops)
    for (int i = 0; i < 2; ++i)
      if (pl->back[i].reserve(size_t(bc) * maxf_back * 4)) return 1;

  HG_CUDA(cudaEventRecord(pl->ev_t0, st));
  const uint8_t* xb = static_cast<const uint8_t*>(d_x);
  auto untile = [&](const float* src, int64_t dim, int64_t w0, int64_t cnt) -> int {
    dim3 g((unsigned)ceil_div(y_cols, 32), (unsigned)ceil_div(cnt, TILE));
    if (y_dtype == HGSFA_F32)
      untile_kernel<float><<<g, 256, 0, st>>>(src, cnt, dim, y_cols, static_cast<float*>(d_y) + size_t(w0) * y_cols);
    else
      untile_kernel<double><<<g, 256, 0, st>>>(src, cnt, dim, y_cols, static_cast<double*>(d_y) + size_t(w0) * y_cols);
    pl->launches++;
    HG_CUDA(cudaGetLastError());
    return 0;
  };
  for (int64_t b0 = 0; b0 < n; b0 += bc) {
    const int64_t bn = std::min(bc, n - b0);
    const void* back_in = nullptr;   // tiled input of the back segment for this back chunk
    bool back_in_u8 = false;
    for (int64_t f0 = b0; f0 < b0 + bn; f0 += fc) {
      const int64_t fn = std::min(fc, b0 + bn - f0);
      const int64_t f_tiles = ceil_div(fn, TILE);
      // --- tiled input of the first op for windows [f0, f0 + fn) ---
      const void* xin;
      if (x_layout == HGSFA_TILED) {
        xin = xb + size_t(f0 / TILE) * pl->input_dim * TILE * in_el;
      } else {
        const void* src = xb + size_t(f0) * ld * dtype_size(x_dtype);
        if (hgsfa_tile_windows_device(src, x_dtype, fn, pl->input_dim, ld, pl->tin.p, in_u8 ? HGSFA_U8 : HGSFA_F32, st))
          return 1;
        pl->launches++;
        xin = pl->tin.p;
      }
      if (split == 0) {  // no front segment (fc == bc): the back segment reads the input directly
        back_in = xin;
        back_in_u8 = in_u8;
        continue;
      }
      float* fout = (split < n_ops) ? static_cast<float*>(pl->mid.p) + size_t((f0 - b0) / TILE) * mid_dim * TILE
                                    : static_cast<float*>(pl->front[(split - 1) & 1].p);
      if (run_ops(pl, 0, split, xin, in_u8, fout, pl->front, f_tiles, st)) return 1;
      if (split == n_ops && untile(fout, mid_dim, f0, fn)) return 1;
      back_in = pl->mid.p;
    }
    if (split < n_ops) {
      float* bout = static_cast<float*>(pl->back[(n_ops - split - 1) & 1].p);
      if (run_ops(pl, split, n_ops, back_in, back_in_u8, bout, pl->back, ceil_div(bn, TILE), st)) return 1;
      if (untile(bout, pl->ops[n_ops - 1].dev.out_dim, b0, bn)) return 1;
    }
  }
  HG_CUDA(cudaEventRecord(pl->ev_t1, st));
  return 0;
}

extern "C" int hgsfa_plan_execute(hgsfa_plan_t pl, const void* x, int x_dtype, int64_t n, int64_t ld, void* y,
                                  int y_dtype, int64_t y_cols, void* stream) {
  HG_CHECK(pl, "hgsfa_plan_execute: null plan");
  HG_CHECK(n >= 0, "hgsfa_plan_execute: negative window count");
  HG_CHECK(ld >= pl->input_dim, "flow input has dimension %lld, should be %lld", (long long)ld, (long long)pl->input_dim);
  HG_CHECK(y_cols > 0 && y_cols <= pl->output_dim, "hgsfa_plan_execute: y_cols=%lld outside (0, %lld]",
           (long long)y_cols, (long long)pl->output_dim);
  HG_CHECK(y_dtype == HGSFA_F32 || y_dtype == HGSFA_F64, "hgsfa_plan_execute: y dtype must be f32 or f64");
  HG_CHECK(x_dtype == HGSFA_U8 || x_dtype == HGSFA_F32 || x_dtype == HGSFA_F64, "hgsfa_plan_execute: bad x dtype %d", x_dtype);
  if (n == 0) return 0;
  HG_CHECK(x && y, "hgsfa_plan_execute: null buffer");
  DeviceGuard guard(pl->device);
  cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : pl->stream;

  // pieces of `piece` windows: H2D of piece i+1 (copy stream) overlaps the kernels of piece i
  int64_t piece = std::min<int64_t>(pl->back_chunk, 65536);
  if (piece > n) piece = n;
  const size_t xel = dtype_size(x_dtype), yel = dtype_size(y_dtype);
  for (int i = 0; i < 2; ++i) {
    if (pl->stage_x[i].reserve(size_t(piece) * pl->input_dim * xel)) return 1;
    if (pl->stage_y[i].reserve(size_t(piece) * y_cols * yel)) return 1;
  }
  const uint8_t* xb = static_cast<const uint8_t*>(x);
  uint8_t* yb = static_cast<uint8_t*>(y);
  int i = 0;
  for (int64_t p0 = 0; p0 < n; p0 += piece, ++i) {
    const int64_t pn = std::min(piece, n - p0);
    const int b = i & 1;
    if (i >= 2) HG_CUDA(cudaStreamWaitEvent(pl->copy_stream, pl->ev_done[b], 0));
    HG_CUDA(cudaMemcpy2DAsync(pl->stage_x[b].p, size_t(pl->input_dim) * xel, xb + size_t(p0) * ld * xel, size_t(ld) * xel,
                              size_t(pl->input_dim) * xel, size_t(pn), cudaMemcpyHostToDevice, pl->copy_stream));
    HG_CUDA(cudaEventRecord(pl->ev_h2d[b], pl->copy_stream));
    HG_CUDA(cudaStreamWaitEvent(st, pl->ev_h2d[b], 0));
    if (hgsfa_plan_execute_device(pl, pl->stage_x[b].p, x_dtype, HGSFA_ROWMAJOR, pn, pl->input_dim, pl->stage_y[b].p,
                                  y_dtype, y_cols, st))
      return 1;
    HG_CUDA(cudaMemcpyAsync(yb + size_t(p0) * y_cols * yel, pl->stage_y[b].p, size_t(pn) * y_cols * yel,
                            cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaEventRecord(pl->ev_done[b], st));
  }
  HG_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// Fused HiGSFA layer kernel, version 2 (sm_100a): TMA bulk loads + in-register expansion + FFMA2.
//
// One launch = one layer operation of the plan (csrc/flow.cu):  for every (node, tile of 128 windows)
//     x0 = X[gather[node]] - x_mean[node]                     Switchboard gather + mean subtraction
//     per pass p:  A = terms_p(x0, S)  (never materialised)   GeneralExpansion term table
//                  Y = A @ W_p[node] + b_p[node]              SFA / PCA / iGSFA projection
//                  Y -> output columns and/or shared rows S   (slow features feed the 2nd iGSFA pass)
//
// Design (what changed against v1 and why, profiles/README_r01.md):
//  * the receptive field of a node is a handful of *contiguous feature runs* of the window-minor
//    activation layout [tile][feature][128 windows]; each run is fetched with ONE cp.async.bulk
//    (TMA bulk copy) into shared memory, completion signalled on an mbarrier.  The node's parameters
//    (x_mean, b, W of every pass) are one more bulk copy.  No thread ever waits on a dependent global
//    load inside the contraction loops (v1: long_scoreboard was the top stall).
//  * loads of node i+1 are issued before node i is computed (two shared-memory stages) when they fit.
//  * expansion terms are evaluated in registers by the warp that consumes them: a lane owns 4
//    consecutive windows, reads x0 rows as one LDS.128, applies |x|^p / products, and feeds the
//    result straight into packed FP32 FMAs (fma.rn.f32x2) against weight rows that are warp-uniform
//    LDS.128 broadcasts.  Expanded features never touch shared or global memory.
//  * warps are independent inside a pass: a warp owns (tile slot, column tile of NT outputs, K-split
//    part) and runs sync-free over its rows; K-split partial sums meet once per pass in shared memory.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "hgsfa.h"

namespace hgsfa {

constexpr int TILE = HGSFA_TILE;   // windows per tile
constexpr int MAX_PASSES = 4;
constexpr int WARPS = 8;          // most warps a CTA is launched with; an op runs with op.warps of them (4 or 8)
constexpr int THREADS = WARPS * 32;

enum TermOp { OP_ID = 0, OP_MUL = 1, OP_ABSPOW = 2, OP_SGNPOW = 3, OP_MUL3 = 4, OP_ABS = 5, OP_CLIP = 6,
              OP_ID_POW = 7 /* segment-only: identity rows followed by |x|^p rows over the same inputs */ };
enum { DST_GLOBAL = 1, DST_ROWS = 2 };

struct Term16 { int16_t i, j, k, pad; };        // source rows of one expansion term
// run of terms [k0, k1) sharing one op (and exponent).  kind: 0 = all operands are receptive-field rows,
// 1 = all operands are shared rows of an earlier pass, 2 = mixed.  ibase >= 0: operand i of term k is
// row ibase + (k - k0) (no term-table lookup), else read from the term table.
struct Seg { int32_t op, k0, k1; float p; int32_t kind, ibase, nomean, pad1; };   // nomean: x_mean is folded into the bias
struct Run { int32_t i0, f0, len, pad; };       // rows [i0, i0+len) of x0 = features [f0, f0+len) of the input

struct PassDev {
  int K, Npad, NT, NTL, KS, TW, SW;   // contraction size, padded columns, columns per warp, column tiles, K-split,
                                      // slot groups per round, tile slots per warp (register tile = SW*4 windows x NT)
  int dst, row0;                  // destination flags, first shared row
  int w_off, b_off;               // float offsets of W[K][Npad] and b[Npad] inside the node parameter block
  int term_off, n_seg;
  const Seg* segs;
  const int* n_valid;             // [n_nodes] columns written to the output buffer
  const int* col_off;             // [n_nodes] first column relative to the node's out_col
};

struct OpDev {
  int n_nodes, d_in, in_dim, out_dim, n_passes, shared, n_rows, twc;
  int npc, n_runs, nstages, param_floats, n_terms, warps, simple;
  float clip_lo, clip_hi;
  const Run* runs;        // [n_nodes][n_runs]
  const int* out_col;     // [n_nodes]
  const float* params;    // [n_w][param_floats]: x_mean[d_in (pad 4)] | per pass b[Npad], W[K][Npad]
  const Term16* terms;    // [n_terms]
  // shared-memory layout (bytes), filled by the host for the input element size in use
  int sm_terms, sm_stage0, sm_stage_bytes, sm_raw_bytes, sm_srows, sm_scratch;
  PassDev pass[MAX_PASSES];
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA), packed FMA
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)      // suspend-time hint: sleep in hardware instead of polling
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a2, unsigned long long w2) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a2), "l"(w2));
}
__device__ __forceinline__ unsigned long long dup2(float a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ unsigned long long pack2(float2 v) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(v.x), "f"(v.y));
  return r;
}
__device__ __forceinline__ float2 unpack2(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}

__device__ __forceinline__ float abspow(float x, float p) {
  // |x|^p, p > 0, on the raw MUFU approximations (lg2(0) = -inf -> ex2(-inf) = 0); .ftz: operands are
  // pixel / feature magnitudes, never denormal in a way that matters at 1e-3 x std
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fabsf(x)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p * l));
  return r;
}
__device__ __forceinline__ float4 clamp4(float4 v, float lo, float hi) {
  return make_float4(fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi), fminf(fmaxf(v.z, lo), hi),
                     fminf(fmaxf(v.w, lo), hi));
}
// 4 packed uint8 -> 4 floats without the XU-pipe I2F: splice each byte into the mantissa of 2^23
__device__ __forceinline__ float4 u8x4_to_float4(uint32_t u) {
  const float magic = 8388608.0f;
  return make_float4(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7650)) - magic,
                     __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7651)) - magic,
                     __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7652)) - magic,
                     __uint_as_float(__byte_perm(u, 0x4B000000u, 0x7653)) - magic);
}

// ------------------------------------------------------------------------------------------------
// operand fetch: row i of the node's source space for this lane's 4 windows
//   raw(i)  : receptive-field row i (IN_T in the staged buffer), minus x_mean[i]
//   srow(r) : shared row r written by an earlier pass (float)
// ------------------------------------------------------------------------------------------------
template <typename IN_T>
struct Rows {
  const IN_T* rawp;     // [d_in][128] of this tile slot, already offset to this lane's 4 windows
  const float* srowp;   // [n_rows][128] of this tile slot, already offset to this lane's 4 windows
  const float* mean;    // [d_in] in the staged parameter block
  int d_in;
  __device__ __forceinline__ float4 raw_nomean(int i) const {
    if (sizeof(IN_T) == 1) return u8x4_to_float4(*reinterpret_cast<const uint32_t*>(rawp + size_t(i) * TILE));
    return *reinterpret_cast<const float4*>(rawp + size_t(i) * TILE);
  }
  __device__ __forceinline__ float4 raw(int i) const {
    float4 v = raw_nomean(i);
    const float m = mean[i];
    v.x -= m; v.y -= m; v.z -= m; v.w -= m;
    return v;
  }
  __device__ __forceinline__ float4 srow(int r) const { return *reinterpret_cast<const float4*>(srowp + size_t(r) * TILE); }
  __device__ __forceinline__ float4 any(int i) const { return i < d_in ? raw(i) : srow(i - d_in); }
};

// acc[s][r][q] += a[s].r * (w[2q], w[2q+1]) for the SW tile slots of this warp; the weight row is read once
template <int NT, int SW>
__device__ __forceinline__ void fma_row(unsigned long long (&acc)[SW][4][NT / 2], const float4 (&a)[SW], const float* wrow) {
  const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(wrow);
#pragma unroll
  for (int q = 0; q < NT / 4; ++q) {
    const ulonglong2 w = w2[q];
#pragma unroll
    for (int s = 0; s < SW; ++s) {
      const unsigned long long ax = dup2(a[s].x), ay = dup2(a[s].y), az = dup2(a[s].z), aw = dup2(a[s].w);
      ffma2(acc[s][0][2 * q], ax, w.x); ffma2(acc[s][0][2 * q + 1], ax, w.y);
      ffma2(acc[s][1][2 * q], ay, w.x); ffma2(acc[s][1][2 * q + 1], ay, w.y);
      ffma2(acc[s][2][2 * q], az, w.x); ffma2(acc[s][2][2 * q + 1], az, w.y);
      ffma2(acc[s][3][2 * q], aw, w.x); ffma2(acc[s][3][2 * q + 1], aw, w.y);
    }
  }
}

__device__ __forceinline__ float4 pow4(const float4 a, float p) {
  return make_float4(abspow(a.x, p), abspow(a.y, p), abspow(a.z, p), abspow(a.w, p));
}
__device__ __forceinline__ float4 mul4(const float4 a, const float4 b) {
  return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}

// ------------------------------------------------------------------------------------------------
// software-pipelined row loop.  A warp issues in order, so an operand chain (LDS -> convert -> pow)
// placed right before the FMAs that need it stalls the whole warp.  Each step therefore
//   (A) issues the shared-memory loads of row it+2,
//   (B) converts / expands the operands of row it+1 (loaded one step earlier) and loads its weights,
//   (C) runs the FMAs of row it,
// so every dependent instruction finds its inputs at least one full row of FMAs old.
// ------------------------------------------------------------------------------------------------
enum SegMode { M_ID_NOMEAN = 0, M_ID_MEAN = 1, M_ID_SROW = 2, M_POW = 3, M_MUL = 4, M_ID_POW = 5 };

template <int NT>
struct WReg { ulonglong2 q[NT / 4]; };

template <int NT>
__device__ __forceinline__ void load_w(WReg<NT>& w, const float* wrow) {
  const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(wrow);
#pragma unroll
  for (int q = 0; q < NT / 4; ++q) w.q[q] = w2[q];
}

template <int NT, int SW>
__device__ __forceinline__ void fma_regs(unsigned long long (&acc)[SW][4][NT / 2], const float4 (&a)[SW], const WReg<NT>& w) {
#pragma unroll
  for (int s = 0; s < SW; ++s) {
    const unsigned long long ax = dup2(a[s].x), ay = dup2(a[s].y), az = dup2(a[s].z), aw = dup2(a[s].w);
#pragma unroll
    for (int q = 0; q < NT / 4; ++q) {
      ffma2(acc[s][0][2 * q], ax, w.q[q].x); ffma2(acc[s][0][2 * q + 1], ax, w.q[q].y);
      ffma2(acc[s][1][2 * q], ay, w.q[q].x); ffma2(acc[s][1][2 * q + 1], ay, w.q[q].y);
      ffma2(acc[s][2][2 * q], az, w.q[q].x); ffma2(acc[s][2][2 * q + 1], az, w.q[q].y);
      ffma2(acc[s][3][2 * q], aw, w.q[q].x); ffma2(acc[s][3][2 * q + 1], aw, w.q[q].y);
    }
  }
}

// Base pointers of the SW tile slots a warp works on (already offset to this lane's 4 windows); slot s
// lives raw_stride / srow_stride elements after slot 0.
template <typename IN_T>
struct SlotPtrs {
  const IN_T* raw0;
  const float* srow0;
  const float* mean;
  int raw_stride, srow_stride;
};

// un-converted operands of one row for SW slots
template <typename IN_T, int SW>
struct RawRow {
  float4 f[SW];     // float sources (f32 rows / shared rows), or first operand of a product
  float4 g[SW];     // second operand of a product
  uint32_t u[SW];   // uint8 sources
  uint32_t v[SW];
  float m, m2;      // x_mean of the row(s)
};

template <typename IN_T, int SW, int MODE>
__device__ __forceinline__ void load_raw(RawRow<IN_T, SW>& r, const SlotPtrs<IN_T>& sp, int i, int j) {
  if (MODE == M_ID_SROW) {
    const float* q = sp.srow0 + i * TILE;
#pragma unroll
    for (int s = 0; s < SW; ++s) r.f[s] = *reinterpret_cast<const float4*>(q + s * sp.srow_stride);
    return;
  }
  const IN_T* q = sp.raw0 + i * TILE;
  const IN_T* q2 = sp.raw0 + j * TILE;
#pragma unroll
  for (int s = 0; s < SW; ++s) {
    if (sizeof(IN_T) == 1) {
      r.u[s] = *reinterpret_cast<const uint32_t*>(q + s * sp.raw_stride);
      if (MODE == M_MUL) r.v[s] = *reinterpret_cast<const uint32_t*>(q2 + s * sp.raw_stride);
    } else {
      r.f[s] = *reinterpret_cast<const float4*>(q + s * sp.raw_stride);
      if (MODE == M_MUL) r.g[s] = *reinterpret_cast<const float4*>(q2 + s * sp.raw_stride);
    }
  }
  if (MODE != M_ID_NOMEAN) r.m = sp.mean[i];
  if (MODE == M_MUL) r.m2 = sp.mean[j];
}

__device__ __forceinline__ float4 sub4(float4 v, float m) { return make_float4(v.x - m, v.y - m, v.z - m, v.w - m); }

// a = operand of the (first) FMA block; b = |x|^p operand of the second block (M_ID_POW only)
template <typename IN_T, int SW, int MODE>
__device__ __forceinline__ void convert_raw(float4 (&a)[SW], float4 (&b)[SW], const RawRow<IN_T, SW>& r, float p) {
#pragma unroll
  for (int s = 0; s < SW; ++s) {
    float4 x = (MODE != M_ID_SROW && sizeof(IN_T) == 1) ? u8x4_to_float4(r.u[s]) : r.f[s];
    if (MODE == M_ID_MEAN || MODE == M_POW || MODE == M_MUL || MODE == M_ID_POW) x = sub4(x, r.m);
    if (MODE == M_ID_POW) b[s] = pow4(x, p);
    if (MODE == M_POW) x = pow4(x, p);
    if (MODE == M_MUL) {
      float4 y = (sizeof(IN_T) == 1) ? u8x4_to_float4(r.v[s]) : r.g[s];
      x = mul4(x, sub4(y, r.m2));
    }
    a[s] = x;
  }
}

// Rows i0, i0+istep, ... (n of them), weight rows w0, w0+wstep, ...  M_MUL takes its operand rows from the
// term table (terms[0], terms[tstep], ...).  M_ID_POW feeds every row to two weight rows: w0 + it*wstep
// (identity) and w0 + pow_off + it*wstep (|x|^p) -- the [identity, |x|^0.8] expansion reads and centres
// each input once.  Indices are advanced with clamped additions: the pipeline fetches up to two rows
// past the end, which must stay inside the segment.
template <typename IN_T, int NT, int SW, int MODE>
__device__ __forceinline__ void seg_pipeline(unsigned long long (&acc)[SW][4][NT / 2], const SlotPtrs<IN_T>& sp,
                                             const float* w0, int wstep, int pow_off, int i0, int istep, int n, float p,
                                             const Term16* terms, int tstep) {
  if (n <= 0) return;
  RawRow<IN_T, SW> R[2];
  float4 a[2][SW], b[2][SW];
  WReg<NT> w[2];
  WReg<NT> wp;
  const int i_last = i0 + (n - 1) * istep;
  const float* w_last = w0 + (n - 1) * wstep;
  const Term16* t_last = terms + (n - 1) * tstep;
  int i_ld = i0;                 // row whose raw operands are loaded next
  const Term16* t_ld = terms;
  const float* w_ld = w0;        // weight row loaded next
  auto load_next_raw = [&](RawRow<IN_T, SW>& dst) {
    if (MODE == M_MUL) {
      const Term16 t = *t_ld;
      load_raw<IN_T, SW, MODE>(dst, sp, t.i, t.j);
      t_ld = (t_ld + tstep <= t_last) ? t_ld + tstep : t_last;
    } else {
      load_raw<IN_T, SW, MODE>(dst, sp, i_ld, 0);
      i_ld = min(i_ld + istep, i_last);
    }
  };
  auto next_w = [&]() {
    const float* r = w_ld;
    w_ld = (w_ld + wstep <= w_last) ? w_ld + wstep : w_last;
    return r;
  };
  load_next_raw(R[0]);
  load_next_raw(R[1]);
  const float* w_cur = next_w();
  load_w<NT>(w[0], w_cur);
  convert_raw<IN_T, SW, MODE>(a[0], b[0], R[0], p);
  int it = 0;
#pragma unroll 1
  for (; it + 1 < n; it += 2) {
    // ---- step it: buffers 0 hold the current row
    if (MODE == M_ID_POW) load_w<NT>(wp, w_cur + pow_off);
    load_next_raw(R[0]);
    w_cur = next_w();
    load_w<NT>(w[1], w_cur);
    fma_regs<NT, SW>(acc, a[0], w[0]);
    convert_raw<IN_T, SW, MODE>(a[1], b[1], R[1], p);
    if (MODE == M_ID_POW) fma_regs<NT, SW>(acc, b[0], wp);
    // ---- step it + 1: buffers 1
    if (MODE == M_ID_POW) load_w<NT>(wp, w_cur + pow_off);
    load_next_raw(R[1]);
    w_cur = next_w();
    load_w<NT>(w[0], w_cur);
    fma_regs<NT, SW>(acc, a[1], w[1]);
    convert_raw<IN_T, SW, MODE>(a[0], b[0], R[0], p);
    if (MODE == M_ID_POW) fma_regs<NT, SW>(acc, b[1], wp);
  }
  if (it < n) {
    if (MODE == M_ID_POW) load_w<NT>(wp, w_cur + pow_off);
    fma_regs<NT, SW>(acc, a[0], w[0]);
    if (MODE == M_ID_POW) fma_regs<NT, SW>(acc, b[0], wp);
  }
}

// bias + saturation + store of 4 windows x 1 column
struct Epilogue {
  const OpDev& op;
  const PassDev& ps;
  const float* bias;   // staged
  float* xout;
  float* srow;         // [n_rows][128] of the tile slot
  int64_t tile;
  int nvalid, col0, lane;
  __device__ __forceinline__ void operator()(int n, float4 v) const {
    const float b = bias[n];
    v.x += b; v.y += b; v.z += b; v.w += b;
    if ((ps.dst & DST_GLOBAL) && n < nvalid)
      reinterpret_cast<float4*>(xout)[(size_t(tile) * op.out_dim + col0 + n) * (TILE / 4) + lane] =
          clamp4(v, op.clip_lo, op.clip_hi);
    if (ps.dst & DST_ROWS) reinterpret_cast<float4*>(srow)[(ps.row0 + n) * (TILE / 4) + lane] = v;
  }
};

template <typename IN_T, int NT, int SW>
__device__ __forceinline__ void run_pass(const OpDev& op, const PassDev& ps, int node, int64_t tile0, int64_t ntiles,
                                         const uint8_t* stage, uint8_t* smem, float* __restrict__ xout) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ks = warp % ps.KS;
  const int nt = (warp / ps.KS) % ps.NTL;
  const int tw = warp / (ps.KS * ps.NTL);
  const int n0 = nt * NT;
  const float* params = reinterpret_cast<const float*>(stage + size_t(op.twc) * op.sm_raw_bytes);
  const float* W = params + ps.w_off + n0;
  const Term16* terms = reinterpret_cast<const Term16*>(smem + op.sm_terms) + ps.term_off;
  float* scratch = reinterpret_cast<float*>(smem + op.sm_scratch);
  const int nvalid = (ps.dst & DST_GLOBAL) ? __ldg(ps.n_valid + node) : 0;
  const int col0 = (ps.dst & DST_GLOBAL) ? (__ldg(op.out_col + node) + __ldg(ps.col_off + node)) : 0;
  const int Npad = ps.Npad, KS = ps.KS;
  const int wstep = KS * Npad;

  for (int g0 = 0; g0 < op.twc; g0 += ps.TW * SW) {
    const int slot0 = g0 + tw * SW;
    const bool active = slot0 < op.twc && tile0 + slot0 < ntiles && n0 < Npad;
    unsigned long long acc[SW][4][NT / 2];
#pragma unroll
    for (int s = 0; s < SW; ++s)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < NT / 2; ++q) acc[s][r][q] = 0ull;

    if (active) {
      // a slot past the end of the batch holds stale shared memory: computed, never stored
      const int d_in = op.d_in, n_rows = op.n_rows;
      SlotPtrs<IN_T> sp;
      sp.raw0 = reinterpret_cast<const IN_T*>(stage + size_t(slot0) * op.sm_raw_bytes) + lane * 4;
      sp.srow0 = reinterpret_cast<const float*>(smem + op.sm_srows) + size_t(slot0) * n_rows * TILE + lane * 4;
      sp.mean = params;
      sp.raw_stride = (slot0 + 1 < op.twc) ? op.sm_raw_bytes / int(sizeof(IN_T)) : 0;
      sp.srow_stride = (slot0 + 1 < op.twc) ? n_rows * TILE : 0;
      Rows<IN_T> rows[SW];   // generic (rare) path
#pragma unroll
      for (int s = 0; s < SW; ++s) {
        rows[s].rawp = sp.raw0 + s * sp.raw_stride;
        rows[s].srowp = sp.srow0 + s * sp.srow_stride;
        rows[s].mean = params;
        rows[s].d_in = d_in;
      }
      const int n_seg = ps.n_seg;
      const Seg* segs = ps.segs;
      for (int sgi = 0; sgi < n_seg; ++sgi) {
        const Seg sg = segs[sgi];
        const int kb = sg.k0 + ks;
        const float* w = W + kb * Npad;
        float4 a[SW];
        if (sg.op == OP_ID_POW) {
          const int half = (sg.k1 - sg.k0) >> 1;          // identity rows [k0, k0+half), |x|^p rows after them
          const int n_here = (sg.k0 + half - kb + KS - 1) / KS;
          seg_pipeline<IN_T, NT, SW, M_ID_POW>(acc, sp, w, wstep, half * Npad, sg.ibase + ks, KS, n_here, sg.p, terms, 0);
          continue;
        }
        const int n_here = (sg.k1 - kb + KS - 1) / KS;   // rows of this segment owned by this K-split part
        if (sg.op == OP_ID && sg.kind == 0 && sg.ibase >= 0 && sg.nomean) {
          seg_pipeline<IN_T, NT, SW, M_ID_NOMEAN>(acc, sp, w, wstep, 0, sg.ibase + ks, KS, n_here, 0.f, terms, 0);
        } else if (sg.op == OP_ID && sg.kind == 0 && sg.ibase >= 0) {
          seg_pipeline<IN_T, NT, SW, M_ID_MEAN>(acc, sp, w, wstep, 0, sg.ibase + ks, KS, n_here, 0.f, terms, 0);
        } else if (sg.op == OP_ID && sg.kind == 1 && sg.ibase >= 0) {
          seg_pipeline<IN_T, NT, SW, M_ID_SROW>(acc, sp, w, wstep, 0, sg.ibase - d_in + ks, KS, n_here, 0.f, terms, 0);
        } else if (sg.op == OP_ABSPOW && sg.kind == 0 && sg.ibase >= 0) {
          seg_pipeline<IN_T, NT, SW, M_POW>(acc, sp, w, wstep, 0, sg.ibase + ks, KS, n_here, sg.p, terms, 0);
        } else if (sg.op == OP_MUL && sg.kind == 0) {
          seg_pipeline<IN_T, NT, SW, M_MUL>(acc, sp, w, wstep, 0, 0, 0, n_here, 0.f, terms + kb, KS);
        } else {
          // everything else (rare): generic operand fetch, one term at a time
          for (int k = kb; k < sg.k1; k += KS, w += wstep) {
            const Term16 t = terms[k];
            const int i = sg.ibase >= 0 ? sg.ibase + (k - sg.k0) : int(t.i);
#pragma unroll
            for (int s = 0; s < SW; ++s) {
              const float4 x = rows[s].any(i);
              float4 v;
              switch (sg.op) {
                case OP_ID: v = x; break;
                case OP_MUL: v = mul4(x, rows[s].any(t.j)); break;
                case OP_ABSPOW: v = pow4(x, sg.p); break;
                case OP_SGNPOW: {
                  const float4 m = pow4(x, sg.p);
                  v = make_float4(copysignf(m.x, x.x), copysignf(m.y, x.y), copysignf(m.z, x.z), copysignf(m.w, x.w));
                  break;
                }
                case OP_MUL3: v = mul4(mul4(x, rows[s].any(t.j)), rows[s].any(t.k)); break;
                case OP_ABS: v = make_float4(fabsf(x.x), fabsf(x.y), fabsf(x.z), fabsf(x.w)); break;
                case OP_CLIP: v = clamp4(x, -sg.p, sg.p); break;
                default: v = make_float4(0.f, 0.f, 0.f, 0.f); break;
              }
              a[s] = v;
            }
            fma_row<NT, SW>(acc, a, w);
          }
        }
      }
    }

    // ---- K-split: pairwise tree over the KS warps of a group (consecutive warps).  In every round the
    // upper half parks its partial sums in shared memory and the lower half adds them, so the scratch
    // area holds at most WARPS / 2 register tiles.
    if (KS > 1) {
      const int grp_slot0 = (warp / KS) * (KS >> 1);
      for (int stride = KS >> 1; stride >= 1; stride >>= 1) {
        if (active && ks >= stride && ks < 2 * stride) {
          float4* slot4 = reinterpret_cast<float4*>(scratch) + size_t(grp_slot0 + ks - stride) * SW * NT * (TILE / 4) + lane;
#pragma unroll
          for (int s = 0; s < SW; ++s)
#pragma unroll
            for (int q = 0; q < NT / 2; ++q) {
              const float2 y0 = unpack2(acc[s][0][q]), y1 = unpack2(acc[s][1][q]), y2 = unpack2(acc[s][2][q]), y3 = unpack2(acc[s][3][q]);
              slot4[(s * NT + 2 * q) * (TILE / 4)] = make_float4(y0.x, y1.x, y2.x, y3.x);
              slot4[(s * NT + 2 * q + 1) * (TILE / 4)] = make_float4(y0.y, y1.y, y2.y, y3.y);
            }
        }
        __syncthreads();
        if (active && ks < stride) {
          const float4* part = reinterpret_cast<const float4*>(scratch) + size_t(grp_slot0 + ks) * SW * NT * (TILE / 4) + lane;
#pragma unroll
          for (int s = 0; s < SW; ++s)
#pragma unroll
            for (int q = 0; q < NT / 2; ++q) {
              const float4 u0 = part[(s * NT + 2 * q) * (TILE / 4)], u1 = part[(s * NT + 2 * q + 1) * (TILE / 4)];
              float2 y0 = unpack2(acc[s][0][q]), y1 = unpack2(acc[s][1][q]), y2 = unpack2(acc[s][2][q]), y3 = unpack2(acc[s][3][q]);
              y0.x += u0.x; y1.x += u0.y; y2.x += u0.z; y3.x += u0.w;
              y0.y += u1.x; y1.y += u1.y; y2.y += u1.z; y3.y += u1.w;
              acc[s][0][q] = pack2(y0); acc[s][1][q] = pack2(y1); acc[s][2][q] = pack2(y2); acc[s][3][q] = pack2(y3);
            }
        }
        __syncthreads();   // scratch slots are reused by the next round / pass
      }
    }
    if (active && ks == 0) {
#pragma unroll
      for (int s = 0; s < SW; ++s) {
        const int slot = slot0 + s;
        const int64_t tile = tile0 + slot;
        if (slot < op.twc && tile < ntiles) {
          Epilogue epi{op, ps, params + ps.b_off, xout,
                       reinterpret_cast<float*>(smem + op.sm_srows) + size_t(slot) * op.n_rows * TILE, tile, nvalid, col0, lane};
#pragma unroll
          for (int q = 0; q < NT / 2; ++q) {
            const float2 y0 = unpack2(acc[s][0][q]), y1 = unpack2(acc[s][1][q]), y2 = unpack2(acc[s][2][q]), y3 = unpack2(acc[s][3][q]);
            epi(n0 + 2 * q, make_float4(y0.x, y1.x, y2.x, y3.x));
            epi(n0 + 2 * q + 1, make_float4(y0.y, y1.y, y2.y, y3.y));
          }
        }
      }
    }
  }
}

// NTMAX bounds the register tile compiled in: 16 (<= 64 accumulator registers) keeps 2 CTAs per SM,
// 32 adds the 128-register tiles (NT 24 / 32, or NT 16 over two tile slots).
template <typename IN_T, int NTMAX>
__global__ void __launch_bounds__(THREADS, (NTMAX <= 16 ? 2 : 1))
    layer_kernel(const OpDev op, const IN_T* __restrict__ xin, float* __restrict__ xout, int64_t ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [4] "stage filled" mbarriers
  int* released = reinterpret_cast<int*>(smem + 64);           // [4] warps done with a stage (simple ops)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  const int64_t tile0 = int64_t(blockIdx.x) * op.twc;
  const int node_begin = blockIdx.y * op.npc;
  const int node_end = min(op.n_nodes, node_begin + op.npc);
  const int64_t tiles_left = ntiles - tile0;
  const int valid_tiles = tiles_left < op.twc ? (int)tiles_left : op.twc;
  const int nst = op.nstages;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&full[i], 1);
      released[i] = 0;
    }
    mbar_fence_init();
  }
  {  // term table of all passes (shared by every node of the op)
    const uint2* src = reinterpret_cast<const uint2*>(op.terms);
    uint2* dst = reinterpret_cast<uint2*>(smem + op.sm_terms);
    for (int i = tid; i < op.n_terms; i += blockDim.x) dst[i] = __ldg(src + i);
  }
  __syncthreads();

  // producer (any one warp): bulk copies of one node's receptive field (valid tile slots) + parameter block
  auto issue = [&](int node, int s) {
    uint8_t* stage = smem + op.sm_stage0 + size_t(s) * op.sm_stage_bytes;
    const Run* runs = op.runs + size_t(node) * op.n_runs;
    const int n_copies = 1 + valid_tiles * op.n_runs;
    if (lane == 0) {
      uint32_t bytes = uint32_t(op.param_floats) * 4u;
      for (int r = 0; r < op.n_runs; ++r) bytes += uint32_t(valid_tiles) * uint32_t(runs[r].len) * TILE * sizeof(IN_T);
      mbar_expect_tx(&full[s], bytes);
    }
    __syncwarp();
    for (int c = lane; c < n_copies; c += 32) {
      if (c == 0) {
        const int nw = op.shared ? 0 : node;
        bulk_g2s(stage + size_t(op.twc) * op.sm_raw_bytes, op.params + size_t(nw) * op.param_floats,
                 uint32_t(op.param_floats) * 4u, &full[s]);
      } else {
        const int slot = (c - 1) / op.n_runs;
        const Run r = runs[(c - 1) % op.n_runs];
        if (r.len > 0)
          bulk_g2s(stage + size_t(slot) * op.sm_raw_bytes + size_t(r.i0) * TILE * sizeof(IN_T),
                   xin + (size_t(tile0 + slot) * op.in_dim + r.f0) * TILE, uint32_t(r.len) * TILE * sizeof(IN_T), &full[s]);
      }
    }
  };

  if (warp == 0)
    for (int i = 0; i < nst && node_begin + i < node_end; ++i) issue(node_begin + i, i);

  int it = 0;
  for (int node = node_begin; node < node_end; ++node, ++it) {
    const int s = it % nst;
    mbar_wait(&full[s], uint32_t(it / nst) & 1u);
    const uint8_t* stage = smem + op.sm_stage0 + size_t(s) * op.sm_stage_bytes;
#pragma unroll 1
    for (int p = 0; p < op.n_passes; ++p) {
      const PassDev& ps = op.pass[p];
      const int code = ps.NT * 4 + ps.SW;
#define HG_PASS(NT_, SW_) run_pass<IN_T, NT_, SW_>(op, ps, node, tile0, ntiles, stage, smem, xout)
      switch (code) {
        // <= 64 accumulator registers per thread
        case 8 * 4 + 1: HG_PASS(8, 1); break;
        case 12 * 4 + 1: HG_PASS(12, 1); break;
        case 16 * 4 + 1: HG_PASS(16, 1); break;
        case 8 * 4 + 2: HG_PASS(8, 2); break;
        // up to 128 accumulator registers per thread
        case 12 * 4 + 2: if constexpr (NTMAX >= 32) HG_PASS(12, 2); break;
        case 16 * 4 + 2: if constexpr (NTMAX >= 32) HG_PASS(16, 2); break;
        case 20 * 4 + 1: if constexpr (NTMAX >= 32) HG_PASS(20, 1); break;
        case 24 * 4 + 1: if constexpr (NTMAX >= 32) HG_PASS(24, 1); break;
        case 28 * 4 + 1: if constexpr (NTMAX >= 32) HG_PASS(28, 1); break;
        case 32 * 4 + 1: if constexpr (NTMAX >= 32) HG_PASS(32, 1); break;
        default: break;
      }
#undef HG_PASS
      if (!op.simple) __syncthreads();   // shared rows of this pass visible; stage consumed after the last pass
    }
    // ---- hand the stage back.  Simple ops (one pass, no K-split) have no CTA-wide barrier at all: warps
    // drift apart by up to nst nodes, and the LAST warp to leave a stage refills it for node + nst.
    if (op.simple) {
      int last = 0;
      __syncwarp();
      if (lane == 0) {
        const int old = atomicAdd(&released[s], 1);
        if (old == nwarps - 1) {
          released[s] = 0;
          last = 1;
        }
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      if (last && node + nst < node_end) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(node + nst, s);
      }
    } else if (warp == 0 && node + nst < node_end) {
      issue(node + nst, s);
    }
  }
}

}  // namespace hgsfa

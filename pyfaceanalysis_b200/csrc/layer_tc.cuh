// Fused HiGSFA layer kernel on the 5th-generation tensor cores (sm_100a: tcgen05 + tensor memory).
//
// Same operation as layer.cuh (one launch = one single-pass layer op of the plan):
//     Y[tile, node] = terms(X[tile, gather[node]] - x_mean[node]) @ W[node] + b[node]
// but the contraction runs as 3xTF32 tcgen05.mma with M = the 128 windows of a tile:
//     E = Ehi + Elo, W = Whi + Wlo (TF32 each);   Y ~= Ehi Whi + Ehi Wlo + Elo Whi   (FP32 accumulate in TMEM)
// which keeps near-FP32 accuracy (5e-7 relative per contraction, tools/tc_probe.cu; 1.7e-4 x std end to end) and
// takes the contraction off the FMA pipe: ~9 ms of tensor-pipe time per 1 Mi windows of U11L_64, so a step is bound
// by the expansion code (34 ms against 61 ms for the FFMA2 kernel).
//
// Roles inside a CTA (320 threads):
//   warps 0-3  expansion: thread = window.  A thread evaluates the expansion terms of ITS window from the
//              staged receptive field (conflict-free LDS: consecutive lanes = consecutive windows), splits
//              every value into TF32 hi/lo and writes them to ITS tensor-memory lane (tcgen05.st 32x32b):
//              the A operand never touches shared memory.
//   warps 4-7  epilogue: thread = window again (same lane quarters).  Drain the accumulators (tcgen05.ld), add the
//              bias, clip and store the window-minor output (coalesced) while the expansion warps are already on
//              the next node: an in-order warp that also had to wait for its own MMAs was the bottleneck
//              (35.7 -> 34.4 ms per step; HGSFA_TC_EPI=0 keeps the epilogue on the expansion warps).
//   warp 8     MMA issue: one elected lane, operands in uniform registers, 3 MMAs per 8 terms.
//   warp 9     producer: cp.async.bulk of receptive-field runs (+ x_mean | b) and of the weight chunks.
// Loop order node -> term chunk (32 terms) -> tile: a weight chunk (hi and lo image, canonical K-major
// no-swizzle core matrices, prepared on the host) is streamed ONCE per node through a small ring and
// shared by the twc tiles of the CTA; the accumulators of all twc tiles stay live in tensor memory.
// All hand-overs are mbarriers (full/free pairs); tcgen05.commit releases A stages, weight stages and
// accumulators.  With two accumulator sets the epilogue of node i runs behind the MMAs of node i+1.
#pragma once
#include <cuda_fp16.h>

#include "layer.cuh"

namespace hgsfa {

#ifndef HGSFA_TC_CK
#define HGSFA_TC_CK 32
#endif
constexpr int TC_CK = HGSFA_TC_CK;   // terms per chunk of the TF32 form (A stage = TC_CK hi + TC_CK lo columns)
// The FP16 form packs two terms per column: a chunk of 2 x TC_CK terms has the SAME footprint -- TC_CK hi + TC_CK lo columns,
// a weight chunk of the same bytes -- and half the hand-overs per term (a hand-over costs ~700 cycles, a 32-term chunk of
// arithmetic ~1 200: profiles/README_r02.md item 14).
constexpr int TC_CK16 = 2 * TC_CK;
__host__ __device__ constexpr int tc_chunk_terms(bool f16) { return f16 ? TC_CK16 : TC_CK; }
#ifndef HGSFA_TC_EPI
#define HGSFA_TC_EPI 1
#endif
// HGSFA_TC_EPI = 1: four dedicated epilogue warps (4-7) drain the accumulators; 0: the expansion warps do it
constexpr int TC_EPI = HGSFA_TC_EPI;
constexpr int TC_MMA_WARP = TC_EPI ? 8 : 4, TC_PROD_WARP = TC_MMA_WARP + 1;
constexpr int TC_THREADS = (TC_PROD_WARP + 1) * 32;
constexpr int TC_MAX_TW = 8;
// ns between barrier tries of the warps that mostly wait (round 2: try_wait returns within tens of ns inside a busy CTA)
// Measured on U11L_64 layers 3-10 (profiles/README_r02.md): 32 / 32 ns 12.41 ms, 512 / 32 12.40, 512 / 128 12.40, 2000 / 64 12.49 --
// the re-tries only fill issue slots nobody else wants; the longer sleeps keep them out of the instruction counts.
#ifndef HGSFA_TC_SLEEP_EPI
#define HGSFA_TC_SLEEP_EPI 512
#endif
#ifndef HGSFA_TC_SLEEP_MMA
#define HGSFA_TC_SLEEP_MMA 64
#endif

#ifndef HGSFA_TC_UNROLL16
#define HGSFA_TC_UNROLL16 1   // unroll factor of the 16-term segment loop (2 measured: see profiles/README_r02.md item 13)
#endif
constexpr int TC_UNROLL16 = HGSFA_TC_UNROLL16;
#ifndef HGSFA_TC_MINB
#define HGSFA_TC_MINB 1     // resident CTAs per SM the register allocation is sized for
#endif

#ifdef HGSFA_TC_TRACE
// development: time stamps of expansion warp 0 of one CTA of the op with HGSFA_TC_TRACE nodes (tools/tc_trace.py)
__device__ unsigned long long tc_trace[16384];
#define TC_T(slot)                                                                                            \
  do {                                                                                                        \
    if (tracing && tcnt < 16380) tc_trace[tcnt++] = ((unsigned long long)clock64() << 8) | (unsigned long long)(slot); \
  } while (0)
#define TC_TRACE_PARAM , bool tracing, int& tcnt
#define TC_TRACE_ARG , tracing, tcnt
#else
#define TC_T(slot) do { } while (0)
#define TC_TRACE_PARAM
#define TC_TRACE_ARG
#endif

struct TcOpDev {
  int n_nodes, d_in, in_dim, out_dim, shared, twc, npc, n_runs;
  int K, Kpad, Npad16, n_chunks, n_terms, n_segs;
  int nd, nstx, nw, na;              // accumulator sets, receptive-field stages, weight-ring stages, A stages
  int head_floats, wchunk_floats;    // x_mean[d_pad4] | b[Npad16] | (m_i, m_j)[n_terms];  one weight chunk = hi[32][Npad16] | lo[32][Npad16]
  int tmem_cols;
  int f16;                           // 1: 2-piece FP16 operands, tcgen05 kind::f16 (K = 16); 0: 3xTF32 (K = 8)
  float scale;                       // accumulator scale of the epilogue (weights are stored times 1 / scale; 1 for TF32)
  float prod_scale;                  // F16: operands of product terms are pre-scaled by this power of two (their weights by its
                                     // inverse square) so that a product of two saturated inputs stays inside FP16's range; else 1
  float clip_lo, clip_hi;
  const Run* runs;
  const int* out_col;
  const int* n_valid;
  const int* col_off;
  const float* head;                 // [n_w][head_floats]
  const float* wimg;                 // [n_w][n_chunks][wchunk_floats]
  const Term16* terms;
  const Seg* segs;                   // split at chunk boundaries
  const int* chunk_seg;              // [n_chunks + 1]
  int sm_terms, sm_toff, sm_segs, sm_chunkseg, sm_bias, sm_x0, sm_xstage_bytes, sm_raw_bytes, sm_w0, sm_wstage_bytes;
};

// ---- barriers (uint64 slots at the start of shared memory) ----
enum { TCB_XFULL = 0, TCB_XFREE = 2, TCB_WFULL = 4, TCB_WFREE = 8, TCB_AFULL = 12, TCB_AFREE = 16, TCB_DFULL = 20,
       TCB_DFREE = 36, TCB_COUNT = 52, TCB_BYTES = 512 };

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol error traps (launch failure reported to the host) instead of hanging the GPU
template <int SLEEP_NS = 32>
__device__ __forceinline__ void mbar_wait_tc(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(0x989680u)      // suspend-time hint: sleep in hardware instead of polling
        : "memory");
    if (ok) return;
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);      // a waiting warp must not eat the issue slots of the working ones
  }
  __trap();
}
// position in a ring of n slots + parity of the current round (no runtime division in the hand-over loops)
struct Ring {
  int idx, n;
  uint32_t par;
  __device__ __forceinline__ explicit Ring(int n_) : idx(0), n(n_), par(0u) {}
  __device__ __forceinline__ void next() {
    if (++idx == n) {
      idx = 0;
      par ^= 1u;
    }
  }
  __device__ __forceinline__ void advance(int k) {      // k <= n
    idx += k;
    if (idx >= n) {
      idx -= n;
      par ^= 1u;
    }
  }
};
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, no swizzle, version 1: start address, leading (K) and stride (MN) byte offsets
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t tc_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
// two FP32 values -> FP16 pair (first value in the low half = the lower K index)
__device__ __forceinline__ uint32_t fr_pack(float a, float b) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ void fr_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = fr_pack(a, b);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  lo = fr_pack(a - hf.x, b - hf.y);
}
// instruction descriptor: D = F32, A = B = F16, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t tc_idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}

// Split of 8 values and store to the hi / lo column groups of an A stage.  `col`: tensor-memory address whose column field
// counts TERMS from the start of the tensor memory (TF32: one column per term; F16: two terms per column, so the column
// field is halved here -- the callers' arithmetic on 8-term groups is the same for both forms).
template <bool F16>
__device__ __forceinline__ void tc_store8(uint32_t col_hi, const float (&v)[8]) {
  if constexpr (F16) {
    // 2-piece FP16: hi = v rounded to FP16, lo = (v - hi) rounded to FP16: 22 significant bits like the TF32 pair, one
    // 32-bit column per two terms (tools/tc_probe2.cu, profiles/README_r02.md)
    const uint32_t t = (col_hi & 0xffff0000u) | ((col_hi & 0xffffu) >> 1);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) fr_split(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
    tmem_st4(t, hi);
    tmem_st4(t + TC_CK16 / 2, lo);
  } else {
    // hi = v rounded to TF32 (nearest, ties away; two ALU ops), lo = v - hi exact in FP32.  The MMA ignores the
    // low 13 bits of its FP32 containers (tools/tc_probe.cu mode 2), i.e. it truncates lo: |error| <= 2^-21 |v|.
    // (hi = raw bits, lo = v - trunc(v) saves one op per term but doubles the error; measured 2.2e-4 x std
    // against 1.1e-4 on U11L_64, profiles/README_r01.md)
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = (__float_as_uint(v[j]) + 0x1000u) & 0xffffe000u;
      lo[j] = __float_as_uint(v[j] - __uint_as_float(hi[j]));
    }
    tmem_st8(col_hi, hi);
    tmem_st8(col_hi + TC_CK, lo);
  }
}

// values that are exactly representable in one piece (uint8 pixels as identity terms; zero padding): no split needed
template <bool F16>
__device__ __forceinline__ void tc_store8_exact(uint32_t col_hi, const float (&v)[8]) {
  if constexpr (F16) {
    const uint32_t t = (col_hi & 0xffff0000u) | ((col_hi & 0xffffu) >> 1);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = fr_pack(v[2 * j], v[2 * j + 1]);
      lo[j] = 0u;
    }
    tmem_st4(t, hi);
    tmem_st4(t + TC_CK16 / 2, lo);
  } else {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __float_as_uint(v[j]);
      lo[j] = 0u;
    }
    tmem_st8(col_hi, hi);
    tmem_st8(col_hi + TC_CK, lo);
  }
}

// one receptive-field value of this thread's window; xp already points at (row, window)
template <typename IN_T>
__device__ __forceinline__ float tc_ld(const IN_T* xp) {
  if (sizeof(IN_T) == 1) return __uint_as_float(0x4B000000u | uint32_t(*reinterpret_cast<const uint8_t*>(xp))) - 8388608.0f;
  return *reinterpret_cast<const float*>(xp);
}

// Segment of terms over consecutive receptive-field rows: MODE 0 identity (x_mean folded into the bias),
// 1 identity, 2 |x|^p.  A segment occupies a multiple of 8 A columns; the terms past `cnt` are padding
// (their weight rows are zero) and repeat the last real term so that the operand stays finite.
template <typename IN_T, int MODE>
__device__ __forceinline__ float tc_row_value(const IN_T* xp, const float* mp, int j, float p) {
  float x = tc_ld<IN_T>(xp + j * TILE);
  if (MODE != 0) x -= mp[j];
  if (MODE == 2) x = abspow(x, p);
  return x;
}
template <typename IN_T, int MODE, bool F16>
__device__ __forceinline__ void tc_seg_rows(const IN_T* xp, const float* mp, int cnt, int ngroups, float p, uint32_t col TC_TRACE_PARAM) {
  // two groups (16 independent operand chains) per iteration while both are full: a warp issues in order, and
  // with one expansion warp per scheduler and CTA the instruction-level parallelism has to come from here
#pragma unroll TC_UNROLL16
  for (; ngroups >= 2 && cnt >= 16; ngroups -= 2, cnt -= 16, xp += 16 * TILE, mp += 16, col += 16) {
    float v0[8], v1[8];
    TC_T(8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v0[j] = tc_row_value<IN_T, MODE>(xp, mp, j, p);
      v1[j] = tc_row_value<IN_T, MODE>(xp, mp, 8 + j, p);
    }
#ifdef HGSFA_TC_TRACE
    asm volatile("" ::"f"(v0[0]), "f"(v0[1]), "f"(v0[2]), "f"(v0[3]), "f"(v0[4]), "f"(v0[5]), "f"(v0[6]), "f"(v0[7]), "f"(v1[0]),
                 "f"(v1[1]), "f"(v1[2]), "f"(v1[3]), "f"(v1[4]), "f"(v1[5]), "f"(v1[6]), "f"(v1[7]));
#endif
    TC_T(9);
    if (sizeof(IN_T) == 1 && MODE == 0) {
      tc_store8_exact<F16>(col, v0);
      tc_store8_exact<F16>(col + 8, v1);
    } else {
      tc_store8<F16>(col, v0);
      tc_store8<F16>(col + 8, v1);
    }
    TC_T(10);
  }
#pragma unroll 1
  for (; ngroups > 0; --ngroups, cnt -= 8, xp += 8 * TILE, mp += 8, col += 8) {
    float v[8];
    if (cnt >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tc_row_value<IN_T, MODE>(xp, mp, j, p);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = tc_row_value<IN_T, MODE>(xp, mp, min(j, cnt - 1), p);
    }
    if (sizeof(IN_T) == 1 && MODE == 0) tc_store8_exact<F16>(col, v);
    else tc_store8<F16>(col, v);
  }
}

// Upper-triangular products x_i x_j (r0 <= i <= j < r0 + N, row-major: the QT expansion) of N centred rows.
// The N operands are loaded once into registers; the enumeration is unrolled at compile time, so a term
// costs one FMUL plus the TF32 split.  A piece covers the 8-term groups [t0 / 8, (t0 + cnt + 7) / 8).
constexpr int OP_TRI = 9;            // segment-only op code: p = N, ibase = r0, nomean = t0
__host__ __device__ constexpr bool tc_tri_size(int n) { return n == 10; }   // cuicuilco's s10 selectors (the shipped thin networks)
__host__ __device__ constexpr int tri_row(int n, int idx) {
  int i = 0, len = n;
  while (idx >= len && len > 0) { idx -= len; --len; ++i; }
  return i;
}
__host__ __device__ constexpr int tri_col(int n, int idx) {
  int i = 0, len = n;
  while (idx >= len && len > 0) { idx -= len; --len; ++i; }
  return i + idx;
}
template <int N, int G, bool F16>
__device__ __forceinline__ void tc_tri_groups(const float (&xc)[N], int g0, int g1, uint32_t col) {
  constexpr int T = N * (N + 1) / 2;
  if constexpr (8 * G < T) {
    if (G >= g0 && G < g1) {
      float v[8];
#define HG_TRI_TERM(J)                                                                       \
  {                                                                                          \
    constexpr int idx = 8 * G + J;                                                           \
    if constexpr (idx < T) v[J] = xc[tri_row(N, idx)] * xc[tri_col(N, idx)];                 \
    else v[J] = 0.f;                                                                         \
  }
      HG_TRI_TERM(0) HG_TRI_TERM(1) HG_TRI_TERM(2) HG_TRI_TERM(3) HG_TRI_TERM(4) HG_TRI_TERM(5) HG_TRI_TERM(6) HG_TRI_TERM(7)
#undef HG_TRI_TERM
      tc_store8<F16>(col + uint32_t(8 * (G - g0)), v);
    }
    tc_tri_groups<N, G + 1, F16>(xc, g0, g1, col);
  }
}
template <typename IN_T, int N, bool F16>
__device__ __forceinline__ void tc_seg_tri(const IN_T* xr, const float* mr, int t0, int cnt, uint32_t col, float sc) {
  float xc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    xc[i] = tc_ld<IN_T>(xr + i * TILE) - mr[i];
    if constexpr (F16) xc[i] *= sc;          // products of saturated inputs must stay inside FP16's range
  }
  tc_tri_groups<N, 0, F16>(xc, t0 >> 3, (t0 + cnt + 7) >> 3, col);
}

template <typename IN_T, bool F16>
__global__ void __launch_bounds__(TC_THREADS, HGSFA_TC_MINB)
    layer_tc_kernel(const TcOpDev op, const IN_T* __restrict__ xin, float* __restrict__ xout, int64_t ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TCB_COUNT * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tile0 = int64_t(blockIdx.x) * op.twc;
  const int node_begin = blockIdx.y * op.npc;
  const int node_end = min(op.n_nodes, node_begin + op.npc);
  const int64_t tiles_left = ntiles - tile0;
  const int vt = tiles_left < op.twc ? (int)tiles_left : op.twc;      // valid tile slots
  const int nd = op.nd, nstx = op.nstx, nw = op.nw, na = op.na, n_chunks = op.n_chunks;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(op.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[TCB_XFULL + i], 1); mbar_init(&bars[TCB_XFREE + i], 4); }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bars[TCB_WFULL + i], 1); mbar_init(&bars[TCB_WFREE + i], 1);
      mbar_init(&bars[TCB_AFULL + i], 4); mbar_init(&bars[TCB_AFREE + i], 1);
    }
    for (int i = 0; i < 16; ++i) { mbar_init(&bars[TCB_DFULL + i], 1); mbar_init(&bars[TCB_DFREE + i], 4); }
    mbar_fence_init();
  }
  {  // term table, segments, chunk index (shared by every node of the op)
    const uint2* src = reinterpret_cast<const uint2*>(op.terms);
    uint2* dst = reinterpret_cast<uint2*>(smem + op.sm_terms);
    int2* toff = reinterpret_cast<int2*>(smem + op.sm_toff);     // operand rows of a product as element offsets
    for (int i = tid; i < op.n_terms; i += TC_THREADS) {
      const uint2 raw = __ldg(src + i);
      dst[i] = raw;
      toff[i] = make_int2(int(int16_t(raw.x & 0xffffu)) * TILE, int(int16_t(raw.x >> 16)) * TILE);
    }
    const uint4* ss = reinterpret_cast<const uint4*>(op.segs);
    uint4* sd = reinterpret_cast<uint4*>(smem + op.sm_segs);
    for (int i = tid; i < op.n_segs * 2; i += TC_THREADS) sd[i] = __ldg(ss + i);
    int* cd = reinterpret_cast<int*>(smem + op.sm_chunkseg);
    for (int i = tid; i <= n_chunks; i += TC_THREADS) cd[i] = __ldg(op.chunk_seg + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t a_col0 = uint32_t(nd * op.twc * op.Npad16);          // A stages follow the accumulator sets
  constexpr int CK = tc_chunk_terms(F16);                              // terms per chunk
  constexpr uint32_t A_STAGE = 2 * TC_CK;                              // columns of one A stage (hi | lo), both forms
  constexpr uint32_t A_LO = TC_CK;                                     // lo pieces follow the hi pieces

  if (warp == TC_PROD_WARP) {
    // ================================ producer ================================
    Ring rx(nstx), rw(nw);
    for (int node = node_begin; node < node_end; ++node, rx.next()) {
      const int sx = rx.idx;
      mbar_wait_tc<256>(&bars[TCB_XFREE + sx], rx.par ^ 1u);
      uint8_t* stage = smem + op.sm_x0 + size_t(sx) * op.sm_xstage_bytes;
      const Run* runs = op.runs + size_t(node) * op.n_runs;
      const int nwi = op.shared ? 0 : node;
      if (lane == 0) {
        uint32_t bytes = uint32_t(op.head_floats) * 4u;
        for (int r = 0; r < op.n_runs; ++r) bytes += uint32_t(vt) * uint32_t(runs[r].len) * TILE * sizeof(IN_T);
        mbar_expect_tx(&bars[TCB_XFULL + sx], bytes);
      }
      __syncwarp();
      const int n_copies = 1 + vt * op.n_runs;
      for (int c = lane; c < n_copies; c += 32) {
        if (c == 0) {
          bulk_g2s(stage + size_t(op.twc) * op.sm_raw_bytes, op.head + size_t(nwi) * op.head_floats,
                   uint32_t(op.head_floats) * 4u, &bars[TCB_XFULL + sx]);
        } else {
          const int slot = (c - 1) / op.n_runs;
          const Run r = runs[(c - 1) % op.n_runs];
          if (r.len > 0)
            bulk_g2s(stage + size_t(slot) * op.sm_raw_bytes + size_t(r.i0) * TILE * sizeof(IN_T),
                     xin + (size_t(tile0 + slot) * op.in_dim + r.f0) * TILE, uint32_t(r.len) * TILE * sizeof(IN_T),
                     &bars[TCB_XFULL + sx]);
        }
      }
      const float* wnode = op.wimg + size_t(nwi) * n_chunks * op.wchunk_floats;
      for (int c = 0; c < n_chunks; ++c, rw.next()) {
        const int sw = rw.idx;
        mbar_wait_tc<256>(&bars[TCB_WFREE + sw], rw.par ^ 1u);
        if (lane == 0) {
          const uint32_t bytes = uint32_t(op.wchunk_floats) * 4u;
          mbar_expect_tx(&bars[TCB_WFULL + sw], bytes);
          bulk_g2s(smem + op.sm_w0 + size_t(sw) * op.sm_wstage_bytes, wnode + size_t(c) * op.wchunk_floats, bytes,
                   &bars[TCB_WFULL + sw]);
        }
        __syncwarp();
      }
    }
  } else if (warp == TC_MMA_WARP) {
    // ================================ MMA issue ================================
    const bool leader = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t idesc = F16 ? tc_idesc_f16(op.Npad16) : tc_idesc(op.Npad16);
    // canonical K-major core matrices (8 rows x 16 bytes): a K step of one MMA (8 TF32 or 16 FP16 terms) spans two of them
    const uint32_t lbo = uint32_t(op.Npad16 / 8) * 128u, sbo = 128u;
    const uint32_t lo_off = uint32_t(TC_CK * op.Npad16) * 4u;        // lo image follows the hi image (CK x Npad16 elements)
    Ring rw(nw), ra(na), rd(nd);
    for (int node = node_begin; node < node_end; ++node, rd.next()) {
      const int set = rd.idx;
      const uint32_t dfree_par = rd.par ^ 1u;
      for (int c = 0; c < n_chunks; ++c, rw.next()) {
        const int sw = rw.idx;
        mbar_wait_tc<HGSFA_TC_SLEEP_MMA>(&bars[TCB_WFULL + sw], rw.par);
        const uint32_t wbase = smem_u32(smem + op.sm_w0 + size_t(sw) * op.sm_wstage_bytes);
        const int kterms = min(op.Kpad - c * CK, CK);
        const int ksteps = F16 ? (kterms + 15) >> 4 : kterms >> 3;        // F16: an odd last 8-term group is zero-filled to 16
        for (int t = 0; t < vt; ++t, ra.next()) {
          if (c == 0) mbar_wait_tc<HGSFA_TC_SLEEP_MMA>(&bars[TCB_DFREE + set * TC_MAX_TW + t], dfree_par);
          const int sa = ra.idx;
          mbar_wait_tc<HGSFA_TC_SLEEP_MMA>(&bars[TCB_AFULL + sa], ra.par);
          tc_fence_after();
          if (leader) {
            const uint32_t d_t = tb + uint32_t((set * op.twc + t) * op.Npad16);
            const uint32_t a_hi = tb + a_col0 + uint32_t(sa) * A_STAGE;
            for (int j = 0; j < ksteps; ++j) {
              const uint64_t bhi = tc_desc(wbase + uint32_t(2 * j) * lbo, lbo, sbo);
              const uint64_t blo = tc_desc(wbase + lo_off + uint32_t(2 * j) * lbo, lbo, sbo);
              if constexpr (F16) {
                tc_mma_f16(d_t, a_hi + 8 * j, bhi, idesc, (c | j) ? 1u : 0u);
                tc_mma_f16(d_t, a_hi + 8 * j, blo, idesc, 1u);
                tc_mma_f16(d_t, a_hi + A_LO + 8 * j, bhi, idesc, 1u);
              } else {
                tc_mma(d_t, a_hi + 8 * j, bhi, idesc, (c | j) ? 1u : 0u);
                tc_mma(d_t, a_hi + 8 * j, blo, idesc, 1u);
                tc_mma(d_t, a_hi + A_LO + 8 * j, bhi, idesc, 1u);
              }
            }
            tc_commit(&bars[TCB_AFREE + sa]);
            if (c == n_chunks - 1) tc_commit(&bars[TCB_DFULL + set * TC_MAX_TW + t]);
          }
          __syncwarp();
        }
        if (leader) tc_commit(&bars[TCB_WFREE + sw]);
        __syncwarp();
      }
    }
  } else {
    // ================================ expansion / epilogue (thread = window) ================================
    const int win = tid & (TILE - 1);
    const uint32_t lane_base = tbase + (uint32_t((warp & 3) * 32) << 16);
    const Term16* terms = reinterpret_cast<const Term16*>(smem + op.sm_terms);
    const Seg* segs = reinterpret_cast<const Seg*>(smem + op.sm_segs);
    const int* chunk_seg = reinterpret_cast<const int*>(smem + op.sm_chunkseg);
    float* bias_buf = reinterpret_cast<float*>(smem + op.sm_bias);   // 8 slots of Npad16 floats

    auto epilogue = [&](int node, int set, uint32_t par, const float* bias) {
      const int nvalid = __ldg(op.n_valid + node);
      const int col0 = __ldg(op.out_col + node) + __ldg(op.col_off + node);
      for (int t = 0; t < vt; ++t) {
        mbar_wait_tc<HGSFA_TC_SLEEP_EPI>(&bars[TCB_DFULL + set * TC_MAX_TW + t], par);
        tc_fence_after();
        float* out = xout + (size_t(tile0 + t) * op.out_dim + col0) * TILE + win;
        const float clo = op.clip_lo, chi = op.clip_hi, sc = op.scale;
        for (int n0 = 0; n0 < op.Npad16; n0 += 16, out += 16 * TILE) {
          uint32_t v[16];
          tmem_ld16(lane_base + uint32_t((set * op.twc + t) * op.Npad16 + n0), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (n0 + 16 >= op.Npad16) {      // accumulator drained: the MMAs of a later node may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[TCB_DFREE + set * TC_MAX_TW + t]);
          }
          const int nleft = nvalid - n0;
          float y[16];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + n0 + 4 * q);
            y[4 * q + 0] = fminf(fmaxf(F16 ? fmaf(__uint_as_float(v[4 * q + 0]), sc, b4.x) : __uint_as_float(v[4 * q + 0]) + b4.x, clo), chi);
            y[4 * q + 1] = fminf(fmaxf(F16 ? fmaf(__uint_as_float(v[4 * q + 1]), sc, b4.y) : __uint_as_float(v[4 * q + 1]) + b4.y, clo), chi);
            y[4 * q + 2] = fminf(fmaxf(F16 ? fmaf(__uint_as_float(v[4 * q + 2]), sc, b4.z) : __uint_as_float(v[4 * q + 2]) + b4.z, clo), chi);
            y[4 * q + 3] = fminf(fmaxf(F16 ? fmaf(__uint_as_float(v[4 * q + 3]), sc, b4.w) : __uint_as_float(v[4 * q + 3]) + b4.w, clo), chi);
          }
          if (nleft >= 16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) out[j * TILE] = y[j];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < nleft) out[j * TILE] = y[j];
          }
        }
      }
    };

    if (TC_EPI && warp >= 4) {
      // ---- dedicated epilogue warps: bias slots are written by expansion warp 0 before the node's first A stage is
      // published, i.e. before any MMA of the node and therefore before its DFULL
      Ring rd(nd);
      int slot = 0;
      for (int node = node_begin; node < node_end; ++node, rd.next(), slot = (slot + 1) & 7)
        epilogue(node, rd.idx, rd.par, bias_buf + slot * op.Npad16);
    } else {
    Ring rx(nstx), ra(na), rd(nd);
#ifdef HGSFA_TC_TRACE
    const bool tracing = op.n_nodes == HGSFA_TC_TRACE && blockIdx.x == 9 && blockIdx.y == 0 && warp == 0 && lane == 0;
    int tcnt = 1;
#endif
    int prev_set = 0, slot = 0;
    uint32_t prev_par = 0u;
    for (int node = node_begin; node < node_end; ++node, rx.next(), rd.next(), slot = (slot + 1) & 7) {
      const int sx = rx.idx;
      TC_T(0);
      mbar_wait_tc(&bars[TCB_XFULL + sx], rx.par);
      TC_T(1);
      const uint8_t* stage = smem + op.sm_x0 + size_t(sx) * op.sm_xstage_bytes;
      const float* head = reinterpret_cast<const float*>(stage + size_t(op.twc) * op.sm_raw_bytes);
      const float* mean = head;
      const int d_pad = (op.d_in + 3) & ~3;
      // bias of this node, kept past the release of the stage (the epilogue runs later, possibly on other warps):
      // with epilogue warps one shared ring of 8 slots written by warp 0 (the epilogue is at most nd + na nodes
      // behind), otherwise a private copy per expansion warp (2 sets x 4 warps = the same 8 slots)
      float* bias_dst = bias_buf + (TC_EPI ? slot : (warp * 2 + rd.idx)) * op.Npad16;
      __syncwarp();
      if (!TC_EPI || warp == 0)
        for (int i = lane; i < op.Npad16; i += 32) bias_dst[i] = head[d_pad + i];
      __syncwarp();

      for (int c = 0; c < n_chunks; ++c) {
        const int sg0 = chunk_seg[c], sg1 = chunk_seg[c + 1];
        for (int t = 0; t < vt; ++t, ra.next()) {
          const int sa = ra.idx;
          TC_T(2);
          mbar_wait_tc(&bars[TCB_AFREE + sa], ra.par ^ 1u);
          tc_fence_after();
          TC_T(3);
          const IN_T* xs = reinterpret_cast<const IN_T*>(stage + size_t(t) * op.sm_raw_bytes);
          // term-unit address of the stage's first term (see tc_store8): TF32 one column per term, F16 two terms per column
          const uint32_t a_stage = F16 ? ((lane_base & 0xffff0000u) | (2u * ((lane_base & 0xffffu) + a_col0 + uint32_t(sa) * A_STAGE)))
                                       : lane_base + a_col0 + uint32_t(sa) * A_STAGE;
          for (int sgi = sg0; sgi < sg1; ++sgi) {
            const Seg sg = segs[sgi];
            const int cnt = sg.kind, ngroups = (sg.k1 - sg.k0) >> 3;     // kind = number of real terms of the piece
            const uint32_t col = a_stage + uint32_t(sg.k0 - c * CK);
            if (sg.ibase >= 0 && (sg.op == OP_ID || sg.op == OP_ABSPOW)) {
              const IN_T* xp = xs + size_t(sg.ibase) * TILE + win;
              const float* mp = mean + sg.ibase;
              if (sg.op == OP_ABSPOW) tc_seg_rows<IN_T, 2, F16>(xp, mp, cnt, ngroups, sg.p, col TC_TRACE_ARG);
              else if (sg.nomean) tc_seg_rows<IN_T, 0, F16>(xp, mp, cnt, ngroups, 0.f, col TC_TRACE_ARG);
              else tc_seg_rows<IN_T, 1, F16>(xp, mp, cnt, ngroups, 0.f, col TC_TRACE_ARG);
              continue;
            }
            if (sg.op == OP_TRI) {
              const IN_T* xr = xs + size_t(sg.ibase) * TILE + win;
              const float* mr = mean + sg.ibase;
              switch (int(sg.p)) {
#define HG_TRI(N_) case N_: tc_seg_tri<IN_T, N_, F16>(xr, mr, sg.nomean, cnt, col, op.prod_scale); break;
                // register-resident form for the sizes in TC_TRI_SIZES only: every instantiation is 0.6-1.5 k instructions, and
                // with all of N = 3..16 the kernel was 148 KB of SASS -- past the instruction cache, 4 % slower on every
                // layer (profiles/README_r02.md item 17); other sizes stay table-driven products (OP_MUL below)
                HG_TRI(10)
#undef HG_TRI
                default: break;
              }
              continue;
            }
            const Term16* tp = terms + sg.pad1;                           // pad1 = first entry of the term table
            const int2* to = reinterpret_cast<const int2*>(smem + op.sm_toff) + sg.pad1;
            const float2* tm2 = reinterpret_cast<const float2*>(head + d_pad + op.Npad16) + sg.pad1;   // (x_mean[i], x_mean[j])
            const IN_T* xt = xs + win;
#pragma unroll 1
            for (int g = 0; g < ngroups; ++g) {
              float v[8];
              if (sg.op == OP_MUL) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int kk = min(8 * g + j, cnt - 1);
                  const int2 o = to[kk];
                  const float2 m = tm2[kk];
                  if constexpr (F16)
                    v[j] = ((tc_ld<IN_T>(xt + o.x) - m.x) * op.prod_scale) * ((tc_ld<IN_T>(xt + o.y) - m.y) * op.prod_scale);
                  else
                    v[j] = (tc_ld<IN_T>(xt + o.x) - m.x) * (tc_ld<IN_T>(xt + o.y) - m.y);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int kk = min(8 * g + j, cnt - 1);
                  const Term16 tm = tp[kk];
                  const int i = sg.ibase >= 0 ? sg.ibase + kk : int(tm.i);
                  const float x = tc_ld<IN_T>(xt + i * TILE) - mean[i];
                  float r;
                  switch (sg.op) {
                    case OP_ID: r = x; break;
                    case OP_ABSPOW: r = abspow(x, sg.p); break;
                    case OP_SGNPOW: r = copysignf(abspow(x, sg.p), x); break;
                    case OP_MUL3:
                      r = x * (tc_ld<IN_T>(xt + int(tm.j) * TILE) - mean[tm.j]) * (tc_ld<IN_T>(xt + int(tm.k) * TILE) - mean[tm.k]);
                      break;
                    case OP_ABS: r = fabsf(x); break;
                    case OP_CLIP: r = fminf(fmaxf(x, -sg.p), sg.p); break;
                    default: r = 0.f; break;
                  }
                  v[j] = r;
                }
              }
              tc_store8<F16>(col + 8 * g, v);
            }
          }
          if constexpr (F16) {
            // a K step of the FP16 MMA covers 16 terms: zero the second half of an odd last group (its weight rows are
            // zero too, but stale tensor-memory bits could be NaN patterns)
            const int kterms = min(op.Kpad - c * CK, CK);
            if (kterms & 8) {
              const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              tc_store8_exact<true>(a_stage + uint32_t(kterms), z);
            }
          }
          TC_T(4);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          TC_T(5);
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars[TCB_AFULL + sa]);
          TC_T(6);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[TCB_XFREE + sx]);     // receptive field consumed
      TC_T(7);
      if (!TC_EPI) {
        if (nd == 2) {
          if (node > node_begin) epilogue(node - 1, prev_set, prev_par, bias_buf + (warp * 2 + prev_set) * op.Npad16);
          prev_set = rd.idx;
          prev_par = rd.par;
        } else {
          epilogue(node, rd.idx, rd.par, bias_buf + (warp * 2 + rd.idx) * op.Npad16);
        }
      }
    }
    if (!TC_EPI && nd == 2 && node_end > node_begin)
      epilogue(node_end - 1, prev_set, prev_par, bias_buf + (warp * 2 + prev_set) * op.Npad16);
#ifdef HGSFA_TC_TRACE
    if (tracing) tc_trace[0] = (unsigned long long)tcnt;
#endif
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(op.tmem_cols));
}

}  // namespace hgsfa

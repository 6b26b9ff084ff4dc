// Fused front of a thin fan-in-2 HiGSFA network: layers 0, 1 and 2 in ONE kernel, lane-resident (sm_100a).
//
// Replaces the first three per-layer launches of mdp.Flow.execute (FaceDetectUpdated.py:699): the 4x4-pixel
// first-layer nodes, the horizontal joins and the vertical joins form 8x8-pixel subtrees (4 + 2 + 1 nodes) that
// never interact, and a window's path through a subtree never leaves its thread:
//
//   thread = window = tensor-memory lane.  Level 0: the thread reads its 16 pixels from the staged pixel box,
//   evaluates [x | |x - m|^p], splits every value into two FP16 pieces and writes them to ITS lane of an A stage
//   (tcgen05.st).  The MMA warp contracts the tile's 128 lanes against the node's weights (tcgen05.mma kind::f16,
//   M = 128, N = 16 / 32, K = 16; three products per K step: Ahi Whi + Ahi Wlo + Alo Whi, FP32 accumulation).
//   Level 1 / 2: the SAME thread reads the two child accumulators of its lane (tcgen05.ld), applies the children's
//   bias + saturation, expands, splits and writes the next A stage.  Only the third level's output is stored to
//   HBM (window-minor tiles, coalesced).  No activation of layers 0 and 1 ever exists in shared or global memory.
//
// Why FP16 pieces and not 3xTF32: profiles/tc_probe2_r02.txt -- an M128 N16 K16 kind::f16 MMA issues every 7.4 ns,
// the M128 N16 K8 kind::tf32 one every 19.3 ns (5x per unit of K); two FP16 pieces carry 22 mantissa bits, the
// relative error of a contraction is 4e-7 (3xTF32: 5e-7).  Operands are range-checked on the host
// (pyfaceanalysis_b200/front.py: pixels, clipped activations, weights rescaled by 2^t per level).
//
// CTA = 19 warps, one per SM (all 512 tensor-memory columns):
//   warps 0-7 / 8-15 expansion of tile group 0 / 1 (two tiles of 128 windows share every weight chunk).  A group's 128
//                    tensor-memory lanes are served by TWO warps per lane quarter, which split every item between them
//                    (a node's pixels 0-7 / 8-15, child 0 / child 1 of a join, output columns 0-15 / 16-31): with 8
//                    expansion warps the kernel was bound by the dependent-issue latency of 2 warps per scheduler
//                    (profiles/ncu_r02_front_*.txt: issue slots 48 % used, top stall "wait"), tensor memory has no room
//                    for a third tile group, so the parallelism comes from inside the tile
//   warps 16 / 17    MMA issue for group 0 / 1 (one elected lane)
//   warp 18          producer: pixel boxes (3-D tensor map over row-major windows, or bulk copies of window-minor
//                    tiles) and weight chunks (cp.async.bulk) through mbarrier rings
// Item order per subtree s (software-pipelined so that no item depends on the one issued just before it):
//   L0ab(s) L2(s-1) L1ab(s) L0cd(s) STORE(s-1) L1cd(s)
// An A stage holds TWO 32-term chunks (two layer-0 nodes, or two chunks of a join) per hand-over; a pair of layer-0 nodes is
// drained by its join before the next pair is issued, so two accumulator slots serve the four nodes.
// Tensor-memory columns per group: 2 x 16 (level-0 accumulators) + 2 x 32 (level 1) + 32 (level 2) + 2 x 64 (A ring).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include "front_api.h"
#include "layer_tc.cuh"

namespace hgsfa {

constexpr int FR_NW = 6;                 // weight-ring stages
constexpr int FR_NA = 2;                 // A-ring stages per group: 64 columns = TWO 32-term chunks (hi: 16 + 16 columns, then lo:
                                         // 16 + 16) per hand-over -- a hand-over costs ~700 cycles (profiles/README_r02.md item 14)
constexpr int FR_NX = 2;                 // pixel-box stages per group
constexpr int FR_WSTAGE = FR_HEAD + 32 * 128;
constexpr int FR_XSTAGE = 16 * 8 * TILE;  // 16 x 8 pixel box of 128 windows
constexpr int FR_EXP_WARPS = 16;         // two tile groups x two warps per lane quarter
constexpr int FR_THREADS = (FR_EXP_WARPS + 3) * 32;
#ifndef HGSFA_FR_SLEEP_MMA
#define HGSFA_FR_SLEEP_MMA 64
#endif
#ifndef HGSFA_FR_SLEEP_PROD
#define HGSFA_FR_SLEEP_PROD 256
#endif
#ifndef HGSFA_FR_LAZY_PUBLISH
#define HGSFA_FR_LAZY_PUBLISH 0
#endif
constexpr bool FR_LAZY_PUBLISH = HGSFA_FR_LAZY_PUBLISH != 0;
constexpr int FR_SLEEP_MMA = HGSFA_FR_SLEEP_MMA, FR_SLEEP_PROD = HGSFA_FR_SLEEP_PROD;   // ns between barrier tries of the service warps
constexpr int FR_GCOLS = 256;            // tensor-memory columns per group
// accumulators: layer 0 two slots of 16 columns (a pair of nodes is drained by its join before the next pair is issued),
// layer 1 two of 32, layer 2 one of 32; then the A ring
constexpr int FR_COL_L0 = 0, FR_COL_L1 = 32, FR_COL_L2 = 96, FR_COL_A = 128, FR_ASTAGE = 64, FR_ALO = 32;
// barriers: WFULL[6] WFREE[6] then per group AFULL[3] AFREE[3] DFULL[7] XFULL[2] XFREE[2]
enum { FRB_WFULL = 0, FRB_WFREE = FR_NW, FRB_G = 2 * FR_NW, FRB_AFULL = 0, FRB_AFREE = FR_NA, FRB_DFULL = 2 * FR_NA,
       FRB_XFULL = 2 * FR_NA + 7, FRB_XFREE = 2 * FR_NA + 7 + FR_NX, FRB_GSTRIDE = 2 * FR_NA + 7 + 2 * FR_NX,
       FRB_COUNT = 2 * FR_NW + 2 * FRB_GSTRIDE };
constexpr int FR_SM_BARS = 0, FR_SM_TMEM = 512, FR_SM_BIAS = 1024;               // bias: 16 warps x 32 floats
constexpr int FR_SM_W = 1024 + 2048, FR_SM_X = FR_SM_W + FR_NW * FR_WSTAGE + 512;        // pixel stages are 1024-aligned (below)
constexpr int FR_SMEM = ((FR_SM_X + 1023) & ~1023) + 2 * FR_NX * FR_XSTAGE + 1024;

__device__ __forceinline__ bool fr_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
// barrier operations on 32-bit shared-window addresses (computed once per role: no generic-to-shared conversion per call)
__device__ __forceinline__ void fr_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void fr_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fr_commit(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void fr_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ void fr_wait(uint32_t addr, uint32_t parity) {
  // try_wait is meant to suspend the warp in hardware (alone it does, ~4 us per call: profiles/tc_probe2_r02.txt) but inside
  // this kernel it returns within tens of ns (profiles/ncu_r02_front_lines.txt: 17 % of the executed instructions were
  // re-tries), so the service warps (MMA issue, producer) sleep between tries instead of stealing issue slots from the
  // expansion warps of their scheduler.  A protocol error traps instead of hanging the GPU.
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    if (fr_try(addr, parity)) return;
    if (SLEEP_NS > 0) {
      __nanosleep(SLEEP_NS);
    } else {
      if (fr_try(addr, parity)) return;
      if (fr_try(addr, parity)) return;
      if (fr_try(addr, parity)) return;
    }
  }
  __trap();
}
template <int SLEEP_NS = 0>
__device__ __forceinline__ void fr_wait(uint64_t* bar, uint32_t parity) { fr_wait<SLEEP_NS>(smem_u32(bar), parity); }
__device__ __forceinline__ uint32_t fr_idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void fr_mma(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
template <int NC>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t* v) {      // NC columns, multiple of 8
#pragma unroll
  for (int c = 0; c < NC; c += 8) tmem_ld8(taddr + c, v + c);
}
__device__ __forceinline__ void fr_tensor_load(void* dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}

// ---- expansion of a join node (levels 1 and 2): a child's accumulator -> this warp's half of every A chunk ----
// Two warps share a lane quarter: warp `half` expands child `half` of the node.  Its 2 NP terms (identity of the child's
// NP columns, then |x - m|^p of the same columns) fill 16 of the 32 terms of each of the NP / 8 chunks, so neither warp
// ever touches the other child's accumulator (pyfaceanalysis_b200/front.py orders the weight rows accordingly).
template <int NP>
struct FrJoin {
  template <typename PUB>
  static __device__ __forceinline__ void run(uint32_t acc, const float* bias, const float* mean, int half, float s, float clo,
                                             float chi, float p, PUB&& publish) {
    float y[NP];
    {
      uint32_t r[NP];
      tmem_ld_cols<NP>(acc, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < NP; k += 4) {
        const float4 b = *reinterpret_cast<const float4*>(bias + k);
        y[k + 0] = fminf(fmaxf(fmaf(__uint_as_float(r[k + 0]), s, b.x), clo), chi);
        y[k + 1] = fminf(fmaxf(fmaf(__uint_as_float(r[k + 1]), s, b.y), clo), chi);
        y[k + 2] = fminf(fmaxf(fmaf(__uint_as_float(r[k + 2]), s, b.z), clo), chi);
        y[k + 3] = fminf(fmaxf(fmaf(__uint_as_float(r[k + 3]), s, b.w), clo), chi);
      }
    }
    constexpr int NCH = NP / 8;
    uint32_t stage = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int q = 0; q < 8; q += 2) {           // four terms per step: one LDS.128 of means serves the power terms
        const int t = 16 * c + 2 * q;            // position in this warp's term sequence, multiple of 4
        float v[4];
        if (t < NP) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = y[t + e];
        } else {
          const float4 m = *reinterpret_cast<const float4*>(mean + (t - NP));
          v[0] = abspow(y[t - NP + 0] - m.x, p);
          v[1] = abspow(y[t - NP + 1] - m.y, p);
          v[2] = abspow(y[t - NP + 2] - m.z, p);
          v[3] = abspow(y[t - NP + 3] - m.w, p);
        }
        fr_split(v[0], v[1], hi[q], lo[q]);
        fr_split(v[2], v[3], hi[q + 1], lo[q + 1]);
      }
      if ((c & 1) == 0) stage = publish.acquire();          // a stage holds two chunks: hi columns 16 (c & 1) + ..., lo after 32
      const uint32_t col = stage + uint32_t(16 * (c & 1) + 8 * half);
      tmem_st8(col, hi);
      tmem_st8(col + FR_ALO, lo);
      if ((c & 1) == 1 || c == NCH - 1) publish.release();
    }
  }
};

template <int NP1, int NP2, int IN_MODE>
__global__ void __launch_bounds__(FR_THREADS, 1)
    front_kernel(const FrontDev fd, const __grid_constant__ CUtensorMap tmap, const uint8_t* __restrict__ xin,
                 float* __restrict__ xout, int64_t ntiles) {
  extern __shared__ __align__(128) uint8_t smem[];   // the pixel ring is aligned by hand below
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FR_SM_BARS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FR_SM_TMEM);
  uint8_t* wring = smem + FR_SM_W;
  uint8_t* xring = smem + (((smem_u32(smem) + FR_SM_X + 1023u) & ~1023u) - smem_u32(smem));    // 1024-aligned
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s_begin = blockIdx.y * fd.sub_per_cta;
  const int s_end = min(fd.n_sub, s_begin + fd.sub_per_cta);
  // the two tile groups of the CTA; an odd tile count lets the last CTA compute its only tile twice (identical stores)
  const int64_t tile_g0 = min(int64_t(blockIdx.x) * 2, ntiles - 1), tile_g1 = min(int64_t(blockIdx.x) * 2 + 1, ntiles - 1);
  constexpr int NCH1 = NP1 / 8, NCH2 = NP2 / 8;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < FR_NW; ++i) { mbar_init(&bars[FRB_WFULL + i], 1); mbar_init(&bars[FRB_WFREE + i], 2); }
    for (int g = 0; g < 2; ++g) {
      uint64_t* gb = bars + FRB_G + g * FRB_GSTRIDE;
      for (int i = 0; i < FR_NA; ++i) { mbar_init(&gb[FRB_AFULL + i], 8); mbar_init(&gb[FRB_AFREE + i], 1); }
      for (int i = 0; i < 7; ++i) mbar_init(&gb[FRB_DFULL + i], 1);
      for (int i = 0; i < FR_NX; ++i) { mbar_init(&gb[FRB_XFULL + i], 1); mbar_init(&gb[FRB_XFREE + i], 8); }
    }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp == FR_EXP_WARPS + 2) {
    // ================================ producer ================================
    if (lane == 0) {
      Ring rw(FR_NW), rx(FR_NX);
      auto load_pixels = [&](int pair) {
        const int2 xy = fd.pair_xy[pair];
        for (int g = 0; g < 2; ++g) {
          uint64_t* gb = bars + FRB_G + g * FRB_GSTRIDE;
          fr_wait<FR_SLEEP_PROD>(&gb[FRB_XFREE + rx.idx], rx.par ^ 1u);
          uint8_t* dst = xring + size_t(g * FR_NX + rx.idx) * FR_XSTAGE;
          mbar_expect_tx(&gb[FRB_XFULL + rx.idx], FR_XSTAGE);
          if (IN_MODE == FR_IN_ROWMAJOR) {
            fr_tensor_load(dst, &tmap, xy.x, int((g ? tile_g1 : tile_g0) * TILE), xy.y, &gb[FRB_XFULL + rx.idx]);
          } else {
            const uint8_t* src = xin + (size_t(g ? tile_g1 : tile_g0) * fd.in_dim + size_t(xy.y) * fd.img_w + xy.x) * TILE;
            for (int r = 0; r < 8; ++r)
              bulk_g2s(dst + r * 16 * TILE, src + size_t(r) * fd.img_w * TILE, 16 * TILE, &gb[FRB_XFULL + rx.idx]);
          }
        }
        rx.next();
      };
      auto load_chunks = [&](int sub, int off, int n, int bytes) {
        const uint8_t* src = fd.wimg + size_t(sub) * fd.sub_bytes + off;
#pragma unroll 1
        for (int c = 0; c < n; ++c, rw.next()) {
          fr_wait<FR_SLEEP_PROD>(&bars[FRB_WFREE + rw.idx], rw.par ^ 1u);
          mbar_expect_tx(&bars[FRB_WFULL + rw.idx], uint32_t(bytes));
          bulk_g2s(wring + size_t(rw.idx) * FR_WSTAGE, src + size_t(c) * bytes, uint32_t(bytes), &bars[FRB_WFULL + rw.idx]);
        }
      };
      const int cb0 = FR_HEAD + fd.nn0 * 128, cb1 = FR_HEAD + fd.nn1 * 128, cb2 = FR_HEAD + fd.nn2 * 128;
      const int off_l1 = 4 * cb0, off_l2 = off_l1 + 2 * NCH1 * cb1;
      load_pixels(s_begin / 2);
      // every role walks the same 6 item slots per subtree iteration, each body instantiated once (code size: the kernel
      // must stay inside the instruction cache):  0 L0ab  1 L2(s-1)  2 L1ab  3 L0cd  4 STORE(s-1)  5 L1cd
      // (a pair of layer-0 nodes shares one A stage; its join runs before the next pair so that two accumulator slots do)
      for (int s = s_begin; s <= s_end; ++s) {
        const bool cur = s < s_end, prev = s > s_begin;
        if (cur && !(s & 1) && s + 2 < s_end) load_pixels(s / 2 + 1);
#pragma unroll 1
        for (int it = 0; it < 6; ++it) {
          if (it == 4) continue;
          const bool l0 = it == 0 || it == 3, join1 = it == 2 || it == 5;
          const bool valid = it == 1 ? prev : cur;
          const int sub = it == 1 ? s - 1 : s;
          const int off = it == 1 ? off_l2 : (join1 ? off_l1 + (it == 5 ? NCH1 * cb1 : 0) : (it == 3 ? 2 * cb0 : 0));
          const int n = it == 1 ? NCH2 : (join1 ? NCH1 : 2);
          const int bytes = it == 1 ? cb2 : (join1 ? cb1 : cb0);
          (void)l0;
          if (valid) load_chunks(sub, off, n, bytes);
        }
      }
    }
    __syncwarp();
  } else if (warp >= FR_EXP_WARPS) {
    // ================================ MMA issue (one warp per tile group) ================================
    const int g = warp - FR_EXP_WARPS;
    const uint32_t gbu = smem_u32(bars + FRB_G + g * FRB_GSTRIDE), barsu = smem_u32(bars);
    const bool leader = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0) + uint32_t(g * FR_GCOLS);
    Ring rw(FR_NW), ra(FR_NA);
    // One item = nch 32-term chunks, two per A stage.  pair = true (layer 0): every chunk is a node of its own -- fresh
    // accumulator dcol + 16 h, barrier dslot + h, its weight stage (with the node's head) released with its MMAs.
    // pair = false (joins): all chunks accumulate into dcol; the first weight stage carries the head the expansion warps
    // read for every chunk, so it is released last.
    auto mma_item = [&](int nch, int nn, uint32_t dcol, int dslot, bool pair) {
      const uint32_t idesc = fr_idesc(nn);
      const uint32_t lbo = uint32_t(nn / 8) * 128u, sbo = 128u, lo_off = uint32_t(nn) * 64u;
      const int ws0 = rw.idx;
#pragma unroll 1
      for (int c = 0; c < nch; ++c, rw.next()) {
        const int h = c & 1;
        const bool last_of_stage = h == 1 || c == nch - 1;
        fr_wait<FR_SLEEP_MMA>(barsu + 8u * uint32_t(FRB_WFULL + rw.idx), rw.par);
        if (h == 0) fr_wait<FR_SLEEP_MMA>(gbu + 8u * uint32_t(FRB_AFULL + ra.idx), ra.par);
        tc_fence_after();
        if (leader) {
          const uint32_t wbase = smem_u32(wring + size_t(rw.idx) * FR_WSTAGE + FR_HEAD);
          const uint32_t a_hi = tb + uint32_t(FR_COL_A + FR_ASTAGE * ra.idx + 16 * h), a_lo = a_hi + uint32_t(FR_ALO);
          const uint32_t d = tb + dcol + (pair ? 16u * uint32_t(h) : 0u);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t bhi = tc_desc(wbase + uint32_t(2 * j) * lbo, lbo, sbo);
            const uint64_t blo = tc_desc(wbase + lo_off + uint32_t(2 * j) * lbo, lbo, sbo);
            fr_mma(d, a_hi + 8u * j, bhi, idesc, ((pair ? 0 : c) | j) ? 1u : 0u);
            fr_mma(d, a_hi + 8u * j, blo, idesc, 1u);
            if (!(pair && j == 0)) fr_mma(d, a_lo + 8u * j, bhi, idesc, 1u);   // pixels are exact in FP16
          }
          if (last_of_stage) fr_commit(gbu + 8u * uint32_t(FRB_AFREE + ra.idx));
          if (pair) {
            fr_commit(barsu + 8u * uint32_t(FRB_WFREE + rw.idx));
            fr_commit(gbu + 8u * uint32_t(FRB_DFULL + dslot + h));
          } else {
            if (c > 0 || nch == 1) fr_commit(barsu + 8u * uint32_t(FRB_WFREE + rw.idx));
            if (c == nch - 1) {
              if (nch > 1) fr_commit(barsu + 8u * uint32_t(FRB_WFREE + ws0));
              fr_commit(gbu + 8u * uint32_t(FRB_DFULL + dslot));
            }
          }
        }
        __syncwarp();
        if (last_of_stage) ra.next();
      }
    };
    for (int s = s_begin; s <= s_end; ++s) {
      const bool cur = s < s_end, prev = s > s_begin;
#pragma unroll 1
      for (int it = 0; it < 6; ++it) {
        if (it == 4) continue;
        const bool l0 = it == 0 || it == 3, join1 = it == 2 || it == 5;
        const bool valid = it == 1 ? prev : cur;
        const int nch = it == 1 ? NCH2 : (join1 ? NCH1 : 2);
        const int nn = it == 1 ? fd.nn2 : (join1 ? fd.nn1 : fd.nn0);
        const uint32_t dcol = it == 1 ? FR_COL_L2 : (join1 ? FR_COL_L1 + (it == 5 ? 32 : 0) : FR_COL_L0);
        const int dslot = it == 1 ? 6 : (join1 ? (it == 5 ? 5 : 4) : (it == 3 ? 2 : 0));
        if (valid) mma_item(nch, nn, dcol, dslot, l0);
      }
    }
  } else {
    // ================================ expansion (thread = window) ================================
    // warps 0-7 tile group 0, 8-15 group 1; inside a group warps 0-3 are "half 0" of the four lane quarters, 4-7 "half 1":
    // the two halves of a quarter split every item between them (pixels 0-7 / 8-15, child 0 / child 1, columns 0-15 / 16-31)
    const int g = warp >> 3, half = (warp >> 2) & 1;
    const int win = (warp & 3) * 32 + lane;
    const uint32_t gbu = smem_u32(bars + FRB_G + g * FRB_GSTRIDE), barsu = smem_u32(bars);
    const uint32_t lane_base = tbase + (uint32_t((warp & 3) * 32) << 16) + uint32_t(g * FR_GCOLS);
    float* bias_out = reinterpret_cast<float*>(smem + FR_SM_BIAS) + warp * 32;
    const int64_t tile = g ? tile_g1 : tile_g0;
    Ring rw(FR_NW), ra(FR_NA), rx(FR_NX);

    // Hand-over of A stages to the MMA warp.  With FR_LAZY_PUBLISH a chunk is handed over one chunk late (release() only
    // notes that its stores were issued; the next acquire() or flush() waits for them and arrives), so that the stores
    // complete behind the next chunk's arithmetic -- measured 3 % SLOWER (profiles/README_r02.md), hence off.
    struct Publisher {
      uint32_t gbu;
      Ring& ra;
      uint32_t lane_base;
      int lane;
      int pending;
      __device__ __forceinline__ void flush() {
        if (pending >= 0) {
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          __syncwarp();
          if (lane == 0) fr_arrive(gbu + 8u * uint32_t(FRB_AFULL + pending));
          pending = -1;
        }
      }
      __device__ __forceinline__ uint32_t acquire() {
        flush();
        fr_wait(gbu + 8u * uint32_t(FRB_AFREE + ra.idx), ra.par ^ 1u);
        tc_fence_after();
        return lane_base + uint32_t(FR_COL_A + FR_ASTAGE * ra.idx);
      }
      __device__ __forceinline__ void release() {
        pending = ra.idx;
        ra.next();
        if (!FR_LAZY_PUBLISH) flush();
      }
    } pub{gbu, ra, lane_base, lane, -1};

    auto head_of = [&](int stage) { return reinterpret_cast<const float*>(wring + size_t(stage) * FR_WSTAGE); };

    auto item_l0 = [&](int sub, int pair) {                    // nodes 2 pair, 2 pair + 1 of the subtree: one A stage, one hand-over
      const int q = sub & 1;
      if (pair == 0 && q == 0) fr_wait(gbu + 8u * uint32_t(FRB_XFULL + rx.idx), rx.par);
      const uint8_t* box = xring + size_t(g * FR_NX + rx.idx) * FR_XSTAGE;
      uint32_t stage = 0;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int off = __ldg(fd.l0_off + sub * 4 + 2 * pair + h);
        const int dy = off & 0xff, dx = off >> 8;
        float px[8];                                           // rows 2 half, 2 half + 1 of the node's 4 x 4 pixels
        if (IN_MODE == FR_IN_ROWMAJOR) {
          // box = [8 rows][128 windows][16 bytes] (tensor dimensions ordered x, window, y): consecutive lanes read words
          // 16 bytes apart -- 4 wavefronts per load, as many as the byte loads of the window-minor form need
          const uint8_t* wrow = box + win * 16 + dx;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(wrow + (dy + 2 * half + r) * (16 * TILE));
            const float4 f = u8x4_to_float4(u);
            px[4 * r + 0] = f.x; px[4 * r + 1] = f.y; px[4 * r + 2] = f.z; px[4 * r + 3] = f.w;
          }
        } else {
          // box = [8 rows][16 pixels][128 windows]
          const uint8_t* col = box + ((dy + 2 * half) * 16 + dx) * TILE + win;
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int x = 0; x < 4; ++x)
              px[4 * r + x] = __uint_as_float(0x4B000000u | uint32_t(col[(r * 16 + x) * TILE])) - 8388608.0f;
        }
        fr_wait(barsu + 8u * uint32_t(FRB_WFULL + rw.idx), rw.par);
        const float* mean = head_of(rw.idx) + 8 * half;
        rw.next();
        uint32_t hi_id[4], hi_pw[4], lo_pw[4];
#pragma unroll
        for (int k = 0; k < 8; k += 4) {
          const float4 m = *reinterpret_cast<const float4*>(mean + k);
          hi_id[k / 2] = fr_pack(px[k], px[k + 1]);            // identity terms: 8-bit integers, exact in FP16
          hi_id[k / 2 + 1] = fr_pack(px[k + 2], px[k + 3]);
          fr_split(abspow(px[k] - m.x, fd.p0), abspow(px[k + 1] - m.y, fd.p0), hi_pw[k / 2], lo_pw[k / 2]);
          fr_split(abspow(px[k + 2] - m.z, fd.p0), abspow(px[k + 3] - m.w, fd.p0), hi_pw[k / 2 + 1], lo_pw[k / 2 + 1]);
        }
        // the node's chunk of the stage: hi columns 16 h + 0-7 identity (K step 0), + 8-15 power (K step 1); lo pieces
        // FR_ALO columns further (the identity's are zero and never read)
        if (h == 0) stage = pub.acquire();
        const uint32_t col = stage + uint32_t(16 * h + 4 * half);
        tmem_st4(col, hi_id);
        tmem_st4(col + 8, hi_pw);
        tmem_st4(col + FR_ALO + 8, lo_pw);
      }
      pub.release();
      if (pair == 1 && q == 1) {                               // pixel box of the pair of subtrees consumed
        __syncwarp();
        if (lane == 0) fr_arrive(gbu + 8u * uint32_t(FRB_XFREE + rx.idx));
        rx.next();
      }
    };
    auto item_l1 = [&](int h, uint32_t par) {
      pub.flush();
      fr_wait(gbu + 8u * uint32_t(FRB_DFULL + 2 * h + half), par);            // this warp expands child `half` of the join
      tc_fence_after();
      fr_wait(barsu + 8u * uint32_t(FRB_WFULL + rw.idx), rw.par);
      const float* head = head_of(rw.idx);
      rw.advance(NCH1);
      FrJoin<NP1>::run(lane_base + FR_COL_L0 + 16 * half, head + NP1 * half, head + 2 * NP1 + NP1 * half, half, fd.s0,
                       fd.clo0, fd.chi0, fd.p1, pub);
    };
    auto item_l2 = [&](uint32_t par) {
      pub.flush();
      fr_wait(gbu + 8u * uint32_t(FRB_DFULL + 4 + half), par);
      tc_fence_after();
      fr_wait(barsu + 8u * uint32_t(FRB_WFULL + rw.idx), rw.par);
      const float* head = head_of(rw.idx);
      rw.advance(NCH2);
      __syncwarp();
      bias_out[lane] = head[4 * NP2 + lane];                   // kept for the store item (the weight stage is recycled)
      __syncwarp();
      FrJoin<NP2>::run(lane_base + FR_COL_L1 + 32 * half, head + NP2 * half, head + 2 * NP2 + NP2 * half, half, fd.s1, fd.clo1,
                       fd.chi1, fd.p2, pub);
    };
    auto item_store = [&](int sub, uint32_t par) {
      pub.flush();
      fr_wait(gbu + 8u * uint32_t(FRB_DFULL + 6), par);
      tc_fence_after();
      uint32_t r[16];                                          // this warp's half of the node's 32 columns
      tmem_ld_cols<16>(lane_base + FR_COL_L2 + 16 * half, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float* out = xout + (size_t(tile) * fd.out_dim + __ldg(fd.out_col + sub) + 16 * half) * TILE + win;
      const int nv = fd.nv_out - 16 * half;
      const float* bias = bias_out + 16 * half;
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        if (k < nv) {                                          // warp-uniform: one branch per four columns
          const float4 b = *reinterpret_cast<const float4*>(bias + k);
          const float y0 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 0]), fd.s2, b.x), fd.clo2), fd.chi2);
          const float y1 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 1]), fd.s2, b.y), fd.clo2), fd.chi2);
          const float y2 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 2]), fd.s2, b.z), fd.clo2), fd.chi2);
          const float y3 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 3]), fd.s2, b.w), fd.clo2), fd.chi2);
          out[size_t(k) * TILE] = y0;
          if (k + 1 < nv) out[size_t(k + 1) * TILE] = y1;
          if (k + 2 < nv) out[size_t(k + 2) * TILE] = y2;
          if (k + 3 < nv) out[size_t(k + 3) * TILE] = y3;
        }
      }
    };

    for (int s = s_begin; s <= s_end; ++s) {
      const bool cur = s < s_end, prev = s > s_begin;
      const uint32_t par = uint32_t(s - s_begin) & 1u, ppar = par ^ 1u;
#pragma unroll 1
      for (int it = 0; it < 6; ++it) {
        if (it == 1) {
          if (prev) item_l2(ppar);
        } else if (it == 4) {
          if (prev) item_store(s - 1, ppar);
        } else if (cur) {
          if (it == 2 || it == 5) item_l1(it == 5 ? 1 : 0, par);
          else item_l0(s, it == 3 ? 1 : 0);
        }
      }
    }
    pub.flush();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

}  // namespace hgsfa

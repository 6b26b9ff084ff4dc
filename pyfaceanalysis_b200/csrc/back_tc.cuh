// Single-layer HiGSFA kernel in the style of the fused front (front_tc.cuh): FP16-split tcgen05, thread = window, the
// expansion warps drain their own accumulators (sm_100a).
//
// Same operation as layer_tc.cuh -- Y[tile, node] = terms(X[tile, gather[node]] - x_mean[node]) @ W[node] + b[node] for a
// single-pass layer op -- for ops whose term table is made of the segment kinds the thin networks use: identity rows,
// |x - m|^p rows and the upper-triangular products of the first 10 centred inputs ("QT" of cuicuilco's s10 selectors).
// What changed against layer_tc.cuh and why (profiles/README_r02.md):
//   * tcgen05.mma kind::f16 on 2-piece FP16 operands (K = 16 per instruction) instead of 3xTF32 (K = 8): per unit of K
//     the tensor pipe is 3-5x cheaper, an A chunk of 32 terms is 32 tensor-memory columns instead of 64, a weight chunk
//     half the bytes.  Operands are range-checked when the plan is created (inputs bounded by the previous layer's
//     saturation; products are evaluated on inputs pre-scaled by 2^-s with the weights rescaled by 2^2s).
//   * no interpreter: an A chunk is four 8-term groups whose kind / first row / count sit in a 16-byte descriptor; the
//     group code is straight-line, means are fetched four at a time, the triangular products are unrolled at compile time.
//   * one CTA per SM with two tile groups sharing every weight chunk, TWO warps per tensor-memory lane quarter and group
//     (16 expansion warps: a pair splits the 8-term groups of every chunk and the output columns, as in the front --
//     two in-order warps per scheduler could not hide their own latencies); the expansion warps store the previous
//     node's outputs themselves (no idle epilogue warps), one node behind so that they never wait for an MMA.
// Roles: warps 0-7 / 8-15 expansion + stores of tile group 0 / 1 (quarter = warp & 3, half = (warp >> 2) & 1), warps
// 16 / 17 MMA issue, warp 18 producer (receptive-field runs and weight chunks by cp.async.bulk).  Tensor memory per
// group: 2 x 64 accumulator columns + 4 x 32 A columns.
#pragma once
#include "front_tc.cuh"

namespace hgsfa {

// upper-triangular products of 8 consecutive terms of group G (row-major enumeration i <= j over N inputs)
template <int N, int G>
__device__ __forceinline__ void bk_tri_group(const float (&xc)[N], int g, float (&v)[8]) {
  constexpr int T = N * (N + 1) / 2;
  if constexpr (8 * G < T) {
    if (g == G) {
#define HG_BK_TERM(J)                                                                        \
  {                                                                                          \
    constexpr int idx = 8 * G + J;                                                           \
    if constexpr (idx < T) v[J] = xc[tri_row(N, idx)] * xc[tri_col(N, idx)];                 \
    else v[J] = 0.f;                                                                         \
  }
      HG_BK_TERM(0) HG_BK_TERM(1) HG_BK_TERM(2) HG_BK_TERM(3) HG_BK_TERM(4) HG_BK_TERM(5) HG_BK_TERM(6) HG_BK_TERM(7)
#undef HG_BK_TERM
      return;
    }
    bk_tri_group<N, G + 1>(xc, g, v);
  }
}

__global__ void __launch_bounds__(BK_THREADS, 1)
    back_kernel(const BackDev bd, const float* __restrict__ xin, float* __restrict__ xout, int64_t ntiles, int smem_x0) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
  BkGroup* groups_s = reinterpret_cast<BkGroup*>(smem + 1024);              // up to 64 chunks x 4 groups
  float* head_s = reinterpret_cast<float*>(smem + BK_SM_HEAD);              // [16 warps][2 x 64 bias | 192 mean]
  uint8_t* wring = smem + BK_SM_W;
  uint8_t* xring = smem + smem_x0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int node_begin = blockIdx.y * bd.npc, node_end = min(bd.n_nodes, node_begin + bd.npc);
  const int64_t tile_g0 = min(int64_t(blockIdx.x) * 2, ntiles - 1), tile_g1 = min(int64_t(blockIdx.x) * 2 + 1, ntiles - 1);
  const int n_chunks = bd.n_chunks;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 32) {
    for (int i = 0; i < BK_NW; ++i) { mbar_init(&bars[BKB_WFULL + i], 1); mbar_init(&bars[BKB_WFREE + i], 2); }
    for (int g = 0; g < 2; ++g) {
      uint64_t* gb = bars + BKB_G + g * BKB_GSTRIDE;
      for (int i = 0; i < BK_NA; ++i) { mbar_init(&gb[BKB_AFULL + i], BK_EXP_WARPS / 2); mbar_init(&gb[BKB_AFREE + i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&gb[BKB_DFULL + i], 1); mbar_init(&gb[BKB_XFULL + i], 1); mbar_init(&gb[BKB_XFREE + i], BK_EXP_WARPS / 2); }
    }
    mbar_fence_init();
  }
  for (int i = tid; i < n_chunks * 4; i += BK_THREADS) groups_s[i] = bd.groups[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;

  if (warp == BK_EXP_WARPS + 2) {
    // ================================ producer ================================
    if (lane == 0) {
      Ring rw(BK_NW), rx(bd.nstx);
      for (int node = node_begin; node < node_end; ++node) {
        const Run* runs = bd.runs + size_t(node) * bd.n_runs;
        for (int g = 0; g < 2; ++g) {
          uint64_t* gb = bars + BKB_G + g * BKB_GSTRIDE;
          fr_wait<FR_SLEEP_PROD>(&gb[BKB_XFREE + rx.idx], rx.par ^ 1u);
          uint8_t* dst = xring + size_t(g * bd.nstx + rx.idx) * bd.x_stage_bytes;
          uint32_t bytes = 0;
          for (int r = 0; r < bd.n_runs; ++r) bytes += uint32_t(runs[r].len) * TILE * 4u;
          mbar_expect_tx(&gb[BKB_XFULL + rx.idx], bytes);
          const float* src = xin + size_t(g ? tile_g1 : tile_g0) * bd.in_dim * TILE;
          for (int r = 0; r < bd.n_runs; ++r)
            if (runs[r].len > 0)
              bulk_g2s(dst + size_t(runs[r].i0) * TILE * 4, src + size_t(runs[r].f0) * TILE, uint32_t(runs[r].len) * TILE * 4u,
                       &gb[BKB_XFULL + rx.idx]);
        }
        rx.next();
        const uint8_t* wsrc = bd.wimg + size_t(bd.shared ? 0 : node) * n_chunks * bd.chunk_bytes;
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c, rw.next()) {
          fr_wait<FR_SLEEP_PROD>(&bars[BKB_WFREE + rw.idx], rw.par ^ 1u);
          mbar_expect_tx(&bars[BKB_WFULL + rw.idx], uint32_t(bd.chunk_bytes));
          bulk_g2s(wring + size_t(rw.idx) * BK_WSTAGE, wsrc + size_t(c) * bd.chunk_bytes, uint32_t(bd.chunk_bytes), &bars[BKB_WFULL + rw.idx]);
        }
      }
    }
    __syncwarp();
  } else if (warp >= BK_EXP_WARPS) {
    // ================================ MMA issue (one warp per tile group) ================================
    const int g = warp - BK_EXP_WARPS;
    uint64_t* gb = bars + BKB_G + g * BKB_GSTRIDE;
    const bool leader = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0) + uint32_t(g * BK_GCOLS);
    const uint32_t idesc = fr_idesc(bd.nn);
    const uint32_t lbo = uint32_t(bd.nn / 8) * 128u, sbo = 128u, lo_off = uint32_t(bd.nn) * 64u;
    Ring rw(BK_NW), ra(BK_NA);
    for (int node = node_begin; node < node_end; ++node) {
      const int slot = (node - node_begin) & 1;
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c, rw.next(), ra.next()) {
        fr_wait<FR_SLEEP_MMA>(&bars[BKB_WFULL + rw.idx], rw.par);
        fr_wait<FR_SLEEP_MMA>(&gb[BKB_AFULL + ra.idx], ra.par);
        tc_fence_after();
        if (leader) {
          const uint32_t wbase = smem_u32(wring + size_t(rw.idx) * BK_WSTAGE + FR_HEAD);
          const uint32_t a_hi = tb + uint32_t(BK_COL_A + 32 * ra.idx), a_lo = a_hi + 16u;
          const uint32_t d_t = tb + uint32_t(BK_COL_ACC + 64 * slot);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint64_t bhi = tc_desc(wbase + uint32_t(2 * j) * lbo, lbo, sbo);
            const uint64_t blo = tc_desc(wbase + lo_off + uint32_t(2 * j) * lbo, lbo, sbo);
            fr_mma(d_t, a_hi + 8u * j, bhi, idesc, (c | j) ? 1u : 0u);
            fr_mma(d_t, a_hi + 8u * j, blo, idesc, 1u);
            fr_mma(d_t, a_lo + 8u * j, bhi, idesc, 1u);
          }
          tc_commit(&gb[BKB_AFREE + ra.idx]);
          tc_commit(&bars[BKB_WFREE + rw.idx]);
          if (c == n_chunks - 1) tc_commit(&gb[BKB_DFULL + slot]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ expansion + stores (thread = window) ================================
    const int g = warp >> 3, half = (warp >> 2) & 1;
    const int win = tid & (TILE - 1);
    uint64_t* gb = bars + BKB_G + g * BKB_GSTRIDE;
    const uint32_t lane_base = tbase + (uint32_t((warp & 3) * 32) << 16) + uint32_t(g * BK_GCOLS);
    float* bias_w = head_s + warp * BK_HEAD_WARP;              // two slots of 64 floats, by node parity
    float* mean = bias_w + 128;                                // this warp's copy of the node's means
    const int64_t tile = g ? tile_g1 : tile_g0;
    const int d_in = bd.d_in;
    Ring rw(BK_NW), ra(BK_NA), rx(bd.nstx);

    auto store_node = [&](int node, int slot, uint32_t par) {
      fr_wait(&gb[BKB_DFULL + slot], par);
      tc_fence_after();
      const int nv = __ldg(bd.n_valid + node);
      float* out = xout + (size_t(tile) * bd.out_dim + __ldg(bd.out_col + node)) * TILE + win;
      const float* bias = bias_w + slot * 64;
#pragma unroll 1
      for (int n0 = 8 * half; n0 < bd.nn; n0 += 16) {          // this warp's 8 of every 16 columns
        uint32_t r[8];
        tmem_ld8(lane_base + uint32_t(BK_COL_ACC + 64 * slot + n0), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int k = 0; k < 8; k += 4) {
          if (n0 + k < nv) {
            const float4 b = *reinterpret_cast<const float4*>(bias + n0 + k);
            const float y0 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 0]), bd.scale, b.x), bd.clo), bd.chi);
            const float y1 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 1]), bd.scale, b.y), bd.clo), bd.chi);
            const float y2 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 2]), bd.scale, b.z), bd.clo), bd.chi);
            const float y3 = fminf(fmaxf(fmaf(__uint_as_float(r[k + 3]), bd.scale, b.w), bd.clo), bd.chi);
            out[size_t(n0 + k) * TILE] = y0;
            if (n0 + k + 1 < nv) out[size_t(n0 + k + 1) * TILE] = y1;
            if (n0 + k + 2 < nv) out[size_t(n0 + k + 2) * TILE] = y2;
            if (n0 + k + 3 < nv) out[size_t(n0 + k + 3) * TILE] = y3;
          }
        }
      }
    };

    for (int node = node_begin; node < node_end; ++node) {
      const int it = node - node_begin, slot = it & 1;
      // ---- expansion of this node ----
      fr_wait(&gb[BKB_XFULL + rx.idx], rx.par);
      const float* xs = reinterpret_cast<const float*>(xring + size_t(g * bd.nstx + rx.idx) * bd.x_stage_bytes) + win;
      fr_wait(&bars[BKB_WFULL + rw.idx], rw.par);
      const float* head = reinterpret_cast<const float*>(wring + size_t(rw.idx) * BK_WSTAGE);
      __syncwarp();
      // the head rides in the first weight chunk, whose stage is recycled as soon as its MMAs are done: keep a copy per
      // warp (bias: for the store one node later; means: for every chunk of this node)
      for (int i = lane; i < bd.nn; i += 32) bias_w[slot * 64 + i] = head[i];
      for (int i = lane; i < bd.mean_floats; i += 32) mean[i] = head[bd.nn + i];
      __syncwarp();
      float xc[BK_TRI_N];                                      // centred, pre-scaled inputs of the product terms
#pragma unroll
      for (int i = 0; i < BK_TRI_N; ++i) {
        const int r = min(bd.tri_row0 + i, d_in - 1);
        xc[i] = (xs[r * TILE] - mean[r]) * bd.tri_scale;
      }
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c, rw.next()) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int gi = 0; gi < 2; ++gi) {
          const BkGroup gr = groups_s[c * 4 + 2 * half + gi];      // this warp's two 8-term groups of the chunk
          float v[8];
          if (gr.kind == BK_TRI) {
            bk_tri_group<BK_TRI_N, 0>(xc, gr.tri, v);
          } else {
            const float* xp = xs + size_t(gr.row0) * TILE;
            if (gr.cnt == 8) {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = xp[k * TILE];
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) v[k] = xp[min(k, gr.cnt - 1) * TILE];       // padding terms repeat the last row (zero weights)
            }
            if (gr.kind != BK_ID_RAW) {
              const float4 m0 = *reinterpret_cast<const float4*>(mean + gr.row0);
              const float4 m1 = *reinterpret_cast<const float4*>(mean + gr.row0 + 4);  // the mean vector is padded by 8
              v[0] -= m0.x; v[1] -= m0.y; v[2] -= m0.z; v[3] -= m0.w;
              v[4] -= m1.x; v[5] -= m1.y; v[6] -= m1.z; v[7] -= m1.w;
              if (gr.kind == BK_POW) {
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = abspow(v[k], bd.p);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 8; k += 2) fr_split(v[k], v[k + 1], hi[gi * 4 + k / 2], lo[gi * 4 + k / 2]);
        }
        fr_wait(&gb[BKB_AFREE + ra.idx], ra.par ^ 1u);
        tc_fence_after();
        const uint32_t col = lane_base + uint32_t(BK_COL_A + 32 * ra.idx + 8 * half);
        tmem_st8(col, hi);
        tmem_st8(col + 16, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&gb[BKB_AFULL + ra.idx]);
        ra.next();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&gb[BKB_XFREE + rx.idx]);     // receptive field consumed
      rx.next();
      // ---- outputs of the previous node (its MMAs have had a whole node's expansion to finish) ----
      if (it > 0) store_node(node - 1, slot ^ 1, uint32_t((it - 1) >> 1) & 1u);
    }
    if (node_end > node_begin) {
      const int it = node_end - 1 - node_begin;
      store_node(node_end - 1, it & 1, uint32_t(it >> 1) & 1u);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

}  // namespace hgsfa

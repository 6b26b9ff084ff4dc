// Flow plan: parses the plan blob written by pyfaceanalysis_b200/plan.py, owns the workspace and
// schedules the fused layer kernels (csrc/layer.cuh) over window chunks (sm_100a).
//
// Replaces mdp.Flow.execute over hinet.Switchboard / Layer / CloneLayer nodes whose children are
// SFANode / PCANode / WhiteningNode / GeneralExpansionNode / iGSFANode (reference call sites
// FaceDetectUpdated.py:699, face_analysis.py:1064,1257).
//
// Data layout in HBM ("TILED", hgsfa.h): activations of every layer are kept window-minor,
//   X[tile][feature][128 windows], so the receptive-field gather of a Switchboard becomes a list of
// contiguous byte ranges (feature runs) that TMA bulk copies fetch, and every warp-level access to one
// feature of one tile is a single 512-byte line set.
//
// Scheduling: layers with many nodes ("front" segment) run over chunks of front_chunk windows so that
// their activations stay L2-resident between launches; the last layers, which have only a few nodes,
// run once per back_chunk windows so that every launch still fills the 148 SMs.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "layer.cuh"
#include "layer_tc.cuh"
#include <cuda_fp16.h>

#include "front_api.h"

namespace hgsfa {

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

// uint8 fast path (what the detector and the benchmark feed): 128 windows x 128 features per CTA.
// Row-major rows are read as 16-byte vectors, the 128 x 128 byte block is transposed in 4 x 4 byte blocks
// with PRMT, and every tiled feature row is written as coalesced 32-bit words.  The staging tile is
// XOR-swizzled so that the column reads of the transpose are bank-conflict free.
__global__ void __launch_bounds__(256) tile_windows_u8_kernel(const uint8_t* __restrict__ src, int64_t n, int64_t dim,
                                                              int64_t ld, uint8_t* __restrict__ dst) {
  __shared__ uint32_t s[TILE][33];
  const int64_t tile = blockIdx.y;
  const int f0 = blockIdx.x * 128;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // phase 1: 128 rows x 8 chunks of 16 bytes
  for (int idx = tid; idx < TILE * 8; idx += 256) {
    const int w = idx >> 3, c = idx & 7;
    const int64_t gw = tile * TILE + w;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gw < n) {
      const uint8_t* p = src + gw * ld + f0 + c * 16;
      if (f0 + c * 16 + 16 <= dim) {
        v = __ldg(reinterpret_cast<const uint4*>(p));
      } else {
        uint32_t tmp[4] = {0u, 0u, 0u, 0u};
        for (int b = 0; b < 16; ++b)
          if (f0 + c * 16 + b < dim) tmp[b >> 2] |= uint32_t(p[b]) << (8 * (b & 3));
        v = make_uint4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
    }
    const int sw = w >> 5;
    s[w][(c * 4 + 0) ^ sw] = v.x; s[w][(c * 4 + 1) ^ sw] = v.y; s[w][(c * 4 + 2) ^ sw] = v.z; s[w][(c * 4 + 3) ^ sw] = v.w;
  }
  __syncthreads();
  // phase 2: lane = group of 4 windows, warp strides over groups of 4 features
  const int q = lane;
  for (int p = warp; p < 32; p += 8) {
    const int col = p ^ (q >> 3);
    const uint32_t r0 = s[4 * q + 0][col], r1 = s[4 * q + 1][col], r2 = s[4 * q + 2][col], r3 = s[4 * q + 3][col];
    // 4 x 4 byte transpose: t_j = byte j of (r0, r1, r2, r3)
    const uint32_t a0 = __byte_perm(r0, r1, 0x5140), a1 = __byte_perm(r0, r1, 0x7362);   // (r0.b0 r1.b0 r0.b1 r1.b1), (b2 b2 b3 b3)
    const uint32_t b0 = __byte_perm(r2, r3, 0x5140), b1 = __byte_perm(r2, r3, 0x7362);
    const uint32_t t0 = __byte_perm(a0, b0, 0x5410), t1 = __byte_perm(a0, b0, 0x7632);
    const uint32_t t2 = __byte_perm(a1, b1, 0x5410), t3 = __byte_perm(a1, b1, 0x7632);
    const int f = f0 + 4 * p;
    uint32_t* o = reinterpret_cast<uint32_t*>(dst + (tile * dim + f) * TILE) + q;
    if (f + 0 < dim) o[0 * (TILE / 4)] = t0;
    if (f + 1 < dim) o[1 * (TILE / 4)] = t1;
    if (f + 2 < dim) o[2 * (TILE / 4)] = t2;
    if (f + 3 < dim) o[3 * (TILE / 4)] = t3;
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
struct OpHost {
  OpDev dev;             // pointers valid on the device; layout fields filled per launch
  int64_t alg_flops, exe_flops;
  bool wide;             // some pass uses a 24 / 32 column register tile
  int scratch_floats;    // K-split scratch (floats)
  size_t smem_bytes[2];  // dynamic shared memory for [f32 input, u8 input]
  int nstages[2];
  bool tc;               // runs on the tensor cores (layer_tc.cuh) instead of the FFMA kernel
  TcOpDev tcd;
  size_t tc_smem[2];
  bool back;             // float inputs also run on the single-layer FP16-split kernel (back_tc.cuh)
  BackDev bkd;
};

// shared-memory layout of layer_tc_kernel for input element size `el`; returns total bytes
static size_t layout_tc(TcOpDev& d, int n_segs, int el) {
  auto up = [](size_t x) { return (x + 127) & ~size_t(127); };
  size_t off = TCB_BYTES;
  d.sm_terms = (int)off;
  off += up(size_t(d.n_terms) * sizeof(Term16));
  d.sm_toff = (int)off;
  off += up(size_t(d.n_terms) * 8);
  d.sm_segs = (int)off;
  off += up(size_t(n_segs) * sizeof(Seg));
  d.sm_chunkseg = (int)off;
  off += up(size_t(d.n_chunks + 1) * 4);
  d.sm_bias = (int)off;
  off += up(size_t(8) * d.Npad16 * 4);
  d.sm_raw_bytes = int(size_t(d.d_in) * TILE * el);
  d.sm_xstage_bytes = (int)up(size_t(d.twc) * d.sm_raw_bytes + size_t(d.head_floats) * 4);
  d.sm_x0 = (int)off;
  off += size_t(d.nstx) * d.sm_xstage_bytes;
  d.sm_wstage_bytes = (int)up(size_t(d.wchunk_floats) * 4);
  d.sm_w0 = (int)off;
  off += size_t(d.nw) * d.sm_wstage_bytes;
  return off;
}

static inline float tf32_rn(float x) {   // round to nearest TF32 (10 mantissa bits), ties away from zero
  uint32_t u;
  std::memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xffffe000u;
  std::memcpy(&x, &u, 4);
  return x;
}

// shared-memory layout for input element size `el`; returns total bytes (0 if nothing fits)
static size_t layout_op(OpDev& d, int scratch_floats, int el, int* nstages_out) {
  const size_t raw = size_t(d.d_in) * TILE * el;                       // per tile slot
  const size_t stage = (size_t(d.twc) * raw + size_t(d.param_floats) * 4 + 127) & ~size_t(127);
  size_t off = 128;                                                     // mbarriers
  d.sm_terms = (int)off;
  off += (size_t(d.n_terms) * sizeof(Term16) + 127) & ~size_t(127);
  d.sm_srows = (int)off;
  off += (size_t(d.twc) * d.n_rows * TILE * 4 + 127) & ~size_t(127);
  d.sm_scratch = (int)off;
  off += (size_t(scratch_floats) * 4 + 127) & ~size_t(127);
  d.sm_stage0 = (int)off;
  d.sm_stage_bytes = (int)stage;
  d.sm_raw_bytes = (int)raw;
  // stages: as many as fit in the CTA's share of shared memory (two 4-warp CTAs per SM, or one 8-warp CTA);
  // ops with a CTA-wide barrier per pass gain nothing beyond two
  const size_t limit = 227 * 1024;
  const size_t budget = (d.warps == 4) ? (limit / 2 - 1024) : limit;
  const int max_stages = d.simple ? 4 : 2;
  int ns = 1;
  while (ns < max_stages && ns < d.npc && off + size_t(ns + 1) * stage <= budget) ++ns;
  if (off + size_t(ns) * stage > limit) return 0;
  d.nstages = ns;
  *nstages_out = ns;
  return off + ns * stage;
}

}  // namespace hgsfa

using namespace hgsfa;

struct hgsfa_plan_s {
  int device = 0;
  cudaStream_t stream = nullptr;     // compute stream owned by the plan (host entry point)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  int64_t input_dim = 0, output_dim = 0;
  std::vector<OpHost> ops;
  DevBuf params;                     // the whole blob: every plan array is an offset into it
  DevBuf tin, front[2], mid, back[2], stage_x[2], stage_y[2];
  std::vector<DevBuf> tc_bufs;       // tensor-core operand images (weights split into TF32 hi / lo, chunked)
  // fused front (csrc/front_tc.cuh): layers 0-2 in one lane-resident kernel when the blob carries its section
  bool front_ok = false;
  FrontDev front_dev{};
  int front_np1 = 0, front_np2 = 0, front_img_h = 0;
  DevBuf fbuf;                       // third-layer output of a front chunk
  double front_ms = 0.0;
  // measured on B200 (profiles/README_r01.md): launches of >= 128 Ki windows hide the wave tail of the
  // one-CTA-per-SM layer kernels; smaller chunks only pay when the batch itself is small
  int64_t front_chunk = 524288, back_chunk = 524288;
  int split = 0;                     // ops [0, split) run per front chunk, [split, n) per back chunk
  int64_t launches = 0;
  double last_ms = 0.0;
  int sm_count = 148;
  int max_npc = 16;
  int max_npc_tc = 32;
  // optional per-op timing (hgsfa_plan_profile)
  bool profile = false;
  struct Stamp { int op; cudaEvent_t e0, e1; };
  std::vector<Stamp> stamps;
  std::vector<double> op_ms;
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

template <typename IN_T>
int launch_layer(hgsfa_plan_s* pl, OpHost& op, const void* xin, float* xout, int64_t ntiles, cudaStream_t st) {
  if (ntiles <= 0) return 0;
  const int v = sizeof(IN_T) == 1 ? 1 : 0;
  OpDev d = op.dev;
  // nodes per CTA, chosen per launch: as many as possible (the first load of a CTA is exposed, later ones
  // are prefetched behind the previous node's FMAs) while the grid still covers the SMs several times over
  {
    const int64_t groups = ceil_div(ntiles, d.twc);
    int64_t npc = (groups * d.n_nodes) / (int64_t(pl->sm_count) * 4);
    if (npc < 1) npc = 1;
    if (npc > pl->max_npc) npc = pl->max_npc;
    if (npc > d.n_nodes) npc = d.n_nodes;
    d.npc = (int)npc;
  }
  int ns = 1;
  const size_t smem = layout_op(d, op.scratch_floats, (int)sizeof(IN_T), &ns);
  HG_CHECK(smem > 0 && smem <= op.smem_bytes[v], "layer launch needs %zu bytes of shared memory, reserved %zu", smem,
           op.smem_bytes[v]);
  dim3 grid((unsigned)ceil_div(ntiles, d.twc), (unsigned)ceil_div(d.n_nodes, d.npc));
  if (op.wide)
    layer_kernel<IN_T, 32><<<grid, 32 * d.warps, smem, st>>>(d, static_cast<const IN_T*>(xin), xout, ntiles);
  else
    layer_kernel<IN_T, 16><<<grid, 32 * d.warps, smem, st>>>(d, static_cast<const IN_T*>(xin), xout, ntiles);
  pl->launches++;
  HG_CUDA(cudaGetLastError());
  return 0;
}

template <typename IN_T>
int launch_layer_tc(hgsfa_plan_s* pl, OpHost& op, const void* xin, float* xout, int64_t ntiles, cudaStream_t st) {
  if (ntiles <= 0) return 0;
  const int v = sizeof(IN_T) == 1 ? 1 : 0;
  TcOpDev d = op.tcd;
  {
    const int64_t groups = ceil_div(ntiles, d.twc);
    int64_t npc = (groups * d.n_nodes) / (int64_t(pl->sm_count) * 6);
    if (npc < 1) npc = 1;
    if (npc > pl->max_npc_tc) npc = pl->max_npc_tc;
    if (npc > d.n_nodes) npc = d.n_nodes;
    d.npc = (int)npc;
  }
  const size_t smem = layout_tc(d, d.n_segs, (int)sizeof(IN_T));
  HG_CHECK(smem <= op.tc_smem[v], "tensor-core layer launch needs %zu bytes of shared memory, reserved %zu", smem, op.tc_smem[v]);
  dim3 grid((unsigned)ceil_div(ntiles, d.twc), (unsigned)ceil_div(d.n_nodes, d.npc));
  if (d.f16) layer_tc_kernel<IN_T, true><<<grid, TC_THREADS, smem, st>>>(d, static_cast<const IN_T*>(xin), xout, ntiles);
  else layer_tc_kernel<IN_T, false><<<grid, TC_THREADS, smem, st>>>(d, static_cast<const IN_T*>(xin), xout, ntiles);
  pl->launches++;
  HG_CUDA(cudaGetLastError());
  return 0;
}

int plan_fail(hgsfa_plan_s* pl, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  pl->params.release();
  for (auto& b : pl->tc_bufs) b.release();
  delete pl;
  return fail("hgsfa_plan_create: %s", buf);
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
  HG_CHECK(std::memcmp(base, "HGSFAPL2", 8) == 0, "hgsfa_plan_create: bad magic (not a version-2 plan blob)");
  const int64_t* hdr = reinterpret_cast<const int64_t*>(base + 8);
  auto pl = new hgsfa_plan_s();
  pl->device = device;
  pl->input_dim = hdr[0];
  pl->output_dim = hdr[1];
  const int64_t n_ops = hdr[2];
  if (n_ops <= 0 || n_ops > 4096 || pl->input_dim <= 0 || pl->output_dim <= 0)
    return plan_fail(pl, "implausible header (n_ops=%lld in=%lld out=%lld)", (long long)n_ops, (long long)hdr[0],
                     (long long)hdr[1]);
  if (pl->params.reserve(nbytes)) { delete pl; return 1; }
  if (cudaMemcpy(pl->params.p, blob, nbytes, cudaMemcpyHostToDevice) != cudaSuccess)
    return plan_fail(pl, "parameter upload failed");
  const uint8_t* dbase = static_cast<const uint8_t*>(pl->params.p);
  auto dev_ptr = [&](const void* host) { return dbase + (static_cast<const uint8_t*>(host) - base); };

  const int64_t front_off = hdr[3], front_bytes = hdr[4];
  if (front_off < 0 || front_bytes < 0 || (front_off > 0 && (front_off % 16 || front_off < 64 || front_off + front_bytes > (int64_t)nbytes)))
    return plan_fail(pl, "fused-front section [%lld, +%lld) outside the blob", (long long)front_off, (long long)front_bytes);
  Cursor cur{base + 64, base + (front_off > 0 ? size_t(front_off) : nbytes)};
  int64_t cur_dim = pl->input_dim;
  for (int64_t o = 0; o < n_ops && cur.ok; ++o) {
    const int64_t* oh = cur.take<int64_t>(16);
    if (!oh) break;
    OpHost op{};
    OpDev& d = op.dev;
    d.n_nodes = (int)oh[0]; d.d_in = (int)oh[1]; d.in_dim = (int)oh[2]; d.out_dim = (int)oh[3];
    d.n_passes = (int)oh[4]; d.shared = (int)(oh[5] & 0xff); d.warps = (int)((oh[5] >> 8) & 0xff); d.n_rows = (int)oh[6];
    op.tc = ((oh[5] >> 16) & 0xff) == 1;
    d.twc = (int)oh[7];
    op.alg_flops = oh[8]; op.exe_flops = oh[9];
    d.npc = (int)oh[10]; d.n_runs = (int)oh[11];
    {
      double clip[2];
      std::memcpy(clip, oh + 12, sizeof(clip));
      d.clip_lo = float(clip[0]); d.clip_hi = float(clip[1]);
    }
    d.param_floats = (int)oh[14]; d.n_terms = (int)oh[15];
    const bool sane = d.n_nodes > 0 && d.n_nodes <= 65535 * 64 && d.d_in > 0 && d.d_in < 32768 && d.in_dim == cur_dim &&
                      d.out_dim > 0 && d.n_passes >= 1 && d.n_passes <= MAX_PASSES &&
                      (op.tc ? (d.twc >= 1 && d.twc <= TC_MAX_TW) : (d.twc == 1 || d.twc == 2 || d.twc == 4 || d.twc == 8 || d.twc == 16)) && d.n_rows >= 0 && d.npc >= 1 &&
                      (d.warps == 4 || d.warps == 8) && d.n_runs >= 1 && d.n_runs <= d.d_in && d.param_floats > 0 && d.param_floats % 4 == 0 &&
                      d.n_terms > 0;
    if (!sane)
      return plan_fail(pl, "op %lld has an inconsistent header (nodes=%d d_in=%d in_dim=%d expected %lld twc=%d)",
                       (long long)o, d.n_nodes, d.d_in, d.in_dim, (long long)cur_dim, d.twc);
    const int n_w = d.shared ? 1 : d.n_nodes;
    const Run* runs = cur.take<Run>(size_t(d.n_nodes) * d.n_runs);
    const int32_t* out_col = cur.take<int32_t>(d.n_nodes);
    const float* params = cur.take<float>(size_t(n_w) * d.param_floats);
    const Term16* terms = cur.take<Term16>(d.n_terms);
    if (!cur.ok) break;
    for (int nd = 0; nd < d.n_nodes; ++nd) {   // the runs of a node must tile its d_in rows inside the input buffer
      int covered = 0;
      for (int r = 0; r < d.n_runs; ++r) {
        const Run& rn = runs[size_t(nd) * d.n_runs + r];
        if (rn.len == 0) continue;
        if (rn.len < 0 || rn.i0 != covered || rn.f0 < 0 || rn.f0 + rn.len > d.in_dim)
          return plan_fail(pl, "op %lld node %d: bad gather run (i0=%d f0=%d len=%d)", (long long)o, nd, rn.i0, rn.f0, rn.len);
        covered += rn.len;
      }
      if (covered != d.d_in) return plan_fail(pl, "op %lld node %d: gather runs cover %d of %d rows", (long long)o, nd, covered, d.d_in);
    }
    d.runs = reinterpret_cast<const Run*>(dev_ptr(runs));
    d.out_col = reinterpret_cast<const int*>(dev_ptr(out_col));
    d.params = reinterpret_cast<const float*>(dev_ptr(params));
    d.terms = reinterpret_cast<const Term16*>(dev_ptr(terms));
    const int n_src = d.d_in + d.n_rows;
    for (int k = 0; k < d.n_terms; ++k)
      if (terms[k].i < 0 || terms[k].i >= n_src || terms[k].j < 0 || terms[k].j >= n_src || terms[k].k < 0 || terms[k].k >= n_src)
        return plan_fail(pl, "op %lld term %d references a row outside [0, %d)", (long long)o, k, n_src);
    op.scratch_floats = 0;
    for (int p = 0; p < d.n_passes && cur.ok; ++p) {
      const int64_t* ph = cur.take<int64_t>(16);
      if (!ph) break;
      PassDev& dp = d.pass[p];
      dp.K = (int)ph[0]; dp.Npad = (int)ph[1]; dp.NT = (int)ph[2]; dp.NTL = (int)ph[3]; dp.KS = (int)ph[4];
      dp.TW = (int)ph[5]; dp.dst = (int)ph[6]; dp.row0 = (int)ph[7]; dp.w_off = (int)ph[8]; dp.b_off = (int)ph[9];
      dp.term_off = (int)ph[10]; dp.n_seg = (int)ph[11]; dp.SW = (int)ph[14];
      const bool tiles_ok =
          op.tc ? (d.n_passes == 1 && d.n_rows == 0 && dp.dst == DST_GLOBAL && dp.Npad >= 1 && dp.Npad % 4 == 0)
                : (dp.NT >= 8 && dp.NT <= 32 && dp.NT % 4 == 0 && dp.NTL >= 1 && dp.KS >= 1 && dp.TW >= 1 &&
                   dp.NTL * dp.KS * dp.TW == d.warps && dp.Npad == dp.NT * dp.NTL && (dp.SW == 1 || dp.SW == 2) &&
                   (dp.SW == 1 || dp.NT <= 16) && d.twc % dp.SW == 0 && (dp.KS & (dp.KS - 1)) == 0);
      const bool psane =
          dp.K > 0 && tiles_ok && (dp.dst & (DST_GLOBAL | DST_ROWS)) && dp.row0 >= 0 &&
          (!(dp.dst & DST_ROWS) || dp.row0 + dp.Npad <= d.n_rows) && dp.w_off >= 0 && dp.w_off % 4 == 0 && dp.b_off >= 0 &&
          dp.w_off + dp.K * dp.Npad <= d.param_floats && dp.b_off + dp.Npad <= d.param_floats && dp.term_off >= 0 &&
          dp.term_off + dp.K <= d.n_terms && dp.n_seg >= 1;
      if (!psane)
        return plan_fail(pl, "op %lld pass %d inconsistent (K=%d Npad=%d NT=%d NTL=%d KS=%d TW=%d dst=%d row0=%d)",
                         (long long)o, p, dp.K, dp.Npad, dp.NT, dp.NTL, dp.KS, dp.TW, dp.dst, dp.row0);
      const Seg* segs = cur.take<Seg>(dp.n_seg);
      const int32_t* n_valid = cur.take<int32_t>(d.n_nodes);
      const int32_t* col_off = cur.take<int32_t>(d.n_nodes);
      if (!cur.ok) break;
      int kcov = 0;
      for (int sgi = 0; sgi < dp.n_seg; ++sgi) {
        if (segs[sgi].op < 0 || segs[sgi].op > OP_ID_POW || segs[sgi].k0 != kcov || segs[sgi].k1 < segs[sgi].k0 ||
            segs[sgi].kind < 0 || segs[sgi].kind > 2 ||
            (segs[sgi].op != OP_ID_POW && segs[sgi].ibase >= 0 && segs[sgi].ibase + (segs[sgi].k1 - segs[sgi].k0) > d.d_in + d.n_rows) ||
            (segs[sgi].op != OP_ID_POW && segs[sgi].ibase >= 0 && segs[sgi].kind == 0 && segs[sgi].ibase + (segs[sgi].k1 - segs[sgi].k0) > d.d_in) ||
            (segs[sgi].ibase >= 0 && segs[sgi].kind == 1 && segs[sgi].ibase < d.d_in) ||
            (segs[sgi].op == OP_ID_POW && (segs[sgi].kind != 0 || segs[sgi].ibase < 0 || ((segs[sgi].k1 - segs[sgi].k0) & 1) ||
                                           segs[sgi].ibase + (segs[sgi].k1 - segs[sgi].k0) / 2 > d.d_in)))
          return plan_fail(pl, "op %lld pass %d: bad term segment %d", (long long)o, p, sgi);
        kcov = segs[sgi].k1;
      }
      if (kcov != dp.K) return plan_fail(pl, "op %lld pass %d: segments cover %d of %d terms", (long long)o, p, kcov, dp.K);
      for (int nd = 0; nd < d.n_nodes; ++nd)
        if ((dp.dst & DST_GLOBAL) && (n_valid[nd] < 0 || n_valid[nd] > dp.Npad || out_col[nd] + col_off[nd] < 0 ||
                                      out_col[nd] + col_off[nd] + n_valid[nd] > d.out_dim))
          return plan_fail(pl, "op %lld pass %d node %d writes outside the output buffer", (long long)o, p, nd);
      dp.segs = reinterpret_cast<const Seg*>(dev_ptr(segs));
      dp.n_valid = reinterpret_cast<const int*>(dev_ptr(n_valid));
      dp.col_off = reinterpret_cast<const int*>(dev_ptr(col_off));
      if (!op.tc && dp.NT * dp.SW > 16) op.wide = true;
      if (!op.tc && dp.KS > 1) op.scratch_floats = std::max(op.scratch_floats, (d.warps / 2) * dp.SW * dp.NT * TILE);
      if (op.tc) {
        // ---- tensor-core operand images (layer_tc.cuh): ph[2..5] = accumulator sets, receptive-field stages,
        // weight-ring stages, A stages chosen by the plan compiler; everything else is derived here
        TcOpDev& t = op.tcd;
        t = TcOpDev{};
        t.n_nodes = d.n_nodes; t.d_in = d.d_in; t.in_dim = d.in_dim; t.out_dim = d.out_dim; t.shared = d.shared;
        t.twc = d.twc; t.npc = 1; t.n_runs = d.n_runs; t.clip_lo = d.clip_lo; t.clip_hi = d.clip_hi;
        t.nd = (int)ph[2]; t.nstx = (int)ph[3]; t.nw = (int)ph[4]; t.na = (int)ph[5];
        int n_max = 1;
        for (int nd = 0; nd < d.n_nodes; ++nd) n_max = std::max(n_max, (int)n_valid[nd]);
        t.K = dp.K; t.Npad16 = (n_max + 15) & ~15;
        t.n_terms = d.n_terms;
        const int d_pad = (d.d_in + 3) & ~3;
        t.head_floats = (d_pad + t.Npad16 + 2 * d.n_terms + 3) & ~3;
        t.wchunk_floats = 2 * TC_CK * t.Npad16;
        const int cols = t.nd * t.twc * t.Npad16 + t.na * 2 * TC_CK;
        t.tmem_cols = 32;
        while (t.tmem_cols < cols) t.tmem_cols *= 2;
        if (!(t.nd == 1 || t.nd == 2) || !(t.nstx == 1 || t.nstx == 2) || t.nw < 2 || t.nw > 4 || !(t.na >= 1 && t.na <= 4) ||
            t.twc > TC_MAX_TW || t.Npad16 > 128 || cols > 512 || n_max > dp.Npad)
          return plan_fail(pl, "op %lld: bad tensor-core configuration (twc=%d Npad16=%d nd=%d nstx=%d nw=%d na=%d cols=%d)",
                           (long long)o, t.twc, t.Npad16, t.nd, t.nstx, t.nw, t.na, cols);
        // segments: identity+power fusions undone; every segment starts at a multiple of 8 A columns and is
        // padded to a multiple of 8 (zero weight rows), then split at chunk boundaries.  For the kernel a
        // piece is (op, k0, k1 = A columns, p, kind = real terms, ibase, nomean, pad1 = first term-table entry).
        std::vector<Seg> tsegs;
        std::vector<int> knew(dp.K, 0);          // term k -> A column
        int cursor = 0;
        int ck = TC_CK;                          // terms per chunk (TC_CK16 for the FP16 form, decided below)
        const char* seg_error = nullptr;
        auto build_segs = [&]() {
        tsegs.clear();
        cursor = 0;
        auto push = [&](Seg sgm) {               // sgm.k0 / k1: ORIGINAL term range
          int real = sgm.k1 - sgm.k0, tk = sgm.k0, ib = sgm.ibase;
          for (int k = sgm.k0; k < sgm.k1; ++k) knew[k] = cursor + (k - sgm.k0);
          int pos = cursor;
          const int end = cursor + ((real + 7) & ~7);
          while (pos < end) {
            Seg piece = sgm;
            piece.k0 = pos;
            piece.k1 = std::min(end, (pos / ck + 1) * ck);
            const int len = piece.k1 - piece.k0;
            piece.kind = std::min(real, len);
            piece.ibase = ib;
            piece.pad1 = tk;
            if (sgm.op == OP_TRI) { piece.ibase = sgm.ibase; piece.nomean = tk - sgm.k0; }
            tsegs.push_back(piece);
            if (ib >= 0) ib += len;
            tk += piece.kind;
            real -= piece.kind;
            pos = piece.k1;
          }
          cursor = end;
        };
        for (int sgi = 0; sgi < dp.n_seg; ++sgi) {
          Seg sgm = segs[sgi];
          if (sgm.kind != 0) { seg_error = "tensor-core ops take receptive-field operands only"; return; }
          if (sgm.op == OP_ID_POW) {
            const int half = (sgm.k1 - sgm.k0) / 2;
            Seg a = sgm, b = sgm;
            a.op = OP_ID; a.k1 = sgm.k0 + half;
            b.op = OP_ABSPOW; b.k0 = sgm.k0 + half; b.nomean = 0;
            push(a); push(b);
          } else {
            if (sgm.op == OP_MUL) {
              // all products x_i x_j, r0 <= i <= j < r0 + n in row-major order (QT expansion)?  -> register-resident form
              const int len = sgm.k1 - sgm.k0, r0 = terms[sgm.k0].i;
              int n = 0;
              for (int c = 3; c <= 16; ++c) if (c * (c + 1) / 2 == len && tc_tri_size(c)) n = c;
              bool tri = n > 0;
              for (int q = 0; tri && q < len; ++q)
                tri = terms[sgm.k0 + q].i == r0 + tri_row(n, q) && terms[sgm.k0 + q].j == r0 + tri_col(n, q);
              if (tri && !getenv("HGSFA_TC_NO_TRI")) {
                sgm.op = OP_TRI;
                sgm.p = float(n);
                sgm.ibase = r0;
              }
            }
            push(sgm);
          }
        }
        };
        build_segs();
        if (seg_error) return plan_fail(pl, "op %lld: %s", (long long)o, seg_error);
        // ---- FP16 pieces instead of TF32 pieces (layer_tc.cuh, F16 = true; HGSFA_TC_F16=0 turns it off) where every operand
        // provably fits FP16's range: float inputs bounded by the previous op's saturation, term kinds with a known bound,
        // operands of products pre-scaled.  Two terms share a tensor-memory column, so a chunk holds 64 terms in the columns
        // and weight bytes the TF32 form needs for 32: half the hand-overs between the expansion warps and the MMA warp,
        // which cost ~700 cycles each against ~1 200 for 32 terms of arithmetic (profiles/README_r02.md items 13-14).
        // Layers 3-10 of U11L_64: 10.8 instead of 12.4 ms per 1 Mi windows; accuracy as 3xTF32 (22 significant bits).
        t.f16 = 0;
        t.scale = 1.0f;
        t.prod_scale = 1.0f;
        double f16_wmul = 1.0, f16_prod_w = 1.0;
        std::vector<uint8_t> is_prod(dp.K, 0);
        {
          const char* env = getenv("HGSFA_TC_F16");
          const char* env_back = getenv("HGSFA_BACK");                     // the single-layer kernel of back_tc.cuh takes these ops
          bool ok = !(env && env[0] == '0') && !(env_back && env_back[0] == '1') && o > 0;
          double bound_c = 0.0;
          if (ok) {
            const OpDev& prev = pl->ops[o - 1].dev;
            const double bound_in = std::max(std::fabs((double)prev.clip_lo), std::fabs((double)prev.clip_hi));
            double max_mean = 0.0;
            for (int w = 0; w < n_w; ++w)
              for (int i = 0; i < d.d_in; ++i) max_mean = std::max(max_mean, std::fabs((double)params[size_t(w) * d.param_floats + i]));
            bound_c = bound_in + max_mean;
            ok = std::isfinite(bound_c) && bound_c > 0.0 && bound_c <= 32768.0;
          }
          for (const Seg& pc : tsegs) {
            if (!ok) break;
            switch (pc.op) {
              case OP_ID: case OP_ABS: case OP_CLIP: break;
              case OP_ABSPOW: case OP_SGNPOW:
                ok = pc.p > 0.f && (pc.p <= 1.f || std::pow(std::max(1.0, bound_c), (double)pc.p) <= 32768.0);
                break;
              case OP_MUL: case OP_TRI: break;                      // operands pre-scaled below
              case OP_MUL3: ok = bound_c * bound_c * bound_c <= 32768.0; break;
              default: ok = false; break;
            }
          }
          double max_w = 0.0;
          int prod_shift = 0;
          if (ok) {
            while (bound_c / double(1 << prod_shift) > 128.0 && prod_shift < 12) ++prod_shift;   // products stay below 2^14
            f16_prod_w = std::ldexp(1.0, 2 * prod_shift);
            for (const Seg& pc : tsegs)
              if (pc.op == OP_TRI || pc.op == OP_MUL)
                for (int q = 0; q < pc.kind; ++q) is_prod[pc.pad1 + q] = 1;
            for (int w = 0; w < n_w; ++w)
              for (int k = 0; k < dp.K; ++k)
                for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n)
                  max_w = std::max(max_w, std::fabs((double)params[size_t(w) * d.param_floats + dp.w_off + size_t(k) * dp.Npad + n]) *
                                              (is_prod[k] ? f16_prod_w : 1.0));
            ok = max_w > 0.0 && std::isfinite(max_w);
          }
          if (ok) {
            const int tpow = (int)std::floor(std::log2(16384.0 / max_w));      // weights stored times 2^tpow: |w| <= 2^14
            t.f16 = 1;
            t.scale = (float)std::ldexp(1.0, -tpow);
            t.prod_scale = (float)std::ldexp(1.0, -prod_shift);
            f16_wmul = std::ldexp(1.0, tpow);
            // chunks of 2 x TC_CK terms: the same tensor-memory columns and weight-chunk bytes as the TF32 form
            ck = TC_CK16;
            build_segs();
          }
        }
        t.Kpad = cursor;
        t.n_chunks = (t.Kpad + ck - 1) / ck;
        std::vector<int32_t> chunk_seg(t.n_chunks + 1, 0);
        {
          size_t sgi = 0;
          for (int c = 0; c < t.n_chunks; ++c) {
            chunk_seg[c] = (int32_t)sgi;
            while (sgi < tsegs.size() && tsegs[sgi].k0 < (c + 1) * ck) ++sgi;
          }
          chunk_seg[t.n_chunks] = (int32_t)tsegs.size();
        }
        t.n_segs = (int)tsegs.size();
        // head = x_mean | bias;  weight chunks = TF32 hi image | lo image, canonical K-major core matrices
        const size_t head_bytes = size_t(n_w) * t.head_floats * 4;
        const size_t wimg_bytes = size_t(n_w) * t.n_chunks * t.wchunk_floats * 4;
        const size_t seg_bytes = (tsegs.size() * sizeof(Seg) + 15) & ~size_t(15);
        const size_t cs_bytes = (chunk_seg.size() * 4 + 15) & ~size_t(15);
        std::vector<uint8_t> host(head_bytes + wimg_bytes + seg_bytes + cs_bytes, 0);
        float* head = reinterpret_cast<float*>(host.data());
        float* wimg = reinterpret_cast<float*>(host.data() + head_bytes);
        const int nb8 = t.Npad16 / 8;
        for (int w = 0; w < n_w; ++w) {
          const float* pw = params + size_t(w) * d.param_floats;
          for (int i = 0; i < d.d_in; ++i) head[size_t(w) * t.head_floats + i] = pw[i];
          for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n) head[size_t(w) * t.head_floats + d_pad + n] = pw[dp.b_off + n];
          for (int k = 0; k < d.n_terms; ++k) {     // operand means of the product terms, in term-table order
            head[size_t(w) * t.head_floats + d_pad + t.Npad16 + 2 * k] = pw[terms[k].i];
            head[size_t(w) * t.head_floats + d_pad + t.Npad16 + 2 * k + 1] = pw[terms[k].j];
          }
          for (int k = 0; k < dp.K; ++k) {
            const int c = knew[k] / ck, kk = knew[k] % ck;
            float* img = wimg + (size_t(w) * t.n_chunks + c) * t.wchunk_floats;
            for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n) {
              const float wv = pw[dp.w_off + size_t(k) * dp.Npad + n];
              if (t.f16) {
                // canonical K-major core matrices of 8 rows x 8 halves; hi image, then lo image
                __half* hi16 = reinterpret_cast<__half*>(img);
                const size_t off = (size_t((kk >> 3) * nb8 + (n >> 3)) * 8 + (n & 7)) * 8 + (kk & 7);
                const double ws = double(wv) * f16_wmul * (is_prod[k] ? f16_prod_w : 1.0);
                const __half h = __float2half_rn(float(ws));
                hi16[off] = h;
                hi16[size_t(TC_CK16) * t.Npad16 + off] = __float2half_rn(float(ws - double(__half2float(h))));
                continue;
              }
              const float hi = tf32_rn(wv);
              const size_t off = (size_t((kk >> 2) * nb8 + (n >> 3)) * 8 + (n & 7)) * 4 + (kk & 3);
              img[off] = hi;
              img[size_t(TC_CK) * t.Npad16 + off] = tf32_rn(wv - hi);
            }
          }
        }
        std::memcpy(host.data() + head_bytes + wimg_bytes, tsegs.data(), tsegs.size() * sizeof(Seg));
        std::memcpy(host.data() + head_bytes + wimg_bytes + seg_bytes, chunk_seg.data(), chunk_seg.size() * 4);
        pl->tc_bufs.emplace_back();
        DevBuf& buf = pl->tc_bufs.back();
        if (buf.reserve(host.size())) return plan_fail(pl, "out of device memory for tensor-core operands");
        if (cudaMemcpy(buf.p, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess)
          return plan_fail(pl, "tensor-core operand upload failed");
        const uint8_t* tb = static_cast<const uint8_t*>(buf.p);
        t.head = reinterpret_cast<const float*>(tb);
        t.wimg = reinterpret_cast<const float*>(tb + head_bytes);
        t.segs = reinterpret_cast<const Seg*>(tb + head_bytes + wimg_bytes);
        t.chunk_seg = reinterpret_cast<const int*>(tb + head_bytes + wimg_bytes + seg_bytes);
        t.runs = d.runs; t.out_col = d.out_col; t.n_valid = dp.n_valid; t.col_off = dp.col_off; t.terms = d.terms;
        for (int v = 0; v < 2; ++v) {
          TcOpDev tmp = t;
          op.tc_smem[v] = layout_tc(tmp, t.n_segs, v ? 1 : 4);
        }
        // ---- single-layer FP16-split kernel (back_tc.cuh): float inputs bounded by the previous op's saturation, term
        // segments of the kinds it has straight-line code for
        op.back = false;
        {
          // opt-in (HGSFA_BACK=1): measured on U11L_64 it is correct and slightly more accurate than the 3xTF32 kernel but
          // 15 % slower with 8 expansion warps per SM (14.5 vs 12.5 ms for ops 3-10, profiles/README_r02.md)
          const char* env = getenv("HGSFA_BACK");
          bool ok = env && env[0] == '1' && !t.f16 && o > 0 && t.Npad16 <= 64 && t.n_chunks <= 64;
          float bound_in = 0.f;
          if (ok) {
            const OpDev& prev = pl->ops[o - 1].dev;
            bound_in = std::max(std::fabs(prev.clip_lo), std::fabs(prev.clip_hi));
            ok = std::isfinite(bound_in) && bound_in > 0.f;
          }
          float pexp = 0.f;
          int tri_row0 = 0, n_tri = 0;
          for (const Seg& pc : tsegs) {
            if (!ok) break;
            if (pc.op == OP_ID || pc.op == OP_ABSPOW) {
              ok = pc.ibase >= 0 && pc.ibase % 4 == 0;
              if (pc.op == OP_ABSPOW) {
                ok = ok && (pexp == 0.f || pexp == pc.p) && pc.p > 0.f && pc.p <= 1.f;
                pexp = pc.p;
              }
            } else if (pc.op == OP_TRI) {
              ok = int(pc.p) == BK_TRI_N && (n_tri == 0 || tri_row0 == pc.ibase);
              tri_row0 = pc.ibase;
              ++n_tri;
            } else {
              ok = false;
            }
          }
          const int mean_floats = ((d.d_in + 7) & ~7) + 8;
          ok = ok && t.Npad16 + mean_floats <= FR_HEAD / 4;
          double max_mean = 0.0, max_w = 0.0;
          int tri_shift = 0;
          std::vector<uint8_t> is_tri(dp.K, 0);
          if (ok) {
            for (int w = 0; w < n_w; ++w)
              for (int i = 0; i < d.d_in; ++i) max_mean = std::max(max_mean, (double)std::fabs(params[size_t(w) * d.param_floats + i]));
            const double bound_c = double(bound_in) + max_mean;
            ok = bound_c <= 16384.0;
            while (bound_c / double(1 << tri_shift) > 128.0 && tri_shift < 12) ++tri_shift;     // products stay below 2^14
            for (const Seg& pc : tsegs)
              if (pc.op == OP_TRI)
                for (int q = 0; q < pc.kind; ++q) is_tri[pc.pad1 + q] = 1;
            const double tri_w = std::ldexp(1.0, 2 * tri_shift);
            for (int w = 0; w < n_w && ok; ++w)
              for (int k = 0; k < dp.K; ++k)
                for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n)
                  max_w = std::max(max_w, std::fabs((double)params[size_t(w) * d.param_floats + dp.w_off + size_t(k) * dp.Npad + n]) *
                                              (is_tri[k] ? tri_w : 1.0));
            ok = ok && max_w > 0.0 && std::isfinite(max_w);
          }
          if (ok) {
            BackDev& b = op.bkd;
            b = BackDev{};
            b.n_nodes = d.n_nodes; b.d_in = d.d_in; b.in_dim = d.in_dim; b.out_dim = d.out_dim; b.shared = d.shared;
            b.n_runs = d.n_runs; b.npc = 1; b.nn = t.Npad16; b.n_chunks = t.n_chunks; b.tri_row0 = tri_row0;
            b.mean_floats = mean_floats; b.chunk_bytes = FR_HEAD + 32 * t.Npad16 * 4;
            const int tpow = (int)std::floor(std::log2(16384.0 / max_w));
            b.scale = (float)std::ldexp(1.0, -tpow);
            b.clo = d.clip_lo; b.chi = d.clip_hi; b.p = pexp > 0.f ? pexp : 1.f;
            b.tri_scale = (float)std::ldexp(1.0, -tri_shift);
            ok = back_layout(b) > 0;
            if (ok) {
              // group table: four 8-term groups per chunk
              std::vector<BkGroup> groups(size_t(t.n_chunks) * 4, BkGroup{BK_ID_RAW, 0, 1, 0});
              for (const Seg& pc : tsegs) {
                const int ngr = (pc.k1 - pc.k0) / 8;
                for (int q = 0; q < ngr; ++q) {
                  BkGroup& gq = groups[pc.k0 / 8 + q];
                  const int cnt = std::max(1, std::min(8, pc.kind - 8 * q));
                  if (pc.op == OP_TRI) gq = BkGroup{BK_TRI, 0, 8, pc.nomean / 8 + q};
                  else gq = BkGroup{pc.op == OP_ABSPOW ? BK_POW : (pc.nomean ? BK_ID_RAW : BK_ID), pc.ibase + 8 * q, cnt, 0};
                }
              }
              const size_t img_bytes = size_t(n_w) * t.n_chunks * b.chunk_bytes;
              const size_t grp_bytes = groups.size() * sizeof(BkGroup);
              const size_t oc_bytes = (size_t(d.n_nodes) * 4 + 15) & ~size_t(15);
              std::vector<uint8_t> hb(img_bytes + grp_bytes + oc_bytes, 0);
              const double wmul = std::ldexp(1.0, tpow), tri_w = std::ldexp(1.0, 2 * tri_shift);
              const int nb8 = t.Npad16 / 8;
              for (int w = 0; w < n_w; ++w) {
                const float* pw = params + size_t(w) * d.param_floats;
                uint8_t* node_img = hb.data() + size_t(w) * t.n_chunks * b.chunk_bytes;
                float* head = reinterpret_cast<float*>(node_img);               // chunk 0: bias[nn] | mean[mean_floats]
                for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n) head[n] = pw[dp.b_off + n];
                for (int i = 0; i < d.d_in; ++i) head[t.Npad16 + i] = pw[i];
                for (int k = 0; k < dp.K; ++k) {
                  const int c = knew[k] / 32, kk = knew[k] % 32;
                  __half* hi = reinterpret_cast<__half*>(node_img + size_t(c) * b.chunk_bytes + FR_HEAD);
                  __half* lo = hi + 32 * t.Npad16;
                  for (int n = 0; n < std::min(dp.Npad, t.Npad16); ++n) {
                    const double wv = double(pw[dp.w_off + size_t(k) * dp.Npad + n]) * (is_tri[k] ? tri_w : 1.0) * wmul;
                    const size_t off = (size_t((kk >> 3) * nb8 + (n >> 3)) * 8 + (n & 7)) * 8 + (kk & 7);
                    const __half h = __float2half_rn(float(wv));
                    hi[off] = h;
                    lo[off] = __float2half_rn(float(wv - double(__half2float(h))));
                  }
                }
              }
              std::memcpy(hb.data() + img_bytes, groups.data(), grp_bytes);
              int32_t* oc = reinterpret_cast<int32_t*>(hb.data() + img_bytes + grp_bytes);
              for (int nd = 0; nd < d.n_nodes; ++nd) oc[nd] = out_col[nd] + col_off[nd];
              pl->tc_bufs.emplace_back();
              DevBuf& bb = pl->tc_bufs.back();
              if (bb.reserve(hb.size())) return plan_fail(pl, "out of device memory for the FP16 operand images");
              if (cudaMemcpy(bb.p, hb.data(), hb.size(), cudaMemcpyHostToDevice) != cudaSuccess)
                return plan_fail(pl, "FP16 operand upload failed");
              const uint8_t* base_b = static_cast<const uint8_t*>(bb.p);
              b.wimg = base_b;
              b.groups = reinterpret_cast<const BkGroup*>(base_b + img_bytes);
              b.out_col = reinterpret_cast<const int*>(base_b + img_bytes + grp_bytes);
              b.n_valid = t.n_valid;
              b.runs = d.runs;
              if (back_set_attributes()) { std::string msg = last_error_ref(); return plan_fail(pl, "%s", msg.c_str()); }
              op.back = true;
            }
          }
        }
        if (op.tc_smem[0] > size_t(227) * 1024)
          return plan_fail(pl, "op %lld: tensor-core layout needs %zu bytes of shared memory", (long long)o, op.tc_smem[0]);
      }
    }
    if (!cur.ok) break;
    // barrier-free hand-over of stages for one-pass, no-K-split ops: measured equal to the per-node barrier on
    // B200 (profiles/README_r01.md), so it stays opt-in
    d.simple = (getenv("HGSFA_SIMPLE") && d.n_passes == 1 && d.pass[0].KS == 1) ? 1 : 0;
    for (int v = 0; v < 2; ++v) {
      OpDev tmp = d;
      tmp.npc = 64;   // worst case for the reservation: as many stages as ever fit
      op.smem_bytes[v] = layout_op(tmp, op.scratch_floats, v ? 1 : 4, &op.nstages[v]);
    }
    if (op.tc) { op.smem_bytes[0] = op.smem_bytes[1] = 1; }
    if (op.smem_bytes[0] == 0)
      return plan_fail(pl, "op %lld does not fit in 227 KB of shared memory (d_in=%d twc=%d)", (long long)o, d.d_in, d.twc);
    cur_dim = d.out_dim;
    pl->ops.push_back(op);
  }
  if (!cur.ok || (int64_t)pl->ops.size() != n_ops || cur_dim != pl->output_dim)
    return plan_fail(pl, "truncated or inconsistent blob (%zu of %lld ops parsed, final dim %lld vs %lld)", pl->ops.size(),
                     (long long)n_ops, (long long)cur_dim, (long long)hdr[1]);
  size_t max_smem = 0, max_tc = 0;
  for (auto& op : pl->ops) {
    if (op.tc) max_tc = std::max(max_tc, std::max(op.tc_smem[0], op.tc_smem[1]));
    else max_smem = std::max(max_smem, std::max(op.smem_bytes[0], op.smem_bytes[1]));
  }
  // The dynamic shared-memory limit of a kernel is per device and shared by every plan of the process: it only ever
  // grows (a plan created later with smaller layers must not shrink it under the feet of an earlier plan -- that
  // made the earlier plan's launches fail with "invalid argument").
  {
    static std::mutex mu;
    static size_t dev_max_tc[64] = {}, dev_max_smem[64] = {};
    std::lock_guard<std::mutex> lock(mu);
    size_t& cur_tc = dev_max_tc[device & 63];
    size_t& cur_smem = dev_max_smem[device & 63];
    if (max_tc > cur_tc) {
      if (cudaFuncSetAttribute(layer_tc_kernel<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_tc) != cudaSuccess ||
          cudaFuncSetAttribute(layer_tc_kernel<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_tc) != cudaSuccess ||
          cudaFuncSetAttribute(layer_tc_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_tc) != cudaSuccess ||
          cudaFuncSetAttribute(layer_tc_kernel<float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_tc) != cudaSuccess)
        return plan_fail(pl, "cannot reserve %zu bytes of dynamic shared memory for the tensor-core kernel", max_tc);
      cur_tc = max_tc;
    }
    if (max_smem > cur_smem) {
      cudaError_t es[4] = {
          cudaFuncSetAttribute(layer_kernel<uint8_t, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
          cudaFuncSetAttribute(layer_kernel<float, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
          cudaFuncSetAttribute(layer_kernel<uint8_t, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem),
          cudaFuncSetAttribute(layer_kernel<float, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem)};
      for (cudaError_t e : es)
        if (e != cudaSuccess)
          return plan_fail(pl, "cannot reserve %zu bytes of dynamic shared memory: %s", max_smem, cudaGetErrorString(e));
      cur_smem = max_smem;
    }
  }
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess && prop.multiProcessorCount > 0) pl->sm_count = prop.multiProcessorCount;
    if (const char* e = getenv("HGSFA_MAX_NPC")) pl->max_npc = std::max(1, atoi(e));
  }
  // back segment = trailing ops with few nodes: they need many windows per launch to fill 148 SMs
  pl->split = (int)pl->ops.size();
  while (pl->split > 0 && pl->ops[pl->split - 1].dev.n_nodes <= 8) pl->split--;
  // ---- fused front section (pyfaceanalysis_b200/front.py): layers 0-2 in one kernel ----
  {
    const char* env = getenv("HGSFA_FRONT");
    if (front_off > 0 && front_bytes > 0 && !(env && env[0] == '0') && pl->ops.size() >= 4 && pl->split >= 3) {
      Cursor fc{base + front_off, base + front_off + front_bytes};
      const uint8_t* fh_raw = fc.take<uint8_t>(8 + 16 * 8 + 12 * 8);
      if (!fh_raw || std::memcmp(fh_raw, "HGSFAFR1", 8) != 0) return plan_fail(pl, "bad fused-front section");
      int64_t fh[16];
      double ff[12];
      std::memcpy(fh, fh_raw + 8, sizeof(fh));
      std::memcpy(ff, fh_raw + 8 + sizeof(fh), sizeof(ff));
      FrontDev& f = pl->front_dev;
      f.n_sub = (int)fh[0]; f.img_w = (int)fh[1]; pl->front_img_h = (int)fh[2];
      pl->front_np1 = (int)fh[3]; pl->front_np2 = (int)fh[4];
      f.nn0 = (int)fh[5]; f.nn1 = (int)fh[6]; f.nn2 = (int)fh[7];
      const int nv0 = (int)fh[8], nv1 = (int)fh[9];
      f.nv_out = (int)fh[10]; f.nch1 = (int)fh[11]; f.nch2 = (int)fh[12];
      f.out_dim = (int)fh[13]; f.sub_bytes = (int)fh[14];
      f.in_dim = (int)pl->input_dim;
      f.s0 = float(ff[0]); f.s1 = float(ff[1]); f.s2 = float(ff[2]);
      f.clo0 = float(ff[3]); f.clo1 = float(ff[4]); f.clo2 = float(ff[5]);
      f.chi0 = float(ff[6]); f.chi1 = float(ff[7]); f.chi2 = float(ff[8]);
      f.p0 = float(ff[9]); f.p1 = float(ff[10]); f.p2 = float(ff[11]);
      const int np1 = pl->front_np1, np2 = pl->front_np2;
      const bool fsane =
          f.n_sub > 0 && f.n_sub % 2 == 0 && f.n_sub <= 65534 && f.img_w >= 16 && f.img_w % 16 == 0 && pl->front_img_h >= 8 &&
          int64_t(f.img_w) * pl->front_img_h == pl->input_dim && fh[15] == FR_HEAD && f.nn0 == 16 && (f.nn1 == 16 || f.nn1 == 32) &&
          (f.nn2 == 16 || f.nn2 == 32) && np1 >= 8 && np1 <= 16 && np2 >= 8 && np2 <= 32 && np1 % 8 == 0 && np2 % 8 == 0 &&
          nv0 >= 1 && nv0 <= np1 && nv1 >= 1 && nv1 <= np2 && np2 <= f.nn1 && f.nv_out >= 1 && f.nv_out <= f.nn2 &&
          f.nch1 == 4 * np1 / 32 && f.nch2 == 4 * np2 / 32 && f.out_dim == pl->ops[2].dev.out_dim &&
          f.sub_bytes == 4 * (FR_HEAD + f.nn0 * 128) + 2 * f.nch1 * (FR_HEAD + f.nn1 * 128) + f.nch2 * (FR_HEAD + f.nn2 * 128) &&
          (2 * np1 + 2 * np1) * 4 <= FR_HEAD && (4 * np2 + 32) * 4 <= FR_HEAD && pl->ops[0].dev.n_nodes == 4 * f.n_sub &&
          pl->ops[1].dev.n_nodes == 2 * f.n_sub && pl->ops[2].dev.n_nodes == f.n_sub;
      if (!fsane) return plan_fail(pl, "inconsistent fused-front header (n_sub=%d np=(%d, %d) nn=(%d, %d, %d))", f.n_sub, np1, np2, f.nn0, f.nn1, f.nn2);
      const int32_t* pair_xy = fc.take<int32_t>(size_t(f.n_sub));          // n_sub / 2 pairs of (x, y)
      const int32_t* l0_off = fc.take<int32_t>(size_t(f.n_sub) * 4);
      const int32_t* out_col = fc.take<int32_t>(size_t(f.n_sub));
      const uint8_t* wimg = fc.take<uint8_t>(size_t(f.n_sub) * f.sub_bytes);
      if (!fc.ok) return plan_fail(pl, "truncated fused-front section");
      for (int k = 0; k < f.n_sub; ++k) {
        const int px = pair_xy[(k / 2) * 2], py = pair_xy[(k / 2) * 2 + 1];
        bool good = px >= 0 && py >= 0 && px % 16 == 0 && px + 16 <= f.img_w && py + 8 <= pl->front_img_h && out_col[k] >= 0 &&
                    out_col[k] + f.nv_out <= f.out_dim;
        for (int i = 0; i < 4 && good; ++i) {
          const int dy = l0_off[k * 4 + i] & 0xff, dx = l0_off[k * 4 + i] >> 8;
          good = dy >= 0 && dy <= 4 && dx >= 0 && dx <= 12 && dx % 4 == 0;
        }
        if (!good) return plan_fail(pl, "fused front: subtree %d has an invalid pixel box or output column", k);
      }
      f.pair_xy = reinterpret_cast<const int2*>(dev_ptr(pair_xy));
      f.l0_off = reinterpret_cast<const int*>(dev_ptr(l0_off));
      f.out_col = reinterpret_cast<const int*>(dev_ptr(out_col));
      f.wimg = dev_ptr(wimg);
      if ((reinterpret_cast<uintptr_t>(f.wimg) & 15) != 0) return plan_fail(pl, "fused front: weight chunks are not 16-byte aligned");
      if (front_set_attributes(np1, np2)) { std::string msg = last_error_ref(); return plan_fail(pl, "%s", msg.c_str()); }
      pl->front_ok = true;
    }
  }
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
  pl->params.release(); pl->tin.release(); pl->mid.release(); pl->fbuf.release();
  for (auto& b : pl->tc_bufs) b.release();
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

#ifdef HGSFA_TC_TRACE
extern "C" int hgsfa_debug_trace(unsigned long long* out, int n) {
  HG_CUDA(cudaDeviceSynchronize());
  HG_CUDA(cudaMemcpyFromSymbol(out, tc_trace, sizeof(unsigned long long) * size_t(n)));
  return 0;
}
#endif

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
  // tile indices ride on gridDim.y (layout kernels) and gridDim.x / twc (layer kernels): a chunk holds at most
  // 65 532 tiles (8.39 M windows); larger requests are clamped, the chunk loop covers the rest
  const int64_t max_chunk = int64_t(65532) * TILE;
  front_chunk = std::min(front_chunk, max_chunk);
  back_chunk = std::min(back_chunk, max_chunk);
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
  PtrDeviceGuard guard(d_dst);
  HG_CHECK(guard.ok, "hgsfa_tile_windows_device: cannot select device %d", guard.device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t max_piece = int64_t(65532) * TILE;       // gridDim.y carries the tile index
  if (n > max_piece) {
    const size_t sel = dtype_size(dtype), del = dtype_size(dst_dtype);
    for (int64_t w0 = 0; w0 < n; w0 += max_piece)
      if (hgsfa_tile_windows_device(static_cast<const uint8_t*>(d_src) + size_t(w0) * ld * sel, dtype, std::min(max_piece, n - w0), dim,
                                    ld, static_cast<uint8_t*>(d_dst) + size_t(w0) * dim * del, dst_dtype, stream))
        return 1;
    return 0;
  }
  dim3 grid((unsigned)ceil_div(dim, 64), (unsigned)ceil_div(n, TILE));
  if (dtype == HGSFA_U8 && dst_dtype == HGSFA_U8 && ld % 16 == 0 && (reinterpret_cast<uintptr_t>(d_src) & 15) == 0) {
    dim3 g8((unsigned)ceil_div(dim, 128), (unsigned)ceil_div(n, TILE));
    tile_windows_u8_kernel<<<g8, 256, 0, st>>>((const uint8_t*)d_src, n, dim, ld, (uint8_t*)d_dst);
  } else if (dtype == HGSFA_U8 && dst_dtype == HGSFA_U8)
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
    OpHost& op = pl->ops[o];
    float* dst = (o == o1 - 1) ? final_out : static_cast<float*>(pp[(o - o0) & 1].p);
    hgsfa_plan_s::Stamp stamp{o, nullptr, nullptr};
    if (pl->profile && cudaEventCreate(&stamp.e0) == cudaSuccess && cudaEventCreate(&stamp.e1) == cudaSuccess)
      cudaEventRecord(stamp.e0, st);
    int rc;
    if (op.back && !cur_u8) {
      rc = back_launch(op.bkd, pl->sm_count, static_cast<const float*>(cur), dst, ntiles, st);
      pl->launches++;
    } else {
      rc = op.tc ? (cur_u8 ? launch_layer_tc<uint8_t>(pl, op, cur, dst, ntiles, st) : launch_layer_tc<float>(pl, op, cur, dst, ntiles, st))
                 : (cur_u8 ? launch_layer<uint8_t>(pl, op, cur, dst, ntiles, st) : launch_layer<float>(pl, op, cur, dst, ntiles, st));
    }
    if (rc) return rc;
    if (pl->profile && stamp.e1) {
      cudaEventRecord(stamp.e1, st);
      pl->stamps.push_back(stamp);
    }
    cur = dst;
    cur_u8 = false;
  }
  return 0;
}

}  // namespace

extern "C" int hgsfa_plan_profile(hgsfa_plan_t pl, int enable) {
  HG_CHECK(pl, "hgsfa_plan_profile: null plan");
  pl->profile = enable != 0;
  pl->op_ms.assign(pl->ops.size(), 0.0);
  return 0;
}

extern "C" int hgsfa_plan_op_stats(hgsfa_plan_t pl, int64_t capacity, double* ms, int32_t* engine, double* alg_flops,
                                   double* exe_flops) {
  HG_CHECK(pl, "hgsfa_plan_op_stats: null plan");
  HG_CHECK(capacity >= (int64_t)pl->ops.size(), "hgsfa_plan_op_stats: capacity %lld < %zu ops", (long long)capacity, pl->ops.size());
  DeviceGuard guard(pl->device);
  HG_CUDA(cudaDeviceSynchronize());
  pl->op_ms.resize(pl->ops.size(), 0.0);
  for (auto& s : pl->stamps) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, s.e0, s.e1) == cudaSuccess) pl->op_ms[s.op] += t;
    cudaEventDestroy(s.e0);
    cudaEventDestroy(s.e1);
  }
  pl->stamps.clear();
  for (size_t o = 0; o < pl->ops.size(); ++o) {
    if (ms) ms[o] = pl->op_ms[o];
    // 2: fused front (uint8 inputs; time booked on op 0), 3: single-layer FP16-split kernel (float inputs)
    if (engine) engine[o] = (pl->front_ok && o < 3) ? 2 : ((pl->ops[o].back || (pl->ops[o].tc && pl->ops[o].tcd.f16)) ? 3 : (pl->ops[o].tc ? 1 : 0));
    if (alg_flops) alg_flops[o] = double(pl->ops[o].alg_flops);
    if (exe_flops) exe_flops[o] = double(pl->ops[o].exe_flops);
  }
  return 0;
}

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
  // plain CUDA semantics: NULL is the (legacy) default stream, so callers that time or order work
  // on stream 0 (e.g. torch's default stream) see these launches in that stream
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const int n_ops = (int)pl->ops.size();
  const int split = pl->split;
  int64_t fc = pl->front_chunk, bc = pl->back_chunk;
  const int64_t n_pad = ceil_div(n, TILE) * TILE;
  if (bc > n_pad) bc = ceil_div(n_pad, 4 * TILE) * 4 * TILE;
  if (fc > bc) fc = bc;
  if (split == 0) fc = bc;        // every op is a back op: one launch sequence per back chunk
  if (split == n_ops) bc = fc;    // no back segment: nothing to accumulate across front chunks
  bc = ceil_div(bc, fc) * fc;

  const bool in_u8 = (x_dtype == HGSFA_U8);
  const size_t in_el = in_u8 ? 1 : 4;
  // fused front: layers 0-2 in one kernel for uint8 windows (the detector's and the benchmark's input); row-major
  // windows are read in place through a tensor map when rows are 16-byte aligned, else tiled first
  const bool use_front = pl->front_ok && in_u8 && split >= 3;
  const bool front_direct = use_front && x_layout == HGSFA_ROWMAJOR && ld % 16 == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15) == 0 &&
                            front_tensor_maps_available();
  const int o_first = use_front ? 3 : 0;      // first op that runs as a per-layer launch
  int64_t maxf_front = 0, maxf_back = 0;
  for (int o = o_first; o < split; ++o) maxf_front = std::max<int64_t>(maxf_front, pl->ops[o].dev.out_dim);
  for (int o = split; o < n_ops; ++o) maxf_back = std::max<int64_t>(maxf_back, pl->ops[o].dev.out_dim);
  if (x_layout == HGSFA_ROWMAJOR && !front_direct && pl->tin.reserve(size_t(fc) * pl->input_dim * in_el)) return 1;
  if (use_front && split > 3 && pl->fbuf.reserve(size_t(fc) * pl->ops[2].dev.out_dim * 4)) return 1;
  if (split > o_first)
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
      } else if (front_direct) {
        xin = xb + size_t(f0) * ld;
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
                                    : static_cast<float*>(pl->front[(split - 1 - o_first) & 1].p);
      if (use_front) {
        float* fdst = split > 3 ? static_cast<float*>(pl->fbuf.p) : fout;
        hgsfa_plan_s::Stamp stamp{0, nullptr, nullptr};
        if (pl->profile && cudaEventCreate(&stamp.e0) == cudaSuccess && cudaEventCreate(&stamp.e1) == cudaSuccess)
          cudaEventRecord(stamp.e0, st);
        if (front_launch(pl->front_dev, pl->front_np1, pl->front_np2, pl->front_img_h, pl->sm_count,
                         front_direct ? FR_IN_ROWMAJOR : FR_IN_TILED, static_cast<const uint8_t*>(xin), ld, fn, fdst, st))
          return 1;
        pl->launches++;
        if (pl->profile && stamp.e1) {
          cudaEventRecord(stamp.e1, st);
          pl->stamps.push_back(stamp);
        }
        if (split > 3 && run_ops(pl, 3, split, fdst, false, fout, pl->front, f_tiles, st)) return 1;
      } else if (run_ops(pl, 0, split, xin, in_u8, fout, pl->front, f_tiles, st)) return 1;
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

// tcgen05 probe, round 2: is a 2-piece FP16 split (kind::f16, K = 16 per instruction) a cheaper route to
// FP32-grade contractions than the 3xTF32 split (kind::tf32, K = 8 per instruction)?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_probe2 tools/tc_probe2.cu
//   timeout 120 build/tc_probe2            (on the B200)
//
// 1. correctness + accuracy: D[128 x N] = A[128 x K] * B[K x N], A = Ahi + Alo and B = Bhi + Blo as FP16 pairs,
//    D = Ahi Bhi + Ahi Blo + Alo Bhi accumulated in FP32, against a float64 product.  A is written to tensor
//    memory by the threads that own the rows (two K elements per 32-bit column), B is the canonical K-major
//    no-swizzle image ([k/8][n/8][n%8][k%8] halves).
// 2. issue rate of M128 x N x K16 kind::f16 MMAs with A in tensor memory, back to back, N = 16..128.
// 3. how long one mbarrier.try_wait blocks in hardware (iterations of a try_wait loop over a known delay),
//    with and without a suspend-time hint.
//
// All waits are bounded; a time-out sets a flag instead of hanging the device.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t phase, int* flag) {
  for (long it = 0; it < (1L << 24); ++it) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
    if (ok) return true;
  }
  if (flag) atomicExch(flag, 1);
  return false;
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// D = F32 (bits 4-5 = 1), A = B = F16 (format 0), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
// half index of B(n, k) in the canonical K-major no-swizzle image of 16-bit elements: [k/8][n/8][n%8][k%8]
__host__ __device__ inline int b_off16(int n, int k, int N) { return (((k >> 3) * (N >> 3) + (n >> 3)) * 8 + (n & 7)) * 8 + (k & 7); }

// hi / lo FP16 pieces of two values, packed (low half = first value)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------------------------------------
// mode 0: D = fp16(A) fp16(B);  mode 1: three products of the 2-piece split
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const __half* __restrict__ Bimg_hi,
                                                    const __half* __restrict__ Bimg_lo, float* __restrict__ D,
                                                    int N, int K, int mode, int* flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16);
  __half* Bhi = reinterpret_cast<__half*>(smem + 128);
  __half* Blo = Bhi + K * N;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < K * N; i += 128) { Bhi[i] = Bimg_hi[i]; Blo[i] = Bimg_lo[i]; }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  // columns: D at [0, N), A hi at [256, 256 + K/2), A lo at [384, 384 + K/2)
  for (int k0 = 0; k0 < K; k0 += 16) {
    uint32_t hi[8], lo[8];
    for (int j = 0; j < 8; ++j) split2(A[(size_t)tid * K + k0 + 2 * j], A[(size_t)tid * K + k0 + 2 * j + 1], hi[j], lo[j]);
    tmem_st8(lane_base + 256 + k0 / 2, hi);
    tmem_st8(lane_base + 384 + k0 / 2, lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128, sbo = 128;
    uint32_t acc = 0;
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint64_t dhi = make_desc(smem_u32(Bhi) + (k0 / 8) * lbo, lbo, sbo);
      uint64_t dlo = make_desc(smem_u32(Blo) + (k0 / 8) * lbo, lbo, sbo);
      mma_f16_ts(tbase, tbase + 256 + k0 / 2, dhi, idesc, acc);
      acc = 1;
      if (mode >= 1) {
        mma_f16_ts(tbase, tbase + 256 + k0 / 2, dlo, idesc, 1);
        mma_f16_ts(tbase, tbase + 384 + k0 / 2, dhi, idesc, 1);
      }
    }
    tc_commit(smem_u32(bar));
  }
  bool ok = mbar_wait_bounded(smem_u32(bar), 0, flag);
  tc_fence_after();
  if (ok) {
    for (int n0 = 0; n0 < N; n0 += 8) {
      uint32_t v[8];
      tmem_ld8(lane_base + n0, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)tid * N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------
// issue rate: every CTA issues `iters` groups of 8 kind::f16 MMAs (K = 16 each), A in tensor memory
__global__ void __launch_bounds__(128) rate_kernel(int N, int iters, int nacc, int* flag, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16);
  __half* B = reinterpret_cast<__half*>(smem + 128);                  // 128 x N image (8 MMAs of K = 16)
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 128 * N; i += 128) B[i] = __float2half(0.001f * (float)(i % 13));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase_v = *tmem_slot;
  const uint32_t lane_base = tbase_v + ((uint32_t)(warp * 32) << 16);
  for (int k0 = 0; k0 < 64; k0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0x38003800u + (uint32_t)((tid + j) & 15);
    tmem_st8(lane_base + 256 + k0, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, tbase_v, 0);
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128, sbo = 128;
    uint64_t db[8];
    uint32_t ta[8], td[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      db[k] = make_desc(smem_u32(B) + (2 * k) * lbo, lbo, sbo);
      ta[k] = tbase + 256 + 8 * k;
      td[k] = tbase + (uint32_t)((k % nacc) * N);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t acc = (it > 0 || k >= nacc) ? 1u : 0u;
        if (leader) mma_f16_ts(td[k], ta[k], db[k], idesc, acc);
      }
    }
    if (leader) tc_commit(smem_u32(bar));
    __syncwarp();
    mbar_wait_bounded(smem_u32(bar), 0, flag);
  }
  __syncthreads();
  tc_fence_after();
  uint32_t v[8];
  tmem_ld8(lane_base, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (v[0] == 0x12345678u) sink[tid] = __uint_as_float(v[1]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase_v), "r"(512));
}

// ---------------------------------------------------------------------------------------------------------
// try_wait: warp 1 arrives after `delay` clocks; warp 0 counts the try_wait calls it needs (hint = 0: no hint)
__global__ void trywait_kernel(long long delay, uint32_t hint, int* out) {
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 32) {
    const long long t0 = clock64();
    while (clock64() - t0 < delay) {}
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(&bar)) : "memory");
  }
  if (threadIdx.x < 32) {
    int n = 0;
    const long long t0 = clock64();
    for (; n < (1 << 22); ++n) {
      uint32_t ok;
      if (hint)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0), "r"(hint) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (ok) break;
    }
    if (threadIdx.x == 0) { out[0] = n + 1; out[1] = (int)(clock64() - t0); }
  }
}

int main() {
  int* flag;
  CK(cudaMalloc(&flag, 4));
  CK(cudaMemset(flag, 0, 4));
  int bad = 0;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  for (int mode = 0; mode < 2; ++mode)
    for (double scale : {1.0, 200.0, 0.01})
      for (int N : {16, 32, 64}) {
        const int K = 48;
        std::vector<float> A(128 * K), B(K * N), D(128 * N);
        std::vector<__half> Bhi(K * N), Blo(K * N);
        srand(1234 + N);
        for (auto& v : A) v = (float)(((double)rand() / RAND_MAX * 2.0 - 1.0) * scale);
        for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (int k = 0; k < K; ++k)
          for (int n = 0; n < N; ++n) {
            const float w = B[k * N + n];
            const __half h = __float2half_rn(w);
            Bhi[b_off16(n, k, N)] = h;
            Blo[b_off16(n, k, N)] = __float2half_rn(w - __half2float(h));
          }
        float *dA, *dD;
        __half *dBh, *dBl;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 2)); CK(cudaMalloc(&dBl, B.size() * 2)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBh, Bhi.data(), B.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBl, Blo.data(), B.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dD, 0, D.size() * 4));
        probe_kernel<<<1, 128, 128 + 2 * K * N * 2>>>(dA, dBh, dBl, dD, N, K, mode, flag);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        double emax = 0, eref = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < N; ++n) {
            double s = 0, st = 0;
            for (int k = 0; k < K; ++k) {
              s += (double)A[m * K + k] * B[k * N + n];
              st += (double)__half2float(__float2half_rn(A[m * K + k])) * __half2float(__float2half_rn(B[k * N + n]));
            }
            const double ref = mode ? s : st;
            emax = fmax(emax, fabs(D[m * N + n] - ref));
            eref = fmax(eref, fabs(ref));
          }
        int f;
        CK(cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost));
        const bool okv = emax <= (mode ? 4e-6 : 1e-5) * fmax(1.0, eref) && !f;
        printf("f16 probe mode=%d scale=%g N=%d K=%d  max|err|=%.3g (max|ref|=%.3g, rel %.2e) timeout=%d %s\n", mode, scale, N, K, emax,
               eref, emax / eref, f, okv ? "OK" : "MISMATCH");
        if (!okv) bad = 1;
        CK(cudaFree(dA)); CK(cudaFree(dBh)); CK(cudaFree(dBl)); CK(cudaFree(dD));
        if (f) { printf("time-out: stopping\n"); return 1; }
      }
  if (bad) printf("f16 probe: MISMATCH somewhere (rates follow anyway)\n");

  float* sink;
  CK(cudaMalloc(&sink, 4096));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int nacc : {1, 2})
    for (int N : {16, 32, 64, 128}) {
      const int iters = 2000;
      const size_t sm = 128 + (size_t)128 * N * 2;
      rate_kernel<<<prop.multiProcessorCount, 128, sm>>>(N, 10, nacc, flag, sink);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      rate_kernel<<<prop.multiProcessorCount, 128, sm>>>(N, iters, nacc, flag, sink);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      int f;
      CK(cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost));
      const double n_mma = (double)iters * 8;
      printf("rate f16 TS nacc=%d N=%3d: %.1f ns per MMA (M128 K16), %.1f TFLOP/s, back to back; timeout=%d\n", nacc, N,
             ms * 1e6 / n_mma, n_mma * 2.0 * 128 * N * 16 * prop.multiProcessorCount / (ms * 1e-3) / 1e12, f);
      if (f) return 1;
    }

  int* out;
  CK(cudaMalloc(&out, 8));
  for (uint32_t hint : {0u, 1000u, 100000u, 10000000u})
    for (long long delay : {2000LL, 20000LL, 200000LL, 2000000LL}) {
      trywait_kernel<<<1, 64>>>(delay, hint, out);
      CK(cudaDeviceSynchronize());
      int h[2];
      CK(cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost));
      printf("try_wait hint=%u ns delay=%lld clk: %d calls, %d clk waited (%.0f clk per call)\n", hint, delay, h[0], h[1],
             (double)h[1] / h[0]);
    }
  return bad;
}

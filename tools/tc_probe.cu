// tcgen05 probe for the HiGSFA layer contraction (round-1 feasibility study, profiles/README_r01.md).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/tc_probe tools/tc_probe.cu
//   timeout 120 build/tc_probe            (on the B200)
//
// 1. correctness: D[128 x N] = A[128 x K] * B[K x N] with kind::tf32, A written to tensor memory by the
//    threads that own the rows (thread = window, tcgen05.st 32x32b), B in shared memory in the canonical
//    K-major no-swizzle form.  Once with plain TF32 operands, once as the 3xTF32 split
//    (Ahi*Bhi + Ahi*Blo + Alo*Bhi) compared with a float64 product.
// 2. throughput: back-to-back MMAs of the layer shapes (M=128, N=16..256, K=8 per instruction) on every SM.
//
// All waits are bounded; a time-out sets a flag instead of hanging the device.
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
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  return d;                 // base offset 0, layout type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :: "r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
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

// byte offset of B(n, k) in the canonical K-major no-swizzle image: [k/4][n/8][n%8][k%4] floats
__host__ __device__ inline int b_off(int n, int k, int N) { return (((k >> 2) * (N >> 3) + (n >> 3)) * 8 + (n & 7)) * 4 + (k & 3); }

// ---------------------------------------------------------------------------------------------------------
// correctness kernel: one CTA of 128 threads.  mode 0: D = tf32(A) tf32(B);  mode 1: 3xTF32 split.
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ Bimg_hi,
                                                    const float* __restrict__ Bimg_lo, float* __restrict__ D,
                                                    int N, int K, int mode, int* flag) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);         // 16 bytes
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16);
  float* Bhi = reinterpret_cast<float*>(smem + 128);
  float* Blo = Bhi + K * N;
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
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_slot;
  const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
  // columns: D at [0, N), A hi at [256, 256+K), A lo at [384, 384+K)   (K <= 128)
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t hi[8], lo[8];
    for (int j = 0; j < 8; ++j) {
      float a = A[(size_t)tid * K + k0 + j];
      uint32_t h;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
      if (mode == 2) {       // does the MMA truncate its FP32 containers?  hi = raw bits, lo = a - trunc(a)
        hi[j] = __float_as_uint(a);
        lo[j] = __float_as_uint(a - __uint_as_float(__float_as_uint(a) & 0xffffe000u));
      } else {
        hi[j] = h;
        lo[j] = __float_as_uint(a - __uint_as_float(h));
      }
    }
    tmem_st8(lane_base + 256 + k0, hi);
    tmem_st8(lane_base + 384 + k0, lo);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128, sbo = 128;
    uint32_t acc = 0;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint64_t dhi = make_desc(smem_u32(Bhi) + (k0 / 4) * lbo, lbo, sbo);
      uint64_t dlo = make_desc(smem_u32(Blo) + (k0 / 4) * lbo, lbo, sbo);
      mma_ts(tbase, tbase + 256 + k0, dhi, idesc, acc);
      acc = 1;
      if (mode >= 1) {
        mma_ts(tbase, tbase + 256 + k0, dlo, idesc, 1);
        mma_ts(tbase, tbase + 384 + k0, dhi, idesc, 1);
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
// throughput kernel: every CTA issues `iters` groups of `kk` MMAs (K = 8 each) and waits for the commit.
// ts = 1: A from tensor memory; ts = 0: A from shared memory (K-major image, 128 rows).
template <int NACC>
__global__ void __launch_bounds__(128) rate_kernel(int N, int kk, int iters, int ts, int waitmode, int nacc, int* flag, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16);
  float* B = reinterpret_cast<float*>(smem + 128);                  // kk*8 x N image
  float* Asm = B + kk * 8 * N;                                      // kk*8 x 128 image
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(smem_u32(bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kk * 8 * N; i += 128) B[i] = 0.001f * (float)(i % 13);
  for (int i = tid; i < kk * 8 * 128; i += 128) Asm[i] = 0.002f * (float)(i % 7);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase_v = *tmem_slot;
  const uint32_t lane_base = tbase_v + ((uint32_t)(warp * 32) << 16);
  for (int k0 = 0; k0 < kk * 8; k0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(0.5f + 0.001f * (float)((tid + j) & 15));
    tmem_st8(lane_base + 256 + k0, v);
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {       // whole warp, one elected lane issues: operands stay in uniform registers
    tc_fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, tbase_v, 0);
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc(N);
    const uint32_t lbo = (uint32_t)(N / 8) * 128, sbo = 128, albo = 16 * 128;
    uint32_t phase = 0;
    uint64_t db[8], da[8];
    uint32_t ta[8], td[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {       // kk == 8: descriptors precomputed, issue loop fully unrolled
      db[k] = make_desc(smem_u32(B) + (2 * k) * lbo, lbo, sbo);
      da[k] = make_desc(smem_u32(Asm) + (2 * k) * albo, albo, sbo);
      ta[k] = tbase + 256 + 8 * k;
      td[k] = tbase + (uint32_t)((k % NACC) * N);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t acc = (it > 0 || k >= NACC) ? 1u : 0u;
        if (leader) {
          if (ts) mma_ts(td[k], ta[k], db[k], idesc, acc);
          else mma_ss(td[k], da[k], db[k], idesc, acc);
        }
      }
      if (waitmode || it == iters - 1) {
        if (leader) tc_commit(smem_u32(bar));
        __syncwarp();
        if (!mbar_wait_bounded(smem_u32(bar), phase, flag)) break;
        phase ^= 1;
      }
    }
  }
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tbase_v;
  uint32_t v[8];
  tmem_ld8(lane_base, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (v[0] == 0x12345678u) sink[tid] = __uint_as_float(v[1]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"(512));
}

static float tf32_round(float x) {     // round to nearest, ties away (cvt.rna)
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x1000u;
  u &= 0xffffe000u;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

int main() {
  int* flag;
  CK(cudaMalloc(&flag, 4));
  CK(cudaMemset(flag, 0, 4));
  int bad = 0;
  for (int mode = 0; mode < 3; ++mode) {
    for (int N : {16, 32, 48, 64}) {
      const int K = 40;
      std::vector<float> A(128 * K), B(K * N), Bhi(K * N), Blo(K * N), D(128 * N);
      srand(1234 + N);
      for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
      for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
      for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
          float w = B[k * N + n], h = tf32_round(w);
          Bhi[b_off(n, k, N)] = h;
          Blo[b_off(n, k, N)] = tf32_round(w - h);
        }
      float *dA, *dBh, *dBl, *dD;
      CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 4)); CK(cudaMalloc(&dBl, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
      CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dBh, Bhi.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dBl, Blo.data(), B.size() * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0, D.size() * 4));
      size_t sm = 128 + 2 * K * N * 4;
      CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      probe_kernel<<<1, 128, sm>>>(dA, dBh, dBl, dD, N, K, mode, flag);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
      double emax = 0, eref = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
          double s = 0, st = 0;
          for (int k = 0; k < K; ++k) {
            s += (double)A[m * K + k] * B[k * N + n];
            st += (double)tf32_round(A[m * K + k]) * tf32_round(B[k * N + n]);
          }
          double ref = mode ? s : st;
          emax = fmax(emax, fabs(D[m * N + n] - ref));
          eref = fmax(eref, fabs(ref));
        }
      int f;
      CK(cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost));
      const double tol = mode ? 2e-5 : 2e-5;   // mode 0: vs tf32-rounded product (hardware may truncate, fp32 accumulation)
      printf("probe mode=%d N=%d K=%d  max|err|=%.3g (max|ref|=%.3g) timeout=%d %s\n", mode, N, K, emax, eref, f,
             (emax <= tol * fmax(1.0, eref) * (mode ? 1 : 50) && !f) ? "OK" : "MISMATCH");
      if (!(emax <= tol * fmax(1.0, eref) * (mode ? 1 : 50)) || f) bad = 1;
      CK(cudaFree(dA)); CK(cudaFree(dBh)); CK(cudaFree(dBl)); CK(cudaFree(dD));
      if (f) { printf("time-out: stopping\n"); return 1; }
    }
  }
  if (bad) { printf("probe FAILED; skipping the rate test\n"); return 1; }

  float* sink;
  CK(cudaMalloc(&sink, 4096));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(rate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int waitmode = 0; waitmode < 2; ++waitmode)
  for (int ts = 1; ts >= 0; --ts)
  for (int nacc : {1, 2, 4, 8})
    for (int N : {16, 32, 64, 128, 256}) {
      if (nacc * N > 256 || (waitmode && nacc != 4 && nacc != 1) || (!ts && nacc == 2)) continue;
      const int kk = 8, iters = 2000;
      size_t sm = 128 + (size_t)kk * 8 * N * 4 + (size_t)kk * 8 * 128 * 4;
      auto launch = [&](int its) {
        if (nacc == 1) rate_kernel<1><<<prop.multiProcessorCount, 128, sm>>>(N, kk, its, ts, waitmode, nacc, flag, sink);
        else if (nacc == 2) rate_kernel<2><<<prop.multiProcessorCount, 128, sm>>>(N, kk, its, ts, waitmode, nacc, flag, sink);
        else if (nacc == 4) rate_kernel<4><<<prop.multiProcessorCount, 128, sm>>>(N, kk, its, ts, waitmode, nacc, flag, sink);
        else rate_kernel<8><<<prop.multiProcessorCount, 128, sm>>>(N, kk, its, ts, waitmode, nacc, flag, sink);
      };
      launch(10);
      CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0));
      launch(iters);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      int f;
      CK(cudaMemcpy(&f, flag, 4, cudaMemcpyDeviceToHost));
      double n_mma = (double)iters * kk;
      double flops = n_mma * 2.0 * 128 * N * 8 * prop.multiProcessorCount;
      printf("rate %s nacc=%d N=%3d: %.1f ns per MMA (M128 K8), %.1f TFLOP/s tf32, %s; timeout=%d\n",
             ts ? "TS" : "SS", nacc, N, ms * 1e6 / n_mma, flops / (ms * 1e-3) / 1e12, waitmode ? "commit+wait every 8 MMAs" : "back to back", f);
      if (f) return 1;
    }
  return 0;
}

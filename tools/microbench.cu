// Pipe-throughput micro-benchmarks for B200 (sm_100a).
//
// Purpose: measure the denominators the fused HiGSFA layer kernels are judged
// against (BASELINE.md section 2 asks for a measured FP32 FFMA peak) and decide
// which instruction forms the inner loops should use:
//   ffma      : 3-register FFMA, 16 independent accumulators per thread
//   ffma2     : packed fma.rn.f32x2 (two FMAs per lane per instruction)
//   ffma_lds  : FFMA whose B operand is a broadcast LDS.128 from shared memory
//               (the weight-broadcast pattern of the layer kernels)
//   mufu      : lg2.approx + ex2.approx pairs (the |x|^0.8 expansion)
//   mma_tf32  : legacy mma.sync.m16n8k8 tf32 (for a 3xTF32 evaluation)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench tools/microbench.cu
// Run  : build/microbench        (prints one JSON object)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;

__global__ void __launch_bounds__(256) k_ffma(float* out, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void ffma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(d) : "l"(a), "l"(b));
}

__global__ void __launch_bounds__(256) k_ffma2(float* out, float a, float b) {
  unsigned long long acc[16];
  unsigned long long a2, b2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(a2) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b));
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = threadIdx.x * 0.001f + i;
    asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(v));
  }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) ffma2(acc[i], a2, b2);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// 4 window rows x 16 outputs register tile; weights broadcast from smem (LDS.128),
// activations read per-lane from smem (conflict-free): the layer-kernel inner loop.
__global__ void __launch_bounds__(128) k_ffma_lds(float* out, int K) {
  extern __shared__ float sm[];
  float* W = sm;                 // [K][16]
  float* A = sm + 64 * 16;       // [K][4][128]
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) W[i] = 0.001f * (i % 37);
  for (int i = threadIdx.x; i < 64 * 4 * 128; i += blockDim.x) A[i] = 0.002f * (i % 51);
  __syncthreads();
  float acc[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[r][i] = 0.f;
  for (int it = 0; it < ITERS / 64; ++it) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      float a0 = A[(k * 4 + 0) * 128 + threadIdx.x];
      float a1 = A[(k * 4 + 1) * 128 + threadIdx.x];
      float a2 = A[(k * 4 + 2) * 128 + threadIdx.x];
      float a3 = A[(k * 4 + 3) * 128 + threadIdx.x];
      const float4* w4 = reinterpret_cast<const float4*>(W + k * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 w = w4[q];
        acc[0][4*q+0] = fmaf(a0, w.x, acc[0][4*q+0]); acc[0][4*q+1] = fmaf(a0, w.y, acc[0][4*q+1]);
        acc[0][4*q+2] = fmaf(a0, w.z, acc[0][4*q+2]); acc[0][4*q+3] = fmaf(a0, w.w, acc[0][4*q+3]);
        acc[1][4*q+0] = fmaf(a1, w.x, acc[1][4*q+0]); acc[1][4*q+1] = fmaf(a1, w.y, acc[1][4*q+1]);
        acc[1][4*q+2] = fmaf(a1, w.z, acc[1][4*q+2]); acc[1][4*q+3] = fmaf(a1, w.w, acc[1][4*q+3]);
        acc[2][4*q+0] = fmaf(a2, w.x, acc[2][4*q+0]); acc[2][4*q+1] = fmaf(a2, w.y, acc[2][4*q+1]);
        acc[2][4*q+2] = fmaf(a2, w.z, acc[2][4*q+2]); acc[2][4*q+3] = fmaf(a2, w.w, acc[2][4*q+3]);
        acc[3][4*q+0] = fmaf(a3, w.x, acc[3][4*q+0]); acc[3][4*q+1] = fmaf(a3, w.y, acc[3][4*q+1]);
        acc[3][4*q+2] = fmaf(a3, w.z, acc[3][4*q+2]); acc[3][4*q+3] = fmaf(a3, w.w, acc[3][4*q+3]);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[r][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same tile with packed FFMA2: the window value is duplicated into a pair, the weight pairs come
// straight out of the LDS.128
__global__ void __launch_bounds__(128) k_ffma2_lds(float* out, int K) {
  extern __shared__ float sm[];
  float* W = sm;
  float* A = sm + 64 * 16;
  for (int i = threadIdx.x; i < 64 * 16; i += blockDim.x) W[i] = 0.001f * (i % 37);
  for (int i = threadIdx.x; i < 64 * 4 * 128; i += blockDim.x) A[i] = 0.002f * (i % 51);
  __syncthreads();
  unsigned long long acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[r][i] = 0ull;
  for (int it = 0; it < ITERS / 64; ++it) {
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      unsigned long long a2[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float a = A[(k * 4 + r) * 128 + threadIdx.x];
        asm("mov.b64 %0, {%1, %1};" : "=l"(a2[r]) : "f"(a));
      }
      const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(W + k * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        ulonglong2 w = w2[q];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r][2*q+0]) : "l"(a2[r]), "l"(w.x));
          asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r][2*q+1]) : "l"(a2[r]), "l"(w.y));
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[r][i]));
      s += lo + hi;
    }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mufu(float* out, float a) {
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 1.0f + threadIdx.x * 0.001f + i;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float l;
      asm volatile("lg2.approx.f32 %0, %1;" : "=f"(l) : "f"(acc[i]));
      l = l * a;
      asm volatile("ex2.approx.f32 %0, %1;" : "=f"(acc[i]) : "f"(l));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mma_tf32(float* out) {
  // 8 independent m16n8k8 accumulator tiles per warp
  float c[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(0.5f + threadIdx.x * 0.01f + i);
  b[0] = __float_as_uint(0.25f); b[1] = __float_as_uint(0.125f);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mma_bf16(float* out) {
  float c[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = 0x3f003f00u + threadIdx.x + i;
  b[0] = 0x3e803e80u; b[1] = 0x3e003e00u;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; ++i) launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const int blocks = sms * 8, threads = 256;
  float* out;
  CK(cudaMalloc(&out, sizeof(float) * blocks * threads));
  const double nthreads = double(blocks) * threads;
  const size_t smem_lds = (64 * 16 + 64 * 4 * 128) * sizeof(float);
  CK(cudaFuncSetAttribute(k_ffma_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lds));
  CK(cudaFuncSetAttribute(k_ffma2_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_lds));

  float t;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", prop.name, sms, prop.clockRate);
  t = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 10);
  printf(", \"ffma_tflops\": %.2f", nthreads * ITERS * 16 * 2 / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }, 10);
  printf(", \"ffma2_tflops\": %.2f", nthreads * ITERS * 16 * 4 / (t * 1e-3) / 1e12);
  {
    const int b2 = sms;  // one 256-thread CTA per SM (smem-limited)
    t = time_ms([&] { k_ffma_lds<<<b2, 128, smem_lds>>>(out, 64); }, 10);
    printf(", \"ffma_lds_4x16_tflops\": %.2f", double(b2) * 128 * (ITERS / 64) * 64 * 64 * 2 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_ffma2_lds<<<b2, 128, smem_lds>>>(out, 64); }, 10);
    printf(", \"ffma2_lds_4x16_tflops\": %.2f", double(b2) * 128 * (ITERS / 64) * 64 * 64 * 2 / (t * 1e-3) / 1e12);
  }
  t = time_ms([&] { k_mufu<<<blocks, threads>>>(out, 0.999f); }, 10);
  printf(", \"mufu_pairs_tops\": %.3f", nthreads * ITERS * 8 / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_mma_tf32<<<blocks, threads>>>(out); }, 10);
  printf(", \"mma_sync_tf32_tflops\": %.1f", double(blocks) * (threads / 32) * ITERS * 8 * (16.0 * 8 * 8 * 2) / (t * 1e-3) / 1e12);
  t = time_ms([&] { k_mma_bf16<<<blocks, threads>>>(out); }, 10);
  printf(", \"mma_sync_bf16_tflops\": %.1f", double(blocks) * (threads / 32) * ITERS * 8 * (16.0 * 8 * 16 * 2) / (t * 1e-3) / 1e12);
  printf("}\n");
  cudaFree(out);
  return 0;
}

// Instruction-throughput micro-benchmarks for the operand-split code of the tensor-core layer kernels (B200).
//
//   pack      : cvt.rn.f16x2.f32 (F2FP.PACK_AB) alone
//   unpack    : cvt.f32.f16 alone
//   split_tf32: hi = (bits + 0x1000) & ~0x1fff, lo = v - hi                                (3xTF32 operands)
//   split_f16 : h2 = pack(v0, v1); lo = pack(v0 - float(h2.x), v1 - float(h2.y))           (2-piece FP16 operands)
//   split_f16i: hi = v rounded to 11 bits with integer ops, lo = v - hi, two packs        (no unpack)
//   pow_f16   : |x - m|^0.8 (lg2, mul, ex2) followed by split_f16 (the expansion inner loop)
// Reported as G values / s over all SMs.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench2 tools/microbench2.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 2048;
constexpr int NV = 16;       // independent values per thread

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float unpack_lo(uint32_t h) {
  float f;
  asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, l;\n\t}" : "=f"(f) : "r"(h));
  return f;
}
__device__ __forceinline__ float unpack_hi(uint32_t h) {
  float f;
  asm volatile("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %1;\n\tcvt.f32.f16 %0, h;\n\t}" : "=f"(f) : "r"(h));
  return f;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_split(uint32_t* out, float seed) {
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = seed + threadIdx.x * 0.01f + i;
  uint32_t sink = 0;
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NV; i += 2) {
      float a = v[i], b = v[i + 1];
      if (MODE == 0) {                       // pack only
        sink ^= pack2(a, b);
      } else if (MODE == 1) {                // unpack only
        a = unpack_lo(__float_as_uint(a));
        b = unpack_hi(__float_as_uint(b));
        sink ^= __float_as_uint(a) ^ __float_as_uint(b);
      } else if (MODE == 2) {                // TF32 split
        const uint32_t ha = (__float_as_uint(a) + 0x1000u) & 0xffffe000u, hb = (__float_as_uint(b) + 0x1000u) & 0xffffe000u;
        sink ^= ha ^ hb ^ __float_as_uint(a - __uint_as_float(ha)) ^ __float_as_uint(b - __uint_as_float(hb));
      } else if (MODE == 3) {                // FP16 split through unpack
        const uint32_t h = pack2(a, b);
        sink ^= h ^ pack2(a - unpack_lo(h), b - unpack_hi(h));
      } else if (MODE == 4) {                // FP16 split with integer rounding of hi
        const uint32_t ha = (__float_as_uint(a) + 0x1000u) & 0xffffe000u, hb = (__float_as_uint(b) + 0x1000u) & 0xffffe000u;
        sink ^= pack2(__uint_as_float(ha), __uint_as_float(hb)) ^ pack2(a - __uint_as_float(ha), b - __uint_as_float(hb));
      } else {                               // pow + FP16 split
        const float pa = exp2f(0.8f * __log2f(fabsf(a - 0.37f))), pb = exp2f(0.8f * __log2f(fabsf(b - 0.37f)));
        const uint32_t h = pack2(pa, pb);
        sink ^= h ^ pack2(pa - unpack_lo(h), pb - unpack_hi(h));
      }
      v[i] = a + 1.0f;
      v[i + 1] = b + 1.0f;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink;
}

template <int MODE>
static int run(const char* name, uint32_t* out, int sms) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int blocks = sms * 8;
  k_split<MODE><<<blocks, 256>>>(out, 1.5f);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_split<MODE><<<blocks, 256>>>(out, 1.5f);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  const double vals = double(blocks) * 256 * ITERS * NV;
  printf("\"%s_gvals\": %.1f, ", name, vals / (ms * 1e-3) / 1e9);
  return 0;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  uint32_t* out;
  CK(cudaMalloc(&out, size_t(prop.multiProcessorCount) * 8 * 256 * 4));
  printf("{\"device\": \"%s\", \"note\": \"each loop iteration also carries one FADD per value (the +1.0 dependency)\", ", prop.name);
  if (run<0>("pack", out, prop.multiProcessorCount)) return 1;
  if (run<1>("unpack", out, prop.multiProcessorCount)) return 1;
  if (run<2>("split_tf32", out, prop.multiProcessorCount)) return 1;
  if (run<3>("split_f16", out, prop.multiProcessorCount)) return 1;
  if (run<4>("split_f16i", out, prop.multiProcessorCount)) return 1;
  if (run<5>("pow_split_f16", out, prop.multiProcessorCount)) return 1;
  printf("\"sms\": %d}\n", prop.multiProcessorCount);
  return 0;
}

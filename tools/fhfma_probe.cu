// Throughput of the sm_100a mixed-precision FMA (PTX fma.rn.f32.bf16 -> SASS FHFMA.BF16: f32 += bf16 * bf16, operands taken
// from either half of a packed register) against plain FFMA and packed HFMA2.BF16.  Standalone: nvcc -arch=sm_100a.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ float fhfma(uint32_t a, uint32_t b, float c, int h) {
  float d;
  if (h == 0) asm volatile("{.reg .b16 al,ah,bl,bh; mov.b32 {al,ah}, %1; mov.b32 {bl,bh}, %2; fma.rn.f32.bf16 %0, al, bl, %3;}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  else asm volatile("{.reg .b16 al,ah,bl,bh; mov.b32 {al,ah}, %1; mov.b32 {bl,bh}, %2; fma.rn.f32.bf16 %0, ah, bh, %3;}" : "=f"(d) : "r"(a), "r"(b), "f"(c));
  return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(const uint32_t* in, float* out, int iters) {
  uint32_t a = in[threadIdx.x], b = in[threadIdx.x + 256];
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)i;
  float fa = __uint_as_float(a), fb = __uint_as_float(b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[i]) : "f"(fa), "f"(fb));
      else if (MODE == 1) acc[i] = fhfma(a, b, acc[i], i & 1);
      else {
        uint32_t r = __float_as_uint(acc[i]);
        asm volatile("fma.rn.bf16x2 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b));
        acc[i] = __uint_as_float(r);
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, const uint32_t* in, float* out, int sms) {
  const int iters = 4096, grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(in, out, iters);
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(in, out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double n = (double)grid * 256 * iters * 16;
  printf("%-14s %8.3f ms  %8.1f G lane-instr/s  (%.1f per clk per SM at 1.965 GHz)\n", name, ms, n / ms / 1e6, n / (ms * 1e-3) / sms / 1.965e9);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  uint32_t* in; float* out;
  cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
  cudaMalloc(&out, (size_t)p.multiProcessorCount * 8 * 256 * 4);
  printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
  run<0>("FFMA", in, out, p.multiProcessorCount);
  run<1>("FHFMA.BF16", in, out, p.multiProcessorCount);
  run<2>("HFMA2.BF16", in, out, p.multiProcessorCount);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

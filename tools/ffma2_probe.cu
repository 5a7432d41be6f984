// FFMA vs FFMA2 (fma.rn.f32x2, sm_100) issue / pipe throughput probe: N independent accumulator chains per thread.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/ffma2_probe.cu -o build/ffma2_probe
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
template <int CH>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>   // CH packed chains = 2*CH scalar chains
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  unsigned long long acc[CH];
  const unsigned long long A = pack(a, a), B = pack(b, b);
#pragma unroll
  for (int i = 0; i < CH; ++i) acc[i] = pack(threadIdx.x + i, threadIdx.x - i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CH; ++i) acc[i] = ffma2(acc[i], A, B);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; ++i) { float2 v = *reinterpret_cast<float2*>(&acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// register-tiled outer product (the inner loop of a SIMT GEMM): 8 x 9 accumulators, every FMA reads three distinct
// registers (one reusable between neighbours) -- the pattern of k_gru_fwd's main loop without its shared-memory loads
__global__ void __launch_bounds__(256) k_outer(float* out, const float* in, int iters) {
  float h[8], w[9], acc[8][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = in[threadIdx.x + i];
#pragma unroll
  for (int j = 0; j < 9; ++j) w[j] = in[threadIdx.x + 8 + j];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 9; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(h[i], w[j], acc[i][j]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) s += acc[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_outer2(float* out, const float* in, int iters) {
  unsigned long long h[8], w[9], acc[8][9];
#pragma unroll
  for (int i = 0; i < 8; ++i) h[i] = pack(in[threadIdx.x + i], in[threadIdx.x + i + 32]);
#pragma unroll
  for (int j = 0; j < 9; ++j) w[j] = pack(in[threadIdx.x + 8 + j], in[threadIdx.x + 40 + j]);
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[i][j] = 0ull;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 9; ++j)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i][j] = ffma2(h[i], w[j], acc[i][j]);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 9; ++j) { float2 v = *reinterpret_cast<float2*>(&acc[i][j]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4 * 4);
  const int iters = 20000;
  for (int threads : {256, 512}) {
    const int grid = 148;
    float ms1 = timeit([&] { k_ffma<32><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); });
    float ms2 = timeit([&] { k_ffma2<16><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); });
    float ms3 = timeit([&] { k_ffma2<32><<<grid, threads>>>(out, iters, 1.0001f, 0.5f); });
    double fl = 2.0 * grid * threads * 32.0 * iters;
    printf("threads/SM %4d: FFMA x32 chains %.3f ms = %.1f TFLOP/s | FFMA2 x16 pairs %.3f ms = %.1f TFLOP/s | FFMA2 x32 pairs %.3f ms = %.1f TFLOP/s\n",
           threads, ms1, fl / ms1 / 1e9, ms2, fl / ms2 / 1e9, ms3, 2 * fl / ms3 / 1e9);
  }
  float* in; cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0, 4096 * 4);
  for (int ctas : {1, 2}) {   // 8 and 16 warps per SM
    const int it2 = 4000;
    float ms1 = timeit([&] { k_outer<<<148 * ctas, 256>>>(out, in, it2); });
    float ms2 = timeit([&] { k_outer2<<<148 * ctas, 256>>>(out, in, it2); });
    double fl = 2.0 * 148 * ctas * 256 * 72.0 * it2;
    printf("outer product 8x9, %2d warps/SM: FFMA %.3f ms = %.1f TFLOP/s | FFMA2 %.3f ms = %.1f TFLOP/s\n", 8 * ctas, ms1,
           fl / ms1 / 1e9, ms2, 2 * fl / ms2 / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

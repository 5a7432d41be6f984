// store_probe.cu -- how fast can W warps per SM stream full 512-byte rows to HBM with st.global.v4 (the finisher's
// store pattern in k_sage_tc)?  148 CTAs (one per SM), W warps each, every warp writes whole rows round-robin.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/store_probe tools/store_probe.cu && build/store_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void k_store(float4* out, long long rows_per_cta, int unroll_dummy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float4 v = make_float4((float)lane, 1.f, 2.f, (float)blockIdx.x);
  float4* base = out + (long long)blockIdx.x * rows_per_cta * 32;
  for (long long r = warp; r < rows_per_cta; r += nw) base[r * 32 + lane] = v;
}
// same, while the CTA also streams `rows_per_cta` rows IN (one extra warp group reads with ld.global.nc.v4)
__global__ void k_store_load(float4* out, const float4* in, long long rows_per_cta, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x >> 5) / 2;
  float4 v = make_float4((float)lane, 1.f, 2.f, (float)blockIdx.x);
  if (warp < nw) {
    float4* base = out + (long long)blockIdx.x * rows_per_cta * 32;
    for (long long r = warp; r < rows_per_cta; r += nw) base[r * 32 + lane] = v;
  } else {
    const float4* base = in + (long long)blockIdx.x * rows_per_cta * 32;
    float acc = 0.f;
    for (long long r = warp - nw; r < rows_per_cta; r += 4 * nw) {
      float4 a = __ldg(base + r * 32 + lane), b = __ldg(base + (r + nw) * 32 + lane);
      float4 c = __ldg(base + (r + 2 * nw) * 32 + lane), d = __ldg(base + (r + 3 * nw) * 32 + lane);
      acc += a.x + b.x + c.x + d.x;
    }
    if (acc == 12345.678f) *sink = acc;
  }
}
int main() {
  const long long rows_per_cta = 44 * 128 * 2;   // two tensors of 44 tiles x 128 rows, like one k_sage_tc CTA
  const long long total = 148 * rows_per_cta * 32;
  float4 *out, *in; float* sink;
  cudaMalloc(&out, total * 16); cudaMalloc(&in, total * 16); cudaMalloc(&sink, 4);
  cudaMemset(in, 0, total * 16);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int W : {1, 2, 4, 8, 16, 32}) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a); k_store<<<148, W * 32>>>(out, rows_per_cta, 0); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("stores only, %2d warps/SM: %.3f ms  %.1f GB/s written\n", W, best, total * 16 / best / 1e6);
  }
  for (int W : {4, 8, 16}) {
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(a); k_store_load<<<148, 2 * W * 32>>>(out, in, rows_per_cta, sink); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    printf("stores + equal loads, %2d + %2d warps/SM: %.3f ms  %.1f GB/s written, %.1f GB/s total\n", W, W, best, total * 16 / best / 1e6, 2 * total * 16 / best / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

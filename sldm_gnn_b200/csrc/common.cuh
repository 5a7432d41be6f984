// common.cuh -- shared helpers for the sm_100a SageBlock kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include "../../include/sldm_sage.h"

namespace sldm {

// ---- error plumbing (thread-local message, C return codes) -----------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);
void count_launch();  // every kernel launch of this library goes through SLDM_LAUNCH_CHECK

#define SLDM_CUDA(expr)                                                \
  do {                                                                 \
    cudaError_t _e = (expr);                                           \
    if (_e != cudaSuccess) return ::sldm::cuda_fail(_e, #expr);        \
  } while (0)

#define SLDM_LAUNCH_CHECK(name)                                        \
  do {                                                                 \
    ::sldm::count_launch();                                            \
    cudaError_t _e = cudaGetLastError();                               \
    if (_e != cudaSuccess) return ::sldm::cuda_fail(_e, name);         \
  } while (0)

#define SLDM_REQUIRE(cond, code, ...)                                  \
  do {                                                                 \
    if (!(cond)) { ::sldm::set_error(__VA_ARGS__); return (code); }    \
  } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

// every section of a workspace / CSR object starts on a 256-byte boundary
constexpr int64_t kAlignI32 = 64;
inline int64_t align_i32(int64_t n) { return round_up<int64_t>(n, kAlignI32); }
inline int64_t align_bytes(int64_t n) { return round_up<int64_t>(n, 256); }

int num_sms();  // cached per process (current device at first call)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: remember it per (call site, device)
// so that a process driving several GPUs sets it on each of them, and the driver call is not repeated per launch.
#define SLDM_OPT_IN_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                    \
    static unsigned long long _done_mask = 0ull;   /* benign race: setting the attribute twice is fine */ \
    int _dev = 0;                                                                                         \
    SLDM_CUDA(cudaGetDevice(&_dev));                                                                      \
    const unsigned long long _bit = 1ull << (_dev & 63);                                                  \
    if (!(_done_mask & _bit)) {                                                                           \
      SLDM_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      _done_mask |= _bit;                                                                                 \
    }                                                                                                     \
  } while (0)

// bump allocator over a caller-provided workspace
struct Arena {
  char* base; int64_t size; int64_t off;
  Arena(void* p, int64_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T> T* take(int64_t count) {
    int64_t bytes = align_bytes(count * (int64_t)sizeof(T));
    if (off + bytes > size) { off = size + 1; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off); off += bytes; return r;
  }
  bool ok() const { return off <= size; }
};

// ---- CSR object views -------------------------------------------------------
struct CsrLayout { int64_t off[8]; };
CsrLayout csr_layout(int64_t N, int64_t E);
inline int64_t hub_capacity(int64_t E) {
  return E / SLDM_HUB_CHUNK + E / SLDM_HUB_DEGREE + 2;
}

// ---- internal launchers shared between translation units --------------------
int segment_reduce_launch(const float* src, int64_t N, int32_t F,
                          const int32_t* rowptr, const int32_t* col,
                          const int32_t* hub_list, const int32_t* hub_count, int64_t hub_cap,
                          bool mean, const float* addend, float* out,
                          float* partials, cudaStream_t s);

int project_forward_launch(const float* agg, const float* x, int64_t N, int32_t Fin, int32_t Fout,
                           const float* W_l, const float* b_l, const float* W_r,
                           const float* ln_w, const float* ln_b, float eps, float slope,
                           float* out, float* xhat, float* rstd,
                           void* ws, int64_t ws_bytes, cudaStream_t s);
int64_t project_forward_ws_bytes(int64_t N, int32_t Fin, int32_t Fout);
// tensor-core (tcgen05, 3xTF32) variant: sage_fwd_tc.cu
bool project_forward_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* agg, const float* x,
                                 const float* out, const float* xhat);
int64_t project_forward_tc_ws_bytes(int32_t Fin, int32_t Fout);
int project_forward_tc_launch(const float* agg, const float* x, int64_t N, int32_t Fin, int32_t Fout,
                              const float* W_l, const float* b_l, const float* W_r,
                              const float* ln_w, const float* ln_b, float eps, float slope,
                              float* out, float* xhat, float* rstd, void* ws, int64_t ws_bytes, cudaStream_t s);

int layer_backward_launch(const float* dout, const float* x, const float* agg,
                          const float* xhat, const float* rstd,
                          int64_t N, int32_t Fin, int32_t Fout,
                          const int32_t* rowptr_dst,
                          const float* W_l, const float* W_r,
                          const float* ln_w, const float* ln_b, float slope,
                          bool need_dx,
                          float* dW_l, float* db_l, float* dW_r, float* dln_w, float* dln_b,
                          float* dz, float* dagg, float* dxroot,
                          void* ws, int64_t ws_bytes, cudaStream_t s, int stages = SLDM_BWD_STAGE_ALL,
                          bool bf16_feats = false);   // bf16_feats: x and agg point at bf16 rows
int64_t layer_backward_ws_bytes(int64_t N, int32_t Fin, int32_t Fout);

bool dgrad_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const float* dagg, const float* dxroot);
int64_t dgrad_tc_ws_bytes(int32_t Fin, int32_t Fout);
int dgrad_tc_launch(const float* dz, int64_t N, int32_t Fin, int32_t Fout, const float* W_l, const float* W_r,
                    const int32_t* rowptr_dst, float* dagg, float* dxroot, void* ws, int64_t ws_bytes,
                    cudaStream_t s);

bool wgrad_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const float* agg, const float* x);
int64_t wgrad_tc_ws_bytes(int32_t Fin, int32_t Fout);
int wgrad_tc_launch(const float* dz, const void* agg, const void* x, int64_t N, int32_t Fin, int32_t Fout,
                    float* part, int* nparts, cudaStream_t s, bool bf16_ops = false);
bool wgrad_bf16_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const void* agg, const void* x);
// bf16 feature storage (sage_tc.cu): agg / x / out bf16, everything else fp32
bool project_forward_bf16_eligible(int64_t N, int32_t Fin, int32_t Fout, const void* agg, const void* x,
                                   const void* out, const float* xhat);
int project_forward_bf16_launch(const void* agg, const void* x, int64_t N, int32_t Fin, int32_t Fout,
                                const float* W_l, const float* b_l, const float* W_r,
                                const float* ln_w, const float* ln_b, float eps, float slope,
                                void* out, float* xhat, float* rstd, void* ws, int64_t ws_bytes, cudaStream_t s);

// ---- device helpers -----------------------------------------------------------
// exclusive scan of one int per thread across a 256-thread CTA; returns this thread's prefix, `total` to every thread.
// smem: 256/32 + 1 ints.  (csr_build.cu has the general-width version.)
__device__ __forceinline__ int block_exclusive_scan_256(int v, int& total, int* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int w = lane < 8 ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < 8) smem[lane] = winc - w;
    if (lane == 7) smem[8] = winc;
  }
  __syncthreads();
  const int prefix = smem[warp] + inc - v;
  total = smem[8];
  __syncthreads();
  return prefix;
}

// round-to-nearest-even fp32 -> bf16 bits (NaN stays NaN through the quiet bit)
__device__ __forceinline__ uint32_t bf16_rn(float f) {
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7F800000u) == 0x7F800000u) return (u >> 16) | ((u & 0xFFFFu) ? 0x40u : 0u);
  return (u + 0x7FFFu + ((u >> 16) & 1u)) >> 16;
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) { return bf16_rn(lo) | (bf16_rn(hi) << 16); }

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4add(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}

}  // namespace sldm

// segment_bf16.cu -- the forward segment mean over bf16 feature rows ("bf16 feature storage", BASELINE configs[4]):
//   agg[i,:] = bf16( sum_{e: dst[e]=i} float(x[src[e],:]) / max(deg_i,1) )
// Same contract, same edge order and the same fp32 accumulation as segment_reduce.cu (which replaces PyG's
// index_select + scatter(reduce='mean'), src/models/blocks/sageblock.py:18); only the storage type of the rows
// changes: 2 bytes per feature in HBM instead of 4, rounded to nearest-even once, after the division.
// A group of LPR lanes owns a row, a lane holds 8 consecutive features (one 128-bit load = 8 bf16).
// Bound: HBM / L2.  Algorithmic bytes per call = E*(F*2 + 4) + 4(N+1) + N*F*2.
#include "common.cuh"
#include <algorithm>

namespace sldm {

__device__ __forceinline__ void add8_bf16(float (&acc)[8], const uint4& v) {
  // bf16 -> fp32 is a 16-bit shift: the low half of a word is the even feature, the high half the odd one
  acc[0] += __uint_as_float(v.x << 16); acc[1] += __uint_as_float(v.x & 0xFFFF0000u);
  acc[2] += __uint_as_float(v.y << 16); acc[3] += __uint_as_float(v.y & 0xFFFF0000u);
  acc[4] += __uint_as_float(v.z << 16); acc[5] += __uint_as_float(v.z & 0xFFFF0000u);
  acc[6] += __uint_as_float(v.w << 16); acc[7] += __uint_as_float(v.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 pack8_bf16(const float (&a)[8]) {
  uint4 r;
  r.x = bf16_rn(a[0]) | (bf16_rn(a[1]) << 16);
  r.y = bf16_rn(a[2]) | (bf16_rn(a[3]) << 16);
  r.z = bf16_rn(a[4]) | (bf16_rn(a[5]) << 16);
  r.w = bf16_rn(a[6]) | (bf16_rn(a[7]) << 16);
  return r;
}
__device__ __forceinline__ float ref_count_bf(int deg) {   // fp32 count of the reference: clamp(min=1), saturates at 2^24
  int c = deg < 1 ? 1 : (deg > 16777216 ? 16777216 : deg);
  return (float)c;
}
__device__ __forceinline__ void mean8(float (&acc)[8], int deg) {
  if (deg <= 1) return;
  if ((deg & (deg - 1)) == 0) {
    const float rc = __fdiv_rn(1.f, (float)deg);            // exact: deg is a power of two
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= rc;
  } else {
    const float c = ref_count_bf(deg);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = __fdiv_rn(acc[j], c);
  }
}

template <int U>
__device__ __forceinline__ void gather_piece_bf(const char* __restrict__ pb, int row_bytes, const int* w, float (&acc)[8]) {
  uint4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(pb + (int64_t)w[u] * row_bytes));
#pragma unroll
  for (int u = 0; u < U; ++u) add8_bf16(acc, v[u]);
}

// rows of degree <= SLDM_HUB_DEGREE: a warp owns RW consecutive rows, their column indices go through a per-warp
// shared-memory window (same scheme as k_segment_rows_lean)
template <int LPR, int RW>
__global__ void __launch_bounds__(256)
k_segment_rows_bf16(const uint4* __restrict__ src, int FV,          // FV = F / 8 vectors per row
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    int64_t N, uint4* __restrict__ out) {
  constexpr int RPW = 32 / LPR;
  constexpr int CH = RPW <= 4 ? 256 * RPW : 1024;
  __shared__ int s_win[8][CH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPR, lig = lane % LPR;
  int* win = s_win[warp];
  const int64_t wbase = ((int64_t)blockIdx.x * 8 + warp) * RW;
  if (wbase >= N) return;
  const int64_t wend = (wbase + RW < N) ? wbase + RW : N;
  const int64_t myrow = (wbase + lane < wend) ? wbase + lane : wend;
  const int rb = __ldg(rowptr + myrow);
  const int re = __ldg(rowptr + (myrow + 1 < wend ? myrow + 1 : wend));
  const int e_last = __shfl_sync(0xffffffffu, re, 31);
  const bool cvalid = lig < FV;
  const char* __restrict__ pb = reinterpret_cast<const char*>(src + (cvalid ? lig : 0));
  const int row_bytes = FV * 16;
  int ws = -CH - 1;
#pragma unroll 1
  for (int it = 0; it < RW / RPW; ++it) {
    const int rfirst = it * RPW;
    if (wbase + rfirst >= wend) break;
    const int ibeg = __shfl_sync(0xffffffffu, rb, rfirst);
    const int iend = __shfl_sync(0xffffffffu, re, rfirst + RPW - 1);
    if (iend > ws + CH) {
      __syncwarp();
      ws = ibeg;
#pragma unroll
      for (int q = 0; q < CH / 32; ++q) {
        const int e = ws + lane + 32 * q;
        if (e < e_last) win[lane + 32 * q] = __ldg(col + e);
      }
      __syncwarp();
    }
    const int beg = __shfl_sync(0xffffffffu, rb, rfirst + g);
    const int end = __shfl_sync(0xffffffffu, re, rfirst + g);
    const int64_t row = wbase + rfirst + g;
    const int deg = end - beg;
    if (row >= wend || deg > SLDM_HUB_DEGREE) continue;     // split rows: hub kernels below
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (beg >= ws && end <= ws + CH) {
      const int* w = win + (beg - ws);
      int rem = deg;
      for (; rem >= 4; rem -= 4, w += 4) gather_piece_bf<4>(pb, row_bytes, w, acc);
      if (rem >= 2) { gather_piece_bf<2>(pb, row_bytes, w, acc); rem -= 2; w += 2; }
      if (rem >= 1) gather_piece_bf<1>(pb, row_bytes, w, acc);
    } else {
#pragma unroll 1
      for (int k = beg; k < end; ++k) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(pb + (int64_t)__ldg(col + k) * row_bytes));
        add8_bf16(acc, v);
      }
    }
    if (cvalid) {
      mean8(acc, deg);
      out[row * FV + lig] = pack8_bf16(acc);
    }
  }
}

// hub rows: one CTA per 2048-edge piece; 256/LPR groups sum fixed sub-ranges in edge order, combined in group order
template <int LPR>
__global__ void __launch_bounds__(256)
k_segment_hub_chunks_bf16(const uint4* __restrict__ src, int FV, const int32_t* __restrict__ rowptr,
                          const int32_t* __restrict__ col, const int4* __restrict__ hub_list,
                          const int32_t* __restrict__ hub_count, int cap, float* __restrict__ partials) {
  constexpr int G = 256 / LPR;
  constexpr int PER = SLDM_HUB_CHUNK / G;
  __shared__ float sm[G][LPR * 8 + 4];
  const int n = min(*hub_count, cap);
  const int g = threadIdx.x / LPR, lig = threadIdx.x % LPR;
  const bool cvalid = lig < FV;
  const int F = FV * 8;
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    const int4 ent = hub_list[c];  // {row, chunk, nchunks, first}
    const int rbeg = __ldg(rowptr + ent.x), rend = __ldg(rowptr + ent.x + 1);
    const int cb = rbeg + ent.y * SLDM_HUB_CHUNK;
    const int ce = min(rend, cb + SLDM_HUB_CHUNK);
    const int b = min(ce, cb + g * PER), e = min(ce, b + PER);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cvalid) {
      int k = b;
      for (; k + 4 <= e; k += 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (int64_t)__ldg(col + k + u) * FV + lig);
#pragma unroll
        for (int u = 0; u < 4; ++u) add8_bf16(acc, v[u]);
      }
      for (; k < e; ++k) add8_bf16(acc, __ldg(src + (int64_t)__ldg(col + k) * FV + lig));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[g][lig * 8 + j] = acc[j];
    __syncthreads();
    for (int f = threadIdx.x; f < F; f += 256) {
      float s = sm[0][f];
#pragma unroll 4
      for (int gg = 1; gg < G; ++gg) s += sm[gg][f];
      partials[(int64_t)c * F + f] = s;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_segment_hub_finalize_bf16(const float* __restrict__ partials, int F, const int32_t* __restrict__ rowptr,
                            const int4* __restrict__ hub_list, const int32_t* __restrict__ hub_count, int cap,
                            uint16_t* __restrict__ out) {
  const int n = min(*hub_count, cap);
  const int lane = threadIdx.x & 31;
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t c = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); c < n; c += nw) {
    const int4 ent = hub_list[c];
    if (ent.y != 0) continue;
    const int64_t row = ent.x;
    const float cnt = ref_count_bf(__ldg(rowptr + row + 1) - __ldg(rowptr + row));
    for (int f = lane; f < F; f += 32) {
      float s = partials[c * F + f];
      for (int j = 1; j < ent.z; ++j) s += partials[(c + j) * F + f];
      out[row * F + f] = (uint16_t)bf16_rn(__fdiv_rn(s, cnt));
    }
  }
}

template <int LPR>
static int launch_bf16(const void* src, int64_t N, int FV, const int32_t* rowptr, const int32_t* col,
                       const int32_t* hub_list, const int32_t* hub_count, int64_t hub_cap, void* out, float* partials,
                       cudaStream_t s) {
  constexpr int RW = (LPR >= 16) ? 8 : 32;       // 256-byte rows and wider: short strips (rows in flight fit the L2)
  const int64_t grid = ceil_div<int64_t>(N, 8 * RW);
  k_segment_rows_bf16<LPR, RW><<<(unsigned)grid, 256, 0, s>>>(static_cast<const uint4*>(src), FV, rowptr, col, N,
                                                             static_cast<uint4*>(out));
  SLDM_LAUNCH_CHECK("k_segment_rows_bf16");
  if (hub_cap > 0 && hub_list != nullptr) {
    const int cap = (int)hub_cap;
    const int g1 = (int)std::min<int64_t>(hub_cap, (int64_t)num_sms() * 8);
    k_segment_hub_chunks_bf16<LPR><<<g1, 256, 0, s>>>(static_cast<const uint4*>(src), FV, rowptr, col,
                                                     reinterpret_cast<const int4*>(hub_list), hub_count, cap, partials);
    SLDM_LAUNCH_CHECK("k_segment_hub_chunks_bf16");
    const int g2 = (int)std::min<int64_t>(ceil_div<int64_t>(hub_cap, 8), (int64_t)num_sms() * 4);
    k_segment_hub_finalize_bf16<<<g2, 256, 0, s>>>(partials, FV * 8, rowptr, reinterpret_cast<const int4*>(hub_list),
                                                   hub_count, cap, static_cast<uint16_t*>(out));
    SLDM_LAUNCH_CHECK("k_segment_hub_finalize_bf16");
  }
  return SLDM_OK;
}

}  // namespace sldm

using namespace sldm;

extern "C" int sldm_segment_mean_bf16(const void* src, int64_t N, int32_t F, const int32_t* csr, int64_t E, void* out,
                                      void* workspace, int64_t workspace_bytes, sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && E >= 0 && F >= 0, SLDM_EINVAL, "sldm_segment_mean_bf16: negative size");
  SLDM_REQUIRE(csr != nullptr, SLDM_EINVAL, "sldm_segment_mean_bf16: csr is NULL");
  if (N == 0 || F == 0) return SLDM_OK;
  SLDM_REQUIRE(src != nullptr && out != nullptr, SLDM_EINVAL, "sldm_segment_mean_bf16: NULL src/out");
  SLDM_REQUIRE(F % 8 == 0 && F <= 256 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 15u) == 0, SLDM_EUNSUPPORTED,
               "sldm_segment_mean_bf16: needs F %% 8 == 0, F <= 256 and 16-byte aligned rows (F = %d)", F);
  const int64_t need = sldm_segment_workspace_bytes(N, E, F);
  SLDM_REQUIRE(workspace_bytes >= need && (workspace != nullptr || need == 0), SLDM_EWORKSPACE,
               "sldm_segment_mean_bf16: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  CsrLayout L = csr_layout(N, E);
  const int32_t* meta = csr + L.off[SLDM_CSR_META];
  const int32_t* rowptr = csr + L.off[SLDM_CSR_ROWPTR_DST];
  const int32_t* col = csr + L.off[SLDM_CSR_COL_SRC];
  const int32_t* hub = E > SLDM_HUB_DEGREE ? csr + L.off[SLDM_CSR_HUB_DST] : nullptr;
  const int64_t cap = E > SLDM_HUB_DEGREE ? hub_capacity(E) : 0;
  const int FV = F / 8;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* part = static_cast<float*>(workspace);
  if (FV <= 4)  return launch_bf16<4>(src, N, FV, rowptr, col, hub, meta, cap, out, part, s);
  if (FV <= 8)  return launch_bf16<8>(src, N, FV, rowptr, col, hub, meta, cap, out, part, s);
  if (FV <= 16) return launch_bf16<16>(src, N, FV, rowptr, col, hub, meta, cap, out, part, s);
  return launch_bf16<32>(src, N, FV, rowptr, col, hub, meta, cap, out, part, s);
}

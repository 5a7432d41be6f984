// segment_reduce.cu -- deterministic CSR segment mean / sum gather.
//
// Forward:  agg[i,:] = sum_{e: dst[e]=i} x[src[e],:] / max(deg_i,1)
//   replaces x.index_select(0, edge_index[0]) + scatter(reduce='mean') of PyG 2.7.0
//   (MessagePassing.propagate / MeanAggregation / utils/_scatter.py::scatter) as
//   reached from src/models/blocks/sageblock.py:18.  No [E,F] intermediate, no atomics.
// Backward: dx[j,:] = dxroot[j,:] + sum_{e: src[e]=j} dagg_scaled[dst[e],:]
//   (the autograd of index_select + scatter_add_), over the transpose CSR.
//
// One group of LPR lanes owns one row; every lane holds VPL 128-bit (or 32-bit
// on the unaligned path) slices of the row.  UNR source rows are in flight per
// group at a time; the adds are issued in edge order, so a row of degree
// <= SLDM_HUB_DEGREE is summed exactly like the reference's sequential CPU
// scatter_add_.  Rows above that are cut into SLDM_HUB_CHUNK-edge pieces, each
// summed by one CTA (fixed sub-ranges per group, fixed combine order) into a
// partial slot, then recombined in piece order: deterministic, no atomics.
//
// Measured (B200, batch workload F=128, profiles/r01b_*): generic kernel 0.38 ms (issue-bound, 362 warp instr/row),
// lean kernel 0.27 ms; config-4 graph 1.20 -> 0.91 ms (6.2 TB/s algorithmic).
// Bound: HBM.  Algorithmic bytes per call = E*(F*4 + 4) + 4(N+1) + N*F*4 (+N*F*4 addend).
// Round 2 (ncu, profiles/r02_ncu_full_summary.txt): the lean kernel is issue bound on the batch workload (issue active
// 77 %, 194 warp instructions per row, DRAM 54 %).  Tried and lost: 8 floats per lane with 256-bit loads (a 512-byte
// row on 16 lanes, two rows per warp at a time: half the instructions per row) -- bit-identical, but 0.267 vs 0.198 ms
// on the batch workload and 1.35 vs 0.90 ms on config 4 (44 registers, fewer resident warps, 32-byte lanes).
#include "common.cuh"
#include <algorithm>
#include <type_traits>

namespace sldm {

template <typename VT> struct Vec;
template <> struct Vec<float4> {
  static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
  static __device__ __forceinline__ float4 ld(const float4* p) { return __ldg(p); }
  static __device__ __forceinline__ void add(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
  static __device__ __forceinline__ float4 div(const float4& a, float c) {
    return make_float4(__fdiv_rn(a.x, c), __fdiv_rn(a.y, c), __fdiv_rn(a.z, c), __fdiv_rn(a.w, c));
  }
};
template <> struct Vec<float> {
  static __device__ __forceinline__ float zero() { return 0.f; }
  static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
  static __device__ __forceinline__ void add(float& a, const float& b) { a += b; }
  static __device__ __forceinline__ float div(const float& a, float c) { return __fdiv_rn(a, c); }
};

// fp32 count of the reference: scatter_add_ of ones saturates at 2^24, then clamp(min=1)
__device__ __forceinline__ float ref_count(int deg) {
  int c = deg < 1 ? 1 : (deg > 16777216 ? 16777216 : deg);
  return (float)c;
}

// acc[q] += sum over edges k in [beg,end) of src[col[k]][fv0 + lig + q*LPR], in edge order
template <typename VT, int LPR, int VPL, int UNR>
__device__ __forceinline__ void accumulate_range(const VT* __restrict__ src, int64_t FV, int fv0,
                                                 const int32_t* __restrict__ col, int beg, int end,
                                                 int lig, unsigned gmask, VT (&acc)[VPL]) {
  bool act[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) act[q] = (int64_t)fv0 + lig + q * LPR < FV;
  for (int k = beg; k < end; k += LPR) {
    const int nk = min(LPR, end - k);
    const int mycol = (lig < nk) ? __ldg(col + k + lig) : 0;
    for (int j = 0; j < nk; j += UNR) {
      VT v[UNR][VPL];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int c = __shfl_sync(gmask, mycol, j + u, LPR);
        const bool ok = (j + u) < nk;
        const VT* p = src + (int64_t)c * FV + fv0 + lig;
#pragma unroll
        for (int q = 0; q < VPL; ++q) v[u][q] = (ok && act[q]) ? Vec<VT>::ld(p + q * LPR) : Vec<VT>::zero();
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
#pragma unroll
        for (int q = 0; q < VPL; ++q) Vec<VT>::add(acc[q], v[u][q]);
      }
    }
  }
}

template <int LPR>
__device__ __forceinline__ unsigned group_mask(int lane) {
  if constexpr (LPR == 32) {
    return 0xffffffffu;
  } else {
    return ((1u << LPR) - 1u) << ((lane / LPR) * LPR);
  }
}

// A CTA owns kRowsPerCta CONSECUTIVE destination rows and its 8 warps walk them interleaved, so at any moment the
// CTA works on a small neighbourhood of the graph.  For block-diagonal batches (whole small graphs per CTA) the
// source rows then hit the SM's L1 (ld.global.nc) instead of L2; for unstructured graphs it changes nothing.
constexpr int kRowsPerCta = 256;

template <typename VT, int LPR, int VPL, int UNR>
__global__ void __launch_bounds__(256)
k_segment_rows(const VT* __restrict__ src, int64_t FV,
               const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
               int64_t N, int mean, const VT* __restrict__ addend, VT* __restrict__ out) {
  constexpr int RPW = 32 / LPR;          // rows per warp step
  constexpr int STEP = 8 * RPW;          // rows per CTA step
  const int lane = threadIdx.x & 31;
  const int lig = lane % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  const int64_t base = (int64_t)blockIdx.x * kRowsPerCta + (threadIdx.x >> 5) * RPW + lane / LPR;
  for (int off = 0; off < kRowsPerCta; off += STEP) {
    const int64_t row = base + off;
    if (row >= N) break;
    const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const int deg = end - beg;
    if (deg > SLDM_HUB_DEGREE) continue;  // split rows: k_segment_hub_*
    const float cnt = ref_count(deg);
    for (int64_t fv0 = 0; fv0 < FV; fv0 += LPR * VPL) {
      VT acc[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) acc[q] = Vec<VT>::zero();
      accumulate_range<VT, LPR, VPL, UNR>(src, FV, (int)fv0, col, beg, end, lig, gmask, acc);
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const int64_t idx = fv0 + lig + q * LPR;
        if (idx < FV) {
          VT r = acc[q];
          if (mean) r = Vec<VT>::div(r, cnt);
          if (addend) { VT a = Vec<VT>::ld(addend + row * FV + idx); Vec<VT>::add(a, r); r = a; }
          out[row * FV + idx] = r;
        }
      }
    }
  }
}

// ---- lean row kernel (128-bit path, FV <= LPR) ------------------------------------------------------------------
// The generic kernel above is issue-bound on low-degree graphs (ncu, profiles/r01b_*: 362 warp instructions per row at
// mean degree 5: padded 8-wide predicated slots, a shuffle and 64-bit index arithmetic per edge, four IEEE divisions
// per lane per row).  This one spends ~7 instructions per edge:
//   * a warp owns 32 CONSECUTIVE rows: their row pointers are two coalesced loads, their column indices one
//     contiguous range of `col` that is copied into a per-warp shared-memory window (coalesced) and then read back
//     as broadcasts -- no shuffles, no per-row dependent index loads;
//   * the edge loop issues exactly deg loads in pieces of 8/4/2/1 (no predicated padding), adds in edge order;
//   * the mean divides with IEEE division only when the count is not a power of two (x * 2^-k is the same
//     correctly rounded value as x / 2^k).
// A group of LPR lanes owns a row (LPR = 32: the whole warp, nothing diverges).  Results are bit-identical to the
// generic kernel: same values, same order of additions.
__device__ __forceinline__ int ldg_stream_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// address of source row c: one IMAD.WIDE (32 x 32 -> 64 bit product added to the 64-bit lane base)
__device__ __forceinline__ const float4* row_ptr(const char* __restrict__ pb, int c, int row_bytes) {
  return reinterpret_cast<const float4*>(pb + (int64_t)c * row_bytes);
}
template <int U>
__device__ __forceinline__ void gather_piece(const char* __restrict__ pb, int row_bytes, const int* w, float4& acc) {
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) v[u] = __ldg(row_ptr(pb, w[u], row_bytes));
#pragma unroll
  for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
}

template <int LPR, int UMAX, int RW>
__global__ void __launch_bounds__(256)
k_segment_rows_lean(const float4* __restrict__ src, int64_t FV,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                    int64_t N, int mean, const float4* __restrict__ addend, float4* __restrict__ out) {
  // RW = consecutive rows per warp (a CTA owns 8*RW).  Short strips keep the rows that are in flight on the whole chip
  // (148 SMs x 64 warps x RW rows) inside the L2: with 32-row strips and 512-byte rows that is 155 MB and every source
  // row was re-read 2.6x from DRAM; RW = 8 measured 0.27 -> 0.20 ms on the batch workload (unchanged on config 4).
  constexpr int rw = RW;
  constexpr int RPW = 32 / LPR;            // rows in flight per warp
  constexpr int CH = RPW <= 4 ? 256 * RPW : 1024;   // window: RPW consecutive non-hub rows fit (RPW <= 4); only a cache
  __shared__ int s_win[8][CH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / LPR, lig = lane % LPR;
  int* win = s_win[warp];
  const int64_t wbase = ((int64_t)blockIdx.x * 8 + warp) * rw;
  if (wbase >= N) return;
  const int64_t wend = (wbase + rw < N) ? wbase + rw : N;   // one past the last row of this warp
  const int64_t myrow = (wbase + lane < wend) ? wbase + lane : wend;
  const int rb = __ldg(rowptr + myrow);
  const int re = __ldg(rowptr + (myrow + 1 < wend ? myrow + 1 : wend));
  const int e_last = __shfl_sync(0xffffffffu, re, 31);      // end of this warp's edge range
  const bool cvalid = lig < FV;
  const char* __restrict__ pb = reinterpret_cast<const char*>(src + (cvalid ? lig : 0));
  const int row_bytes = (int)FV * 16;
  int ws = -CH - 1;                        // window covers col[ws, ws + CH)
#pragma unroll 1
  for (int it = 0; it < rw / RPW; ++it) {
    const int rfirst = it * RPW;
    if (wbase + rfirst >= wend) break;
    const int ibeg = __shfl_sync(0xffffffffu, rb, rfirst);
    const int iend = __shfl_sync(0xffffffffu, re, rfirst + RPW - 1);
    if (iend > ws + CH) {                  // refill from the first edge of this iteration's rows
      __syncwarp();
      ws = ibeg;
#pragma unroll
      for (int q = 0; q < CH / 32; ++q) {
        const int e = ws + lane + 32 * q;
        if (e < e_last) win[lane + 32 * q] = ldg_stream_i32(col + e);
      }
      __syncwarp();
    }
    const int beg = __shfl_sync(0xffffffffu, rb, rfirst + g);
    const int end = __shfl_sync(0xffffffffu, re, rfirst + g);
    const int64_t row = wbase + rfirst + g;
    const int deg = end - beg;
    if (row >= wend || deg > SLDM_HUB_DEGREE) continue;     // split rows: k_segment_hub_*
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beg >= ws && end <= ws + CH) {
      const int* w = win + (beg - ws);
      int rem = deg;
      if constexpr (UMAX >= 8) {
        for (; rem >= 8; rem -= 8, w += 8) gather_piece<8>(pb, row_bytes, w, acc);
        if (rem >= 4) { gather_piece<4>(pb, row_bytes, w, acc); rem -= 4; w += 4; }
      } else {
        for (; rem >= 4; rem -= 4, w += 4) gather_piece<4>(pb, row_bytes, w, acc);
      }
      if (rem >= 2) { gather_piece<2>(pb, row_bytes, w, acc); rem -= 2; w += 2; }
      if (rem >= 1) gather_piece<1>(pb, row_bytes, w, acc);
    } else {                               // row not covered by the window (next to a hub row, RPW > 1): direct index loads
#pragma unroll 1
      for (int k = beg; k < end; ++k) {
        const float4 v = __ldg(row_ptr(pb, __ldg(col + k), row_bytes));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (cvalid) {
      if (mean && deg > 1) {
        if ((deg & (deg - 1)) == 0) {
          const float rc = __fdiv_rn(1.f, (float)deg);       // exact: deg is a power of two
          acc.x *= rc; acc.y *= rc; acc.z *= rc; acc.w *= rc;
        } else {
          acc = Vec<float4>::div(acc, ref_count(deg));
        }
      }
      if (addend != nullptr) { float4 a = ldg_stream_f4(addend + row * FV + lig); Vec<float4>::add(a, acc); acc = a; }
      out[row * FV + lig] = acc;
    }
  }
}

template <typename VT, int LPR, int VPL, int UNR>
__global__ void __launch_bounds__(256)
k_segment_hub_chunks(const VT* __restrict__ src, int64_t FV,
                     const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                     const int4* __restrict__ hub_list, const int32_t* __restrict__ hub_count,
                     int cap, VT* __restrict__ partials) {
  constexpr int G = 256 / LPR;
  constexpr int PER = SLDM_HUB_CHUNK / G;
  __shared__ VT sm[G][LPR * VPL];
  const int n = min(*hub_count, cap);
  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x / LPR, lig = threadIdx.x % LPR;
  const unsigned gmask = group_mask<LPR>(lane);
  for (int c = blockIdx.x; c < n; c += gridDim.x) {
    const int4 ent = hub_list[c];  // {row, chunk, nchunks, first}
    const int rbeg = __ldg(rowptr + ent.x), rend = __ldg(rowptr + ent.x + 1);
    const int cb = rbeg + ent.y * SLDM_HUB_CHUNK;
    const int ce = min(rend, cb + SLDM_HUB_CHUNK);
    const int b = min(ce, cb + g * PER), e = min(ce, b + PER);
    for (int64_t fv0 = 0; fv0 < FV; fv0 += LPR * VPL) {
      VT acc[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) acc[q] = Vec<VT>::zero();
      accumulate_range<VT, LPR, VPL, UNR>(src, FV, (int)fv0, col, b, e, lig, gmask, acc);
#pragma unroll
      for (int q = 0; q < VPL; ++q) sm[g][lig + q * LPR] = acc[q];
      __syncthreads();
      if (threadIdx.x < LPR * VPL) {
        const int64_t idx = fv0 + threadIdx.x;
        if (idx < FV) {
          VT s = sm[0][threadIdx.x];
#pragma unroll 4
          for (int gg = 1; gg < G; ++gg) Vec<VT>::add(s, sm[gg][threadIdx.x]);
          partials[(int64_t)c * FV + idx] = s;
        }
      }
      __syncthreads();
    }
  }
}

template <typename VT, int LPR>
__global__ void __launch_bounds__(256)
k_segment_hub_finalize(const VT* __restrict__ partials, int64_t FV,
                       const int32_t* __restrict__ rowptr,
                       const int4* __restrict__ hub_list, const int32_t* __restrict__ hub_count,
                       int cap, int mean, const VT* __restrict__ addend, VT* __restrict__ out) {
  const int n = min(*hub_count, cap);
  const int64_t groups = (int64_t)gridDim.x * (256 / LPR);
  const int lig = threadIdx.x % LPR;
  for (int64_t c = (int64_t)blockIdx.x * (256 / LPR) + threadIdx.x / LPR; c < n; c += groups) {
    const int4 ent = hub_list[c];
    if (ent.y != 0) continue;
    const int64_t row = ent.x;
    const float cnt = ref_count(__ldg(rowptr + row + 1) - __ldg(rowptr + row));
    for (int64_t idx = lig; idx < FV; idx += LPR) {
      VT s = partials[c * FV + idx];
      for (int j = 1; j < ent.z; ++j) Vec<VT>::add(s, partials[(c + j) * FV + idx]);
      if (mean) s = Vec<VT>::div(s, cnt);
      if (addend) { VT a = Vec<VT>::ld(addend + row * FV + idx); Vec<VT>::add(a, s); s = a; }
      out[row * FV + idx] = s;
    }
  }
}

template <typename VT, int LPR, int VPL, int UNR>
static int launch_all(const float* src, int64_t N, int64_t FV,
                      const int32_t* rowptr, const int32_t* col,
                      const int32_t* hub_list, const int32_t* hub_count, int64_t hub_cap,
                      bool mean, const float* addend, float* out, float* partials, cudaStream_t s) {
  const VT* vsrc = reinterpret_cast<const VT*>(src);
  const VT* vadd = reinterpret_cast<const VT*>(addend);
  VT* vout = reinterpret_cast<VT*>(out);
  VT* vpart = reinterpret_cast<VT*>(partials);
  const int64_t grid = ceil_div<int64_t>(N, kRowsPerCta);
  if constexpr (std::is_same<VT, float4>::value && VPL == 1) {
    static const int lean = [] { const char* e = getenv("SLDM_SEG_LEAN"); return e ? atoi(e) : 4; }();   // 0: generic kernel, 4|8: lean kernel with that many loads per piece
    constexpr int RW = (LPR == 32) ? 8 : 32;      // 512-byte rows: short strips (L2 footprint); narrower rows: long strips
    const int64_t lgrid = ceil_div<int64_t>(N, 8 * RW);
    if (lean == 4 || lean == 1)
      k_segment_rows_lean<LPR, 4, RW><<<(unsigned)lgrid, 256, 0, s>>>(vsrc, FV, rowptr, col, N, mean ? 1 : 0, vadd, vout);
    else if (lean != 0)
      k_segment_rows_lean<LPR, 8, RW><<<(unsigned)lgrid, 256, 0, s>>>(vsrc, FV, rowptr, col, N, mean ? 1 : 0, vadd, vout);
    else
      k_segment_rows<VT, LPR, VPL, UNR><<<(unsigned)grid, 256, 0, s>>>(vsrc, FV, rowptr, col, N, mean ? 1 : 0, vadd, vout);
  } else {
    k_segment_rows<VT, LPR, VPL, UNR><<<(unsigned)grid, 256, 0, s>>>(vsrc, FV, rowptr, col, N, mean ? 1 : 0, vadd, vout);
  }
  SLDM_LAUNCH_CHECK("k_segment_rows");
  if (hub_cap > 0 && hub_list != nullptr) {
    const int cap = (int)hub_cap;
    const int g1 = (int)std::min<int64_t>(hub_cap, (int64_t)num_sms() * 8);
    k_segment_hub_chunks<VT, LPR, VPL, UNR><<<g1, 256, 0, s>>>(
        vsrc, FV, rowptr, col, reinterpret_cast<const int4*>(hub_list), hub_count, cap, vpart);
    SLDM_LAUNCH_CHECK("k_segment_hub_chunks");
    const int g2 = (int)std::min<int64_t>(ceil_div<int64_t>(hub_cap, 256 / LPR), (int64_t)num_sms() * 4);
    k_segment_hub_finalize<VT, LPR><<<g2, 256, 0, s>>>(
        vpart, FV, rowptr, reinterpret_cast<const int4*>(hub_list), hub_count, cap, mean ? 1 : 0, vadd, vout);
    SLDM_LAUNCH_CHECK("k_segment_hub_finalize");
  }
  return SLDM_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int segment_reduce_launch(const float* src, int64_t N, int32_t F,
                          const int32_t* rowptr, const int32_t* col,
                          const int32_t* hub_list, const int32_t* hub_count, int64_t hub_cap,
                          bool mean, const float* addend, float* out,
                          float* partials, cudaStream_t s) {
  if (N == 0 || F == 0) return SLDM_OK;
  const bool vec = (F % 4 == 0) && aligned16(src) && aligned16(out) && aligned16(addend) && aligned16(partials);
  if (vec) {
    const int64_t FV = F / 4;
    if (FV <= 4)  return launch_all<float4, 4, 1, 8>(src, N, FV, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
    if (FV <= 8)  return launch_all<float4, 8, 1, 8>(src, N, FV, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
    if (FV <= 16) return launch_all<float4, 16, 1, 8>(src, N, FV, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
    if (FV <= 32) return launch_all<float4, 32, 1, 8>(src, N, FV, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
    return launch_all<float4, 32, 2, 4>(src, N, FV, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
  }
  // unaligned / F % 4 != 0: same algorithm on 32-bit slices
  if (F <= 32) return launch_all<float, 32, 1, 8>(src, N, F, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
  return launch_all<float, 32, 4, 4>(src, N, F, rowptr, col, hub_list, hub_count, hub_cap, mean, addend, out, partials, s);
}

}  // namespace sldm

using namespace sldm;

extern "C" int64_t sldm_segment_workspace_bytes(int64_t N, int64_t E, int32_t F) {
  if (N < 0 || E < 0 || F < 0) return -1;
  return align_bytes(hub_capacity(E) * (int64_t)F * 4);
}

extern "C" int sldm_segment_reduce(const float* src, int64_t N, int32_t F,
                                   const int32_t* csr, int64_t E, int32_t transpose,
                                   int32_t mean, const float* addend, float* out,
                                   void* workspace, int64_t workspace_bytes,
                                   sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && E >= 0 && F >= 0, SLDM_EINVAL, "sldm_segment_reduce: negative size");
  SLDM_REQUIRE(csr != nullptr, SLDM_EINVAL, "sldm_segment_reduce: csr is NULL");
  SLDM_REQUIRE(N == 0 || F == 0 || (src != nullptr && out != nullptr), SLDM_EINVAL, "sldm_segment_reduce: NULL src/out");
  const int64_t need = sldm_segment_workspace_bytes(N, E, F);
  SLDM_REQUIRE(workspace_bytes >= need && (workspace != nullptr || need == 0), SLDM_EWORKSPACE,
               "sldm_segment_reduce: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  CsrLayout L = csr_layout(N, E);
  const int32_t* meta = csr + L.off[SLDM_CSR_META];
  const int32_t* rowptr = csr + L.off[transpose ? SLDM_CSR_ROWPTR_SRC : SLDM_CSR_ROWPTR_DST];
  const int32_t* col = csr + L.off[transpose ? SLDM_CSR_COL_DST : SLDM_CSR_COL_SRC];
  const int32_t* hub = csr + L.off[transpose ? SLDM_CSR_HUB_SRC : SLDM_CSR_HUB_DST];
  return segment_reduce_launch(src, N, F, rowptr, col, E > SLDM_HUB_DEGREE ? hub : nullptr,
                               meta + (transpose ? 1 : 0), E > SLDM_HUB_DEGREE ? hub_capacity(E) : 0,
                               mean != 0, addend, out, static_cast<float*>(workspace),
                               static_cast<cudaStream_t>(stream));
}

// readout.cu -- graph-level readout that consumes the SageBlock output:
//   global_mean_pool(x, batch), global_max_pool(x, batch) and their concatenation ('double'),
//   src/models/grusage.py:113-120 (choice of pooling) and :185 (x = self.global_pool(x, batch)).
// PyG 2.7.0 semantics restated (nn/pool/glob.py + utils/_scatter.py):
//   mean[g,:] = sum_{i: batch[i]=g} x[i,:] / max(count_g, 1)          (empty graph -> 0)
//   max[g,:]  = max_{i: batch[i]=g} x[i,:]   via scatter_reduce_(amax, include_self=False) into zeros
//               (empty graph -> 0)
//   backward : dx[i,:] = dmean[g,:]/max(count_g,1) + [x[i,:] == max[g,:]] * dmax[g,:] / ties[g,:]
//               (torch's amax backward shares the gradient evenly among the elements equal to the maximum)
// The reference path does a host sync (batch.max().item()) and four scatter passes; here the graph segments come
// from the same device-side CSR machinery as the aggregation (members of graph g = one CSR row, in node order, so an
// unsorted batch vector is handled too), one CTA per graph streams its rows once for both statistics, and the backward
// is one kernel, a CTA per graph: a sweep that counts the ties, then a sweep over the same rows (from L2) that writes
// dx.  No atomics: every result is deterministic.
//
// Bound: HBM.  Algorithmic bytes: forward N*F*4 read + G*2F*4 written; backward N*F*4 read + N*F*4 written (x a second
// time from L2).
#include "common.cuh"
#include <algorithm>
#include <math.h>

namespace sldm {

constexpr int kRoThreads = 256;
constexpr int kRoWarps = kRoThreads / 32;

// out_mean / out_max: [G, ld] row-major views (ld >= F), either may be NULL.
// members[ptr[g] .. ptr[g+1]) are the node ids of graph g (ascending: the CSR build is stable).
template <bool VEC>
__global__ void __launch_bounds__(kRoThreads)
k_readout_fwd(const float* __restrict__ x, int32_t F, const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ members, float* __restrict__ out_mean, float* __restrict__ out_max,
              int64_t ld) {
  extern __shared__ float sm[];                 // [kRoWarps][2][F]
  const int g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int beg = __ldg(ptr + g), end = __ldg(ptr + g + 1);
  const int W = VEC ? 4 : 1;                    // floats per lane per column step
  for (int c0 = 0; c0 < F; c0 += 32 * W * 2) { // two column slices per lane per sweep
    float s[2][4], m[2][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int e = 0; e < 4; ++e) { s[q][e] = 0.f; m[q][e] = -INFINITY; }
    // a warp walks rows beg+warp, +8, ...: four member ids are fetched first, then their four rows (independent loads
    // in flight: one row at a time behind a dependent index load ran at 2.7 TB/s), accumulated in the same row order
    int i = beg + warp;
    for (; i + 3 * kRoWarps < end; i += 4 * kRoWarps) {
      const float* row[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) row[k] = x + (int64_t)__ldg(members + i + k * kRoWarps) * F;
      if (VEC) {
        float4 v[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + (q * 32 + lane) * W;
            if (c < F) v[k][q] = ldg4(row[k] + c);
          }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + (q * 32 + lane) * W;
            if (c < F) {
              s[q][0] += v[k][q].x; s[q][1] += v[k][q].y; s[q][2] += v[k][q].z; s[q][3] += v[k][q].w;
              m[q][0] = fmaxf(m[q][0], v[k][q].x); m[q][1] = fmaxf(m[q][1], v[k][q].y);
              m[q][2] = fmaxf(m[q][2], v[k][q].z); m[q][3] = fmaxf(m[q][3], v[k][q].w);
            }
          }
      } else {
        float v[4][2];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + (q * 32 + lane) * W;
            if (c < F) v[k][q] = __ldg(row[k] + c);
          }
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c = c0 + (q * 32 + lane) * W;
            if (c < F) { s[q][0] += v[k][q]; m[q][0] = fmaxf(m[q][0], v[k][q]); }
          }
      }
    }
    for (; i < end; i += kRoWarps) {
      const float* row = x + (int64_t)__ldg(members + i) * F;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = c0 + (q * 32 + lane) * W;
        if (c < F) {
          if (VEC) {
            const float4 v = ldg4(row + c);
            s[q][0] += v.x; s[q][1] += v.y; s[q][2] += v.z; s[q][3] += v.w;
            m[q][0] = fmaxf(m[q][0], v.x); m[q][1] = fmaxf(m[q][1], v.y);
            m[q][2] = fmaxf(m[q][2], v.z); m[q][3] = fmaxf(m[q][3], v.w);
          } else {
            const float v = __ldg(row + c);
            s[q][0] += v; m[q][0] = fmaxf(m[q][0], v);
          }
        }
      }
    }
    // combine the 8 warps in warp order (fixed: deterministic)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = (q * 32 + lane) * W;       // column inside this sweep
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (e < W && c0 + c + e < F) {
          sm[(warp * 2 + 0) * F + c0 + c + e] = s[q][e];
          sm[(warp * 2 + 1) * F + c0 + c + e] = m[q][e];
        }
      }
    }
  }
  __syncthreads();
  const int cnt = end - beg;
  const float fc = (float)(cnt < 1 ? 1 : cnt);
  for (int c = threadIdx.x; c < F; c += kRoThreads) {
    float s = sm[c], m = sm[F + c];
#pragma unroll
    for (int w = 1; w < kRoWarps; ++w) { s += sm[(w * 2) * F + c]; m = fmaxf(m, sm[(w * 2 + 1) * F + c]); }
    if (out_mean) out_mean[(int64_t)g * ld + c] = __fdiv_rn(s, fc);
    if (out_max) out_max[(int64_t)g * ld + c] = cnt > 0 ? m : 0.f;
  }
}

// NaN handling: fmaxf drops NaNs while torch's amax propagates them; features reaching the readout are finite
// (LayerNorm + (Leaky)ReLU outputs), the parity tests use finite inputs.

// Per-graph backward coefficients, one CTA per graph (same row walk as the forward):
//   coef[g, c]     = dmean[g,c] / max(count_g, 1)
//   coef[g, F + c] = dmax[g,c] / ties[g,c],   ties[g,c] = #{i in graph g : x[i,c] == max[g,c]}
template <bool VEC>
__global__ void __launch_bounds__(kRoThreads)
k_readout_coef(const float* __restrict__ x, int32_t F, const int32_t* __restrict__ ptr,
               const int32_t* __restrict__ members, const float* __restrict__ out_max, int64_t ld_max,
               const float* __restrict__ dmean, const float* __restrict__ dmax, int64_t ld_d,
               float* __restrict__ coef) {
  extern __shared__ float sm[];                 // [kRoWarps][F]
  const int g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int beg = __ldg(ptr + g), end = __ldg(ptr + g + 1);
  constexpr int W = VEC ? 4 : 1;
  if (dmax != nullptr) {
    for (int c = lane * W; c < F; c += 32 * W) {
      float mx[4], n[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < W; ++e) mx[e] = __ldg(out_max + (int64_t)g * ld_max + c + e);
      int i = beg + warp;
      for (; i + 3 * kRoWarps < end; i += 4 * kRoWarps) {   // four independent rows in flight (counts: order-free)
        const float* row[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) row[k] = x + (int64_t)__ldg(members + i + k * kRoWarps) * F + c;
        if (VEC) {
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = ldg4(row[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            n[0] += v[k].x == mx[0] ? 1.f : 0.f; n[1] += v[k].y == mx[1] ? 1.f : 0.f;
            n[2] += v[k].z == mx[2] ? 1.f : 0.f; n[3] += v[k].w == mx[3] ? 1.f : 0.f;
          }
        } else {
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldg(row[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) n[0] += v[k] == mx[0] ? 1.f : 0.f;
        }
      }
      for (; i < end; i += kRoWarps) {
        const float* row = x + (int64_t)__ldg(members + i) * F + c;
        if (VEC) {
          const float4 v = ldg4(row);
          n[0] += v.x == mx[0] ? 1.f : 0.f; n[1] += v.y == mx[1] ? 1.f : 0.f;
          n[2] += v.z == mx[2] ? 1.f : 0.f; n[3] += v.w == mx[3] ? 1.f : 0.f;
        } else {
          n[0] += __ldg(row) == mx[0] ? 1.f : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < W; ++e) sm[warp * F + c + e] = n[e];
    }
  }
  __syncthreads();
  const int cnt = end - beg;
  const float fc = (float)(cnt < 1 ? 1 : cnt);
  for (int c = threadIdx.x; c < F; c += kRoThreads) {
    coef[(int64_t)g * 2 * F + c] = dmean ? __fdiv_rn(__ldg(dmean + (int64_t)g * ld_d + c), fc) : 0.f;
    float b = 0.f;
    if (dmax != nullptr) {
      float n = 0.f;
#pragma unroll
      for (int w = 0; w < kRoWarps; ++w) n += sm[w * F + c];
      b = n > 0.f ? __fdiv_rn(__ldg(dmax + (int64_t)g * ld_d + c), n) : 0.f;
    }
    coef[(int64_t)g * 2 * F + F + c] = b;
  }
}

// The whole backward of one graph in one CTA: sweep 1 counts the ties per feature (k_readout_coef's walk) and leaves
// the two coefficient rows in shared memory, sweep 2 walks the same rows again -- ~100 KB per graph, still in L2 --
// and writes dx.  HBM traffic: x read once, dx written once (the two-kernel version read x twice: 76 + 182 us at the
// bench shape).  Every node is a member of exactly one CSR row, so every dx row is written exactly once.
template <bool VEC>
__global__ void __launch_bounds__(kRoThreads)
k_readout_bwd_graph(const float* __restrict__ x, int32_t F, const int32_t* __restrict__ ptr,
                    const int32_t* __restrict__ members, const float* __restrict__ out_max, int64_t ld_max,
                    const float* __restrict__ dmean, const float* __restrict__ dmax, int64_t ld_d,
                    float* __restrict__ dx) {
  extern __shared__ float sm[];                 // [kRoWarps][F] tie counts, then [3][F]: coefA | coefB | max
  const int g = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int beg = __ldg(ptr + g), end = __ldg(ptr + g + 1);
  if (end <= beg) return;
  constexpr int W = VEC ? 4 : 1;
  if (dmax != nullptr) {
    for (int c = lane * W; c < F; c += 32 * W) {
      float mx[4], n[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int e = 0; e < W; ++e) mx[e] = __ldg(out_max + (int64_t)g * ld_max + c + e);
      int i = beg + warp;
      for (; i + 3 * kRoWarps < end; i += 4 * kRoWarps) {   // four independent rows in flight (counts: order-free)
        const float* row[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) row[k] = x + (int64_t)__ldg(members + i + k * kRoWarps) * F + c;
        if (VEC) {
          float4 v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = ldg4(row[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            n[0] += v[k].x == mx[0] ? 1.f : 0.f; n[1] += v[k].y == mx[1] ? 1.f : 0.f;
            n[2] += v[k].z == mx[2] ? 1.f : 0.f; n[3] += v[k].w == mx[3] ? 1.f : 0.f;
          }
        } else {
          float v[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) v[k] = __ldg(row[k]);
#pragma unroll
          for (int k = 0; k < 4; ++k) n[0] += v[k] == mx[0] ? 1.f : 0.f;
        }
      }
      for (; i < end; i += kRoWarps) {
        const float* row = x + (int64_t)__ldg(members + i) * F + c;
        if (VEC) {
          const float4 v = ldg4(row);
          n[0] += v.x == mx[0] ? 1.f : 0.f; n[1] += v.y == mx[1] ? 1.f : 0.f;
          n[2] += v.z == mx[2] ? 1.f : 0.f; n[3] += v.w == mx[3] ? 1.f : 0.f;
        } else {
          n[0] += __ldg(row) == mx[0] ? 1.f : 0.f;
        }
      }
#pragma unroll
      for (int e = 0; e < W; ++e) sm[warp * F + c + e] = n[e];
    }
  }
  __syncthreads();
  const float fc = (float)(end - beg);
  float ca[4], cb[4], cm[4];                     // F <= 1024 = 4 columns per thread
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = threadIdx.x + j * kRoThreads;
    ca[j] = cb[j] = cm[j] = 0.f;
    if (c < F) {
      ca[j] = dmean ? __fdiv_rn(__ldg(dmean + (int64_t)g * ld_d + c), fc) : 0.f;
      if (dmax != nullptr) {
        float n = 0.f;
#pragma unroll
        for (int w = 0; w < kRoWarps; ++w) n += sm[w * F + c];
        cb[j] = n > 0.f ? __fdiv_rn(__ldg(dmax + (int64_t)g * ld_d + c), n) : 0.f;
        cm[j] = __ldg(out_max + (int64_t)g * ld_max + c);
      }
    }
  }
  __syncthreads();                               // the tie counts have been read: reuse the buffer
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = threadIdx.x + j * kRoThreads;
    if (c < F) { sm[c] = ca[j]; sm[F + c] = cb[j]; sm[2 * F + c] = cm[j]; }
  }
  __syncthreads();
  const bool use_max = dmax != nullptr;
  int i = beg + warp;
  if (VEC && use_max) {                          // four rows in flight per warp, like sweep 1
    for (; i + 3 * kRoWarps < end; i += 4 * kRoWarps) {
      int64_t id[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) id[k] = __ldg(members + i + k * kRoWarps);
      for (int c = lane * 4; c < F; c += 128) {
        float4 xv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xv[k] = ldg4(x + id[k] * F + c);
        const float4 a = *reinterpret_cast<const float4*>(sm + c), b = *reinterpret_cast<const float4*>(sm + F + c),
                     mx = *reinterpret_cast<const float4*>(sm + 2 * F + c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 r = a;
          r.x += xv[k].x == mx.x ? b.x : 0.f; r.y += xv[k].y == mx.y ? b.y : 0.f;
          r.z += xv[k].z == mx.z ? b.z : 0.f; r.w += xv[k].w == mx.w ? b.w : 0.f;
          st4(dx + id[k] * F + c, r);
        }
      }
    }
  }
  for (; i < end; i += kRoWarps) {
    const int64_t id = __ldg(members + i);
    for (int c = lane * W; c < F; c += 32 * W) {
      if (VEC) {
        float4 r = *reinterpret_cast<const float4*>(sm + c);
        if (use_max) {
          const float4 xv = ldg4(x + id * F + c);
          const float4 b = *reinterpret_cast<const float4*>(sm + F + c), mx = *reinterpret_cast<const float4*>(sm + 2 * F + c);
          r.x += xv.x == mx.x ? b.x : 0.f; r.y += xv.y == mx.y ? b.y : 0.f;
          r.z += xv.z == mx.z ? b.z : 0.f; r.w += xv.w == mx.w ? b.w : 0.f;
        }
        st4(dx + id * F + c, r);
      } else {
        float r = sm[c];
        if (use_max && __ldg(x + id * F + c) == sm[2 * F + c]) r += sm[F + c];
        dx[id * F + c] = r;
      }
    }
  }
}

// dx[i,c] = coef[g,c] + (x[i,c] == max[g,c] ? coef[g,F+c] : 0);  one warp per node row, 128-bit accesses when aligned
template <bool VEC>
__global__ void __launch_bounds__(256)
k_readout_bwd(const float* __restrict__ x, int64_t N, int32_t F, const int64_t* __restrict__ batch, int64_t G,
              const float* __restrict__ out_max, int64_t ld_max, const float* __restrict__ coef, int use_max,
              float* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  constexpr int W = VEC ? 4 : 1;
  const int64_t nw = (int64_t)gridDim.x * 8;
  for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < N; i += nw) {
    int64_t g = batch ? batch[i] : 0;
    // out-of-range graph ids: clamped exactly like the membership CSR of the forward (k_convert), so forward and
    // backward agree; the build has raised meta[2] and the host reports it (ops._IndexChecks)
    const bool ok = G > 0;
    g = g < 0 ? 0 : (g >= G ? G - 1 : g);
    g = ok ? g : 0;
    const float* ca = coef + g * 2 * F;
    for (int c = lane * W; c < F; c += 32 * W) {
      if (VEC) {
        float4 r = ok ? ldg4(ca + c) : f4zero();
        if (ok && use_max) {
          const float4 xv = ldg4(x + i * F + c), mx = ldg4(out_max + g * ld_max + c), b = ldg4(ca + F + c);
          r.x += xv.x == mx.x ? b.x : 0.f; r.y += xv.y == mx.y ? b.y : 0.f;
          r.z += xv.z == mx.z ? b.z : 0.f; r.w += xv.w == mx.w ? b.w : 0.f;
        }
        st4(dx + i * F + c, r);
      } else {
        float r = ok ? __ldg(ca + c) : 0.f;
        if (ok && use_max && __ldg(x + i * F + c) == __ldg(out_max + g * ld_max + c)) r += __ldg(ca + F + c);
        dx[i * F + c] = r;
      }
    }
  }
}

}  // namespace sldm

using namespace sldm;

extern "C" int64_t sldm_readout_workspace_bytes(int64_t G, int32_t F) {
  if (G < 0 || F < 0) return -1;
  return align_bytes((G > 0 ? G : 1) * (int64_t)F * 2 * 4);   // per-graph backward coefficients [G, 2F]
}

// csr: membership CSR built by sldm_csr_build_pairs(NULL, batch, N, csr_nodes >= G): row g of (rowptr_dst, col_src) lists the
// nodes of graph g.  out_mean / out_max are [G, ld] views (pass the two halves of one [G,2F] buffer with ld = 2F for
// the 'double' readout); either may be NULL.
extern "C" int sldm_readout_forward(const float* x, int64_t N, int32_t F, const int32_t* csr, int64_t csr_nodes,
                                    int64_t G, float* out_mean, float* out_max, int64_t ld, sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && G >= 0 && F >= 1, SLDM_EINVAL, "sldm_readout_forward: bad sizes N=%lld G=%lld F=%d",
               (long long)N, (long long)G, F);
  SLDM_REQUIRE(csr_nodes >= G, SLDM_ESHAPE, "sldm_readout_forward: membership CSR covers %lld rows, need %lld",
               (long long)csr_nodes, (long long)G);
  SLDM_REQUIRE(ld >= F, SLDM_ESHAPE, "sldm_readout_forward: ld=%lld < F=%d", (long long)ld, F);
  SLDM_REQUIRE(F <= 1024, SLDM_EUNSUPPORTED, "sldm_readout_forward: F=%d > 1024", F);
  if (G == 0) return SLDM_OK;
  SLDM_REQUIRE(csr != nullptr && (N == 0 || x != nullptr), SLDM_EINVAL, "sldm_readout_forward: NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CsrLayout L = csr_layout(csr_nodes, N);
  const int32_t* ptr = csr + L.off[SLDM_CSR_ROWPTR_DST];
  const int32_t* members = csr + L.off[SLDM_CSR_COL_SRC];
  const size_t smem = (size_t)kRoWarps * 2 * F * sizeof(float);
  const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  if (vec) {
    SLDM_OPT_IN_SMEM(k_readout_fwd<true>, 160 * 1024);
    k_readout_fwd<true><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_mean, out_max, ld);
  } else {
    SLDM_OPT_IN_SMEM(k_readout_fwd<false>, 160 * 1024);
    k_readout_fwd<false><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_mean, out_max, ld);
  }
  SLDM_LAUNCH_CHECK("k_readout_fwd");
  return SLDM_OK;
}

// dmean / dmax are [G, ld_d] views of the upstream gradient (either may be NULL); out_max is the forward result
// ([G, ld_max] view; required iff dmax != NULL).  dx [N,F] is overwritten.
extern "C" int sldm_readout_backward(const float* x, int64_t N, int32_t F, const int64_t* batch,
                                     const int32_t* csr, int64_t csr_nodes, int64_t G,
                                     const float* out_max, int64_t ld_max,
                                     const float* dmean, const float* dmax, int64_t ld_d,
                                     float* dx, void* workspace, int64_t workspace_bytes, sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && G >= 0 && F >= 1, SLDM_EINVAL, "sldm_readout_backward: bad sizes");
  SLDM_REQUIRE(F <= 1024, SLDM_EUNSUPPORTED, "sldm_readout_backward: F=%d > 1024", F);
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(x && csr && dx, SLDM_EINVAL, "sldm_readout_backward: NULL pointer");
  SLDM_REQUIRE(dmax == nullptr || out_max != nullptr, SLDM_EINVAL, "sldm_readout_backward: dmax without out_max");
  SLDM_REQUIRE(G >= 1, SLDM_ESHAPE, "sldm_readout_backward: %lld nodes but 0 graphs", (long long)N);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CsrLayout L = csr_layout(csr_nodes, N);
  const int32_t* ptr = csr + L.off[SLDM_CSR_ROWPTR_DST];
  const int32_t* members = csr + L.off[SLDM_CSR_COL_SRC];
  SLDM_REQUIRE(workspace != nullptr && workspace_bytes >= sldm_readout_workspace_bytes(G, F), SLDM_EWORKSPACE,
               "sldm_readout_backward: workspace too small");
  float* coef = static_cast<float*>(workspace);
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = (F % 4 == 0) && a16(x) && a16(dx) && a16(out_max) && (ld_max % 4 == 0);
  const size_t smem = (size_t)kRoWarps * F * sizeof(float);
  static const bool fused = [] { const char* e = getenv("SLDM_READOUT_TWO_KERNELS"); return !(e && e[0] == '1'); }();
  if (fused) {   // one CTA per graph does both sweeps (SLDM_READOUT_TWO_KERNELS=1: the coefficient kernel + the row-parallel kernel)
    if (vec) k_readout_bwd_graph<true><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_max, ld_max, dmean, dmax, ld_d, dx);
    else     k_readout_bwd_graph<false><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_max, ld_max, dmean, dmax, ld_d, dx);
    SLDM_LAUNCH_CHECK("k_readout_bwd_graph");
    return SLDM_OK;
  }
  if (vec) k_readout_coef<true><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_max, ld_max, dmean, dmax, ld_d, coef);
  else     k_readout_coef<false><<<(unsigned)G, kRoThreads, smem, s>>>(x, F, ptr, members, out_max, ld_max, dmean, dmax, ld_d, coef);
  SLDM_LAUNCH_CHECK("k_readout_coef");
  const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(N, 8), (int64_t)num_sms() * 16);
  if (vec) k_readout_bwd<true><<<grid, 256, 0, s>>>(x, N, F, batch, G, out_max, ld_max, coef, dmax != nullptr, dx);
  else     k_readout_bwd<false><<<grid, 256, 0, s>>>(x, N, F, batch, G, out_max, ld_max, coef, dmax != nullptr, dx);
  SLDM_LAUNCH_CHECK("k_readout_bwd");
  return SLDM_OK;
}

// edge_build.cu -- proximity edges between vehicle trajectories.
//
// Reference: the O(V^2 T) Python double loop of src/gbuilder.py:88-112 (offline packs) and :244-268
// (GraphOnlineCreator, the per-window cost of rcv.py:77).  For every ORDERED pair (i, j), i != j:
//   dists = || x[i,:,0:2] - x[j,:,0:2] ||_2 over the frames where both presence flags x[.,:,4] exceed 0.5
//   edge (i, j) exists iff dists is non-empty and min(dists) <= m_radius
//   edge_attr = [min, max, mean, mean of squares] of dists (float32, numpy's pairwise summation)
// Edges come out in (i, j) lexicographic order.  Here: one CTA per source vehicle i; pass 1 counts the edges of every
// row, a single-CTA scan turns the counts into offsets, pass 2 recomputes the pair statistics and writes each row's edges
// in ascending j (block-wide exclusive scan of the flags per chunk of 256 targets): deterministic, no atomics.
// Bound: FP32 issue (V^2 T distance evaluations); V is a few hundred, so the whole build is a few microseconds of GPU time.
#include "common.cuh"
#include <math.h>

namespace sldm {

constexpr int kEbMaxT = 128;   // frames per trajectory (numpy sums <= 128 elements with its unrolled 8-way loop, no recursion)

// numpy's pairwise_sum for n <= 128 (numpy/core/src/umath/loops_utils.h.src): n < 8 sequential from 0; else eight
// running sums over blocks of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the tail sequentially
__device__ __forceinline__ float np_pairwise_sum(const float* a, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, a[i]);
    return r;
  }
  float r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], a[i + k]);
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, a[i]);
  return res;
}

struct PairStats { int n; float mn, mx, mean, msq; };

// x: [V, T, F] row-major fp32; s_xi: this CTA's source trajectory {X, Y, presence} per frame in shared memory
__device__ __forceinline__ PairStats pair_stats(const float* __restrict__ x, int T, int F, const float* s_xi, int j,
                                                bool want_means) {
  float d[kEbMaxT];
  PairStats ps; ps.n = 0; ps.mn = INFINITY; ps.mx = -INFINITY; ps.mean = 0.f; ps.msq = 0.f;
  const float* xj = x + (int64_t)j * T * F;
  for (int t = 0; t < T; ++t) {
    if (s_xi[3 * t + 2] > 0.5f && __ldg(xj + t * F + 4) > 0.5f) {
      const float dx = __fsub_rn(s_xi[3 * t], __ldg(xj + t * F)), dy = __fsub_rn(s_xi[3 * t + 1], __ldg(xj + t * F + 1));
      const float v = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // np.linalg.norm(axis=1), unfused
      ps.mn = fminf(ps.mn, v); ps.mx = fmaxf(ps.mx, v);
      if (want_means) d[ps.n] = v;
      ++ps.n;
    }
  }
  if (want_means && ps.n > 0) {
    ps.mean = __fdiv_rn(np_pairwise_sum(d, ps.n), (float)ps.n);
    for (int k = 0; k < ps.n; ++k) d[k] = __fmul_rn(d[k], d[k]);
    ps.msq = __fdiv_rn(np_pairwise_sum(d, ps.n), (float)ps.n);
  }
  return ps;
}

template <bool FILL>
__global__ void __launch_bounds__(256)
k_edge_rows(const float* __restrict__ x, int V, int T, int F, float radius, int32_t* __restrict__ rowcnt,
            const int32_t* __restrict__ rowoff, int64_t* __restrict__ edge_index, int64_t E_cap,
            float* __restrict__ edge_attr) {
  extern __shared__ float s_xi[];                 // [T][3]
  __shared__ int s_scan[256 / 32 + 1];
  const int i = blockIdx.x;
  for (int t = threadIdx.x; t < T; t += 256) {
    const float* xi = x + ((int64_t)i * T + t) * F;
    s_xi[3 * t] = xi[0]; s_xi[3 * t + 1] = xi[1]; s_xi[3 * t + 2] = xi[4];
  }
  __syncthreads();
  int run = FILL ? rowoff[i] : 0;                 // edges of this row written so far (uniform across the CTA)
  for (int j0 = 0; j0 < V; j0 += 256) {
    const int j = j0 + threadIdx.x;
    PairStats ps; ps.n = 0; ps.mn = INFINITY;
    if (j < V && j != i) ps = pair_stats(x, T, F, s_xi, j, FILL);
    const int flag = (ps.n > 0 && ps.mn <= radius) ? 1 : 0;
    int total;
    const int pre = block_exclusive_scan_256(flag, total, s_scan);
    if (FILL && flag) {
      const int64_t e = run + pre;
      edge_index[e] = i;
      edge_index[E_cap + e] = j;
      float4 a = make_float4(ps.mn, ps.mx, ps.mean, ps.msq);
      *reinterpret_cast<float4*>(edge_attr + 4 * e) = a;
    }
    run += total;
  }
  if (!FILL && threadIdx.x == 0) rowcnt[i] = run;
}

// offsets[i] = exclusive prefix of counts; offsets[V] = total; single CTA (V is a few hundred .. a few thousand)
__global__ void __launch_bounds__(256)
k_edge_offsets(const int32_t* __restrict__ cnt, int V, int32_t* __restrict__ off) {
  __shared__ int s_scan[256 / 32 + 1];
  int carry = 0;
  for (int c = 0; c < V; c += 256) {
    const int i = c + threadIdx.x;
    const int v = i < V ? cnt[i] : 0;
    int total;
    const int p = block_exclusive_scan_256(v, total, s_scan);
    if (i < V) off[i] = carry + p;
    carry += total;
  }
  if (threadIdx.x == 0) off[V] = carry;
}

}  // namespace sldm

using namespace sldm;

// Pass 1: counts[V] and offsets[V+1] (offsets[V] = number of edges E; the host reads it to size the outputs).
extern "C" int sldm_edge_build_count(const float* x, int64_t V, int32_t T, int32_t F, float m_radius,
                                     int32_t* counts, int32_t* offsets, sldm_stream_t stream) {
  SLDM_REQUIRE(V >= 0 && T >= 0, SLDM_EINVAL, "sldm_edge_build_count: negative size");
  SLDM_REQUIRE(F >= 5, SLDM_ESHAPE, "sldm_edge_build_count: need >= 5 temporal features (X, Y, ., ., presence), got %d", F);
  SLDM_REQUIRE(T <= kEbMaxT, SLDM_EUNSUPPORTED, "sldm_edge_build_count: %d frames > %d", T, kEbMaxT);
  SLDM_REQUIRE(V < ((int64_t)1 << 24), SLDM_EUNSUPPORTED, "sldm_edge_build_count: too many vehicles");
  SLDM_REQUIRE(offsets != nullptr && (V == 0 || (x != nullptr && counts != nullptr)), SLDM_EINVAL, "sldm_edge_build_count: NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (V > 0) {
    k_edge_rows<false><<<(unsigned)V, 256, (size_t)3 * T * sizeof(float), s>>>(x, (int)V, T, F, m_radius, counts, nullptr, nullptr, 0, nullptr);
    SLDM_LAUNCH_CHECK("k_edge_rows<count>");
  }
  k_edge_offsets<<<1, 256, 0, s>>>(counts, (int)V, offsets);
  SLDM_LAUNCH_CHECK("k_edge_offsets");
  return SLDM_OK;
}

// Pass 2: edge_index int64 [2, E] and edge_attr fp32 [E, 4] (16-byte aligned), E = offsets[V] as read by the host.
extern "C" int sldm_edge_build_fill(const float* x, int64_t V, int32_t T, int32_t F, float m_radius,
                                    const int32_t* offsets, int64_t E, int64_t* edge_index, float* edge_attr,
                                    sldm_stream_t stream) {
  SLDM_REQUIRE(V >= 0 && T >= 0 && E >= 0, SLDM_EINVAL, "sldm_edge_build_fill: negative size");
  SLDM_REQUIRE(F >= 5 && T <= kEbMaxT, SLDM_ESHAPE, "sldm_edge_build_fill: bad F / T");
  if (V == 0 || E == 0) return SLDM_OK;
  SLDM_REQUIRE(x && offsets && edge_index && edge_attr, SLDM_EINVAL, "sldm_edge_build_fill: NULL pointer");
  SLDM_REQUIRE((reinterpret_cast<uintptr_t>(edge_attr) & 15u) == 0, SLDM_EINVAL, "sldm_edge_build_fill: edge_attr must be 16-byte aligned");
  k_edge_rows<true><<<(unsigned)V, 256, (size_t)3 * T * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
      x, (int)V, T, F, m_radius, nullptr, offsets, edge_index, E, edge_attr);
  SLDM_LAUNCH_CHECK("k_edge_rows<fill>");
  return SLDM_OK;
}

// csr_build.cu -- device-side edge_index -> CSR (+ transpose CSR) build.
//
// Replaces the implicit "scatter by edge_index[1]" of the reference
// (PyG utils/_scatter.py::scatter reached from src/models/blocks/sageblock.py:18)
// by an explicit, deterministic CSR:  a stable LSD radix sort of the edges by
// destination (and by source for the transpose used in backward) plus a degree
// histogram and exclusive prefix sum.  Stability keeps every segment in edge
// order, which is the summation order of the reference's CPU scatter_add_.
//
// All of this is HBM-bound int32 work: 16E bytes read (int64 pairs), 8E written
// per sorted column array, 8(N+1) for the row pointers.
#include "common.cuh"
#include <algorithm>

namespace sldm {

// ------------------------------------------------------------------ layout --
CsrLayout csr_layout(int64_t N, int64_t E) {
  CsrLayout L;
  int64_t o = 0;
  int64_t cap = hub_capacity(E);
  // meta, rowptr_dst, rowptr_src are contiguous so one memset clears them
  L.off[SLDM_CSR_META] = o;        o += align_i32(64);
  L.off[SLDM_CSR_ROWPTR_DST] = o;  o += align_i32(N + 1);
  L.off[SLDM_CSR_ROWPTR_SRC] = o;  o += align_i32(N + 1);
  L.off[SLDM_CSR_COL_SRC] = o;     o += align_i32(E > 0 ? E : 1);
  L.off[SLDM_CSR_COL_DST] = o;     o += align_i32(E > 0 ? E : 1);
  L.off[SLDM_CSR_HUB_DST] = o;     o += align_i32(cap * 4);
  L.off[SLDM_CSR_HUB_SRC] = o;     o += align_i32(cap * 4);
  L.off[SLDM_CSR_TOTAL] = o;
  return L;
}

// -------------------------------------------------------------------- scan --
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

int64_t scan_spine_elems(int64_t n) { return ceil_div<int64_t>(n, kScanTile) + 2; }

// exclusive scan of one value per thread across the block; returns the prefix
// of this thread and (to every thread) the block total.
template <int THREADS>
__device__ __forceinline__ int block_exclusive_scan(int v, int& total, int* smem /*[THREADS/32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  constexpr int NW = THREADS / 32;
  if (warp == 0) {
    int w = (lane < NW) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < NW) smem[lane] = winc - w;  // exclusive warp prefix
    if (lane == 31) smem[NW] = winc;       // total (lanes >= NW contribute 0)
  }
  __syncthreads();
  int prefix = smem[warp] + inc - v;
  total = smem[NW];
  __syncthreads();
  return prefix;
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_reduce(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ spine) {
  __shared__ int sm[kScanThreads / 32 + 1];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) s += in[idx];
  }
  int total;
  block_exclusive_scan<kScanThreads>(s, total, sm);
  if (threadIdx.x == 0) spine[blockIdx.x] = total;
}

// single block: exclusive scan of spine[0..nb) in place
__global__ void __launch_bounds__(1024)
k_scan_spine(int32_t* __restrict__ spine, int64_t nb) {
  __shared__ int sm[1024 / 32 + 1];
  int carry = 0;
  for (int64_t c = 0; c < nb; c += 1024) {
    int64_t idx = c + threadIdx.x;
    int v = (idx < nb) ? spine[idx] : 0;
    int total;
    int p = block_exclusive_scan<1024>(v, total, sm);
    if (idx < nb) spine[idx] = carry + p;
    carry += total;
  }
}

__global__ void __launch_bounds__(kScanThreads)
k_scan_apply(const int32_t* in, int32_t* out, int64_t n, const int32_t* __restrict__ spine) {
  __shared__ int sm[kScanThreads / 32 + 1];
  int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = (idx < n) ? in[idx] : 0;
    s += v[i];
  }
  int total;
  int p = block_exclusive_scan<kScanThreads>(s, total, sm) + spine[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = p;
    p += v[i];
  }
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* spine, cudaStream_t s) {
  if (n <= 0) return SLDM_OK;
  int64_t nb = ceil_div<int64_t>(n, kScanTile);
  k_scan_reduce<<<(unsigned)nb, kScanThreads, 0, s>>>(in, n, spine);
  SLDM_LAUNCH_CHECK("k_scan_reduce");
  k_scan_spine<<<1, 1024, 0, s>>>(spine, nb);
  SLDM_LAUNCH_CHECK("k_scan_spine");
  k_scan_apply<<<(unsigned)nb, kScanThreads, 0, s>>>(in, out, n, spine);
  SLDM_LAUNCH_CHECK("k_scan_apply");
  return SLDM_OK;
}

// ------------------------------------------------------- convert + degrees --
__global__ void __launch_bounds__(256)
k_convert_count(const int64_t* __restrict__ ei, int64_t E, int32_t N,
                int32_t* __restrict__ src32, int32_t* __restrict__ dst32,
                int32_t* __restrict__ deg_dst, int32_t* __restrict__ deg_src,
                int32_t* __restrict__ meta) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    int64_t s = ei[e], d = ei[E + e];
    if (e > 0) {
      if (ei[e - 1] > s) meta[3] = 1;
      if (ei[E + e - 1] > d) meta[4] = 1;
    }
    if (s < 0 || s >= N || d < 0 || d >= N) {
      meta[2] = 1;  // reported through meta; clamped so nothing goes out of bounds
      s = s < 0 ? 0 : (s >= N ? N - 1 : s);
      d = d < 0 ? 0 : (d >= N ? N - 1 : d);
    }
    src32[e] = (int32_t)s;
    dst32[e] = (int32_t)d;
    atomicAdd(deg_dst + d, 1);
    atomicAdd(deg_src + s, 1);
  }
}

// ----------------------------------------------------- stable LSD radix sort --
constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;
constexpr int kRsRounds = 8;
constexpr int kRsTile = kRsThreads * kRsRounds;  // 2048 keys per block

// block_hist is digit-major: [256][nb]
__global__ void __launch_bounds__(kRsThreads)
k_radix_hist(const int32_t* __restrict__ keys, int64_t n, int shift,
             int32_t* __restrict__ block_hist, int nb,
             const int32_t* __restrict__ unsorted_flag) {
  if (unsorted_flag && *unsorted_flag == 0) return;  // input already ordered
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t base = (int64_t)blockIdx.x * kRsTile + warp * (32 * kRsRounds) + lane;
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    int64_t idx = base + r * 32;
    if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255], 1);
  }
  __syncthreads();
  block_hist[(int64_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(kRsThreads)
k_radix_scatter(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in,
                int64_t n, int shift, const int32_t* __restrict__ block_off, int nb,
                int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                const int32_t* __restrict__ unsorted_flag) {
  if (unsorted_flag && *unsorted_flag == 0) return;
  __shared__ int wh[kRsWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int w = 0; w < kRsWarps; ++w) wh[w][threadIdx.x] = 0;
  __syncthreads();

  int64_t base = (int64_t)blockIdx.x * kRsTile + warp * (32 * kRsRounds) + lane;
  int k[kRsRounds], v[kRsRounds];
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    int64_t idx = base + r * 32;
    bool valid = idx < n;
    k[r] = valid ? keys_in[idx] : 0;
    v[r] = valid ? vals_in[idx] : 0;
    if (valid) atomicAdd(&wh[warp][(k[r] >> shift) & 255], 1);
  }
  __syncthreads();
  {  // per digit: turn per-warp counts into per-warp start offsets (warp order = key order)
    int d = threadIdx.x;
    int run = block_off[(int64_t)d * nb + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRsWarps; ++w) {
      int c = wh[w][d];
      wh[w][d] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRsRounds; ++r) {
    int64_t idx = base + r * 32;
    bool valid = idx < n;
    unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      int d = (k[r] >> shift) & 255;
      unsigned m = __match_any_sync(vm, d);
      int rank = __popc(m & ((1u << lane) - 1u));
      int leader = __ffs(m) - 1;
      int old = 0;
      if (lane == leader) { old = wh[warp][d]; wh[warp][d] = old + __popc(m); }
      old = __shfl_sync(m, old, leader);
      int pos = old + rank;
      if (keys_out) keys_out[pos] = k[r];
      vals_out[pos] = v[r];
    }
    __syncwarp();
  }
}

// when the keys were already non-decreasing the stable sort is the identity
__global__ void __launch_bounds__(256)
k_copy_if_sorted(const int32_t* __restrict__ vals, int64_t n, int32_t* __restrict__ out,
                 const int32_t* __restrict__ unsorted_flag) {
  if (*unsorted_flag != 0) return;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vals[i];
}

static int radix_sort_pairs(const int32_t* keys0, const int32_t* vals0, int64_t n, int npass,
                            int32_t* kA, int32_t* vA, int32_t* kB, int32_t* vB,
                            int32_t* final_vals, int32_t* hist, int32_t* spine,
                            const int32_t* unsorted_flag, cudaStream_t s) {
  int nb = (int)ceil_div<int64_t>(n, kRsTile);
  const int32_t* in_k = keys0;
  const int32_t* in_v = vals0;
  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256 * 4), (int64_t)num_sms() * 8);
    k_copy_if_sorted<<<grid, 256, 0, s>>>(vals0, n, final_vals, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_copy_if_sorted");
  }
  for (int p = 0; p < npass; ++p) {
    bool last = (p == npass - 1);
    int32_t* out_k = last ? nullptr : ((p & 1) ? kB : kA);
    int32_t* out_v = last ? final_vals : ((p & 1) ? vB : vA);
    k_radix_hist<<<nb, kRsThreads, 0, s>>>(in_k, n, 8 * p, hist, nb, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_radix_hist");
    int rc = exclusive_scan_i32(hist, hist, (int64_t)256 * nb, spine, s);
    if (rc) return rc;
    k_radix_scatter<<<nb, kRsThreads, 0, s>>>(in_k, in_v, n, 8 * p, hist, nb, out_k, out_v, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_radix_scatter");
    in_k = out_k; in_v = out_v;
  }
  return SLDM_OK;
}

// ----------------------------------------------------------- hub work list --
__global__ void __launch_bounds__(256)
k_plan_hubs(const int32_t* __restrict__ rowptr, int32_t N, int32_t* __restrict__ hub_list,
            int32_t* __restrict__ counter, int32_t cap) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int deg = rowptr[i + 1] - rowptr[i];
  if (deg <= SLDM_HUB_DEGREE) return;
  int nch = (deg + SLDM_HUB_CHUNK - 1) / SLDM_HUB_CHUNK;
  int k = atomicAdd(counter, nch);
  if (k + nch > cap) return;  // cannot happen: cap >= E/CHUNK + E/HUB_DEGREE
  for (int c = 0; c < nch; ++c) {
    int4 ent = make_int4((int)i, c, nch, k);
    reinterpret_cast<int4*>(hub_list)[k + c] = ent;
  }
}

}  // namespace sldm

// ================================================================== C ABI ==
using namespace sldm;

static int radix_passes(int64_t N) {
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) < N) ++bits;
  return (bits + 7) / 8;
}

struct CsrWs { int64_t src32, dst32, kA, vA, kB, vB, hist, spine, total; };
static CsrWs csr_ws_layout(int64_t N, int64_t E) {
  CsrWs w; int64_t o = 0;
  int64_t e = align_bytes((E > 0 ? E : 1) * 4);
  int64_t nb = ceil_div<int64_t>(E > 0 ? E : 1, kRsTile);
  w.src32 = o; o += e;  w.dst32 = o; o += e;
  w.kA = o; o += e;  w.vA = o; o += e;  w.kB = o; o += e;  w.vB = o; o += e;
  w.hist = o; o += align_bytes(256 * nb * 4);
  int64_t sp = scan_spine_elems(256 * nb);
  int64_t sp2 = scan_spine_elems(N + 1);
  w.spine = o; o += align_bytes((sp > sp2 ? sp : sp2) * 4);
  w.total = o;
  return w;
}

extern "C" int sldm_csr_layout(int64_t N, int64_t E, int64_t* out8) {
  SLDM_REQUIRE(out8 != nullptr, SLDM_EINVAL, "sldm_csr_layout: out8 is NULL");
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_csr_layout: negative N or E");
  CsrLayout L = csr_layout(N, E);
  for (int i = 0; i < 8; ++i) out8[i] = L.off[i];
  return SLDM_OK;
}

extern "C" int64_t sldm_csr_workspace_bytes(int64_t N, int64_t E) {
  if (N < 0 || E < 0) return -1;
  return csr_ws_layout(N, E).total;
}

extern "C" int sldm_csr_build(const int64_t* edge_index, int64_t E, int64_t N,
                              int32_t* csr, void* workspace, int64_t workspace_bytes,
                              sldm_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t kMax = ((int64_t)1 << 31) - ((int64_t)1 << 20);
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_csr_build: negative N (%lld) or E (%lld)", (long long)N, (long long)E);
  SLDM_REQUIRE(N < kMax && E < kMax, SLDM_EUNSUPPORTED, "sldm_csr_build: N=%lld / E=%lld exceed the int32 CSR", (long long)N, (long long)E);
  SLDM_REQUIRE(!(E > 0 && N == 0), SLDM_EINVAL, "sldm_csr_build: %lld edges but 0 nodes", (long long)E);
  SLDM_REQUIRE(csr != nullptr, SLDM_EINVAL, "sldm_csr_build: csr is NULL");
  SLDM_REQUIRE(E == 0 || edge_index != nullptr, SLDM_EINVAL, "sldm_csr_build: edge_index is NULL");
  CsrLayout L = csr_layout(N, E);
  CsrWs W = csr_ws_layout(N, E);
  SLDM_REQUIRE(E == 0 || (workspace != nullptr && workspace_bytes >= W.total), SLDM_EWORKSPACE,
               "sldm_csr_build: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)W.total);

  int32_t* meta = csr + L.off[SLDM_CSR_META];
  int32_t* rp_d = csr + L.off[SLDM_CSR_ROWPTR_DST];
  int32_t* rp_s = csr + L.off[SLDM_CSR_ROWPTR_SRC];
  int32_t* col_s = csr + L.off[SLDM_CSR_COL_SRC];
  int32_t* col_d = csr + L.off[SLDM_CSR_COL_DST];
  int32_t* hub_d = csr + L.off[SLDM_CSR_HUB_DST];
  int32_t* hub_s = csr + L.off[SLDM_CSR_HUB_SRC];

  // meta + both rowptr arrays are contiguous
  SLDM_CUDA(cudaMemsetAsync(meta, 0, (size_t)(L.off[SLDM_CSR_COL_SRC] - L.off[SLDM_CSR_META]) * 4, s));
  if (E == 0) return SLDM_OK;

  char* wb = static_cast<char*>(workspace);
  int32_t* src32 = reinterpret_cast<int32_t*>(wb + W.src32);
  int32_t* dst32 = reinterpret_cast<int32_t*>(wb + W.dst32);
  int32_t* kA = reinterpret_cast<int32_t*>(wb + W.kA);
  int32_t* vA = reinterpret_cast<int32_t*>(wb + W.vA);
  int32_t* kB = reinterpret_cast<int32_t*>(wb + W.kB);
  int32_t* vB = reinterpret_cast<int32_t*>(wb + W.vB);
  int32_t* hist = reinterpret_cast<int32_t*>(wb + W.hist);
  int32_t* spine = reinterpret_cast<int32_t*>(wb + W.spine);

  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(E, 256), (int64_t)num_sms() * 16);
    k_convert_count<<<grid, 256, 0, s>>>(edge_index, E, (int32_t)N, src32, dst32, rp_d, rp_s, meta);
    SLDM_LAUNCH_CHECK("k_convert_count");
  }
  int rc;
  if ((rc = exclusive_scan_i32(rp_d, rp_d, N + 1, spine, s))) return rc;
  if ((rc = exclusive_scan_i32(rp_s, rp_s, N + 1, spine, s))) return rc;

  int npass = radix_passes(N);
  // by destination: keys = dst, payload = src  -> col_src   (skipped when dst already ordered)
  if ((rc = radix_sort_pairs(dst32, src32, E, npass, kA, vA, kB, vB, col_s, hist, spine, meta + 4, s))) return rc;
  // by source (transpose): keys = src, payload = dst -> col_dst
  if ((rc = radix_sort_pairs(src32, dst32, E, npass, kA, vA, kB, vB, col_d, hist, spine, meta + 3, s))) return rc;

  int cap = (int)hub_capacity(E);
  int grid = (int)ceil_div<int64_t>(N, 256);
  k_plan_hubs<<<grid, 256, 0, s>>>(rp_d, (int32_t)N, hub_d, meta + 0, cap);
  SLDM_LAUNCH_CHECK("k_plan_hubs(dst)");
  k_plan_hubs<<<grid, 256, 0, s>>>(rp_s, (int32_t)N, hub_s, meta + 1, cap);
  SLDM_LAUNCH_CHECK("k_plan_hubs(src)");
  return SLDM_OK;
}

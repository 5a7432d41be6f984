// csr_build.cu -- device-side edge_index -> CSR (+ transpose CSR) build.
//
// Replaces the implicit "scatter by edge_index[1]" of the reference
// (PyG utils/_scatter.py::scatter reached from src/models/blocks/sageblock.py:18)
// by an explicit, deterministic CSR:  a stable LSD radix sort of the edges by
// destination (and by source for the transpose used in backward); the row pointers
// are read off the sorted keys (no atomics anywhere).  Stability keeps every
// segment in edge order, which is the summation order of the reference's CPU
// scatter_add_.
//
// Kernels per build: k_convert (int64 -> int32, range / sortedness flags), then per
// key array k_digit_hist (all digits in one read), k_digit_offsets, one
// k_onesweep_pass per 8-bit digit, k_rowptr_from_sorted; k_plan_hubs for the split
// rows.  A key array that is already non-decreasing (edge_index[0] as the
// reference's builders emit it, src/gbuilder.py:88-112) skips its sort on the device.
//
// All of this is HBM-bound int32 work: 16E bytes read (int64 pairs), 8E written
// per sorted column array, 8(N+1) for the row pointers.  Measured on B200
// (profiles/r01g_csr_launches.txt): 4.1 M edges 0.29 ms, 10 M unsorted edges 0.81 ms; the
// passes are bound by the stable ranking (warp votes on the ADU pipe, or shared-memory
// atomics emulating match.any), not by HBM.
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

// ---------------------------------------------------------------- convert --
// int64 [2,E] -> two int32 arrays; range check (reported through meta, indices clamped) and the two
// "already non-decreasing" flags.  Pure streaming: 16E bytes read, 8E written.
__global__ void __launch_bounds__(256)
k_convert(const int64_t* __restrict__ ei_src, const int64_t* __restrict__ ei_dst, int64_t E, int32_t N,
          int32_t* __restrict__ src32, int32_t* __restrict__ dst32, int32_t* __restrict__ meta) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  bool bad = false, us = false, ud = false;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += stride) {
    int64_t s = ei_src ? ei_src[e] : e, d = ei_dst[e];   // ei_src == NULL: source ids are 0..E-1 (membership lists)
    if (e > 0) {
      us |= ei_src != nullptr && ei_src[e - 1] > s;
      ud |= ei_dst[e - 1] > d;
    }
    if ((ei_src != nullptr && (s < 0 || s >= N)) || d < 0 || d >= N) {
      bad = true;  // reported through meta; clamped so nothing goes out of bounds
      if (ei_src != nullptr) s = s < 0 ? 0 : (s >= N ? N - 1 : s);
      d = d < 0 ? 0 : (d >= N ? N - 1 : d);
    }
    src32[e] = (int32_t)s;
    dst32[e] = (int32_t)d;
  }
  if (bad) meta[2] = 1;
  if (us) meta[3] = 1;
  if (ud) meta[4] = 1;
}

// ----------------------------------------------------- stable LSD radix sort --
// One kernel per 8-bit digit ("onesweep"): a CTA takes a tile of 4096 consecutive keys, ranks them stably by digit
// (per-warp match_any ranking + prefix over warps), learns how many keys with the same digit precede its tile from
// the tiles before it (decoupled look-back over per-tile state words) and writes the tile out grouped by digit
// through shared memory, so consecutive threads write consecutive addresses.  The global start of every digit comes
// from a histogram of all digits taken in ONE read of the keys before the first pass.  Pure integer work, no float:
// the result is the unique stable sort, bit-exact against argsort(stable).
constexpr int kOsThreads = 256;
constexpr int kOsWarps = kOsThreads / 32;
constexpr int kOsKpt = 16;                       // keys per thread
constexpr int kOsTile = kOsThreads * kOsKpt;     // 4096 keys per CTA
constexpr int kMaxPass = 4;
#ifndef SLDM_RANK_ATOMIC_OR
#define SLDM_RANK_ATOMIC_OR 1
#endif
constexpr bool kRankAtomicOr = SLDM_RANK_ATOMIC_OR != 0;   // stable ranking: shared-memory atomicOr peer masks (1) or warp ballots (0)

// ghist[p][d] += #keys whose p-th digit is d, for all passes at once.  High digits of clustered keys are uniform
// across a warp: one shared-memory atomic per warp instead of 32.
__global__ void __launch_bounds__(256)
k_digit_hist(const int32_t* __restrict__ keys, int64_t n, int npass, int32_t* __restrict__ ghist,
             const int32_t* __restrict__ unsorted_flag) {
  if (unsorted_flag && *unsorted_flag == 0) return;  // input already ordered: the sort is skipped
  __shared__ int h[kMaxPass][256];
  for (int i = threadIdx.x; i < kMaxPass * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t nround = ceil_div<int64_t>(n, stride);
  for (int64_t r = 0; r < nround; ++r) {
    const int64_t idx = r * stride + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool valid = idx < n;
    const int k = valid ? keys[idx] : 0;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (vm == 0) continue;
    const int first = __ffs(vm) - 1;
    for (int p = 0; p < npass; ++p) {
      const int d = (k >> (8 * p)) & 255;
      const int d0 = __shfl_sync(0xffffffffu, d, first);
      const bool uniform = __ballot_sync(0xffffffffu, !valid || d == d0) == 0xffffffffu;
      if (uniform) { if (lane == first) atomicAdd(&h[p][d0], __popc(vm)); }
      else if (valid) atomicAdd(&h[p][d], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < npass * 256; i += 256) {
    const int c = (&h[0][0])[i];
    if (c) atomicAdd(ghist + i, c);
  }
}

// gbase[p][d] = exclusive prefix of ghist[p][*]; one CTA of 256 threads per key array, all passes
__global__ void __launch_bounds__(256)
k_digit_offsets(const int32_t* __restrict__ ghist, int32_t* __restrict__ gbase, int npass) {
  __shared__ int sm[256 / 32 + 1];
  for (int p = 0; p < npass; ++p) {
    int total;
    const int v = ghist[p * 256 + threadIdx.x];
    gbase[p * 256 + threadIdx.x] = block_exclusive_scan<256>(v, total, sm);
  }
}

// lanes of `vm` whose 8-bit digit equals this lane's: eight ballots.  (__match_any_sync does the same in one
// instruction, but MATCH.ANY occupies the ADU pipe for ~140 cycles on sm_100 -- ncu: sm__inst_executed_pipe_adu 76%,
// 206 us per pass on 10 M keys -- so the ranking loop was the whole cost of the pass.)
__device__ __forceinline__ unsigned peers_same_digit(unsigned vm, int d, int nbits) {
  unsigned m = vm;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    if (b >= nbits) break;                         // the top digit of the key range is narrower than 8 bits
    const bool bit = (d >> b) & 1;
    const unsigned bal = __ballot_sync(vm, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

__device__ __forceinline__ unsigned long long ld_state(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_state(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// tile_state: [ntiles][256] 64-bit words, {flag (2 bits) | count}; flag 1 = this tile's own count, 2 = inclusive
// prefix over tiles 0..t.  ticket: tiles are handed out in launch order so every earlier tile is running or done.
__global__ void __launch_bounds__(kOsThreads)
k_onesweep_pass(const int32_t* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n, int shift, int nbits,
                const int32_t* __restrict__ gbase, unsigned long long* __restrict__ tile_state,
                int32_t* __restrict__ ticket, int32_t* __restrict__ keys_out, int32_t* __restrict__ vals_out,
                const int32_t* __restrict__ unsorted_flag) {
  if (unsorted_flag && *unsorted_flag == 0) return;
  __shared__ int wh[kOsWarps][256];          // per-warp digit counts, then per-warp start offsets
  __shared__ int s_lstart[256];              // start of digit d inside the digit-grouped tile
  __shared__ int s_gdelta[256];              // global position of element i of the grouped tile = s_gdelta[d] + i
  __shared__ int s_scan[kOsThreads / 32 + 1];
  __shared__ int s_tile;
  __shared__ int skeys[kOsTile];
  __shared__ int svals[kOsTile];
  constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagInc = 2ull << 62, kMask = (1ull << 62) - 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(ticket, 1);
#pragma unroll
  for (int w = 0; w < kOsWarps; ++w) wh[w][tid] = 0;
  __syncthreads();
  const int tile = s_tile;
  const int64_t tbase = (int64_t)tile * kOsTile;
  const int tile_n = (int)((n - tbase < kOsTile) ? (n - tbase) : kOsTile);

  // ---- load + stable rank inside the warp's 512-key slice (slice order = key order) ----
  int k[kOsKpt], v[kOsKpt];
  unsigned short rl[kOsKpt];
  const int wbase = warp * (32 * kOsKpt) + lane;
#pragma unroll
  for (int r = 0; r < kOsKpt; ++r) {
    const int i = wbase + r * 32;
    const bool valid = i < tile_n;
    k[r] = valid ? keys_in[tbase + i] : 0;
    v[r] = valid ? vals_in[tbase + i] : 0;
  }
  if (kRankAtomicOr && nbits > 5) {   // narrow digits: a few ballots are cheaper
    // peer mask through shared memory: every lane ORs its bit into the word of its digit, then reads the word back
    // (an emulated match.any on the LSU pipe instead of eight votes on the ADU pipe)
    unsigned* mk = reinterpret_cast<unsigned*>(skeys) + warp * 256;   // skeys is not used before the regroup step
#pragma unroll
    for (int j = 0; j < 8; ++j) mk[lane + 32 * j] = 0u;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < kOsKpt; ++r) {
      const bool valid = wbase + r * 32 < tile_n;
      const int d = (k[r] >> shift) & 255;
      if (valid) atomicOr(&mk[d], 1u << lane);
      __syncwarp();
      unsigned m = 0u;
      int old = 0;
      if (valid) { m = mk[d]; old = wh[warp][d]; }
      __syncwarp();
      if (valid && lane == __ffs(m) - 1) { wh[warp][d] = old + __popc(m); mk[d] = 0u; }
      __syncwarp();
      rl[r] = (unsigned short)(old + __popc(m & ((1u << lane) - 1u)));
    }
  } else {
#pragma unroll
  for (int r = 0; r < kOsKpt; ++r) {
    const bool valid = wbase + r * 32 < tile_n;
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const int d = (k[r] >> shift) & 255;
      const unsigned m = peers_same_digit(vm, d, nbits);
      const int leader = __ffs(m) - 1;
      int old = 0;
      if (lane == leader) { old = wh[warp][d]; wh[warp][d] = old + __popc(m); }
      old = __shfl_sync(m, old, leader);
      rl[r] = (unsigned short)(old + __popc(m & ((1u << lane) - 1u)));
    }
    __syncwarp();
  }
  }
  __syncthreads();

  // ---- per digit (thread = digit): offsets over warps, tile count, position among the tiles ----
  {
    const int d = tid;
    int run = 0;
#pragma unroll
    for (int w = 0; w < kOsWarps; ++w) { const int c = wh[w][d]; wh[w][d] = run; run += c; }
    const int cnt = run;
    int total;
    const int lstart = block_exclusive_scan<kOsThreads>(cnt, total, s_scan);
    unsigned long long* mine = tile_state + (int64_t)tile * 256 + d;
    long long excl = 0;
    if (tile == 0) {
      st_state(mine, kFlagInc | (unsigned long long)cnt);
    } else {
      st_state(mine, kFlagAgg | (unsigned long long)cnt);
      // Walk back over the earlier tiles until one with an inclusive prefix is found.  kLook state words are fetched
      // per step (independent loads), so the walk costs one L2 round trip per kLook tiles; in the first wave, where
      // no tile is inclusive yet, that walk is the critical path of the pass.
      constexpr int kLook = 8;
      int t = tile - 1;
      bool done = false;
      while (!done) {
        unsigned long long sv[kLook];
#pragma unroll
        for (int j = 0; j < kLook; ++j)
          sv[j] = (t - j >= 0) ? ld_state(tile_state + (int64_t)(t - j) * 256 + d) : kFlagInc;
#pragma unroll
        for (int j = 0; j < kLook; ++j) {
          if (done) break;
          const unsigned long long fl = sv[j] >> 62;
          if (fl == 0) break;                      // not published yet: poll again from this tile
          excl += (long long)(sv[j] & kMask);
          --t;
          if (fl == 2) done = true;
        }
      }
      st_state(mine, kFlagInc | (unsigned long long)(excl + cnt));
    }
    s_lstart[d] = lstart;
    s_gdelta[d] = gbase[d] + (int)excl - lstart;
  }
  __syncthreads();

  // ---- group the tile by digit in shared memory, then write runs of equal digits with consecutive threads ----
#pragma unroll
  for (int r = 0; r < kOsKpt; ++r) {
    if (wbase + r * 32 < tile_n) {
      const int d = (k[r] >> shift) & 255;
      const int pos = s_lstart[d] + wh[warp][d] + rl[r];
      skeys[pos] = k[r];
      svals[pos] = v[r];
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kOsKpt; ++j) {
    const int i = j * kOsThreads + tid;
    if (i < tile_n) {
      const int key = skeys[i];
      const int pos = s_gdelta[(key >> shift) & 255] + i;
      keys_out[pos] = key;
      vals_out[pos] = svals[i];
    }
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

// rowptr[k] = first position whose key is >= k, from the sorted keys (the input itself when it was already ordered):
// rowptr == cumsum(bincount(keys)) without a single atomic.  rowptr[N] = n.
__global__ void __launch_bounds__(256)
k_rowptr_from_sorted(const int32_t* __restrict__ keys_if_sorted, const int32_t* __restrict__ keys_after_sort,
                     int64_t n, int32_t N, int32_t* __restrict__ rowptr, const int32_t* __restrict__ unsorted_flag) {
  // Position i writes rowptr[k] = i for every k in (keys[i-1], keys[i]].  Runs of absent keys are normally short; a long
  // one (trailing isolated nodes, membership lists whose keys only span the graphs) is parked in shared memory and
  // filled by the whole CTA, so no single thread ever writes more than kInline entries in a row.
  constexpr int kInline = 16, kMaxGaps = 64;
  __shared__ int g_lo[kMaxGaps], g_hi[kMaxGaps], g_val[kMaxGaps];
  __shared__ int g_n;
  const int32_t* __restrict__ keys = (*unsorted_flag == 0) ? keys_if_sorted : keys_after_sort;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t rounds = ceil_div<int64_t>(n + 1, stride);
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i = r * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int lo = 0, hi = 0;
    if (i <= n) {
      lo = (i == 0) ? -1 : keys[i - 1];
      hi = (i == n) ? N : keys[i];
    }
    const bool big = hi - lo > kInline;
    if (!big) for (int kk = lo + 1; kk <= hi; ++kk) rowptr[kk] = (int32_t)i;
    if (!__syncthreads_or(big)) continue;          // the common case: no long run in this CTA's slice
    if (threadIdx.x == 0) g_n = 0;
    __syncthreads();
    if (big) {
      const int slot = atomicAdd(&g_n, 1);
      if (slot < kMaxGaps) { g_lo[slot] = lo; g_hi[slot] = hi; g_val[slot] = (int32_t)i; }
      else for (int kk = lo + 1; kk <= hi; ++kk) rowptr[kk] = (int32_t)i;   // list full: fall back to the serial fill
    }
    __syncthreads();
    const int ng = min(g_n, kMaxGaps);
    for (int q = 0; q < ng; ++q)
      for (int kk = g_lo[q] + 1 + threadIdx.x; kk <= g_hi[q]; kk += blockDim.x) rowptr[kk] = g_val[q];
    __syncthreads();
  }
}

struct SortWs {            // per key array
  int32_t* ghist;          // [kMaxPass][256]
  int32_t* gbase;          // [kMaxPass][256]
  int32_t* tickets;        // [kMaxPass]
  unsigned long long* tile_state;   // [npass][ntiles][256]
};

// keys0/vals0 -> final_vals (stable by key); the sorted keys end in *sorted_keys_out (kA or kB)
static int radix_sort_pairs(const int32_t* keys0, const int32_t* vals0, int64_t n, int npass, int key_bits,
                            int32_t* kA, int32_t* vA, int32_t* kB, int32_t* vB,
                            int32_t* final_vals, const SortWs& W, const int32_t** sorted_keys_out,
                            const int32_t* unsorted_flag, cudaStream_t s) {
  const int64_t ntiles = ceil_div<int64_t>(n, kOsTile);
  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256 * 4), (int64_t)num_sms() * 8);
    k_copy_if_sorted<<<grid, 256, 0, s>>>(vals0, n, final_vals, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_copy_if_sorted");
    grid = (int)std::min<int64_t>(ceil_div<int64_t>(n, 256 * 8), (int64_t)num_sms() * 8);
    k_digit_hist<<<grid, 256, 0, s>>>(keys0, n, npass, W.ghist, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_digit_hist");
    k_digit_offsets<<<1, 256, 0, s>>>(W.ghist, W.gbase, npass);
    SLDM_LAUNCH_CHECK("k_digit_offsets");
  }
  const int32_t* in_k = keys0;
  const int32_t* in_v = vals0;
  for (int p = 0; p < npass; ++p) {
    const bool last = (p == npass - 1);
    int32_t* out_k = (p & 1) ? kB : kA;
    int32_t* out_v = last ? final_vals : ((p & 1) ? vB : vA);
    const int nbits = std::min(8, key_bits - 8 * p);
    k_onesweep_pass<<<(unsigned)ntiles, kOsThreads, 0, s>>>(in_k, in_v, n, 8 * p, nbits, W.gbase + p * 256,
                                                           W.tile_state + (int64_t)p * ntiles * 256, W.tickets + p,
                                                           out_k, out_v, unsorted_flag);
    SLDM_LAUNCH_CHECK("k_onesweep_pass");
    in_k = out_k; in_v = out_v;
  }
  *sorted_keys_out = in_k;
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

static int key_bits(int64_t N) {
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) < N) ++bits;
  return bits;
}
static int radix_passes(int64_t N) { return (key_bits(N) + 7) / 8; }

// workspace: int32 copies of the two index rows, two ping-pong (key, value) pairs, and the sort state.  The sort
// state (histograms, tickets, per-tile look-back words for every pass of both sorts) is one contiguous region that a
// single memset clears.
struct CsrWs { int64_t src32, dst32, kA, vA, kB, vB, state, state_bytes, total; int64_t ntiles; int npass; };
static int64_t sort_state_bytes(int64_t ntiles, int npass) {
  return align_bytes(2 * kMaxPass * 256 * 4) + align_bytes(kMaxPass * 4) + align_bytes((int64_t)npass * ntiles * 256 * 8);
}
static CsrWs csr_ws_layout(int64_t N, int64_t E) {
  CsrWs w; int64_t o = 0;
  int64_t e = align_bytes((E > 0 ? E : 1) * 4);
  w.ntiles = ceil_div<int64_t>(E > 0 ? E : 1, kOsTile);
  w.npass = radix_passes(N);
  w.src32 = o; o += e;  w.dst32 = o; o += e;
  w.kA = o; o += e;  w.vA = o; o += e;  w.kB = o; o += e;  w.vB = o; o += e;
  w.state = o; w.state_bytes = 2 * sort_state_bytes(w.ntiles, w.npass); o += w.state_bytes;
  w.total = o;
  return w;
}
static SortWs sort_ws_at(char* base, int64_t ntiles, int npass) {
  SortWs W;
  W.ghist = reinterpret_cast<int32_t*>(base);
  W.gbase = W.ghist + kMaxPass * 256;
  char* p = base + align_bytes(2 * kMaxPass * 256 * 4);
  W.tickets = reinterpret_cast<int32_t*>(p);
  p += align_bytes(kMaxPass * 4);
  W.tile_state = reinterpret_cast<unsigned long long*>(p);
  (void)ntiles; (void)npass;
  return W;
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

// meta[2] (1 if any index was outside [0, N)) -> *status_host_pinned, asynchronously behind the build on `stream`:
// the host presets the word to -1 and reads 0 / 1 once the copy has landed -- no event, no synchronisation.
extern "C" int sldm_csr_status_async(const int32_t* csr, int32_t* status_host_pinned, sldm_stream_t stream) {
  SLDM_REQUIRE(csr != nullptr && status_host_pinned != nullptr, SLDM_EINVAL, "sldm_csr_status_async: NULL pointer");
  SLDM_CUDA(cudaMemcpyAsync(status_host_pinned, csr + 2, sizeof(int32_t), cudaMemcpyDeviceToHost,
                            static_cast<cudaStream_t>(stream)));
  return SLDM_OK;
}

extern "C" int sldm_csr_build(const int64_t* edge_index, int64_t E, int64_t N,
                              int32_t* csr, void* workspace, int64_t workspace_bytes,
                              sldm_stream_t stream) {
  SLDM_REQUIRE(E == 0 || edge_index != nullptr, SLDM_EINVAL, "sldm_csr_build: edge_index is NULL");
  return sldm_csr_build_pairs(edge_index, edge_index ? edge_index + E : nullptr, E, N, csr, workspace, workspace_bytes, stream);
}

extern "C" int sldm_csr_build_pairs(const int64_t* edge_src, const int64_t* edge_dst, int64_t E, int64_t N,
                                    int32_t* csr, void* workspace, int64_t workspace_bytes,
                                    sldm_stream_t stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t kMax = ((int64_t)1 << 31) - ((int64_t)1 << 20);
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_csr_build: negative N (%lld) or E (%lld)", (long long)N, (long long)E);
  SLDM_REQUIRE(N < kMax && E < kMax, SLDM_EUNSUPPORTED, "sldm_csr_build: N=%lld / E=%lld exceed the int32 CSR", (long long)N, (long long)E);
  SLDM_REQUIRE(!(E > 0 && N == 0), SLDM_EINVAL, "sldm_csr_build: %lld edges but 0 nodes", (long long)E);
  SLDM_REQUIRE(csr != nullptr, SLDM_EINVAL, "sldm_csr_build: csr is NULL");
  SLDM_REQUIRE(E == 0 || edge_dst != nullptr, SLDM_EINVAL, "sldm_csr_build: the destination row is NULL");
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
  if (E == 0) {  // no edges: all row pointers are zero (meta + both rowptr arrays are contiguous)
    SLDM_CUDA(cudaMemsetAsync(meta, 0, (size_t)(L.off[SLDM_CSR_COL_SRC] - L.off[SLDM_CSR_META]) * 4, s));
    return SLDM_OK;
  }
  SLDM_CUDA(cudaMemsetAsync(meta, 0, 64 * 4, s));   // flags and hub counters; every rowptr entry is written below

  char* wb = static_cast<char*>(workspace);
  int32_t* src32 = reinterpret_cast<int32_t*>(wb + W.src32);
  int32_t* dst32 = reinterpret_cast<int32_t*>(wb + W.dst32);
  int32_t* kA = reinterpret_cast<int32_t*>(wb + W.kA);
  int32_t* vA = reinterpret_cast<int32_t*>(wb + W.vA);
  int32_t* kB = reinterpret_cast<int32_t*>(wb + W.kB);
  int32_t* vB = reinterpret_cast<int32_t*>(wb + W.vB);
  SLDM_CUDA(cudaMemsetAsync(wb + W.state, 0, (size_t)W.state_bytes, s));
  const SortWs Wd = sort_ws_at(wb + W.state, W.ntiles, W.npass);
  const SortWs Ws = sort_ws_at(wb + W.state + W.state_bytes / 2, W.ntiles, W.npass);

  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(E, 256 * 4), (int64_t)num_sms() * 8);
    k_convert<<<grid, 256, 0, s>>>(edge_src, edge_dst, E, (int32_t)N, src32, dst32, meta);
    SLDM_LAUNCH_CHECK("k_convert");
  }
  int rc;
  const int npass = W.npass;
  const int32_t* sorted_d = nullptr;
  const int32_t* sorted_s = nullptr;
  // by destination: keys = dst, payload = src  -> col_src   (skipped on the device when dst is already ordered)
  if ((rc = radix_sort_pairs(dst32, src32, E, npass, key_bits(N), kA, vA, kB, vB, col_s, Wd, &sorted_d, meta + 4, s))) return rc;
  {
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(E + 1, 256), (int64_t)num_sms() * 16);
    k_rowptr_from_sorted<<<grid, 256, 0, s>>>(dst32, sorted_d, E, (int32_t)N, rp_d, meta + 4);
    SLDM_LAUNCH_CHECK("k_rowptr_from_sorted(dst)");
  }
  const bool membership = (edge_src == nullptr);   // only (rowptr_dst, col_src) are produced; source ids may exceed N
  if (!membership) {
    // by source (transpose): keys = src, payload = dst -> col_dst.  The ping-pong buffers are reused: stream order
    // guarantees the row pointers above were derived before they are overwritten.
    if ((rc = radix_sort_pairs(src32, dst32, E, npass, key_bits(N), kA, vA, kB, vB, col_d, Ws, &sorted_s, meta + 3, s))) return rc;
    int grid = (int)std::min<int64_t>(ceil_div<int64_t>(E + 1, 256), (int64_t)num_sms() * 16);
    k_rowptr_from_sorted<<<grid, 256, 0, s>>>(src32, sorted_s, E, (int32_t)N, rp_s, meta + 3);
    SLDM_LAUNCH_CHECK("k_rowptr_from_sorted(src)");
  }

  int cap = (int)hub_capacity(E);
  int grid = (int)ceil_div<int64_t>(N, 256);
  k_plan_hubs<<<grid, 256, 0, s>>>(rp_d, (int32_t)N, hub_d, meta + 0, cap);
  SLDM_LAUNCH_CHECK("k_plan_hubs(dst)");
  if (!membership) {
    k_plan_hubs<<<grid, 256, 0, s>>>(rp_s, (int32_t)N, hub_s, meta + 1, cap);
    SLDM_LAUNCH_CHECK("k_plan_hubs(src)");
  }
  return SLDM_OK;
}

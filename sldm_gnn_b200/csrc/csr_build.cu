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
// Kernels per build (7 launches + 2 memsets; round 1: 17 + 2): k_convert_hist (int64 -> int32, range / sortedness
// flags, digit histograms of the destination keys in the same read), k_digit_hist (source keys; exits at once when
// they are ordered), one k_onesweep_pass per 8-bit digit serving BOTH sorts (blockIdx.y), k_rowptr_from_sorted and
// k_plan_hubs (both arrays each).  A key array that is already non-decreasing (edge_index[0] as the reference's
// builders emit it, src/gbuilder.py:88-112) skips its sort on the device.
//
// All of this is HBM-bound int32 work by its bytes -- 16E read (int64 pairs), 8E written per sorted column array,
// 8(N+1) for the row pointers -- but not by its time.  Measured on B200 (profiles/r02j_csr_launches.txt): 4.1 M edges
// 0.193 ms (round 1: 0.27), 10 M unsorted edges 0.60 ms (0.78), 32 k edges 0.054 ms (0.077).  A pass runs at
// ~115 G keys/s: it is instruction-bound (ncu: issue slots 62 % busy, 209 warp instructions per 32 keys -- stable
// ranking, regrouping through shared memory, look-back), the convert kernel is bound by its shared-memory atomics
// (~64 LSU cycles per warp instruction: 1.4 per 32 keys = 40 us of the 41).
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

// ------------------------------------------------------------------ plan --
// Stable LSD radix sort of (key, value) pairs, keys < N, 8-bit digits.  A 10-bit variant of the same kernels exists
// behind SLDM_CSR_DIGIT_BITS=10 (tests run both): it saves a whole pass for 17..20 key bits (0.8 M .. 1 M nodes) and
// still LOSES -- measured on B200, 4.1 M / 10 M edges: 0.262 / 0.83 ms against 0.234 / 0.61 ms, a 1024-bin pass costs
// almost twice a 256-bin one (four digits per thread in the prefix phase, 73 KB of shared memory per CTA).
constexpr int kOsThreads = 256;
constexpr int kOsWarps = kOsThreads / 32;
constexpr int kOsKpt = 16;                       // keys per thread
constexpr int kOsTile = kOsThreads * kOsKpt;     // 4096 keys per CTA
constexpr int kMaxPass = 4;
#ifndef SLDM_OS_MIN_CTAS
#define SLDM_OS_MIN_CTAS 4
#endif
constexpr int kOsMinCtas = SLDM_OS_MIN_CTAS;      // resident CTAs per SM the pass kernel is compiled for (64 registers at 4)
constexpr int kHistInts = 4096;                  // per sort: npass * bins <= max(4 * 256, 3 * 1024)
#ifndef SLDM_RANK_ATOMIC_OR
#define SLDM_RANK_ATOMIC_OR 1
#endif
constexpr bool kRankAtomicOr = SLDM_RANK_ATOMIC_OR != 0;   // stable ranking: shared-memory atomicOr peer masks (1) or warp ballots (0)

struct SortPlan { int db, npass, key_bits; };
static int key_bits(int64_t N) {
  int bits = 1;
  while (bits < 31 && ((int64_t)1 << bits) < N) ++bits;
  return bits;
}
// 64-bit look-back words are needed from 2^30 keys on; SLDM_CSR_WIDE_STATE=1 forces them for any size (tests: the
// variant would otherwise only ever run on inputs of 17 GB and more)
static bool wide_state(int64_t n) {
  const char* e = getenv("SLDM_CSR_WIDE_STATE");
  return n >= ((int64_t)1 << 30) || (e && e[0] == '1');
}
static SortPlan sort_plan(int64_t N) {
  SortPlan P;
  P.key_bits = key_bits(N);
  P.db = 8;
  if (const char* e = getenv("SLDM_CSR_DIGIT_BITS")) { const int v = atoi(e); if (v == 8 || v == 10) P.db = v; }
  P.npass = (P.key_bits + P.db - 1) / P.db;
  return P;
}

// ---------------------------------------------------------------- convert --
// Digit histogram of one warp-round of keys into shared-memory counters h[pass][digit].  The cost is the number of
// shared-memory atomic instructions (~50-64 cycles per warp instruction on the LSU pipe, same-address or not), so a
// digit that is uniform across the warp -- the high digits of clustered keys -- is counted by one lane.  (Leader
// rounds for the first two distinct values of a warp were tried and lost: 57 -> 68 us on the batch, 114 -> 192 us on
// random keys.)
template <int DB>
__device__ __forceinline__ void hist_round(int* __restrict__ h, int k, bool valid, unsigned vm, int npass, int lane) {
  constexpr int NB = 1 << DB;
  const int first = __ffs(vm) - 1;
  for (int p = 0; p < npass; ++p) {
    const int dg = (k >> (DB * p)) & (NB - 1);
    const int d0 = __shfl_sync(0xffffffffu, dg, first);
    const bool uniform = __ballot_sync(0xffffffffu, !valid || dg == d0) == 0xffffffffu;
    if (uniform) { if (lane == first) atomicAdd(&h[p * NB + d0], __popc(vm)); }
    else if (valid) atomicAdd(&h[p * NB + dg], 1);
  }
}

// int64 [2,E] -> two int32 arrays; range check (reported through meta, indices clamped), the two "already
// non-decreasing" flags, and -- in the same read -- the digit histograms of every pass of the sort by DESTINATION
// (ghist[pass][digit], shared-memory counters flushed once per CTA): the copy is HBM-bound, the counting LSU-bound, so
// the two overlap.  The histogram of the source keys is k_digit_hist below: edge_index[0] is normally ordered already
// (src/gbuilder.py:88-112) and that kernel then exits at once.  16E bytes read, 8E written.
template <int DB>
__global__ void __launch_bounds__(256)
k_convert_hist(const int64_t* __restrict__ ei_src, const int64_t* __restrict__ ei_dst, int64_t E, int32_t N, int npass,
               int32_t* __restrict__ src32, int32_t* __restrict__ dst32, int32_t* __restrict__ meta,
               int32_t* __restrict__ ghist_dst) {
  constexpr int NB = 1 << DB;
  extern __shared__ int h[];                      // [npass][NB]
  const int nh = npass * NB;
  for (int i = threadIdx.x; i < nh; i += 256) h[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * 256;
  const int64_t nround = ceil_div<int64_t>(E, stride);
  bool bad = false, us = false, ud = false;
  for (int64_t r = 0; r < nround; ++r) {
    const int64_t e = r * stride + (int64_t)blockIdx.x * 256 + threadIdx.x;
    const bool valid = e < E;
    int64_t s = 0, d = 0;
    if (valid) {
      s = ei_src ? ei_src[e] : e;   // ei_src == NULL: source ids are 0..E-1 (membership lists)
      d = ei_dst[e];
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
    const unsigned vm = __ballot_sync(0xffffffffu, valid);
    if (vm == 0) continue;
    hist_round<DB>(h, (int)d, valid, vm, npass, lane);
  }
  if (bad) meta[2] = 1;
  if (us) meta[3] = 1;
  if (ud) meta[4] = 1;
  __syncthreads();
  for (int i = threadIdx.x; i < nh; i += 256) {
    const int c = h[i];
    if (c) atomicAdd(ghist_dst + i, c);
  }
}

// ghist[p][d] += #keys whose p-th digit is d, for all passes at once (the sort by source; skipped when ordered)
template <int DB>
__global__ void __launch_bounds__(256)
k_digit_hist(const int32_t* __restrict__ keys, int64_t n, int npass, int32_t* __restrict__ ghist,
             const int32_t* __restrict__ unsorted_flag) {
  if (*unsorted_flag == 0) return;  // input already ordered: the sort is skipped
  constexpr int NB = 1 << DB;
  extern __shared__ int h[];
  const int nh = npass * NB;
  for (int i = threadIdx.x; i < nh; i += 256) h[i] = 0;
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
    hist_round<DB>(h, k, valid, vm, npass, lane);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nh; i += 256) {
    const int c = h[i];
    if (c) atomicAdd(ghist + i, c);
  }
}

// ----------------------------------------------------- stable LSD radix sort --
// One kernel per digit ("onesweep"), both sorts of a build in ONE launch (blockIdx.y; a sort whose keys were already
// non-decreasing exits at once): a CTA takes a tile of 4096 consecutive keys, ranks them stably by digit (per-warp
// ranking + prefix over warps), regroups the tile by digit in shared memory, learns how many keys with the same digit
// precede its tile from the tiles before it (decoupled look-back over per-tile state words) and writes the tile out
// so that consecutive threads write consecutive addresses.  The global start of every digit comes from the histogram
// the convert kernel took: tile 0 scans it and folds it into the prefix it publishes.  Pure integer work, no float:
// the result is the unique stable sort, bit-exact against argsort(stable).
//
// Look-back: the tile's keys are already parked in shared memory when it starts, so the registers are free for a WIDE
// window -- 32 state words per thread and round trip.  In the first wave no tile is inclusive yet and the walk is the
// critical path of the pass: ~T / (2 * window) L2 round trips for T resident tiles (window 8: ~28 round trips of the
// 40 us pass at 444 resident tiles; window 32: 7).

// lanes of `vm` whose digit equals this lane's: one ballot per digit bit.  (__match_any_sync does the same in one
// instruction, but MATCH.ANY occupies the ADU pipe for ~140 cycles on sm_100 -- ncu: sm__inst_executed_pipe_adu 76%,
// 206 us per pass on 10 M keys -- so the ranking loop was the whole cost of the pass.)
__device__ __forceinline__ unsigned peers_same_digit(unsigned vm, int d, int nbits) {
  unsigned m = vm;
#pragma unroll
  for (int b = 0; b < 10; ++b) {
    if (b >= nbits) break;                         // the top digit of the key range may be narrower
    const bool bit = (d >> b) & 1;
    const unsigned bal = __ballot_sync(vm, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// per-tile state word: {flag (2 bits) | count}; flag 1 = this tile's own count, 2 = inclusive prefix over tiles 0..t
// (tile 0: plus the global start of the digit).  32-bit words while every count fits 30 bits, else 64-bit.
template <typename ST> struct StateWord;
template <> struct StateWord<unsigned> {
  static constexpr int kShift = 30;
  static __device__ __forceinline__ unsigned ld(const unsigned* p) {
    unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
  }
  static __device__ __forceinline__ void st(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  }
};
template <> struct StateWord<unsigned long long> {
  static constexpr int kShift = 62;
  static __device__ __forceinline__ unsigned long long ld(const unsigned long long* p) {
    unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
  }
  static __device__ __forceinline__ void st(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
  }
};

struct PassArgs {                     // one sort's view of one pass
  const int32_t* keys_in; const int32_t* vals_in;
  int32_t* keys_out; int32_t* vals_out;
  const int32_t* ghist;               // [NB] counts of this pass's digit over all keys
  void* tile_state;                   // [ntiles][NB] state words of this pass
  int32_t* ticket;                    // tiles are handed out in launch order so every earlier tile is running or done
  const int32_t* unsorted_flag;
};

template <int DB>
constexpr int onesweep_smem_bytes() { return (kOsWarps * (1 << DB) + 2 * kOsTile + 2 * (1 << DB) + 16) * 4; }

template <int DB, typename ST>
__global__ void __launch_bounds__(kOsThreads, DB == 10 ? 3 : kOsMinCtas)   // 10-bit digits: 73 KB of shared memory, three CTAs fit anyway
k_onesweep_pass(PassArgs a0, PassArgs a1, int64_t n, int shift, int nbits) {
  const PassArgs& A = blockIdx.y == 0 ? a0 : a1;
  if (A.unsorted_flag && *A.unsorted_flag == 0) return;   // input already ordered: the sort is skipped
  constexpr int NB = 1 << DB;
  constexpr int DPT = NB / kOsThreads;               // digits per thread in the per-digit phase (consecutive digits)
  constexpr int kLook = 32 / DPT;                    // look-back window per digit
  using SW = StateWord<ST>;
  constexpr ST kFlagAgg = (ST)1 << SW::kShift, kFlagInc = (ST)2 << SW::kShift, kMask = ((ST)1 << SW::kShift) - 1;
  extern __shared__ int sm_os[];
  int (*wh)[NB] = reinterpret_cast<int (*)[NB]>(sm_os);      // per-warp digit counts, then per-warp start offsets
  int* const skeys = sm_os + kOsWarps * NB;                   // the tile grouped by digit (the ranking's peer masks before that)
  int* const svals = skeys + kOsTile;
  int* const s_lstart = svals + kOsTile;                      // start of digit d inside the grouped tile
  int* const s_gdelta = s_lstart + NB;                        // global position of element i of the grouped tile = s_gdelta[d] + i
  int* const s_scan = s_gdelta + NB;                          // [kOsWarps + 1] + the ticket
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_scan[12] = atomicAdd(A.ticket, 1);
  for (int i = tid; i < kOsWarps * NB; i += kOsThreads) { sm_os[i] = 0; skeys[i] = 0; }   // counts and peer masks
  __syncthreads();
  const int tile = s_scan[12];
  const int64_t tbase = (int64_t)tile * kOsTile;
  const int tile_n = (int)((n - tbase < kOsTile) ? (n - tbase) : kOsTile);
  const int dmask = NB - 1;

  // ---- load + stable rank inside the warp's 512-key slice (slice order = key order) ----
  int k[kOsKpt], v[kOsKpt];
  unsigned short rl[kOsKpt];
  const int wbase = warp * (32 * kOsKpt) + lane;
#pragma unroll
  for (int r = 0; r < kOsKpt; ++r) {
    const int i = wbase + r * 32;
    const bool valid = i < tile_n;
    k[r] = valid ? A.keys_in[tbase + i] : 0;
    v[r] = valid ? A.vals_in[tbase + i] : 0;
  }
  if (kRankAtomicOr && nbits > 5) {   // narrow digits: a few ballots are cheaper
    // peer mask through shared memory: every lane ORs its bit into the word of its digit, then reads the word back
    // (an emulated match.any on the LSU pipe instead of one vote per digit bit on the ADU pipe).  Letting every 2nd /
    // 3rd / 4th round take the hardware MATCH.ANY, to use both pipes side by side, changed nothing (0.193-0.199 ms per
    // build against 0.193 without).
    unsigned* mk = reinterpret_cast<unsigned*>(skeys) + warp * NB;
#pragma unroll
    for (int r = 0; r < kOsKpt; ++r) {
      const bool valid = wbase + r * 32 < tile_n;
      const int d = (k[r] >> shift) & dmask;
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
        const int d = (k[r] >> shift) & dmask;
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

  // ---- per digit (thread = DPT consecutive digits): offsets over warps, tile counts, position inside the tile ----
  const int d0 = tid * DPT;
  int cnt[DPT];
  {
    int tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) cnt[j] = 0;
#pragma unroll
    for (int w = 0; w < kOsWarps; ++w) {
      if constexpr (DPT == 4) {
        int4 c = *reinterpret_cast<int4*>(&wh[w][d0]);
        *reinterpret_cast<int4*>(&wh[w][d0]) = make_int4(cnt[0], cnt[1], cnt[2], cnt[3]);
        cnt[0] += c.x; cnt[1] += c.y; cnt[2] += c.z; cnt[3] += c.w;
      } else {
#pragma unroll
        for (int j = 0; j < DPT; ++j) { const int c = wh[w][d0 + j]; wh[w][d0 + j] = cnt[j]; cnt[j] += c; }
      }
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) tsum += cnt[j];
    int total;
    int lstart = block_exclusive_scan<kOsThreads>(tsum, total, s_scan);
#pragma unroll
    for (int j = 0; j < DPT; ++j) { s_lstart[d0 + j] = lstart; lstart += cnt[j]; }
  }
  ST* const state = static_cast<ST*>(A.tile_state);
  ST* const mine = state + (int64_t)tile * NB + d0;
  long long excl[DPT];
  if (tile == 0) {
    // global start of every digit = exclusive prefix of the histogram; folded into the inclusive prefix of tile 0
    int g[DPT], gsum = 0, total;
#pragma unroll
    for (int j = 0; j < DPT; ++j) { g[j] = A.ghist[d0 + j]; gsum += g[j]; }
    int gb = block_exclusive_scan<kOsThreads>(gsum, total, s_scan);
#pragma unroll
    for (int j = 0; j < DPT; ++j) { excl[j] = gb; gb += g[j]; SW::st(mine + j, kFlagInc | (ST)(excl[j] + cnt[j])); }
  } else {
#pragma unroll
    for (int j = 0; j < DPT; ++j) SW::st(mine + j, kFlagAgg | (ST)cnt[j]);
  }
  __syncthreads();   // s_lstart, wh offsets

  // ---- group the tile by digit in shared memory (the registers are free after this) ----
#pragma unroll
  for (int r = 0; r < kOsKpt; ++r) {
    if (wbase + r * 32 < tile_n) {
      const int d = (k[r] >> shift) & dmask;
      const int pos = s_lstart[d] + wh[warp][d] + rl[r];
      skeys[pos] = k[r];
      svals[pos] = v[r];
    }
  }

  // ---- look-back: how many keys with my digits do the tiles before mine hold? ----
  if (tile != 0) {
    // A digit this tile does not hold needs no position: its thread leaves the aggregate word (count 0) in place and
    // later tiles simply walk past it.  (The top digit of a 20-bit key range has 13 live values of 256.)  Every 64th
    // tile walks for all digits and publishes inclusive prefixes, which bounds the walk for a digit that is rare.
    int t[DPT]; bool done[DPT]; bool all = true;
#pragma unroll
    for (int j = 0; j < DPT; ++j) { t[j] = tile - 1; done[j] = cnt[j] == 0 && (tile & 63) != 0; excl[j] = 0; all &= done[j]; }
    bool skip[DPT];
#pragma unroll
    for (int j = 0; j < DPT; ++j) skip[j] = done[j];
    bool blocked = false;
    while (!all) {
      if (blocked) __nanosleep(64);                  // a predecessor had not published: do not steal issue slots
      blocked = false;
      ST sv[DPT][kLook];
#pragma unroll
      for (int j = 0; j < DPT; ++j)
#pragma unroll
        for (int w = 0; w < kLook; ++w)
          sv[j][w] = (!done[j] && t[j] - w >= 0) ? SW::ld(state + (int64_t)(t[j] - w) * NB + d0 + j) : kFlagInc;
      all = true;
#pragma unroll
      for (int j = 0; j < DPT; ++j) {
#pragma unroll
        for (int w = 0; w < kLook; ++w) {
          if (done[j]) break;
          const ST fl = sv[j][w] >> SW::kShift;
          if (fl == 0) { blocked = true; break; }  // not published yet: poll again from this tile
          excl[j] += (long long)(sv[j][w] & kMask);
          --t[j];
          if (fl == 2) done[j] = true;
        }
        all &= done[j];
      }
    }
#pragma unroll
    for (int j = 0; j < DPT; ++j) if (!skip[j]) SW::st(mine + j, kFlagInc | (ST)(excl[j] + cnt[j]));
  }
#pragma unroll
  for (int j = 0; j < DPT; ++j) s_gdelta[d0 + j] = (int)excl[j] - s_lstart[d0 + j];
  __syncthreads();

  // ---- write runs of equal digits with consecutive threads ----
#pragma unroll
  for (int j = 0; j < kOsKpt; ++j) {
    const int i = j * kOsThreads + tid;
    if (i < tile_n) {
      const int key = skeys[i];
      const int pos = s_gdelta[(key >> shift) & dmask] + i;
      A.keys_out[pos] = key;
      A.vals_out[pos] = svals[i];
    }
  }
}

// rowptr[k] = first position whose key is >= k, from the sorted keys (the input itself when it was already ordered):
// rowptr == cumsum(bincount(keys)) without a single atomic.  rowptr[N] = n.  Both sorts in one launch (blockIdx.y);
// when the keys were already non-decreasing the stable sort is the identity and the values are copied here.
struct RowptrArgs {
  const int32_t* keys_if_sorted; const int32_t* keys_after_sort; const int32_t* vals0;
  int32_t* vals_out; int32_t* rowptr; const int32_t* unsorted_flag;
};
__global__ void __launch_bounds__(256)
k_rowptr_from_sorted(RowptrArgs a0, RowptrArgs a1, int64_t n, int32_t N) {
  // Position i writes rowptr[k] = i for every k in (keys[i-1], keys[i]].  Runs of absent keys are normally short; a long
  // one (trailing isolated nodes, membership lists whose keys only span the graphs) is parked in shared memory and
  // filled by the whole CTA, so no single thread ever writes more than kInline entries in a row.  A thread owns kPer
  // consecutive positions (two 128-bit loads; one position per thread and round was latency-bound: 17 us per array
  // for 4.1 M keys).
  constexpr int kInline = 16, kMaxGaps = 64, kPer = 8;
  __shared__ int g_lo[kMaxGaps], g_hi[kMaxGaps], g_val[kMaxGaps];
  __shared__ int g_n;
  const RowptrArgs& A = blockIdx.y == 0 ? a0 : a1;
  const bool was_sorted = *A.unsorted_flag == 0;
  const int32_t* __restrict__ keys = was_sorted ? A.keys_if_sorted : A.keys_after_sort;
  int32_t* __restrict__ rowptr = A.rowptr;
  // sections of the workspace / CSR object are 256-byte aligned whenever the caller's buffers are 16-byte aligned
  const bool vec = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(A.vals0) |
                     reinterpret_cast<uintptr_t>(A.vals_out)) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * kPer;
  const int64_t rounds = ceil_div<int64_t>(n + 1, stride);
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t i0 = r * stride + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * kPer;
    int kv[kPer + 1];                               // kv[j] = keys[i0 + j - 1], with keys[-1] = -1 and keys[n] = N
    kv[0] = (i0 == 0) ? -1 : (i0 - 1 < n ? keys[i0 - 1] : N);
    if (i0 + kPer <= n && vec) {                    // whole strip inside the array
      const int4 q0 = *reinterpret_cast<const int4*>(keys + i0), q1 = *reinterpret_cast<const int4*>(keys + i0 + 4);
      kv[1] = q0.x; kv[2] = q0.y; kv[3] = q0.z; kv[4] = q0.w; kv[5] = q1.x; kv[6] = q1.y; kv[7] = q1.z; kv[8] = q1.w;
      if (was_sorted) {
        const int4 v0 = *reinterpret_cast<const int4*>(A.vals0 + i0), v1 = *reinterpret_cast<const int4*>(A.vals0 + i0 + 4);
        *reinterpret_cast<int4*>(A.vals_out + i0) = v0;
        *reinterpret_cast<int4*>(A.vals_out + i0 + 4) = v1;
      }
    } else {
#pragma unroll
      for (int j = 1; j <= kPer; ++j) kv[j] = (i0 + j - 1 < n) ? keys[i0 + j - 1] : N;
      if (was_sorted) {
#pragma unroll
        for (int j = 0; j < kPer; ++j) if (i0 + j < n) A.vals_out[i0 + j] = A.vals0[i0 + j];
      }
    }
    bool big = false;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (i0 + j > n) break;
      const int lo = kv[j], hi = kv[j + 1];
      if (hi - lo > kInline) big = true;
      else for (int kk = lo + 1; kk <= hi; ++kk) rowptr[kk] = (int32_t)(i0 + j);
    }
    if (!__syncthreads_or(big)) continue;          // the common case: no long run in this CTA's slice
    if (threadIdx.x == 0) g_n = 0;
    __syncthreads();
    if (big) {
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        if (i0 + j > n) break;
        const int lo = kv[j], hi = kv[j + 1];
        if (hi - lo <= kInline) continue;
        const int slot = atomicAdd(&g_n, 1);
        if (slot < kMaxGaps) { g_lo[slot] = lo; g_hi[slot] = hi; g_val[slot] = (int32_t)(i0 + j); }
        else for (int kk = lo + 1; kk <= hi; ++kk) rowptr[kk] = (int32_t)(i0 + j);   // list full: serial fill
      }
    }
    __syncthreads();
    const int ng = min(g_n, kMaxGaps);
    for (int q = 0; q < ng; ++q)
      for (int kk = g_lo[q] + 1 + threadIdx.x; kk <= g_hi[q]; kk += blockDim.x) rowptr[kk] = g_val[q];
    __syncthreads();
  }
}

struct SortWs {            // per sort
  int32_t* ghist;          // [npass][bins] (kHistInts)
  int32_t* tickets;        // [kMaxPass]
  char* tile_state;        // [npass][ntiles][bins] state words
  int32_t *kA, *vA, *kB, *vB;
};

template <int DB, typename ST>
static int launch_pass(const PassArgs& a0, const PassArgs& a1, int64_t n, int nsort, int shift, int nbits, cudaStream_t s) {
  constexpr int smem = onesweep_smem_bytes<DB>();
  SLDM_OPT_IN_SMEM((k_onesweep_pass<DB, ST>), smem);
  const dim3 grid((unsigned)ceil_div<int64_t>(n, kOsTile), (unsigned)nsort);
  k_onesweep_pass<DB, ST><<<grid, kOsThreads, smem, s>>>(a0, a1, n, shift, nbits);
  SLDM_LAUNCH_CHECK("k_onesweep_pass");
  return SLDM_OK;
}

// both sorts pass by pass: sort i takes (keys0[i], vals0[i]) -> final_vals[i] (stable by key); the sorted keys end in
// sorted_keys_out[i] (its kA or kB)
static int radix_sort_pairs2(int nsort, const int32_t* const keys0[2], const int32_t* const vals0[2], int64_t n,
                             const SortPlan& P, const SortWs W[2], int32_t* const final_vals[2],
                             const int32_t* sorted_keys_out[2], const int32_t* const unsorted_flag[2], cudaStream_t s) {
  const int64_t ntiles = ceil_div<int64_t>(n, kOsTile);
  const int NB = 1 << P.db;
  const bool wide = wide_state(n);
  const int64_t word = wide ? 8 : 4;
  const int32_t* in_k[2] = {keys0[0], keys0[1]};
  const int32_t* in_v[2] = {vals0[0], vals0[1]};
  for (int p = 0; p < P.npass; ++p) {
    const bool last = (p == P.npass - 1);
    PassArgs a[2];
    for (int i = 0; i < 2; ++i) {
      const int j = i < nsort ? i : 0;
      a[i].keys_in = in_k[j]; a[i].vals_in = in_v[j];
      a[i].keys_out = (p & 1) ? W[j].kB : W[j].kA;
      a[i].vals_out = last ? final_vals[j] : ((p & 1) ? W[j].vB : W[j].vA);
      a[i].ghist = W[j].ghist + p * NB;
      a[i].tile_state = W[j].tile_state + (int64_t)p * ntiles * NB * word;
      a[i].ticket = W[j].tickets + p;
      a[i].unsorted_flag = unsorted_flag[j];
    }
    const int nbits = std::min(P.db, P.key_bits - P.db * p);
    int rc;
    if (P.db == 10) rc = wide ? launch_pass<10, unsigned long long>(a[0], a[1], n, nsort, P.db * p, nbits, s)
                              : launch_pass<10, unsigned>(a[0], a[1], n, nsort, P.db * p, nbits, s);
    else            rc = wide ? launch_pass<8, unsigned long long>(a[0], a[1], n, nsort, P.db * p, nbits, s)
                              : launch_pass<8, unsigned>(a[0], a[1], n, nsort, P.db * p, nbits, s);
    if (rc) return rc;
    for (int i = 0; i < nsort; ++i) { in_k[i] = a[i].keys_out; in_v[i] = a[i].vals_out; }
  }
  for (int i = 0; i < nsort; ++i) sorted_keys_out[i] = in_k[i];
  return SLDM_OK;
}

// ----------------------------------------------------------- hub work list --
struct HubArgs { const int32_t* rowptr; int32_t* hub_list; int32_t* counter; };
__global__ void __launch_bounds__(256)
k_plan_hubs(HubArgs a0, HubArgs a1, int32_t N, int32_t cap) {
  const HubArgs& A = blockIdx.y == 0 ? a0 : a1;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int deg = A.rowptr[i + 1] - A.rowptr[i];
  if (deg <= SLDM_HUB_DEGREE) return;
  int nch = (deg + SLDM_HUB_CHUNK - 1) / SLDM_HUB_CHUNK;
  int k = atomicAdd(A.counter, nch);
  if (k + nch > cap) return;  // cannot happen: cap >= E/CHUNK + E/HUB_DEGREE
  for (int c = 0; c < nch; ++c) {
    int4 ent = make_int4((int)i, c, nch, k);
    reinterpret_cast<int4*>(A.hub_list)[k + c] = ent;
  }
}

}  // namespace sldm

// ================================================================== C ABI ==
using namespace sldm;

// workspace: int32 copies of the two index rows, per sort two ping-pong (key, value) pairs, and the sort state.  The
// sort state (histograms, tickets, per-tile look-back words for every pass of both sorts) is one contiguous region
// that a single memset clears.
struct CsrWs { int64_t src32, dst32, pp[2][4], state, state_bytes, total; int64_t ntiles; SortPlan plan; int64_t word; };
static int64_t sort_state_bytes(int64_t ntiles, const SortPlan& P, int64_t word) {
  return align_bytes(kHistInts * 4) + align_bytes(kMaxPass * 4) + align_bytes((int64_t)P.npass * ntiles * (1 << P.db) * word);
}
static CsrWs csr_ws_layout(int64_t N, int64_t E) {
  CsrWs w; int64_t o = 0;
  int64_t e = align_bytes((E > 0 ? E : 1) * 4);
  w.ntiles = ceil_div<int64_t>(E > 0 ? E : 1, kOsTile);
  w.plan = sort_plan(N);
  w.word = wide_state(E) ? 8 : 4;
  w.src32 = o; o += e;  w.dst32 = o; o += e;
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) { w.pp[i][j] = o; o += e; }
  w.state = o; w.state_bytes = 2 * sort_state_bytes(w.ntiles, w.plan, w.word); o += w.state_bytes;
  w.total = o;
  return w;
}
static SortWs sort_ws_at(char* wb, const CsrWs& W, int i) {
  SortWs S;
  char* base = wb + W.state + i * (W.state_bytes / 2);
  S.ghist = reinterpret_cast<int32_t*>(base);
  char* p = base + align_bytes(kHistInts * 4);
  S.tickets = reinterpret_cast<int32_t*>(p);
  p += align_bytes(kMaxPass * 4);
  S.tile_state = p;
  S.kA = reinterpret_cast<int32_t*>(wb + W.pp[i][0]);  S.vA = reinterpret_cast<int32_t*>(wb + W.pp[i][1]);
  S.kB = reinterpret_cast<int32_t*>(wb + W.pp[i][2]);  S.vB = reinterpret_cast<int32_t*>(wb + W.pp[i][3]);
  return S;
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
  SLDM_CUDA(cudaMemsetAsync(wb + W.state, 0, (size_t)W.state_bytes, s));
  const SortWs Wsort[2] = {sort_ws_at(wb, W, 0), sort_ws_at(wb, W, 1)};
  const SortPlan& P = W.plan;
  const bool membership = (edge_src == nullptr);   // only (rowptr_dst, col_src) are produced; source ids may exceed N
  const int nsort = membership ? 1 : 2;

  {
    const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(E, 256 * 4), (int64_t)num_sms() * 8);
    const size_t smem = (size_t)P.npass * (1 << P.db) * 4;
    if (P.db == 10) k_convert_hist<10><<<grid, 256, smem, s>>>(edge_src, edge_dst, E, (int32_t)N, P.npass, src32, dst32, meta, Wsort[0].ghist);
    else            k_convert_hist<8><<<grid, 256, smem, s>>>(edge_src, edge_dst, E, (int32_t)N, P.npass, src32, dst32, meta, Wsort[0].ghist);
    SLDM_LAUNCH_CHECK("k_convert_hist");
    if (!membership) {
      const int hgrid = (int)std::min<int64_t>(ceil_div<int64_t>(E, 256 * 8), (int64_t)num_sms() * 8);
      if (P.db == 10) k_digit_hist<10><<<hgrid, 256, smem, s>>>(src32, E, P.npass, Wsort[1].ghist, meta + 3);
      else            k_digit_hist<8><<<hgrid, 256, smem, s>>>(src32, E, P.npass, Wsort[1].ghist, meta + 3);
      SLDM_LAUNCH_CHECK("k_digit_hist");
    }
  }
  // sort 0, by destination: keys = dst, payload = src -> col_src; sort 1, by source (transpose): keys = src,
  // payload = dst -> col_dst.  Each is skipped on the device when its keys are already ordered.
  const int32_t* const keys0[2] = {dst32, src32};
  const int32_t* const vals0[2] = {src32, dst32};
  int32_t* const final_vals[2] = {col_s, col_d};
  const int32_t* const flags[2] = {meta + 4, meta + 3};
  const int32_t* sorted[2] = {nullptr, nullptr};
  int rc;
  if ((rc = radix_sort_pairs2(nsort, keys0, vals0, E, P, Wsort, final_vals, sorted, flags, s))) return rc;
  {
    const RowptrArgs r0{dst32, sorted[0], src32, col_s, rp_d, meta + 4};
    const RowptrArgs r1{src32, sorted[1], dst32, col_d, rp_s, meta + 3};
    const dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(E + 1, 256 * 8), (int64_t)num_sms() * 8), (unsigned)nsort);
    k_rowptr_from_sorted<<<grid, 256, 0, s>>>(r0, membership ? r0 : r1, E, (int32_t)N);
    SLDM_LAUNCH_CHECK("k_rowptr_from_sorted");
  }
  {
    const int cap = (int)hub_capacity(E);
    const HubArgs h0{rp_d, hub_d, meta + 0}, h1{rp_s, hub_s, meta + 1};
    const dim3 grid((unsigned)ceil_div<int64_t>(N, 256), (unsigned)nsort);
    k_plan_hubs<<<grid, 256, 0, s>>>(h0, membership ? h0 : h1, (int32_t)N, cap);
    SLDM_LAUNCH_CHECK("k_plan_hubs");
  }
  return SLDM_OK;
}

// sage_tc.cu -- tensor-core (tcgen05) GEMM engine of the SageBlock layer, two uses:
//   FWD   : z = agg W_l^T + b_l + x W_r^T ; out = act(LN(z))   (same contract as k_sage_proj_fwd, sage_fwd.cu;
//           src/models/blocks/sageblock.py:18-19)
//   DGRAD : dagg = (dz W_l) / max(deg,1) ; dxroot = dz W_r      (the data-gradient GEMMs of sage_bwd.cu)
//
// Why it can be used under an fp32 parity bar.  Plain TF32 keeps 11 significand bits
// (measured: the hardware TRUNCATES fp32 operands, tools/tc_probe.cu) -- 1e-3 errors.  Here
// every operand is split a = a_hi + a_lo (a_hi = what the hardware keeps, a_lo = a - a_hi,
// rounded to tf32) and three products a_lo*b_hi + a_hi*b_lo + a_hi*b_hi are issued, the small
// ones FIRST.  The tensor core adds into its fp32 accumulator with round-toward-zero
// (measured bias), so the TMEM accumulator only ever holds ONE 32-wide K chunk (12 MMAs) and
// is flushed into fp32 registers on the CUDA cores (round-to-nearest adds).  Measured against
// fp64 (tools/tc_probe2.cu, K=128): max 3.2e-7 / rms 7.3e-8 versus 7.8e-7 / 1.1e-7 for a
// sequential fp32 FMA chain -- i.e. at least as accurate as the SIMT kernel.
//
// Pipeline (one persistent CTA per SM, 896 threads = 7 warpgroups, 128 rows x Fout per tile; setmaxnreg 56 / 72 / 80 / 56
// of the 72 x 896 register pool):
//   warp 0        TMA producer A: raw activation chunks [128 x 32] (agg, then x) into a 3-deep ring -- the HBM stream;
//                                a stage is handed back by the CONVERTER as soon as the tile is in registers
//   warp 2        TMA producer B: pre-split weight tiles B_hi, B_lo [Fout x 32] of the chunk into a 3-deep ring (L2 hits)
//   warps 4-7     converter    : reads its row of the raw tile and parks a (the tensor core truncates it to tf32 = a_hi
//                                itself) and a_lo = rna_tf32(a - a_hi) in TENSOR MEMORY with tcgen05.st (lane = row,
//                                column = k; two 64-column slots with their own free barriers) -- the MMAs take A from
//                                TMEM, so shared memory only carries the raw tile once and the weights
//   warps 1, 3    MMA issuers  : alternate K chunks; 12 x tcgen05.mma.kind::tf32 (M=128, N=Fout, K=8), A from TMEM, B
//                                from smem descriptors, into one of three TMEM accumulators; tcgen05.commit frees the
//                                weight stage and the TMEM A slot and signals the drain.  Warp-uniform control flow,
//                                one elected lane issues (12 back-to-back UTCHMMA, operands on the uniform datapath)
//   warps 8-23    drain        : tcgen05.ld the chunk accumulator (thread = row x quarter of the columns), add into
//                                registers (round to nearest); after the last chunk: + bias, LayerNorm statistics (thread
//                                per row x quarter: one warp instruction serves 32 rows; one exchange of per-quarter
//                                (sum, M2) through smem, Chan's merge) -- DGRAD: / max(deg,1) -- and the normalised row
//                                is PARKED in a swizzled shared-memory tile [128 x Fout]; on to the next tile's chunks
//   warps 24-27   finisher     : one warp per 32-row quadrant of the parked tile, LANE = 4 CONSECUTIVE COLUMNS: xhat out
//                                (training), affine + (Leaky)ReLU, out -- every store instruction writes whole 512-byte
//                                rows; off the MMA -> drain critical path (DGRAD: plain copy to dagg | dxroot)
// Tensor memory (512 columns): three accumulators [0, 384), two A slots [384, 512).
// Bound: HBM at large N (N*4*(2Fin + 2Fout) bytes).  Measured on B200, batch shape (0.82 M rows, 128 -> 128), per
// 128-row tile in SM cycles (clock64 traces under profiles/r02_trace_*; HBM floor ~8.9k at the ~1.5 GHz the SMs run at):
//   round 1 (epilogue warps do statistics + two TMA-store pushes)                         14.7k   0.392 ms
//   + finisher warpgroup (stores leave the drain), still `if (lane == 0)` MMA issue         13.8k   0.380 ms
//     -> trace: 150 cycles per tcgen05.mma: under a divergent branch ptxas wraps every MMA in an ELECT / 4x R2UR /
//        BRA.U.ANY broadcast loop (15 SASS instructions each); the kernel was bound by MMA ISSUE
//   + warp-uniform issue with an elected lane (this file)                                   12.3k   0.355 ms
//   + four rows in flight per finisher warp                                                         0.349 ms
// What was tried and lost: 256-bit st.global straight from the accumulator registers (lane = row, 32 B per lane:
// partial-line writes, 0.507 ms); LayerNorm in a lane-per-column finisher (2.9k cycles per 4 rows of shuffles and
// divisions: 0.587 ms) and in a thread-per-row finisher over the parked tile (0.470 ms: four warps cannot absorb it);
// deeper rings (A 3..8, B 2..3: no change); per-quarter statistics in the drain + merge and normalisation in the finisher
// (the drain's tail drops to 0.5k cycles but the finisher, whose stores already take ~11k per tile, becomes the pace
// maker: 0.400 ms); an L2 prefetch of the next tile's rows (0.370 ms); mbarrier suspend-time hints (no change).  Timing-only experiments (wrong results, -DSLDM_TC_NOSTORE / weights loaded
// once): without the 128 KB of stores per tile 0.300 ms (the finisher's 64 STG.128 per warp and tile take ~11k cycles
// here, with or without TMA bulk stores; a stand-alone probe of the same pattern, tools/store_probe.cu, writes 6.2 TB/s
// from two warps per SM and 2.5 + 2.5 TB/s next to an equal load stream -- so it is this kernel's own operand traffic
// that slows them: 2.4 TB/s of HBM reads + 4.6 TB/s of weight re-reads from L2 + 2.2 TB/s of stores = 9.2 TB/s through
// the L2 <-> SM fabric), without the weight re-streaming 0.339 ms, without both 0.278 ms (then the drain tail --
// statistics 2.6k + park 0.8k during which the MMA stream runs out of accumulators -- sets the pace).  Three limits sit
// within 10 % of each other; the next step is a CTA pair sharing the weight stages (cluster multicast).
#include "common.cuh"
#include "tc_common.cuh"

namespace sldm {
using namespace tc;

constexpr int kTcThreads = 896;   // 7 warpgroups: {TMA, MMA, alloc+TMA B, MMA} {converter x4} {drain x4} x 4 {finisher x4}
constexpr int kWgThreads = 512;   // k_wgrad_tc: {TMA, MMA, alloc, idle} {converter x4} {epilogue x4} x 2
// Ring depths.  Round 1 ran A 5 / B 2; once the MMA issue was fixed (round 2) the weight ring became the pace maker of
// DGRAD: a stage is free only when the MMAs that read it have completed, the reload then takes ~2k cycles (L2 hit, but
// queued behind the operand loads), so with two stages a chunk took (0.8k MMA + 2k reload) / 2.  A 3 / B 3 (the same
// 209 KB): dgrad 0.329 -> 0.302 ms, forward 0.357 -> 0.351 ms at the batch shape.  (A 4 / B 3 does not fit next to the
// 64 KB hand-off tile.)
#ifndef SLDM_TC_STAGES
#define SLDM_TC_STAGES 3
#endif
#ifndef SLDM_TC_BSTAGES
#define SLDM_TC_BSTAGES 3
#endif
constexpr int kTcStages = SLDM_TC_STAGES;     // A ring: raw activation chunks [128 x 32] (16 KB each) -- the HBM operand, prefetched deep
constexpr int kTcBStages = SLDM_TC_BSTAGES;   // B ring: weight chunks B_hi | B_lo (32 KB each at Fout = 128; L2 resident)
constexpr int kTcBM = 128;
constexpr int kTcAcc = 3;          // TMEM accumulator ring (one K chunk each), columns [0, 3*32*NT): look-ahead of the MMA stream
constexpr int kTcASlots = 2;       // TMEM A slots (own ring, own "free" barriers: shorter than the smem stage ring)
constexpr int kTcZBytes = kTcBM * 128 * 4;     // hand-off tile drain -> finisher: [128 rows][<= 128 fp32], chunk-swizzled
constexpr int kTcACol0 = 384;      // TMEM columns of the A ring: kTcASlots x {A_hi[32] | A_lo[32]}
constexpr int MODE_FWD = 0, MODE_DGRAD = 1;

// One work item = (128-row tile, output group).  Its K loop runs over nsrc A sources x Kc 32-wide chunks.
//   FWD  : nsrc = 2 (agg, x), ngroups = 1, Nout = Fout, Kc = Fin/32
//   DGRAD: nsrc = 1 (dz),     ngroups = 2 (W_l -> dagg, W_r -> dxroot), Nout = Fin, Kc = Fout/32
// Weight tiles come from one packed array [4 * Nout][K]: block (2*(g*nsrc+src) + lo) holds the hi / lo part.
struct TcProblem { int64_t N; int Kc, nsrc, ngroups, Nout; int rev = 0; };   // rev: tiles are walked from the last row block to the first

__global__ void __launch_bounds__(256)
k_split_weights(const float* __restrict__ W_l, const float* __restrict__ W_r, int count, float* __restrict__ out) {
  // out = [W_l hi | W_l lo | W_r hi | W_r lo], each `count` floats; hi = rna_tf32(w), lo = rna_tf32(w - hi)
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * count; i += gridDim.x * blockDim.x) {
    const float w = i < count ? W_l[i] : W_r[i - count];
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(w));
    const float hf = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(w - hf));
    const int base = i < count ? i : (i - count) + 2 * count;
    out[base] = hf;
    out[base + count] = __uint_as_float(l);
  }
}

__device__ __forceinline__ float tf32_lo(float a) {
  const float hi = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);  // what the tensor core keeps
  // a - hi is exact; round it to tf32 (nearest, ties away) with integer ops: add half an ulp, truncate
  return __uint_as_float((__float_as_uint(a - hi) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// debug timeline (SLDM_TC_TRACE=<file>): CTA 0 stamps clock64() per role / chunk / event
constexpr int kTraceIts = 96, kTraceEv = 32;
#define TC_TRACE(ev, itv) \
  do { if (trace != nullptr && blockIdx.x == 0 && (itv) < kTraceIts) trace[(itv) * kTraceEv + (ev)] = clock64(); } while (0)

template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// Hand-off tile in shared memory, laid out so that the TMA engine can store it as it lies: NT column slabs of
// [128 rows][128 bytes] (32 fp32 columns each), 128-byte swizzle -- 16-byte chunk c of row r sits at position
// c ^ (r & 7) of its 128-byte row.  j = 16-byte chunk index inside the full row (0 .. 8*NT-1).  Drain writes (8 lanes
// of a quarter warp = 8 consecutive rows, same j) and finisher reads (the lanes of a row = consecutive j) both touch
// every bank group exactly once per phase.
template <int NT>
__device__ __forceinline__ uint32_t z_chunk_addr(uint32_t zbase, int r, int j) {
  return zbase + (uint32_t)(j >> 3) * (kTcBM * 128u) + (uint32_t)r * 128u + ((uint32_t)((j & 7) ^ (r & 7)) << 4);
}

struct DrainArgs {
  int64_t N; int Fout; int64_t ntiles; int nchunks; int ngroups; float eps; uint32_t tmem_base; uint32_t zbase;
  float* rstd; const int32_t* rowptr; const float* s_bias; float* s_sum; float* s_var;
  uint64_t* bar_acc_full; uint64_t* bar_acc_empty; uint64_t* bar_z_full; uint64_t* bar_z_empty; long long* trace;
  int64_t tbase; int tstep;      // row block of loop index t = tbase + tstep * t (walk direction)
};

// Drain role: 16 warps (512 threads).  Thread (quadrant q, lane, column quarter CQ) owns tile row q*32+lane and the
// columns [CQ*8*NT, (CQ+1)*8*NT): one tcgen05.ld round trip per K chunk, fp32 register accumulation.  After the last
// chunk: FWD + bias and the LayerNorm statistics (thread per row x quarter: every warp instruction serves 32 rows --
// a lane-per-column finisher spent 2.9k cycles per 4 rows on the divisions and shuffles of the statistics alone),
// DGRAD the division by the in-degree; the normalised row is parked for the finisher.  FULL = (Fout == 32*NT).
template <int NT, bool FULL, int MODE>
__device__ __forceinline__ void drain_role(const DrainArgs a) {
  constexpr int HC = 8 * NT;               // columns per thread
  constexpr int ACC_COLS = 32 * NT;
  long long* trace = a.trace;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3;                  // TMEM lane quadrant of this warp
  const int rloc = q * 32 + lane;
  const int cq = (warp - 8) >> 2;          // column quarter
  const int c_lo = cq * HC;
  const uint32_t tq = a.tmem_base + ((uint32_t)(q * 32) << 16) + c_lo;
  const float fF = (float)a.Fout;
  const int Fout = a.Fout;
  const float* const bias = a.s_bias + c_lo;
  uint32_t it = 0, hand = 0;
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    for (int grp = 0; grp < a.ngroups; ++grp, ++hand) {
      float z[HC];
      for (int c = 0; c < a.nchunks; ++c, ++it) {
        const uint32_t ab = it % kTcAcc, aph = (it / kTcAcc) & 1;
        if (tid == 256) TC_TRACE(8, it);
        mbar_wait(&a.bar_acc_full[ab], aph);
        if (tid == 256) TC_TRACE(9, it);
        tc_fence_after();
        uint32_t rr[NT][8];
#pragma unroll
        for (int g0 = 0; g0 < NT; ++g0) tmem_ld_32x8(tq + ab * ACC_COLS + g0 * 8, rr[g0]);
        tmem_ld_wait();
        tc_fence_before();                   // values are in registers: release the accumulator
        __syncwarp();
        if (lane == 0) mbar_arrive(&a.bar_acc_empty[ab]);
        if (tid == 256) TC_TRACE(10, it);
        if (c == 0) {
#pragma unroll
          for (int g0 = 0; g0 < NT; ++g0)
#pragma unroll
            for (int j = 0; j < 8; ++j) z[g0 * 8 + j] = __uint_as_float(rr[g0][j]);
        } else {
#pragma unroll
          for (int g0 = 0; g0 < NT; ++g0)
#pragma unroll
            for (int j = 0; j < 8; ++j) z[g0 * 8 + j] += __uint_as_float(rr[g0][j]);
        }
      }
      if (tid == 256) TC_TRACE(11, it - 1);
      const int64_t row = (a.tbase + a.tstep * tile) * kTcBM + rloc;
      if constexpr (MODE == MODE_FWD) {
        // ---- bias + LayerNorm statistics; the four column quarters of a row live in warps 8+q, 12+q, 16+q, 20+q.
        //      Each quarter computes its own (sum, M2 about its own mean); ONE exchange through smem on a 128-thread
        //      named barrier and Chan's merge give the row mean and variance:
        //         M2 = sum_q M2_q + sum_q n_q (m_q - mean)^2        (two-pass accuracy, one barrier instead of two)
        float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j4 = 0; j4 < HC / 4; ++j4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * j4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = 4 * j4 + e;
            z[j] += bb[e];
            ps[e] += (FULL || c_lo + j < Fout) ? z[j] : 0.f;
          }
        }
        const int nq_i = FULL ? HC : max(0, min(HC, Fout - c_lo));     // valid columns of this quarter
        const float nq = (float)nq_i;
        const float sq = (ps[0] + ps[1]) + (ps[2] + ps[3]);
        const float mq = nq_i > 0 ? __fdiv_rn(sq, nq) : 0.f;
        float pv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < HC; ++j) {
          const float d = z[j] - mq;
          pv[j & 3] += (FULL || c_lo + j < Fout) ? d * d : 0.f;
        }
        // buffers alternate per tile: a warp can only be one barrier ahead of its three partners, so the values of
        // tile t are never overwritten (by tile t+2) before everybody has read them
        float* const sS = a.s_sum + (hand & 1) * 512;
        float* const sV = a.s_var + (hand & 1) * 512;
        sS[cq * 128 + rloc] = mq;                           // the quarter's MEAN (the merge needs means, not sums)
        sV[cq * 128 + rloc] = (pv[0] + pv[1]) + (pv[2] + pv[3]);
        named_bar_sync(2 + q, 128);
        float m4[4], msum = 0.f, m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          m4[k] = sS[k * 128 + rloc];
          const int nk_i = FULL ? HC : max(0, min(HC, Fout - k * HC));
          msum = fmaf((float)nk_i, m4[k], msum);            // sum_q n_q m_q
        }
        const float mean = __fdiv_rn(msum, fF);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int nk_i = FULL ? HC : max(0, min(HC, Fout - k * HC));
          const float dm = (nk_i > 0 ? m4[k] : mean) - mean;
          m2 += sV[k * 128 + rloc] + (float)nk_i * dm * dm;
        }
        const float rs = __fdiv_rn(1.f, __fsqrt_rn(__fdiv_rn(m2, fF) + a.eps));
        const float nmr = -mean * rs;
#pragma unroll
        for (int j = 0; j < HC; ++j) z[j] = fmaf(z[j], rs, nmr);     // z now holds xhat = (z - mean) * rstd
        if (row < a.N && a.rstd != nullptr && cq == 0) a.rstd[row] = rs;
        if (tid == 256) TC_TRACE(14, it - 1);
      } else {
        if (grp == 0) {
          float cnt = 1.f;
          if (row < a.N) {
            int deg = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
            deg = deg < 1 ? 1 : (deg > 16777216 ? 16777216 : deg);
            cnt = (float)deg;
          }
          // z / cnt as a reciprocal multiply with one FMA correction step: q = z*r, q += (z - q*cnt)*r.  Correctly
          // rounded in all but rare double-rounding cases (and exact for cnt = 2^k) at 3 instructions per element
          // instead of the ~10 of an IEEE division -- 32 divisions per thread were most of this tail.
          const float rc = __fdiv_rn(1.f, cnt);
#pragma unroll
          for (int j = 0; j < HC; ++j) {
            const float qv = z[j] * rc;
            z[j] = fmaf(fmaf(-qv, cnt, z[j]), rc, qv);
          }
        }
      }
      // ---- park the finished row for the finisher warp of this quadrant and move on
      mbar_wait(&a.bar_z_empty[q], (hand & 1) ^ 1);      // the previous tile of this quadrant has been read
      if (tid == 256) TC_TRACE(12, it - 1);
#pragma unroll
      for (int c = 0; c < 2 * NT; ++c)
        sts128(z_chunk_addr<NT>(a.zbase, rloc, cq * 2 * NT + c), make_float4(z[4 * c], z[4 * c + 1], z[4 * c + 2], z[4 * c + 3]));
      fence_proxy_async_smem();                           // the tile is also read by the TMA engine (bulk stores)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a.bar_z_full[q]);       // release: the finisher's wait acquires these writes
      if (tid == 256) TC_TRACE(13, it - 1);
    }
  }
}

struct FinArgs {
  int64_t N; int Fout; int64_t ntiles; int ngroups; float slope;
  float* out; float* xhat; uint32_t zbase;
  const float* gamma; const float* beta;       // global pointers (FWD)
  const CUtensorMap* tm_s0; const CUtensorMap* tm_s1;   // bulk-store maps, box [32 rows][32 fp32]: FWD xhat | DGRAD dagg, dxroot
  uint64_t* bar_z_full; uint64_t* bar_z_empty; long long* trace;
  int64_t tbase; int tstep;
};

// Finisher role: 4 warps, warp q owns rows [32q, 32q+32) of the parked tile.  A row is spread over LPRW = 8*NT lanes
// (lane = 4 consecutive columns), RPI = 32 / LPRW rows per instruction (NT = 3: 24 of 32 lanes active), U of them in
// flight.  FWD: the parked values are xhat: store it (training), then the affine + (Leaky)ReLU and the output store.
// DGRAD: plain copy to dagg | dxroot.  Every store instruction writes whole rows (512 B at Fout = 128).
template <int NT, int MODE, bool ABF>
__device__ __forceinline__ void finisher_role(const FinArgs a) {
  constexpr int LPRW = (NT == 1) ? 8 : (NT == 2) ? 16 : 32;
  constexpr int RPI = 32 / LPRW;
  #ifndef SLDM_TC_FIN_U
#define SLDM_TC_FIN_U 4
#endif
  constexpr int U = SLDM_TC_FIN_U;          // row instructions in flight per warp
  long long* trace = a.trace;
  const int tid = threadIdx.x, lane = tid & 31;
  const int q = (tid >> 5) & 3;
  const int lig = lane % LPRW, sub = lane / LPRW;
  const int Fout = a.Fout;
  const int c0 = 4 * lig;
  const bool cvalid = c0 < Fout;                      // Fout % 4 == 0: a lane's four columns are valid together
  float4 gam4 = make_float4(0.f, 0.f, 0.f, 0.f), bet4 = gam4;
  if (MODE == MODE_FWD && cvalid) {
    gam4 = __ldg(reinterpret_cast<const float4*>(a.gamma + c0));
    bet4 = __ldg(reinterpret_cast<const float4*>(a.beta + c0));
  }
  uint32_t hand = 0;
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    for (int grp = 0; grp < a.ngroups; ++grp, ++hand) {
      if (tid == 768) TC_TRACE(16, hand);
      mbar_wait(&a.bar_z_full[q], hand & 1);
      if (tid == 768) TC_TRACE(17, hand);
      float* const o_main = (MODE == MODE_FWD) ? a.out : (grp == 0 ? a.out : a.xhat);   // DGRAD: dagg | dxroot
      // Whatever leaves unchanged -- xhat in training, dagg / dxroot in DGRAD -- is stored by the TMA engine straight
      // from the parked tile (one [32 rows x 128 B] box per column slab): a warp's own st.global stream gets one
      // 512-byte store through every ~190 cycles, and 128 KB per tile through four warps was the slowest stage of
      // the kernel (profiles/r02_ab_kernel_variants.jsonl, "nostore").  Only the activated output takes that path.
      // (whole 32-column slabs only: Fout == 32*NT; other widths keep the st.global path.  Measured at the batch
      // shape: DGRAD 0.360 -> 0.353 ms; for the forward's xhat it was neutral to slightly worse (0.349 -> 0.357 ms: the
      // finisher then waits for the engine before handing the tile back), so the forward keeps st.global for both.)
      const bool bulk = (Fout == 32 * NT) && (MODE == MODE_DGRAD);
      if (bulk && lane == 0) {
        const CUtensorMap* tm = (MODE == MODE_FWD) ? a.tm_s0 : (grp == 0 ? a.tm_s0 : a.tm_s1);
        const int grow0 = (int)((a.tbase + a.tstep * tile) * kTcBM) + q * 32;
#pragma unroll
        for (int sl = 0; sl < NT; ++sl)
          if (sl * 32 < Fout)
            tma_store_2d_u32(tm, a.zbase + (uint32_t)sl * (kTcBM * 128u) + (uint32_t)q * 4096u, sl * 32, grow0);
        tma_store_commit();
      }
      if (MODE == MODE_DGRAD && bulk) {
        if (lane == 0) {
          tma_store_wait_read<0>();                     // the engine has read the tile: hand it back
          mbar_arrive(&a.bar_z_empty[q]);
        }
        __syncwarp();
        if (tid == 768) TC_TRACE(18, hand);
        continue;
      }
#pragma unroll 1
      for (int i = 0; i < 32; i += U * RPI) {
        float4 v[U];
        int rl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          rl[u] = q * 32 + i + u * RPI + sub;
          v[u] = cvalid ? lds128(z_chunk_addr<NT>(a.zbase, rl[u], lig)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (i + U * RPI >= 32) {
          // last rows of the quadrant are in registers: hand the tile back (the loads must have RETURNED first)
          float guard = 0.f;
#pragma unroll
          for (int u = 0; u < U; ++u) guard += v[u].x;
          asm volatile("" ::"f"(guard) : "memory");
          __syncwarp();
          if (lane == 0) {
            if (bulk) tma_store_wait_read<0>();          // ... and the engine has read it too
            mbar_arrive(&a.bar_z_empty[q]);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t row = (a.tbase + a.tstep * tile) * kTcBM + rl[u];
#ifdef SLDM_TC_NOSTORE   /* timing experiment only (no results): the finisher issues no global stores */
          if (row < 0 && cvalid) {
#else
          if (row < a.N && cvalid) {
#endif
            if constexpr (MODE == MODE_FWD) {
              if (!bulk && a.xhat != nullptr) *reinterpret_cast<float4*>(a.xhat + row * Fout + c0) = v[u];
              float4 y;
              y.x = fmaf(v[u].x, gam4.x, bet4.x); y.y = fmaf(v[u].y, gam4.y, bet4.y);
              y.z = fmaf(v[u].z, gam4.z, bet4.z); y.w = fmaf(v[u].w, gam4.w, bet4.w);
              y.x = y.x > 0.f ? y.x : a.slope * y.x; y.y = y.y > 0.f ? y.y : a.slope * y.y;
              y.z = y.z > 0.f ? y.z : a.slope * y.z; y.w = y.w > 0.f ? y.w : a.slope * y.w;
              if constexpr (ABF) {     // bf16 feature storage: the layer output is rounded once, here
                uint16_t* const ob = reinterpret_cast<uint16_t*>(o_main);
                *reinterpret_cast<uint2*>(ob + row * Fout + c0) = make_uint2(pack2_bf16(y.x, y.y), pack2_bf16(y.z, y.w));
              } else {
                *reinterpret_cast<float4*>(o_main + row * Fout + c0) = y;
              }
            } else {
              *reinterpret_cast<float4*>(o_main + row * Fout + c0) = v[u];
            }
          }
        }
      }
      if (tid == 768) TC_TRACE(18, hand);
    }
  }
  if (lane == 0) tma_store_wait_all<0>();   // all global writes of this warp's bulk stores are complete before the CTA exits
}

// ABF (forward only): the A sources (agg, x) are bf16 rows.  TMA moves them as 32-bit words (a raw chunk [128 x 32
// words] carries 64 K values), the converter unpacks a word into its two bf16 values -- both EXACT in tf32, so there
// is no a_lo: a TMEM slot holds the 64 K values of the chunk and every k-step costs two MMAs (a*b_lo, a*b_hi)
// instead of three; the weight tiles stay 32-K stages, two per raw chunk.  The output is stored as bf16.
template <int NT, int MODE, bool ABF>  // NT = ceil(Nout / 32) in 1..4
__global__ void __launch_bounds__(kTcThreads, 1)
k_sage_tc(const __grid_constant__ CUtensorMap tm_agg, const __grid_constant__ CUtensorMap tm_x,
          const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_s0,
          const __grid_constant__ CUtensorMap tm_s1, const TcProblem pb,
          const float* __restrict__ b_l, const float* __restrict__ gamma,
          const float* __restrict__ beta, float eps, float slope,
          float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd,
          const int32_t* __restrict__ rowptr, long long* __restrict__ trace) {
  const int64_t N = pb.N;
  const int Fout = pb.Nout;   // MMA N / output width of this problem
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[kTcStages], bar_empty[kTcStages];       // A ring: TMA -> converter -> TMA
  __shared__ uint64_t bar_bfull[kTcBStages], bar_bempty[kTcBStages];   // B ring: TMA -> MMA -> TMA
  __shared__ uint64_t bar_conv[kTcASlots], bar_afree[kTcASlots];       // TMEM A slots: converter -> MMA -> converter
  __shared__ uint64_t bar_acc_full[kTcAcc], bar_acc_empty[kTcAcc];
  __shared__ uint64_t bar_z_full[4], bar_z_empty[4];                   // hand-off tile, per 32-row quadrant: drain -> finisher -> drain
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float s_bias[128];
  __shared__ float s_sum[2][4][128], s_var[2][4][128];   // double buffered by tile parity (one barrier per tile)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = kTcBM * 128;
  const uint32_t b_bytes = (uint32_t)Fout * 128;
  // shared memory: A ring (kTcStages raw tiles) | B ring (kTcBStages x {B_hi, B_lo}) | hand-off tile [128 x 32*NT]
  uint8_t* const smem_b = smem + (size_t)kTcStages * a_bytes;
  uint8_t* const smem_z = smem_b + (size_t)kTcBStages * 2 * b_bytes;
  constexpr int ACC_COLS = 32 * NT;
  constexpr uint32_t TMEM_COLS = 512;                   // accumulators [0, 384) + A slots [384, 512)
  static_assert(kTcAcc * 128 <= kTcACol0 && kTcACol0 + 64 * kTcASlots <= 512, "TMEM budget");
  // Two MMA issuers alternate chunks; an mbarrier parity wait is only safe when the waiter cannot be two phases ahead.
  // B = 2: a warp always meets the same stage (chunk parity == stage).  B == kTcAcc: before a warp waits for stage
  // it % B it has waited for the accumulator of chunk it - kTcAcc to be drained, i.e. for the previous use of that very
  // stage to be complete.  Other depths would need per-warp stage tracking.
  static_assert(kTcBStages == 2 || kTcBStages == kTcAcc, "weight ring depth vs. the two-issuer parity waits");

  const int half = pb.Kc;
  const int nchunks = pb.nsrc * pb.Kc;
  const int ngroups = pb.ngroups;
  const int64_t ntiles = (N + kTcBM - 1) / kTcBM;

  if (MODE == MODE_FWD && tid < 128) s_bias[tid] = tid < Fout ? b_l[tid] : 0.f;
  if (tid == 0) {
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 4);       // released by the four converter warps once the raw tile is in registers
    }
    for (int s = 0; s < kTcBStages; ++s) {
      mbar_init(&bar_bfull[s], 1);
      mbar_init(&bar_bempty[s], 1);      // released by tcgen05.commit
    }
    for (int a = 0; a < kTcASlots; ++a) mbar_init(&bar_conv[a], 4);
    for (int a = 0; a < kTcASlots; ++a) mbar_init(&bar_afree[a], 1);
    for (int a = 0; a < kTcAcc; ++a) {
      mbar_init(&bar_acc_full[a], 1);
      mbar_init(&bar_acc_empty[a], 16);
    }
    for (int qd = 0; qd < 4; ++qd) {
      mbar_init(&bar_z_full[qd], 4);     // the four column-quarter drain warps of the quadrant
      mbar_init(&bar_z_empty[qd], 1);    // its finisher warp
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_agg);
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  // warpgroups 0 and 1 hand registers to the two epilogue warpgroups (setmaxnreg is per warpgroup)
  if (warp < 4) {
   reg_dec<56>();
   if (warp == 0) {
    // ---------------------------------------------------------- TMA producer: A --
    // raw activation chunks, kTcStages deep: this is the HBM stream, it keeps running while the epilogue finishes a tile
    {   // whole warp, uniform control flow; one elected lane issues (operands stay on the uniform datapath)
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = (int)((pb.rev ? ntiles - 1 - tile : tile) * kTcBM);
        for (int g = 0; g < ngroups; ++g) {
          for (int c = 0; c < nchunks; ++c, ++it) {
            const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1;
            if (lane == 0) TC_TRACE(0, it);
            mbar_wait(&bar_empty[s], ph ^ 1);
            if (lane == 0) TC_TRACE(1, it);
            const int src = c / half;
            const int k0 = (c - src * half) * 32;
            if (elect_one()) {
              mbar_expect_tx(&bar_full[s], a_bytes);
              tma_load_2d_u32(smem_u + s * a_bytes, src == 0 ? &tm_agg : &tm_x, k0, row0, &bar_full[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 2) {
    // ---------------------------------------------------------- TMA producer: B --
    // pre-split weight tiles B_hi, B_lo [Fout x 32] of the chunk (L2 resident; re-streaming them is not a limiter:
    // skipping these loads changed the kernel time by < 5% on B200)
    {
      const uint32_t smem_b_u = __shfl_sync(0xffffffffu, smem_u32(smem_b), 0);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int g = 0; g < ngroups; ++g) {
          for (int c = 0; c < nchunks * (ABF ? 2 : 1); ++c, ++it) {    // ABF: two 32-K weight stages per raw chunk
            const uint32_t s = it % kTcBStages, ph = (it / kTcBStages) & 1;
            if (lane == 0) TC_TRACE(29, it);
            mbar_wait(&bar_bempty[s], ph ^ 1);
            if (lane == 0) TC_TRACE(30, it);
            const uint32_t st = smem_b_u + s * 2u * b_bytes;
            const int cc = ABF ? (c >> 1) : c;
            const int src = cc / half;
            const int k0 = ABF ? ((cc - src * half) * 64 + 32 * (c & 1)) : ((cc - src * half) * 32);
            const int wrow = 2 * (g * pb.nsrc + src) * Fout;
            if (elect_one()) {
              mbar_expect_tx(&bar_bfull[s], 2 * b_bytes);
              tma_load_2d_u32(st, &tm_w, k0, wrow, &bar_bfull[s]);
              tma_load_2d_u32(st + b_bytes, &tm_w, k0, wrow + Fout, &bar_bfull[s]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ------------------------------------------------------------- MMA issuers --
    // Two issuing warps alternate K chunks (warp 1: even, warp 3: odd).  A chunk has its own accumulator, smem stage
    // and TMEM A slot, so the two streams are independent; while one sits in the barrier waits of its next chunk the
    // tensor pipe is fed by the other (one issuer left ~35% bubbles: profiles/r01c trace).
    // The WHOLE warp runs this loop in uniform control flow and one elected lane issues.  Round 1 ran it under
    // `if (lane == 0)`: every operand then lives in a per-thread register and ptxas wraps each tcgen05.mma in an
    // ELECT / 4x R2UR / BRA.U.ANY broadcast loop (15 SASS instructions per MMA; measured 150 cycles per MMA against
    // the 64-cycle tensor-pipe floor -- the kernel was bound by MMA ISSUE, profiles/r02 trace).  With warp-uniform
    // operands (tmem base through a lane-0 shuffle) they are computed on the uniform datapath.
    {
      const uint32_t my_parity = (warp == 3) ? 1u : 0u;
      const uint32_t idesc = make_idesc_tf32(kTcBM, Fout, 0, 0);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_b_u = __shfl_sync(0xffffffffu, smem_u32(smem_b), 0);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int c = 0; c < ngroups * nchunks; ++c, ++it) {
          const uint32_t sb = it % kTcBStages, bph = (it / kTcBStages) & 1;
          const uint32_t asl = it % kTcASlots, aslph = (it / kTcASlots) & 1;
          const uint32_t ab = it % kTcAcc, aph = (it / kTcAcc) & 1;
          if (ABF ? (my_parity != 0u) : ((it & 1u) != my_parity)) continue;   // ABF: warp 1 issues every chunk (see below)
          if (lane == 0) TC_TRACE(2, it);
          mbar_wait(&bar_acc_empty[ab], aph ^ 1);   // epilogue drained this accumulator (three chunks ago)
          if (lane == 0) TC_TRACE(3, it);
          mbar_wait(&bar_conv[asl], aslph);         // a_hi | a_lo of this chunk are in the TMEM slot
          if (lane == 0) TC_TRACE(28, it);
          const uint32_t a_hi = tmem_u + kTcACol0 + asl * 64;
          const uint32_t a_lo = a_hi + 32;
          const uint32_t d = tmem_u + ab * ACC_COLS;
          if constexpr (ABF) {
            // One issuer only: a raw chunk consumes BOTH weight stages, and a parity wait is only safe when the waiter
            // is at most one phase away from the barrier -- two warps alternating chunks would each wait on stages
            // whose previous phase they never observed (seen on B200: 3 of 9 tiles wrong at Fin = Fout = 128).
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t itb = 2u * it + h;
              const uint32_t sbh = itb % kTcBStages, bphh = (itb / kTcBStages) & 1;
              mbar_wait(&bar_bfull[sbh], bphh);
              tc_fence_after();
              const uint64_t db_hi = make_smem_desc_sw128(smem_b_u + sbh * 2u * b_bytes, 16, 1024);
              const uint64_t db_lo = db_hi + (b_bytes >> 4);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, a_hi + 32 * h + 8 * ks, db_lo + 2 * ks, idesc, (h | ks) ? 1u : 0u);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, a_hi + 32 * h + 8 * ks, db_hi + 2 * ks, idesc, 1u);
                mma_commit(&bar_bempty[sbh]);
                if (h == 1) {
                  mma_commit(&bar_afree[asl]);
                  mma_commit(&bar_acc_full[ab]);
                }
              }
              __syncwarp();
            }
            if (lane == 0) TC_TRACE(5, it);
            continue;
          }
          mbar_wait(&bar_bfull[sb], bph);           // weight tiles of this chunk have landed
          if (lane == 0) TC_TRACE(4, it);
          tc_fence_after();
          const uint32_t sa = smem_b_u + sb * 2u * b_bytes;
          // B descriptors differ only in the 14-bit start-address field (bytes >> 4); A comes from tensor memory
          const uint64_t d_b_hi = make_smem_desc_sw128(sa, 16, 1024);
          const uint64_t d_b_lo = d_b_hi + (b_bytes >> 4);
          if (elect_one()) {
            // small products first: the accumulator rounds toward zero
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, a_lo + 8 * ks, d_b_hi + 2 * ks, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, a_hi + 8 * ks, d_b_lo + 2 * ks, idesc, 1u);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) mma_tf32_ts(d, a_hi + 8 * ks, d_b_hi + 2 * ks, idesc, 1u);
            mma_commit(&bar_bempty[sb]);     // weight stage reusable once these MMAs have read it
            mma_commit(&bar_afree[asl]);     // ... and the TMEM A slot
            mma_commit(&bar_acc_full[ab]);   // chunk accumulator complete
          }
          __syncwarp();
          if (lane == 0) TC_TRACE(5, it);
        }
      }
    }
   }  // warp 2 (TMEM allocator) idles until teardown
  } else if (warp < 8) {
    reg_dec<72>();
    // --------------------------------------------------------------- converter --
    const int r = tid - 128;  // tile row 0..127
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < ngroups * nchunks; ++c, ++it) {
        const uint32_t s = it % kTcStages, ph = (it / kTcStages) & 1;
        mbar_wait(&bar_full[s], ph);
        if (tid == 128) TC_TRACE(6, it);
        const uint32_t a_raw = smem_u32(smem + (size_t)s * a_bytes) + (uint32_t)r * 128;
        // this thread's row of the chunk: split every value into the part the tensor core keeps (hi) and the
        // rounded remainder (lo) and park both in a tensor-memory slot (lane = row, column = k).
        const uint32_t asl = it % kTcASlots, aslph = (it / kTcASlots) & 1;
        const uint32_t t_hi = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTcACol0 + asl * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = lds128(a_raw + (uint32_t)(((4 * h + j) ^ (r & 7)) << 4));
          if (h == 1) {                            // the whole row is in registers: hand the raw stage back to TMA
            // (the loads must have RETURNED before the stage is released: make the release depend on their values)
            float guard = v[0].x + v[1].x + v[2].x + v[3].x;
            asm volatile("" ::"f"(guard) : "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float f[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if constexpr (ABF) {
                // a 32-bit word = two bf16 values (low half = even K index); hi[] takes K values 16h .. 16h+15 of
                // this half's 32, lo[] the next 16 -- both exact, no remainder term
                const uint32_t w = __float_as_uint(f[e]);
                uint32_t* const dst = (j < 2) ? hi : lo;
                dst[8 * (j & 1) + 2 * e] = w << 16;
                dst[8 * (j & 1) + 2 * e + 1] = w & 0xFFFF0000u;
              } else {
                hi[4 * j + e] = __float_as_uint(f[e]);    // unmasked: the tensor core truncates to tf32 itself
                lo[4 * j + e] = __float_as_uint(tf32_lo(f[e]));
              }
            }
          }
          if (h == 0) {
            // the first half is converted while the MMAs that still read this slot (two chunks ago) finish; only
            // the tensor-memory stores wait for them
            mbar_wait(&bar_afree[asl], aslph ^ 1);
            tc_fence_after();
          }
          if constexpr (ABF) {           // K values 32h .. 32h+31 of the chunk's 64
            tmem_st_32x16(t_hi + 32 * h, hi);
            tmem_st_32x16(t_hi + 32 * h + 16, lo);
          } else {
            tmem_st_32x16(t_hi + 16 * h, hi);
            tmem_st_32x16(t_hi + 32 + 16 * h, lo);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_conv[asl]);
        if (tid == 128) TC_TRACE(7, it);
      }
    }
  } else if (warp < 24) {
    reg_inc<80>();
    // ------------------------------------------------------------------- drain --
    DrainArgs da{N, Fout, ntiles, nchunks, ngroups, eps, tmem_base, smem_u32(smem_z), rstd, rowptr, s_bias,
                 &s_sum[0][0][0], &s_var[0][0][0], bar_acc_full, bar_acc_empty, bar_z_full, bar_z_empty, trace,
                 pb.rev ? ntiles - 1 : 0, pb.rev ? -1 : 1};
    if (Fout == 32 * NT) drain_role<NT, true, MODE>(da); else drain_role<NT, false, MODE>(da);
  } else {
    reg_dec<56>();
    // ---------------------------------------------------------------- finisher --
    FinArgs fa{N, Fout, ntiles, ngroups, slope, out, xhat, smem_u32(smem_z), gamma, beta, &tm_s0, &tm_s1,
               bar_z_full, bar_z_empty, trace, pb.rev ? ntiles - 1 : 0, pb.rev ? -1 : 1};
    finisher_role<NT, MODE, ABF>(fa);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================== weight gradient ==
// dW_l = dz^T agg, dW_r = dz^T x as one GEMM  D[m][n] = sum_r dz[r][m] * [agg | x][r][n]   (M = Fout <= 128 lanes,
// N = 2*32*NB columns, K = node rows).  Both operands have the reduction index as their slow dimension, i.e. they are
// "MN-major": the tiles are TMA-loaded as [32 rows][32 floats] boxes with the 128B_ATOM_32B swizzle and described with
// layout SWIZZLE_128B_BASE32B (lbo = bytes of one [rows x 32] block between 32-column blocks, sbo = 512, 1024 bytes per
// K=8 step) -- found with
// tools/tc_probe.cu.  Split precision as in the forward: a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, small terms first.
// Every CTA owns a contiguous slab of rows; the TMEM accumulator is flushed into fp32 registers every kWgFlush chunks
// (the tensor core accumulates with round-toward-zero) and the per-CTA partial is written once at the end;
// k_reduce_parts sums the partials in CTA order (deterministic).
constexpr int kWgRows = 16;                 // node rows per pipeline chunk (the K extent of one stage)
constexpr int kWgStages = 4;                // 4 x 48 KB: the fixed TMA latency is spread over more, smaller stages
constexpr int kWgFlush = 128 / kWgRows;     // chunks per TMEM accumulator (128 rows)
constexpr uint32_t kWgBlk = kWgRows * 128;  // bytes of one [kWgRows x 32 floats] operand block

// BBF: agg and x are bf16 rows.  TMA moves them as 32-bit words ([kWgRows x 32 words] blocks, NB/2 per source); the
// converter splits every word IN PLACE into its odd-column value (w & 0xFFFF0000) and writes the even-column value
// (w << 16) NB blocks further on -- both exact in tf32, so the operand needs no lo part and a k-step costs two MMAs.
// The accumulator columns then come out as [odd(agg) | odd(x) | even(agg) | even(x)]; the epilogue undoes that
// permutation when it writes the partial.  (No knowledge of the swizzle pattern is needed: a word stays where TMA
// put it.)  Needs Fin % 64 == 0.
template <int NB, bool BBF>   // NB = ceil(Fin / 32) in 1..4
__global__ void __launch_bounds__(kWgThreads, 1)
k_wgrad_tc(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_agg,
           const __grid_constant__ CUtensorMap tm_x, int64_t N, int Fin, int Fout, int chunks_per_cta,
           float* __restrict__ part) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full[kWgStages], bar_conv[kWgStages], bar_empty[kWgStages];
  __shared__ uint64_t bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t A_BYTES = 4 * kWgBlk;            // dz chunk: kWgRows rows x 128 columns
  constexpr uint32_t B_BYTES = 2 * NB * kWgBlk;       // [agg | x] chunk: kWgRows rows x 2*32*NB columns
  constexpr uint32_t RAW_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t STAGE_BYTES = 2 * RAW_BYTES;     // raw (= hi) followed by lo, same layout
  constexpr int NCOLS = 64 * NB;                      // accumulator columns
  constexpr uint32_t TMEM_COLS = (2 * NCOLS <= 128) ? 128 : ((2 * NCOLS <= 256) ? 256 : 512);

  const int64_t total_chunks = (N + kWgRows - 1) / kWgRows;
  const int64_t c_beg = (int64_t)blockIdx.x * chunks_per_cta;
  int64_t c_end = c_beg + chunks_per_cta;
  if (c_end > total_chunks) c_end = total_chunks;
  const int nch = c_end > c_beg ? (int)(c_end - c_beg) : 0;

  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_conv[s], 4); mbar_init(&bar_empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bar_acc_full[a], 1); mbar_init(&bar_acc_empty[a], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&tm_dz); tma_prefetch_desc(&tm_agg); tma_prefetch_desc(&tm_x);
  }
  if (warp == 2) tmem_alloc(&tmem_base_s, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    reg_dec<72>();
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer --
      // (whole warp, uniform control flow, one elected lane issues: operands stay on the uniform datapath)
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      for (int it = 0; it < nch; ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
        mbar_wait(&bar_empty[s], ph ^ 1);
        const uint32_t st = smem_u + s * STAGE_BYTES;
        const int r0 = (int)((c_beg + it) * kWgRows);
        if (elect_one()) {
          constexpr int NBS = BBF ? NB / 2 : NB;         // TMA blocks per source (bf16 rows are half as many words)
          mbar_expect_tx(&bar_full[s], A_BYTES + 2 * NBS * kWgBlk);
#pragma unroll
          for (int b = 0; b < 4; ++b) tma_load_2d_u32(st + b * kWgBlk, &tm_dz, b * 32, r0, &bar_full[s]);
#pragma unroll
          for (int b = 0; b < NBS; ++b) tma_load_2d_u32(st + A_BYTES + b * kWgBlk, &tm_agg, b * 32, r0, &bar_full[s]);
#pragma unroll
          for (int b = 0; b < NBS; ++b) tma_load_2d_u32(st + A_BYTES + (NBS + b) * kWgBlk, &tm_x, b * 32, r0, &bar_full[s]);
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer --
      // whole warp in uniform control flow, one elected lane issues (see k_sage_tc: under `if (lane == 0)` ptxas wraps
      // every tcgen05.mma in a 15-instruction broadcast loop and the issue rate, not the tensor pipe, sets the pace)
      const uint32_t idesc = make_idesc_tf32(128, NCOLS, 1, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      for (int it = 0; it < nch; ++it) {
        const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
        const uint32_t fl = it / kWgFlush, ab = fl & 1, aph = (fl >> 1) & 1;
        const bool first = (it % kWgFlush) == 0;
        if (first) mbar_wait(&bar_acc_empty[ab], aph ^ 1);
        mbar_wait(&bar_conv[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u + s * STAGE_BYTES;
        const uint64_t d_a_hi = make_smem_desc(sa, kWgBlk, 512, 1);
        const uint64_t d_b_hi = d_a_hi + (A_BYTES >> 4);
        const uint64_t d_a_lo = d_a_hi + (RAW_BYTES >> 4);
        const uint64_t d_b_lo = d_b_hi + (RAW_BYTES >> 4);
        const uint32_t d = tmem_u + ab * NCOLS;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < kWgRows / 8; ++ks) mma_tf32_ss(d, d_a_lo + 64 * ks, d_b_hi + 64 * ks, idesc, (first && ks == 0) ? 0u : 1u);
          if constexpr (!BBF) {
#pragma unroll
            for (int ks = 0; ks < kWgRows / 8; ++ks) mma_tf32_ss(d, d_a_hi + 64 * ks, d_b_lo + 64 * ks, idesc, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < kWgRows / 8; ++ks) mma_tf32_ss(d, d_a_hi + 64 * ks, d_b_hi + 64 * ks, idesc, 1u);
          mma_commit(&bar_empty[s]);
          if ((it % kWgFlush) == kWgFlush - 1 || it == nch - 1) mma_commit(&bar_acc_full[ab]);
        }
        __syncwarp();
      }
    }
  } else if (warp < 8) {
    reg_dec<72>();
    // ------------------------------------------------------------------- converter --
    const int t = tid - 128;
    for (int it = 0; it < nch; ++it) {
      const uint32_t s = it % kWgStages, ph = (it / kWgStages) & 1;
      mbar_wait(&bar_full[s], ph);
      const uint32_t raw = smem_u32(smem + (size_t)s * STAGE_BYTES);
      if constexpr (BBF) {
        constexpr int NVA = A_BYTES / 16 / 128;          // dz: lo part as usual
#pragma unroll
        for (int i = 0; i < NVA; ++i) {
          const uint32_t off = (uint32_t)(t + i * 128) * 16;
          const float4 v = lds128(raw + off);
          float4 l;
          l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
          sts128(raw + RAW_BYTES + off, l);
        }
        constexpr int NVB = NB * kWgBlk / 16 / 128;      // [agg | x] words: odd values in place, even values NB blocks on
#pragma unroll
        for (int i = 0; i < NVB; ++i) {
          const uint32_t off = A_BYTES + (uint32_t)(t + i * 128) * 16;
          const float4 v = lds128(raw + off);
          const uint32_t w[4] = {__float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w)};
          sts128(raw + off, make_float4(__uint_as_float(w[0] & 0xFFFF0000u), __uint_as_float(w[1] & 0xFFFF0000u),
                                        __uint_as_float(w[2] & 0xFFFF0000u), __uint_as_float(w[3] & 0xFFFF0000u)));
          sts128(raw + off + NB * kWgBlk, make_float4(__uint_as_float(w[0] << 16), __uint_as_float(w[1] << 16),
                                                      __uint_as_float(w[2] << 16), __uint_as_float(w[3] << 16)));
        }
      } else {
      constexpr int NV = RAW_BYTES / 16 / 128;   // float4 per thread
#pragma unroll 4
      for (int i = 0; i < NV; ++i) {
        const uint32_t off = (uint32_t)(t + i * 128) * 16;
        const float4 v = lds128(raw + off);
        float4 l;
        l.x = tf32_lo(v.x); l.y = tf32_lo(v.y); l.z = tf32_lo(v.z); l.w = tf32_lo(v.w);
        sts128(raw + RAW_BYTES + off, l);
      }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_conv[s]);
    }
  } else {
    reg_inc<184>();
    // -------------------------------------------------------------------- epilogue --
    constexpr int HC = 32 * NB;              // columns per thread: half of the accumulator
    const int q = warp & 3;
    const int hf = (warp - 8) >> 2;          // 0: agg half (dW_l), 1: x half (dW_r)
    const int m = q * 32 + lane;             // output feature (row of dW)
    const uint32_t tq = tmem_base + ((uint32_t)(q * 32) << 16) + hf * HC;
    float acc[HC];
#pragma unroll
    for (int j = 0; j < HC; ++j) acc[j] = 0.f;
    const int nfl = (nch + kWgFlush - 1) / kWgFlush;
    for (int fl = 0; fl < nfl; ++fl) {
      const uint32_t ab = fl & 1, aph = (fl >> 1) & 1;
      mbar_wait(&bar_acc_full[ab], aph);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < HC / 16; ++g) {
        uint32_t rr[16];
        tmem_ld_32x16(tq + ab * NCOLS + g * 16, rr);
        tmem_ld_wait();
        if (g == HC / 16 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[ab]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[g * 16 + j] += __uint_as_float(rr[j]);
      }
    }
    // partial of this CTA: part[cta][hf][m][n], n < Fin
    if (m < Fout) {
      if constexpr (BBF) {
        // accumulator half hf = odd (0) / even (1) feature columns; inside: [agg blocks (NB/2) | x blocks (NB/2)]
#pragma unroll
        for (int j = 0; j < HC; ++j) {
          const int blk = j >> 5, src = blk / (NB / 2), word = (blk % (NB / 2)) * 32 + (j & 31);
          const int col = 2 * word + (hf == 0 ? 1 : 0);
          if (col < Fin) part[(((int64_t)blockIdx.x * 2 + src) * Fout + m) * Fin + col] = acc[j];
        }
      } else {
        float* dst = part + (((int64_t)blockIdx.x * 2 + hf) * Fout + m) * Fin;
#pragma unroll
        for (int j = 0; j < HC; ++j)
          if (j < Fin) dst[j] = acc[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------ host --
static bool tc_disabled() {
  static int disabled = -1;
  if (disabled < 0) { const char* e = getenv("SLDM_DISABLE_TC"); disabled = (e && e[0] == '1') ? 1 : 0; }
  return disabled != 0;
}
static bool p16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

bool project_forward_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* agg, const float* x,
                                 const float* out, const float* xhat) {
  return !tc_disabled() && N >= 1 && N < ((int64_t)1 << 31) - 256 && Fin % 32 == 0 && Fin >= 32 && Fout % 16 == 0 &&
         Fout >= 16 && Fout <= 128 && p16(agg) && p16(x) && p16(out) && p16(xhat);
}
bool dgrad_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const float* dagg, const float* dxroot) {
  return !tc_disabled() && N >= 1 && N < ((int64_t)1 << 31) - 256 && Fout % 32 == 0 && Fout >= 32 && Fin % 16 == 0 &&
         Fin >= 16 && Fin <= 128 && p16(dz) && p16(dagg) && p16(dxroot);
}

int64_t project_forward_tc_ws_bytes(int32_t Fin, int32_t Fout) { return align_bytes((int64_t)4 * Fin * Fout * 4); }
int64_t dgrad_tc_ws_bytes(int32_t Fin, int32_t Fout) { return align_bytes((int64_t)4 * Fin * Fout * 4); }

// transposed split for DGRAD: out block b in {W_l^T hi, W_l^T lo, W_r^T hi, W_r^T lo}, each [Fin][Fout]
__global__ void __launch_bounds__(256)
k_split_weights_t(const float* __restrict__ W_l, const float* __restrict__ W_r, int Fin, int Fout,
                  float* __restrict__ out) {
  const int count = Fin * Fout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * count; i += gridDim.x * blockDim.x) {
    const int which = i >= count;
    const int j = which ? i - count : i;      // index into the transposed [Fin][Fout] matrix
    const int n = j / Fout, k = j % Fout;     // n = input feature, k = output feature
    const float w = (which ? W_r : W_l)[(int64_t)k * Fin + n];
    const float hf = __uint_as_float((__float_as_uint(w) + 0x1000u) & 0xFFFFE000u);
    const float lf = __uint_as_float((__float_as_uint(w - hf) + 0x1000u) & 0xFFFFE000u);
    out[(int64_t)(2 * which) * count + j] = hf;
    out[(int64_t)(2 * which + 1) * count + j] = lf;
  }
}

// SLDM_TC_REVERSE: bit 0 = the forward projection, bit 1 = DGRAD walk their row blocks from the last to the first, so
// that the rows the PREVIOUS kernel of the chain wrote last (the tail of agg / dz, still in the 126 MB L2) are read
// first.  Measured on the batch step (three alternating runs each): forward reversed 3.518 vs 3.542 ms (inference
// 1.283 vs 1.294), DGRAD reversed 3.574 (worse: k_ln_bwd_rows does not finish with the last rows), both 3.532.
// Default: the forward only.
static int tc_reverse_mask() {
  static const int m = [] { const char* e = getenv("SLDM_TC_REVERSE"); return e ? atoi(e) : 1; }();
  return m;
}

template <int NT, int MODE, bool ABF = false>
static int launch_tc(const CUtensorMap& ma, const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& ms0,
                     const CUtensorMap& ms1, const TcProblem& pb,
                     const float* b_l, const float* g, const float* b, float eps, float slope,
                     float* out, float* xhat, float* rstd, const int32_t* rowptr, cudaStream_t s) {
  const size_t smem = (size_t)kTcStages * kTcBM * 128 + (size_t)kTcBStages * 2 * pb.Nout * 128 + kTcZBytes + 1024;
  SLDM_OPT_IN_SMEM((k_sage_tc<NT, MODE, ABF>), kTcStages * kTcBM * 128 + kTcBStages * 2 * 128 * 128 + kTcZBytes + 1024);
  const int64_t ntiles = ceil_div<int64_t>(pb.N, kTcBM);
  const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
  long long* trace = nullptr;
  const char* tf = getenv("SLDM_TC_TRACE");
  if (tf && tf[0]) {
    SLDM_CUDA(cudaMalloc(&trace, sizeof(long long) * kTraceIts * kTraceEv));
    SLDM_CUDA(cudaMemsetAsync(trace, 0, sizeof(long long) * kTraceIts * kTraceEv, s));
  }
  k_sage_tc<NT, MODE, ABF><<<grid, kTcThreads, smem, s>>>(ma, mx, mw, ms0, ms1, pb, b_l, g, b, eps, slope, out, xhat, rstd, rowptr, trace);
  SLDM_LAUNCH_CHECK("k_sage_tc");
  if (trace) {
    static long long h[kTraceIts * kTraceEv];
    SLDM_CUDA(cudaStreamSynchronize(s));
    SLDM_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(trace);
    FILE* f = fopen(tf, "w");
    if (f) {
      fprintf(f, "# cycles from t0; ev: 0 tma_wait_empty 1 tma_got_empty 2 mma_top 3 mma_accempty 4 mma_conv 5 mma_committed "
                 "6 conv_full 7 conv_done 8 drain_wait 9 drain_accfull 10 drain_released 11 drain_tile_done 12 z_empty 13 z_parked "
                 "16 fin_wait 17 fin_got 18 fin_done (16-18 indexed by hand-off number) 28 mma_conv_only 29 tmaB_wait_empty 30 tmaB_got_empty\n");
      const long long t0 = h[0];
      for (int i = 0; i < kTraceIts; ++i) {
        fprintf(f, "%d", i);
        for (int e = 0; e < kTraceEv; ++e) fprintf(f, " %lld", h[i * kTraceEv + e] ? h[i * kTraceEv + e] - t0 : -1);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  return SLDM_OK;
}

template <int MODE, bool ABF = false>
static int dispatch_tc(const CUtensorMap& ma, const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& ms0,
                       const CUtensorMap& ms1, const TcProblem& pb,
                       const float* b_l, const float* g, const float* b, float eps, float slope,
                       float* out, float* xhat, float* rstd, const int32_t* rowptr, cudaStream_t s) {
  switch (ceil_div(pb.Nout, 32)) {
    case 1: return launch_tc<1, MODE, ABF>(ma, mx, mw, ms0, ms1, pb, b_l, g, b, eps, slope, out, xhat, rstd, rowptr, s);
    case 2: return launch_tc<2, MODE, ABF>(ma, mx, mw, ms0, ms1, pb, b_l, g, b, eps, slope, out, xhat, rstd, rowptr, s);
    case 3: return launch_tc<3, MODE, ABF>(ma, mx, mw, ms0, ms1, pb, b_l, g, b, eps, slope, out, xhat, rstd, rowptr, s);
    default: return launch_tc<4, MODE, ABF>(ma, mx, mw, ms0, ms1, pb, b_l, g, b, eps, slope, out, xhat, rstd, rowptr, s);
  }
}

int project_forward_tc_launch(const float* agg, const float* x, int64_t N, int32_t Fin, int32_t Fout,
                              const float* W_l, const float* b_l, const float* W_r,
                              const float* ln_w, const float* ln_b, float eps, float slope,
                              float* out, float* xhat, float* rstd, void* ws, int64_t ws_bytes, cudaStream_t s) {
  SLDM_REQUIRE(ws != nullptr && ws_bytes >= project_forward_tc_ws_bytes(Fin, Fout), SLDM_EWORKSPACE,
               "projection (tensor path): workspace too small");
  float* wsplit = static_cast<float*>(ws);
  const int count = Fin * Fout;
  k_split_weights<<<ceil_div(2 * count, 256 * 4), 256, 0, s>>>(W_l, W_r, count, wsplit);
  SLDM_LAUNCH_CHECK("k_split_weights");
  CUtensorMap ma, mx, mw;
  int rc;
  if ((rc = make_tmap_2d_f32(&ma, agg, (uint64_t)N, Fin, Fin, kTcBM, 32))) return rc;
  if ((rc = make_tmap_2d_f32(&mx, x, (uint64_t)N, Fin, Fin, kTcBM, 32))) return rc;
  if ((rc = make_tmap_2d_f32(&mw, wsplit, (uint64_t)4 * Fout, Fin, Fin, Fout, 32))) return rc;
  CUtensorMap ms = ma;   // bulk-store map of xhat (training, whole 32-column slabs): [32 rows][32 fp32] boxes out of the parked tile
  if (xhat != nullptr && Fout % 32 == 0 && (rc = make_tmap_2d_f32(&ms, xhat, (uint64_t)N, Fout, Fout, 32, 32))) return rc;
  TcProblem pb{N, Fin / 32, 2, 1, Fout};
  pb.rev = tc_reverse_mask() & 1;
  return dispatch_tc<MODE_FWD>(ma, mx, mw, ms, ms, pb, b_l, ln_w, ln_b, eps, slope, out, xhat, rstd, nullptr, s);
}

// bf16 feature storage: agg, x and out are bf16 rows ([N,Fin] / [N,Fout] uint16), weights / LayerNorm / xhat / rstd fp32
bool project_forward_bf16_eligible(int64_t N, int32_t Fin, int32_t Fout, const void* agg, const void* x,
                                   const void* out, const float* xhat) {
  return !tc_disabled() && N >= 1 && N < ((int64_t)1 << 31) - 256 && Fin % 64 == 0 && Fin >= 64 && Fin <= 256 &&
         Fout % 16 == 0 && Fout >= 16 && Fout <= 128 && p16(agg) && p16(x) && p16(out) && p16(xhat);
}
int project_forward_bf16_launch(const void* agg, const void* x, int64_t N, int32_t Fin, int32_t Fout,
                                const float* W_l, const float* b_l, const float* W_r,
                                const float* ln_w, const float* ln_b, float eps, float slope,
                                void* out, float* xhat, float* rstd, void* ws, int64_t ws_bytes, cudaStream_t s) {
  SLDM_REQUIRE(ws != nullptr && ws_bytes >= project_forward_tc_ws_bytes(Fin, Fout), SLDM_EWORKSPACE,
               "projection (bf16 features): workspace too small");
  float* wsplit = static_cast<float*>(ws);
  const int count = Fin * Fout;
  k_split_weights<<<ceil_div(2 * count, 256 * 4), 256, 0, s>>>(W_l, W_r, count, wsplit);
  SLDM_LAUNCH_CHECK("k_split_weights");
  CUtensorMap ma, mx, mw;
  int rc;
  // the bf16 rows travel as 32-bit words: [N][Fin/2]
  if ((rc = make_tmap_2d_f32(&ma, static_cast<const float*>(agg), (uint64_t)N, Fin / 2, Fin / 2, kTcBM, 32))) return rc;
  if ((rc = make_tmap_2d_f32(&mx, static_cast<const float*>(x), (uint64_t)N, Fin / 2, Fin / 2, kTcBM, 32))) return rc;
  if ((rc = make_tmap_2d_f32(&mw, wsplit, (uint64_t)4 * Fout, Fin, Fin, Fout, 32))) return rc;
  CUtensorMap ms = ma;
  if (xhat != nullptr && Fout % 32 == 0 && (rc = make_tmap_2d_f32(&ms, xhat, (uint64_t)N, Fout, Fout, 32, 32))) return rc;
  TcProblem pb{N, Fin / 64, 2, 1, Fout};
  return dispatch_tc<MODE_FWD, true>(ma, mx, mw, ms, ms, pb, b_l, ln_w, ln_b, eps, slope, static_cast<float*>(out), xhat, rstd,
                                     nullptr, s);
}

// dagg[N,Fin] = (dz W_l) / max(deg,1) ; dxroot[N,Fin] = dz W_r          (dz is [N,Fout])
int dgrad_tc_launch(const float* dz, int64_t N, int32_t Fin, int32_t Fout, const float* W_l, const float* W_r,
                    const int32_t* rowptr_dst, float* dagg, float* dxroot, void* ws, int64_t ws_bytes,
                    cudaStream_t s) {
  SLDM_REQUIRE(ws != nullptr && ws_bytes >= dgrad_tc_ws_bytes(Fin, Fout), SLDM_EWORKSPACE,
               "dgrad (tensor path): workspace too small");
  float* wsplit = static_cast<float*>(ws);
  k_split_weights_t<<<ceil_div(2 * Fin * Fout, 256 * 4), 256, 0, s>>>(W_l, W_r, Fin, Fout, wsplit);
  SLDM_LAUNCH_CHECK("k_split_weights_t");
  CUtensorMap mz, mw;
  int rc;
  if ((rc = make_tmap_2d_f32(&mz, dz, (uint64_t)N, Fout, Fout, kTcBM, 32))) return rc;
  if ((rc = make_tmap_2d_f32(&mw, wsplit, (uint64_t)4 * Fin, Fout, Fout, Fin, 32))) return rc;
  CUtensorMap ms0 = mz, ms1 = mz;   // bulk-store maps of dagg / dxroot (whole 32-column slabs)
  if (Fin % 32 == 0) {
    if ((rc = make_tmap_2d_f32(&ms0, dagg, (uint64_t)N, Fin, Fin, 32, 32))) return rc;
    if ((rc = make_tmap_2d_f32(&ms1, dxroot, (uint64_t)N, Fin, Fin, 32, 32))) return rc;
  }
  TcProblem pb{N, Fout / 32, 1, 2, Fin};
  pb.rev = (tc_reverse_mask() >> 1) & 1;
  return dispatch_tc<MODE_DGRAD>(mz, mz, mw, ms0, ms1, pb, nullptr, nullptr, nullptr, 0.f, 0.f, dagg, dxroot, nullptr, rowptr_dst, s);
}

bool wgrad_tc_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const float* agg, const float* x) {
  return !tc_disabled() && N >= 1 && N < ((int64_t)1 << 31) - 256 && Fin % 4 == 0 && Fout % 4 == 0 && Fin >= 16 &&
         Fin <= 128 && Fout >= 16 && Fout <= 128 && p16(dz) && p16(agg) && p16(x);
}
int wgrad_tc_grid() { return num_sms(); }
int64_t wgrad_tc_ws_bytes(int32_t Fin, int32_t Fout) { return align_bytes((int64_t)wgrad_tc_grid() * 2 * Fin * Fout * 4); }

template <int NB, bool BBF = false>
static int launch_wgrad(const CUtensorMap& mz, const CUtensorMap& ma, const CUtensorMap& mx, int64_t N, int Fin,
                        int Fout, int cpc, int grid, float* part, cudaStream_t s) {
  const size_t smem = (size_t)kWgStages * 2 * (4 + 2 * NB) * kWgBlk + 1024;
  SLDM_OPT_IN_SMEM((k_wgrad_tc<NB, BBF>), smem);
  k_wgrad_tc<NB, BBF><<<grid, kWgThreads, smem, s>>>(mz, ma, mx, N, Fin, Fout, cpc, part);
  SLDM_LAUNCH_CHECK("k_wgrad_tc");
  return SLDM_OK;
}

bool wgrad_bf16_eligible(int64_t N, int32_t Fin, int32_t Fout, const float* dz, const void* agg, const void* x) {
  return !tc_disabled() && N >= 1 && N < ((int64_t)1 << 31) - 256 && Fin % 64 == 0 && Fout % 4 == 0 && Fin >= 64 &&
         Fin <= 128 && Fout >= 16 && Fout <= 128 && p16(dz) && p16(agg) && p16(x);
}

// part[grid][2][Fout][Fin]; returns the number of partials through *nparts.  bf16_ops: agg and x are bf16 rows.
int wgrad_tc_launch(const float* dz, const void* agg_v, const void* x_v, int64_t N, int32_t Fin, int32_t Fout,
                    float* part, int* nparts, cudaStream_t s, bool bf16_ops) {
  CUtensorMap mz, ma, mx;
  int rc;
  const float* agg = static_cast<const float*>(agg_v);
  const float* x = static_cast<const float*>(x_v);
  const int Fw = bf16_ops ? Fin / 2 : Fin;          // 32-bit words per operand row
  if ((rc = make_tmap_2d_f32(&mz, dz, (uint64_t)N, Fout, Fout, kWgRows, 32, 1))) return rc;
  if ((rc = make_tmap_2d_f32(&ma, agg, (uint64_t)N, Fw, Fw, kWgRows, 32, 1))) return rc;
  if ((rc = make_tmap_2d_f32(&mx, x, (uint64_t)N, Fw, Fw, kWgRows, 32, 1))) return rc;
  const int64_t total_chunks = ceil_div<int64_t>(N, kWgRows);
  int grid = wgrad_tc_grid();
  if (grid > total_chunks) grid = (int)total_chunks;
  // whole flush groups per CTA keep the slab boundaries independent of the grid rounding
  int64_t cpc = round_up<int64_t>(ceil_div<int64_t>(total_chunks, grid), kWgFlush);
  grid = (int)ceil_div<int64_t>(total_chunks, cpc);
  *nparts = grid;
  if (bf16_ops)
    return Fin == 64 ? launch_wgrad<2, true>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s)
                     : launch_wgrad<4, true>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s);
  switch (ceil_div(Fin, 32)) {
    case 1: return launch_wgrad<1>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s);
    case 2: return launch_wgrad<2>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s);
    case 3: return launch_wgrad<3>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s);
    default: return launch_wgrad<4>(mz, ma, mx, N, Fin, Fout, (int)cpc, grid, part, s);
  }
}

}  // namespace sldm

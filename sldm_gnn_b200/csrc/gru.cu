// gru.cu -- fused single-layer GRU over short sequences (the sequence head in front of the vehicle-graph SageBlock).
//
// Replaces `gru_out, hlast = self.gru(x); x = hlast[-1]` (src/models/grusage.py:160-161; nn.GRU(input 6, hidden 96,
// one layer, batch_first) built at grusage.py:55-60) and its autograd backward.  torch runs that as 16 x (2 GEMMs +
// a cell kernel) forward and as many again backward, with every gate tensor making a round trip through HBM; here the
// whole recurrence of a tile of 64 sequences stays on one SM:
//
//   forward  (k_gru_fwd):  W_hh [3H,H] lives in shared memory (row stride H+4 floats: conflict-free 128-bit loads with
//            lane = hidden unit), a warp owns 8 sequences for all T steps (its h rows are read as shared-memory
//            broadcasts), a lane owns H/32 hidden units x 3 gates x 8 rows = 72 fp32 accumulators at H = 96; the gate
//            math of a (row, unit) is local to its thread, so the time loop needs no block-level barrier at all.
//            When training it writes h_{t-1}, r, z, n and (W_hn h + b_hn) per step -- what torch's fused cell saves --
//            as one [T, N, 5, H] buffer: a warp's output of a step is 8 x 5H contiguous floats (with five [N, T, H]
//            arrays the 128-byte pieces were 6 KB apart and the stores cost 3.6 ms on top of 5.7 ms of arithmetic).
//   backward (k_gru_bwd):  the same tiling walks t = T-1 .. 0 with W_hh transposed in shared memory
//            (dh_{t-1} = dh_t z + [dr dz dn r] W_hh); it emits the hidden-gate gradients dgh [T,N,3H] and accumulates
//            dW_ih, db_ih, db_hh per CTA in registers (fixed order; the per-CTA partials are summed by the caller).
//   wgrad    (k_gru_wgrad): dW_hh = dgh^T h_prev over the T*N rows, one persistent CTA per SM (below).
//
// Arithmetic: FP32 FMA in sequential k order, expf / tanhf / IEEE division like ATen's gru_cell_forward, i.e. the
// reference's own formulas (aten/src/ATen/native/cuda/RNN.cu); bound by the FP32 pipe (2*3H*H flops per sequence and
// step each way), not by HBM.  Limits: H in {32, 64, 96}, input width <= 8, the tile's x must fit shared memory.
#include "common.cuh"
#include <algorithm>

namespace sldm {
namespace {

constexpr int kGruRows = 64;      // sequences per CTA
constexpr int kGruRpw = 8;        // sequences per warp
constexpr int kGruThreads = 256;
constexpr int kGruMaxI = 8;
constexpr int kGruLdi = 9;        // W_ih row stride in shared memory (odd: conflict-free scalar loads)
constexpr int64_t kGruSmemMax = 227 * 1024;

__device__ __forceinline__ float gru_sigmoid(float v) { return 1.0f / (1.0f + expf(-v)); }

__device__ __forceinline__ void fma4(float& acc, const float4& a, const float4& b) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
}

// Packed FP32 FMA of sm_100 (fma.rn.f32x2 -> FFMA2): two FMAs per issue slot at the same pipe rate (tools/ffma2_probe.cu:
// a register-tiled 8 x 9 outer product reaches 56 TFLOP/s with FFMA and 65 with FFMA2 on B200).  Used by k_gru_wgrad
// (4.03 -> 3.75 ms).  In the recurrence kernels a k-paired FFMA2 main loop (even k in the low half of an accumulator pair,
// odd k in the high half) measured slower than scalar FFMA (forward 5.70 vs 5.34 ms: 144 accumulator registers and ~80
// pair-shuffling MOVs per 288 FFMA2), so they keep the scalar loop.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f32x2 f32x2_dup(float v) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void load_x_tile(float* xs, const float* __restrict__ x, int64_t row0, int nrows, int TI) {
  const float* xg = x + row0 * TI;
  for (int i = threadIdx.x; i < kGruRows * TI; i += kGruThreads) xs[i] = i < nrows * TI ? __ldg(xg + i) : 0.f;
}

template <int H, bool SAVE>
__global__ void __launch_bounds__(kGruThreads, 1)
k_gru_fwd(const float* __restrict__ x, int64_t N, int T, int I,
          const float* __restrict__ W_ih, const float* __restrict__ W_hh,
          const float* __restrict__ b_ih, const float* __restrict__ b_hh,
          float* __restrict__ h_last, float* __restrict__ saved) {
  constexpr int U = H / 32, LDW = H + 4;
  extern __shared__ __align__(16) float sm[];
  float* Ws = sm;                        // [3H][LDW]
  float* hs = Ws + 3 * H * LDW;          // [64][H]   current hidden state of the tile
  float* Wi = hs + kGruRows * H;         // [3H][9]
  float* bs = Wi + 3 * H * kGruLdi;      // b_ir+b_hr | b_iz+b_hz | b_in | b_hn
  float* xs = bs + 4 * H;                // [64][T*I]
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kGruRows;
  const int nrows = (int)min((int64_t)kGruRows, N - row0);
  const int TI = T * I;

  for (int i = tid; i < 3 * H * (H / 4); i += kGruThreads) {
    const int r = i / (H / 4), c = i % (H / 4);
    *reinterpret_cast<float4*>(Ws + r * LDW + 4 * c) = __ldg(reinterpret_cast<const float4*>(W_hh) + i);
  }
  for (int i = tid; i < 3 * H * I; i += kGruThreads) Wi[(i / I) * kGruLdi + (i % I)] = __ldg(W_ih + i);
  for (int i = tid; i < H; i += kGruThreads) {
    bs[i] = __ldg(b_ih + i) + __ldg(b_hh + i);
    bs[H + i] = __ldg(b_ih + H + i) + __ldg(b_hh + H + i);
    bs[2 * H + i] = __ldg(b_ih + 2 * H + i);
    bs[3 * H + i] = __ldg(b_hh + 2 * H + i);
  }
  for (int i = tid; i < kGruRows * H; i += kGruThreads) hs[i] = 0.f;
  load_x_tile(xs, x, row0, nrows, TI);
  __syncthreads();

  const int r0 = w * kGruRpw;
  const float* hrow = hs + r0 * H;
  const float* wrow = Ws + l * LDW;
  float hreg[kGruRpw][U];
#pragma unroll
  for (int i = 0; i < kGruRpw; ++i)
#pragma unroll
    for (int u = 0; u < U; ++u) hreg[i][u] = 0.f;

  for (int t = 0; t < T; ++t) {
    float ar[kGruRpw][U], az[kGruRpw][U], an[kGruRpw][U];
#pragma unroll
    for (int i = 0; i < kGruRpw; ++i)
#pragma unroll
      for (int u = 0; u < U; ++u) ar[i][u] = az[i][u] = an[i][u] = 0.f;

    // hidden part: [8 rows] x [3 gates x U units] over k, h rows as broadcasts, weights one row per lane
#pragma unroll 2
    for (int k = 0; k < H; k += 4) {
      float4 hv[kGruRpw];
#pragma unroll
      for (int i = 0; i < kGruRpw; ++i) hv[i] = *reinterpret_cast<const float4*>(hrow + i * H + k);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 w0 = *reinterpret_cast<const float4*>(wrow + (u * 32) * LDW + k);
        const float4 w1 = *reinterpret_cast<const float4*>(wrow + (H + u * 32) * LDW + k);
        const float4 w2 = *reinterpret_cast<const float4*>(wrow + (2 * H + u * 32) * LDW + k);
#pragma unroll
        for (int i = 0; i < kGruRpw; ++i) {
          fma4(ar[i][u], hv[i], w0);
          fma4(az[i][u], hv[i], w1);
          fma4(an[i][u], hv[i], w2);
        }
      }
    }
    if (SAVE) {
#pragma unroll
      for (int i = 0; i < kGruRpw; ++i)
        if (r0 + i < nrows) {
          float* o = saved + (((int64_t)t * N + row0 + r0 + i) * 5) * H + l;
#pragma unroll
          for (int u = 0; u < U; ++u) o[32 * u] = hreg[i][u];
        }
    }
    // input part (I <= 8): r and z continue their accumulators, the n gate keeps its input half apart
    float ai[kGruRpw][U];
#pragma unroll
    for (int i = 0; i < kGruRpw; ++i)
#pragma unroll
      for (int u = 0; u < U; ++u) ai[i][u] = 0.f;
    for (int i2 = 0; i2 < I; ++i2) {
      float xv[kGruRpw];
#pragma unroll
      for (int i = 0; i < kGruRpw; ++i) xv[i] = xs[(r0 + i) * TI + t * I + i2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float wr = Wi[(l + 32 * u) * kGruLdi + i2];
        const float wz = Wi[(H + l + 32 * u) * kGruLdi + i2];
        const float wn = Wi[(2 * H + l + 32 * u) * kGruLdi + i2];
#pragma unroll
        for (int i = 0; i < kGruRpw; ++i) {
          ar[i][u] = fmaf(xv[i], wr, ar[i][u]);
          az[i][u] = fmaf(xv[i], wz, az[i][u]);
          ai[i][u] = fmaf(xv[i], wn, ai[i][u]);
        }
      }
    }
    // gates (ATen gru_cell_forward): r, z = sigmoid, n = tanh(i_n + b_in + r (h_n + b_hn)), h' = n + z (h - n)
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int j = l + 32 * u;
      const float br = bs[j], bz = bs[H + j], bin = bs[2 * H + j], bhn = bs[3 * H + j];
#pragma unroll
      for (int i = 0; i < kGruRpw; ++i) {
        const float rg = gru_sigmoid(ar[i][u] + br);
        const float zg = gru_sigmoid(az[i][u] + bz);
        const float hn = an[i][u] + bhn;
        const float ng = tanhf(ai[i][u] + bin + rg * hn);
        if (SAVE && r0 + i < nrows) {
          float* o = saved + (((int64_t)t * N + row0 + r0 + i) * 5) * H + j;
          o[H] = rg; o[2 * H] = zg; o[3 * H] = ng; o[4 * H] = hn;
        }
        hreg[i][u] = ng + zg * (hreg[i][u] - ng);
      }
    }
    __syncwarp();   // every lane has finished reading the warp's h rows
#pragma unroll
    for (int i = 0; i < kGruRpw; ++i)
#pragma unroll
      for (int u = 0; u < U; ++u) hs[(r0 + i) * H + l + 32 * u] = hreg[i][u];
    __syncwarp();
  }
#pragma unroll
  for (int i = 0; i < kGruRpw; ++i)
    if (r0 + i < nrows) {
#pragma unroll
      for (int u = 0; u < U; ++u) h_last[(row0 + r0 + i) * H + l + 32 * u] = hreg[i][u];
    }
}

// per-CTA partial layout (floats): [28*U][32 lanes]: v = (u*3+g)*8 + i2 -> dW_ih[g*H + 32u + lane][i2];
// v = 24U + u*3 + g -> db_ih[g*H + 32u + lane]; v = 27U + u -> db_hh[2H + 32u + lane] (its r and z thirds equal db_ih's)
template <int H>
__global__ void __launch_bounds__(kGruThreads, 1)
k_gru_bwd(const float* __restrict__ x, int64_t N, int T, int I, const float* __restrict__ W_hh,
          const float* __restrict__ dh_last, const float* __restrict__ saved,
          float* __restrict__ dgh, float* __restrict__ gin_out, float* __restrict__ parts) {
  constexpr int U = H / 32, G3 = 3 * H, LDT = 3 * H + 4, V = 28 * U;
  extern __shared__ __align__(16) float sm[];
  float* Wt = sm;                       // [H][LDT]: Wt[j][k] = W_hh[k][j]
  float* gs = Wt + H * LDT;             // [64][3H]  hidden-gate gradients of the current step
  float* xs = gs + kGruRows * G3;       // [64][T*I]
  const int tid = threadIdx.x, w = tid >> 5, l = tid & 31;
  const int64_t row0 = (int64_t)blockIdx.x * kGruRows;
  const int nrows = (int)min((int64_t)kGruRows, N - row0);
  const int TI = T * I;

  for (int i = tid; i < G3 * H; i += kGruThreads) Wt[(i % H) * LDT + (i / H)] = __ldg(W_hh + i);
  load_x_tile(xs, x, row0, nrows, TI);
  __syncthreads();

  const int r0 = w * kGruRpw;
  float dh[kGruRpw][U];
#pragma unroll
  for (int i = 0; i < kGruRpw; ++i)
#pragma unroll
    for (int u = 0; u < U; ++u) dh[i][u] = r0 + i < nrows ? __ldg(dh_last + (row0 + r0 + i) * H + l + 32 * u) : 0.f;
  float dWi[U][3][kGruMaxI], dbi[U][3], dbn[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    dbn[u] = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      dbi[u][g] = 0.f;
#pragma unroll
      for (int i2 = 0; i2 < kGruMaxI; ++i2) dWi[u][g][i2] = 0.f;
    }
  }
  const float* grow = gs + r0 * G3;
  const float* wrow = Wt + l * LDT;

  for (int t = T - 1; t >= 0; --t) {
    // cell backward (ATen gru_cell_backward) for the thread's (row, unit) pairs
#pragma unroll
    for (int i = 0; i < kGruRpw; ++i) {
      const bool valid = r0 + i < nrows;
      float xv[kGruMaxI];
#pragma unroll
      for (int i2 = 0; i2 < kGruMaxI; ++i2) xv[i2] = i2 < I ? xs[(r0 + i) * TI + t * I + i2] : 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = l + 32 * u;
        const int64_t tr = (int64_t)t * N + row0 + r0 + i;     // row of the [T, N, ...] buffers
        float rg = 0.f, zg = 0.f, ng = 0.f, hn = 0.f, hpv = 0.f;
        if (valid) {
          const float* o = saved + tr * 5 * H + j;
          hpv = __ldg(o); rg = __ldg(o + H); zg = __ldg(o + 2 * H); ng = __ldg(o + 3 * H); hn = __ldg(o + 4 * H);
        }
        const float g = dh[i][u];
        const float gig = g * (hpv - ng) * (1.f - zg) * zg;
        const float ghx = g * zg;
        const float gin = g * (1.f - zg) * (1.f - ng * ng);
        const float ghn = gin * rg;
        const float grg = gin * hn * (1.f - rg) * rg;
        if (valid) {
          float* og = dgh + tr * G3 + j;
          og[0] = grg; og[H] = gig; og[2 * H] = ghn;
          if (gin_out != nullptr) gin_out[tr * H + j] = gin;
        }
        float* gr = gs + (r0 + i) * G3 + j;
        gr[0] = grg; gr[H] = gig; gr[2 * H] = ghn;
        dh[i][u] = ghx;
        dbi[u][0] += grg; dbi[u][1] += gig; dbi[u][2] += gin; dbn[u] += ghn;
#pragma unroll
        for (int i2 = 0; i2 < kGruMaxI; ++i2) {
          dWi[u][0][i2] = fmaf(grg, xv[i2], dWi[u][0][i2]);
          dWi[u][1][i2] = fmaf(gig, xv[i2], dWi[u][1][i2]);
          dWi[u][2][i2] = fmaf(gin, xv[i2], dWi[u][2][i2]);
        }
      }
    }
    __syncwarp();
    // dh_{t-1} = dh_t z + dgh W_hh : [8 rows] x [U units] over k = 0..3H
#pragma unroll 2
    for (int k = 0; k < G3; k += 4) {
      float4 gv[kGruRpw];
#pragma unroll
      for (int i = 0; i < kGruRpw; ++i) gv[i] = *reinterpret_cast<const float4*>(grow + i * G3 + k);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float4 wv = *reinterpret_cast<const float4*>(wrow + (32 * u) * LDT + k);
#pragma unroll
        for (int i = 0; i < kGruRpw; ++i) fma4(dh[i][u], gv[i], wv);
      }
    }
    __syncwarp();
  }

  // parameter-gradient partials: warps combined in index order
  __syncthreads();
  float* red = sm;   // [8 warps][V][32]  (fits in Wt + gs for every supported H)
  {
    float* mine = red + (w * V) * 32 + l;
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {
#pragma unroll
        for (int i2 = 0; i2 < kGruMaxI; ++i2) mine[((u * 3 + g) * 8 + i2) * 32] = dWi[u][g][i2];
        mine[(24 * U + u * 3 + g) * 32] = dbi[u][g];
      }
      mine[(27 * U + u) * 32] = dbn[u];
    }
  }
  __syncthreads();
  for (int o = tid; o < V * 32; o += kGruThreads) {
    float s = 0.f;
#pragma unroll
    for (int ww = 0; ww < kGruThreads / 32; ++ww) s += red[ww * V * 32 + o];
    parts[(int64_t)blockIdx.x * (V * 32) + o] = s;
  }
}

// dW_hh = dgh^T . h_prev over the T*N rows: [3H, H] accumulated in registers by a persistent CTA per SM (thread = 6 gate
// rows x H/8 hidden columns), 32-row chunks of both operands double-buffered in shared memory with cp.async; h_prev is
// read in place from the saved buffer (row stride 5H).  cuBLAS takes 6.1 ms for this 288 x 96 x 3.3 M product at the
// C2 shape (30 TFLOP/s); rows are summed in index order per CTA, the per-CTA tiles by the caller.
constexpr int kGruWgChunk = 32;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int H>
__global__ void __launch_bounds__(4 * H, 1)
k_gru_wgrad(const float* __restrict__ dgh, const float* __restrict__ saved, int64_t rows, float* __restrict__ parts) {
  constexpr int G3 = 3 * H, CW = H / 8, CH = kGruWgChunk, NT = 4 * H;
  extern __shared__ __align__(16) float sm[];
  float* Gs = sm;                  // [2][CH][3H]
  float* Ps = sm + 2 * CH * G3;    // [2][CH][H]
  const int tid = threadIdx.x, b = tid & 7, a = tid >> 3;
  f32x2 acc[6][CW / 2];          // pairs of adjacent hidden columns (FFMA2: the gate value is duplicated, h_prev pairs are natural)
#pragma unroll
  for (int m = 0; m < 6; ++m)
#pragma unroll
    for (int n = 0; n < CW / 2; ++n) acc[m][n] = 0ull;
  const int64_t nchunks = ceil_div<int64_t>(rows, CH);

  auto load = [&](int stage, int64_t c) {
    const int64_t r0 = c * CH;
    const int nr = (int)min((int64_t)CH, rows - r0);
    float* gd = Gs + stage * CH * G3;
    const float* gsrc = dgh + r0 * G3;
    for (int i = tid; i < CH * G3 / 4; i += NT) {
      if (i / (G3 / 4) < nr) cp_async16(gd + 4 * i, gsrc + 4 * i);
      else *reinterpret_cast<float4*>(gd + 4 * i) = f4zero();
    }
    float* pd = Ps + stage * CH * H;
    for (int i = tid; i < CH * H / 4; i += NT) {
      const int r = i / (H / 4), c4 = i % (H / 4);
      if (r < nr) cp_async16(pd + 4 * i, saved + (r0 + r) * 5 * H + 4 * c4);
      else *reinterpret_cast<float4*>(pd + 4 * i) = f4zero();
    }
    cp_async_commit();
  };

  int stage = 0;
  if ((int64_t)blockIdx.x < nchunks) load(0, blockIdx.x);
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t nxt = c + gridDim.x;
    if (nxt < nchunks) { load(stage ^ 1, nxt); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
    const float* g = Gs + stage * CH * G3 + 6 * a;
    const float* p = Ps + stage * CH * H + CW * b;
#pragma unroll 4
    for (int r = 0; r < CH; ++r) {
      const float2 g0 = *reinterpret_cast<const float2*>(g + r * G3);
      const float2 g1 = *reinterpret_cast<const float2*>(g + r * G3 + 2);
      const float2 g2 = *reinterpret_cast<const float2*>(g + r * G3 + 4);
      const f32x2 gv[6] = {f32x2_dup(g0.x), f32x2_dup(g0.y), f32x2_dup(g1.x), f32x2_dup(g1.y), f32x2_dup(g2.x), f32x2_dup(g2.y)};
#pragma unroll
      for (int q = 0; q < CW / 4; ++q) {
        const ulonglong2 pv = *reinterpret_cast<const ulonglong2*>(p + r * H + 4 * q);
#pragma unroll
        for (int m = 0; m < 6; ++m) {
          acc[m][2 * q + 0] = ffma2(gv[m], pv.x, acc[m][2 * q + 0]);
          acc[m][2 * q + 1] = ffma2(gv[m], pv.y, acc[m][2 * q + 1]);
        }
      }
    }
    __syncthreads();
    stage ^= 1;
  }
  float* out = parts + (int64_t)blockIdx.x * G3 * H;
#pragma unroll
  for (int m = 0; m < 6; ++m)
#pragma unroll
    for (int q = 0; q < CW / 4; ++q) {
      ulonglong2 v;
      v.x = acc[m][2 * q]; v.y = acc[m][2 * q + 1];
      *reinterpret_cast<ulonglong2*>(out + (6 * a + m) * H + CW * b + 4 * q) = v;
    }
}

int64_t gru_wgrad_tiles(int64_t rows) {
  return std::min<int64_t>(num_sms(), ceil_div<int64_t>(rows, kGruWgChunk));
}

template <int H>
int gru_wgrad_launch(const float* dgh, const float* saved, int64_t rows, float* parts, cudaStream_t s) {
  const int64_t smem = 2ll * kGruWgChunk * 4 * H * 4;
  SLDM_OPT_IN_SMEM((k_gru_wgrad<H>), smem);
  k_gru_wgrad<H><<<(unsigned)gru_wgrad_tiles(rows), 4 * H, (size_t)smem, s>>>(dgh, saved, rows, parts);
  SLDM_LAUNCH_CHECK("k_gru_wgrad");
  return SLDM_OK;
}

int64_t gru_fwd_smem(int H, int T, int I) {
  return 4ll * (3ll * H * (H + 4) + (int64_t)kGruRows * H + 3ll * H * kGruLdi + 4ll * H + (int64_t)kGruRows * T * I);
}
int64_t gru_bwd_smem(int H, int T, int I) {
  const int64_t main_part = (int64_t)H * (3 * H + 4) + (int64_t)kGruRows * 3 * H;
  const int64_t red = 8ll * 28 * (H / 32) * 32;
  return 4ll * (std::max(main_part, red) + (int64_t)kGruRows * T * I);
}

int gru_check(const char* who, int64_t N, int32_t T, int32_t I, int32_t H) {
  SLDM_REQUIRE(N >= 0 && T >= 1 && I >= 1, SLDM_EINVAL, "%s: bad sizes N=%lld T=%d I=%d", who, (long long)N, T, I);
  SLDM_REQUIRE(H == 32 || H == 64 || H == 96, SLDM_EUNSUPPORTED, "%s: hidden size %d (fused path: 32, 64, 96)", who, H);
  SLDM_REQUIRE(I <= kGruMaxI, SLDM_EUNSUPPORTED, "%s: input width %d > %d", who, I, kGruMaxI);
  SLDM_REQUIRE(std::max(gru_fwd_smem(H, T, I), gru_bwd_smem(H, T, I)) <= kGruSmemMax, SLDM_EUNSUPPORTED,
               "%s: %d steps x %d inputs do not fit the shared-memory tile", who, T, I);
  SLDM_REQUIRE(ceil_div<int64_t>(N, kGruRows) < (1ll << 31), SLDM_EUNSUPPORTED, "%s: N too large", who);
  return SLDM_OK;
}

template <int H>
int gru_fwd_launch(const float* x, int64_t N, int T, int I, const float* W_ih, const float* W_hh, const float* b_ih,
                   const float* b_hh, float* h_last, float* saved, cudaStream_t s) {
  const int64_t smem = gru_fwd_smem(H, T, I);
  const unsigned grid = (unsigned)ceil_div<int64_t>(N, kGruRows);
  if (saved != nullptr) {
    SLDM_OPT_IN_SMEM((k_gru_fwd<H, true>), kGruSmemMax);
    k_gru_fwd<H, true><<<grid, kGruThreads, (size_t)smem, s>>>(x, N, T, I, W_ih, W_hh, b_ih, b_hh, h_last, saved);
  } else {
    SLDM_OPT_IN_SMEM((k_gru_fwd<H, false>), kGruSmemMax);
    k_gru_fwd<H, false><<<grid, kGruThreads, (size_t)smem, s>>>(x, N, T, I, W_ih, W_hh, b_ih, b_hh, h_last, nullptr);
  }
  SLDM_LAUNCH_CHECK("k_gru_fwd");
  return SLDM_OK;
}

template <int H>
int gru_bwd_launch(const float* x, int64_t N, int T, int I, const float* W_hh, const float* dh_last,
                   const float* saved, float* dgh, float* gin, float* parts, cudaStream_t s) {
  const int64_t smem = gru_bwd_smem(H, T, I);
  const unsigned grid = (unsigned)ceil_div<int64_t>(N, kGruRows);
  SLDM_OPT_IN_SMEM((k_gru_bwd<H>), kGruSmemMax);
  k_gru_bwd<H><<<grid, kGruThreads, (size_t)smem, s>>>(x, N, T, I, W_hh, dh_last, saved, dgh, gin, parts);
  SLDM_LAUNCH_CHECK("k_gru_bwd");
  return SLDM_OK;
}

}  // namespace
}  // namespace sldm

using namespace sldm;

extern "C" int sldm_gru_supported(int32_t T, int32_t I, int32_t H) {
  return (H == 32 || H == 64 || H == 96) && T >= 1 && I >= 1 && I <= kGruMaxI &&
         std::max(gru_fwd_smem(H, T, I), gru_bwd_smem(H, T, I)) <= kGruSmemMax;
}

extern "C" int64_t sldm_gru_partial_rows(int64_t N) { return N < 0 ? -1 : ceil_div<int64_t>(N, kGruRows); }
extern "C" int64_t sldm_gru_partial_width(int32_t H) { return H > 0 && H % 32 == 0 ? 28ll * H : -1; }

extern "C" int sldm_gru_forward(const float* x, int64_t N, int32_t T, int32_t I, int32_t H,
                                const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                                float* h_last, float* saved, sldm_stream_t stream) {
  int rc = gru_check("sldm_gru_forward", N, T, I, H);
  if (rc) return rc;
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(x && W_ih && W_hh && b_ih && b_hh && h_last, SLDM_EINVAL, "sldm_gru_forward: NULL pointer");
  SLDM_REQUIRE((reinterpret_cast<uintptr_t>(W_hh) & 15u) == 0, SLDM_EINVAL, "sldm_gru_forward: W_hh must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (H) {
    case 32: return gru_fwd_launch<32>(x, N, T, I, W_ih, W_hh, b_ih, b_hh, h_last, saved, s);
    case 64: return gru_fwd_launch<64>(x, N, T, I, W_ih, W_hh, b_ih, b_hh, h_last, saved, s);
    default: return gru_fwd_launch<96>(x, N, T, I, W_ih, W_hh, b_ih, b_hh, h_last, saved, s);
  }
}

extern "C" int sldm_gru_backward(const float* x, int64_t N, int32_t T, int32_t I, int32_t H, const float* W_hh,
                                 const float* dh_last, const float* saved, float* dgh, float* dgi_n, float* partials,
                                 int64_t partial_rows, sldm_stream_t stream) {
  int rc = gru_check("sldm_gru_backward", N, T, I, H);
  if (rc) return rc;
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(x && W_hh && dh_last && saved && dgh && partials, SLDM_EINVAL, "sldm_gru_backward: NULL pointer");
  SLDM_REQUIRE(partial_rows >= ceil_div<int64_t>(N, kGruRows), SLDM_EWORKSPACE,
               "sldm_gru_backward: %lld partial rows < %lld", (long long)partial_rows,
               (long long)ceil_div<int64_t>(N, kGruRows));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (H) {
    case 32: return gru_bwd_launch<32>(x, N, T, I, W_hh, dh_last, saved, dgh, dgi_n, partials, s);
    case 64: return gru_bwd_launch<64>(x, N, T, I, W_hh, dh_last, saved, dgh, dgi_n, partials, s);
    default: return gru_bwd_launch<96>(x, N, T, I, W_hh, dh_last, saved, dgh, dgi_n, partials, s);
  }
}

extern "C" int64_t sldm_gru_wgrad_tiles(int64_t N, int32_t T) {
  if (N < 0 || T < 1) return -1;
  return gru_wgrad_tiles(N * (int64_t)T);
}

extern "C" int sldm_gru_wgrad(const float* dgh, const float* saved, int64_t N, int32_t T, int32_t H,
                              float* partials, int64_t partial_tiles, sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && T >= 1, SLDM_EINVAL, "sldm_gru_wgrad: bad sizes N=%lld T=%d", (long long)N, T);
  SLDM_REQUIRE(H == 32 || H == 64 || H == 96, SLDM_EUNSUPPORTED, "sldm_gru_wgrad: hidden size %d (fused path: 32, 64, 96)", H);
  const int64_t rows = N * (int64_t)T;
  if (rows == 0) return SLDM_OK;
  SLDM_REQUIRE(dgh && saved && partials, SLDM_EINVAL, "sldm_gru_wgrad: NULL pointer");
  SLDM_REQUIRE(partial_tiles >= gru_wgrad_tiles(rows), SLDM_EWORKSPACE, "sldm_gru_wgrad: %lld partial tiles < %lld",
               (long long)partial_tiles, (long long)gru_wgrad_tiles(rows));
  SLDM_REQUIRE(((reinterpret_cast<uintptr_t>(dgh) | reinterpret_cast<uintptr_t>(saved) |
                 reinterpret_cast<uintptr_t>(partials)) & 15u) == 0, SLDM_EINVAL, "sldm_gru_wgrad: buffers must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (H) {
    case 32: return gru_wgrad_launch<32>(dgh, saved, rows, partials, s);
    case 64: return gru_wgrad_launch<64>(dgh, saved, rows, partials, s);
    default: return gru_wgrad_launch<96>(dgh, saved, rows, partials, s);
  }
}

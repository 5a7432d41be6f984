// sage_bwd.cu -- backward of one SageBlock layer (everything except the
// transpose gather, which is segment_reduce.cu run over the transpose CSR).
//
// Restates the autograd of src/models/blocks/sageblock.py:18-19:
//   y   = xhat*gamma + beta ;  dy = dout * (y > 0 ? 1 : slope)
//   dgamma = sum_i dy*xhat ;   dbeta = sum_i dy
//   dzh = dy*gamma ;  dz = rstd * (dzh - mean(dzh) - xhat*mean(dzh*xhat))
//   db_l = sum_i dz ; dW_l = dz^T agg ; dW_r = dz^T x
//   dagg = dz W_l   (stored pre-divided by max(deg,1): the mean's backward)
//   dxroot = dz W_r
// Kernels:
//   k_ln_bwd_dgrad : LN/activation backward for a tile of rows kept in shared
//                    memory, column partial sums, then the two data-gradient
//                    GEMMs straight out of that tile.
//   k_wgrad        : split-K (over node slabs) weight-gradient GEMM.
//   k_reduce_parts : fixed-order reduction of the per-CTA partials (no atomics).
#include "common.cuh"
#include <algorithm>

namespace sldm {

constexpr int kBKb = 16;

__device__ __forceinline__ float f4at(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

template <int TX, int TN, int TM>
struct BTile {
  static constexpr int TY = 256 / TX;
  static constexpr int BN = TX * TN;
  static constexpr int BM = TY * TM;
  __device__ static __forceinline__ int rowl(int ty, int i) {
    return (TM % 4 == 0) ? (i / 4) * (TY * 4) + ty * 4 + (i % 4) : ty * TM + i;
  }
  __device__ static __forceinline__ int coll(int tx, int j) {
    return (j / 4) * (TX * 4) + tx * 4 + (j % 4);
  }
};

constexpr int kQMax = 8;  // Fout <= 256 -> at most 8 columns per lane

// dynamic smem: Az[BM][KP+4] | Ws[2][kBKb][BN] ; Az is reused for the column partials
template <int TX, int TN, int TM>
__global__ void __launch_bounds__(256)
k_ln_bwd_dgrad(const float* __restrict__ dout, const float* __restrict__ xhat,
               const float* __restrict__ rstd, const float* __restrict__ gamma,
               const float* __restrict__ beta, float slope,
               int64_t N, int Fin, int Fout, int KP,
               const float* __restrict__ W_l, const float* __restrict__ W_r,
               const int32_t* __restrict__ rowptr_dst,
               float* __restrict__ dz, float* __restrict__ dagg, float* __restrict__ dxroot,
               float* __restrict__ colpart, int need_dx, int vec_w, int vec_o) {
  using T = BTile<TX, TN, TM>;
  constexpr int BM = T::BM, BN = T::BN;
  constexpr int W_F4 = kBKb * (BN / 4);
  constexpr int W_PT = (W_F4 + 255) / 256;
  extern __shared__ __align__(16) float smem[];
  const int AZS = KP + 4;
  float* Az = smem;
  float* Ws = smem + (size_t)BM * AZS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid % TX, ty = tid / TX;
  const int Q = (Fout + 31) / 32;
  const float fF = (float)Fout;

  float g[kQMax], be[kQMax], pg[kQMax], pb[kQMax], pdb[kQMax];
#pragma unroll
  for (int q = 0; q < kQMax; ++q) {
    const int c = lane + 32 * q;
    const bool cv = q < Q && c < Fout;
    g[q] = cv ? __ldg(gamma + c) : 0.f;
    be[q] = cv ? __ldg(beta + c) : 0.f;
    pg[q] = pb[q] = pdb[q] = 0.f;
  }

  const int64_t ntiles = ceil_div<int64_t>(N, BM);
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * BM;
    // ---------------- phase 1: dz tile ----------------
    for (int r = warp; r < BM; r += 8) {
      const int64_t row = row0 + r;
      float* az = Az + (size_t)r * AZS;
      if (row < N) {
        const float rs = __ldg(rstd + row);
        float dzh[kQMax], xh[kQMax];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < kQMax; ++q) {
          const int c = lane + 32 * q;
          dzh[q] = 0.f; xh[q] = 0.f;
          if (q < Q && c < Fout) {
            const float d = __ldg(dout + row * Fout + c);
            const float h = __ldg(xhat + row * Fout + c);
            const float y = fmaf(h, g[q], be[q]);
            const float dy = y > 0.f ? d : d * slope;
            pg[q] = fmaf(dy, h, pg[q]);
            pb[q] += dy;
            const float t = dy * g[q];
            dzh[q] = t; xh[q] = h;
            s1 += t;
            s2 = fmaf(t, h, s2);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          s1 += __shfl_xor_sync(0xffffffffu, s1, o);
          s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float c1 = __fdiv_rn(s1, fF), c2 = __fdiv_rn(s2, fF);
#pragma unroll
        for (int q = 0; q < kQMax; ++q) {
          const int c = lane + 32 * q;
          if (q < Q && c < KP) {
            float v = 0.f;
            if (c < Fout) {
              v = rs * (dzh[q] - c1 - xh[q] * c2);
              pdb[q] += v;
              dz[row * Fout + c] = v;
            }
            az[c] = v;
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < kQMax; ++q) {
          const int c = lane + 32 * q;
          if (q < Q && c < KP) az[c] = 0.f;
        }
      }
    }
    __syncthreads();
    if (!need_dx) { __syncthreads(); continue; }

    // ---------------- phase 2: [dagg | dxroot] = dz_tile * [W_l | W_r] ----------------
    const int nkc = KP / kBKb;
    const int nnb = (Fin + BN - 1) / BN;
    for (int pass = 0; pass < 2; ++pass) {
      const float* __restrict__ W = pass == 0 ? W_l : W_r;
      float* __restrict__ O = pass == 0 ? dagg : dxroot;
      for (int nb = 0; nb < nnb; ++nb) {
        const int n0 = nb * BN;
        float acc[TM][TN];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
        float4 rw[W_PT];
        auto fetch = [&](int kc) {
#pragma unroll
          for (int i = 0; i < W_PT; ++i) {
            const int f = tid + i * 256;
            if (f < W_F4) {
              const int k = kc * kBKb + f / (BN / 4);
              const int n = n0 + (f % (BN / 4)) * 4;
              float4 v = f4zero();
              if (k < Fout && n < Fin) {
                const float* p = W + (int64_t)k * Fin + n;
                if (vec_w) v = ldg4(p);
                else {
                  v.x = __ldg(p);
                  v.y = (n + 1 < Fin) ? __ldg(p + 1) : 0.f;
                  v.z = (n + 2 < Fin) ? __ldg(p + 2) : 0.f;
                  v.w = (n + 3 < Fin) ? __ldg(p + 3) : 0.f;
                }
              }
              rw[i] = v;
            }
          }
        };
        auto stash = [&](int buf) {
#pragma unroll
          for (int i = 0; i < W_PT; ++i) {
            const int f = tid + i * 256;
            if (f < W_F4) st4(Ws + (size_t)buf * kBKb * BN + (size_t)f * 4, rw[i]);
          }
        };
        fetch(0);
        stash(0);
        __syncthreads();
        for (int kc = 0; kc < nkc; ++kc) {
          const int buf = kc & 1;
          if (kc + 1 < nkc) fetch(kc + 1);
          const float* wsb = Ws + (size_t)buf * kBKb * BN;
#pragma unroll
          for (int k4 = 0; k4 < kBKb; k4 += 4) {
            float4 a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i)
              a[i] = *reinterpret_cast<const float4*>(Az + (size_t)T::rowl(ty, i) * AZS + kc * kBKb + k4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              float4 w[TN / 4];
#pragma unroll
              for (int j4 = 0; j4 < TN / 4; ++j4)
                w[j4] = *reinterpret_cast<const float4*>(wsb + (k4 + kk) * BN + T::coll(tx, j4 * 4));
#pragma unroll
              for (int i = 0; i < TM; ++i) {
                const float av = f4at(a[i], kk);
#pragma unroll
                for (int j4 = 0; j4 < TN / 4; ++j4) {
                  acc[i][j4 * 4 + 0] = fmaf(av, w[j4].x, acc[i][j4 * 4 + 0]);
                  acc[i][j4 * 4 + 1] = fmaf(av, w[j4].y, acc[i][j4 * 4 + 1]);
                  acc[i][j4 * 4 + 2] = fmaf(av, w[j4].z, acc[i][j4 * 4 + 2]);
                  acc[i][j4 * 4 + 3] = fmaf(av, w[j4].w, acc[i][j4 * 4 + 3]);
                }
              }
            }
          }
          if (kc + 1 < nkc) stash(buf ^ 1);
          __syncthreads();
        }
        // store
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const int64_t row = row0 + T::rowl(ty, i);
          if (row >= N) continue;
          float inv_cnt = 1.f;
          if (pass == 0) {
            int deg = __ldg(rowptr_dst + row + 1) - __ldg(rowptr_dst + row);
            deg = deg < 1 ? 1 : (deg > 16777216 ? 16777216 : deg);
            inv_cnt = (float)deg;
          }
          float* orow = O + row * Fin;
#pragma unroll
          for (int j4 = 0; j4 < TN / 4; ++j4) {
            const int n = n0 + T::coll(tx, j4 * 4);
            float o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              o[q] = pass == 0 ? __fdiv_rn(acc[i][j4 * 4 + q], inv_cnt) : acc[i][j4 * 4 + q];
            if (vec_o) { if (n < Fin) st4(orow + n, make_float4(o[0], o[1], o[2], o[3])); }
            else {
#pragma unroll
              for (int q = 0; q < 4; ++q) if (n + q < Fin) orow[n + q] = o[q];
            }
          }
        }
      }
    }
    __syncthreads();  // Az is rewritten by the next tile
  }

  // ---------------- column partials: fixed warp order ----------------
  __syncthreads();
  float* cp = smem;  // [8 warps][3][32*Q]
  const int CW = 32 * Q;
#pragma unroll
  for (int q = 0; q < kQMax; ++q) {
    if (q < Q) {
      cp[(warp * 3 + 0) * CW + lane + 32 * q] = pg[q];
      cp[(warp * 3 + 1) * CW + lane + 32 * q] = pb[q];
      cp[(warp * 3 + 2) * CW + lane + 32 * q] = pdb[q];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < 3 * Fout; idx += 256) {
    const int which = idx / Fout, c = idx % Fout;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += cp[(w * 3 + which) * CW + c];
    colpart[(int64_t)blockIdx.x * 3 * Fout + idx] = s;
  }
}

// ---- LayerNorm + activation backward alone (the tensor path does the GEMMs elsewhere) ----
// Streaming kernel: a warp takes 4 consecutive rows per step, a lane holds VPL float4 column slices of each row.
// dz is written, the column sums (dgamma, dbeta, db_l) are kept per lane and reduced per CTA in warp order.
// Bound: HBM, 3*N*Fout*4 bytes.
template <int VPL>
__global__ void __launch_bounds__(256)
k_ln_bwd_rows(const float* __restrict__ dout, const float* __restrict__ xhat, const float* __restrict__ rstd,
              const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
              int64_t N, int Fout, float* __restrict__ dz, float* __restrict__ colpart) {
  constexpr int R = 4;
  __shared__ float cp[8][3][128 * VPL];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float fF = (float)Fout;
  float4 g[VPL], be[VPL], pg[VPL], pb[VPL], pdb[VPL];
  bool cv[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int c = (lane + 32 * q) * 4;
    cv[q] = c < Fout;
    g[q] = cv[q] ? ldg4(gamma + c) : f4zero();
    be[q] = cv[q] ? ldg4(beta + c) : f4zero();
    pg[q] = pb[q] = pdb[q] = f4zero();
  }
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + warp) * R; r0 < N; r0 += nwarps * R) {
    float4 d[R][VPL], h[R][VPL];
    float rs[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int64_t row = r0 + i;
      const bool rv = row < N;
      rs[i] = rv ? __ldg(rstd + row) : 0.f;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const bool v = rv && cv[q];
        d[i][q] = v ? ldg4(dout + row * Fout + (lane + 32 * q) * 4) : f4zero();
        h[i][q] = v ? ldg4(xhat + row * Fout + (lane + 32 * q) * 4) : f4zero();
      }
    }
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int64_t row = r0 + i;
      float s1 = 0.f, s2 = 0.f;
      float4 t[VPL];
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        const float* dd = reinterpret_cast<const float*>(&d[i][q]);
        const float* hh = reinterpret_cast<const float*>(&h[i][q]);
        const float* gg = reinterpret_cast<const float*>(&g[q]);
        const float* bb = reinterpret_cast<const float*>(&be[q]);
        float* tt = reinterpret_cast<float*>(&t[q]);
        float* ppg = reinterpret_cast<float*>(&pg[q]);
        float* ppb = reinterpret_cast<float*>(&pb[q]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float y = fmaf(hh[e], gg[e], bb[e]);
          const float dy = y > 0.f ? dd[e] : dd[e] * slope;
          ppg[e] = fmaf(dy, hh[e], ppg[e]);
          ppb[e] += dy;
          tt[e] = dy * gg[e];
          s1 += tt[e];
          s2 = fmaf(tt[e], hh[e], s2);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      const float c1 = __fdiv_rn(s1, fF), c2 = __fdiv_rn(s2, fF);
      if (row < N) {
#pragma unroll
        for (int q = 0; q < VPL; ++q) {
          if (cv[q]) {
            const float* hh = reinterpret_cast<const float*>(&h[i][q]);
            const float* tt = reinterpret_cast<const float*>(&t[q]);
            float* pp = reinterpret_cast<float*>(&pdb[q]);
            float4 o4;
            float* oo = reinterpret_cast<float*>(&o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              oo[e] = rs[i] * (tt[e] - c1 - hh[e] * c2);
              pp[e] += oo[e];
            }
            st4(dz + row * Fout + (lane + 32 * q) * 4, o4);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int c = (lane + 32 * q) * 4;
    st4(&cp[warp][0][c], pg[q]);
    st4(&cp[warp][1][c], pb[q]);
    st4(&cp[warp][2][c], pdb[q]);
  }
  __syncthreads();
  for (int idx = tid; idx < 3 * Fout; idx += 256) {
    const int which = idx / Fout, c = idx % Fout;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += cp[w][which][c];
    colpart[(int64_t)blockIdx.x * 3 * Fout + idx] = s;
  }
}

// ---- weight gradient: part[s][which][m][n] = sum_{r in slab s} dz[r][m] * cat_which[r][n] ----
__global__ void __launch_bounds__(256)
k_wgrad(const float* __restrict__ dz, const float* __restrict__ agg, const float* __restrict__ x,
        int64_t N, int Fin, int Fout, int64_t rows_per_slab, int MB, int NB,
        float* __restrict__ part, int vec_m, int vec_n) {
  __shared__ __align__(16) float As[2][kBKb][128];
  __shared__ __align__(16) float Bs[2][kBKb][128];
  const int tid = threadIdx.x, tn = tid % 16, tm = tid / 16;
  const int s = blockIdx.x;
  int t = blockIdx.y;
  const int nb = t % NB; t /= NB;
  const int mb = t % MB; t /= MB;
  const int which = t;  // 0: agg (dW_l), 1: x (dW_r)
  const float* __restrict__ B = which == 0 ? agg : x;
  const int m0 = mb * 128, n0 = nb * 128;
  const int64_t r0 = (int64_t)s * rows_per_slab;
  const int64_t r1 = (r0 + rows_per_slab < N) ? r0 + rows_per_slab : N;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto ld_guard = [&](const float* __restrict__ P, int64_t row, int ld, int c, int cmax, int vec) -> float4 {
    float4 v = f4zero();
    if (row < r1 && c < cmax) {
      const float* p = P + row * ld + c;
      if (vec) v = ldg4(p);
      else {
        v.x = __ldg(p);
        v.y = (c + 1 < cmax) ? __ldg(p + 1) : 0.f;
        v.z = (c + 2 < cmax) ? __ldg(p + 2) : 0.f;
        v.w = (c + 3 < cmax) ? __ldg(p + 3) : 0.f;
      }
    }
    return v;
  };
  auto fetch = [&](int64_t rc) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * 256;
      const int k = f / 32, c4 = (f % 32) * 4;
      ra[i] = ld_guard(dz, rc + k, Fout, m0 + c4, Fout, vec_m);
      rb[i] = ld_guard(B, rc + k, Fin, n0 + c4, Fin, vec_n);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = tid + i * 256;
      st4(&As[buf][0][0] + f * 4, ra[i]);
      st4(&Bs[buf][0][0] + f * 4, rb[i]);
    }
  };
  const int64_t nch = r1 > r0 ? ceil_div<int64_t>(r1 - r0, kBKb) : 0;
  if (nch > 0) {
    fetch(r0);
    stash(0);
  }
  __syncthreads();
  for (int64_t ch = 0; ch < nch; ++ch) {
    const int buf = (int)(ch & 1);
    if (ch + 1 < nch) fetch(r0 + (ch + 1) * kBKb);
#pragma unroll
    for (int k = 0; k < kBKb; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + tm * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (ch + 1 < nch) stash(buf ^ 1);
    __syncthreads();
  }
  float* P = part + ((int64_t)s * 2 + which) * Fout * Fin;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i / 4) * 64 + tm * 4 + (i % 4);
    if (m >= Fout) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j / 4) * 64 + tn * 4 + (j % 4);
      if (n < Fin) P[(int64_t)m * Fin + n] = acc[i][j];
    }
  }
}

// out[i] = sum_{s < S} part[s*stride + i], deterministic: a CTA owns 32 consecutive outputs, its 8 warps each add the
// slices s = w, w+8, ... in ascending order, the 8 partial sums are combined in warp order.  (One thread per output
// walking all S slices was latency bound: 592 dependent loads for the 384 column sums took 52 us -- profiles/r01i.)
__global__ void __launch_bounds__(256)
k_reduce_parts(const float* __restrict__ part, int S, int64_t stride, int64_t count,
               float* __restrict__ out0, int64_t split, float* __restrict__ out1,
               int64_t split2, float* __restrict__ out2) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < count)
    for (int k = w; k < S; k += 8) s += __ldg(part + (int64_t)k * stride + i);
  sm[w][lane] = s;
  __syncthreads();
  if (w == 0 && i < count) {
    float t = sm[0][lane];
#pragma unroll
    for (int q = 1; q < 8; ++q) t += sm[q][lane];
    if (i < split) out0[i] = t;
    else if (i < split2) out1[i - split] = t;
    else out2[i - split2] = t;
  }
}

// ------------------------------------------------------------------ launch --
struct BwdPlan { int grid1; int grid_ln; int S; int64_t rows_per_slab; int MB, NB; int64_t colpart_off, part_off, tc_off, wg_off, total; };

static BwdPlan bwd_plan(int64_t N, int Fin, int Fout) {
  BwdPlan p;
  const int sms = num_sms();
  const int64_t ntiles = ceil_div<int64_t>(N > 0 ? N : 1, 64);
  p.grid1 = (int)std::min<int64_t>(ntiles, (int64_t)sms * 2);
  p.MB = ceil_div(Fout, 128);
  p.NB = ceil_div(Fin, 128);
  const int per = 2 * p.MB * p.NB;
  const int target = std::max(1, 4 * sms / per);
  p.S = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(N > 0 ? N : 1, 128), target));
  p.rows_per_slab = round_up<int64_t>(ceil_div<int64_t>(N > 0 ? N : 1, p.S), kBKb);
  p.S = (int)ceil_div<int64_t>(N > 0 ? N : 1, p.rows_per_slab);
  int64_t o = 0;
  p.grid_ln = (int)std::min<int64_t>(ceil_div<int64_t>(N > 0 ? N : 1, 32), (int64_t)sms * 4);
  p.colpart_off = o; o += align_bytes((int64_t)std::max(p.grid1, p.grid_ln) * 3 * Fout * 4);
  p.part_off = o;    o += align_bytes((int64_t)p.S * 2 * Fout * Fin * 4);
  p.tc_off = o;      o += dgrad_tc_ws_bytes(Fin, Fout);
  p.wg_off = o;      o += wgrad_tc_ws_bytes(Fin, Fout);
  p.total = o;
  return p;
}

int64_t layer_backward_ws_bytes(int64_t N, int32_t Fin, int32_t Fout) {
  return bwd_plan(N, Fin, Fout).total;
}

static bool b16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int TX, int TN, int TM>
static int launch_b1(int grid, size_t smem, cudaStream_t s,
                     const float* dout, const float* xhat, const float* rstd, const float* gamma,
                     const float* beta, float slope, int64_t N, int Fin, int Fout, int KP,
                     const float* W_l, const float* W_r, const int32_t* rowptr_dst,
                     float* dz, float* dagg, float* dxroot, float* colpart, int need_dx, int vec_w, int vec_o) {
  SLDM_OPT_IN_SMEM((k_ln_bwd_dgrad<TX, TN, TM>), 160 * 1024);
  k_ln_bwd_dgrad<TX, TN, TM><<<grid, 256, smem, s>>>(dout, xhat, rstd, gamma, beta, slope, N, Fin, Fout, KP,
                                                     W_l, W_r, rowptr_dst, dz, dagg, dxroot, colpart,
                                                     need_dx, vec_w, vec_o);
  SLDM_LAUNCH_CHECK("k_ln_bwd_dgrad");
  return SLDM_OK;
}

int layer_backward_launch(const float* dout, const float* x, const float* agg,
                          const float* xhat, const float* rstd,
                          int64_t N, int32_t Fin, int32_t Fout,
                          const int32_t* rowptr_dst,
                          const float* W_l, const float* W_r,
                          const float* ln_w, const float* ln_b, float slope,
                          bool need_dx,
                          float* dW_l, float* db_l, float* dW_r, float* dln_w, float* dln_b,
                          float* dz, float* dagg, float* dxroot,
                          void* ws, int64_t ws_bytes, cudaStream_t s, int stages, bool bf16_feats) {
  SLDM_REQUIRE(Fin >= 1 && Fout >= 1, SLDM_ESHAPE, "backward: Fin=%d Fout=%d must be >= 1", Fin, Fout);
  // bf16 feature storage: x and agg are bf16 rows (only the weight gradient reads them); needs the tensor path
  SLDM_REQUIRE(!bf16_feats || N == 0 || (wgrad_bf16_eligible(N, Fin, Fout, dz, agg, x) &&
                                         (!need_dx || dgrad_tc_eligible(N, Fin, Fout, dz, dagg, dxroot))),
               SLDM_EUNSUPPORTED, "backward (bf16 features): needs Fin %% 64 == 0, Fin <= 128, Fout %% 32 == 0, Fout <= 128");
  SLDM_REQUIRE(Fout <= 256, SLDM_EUNSUPPORTED, "backward: Fout=%d > 256 is not covered by the kernels", Fout);
  const int64_t wcount = (int64_t)Fout * Fin;
  if (N == 0) {  // empty batch: all parameter gradients are zero
    SLDM_CUDA(cudaMemsetAsync(dW_l, 0, wcount * 4, s));
    SLDM_CUDA(cudaMemsetAsync(dW_r, 0, wcount * 4, s));
    SLDM_CUDA(cudaMemsetAsync(db_l, 0, Fout * 4, s));
    SLDM_CUDA(cudaMemsetAsync(dln_w, 0, Fout * 4, s));
    SLDM_CUDA(cudaMemsetAsync(dln_b, 0, Fout * 4, s));
    return SLDM_OK;
  }
  BwdPlan p = bwd_plan(N, Fin, Fout);
  SLDM_REQUIRE(ws != nullptr && ws_bytes >= p.total, SLDM_EWORKSPACE, "backward: workspace %lld < %lld bytes",
               (long long)ws_bytes, (long long)p.total);
  float* colpart = reinterpret_cast<float*>(static_cast<char*>(ws) + p.colpart_off);
  float* part = reinterpret_cast<float*>(static_cast<char*>(ws) + p.part_off);

  const int KP = round_up<int>(Fout, kBKb);
  const int vec_w = (Fin % 4 == 0) && b16(W_l) && b16(W_r);
  const int vec_o = (Fin % 4 == 0) && b16(dagg) && b16(dxroot);
  const int bn = Fin <= 32 ? 32 : (Fin <= 64 ? 64 : (Fin <= 96 ? 96 : 128));
  const int Q = ceil_div(Fout, 32);
  size_t smem_az = (size_t)64 * (KP + 4) * 4;
  size_t smem_cp = (size_t)8 * 3 * 32 * Q * 4;
  size_t smem = std::max(smem_az, smem_cp) + (size_t)2 * kBKb * bn * 4;
  // tensor path: the SIMT kernel only does the LayerNorm/activation backward (dz + column partials),
  // the two data-gradient GEMMs run on tcgen05 (sage_tc.cu, MODE_DGRAD)
  const bool tc_dgrad = need_dx && dgrad_tc_eligible(N, Fin, Fout, dz, dagg, dxroot);
  const bool simt_dx = need_dx && !tc_dgrad;
  int rc = SLDM_OK;
  int ncolparts = p.grid1;
  const bool ln_stream = !simt_dx && (Fout % 4 == 0) && b16(dout) && b16(xhat) && b16(dz) && b16(ln_w) && b16(ln_b);
  if (!(stages & SLDM_BWD_STAGE_LN)) {
    ncolparts = ln_stream ? p.grid_ln : p.grid1;   // profiling: dz / column partials of an earlier full call are reused
  } else if (ln_stream) {
    ncolparts = p.grid_ln;
    if (Fout <= 128)
      k_ln_bwd_rows<1><<<p.grid_ln, 256, 0, s>>>(dout, xhat, rstd, ln_w, ln_b, slope, N, Fout, dz, colpart);
    else
      k_ln_bwd_rows<2><<<p.grid_ln, 256, 0, s>>>(dout, xhat, rstd, ln_w, ln_b, slope, N, Fout, dz, colpart);
    SLDM_LAUNCH_CHECK("k_ln_bwd_rows");
  } else {
#define SLDM_B1(TX, TN, TM) \
  rc = launch_b1<TX, TN, TM>(p.grid1, smem, s, dout, xhat, rstd, ln_w, ln_b, slope, N, Fin, Fout, KP, W_l, W_r, \
                             rowptr_dst, dz, dagg, dxroot, colpart, simt_dx ? 1 : 0, vec_w, vec_o)
  switch (bn) {
    case 32: SLDM_B1(8, 4, 2); break;
    case 64: SLDM_B1(16, 4, 4); break;
    case 96: SLDM_B1(8, 12, 2); break;
    default: SLDM_B1(16, 8, 4); break;
  }
#undef SLDM_B1
  }
  if (rc) return rc;
  if (tc_dgrad && (stages & SLDM_BWD_STAGE_DGRAD)) {
    rc = dgrad_tc_launch(dz, N, Fin, Fout, W_l, W_r, rowptr_dst, dagg, dxroot, static_cast<char*>(ws) + p.tc_off,
                         p.total - p.tc_off, s);
    if (rc) return rc;
  }

  if (!(stages & SLDM_BWD_STAGE_WGRAD)) return SLDM_OK;
  int nparts = p.S;
  if (bf16_feats || wgrad_tc_eligible(N, Fin, Fout, dz, agg, x)) {
    part = reinterpret_cast<float*>(static_cast<char*>(ws) + p.wg_off);
    if ((rc = wgrad_tc_launch(dz, agg, x, N, Fin, Fout, part, &nparts, s, bf16_feats))) return rc;
  } else {
    dim3 grid(p.S, 2 * p.MB * p.NB);
    const int vec_m = (Fout % 4 == 0) && b16(dz);
    const int vec_n = (Fin % 4 == 0) && b16(agg) && b16(x);
    k_wgrad<<<grid, 256, 0, s>>>(dz, agg, x, N, Fin, Fout, p.rows_per_slab, p.MB, p.NB, part, vec_m, vec_n);
    SLDM_LAUNCH_CHECK("k_wgrad");
  }
  {
    const int64_t count = 2 * wcount;
    k_reduce_parts<<<(unsigned)ceil_div<int64_t>(count, 32), 256, 0, s>>>(part, nparts, 2 * wcount, count,
                                                                         dW_l, wcount, dW_r, count, nullptr);
    SLDM_LAUNCH_CHECK("k_reduce_parts(dW)");
    const int64_t c3 = 3 * (int64_t)Fout;
    k_reduce_parts<<<(unsigned)ceil_div<int64_t>(c3, 32), 256, 0, s>>>(colpart, ncolparts, c3, c3,
                                                                      dln_w, Fout, dln_b, 2 * (int64_t)Fout, db_l);
    SLDM_LAUNCH_CHECK("k_reduce_parts(cols)");
  }
  return SLDM_OK;
}

}  // namespace sldm

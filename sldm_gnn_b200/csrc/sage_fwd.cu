// sage_fwd.cu -- fused projection + root add + bias + LayerNorm + activation.
//
//   z    = agg W_l^T + b_l + x W_r^T          (PyG SAGEConv.forward: lin_l(out) + lin_r(x))
//   xhat = (z - mean(z)) * rstd(z)             (torch.nn.LayerNorm, biased variance, eps)
//   out  = act(xhat * gamma + beta)            (LeakyReLU(slope) | ReLU == slope 0)
// i.e. everything src/models/blocks/sageblock.py:18-19 does after the aggregation,
// in one pass over [N,Fin] x2 -> [N,Fout]; the reference runs 2 GEMMs + add + 3
// elementwise launches.
//
// FP32 FMA path.  The contraction is treated as one GEMM with K = 2*Fin over the
// concatenated operand [agg | x] and the packed, transposed weight
// Wt[k][n] = (k < FinP ? W_l[n][k] : W_r[n][k-FinP]), zero padded to the tile.
// One CTA owns BM full rows, so the LayerNorm statistics are an in-register /
// shuffle reduction over the TX lanes that share a row.
//
// Bound at F >= 64: FP32 pipe (4*N*Fin*Fout flop); HBM bytes N*4*(2 Fin + Fout)
// (+ N*4*Fout + 4N when xhat/rstd are saved).
#include "common.cuh"

namespace sldm {

constexpr int kBK = 16;  // K chunk

__global__ void __launch_bounds__(256)
k_pack_weights(const float* __restrict__ W_l, const float* __restrict__ W_r,
               int Fin, int Fout, int FinP, int BN, float* __restrict__ Wt) {
  const int64_t total = (int64_t)2 * FinP * BN;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % BN);
    const int k = (int)(i / BN);
    const int kk = k < FinP ? k : k - FinP;
    const float* W = k < FinP ? W_l : W_r;
    Wt[i] = (n < Fout && kk < Fin) ? W[(int64_t)n * Fin + kk] : 0.f;
  }
}

template <int TX, int TN, int TM>
struct Tile {
  static constexpr int TY = 256 / TX;
  static constexpr int BN = TX * TN;
  static constexpr int BM = TY * TM;
  static_assert(TN % 4 == 0, "TN must be a multiple of 4");
  __device__ static __forceinline__ int rowl(int ty, int i) {
    return (TM % 4 == 0) ? (i / 4) * (TY * 4) + ty * 4 + (i % 4) : ty * TM + i;
  }
  __device__ static __forceinline__ int coll(int tx, int j) {
    return (j / 4) * (TX * 4) + tx * 4 + (j % 4);
  }
};

__device__ __forceinline__ float f4get(const float4& v, int k) {
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// 4 consecutive k of one operand row, zero filled outside [0,Fin) x [0,N)
__device__ __forceinline__ float4 load_a4(const float* __restrict__ A, int64_t row, int64_t N,
                                          int k, int Fin, bool vec) {
  if (row >= N || k >= Fin) return f4zero();
  const float* p = A + row * Fin + k;
  if (vec) return ldg4(p);
  float4 r;
  r.x = __ldg(p);
  r.y = (k + 1 < Fin) ? __ldg(p + 1) : 0.f;
  r.z = (k + 2 < Fin) ? __ldg(p + 2) : 0.f;
  r.w = (k + 3 < Fin) ? __ldg(p + 3) : 0.f;
  return r;
}

template <int TX, int TN, int TM>
__global__ void __launch_bounds__(256)
k_sage_proj_fwd(const float* __restrict__ agg, const float* __restrict__ x, int64_t N,
                int Fin, int Fout, int FinP, const float* __restrict__ Wt,
                const float* __restrict__ b_l, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, float slope,
                float* __restrict__ out, float* __restrict__ xhat, float* __restrict__ rstd,
                int vec_in, int vec_out) {
  using T = Tile<TX, TN, TM>;
  constexpr int BM = T::BM, BN = T::BN;
  constexpr int AS = kBK + 4;
  constexpr int A_F4 = BM * (kBK / 4);          // float4 per A chunk
  constexpr int W_F4 = kBK * (BN / 4);          // float4 per W chunk
  constexpr int A_PT = (A_F4 + 255) / 256;
  constexpr int W_PT = (W_F4 + 255) / 256;
  __shared__ __align__(16) float As[2][BM][AS];
  __shared__ __align__(16) float Ws[2][kBK][BN];

  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int half = FinP / kBK, nchunks = 2 * half;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_PT], rw[W_PT];
  auto fetch = [&](int c) {
    const float* A = c < half ? agg : x;
    const int k0 = (c < half ? c : c - half) * kBK;
#pragma unroll
    for (int i = 0; i < A_PT; ++i) {
      const int f = tid + i * 256;
      if (f < A_F4) ra[i] = load_a4(A, row0 + f / (kBK / 4), N, k0 + (f % (kBK / 4)) * 4, Fin, vec_in);
    }
    const float* Wc = Wt + (int64_t)c * kBK * BN;
#pragma unroll
    for (int i = 0; i < W_PT; ++i) {
      const int f = tid + i * 256;
      if (f < W_F4) rw[i] = ldg4(Wc + (int64_t)f * 4);
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_PT; ++i) {
      const int f = tid + i * 256;
      if (f < A_F4) st4(&As[buf][f / (kBK / 4)][(f % (kBK / 4)) * 4], ra[i]);
    }
#pragma unroll
    for (int i = 0; i < W_PT; ++i) {
      const int f = tid + i * 256;
      if (f < W_F4) st4(&Ws[buf][0][0] + (int64_t)f * 4, rw[i]);
    }
  };

  fetch(0);
  stash(0);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) fetch(c + 1);
#pragma unroll
    for (int k4 = 0; k4 < kBK; k4 += 4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(&As[buf][T::rowl(ty, i)][k4]);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float4 w[TN / 4];
#pragma unroll
        for (int j4 = 0; j4 < TN / 4; ++j4)
          w[j4] = *reinterpret_cast<const float4*>(&Ws[buf][k4 + kk][T::coll(tx, j4 * 4)]);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float av = f4get(a[i], kk);
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
    if (c + 1 < nchunks) stash(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: bias, LayerNorm over the row, activation ----
  float bias[TN], g[TN], be[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int col = T::coll(tx, j);
    const bool cv = col < Fout;
    bias[j] = cv ? __ldg(b_l + col) : 0.f;
    g[j] = cv ? __ldg(gamma + col) : 0.f;
    be[j] = cv ? __ldg(beta + col) : 0.f;
  }
  const float fF = (float)Fout;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t row = row0 + T::rowl(ty, i);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float z = acc[i][j] + bias[j];
      acc[i][j] = z;
      s += (T::coll(tx, j) < Fout) ? z : 0.f;
    }
#pragma unroll
    for (int o = TX / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = __fdiv_rn(s, fF);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const float d = acc[i][j] - mean;
      v += (T::coll(tx, j) < Fout) ? d * d : 0.f;
    }
#pragma unroll
    for (int o = TX / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rs = __fdiv_rn(1.f, __fsqrt_rn(__fdiv_rn(v, fF) + eps));
    if (row < N) {
      if (rstd != nullptr && tx == 0) rstd[row] = rs;
      float* orow = out + row * Fout;
      float* hrow = xhat ? xhat + row * Fout : nullptr;
#pragma unroll
      for (int j4 = 0; j4 < TN / 4; ++j4) {
        const int col = T::coll(tx, j4 * 4);
        float h[4], o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          h[q] = (acc[i][j4 * 4 + q] - mean) * rs;
          const float y = fmaf(h[q], g[j4 * 4 + q], be[j4 * 4 + q]);
          o[q] = y > 0.f ? y : slope * y;
        }
        if (vec_out) {
          if (col < Fout) {
            st4(orow + col, make_float4(o[0], o[1], o[2], o[3]));
            if (hrow) st4(hrow + col, make_float4(h[0], h[1], h[2], h[3]));
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (col + q < Fout) { orow[col + q] = o[q]; if (hrow) hrow[col + q] = h[q]; }
        }
      }
    }
  }
}

template <int TX, int TN, int TM>
static int launch_proj(const float* agg, const float* x, int64_t N, int Fin, int Fout, int FinP,
                       const float* Wt, const float* b_l, const float* g, const float* b,
                       float eps, float slope, float* out, float* xhat, float* rstd,
                       int vec_in, int vec_out, cudaStream_t s) {
  using T = Tile<TX, TN, TM>;
  const int64_t grid = ceil_div<int64_t>(N, T::BM);
  k_sage_proj_fwd<TX, TN, TM><<<(unsigned)grid, 256, 0, s>>>(agg, x, N, Fin, Fout, FinP, Wt, b_l, g, b,
                                                            eps, slope, out, xhat, rstd, vec_in, vec_out);
  SLDM_LAUNCH_CHECK("k_sage_proj_fwd");
  return SLDM_OK;
}

static int proj_bn(int Fout) {
  return Fout <= 32 ? 32 : (Fout <= 64 ? 64 : (Fout <= 96 ? 96 : (Fout <= 128 ? 128 : 256)));
}

int64_t project_forward_ws_bytes(int64_t /*N*/, int32_t Fin, int32_t Fout) {
  const int FinP = round_up<int>(Fin > 0 ? Fin : 1, kBK);
  const int64_t simt = align_bytes((int64_t)2 * FinP * proj_bn(Fout) * 4);
  const int64_t tcb = project_forward_tc_ws_bytes(Fin, Fout);
  return simt > tcb ? simt : tcb;
}

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int project_forward_launch(const float* agg, const float* x, int64_t N, int32_t Fin, int32_t Fout,
                           const float* W_l, const float* b_l, const float* W_r,
                           const float* ln_w, const float* ln_b, float eps, float slope,
                           float* out, float* xhat, float* rstd,
                           void* ws, int64_t ws_bytes, cudaStream_t s) {
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(Fin >= 1 && Fout >= 1, SLDM_ESHAPE, "projection: Fin=%d Fout=%d must be >= 1", Fin, Fout);
  SLDM_REQUIRE(Fout <= 256, SLDM_EUNSUPPORTED, "projection: Fout=%d > 256 is not covered by the kernels", Fout);
  SLDM_REQUIRE(ws_bytes >= project_forward_ws_bytes(N, Fin, Fout) && ws != nullptr, SLDM_EWORKSPACE,
               "projection: workspace too small");
  if (project_forward_tc_eligible(N, Fin, Fout, agg, x, out, xhat))
    return project_forward_tc_launch(agg, x, N, Fin, Fout, W_l, b_l, W_r, ln_w, ln_b, eps, slope, out, xhat, rstd,
                                     ws, ws_bytes, s);
  const int FinP = round_up<int>(Fin, kBK);
  const int BN = proj_bn(Fout);
  float* Wt = static_cast<float*>(ws);
  {
    const int64_t total = (int64_t)2 * FinP * BN;
    const int grid = (int)ceil_div<int64_t>(total, 256 * 4);
    k_pack_weights<<<grid, 256, 0, s>>>(W_l, W_r, Fin, Fout, FinP, BN, Wt);
    SLDM_LAUNCH_CHECK("k_pack_weights");
  }
  const int vec_in = (Fin % 4 == 0) && a16(agg) && a16(x);
  const int vec_out = (Fout % 4 == 0) && a16(out) && a16(xhat);
  const bool small = N < (int64_t)num_sms() * 2 * 128;
#define SLDM_PROJ(TX, TN, TM) \
  return launch_proj<TX, TN, TM>(agg, x, N, Fin, Fout, FinP, Wt, b_l, ln_w, ln_b, eps, slope, out, xhat, rstd, vec_in, vec_out, s)
  switch (BN) {
    case 32:  if (small) SLDM_PROJ(8, 4, 1);  else SLDM_PROJ(8, 4, 4);
    case 64:  if (small) SLDM_PROJ(16, 4, 2); else SLDM_PROJ(16, 4, 8);
    case 96:  if (small) SLDM_PROJ(8, 12, 1); else SLDM_PROJ(8, 12, 4);
    case 128: if (small) SLDM_PROJ(16, 8, 2); else SLDM_PROJ(16, 8, 8);
    default:  if (small) SLDM_PROJ(32, 8, 4); else SLDM_PROJ(32, 8, 8);
  }
#undef SLDM_PROJ
}

}  // namespace sldm

using namespace sldm;

extern "C" int64_t sldm_sage_project_workspace_bytes(int64_t N, int32_t Fin, int32_t Fout) {
  if (N < 0 || Fin < 0 || Fout < 0) return -1;
  return project_forward_ws_bytes(N, Fin, Fout);
}

extern "C" int sldm_sage_project_forward(const float* agg, const float* x, int64_t N,
                                         int32_t Fin, int32_t Fout,
                                         const float* W_l, const float* b_l, const float* W_r,
                                         const float* ln_w, const float* ln_b,
                                         float eps, float slope,
                                         float* out, float* xhat_out, float* rstd_out,
                                         void* workspace, int64_t workspace_bytes,
                                         sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0, SLDM_EINVAL, "sldm_sage_project_forward: negative N");
  SLDM_REQUIRE(N == 0 || (agg && x && W_l && b_l && W_r && ln_w && ln_b && out), SLDM_EINVAL,
               "sldm_sage_project_forward: NULL pointer");
  return project_forward_launch(agg, x, N, Fin, Fout, W_l, b_l, W_r, ln_w, ln_b, eps, slope,
                                out, xhat_out, rstd_out, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

// tc_probe2.cu -- accumulation-accuracy probe for split-precision TF32 on tcgen05 (bring-up tool).
// D[128 x 64] = A[128 x 128] * B[64 x 128]^T with A = Ah + Al, B = Bh + Bl.
// Variants: order of the three products, flushing the TMEM accumulator to fp32 registers
// every `flush` chunks (RN adds on the CUDA cores), 3 vs 4 products, RN vs truncation split.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../sldm_gnn_b200/csrc/tc_common.cuh"
using namespace sldm;
using namespace sldm::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

struct Maps4 { CUtensorMap ah, al, bh, bl; };
constexpr int N_ = 64, K_ = 128, NCH = K_ / 32;

// order: 0 = main (hh) first, 1 = corrections (lh, hl[, ll]) first.  nprod 3 or 4.  flush = chunks per accumulator.
__global__ void __launch_bounds__(128) k_probe2(const __grid_constant__ Maps4 maps, int order, int nprod, int flush, float* D) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_tile = 128 * 128, b_tile = N_ * 128;
  uint8_t* Ah = smem; uint8_t* Al = Ah + NCH * a_tile; uint8_t* Bh = Al + NCH * a_tile; uint8_t* Bl = Bh + NCH * b_tile;
  if (tid == 0) { mbar_init(&bar_full, 1); mbar_init(&bar_mma, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  const int nacc = NCH / flush;   // accumulators, each N_ columns
  if (tid == 0) {
    mbar_expect_tx(&bar_full, NCH * 2 * (a_tile + b_tile));
    for (int c = 0; c < NCH; ++c) {
      tma_load_2d(Ah + c * a_tile, &maps.ah, c * 32, 0, &bar_full);
      tma_load_2d(Al + c * a_tile, &maps.al, c * 32, 0, &bar_full);
      tma_load_2d(Bh + c * b_tile, &maps.bh, c * 32, 0, &bar_full);
      tma_load_2d(Bl + c * b_tile, &maps.bl, c * 32, 0, &bar_full);
    }
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, N_, 0, 0);
    for (int a = 0; a < nacc; ++a) {
      uint32_t acc = 0;
      for (int cc = 0; cc < flush; ++cc) {
        const int c = a * flush + cc;
        // product list: (A,B) in {h,l}
        int pa[4], pb[4], np = 0;
        if (order == 0) { pa[np] = 0; pb[np++] = 0; }
        pa[np] = 1; pb[np++] = 0;
        pa[np] = 0; pb[np++] = 1;
        if (nprod == 4) { pa[np] = 1; pb[np++] = 1; }
        if (order == 1) { pa[np] = 0; pb[np++] = 0; }
        for (int p = 0; p < np; ++p)
          for (int ks = 0; ks < 4; ++ks) {
            uint32_t aa = smem_u32((pa[p] ? Al : Ah) + c * a_tile) + ks * 32;
            uint32_t bb = smem_u32((pb[p] ? Bl : Bh) + c * b_tile) + ks * 32;
            mma_tf32_ss(tmem_base + a * N_, make_smem_desc_sw128(aa, 16, 1024), make_smem_desc_sw128(bb, 16, 1024), idesc, acc);
            acc = 1;
          }
      }
    }
    mma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N_; c0 += 32) {
    float sum[32];
    for (int j = 0; j < 32; ++j) sum[j] = 0.f;
    for (int a = 0; a < nacc; ++a) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + a * N_ + c0, r);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) sum[j] += __uint_as_float(r[j]);
    }
    for (int j = 0; j < 32; ++j) D[(size_t)row * N_ + c0 + j] = sum[j];
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
  const int M = 128;
  std::vector<float> A((size_t)M * K_), B((size_t)N_ * K_);
  srand(7);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.15f;
  std::vector<double> exact((size_t)M * N_); std::vector<float> chain((size_t)M * N_);
  double e_chain = 0, dmax = 0, rms_chain = 0;
  for (int m = 0; m < M; ++m) for (int n = 0; n < N_; ++n) {
    double s = 0; float f = 0.f;
    for (int k = 0; k < K_; ++k) { s += (double)A[(size_t)m * K_ + k] * B[(size_t)n * K_ + k]; f = fmaf(A[(size_t)m * K_ + k], B[(size_t)n * K_ + k], f); }
    exact[(size_t)m * N_ + n] = s; chain[(size_t)m * N_ + n] = f;
    e_chain = fmax(e_chain, fabs(f - s)); dmax = fmax(dmax, fabs(s)); rms_chain += (f - s) * (f - s);
  }
  printf("K=%d N=%d max|D|=%.3f ; fp32 FMA chain vs exact: max %.3e rms %.3e\n", K_, N_, dmax, e_chain, sqrt(rms_chain / (M * N_)));
  float *dAh, *dAl, *dBh, *dBl, *dD;
  CK(cudaMalloc(&dAh, A.size() * 4)); CK(cudaMalloc(&dAl, A.size() * 4)); CK(cudaMalloc(&dBh, B.size() * 4)); CK(cudaMalloc(&dBl, B.size() * 4));
  CK(cudaMalloc(&dD, (size_t)M * N_ * 4));
  size_t smem = (size_t)NCH * 2 * (128 * 128 + N_ * 128) + 1024;
  CK(cudaFuncSetAttribute(k_probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int split = 0; split < 2; ++split) {   // 0: RN split, 1: truncation split (A raw as hi)
    std::vector<float> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
    for (size_t i = 0; i < A.size(); ++i) {
      if (split == 0) { Ah[i] = tf32_rna(A[i]); Al[i] = tf32_rna(A[i] - Ah[i]); }
      else            { Ah[i] = A[i];           Al[i] = tf32_rna(A[i] - tf32_trunc(A[i])); }
    }
    for (size_t i = 0; i < B.size(); ++i) { Bh[i] = tf32_rna(B[i]); Bl[i] = tf32_rna(B[i] - Bh[i]); }
    CK(cudaMemcpy(dAh, Ah.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dAl, Al.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dBh, Bh.data(), B.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dBl, Bl.data(), B.size() * 4, cudaMemcpyHostToDevice));
    Maps4 maps;
    if (make_tmap_2d_f32(&maps.ah, dAh, M, K_, K_, 128, 32) || make_tmap_2d_f32(&maps.al, dAl, M, K_, K_, 128, 32) ||
        make_tmap_2d_f32(&maps.bh, dBh, N_, K_, K_, N_, 32) || make_tmap_2d_f32(&maps.bl, dBl, N_, K_, K_, N_, 32)) { printf("tmap failed\n"); return 1; }
    for (int nprod = 3; nprod <= 4; ++nprod)
      for (int order = 0; order < 2; ++order)
        for (int flush = 1; flush <= NCH; flush *= 2) {
          CK(cudaMemset(dD, 0xFF, (size_t)M * N_ * 4));
          k_probe2<<<1, 128, smem>>>(maps, order, nprod, flush, dD);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 1; }
          std::vector<float> D((size_t)M * N_);
          CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
          double emax = 0, rms = 0, bias = 0;
          for (size_t i = 0; i < D.size(); ++i) { double d = D[i] - exact[i]; emax = fmax(emax, fabs(d)); rms += d * d; bias += d * (exact[i] >= 0 ? 1 : -1); }
          printf("split=%s nprod=%d order=%s chunks/acc=%d : max %.3e rms %.3e signed-bias %.3e\n", split ? "trunc" : "rn   ", nprod,
                 order ? "corr-first" : "main-first", flush, emax, sqrt(rms / D.size()), bias / D.size());
        }
  }
  return 0;
}

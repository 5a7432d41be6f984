// tc_probe.cu -- bring-up probe for the tcgen05 building blocks (not part of the product).
// One CTA computes D[128 x N] = sum_p A_p * B_p^T with kind::tf32 MMAs from TMA-loaded,
// 128B-swizzled shared-memory tiles, for every operand-major combination the SageBlock
// GEMMs need, and checks against a CPU model.  Answers: does the hardware truncate or
// round fp32 -> tf32?  how exact is the fp32 accumulation?  do the descriptors work?
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include "../sldm_gnn_b200/csrc/tc_common.cuh"

using namespace sldm;
using namespace sldm::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2);} } while (0)

struct Maps { CUtensorMap a[3]; CUtensorMap b[3]; };
struct MnCfg { uint32_t layout, lbo, sbo, kadv; int atom32; };

template <int A_MN, int B_MN>
__global__ void __launch_bounds__(128) k_probe(const __grid_constant__ Maps maps, int npairs, int N, int K, float* D, MnCfg mn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar_full, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nch = K / 32;
  const uint32_t a_tile = 128 * 128;          // bytes of one A chunk (128 rows x 128 B or 4 blocks x 32 x 128 B)
  const uint32_t b_tile = (uint32_t)N * 128;  // bytes of one B chunk
  uint8_t* Abuf = smem;                                     // [pair][chunk]
  uint8_t* Bbuf = smem + (size_t)3 * nch * a_tile;          // [pair][chunk]

  if (tid == 0) {
    mbar_init(&bar_full, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    mbar_expect_tx(&bar_full, (uint32_t)npairs * nch * (a_tile + b_tile));
    for (int p = 0; p < npairs; ++p)
      for (int c = 0; c < nch; ++c) {
        uint8_t* ad = Abuf + ((size_t)p * nch + c) * a_tile;
        uint8_t* bd = Bbuf + ((size_t)p * nch + c) * b_tile;
        if (A_MN) { for (int b = 0; b < 4; ++b) tma_load_2d(ad + b * 4096, &maps.a[p], b * 32, c * 32, &bar_full); }
        else      { tma_load_2d(ad, &maps.a[p], c * 32, 0, &bar_full); }
        if (B_MN) { for (int b = 0; b < N / 32; ++b) tma_load_2d(bd + b * 4096, &maps.b[p], b * 32, c * 32, &bar_full); }
        else      { tma_load_2d(bd, &maps.b[p], c * 32, 0, &bar_full); }
      }
    mbar_wait(&bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_tf32(128, N, A_MN, B_MN);
    uint32_t acc = 0;
    for (int p = 0; p < npairs; ++p)
      for (int c = 0; c < nch; ++c)
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t aa = smem_u32(Abuf + ((size_t)p * nch + c) * a_tile);
          uint32_t bb = smem_u32(Bbuf + ((size_t)p * nch + c) * b_tile);
          uint64_t ad = A_MN ? make_smem_desc(aa + ks * mn.kadv, mn.lbo, mn.sbo, mn.layout) : make_smem_desc_sw128(aa + ks * 32, 16, 1024);
          uint64_t bd = B_MN ? make_smem_desc(bb + ks * mn.kadv, mn.lbo, mn.sbo, mn.layout) : make_smem_desc_sw128(bb + ks * 32, 16, 1024);
          mma_tf32_ss(tmem_base, ad, bd, idesc, acc);
          acc = 1;
        }
    mma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    const int row = warp * 32 + lane;
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static std::vector<float> transpose(const std::vector<float>& a, int r, int c) {
  std::vector<float> t((size_t)r * c);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) t[(size_t)j * r + i] = a[(size_t)i * c + j];
  return t;
}

template <int A_MN, int B_MN>
static int run_case(const char* name, int N, int K, int mode /*0: raw 1xTF32, 1: 3xTF32 host split*/, MnCfg mn = MnCfg{2, 4096, 1024, 1024, 0}) {
  const int M = 128;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  srand(1234 + N + K + mode);
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
  std::vector<float> Ah(A.size()), Al(A.size()), Bh(B.size()), Bl(B.size());
  for (size_t i = 0; i < A.size(); ++i) { Ah[i] = tf32_rna(A[i]); Al[i] = tf32_rna(A[i] - Ah[i]); }
  for (size_t i = 0; i < B.size(); ++i) { Bh[i] = tf32_rna(B[i]); Bl[i] = tf32_rna(B[i] - Bh[i]); }
  const int npairs = mode == 0 ? 1 : 3;
  const std::vector<float>* Ap[3] = {mode == 0 ? &A : &Ah, &Al, &Ah};
  const std::vector<float>* Bp[3] = {mode == 0 ? &B : &Bh, &Bh, &Bl};
  float* dA[3]; float* dB[3]; float* dD;
  Maps maps;
  for (int p = 0; p < npairs; ++p) {
    std::vector<float> a = A_MN ? transpose(*Ap[p], M, K) : *Ap[p];   // MN-major: stored [K][M]
    std::vector<float> b = B_MN ? transpose(*Bp[p], N, K) : *Bp[p];
    CK(cudaMalloc(&dA[p], a.size() * 4)); CK(cudaMalloc(&dB[p], b.size() * 4));
    CK(cudaMemcpy(dA[p], a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB[p], b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    int rc;
    rc = A_MN ? make_tmap_2d_f32(&maps.a[p], dA[p], K, M, M, 32, 32, mn.atom32) : make_tmap_2d_f32(&maps.a[p], dA[p], M, K, K, 128, 32);
    if (rc) { printf("tmap A failed\n"); return 1; }
    rc = B_MN ? make_tmap_2d_f32(&maps.b[p], dB[p], K, N, N, 32, 32, mn.atom32) : make_tmap_2d_f32(&maps.b[p], dB[p], N, K, K, N, 32);
    if (rc) { printf("tmap B failed\n"); return 1; }
  }
  for (int p = npairs; p < 3; ++p) { maps.a[p] = maps.a[0]; maps.b[p] = maps.b[0]; }
  CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
  size_t smem = (size_t)3 * (K / 32) * (128 * 128 + N * 128) + 1024;
  CK(cudaFuncSetAttribute(k_probe<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_probe<A_MN, B_MN><<<1, 128, smem>>>(maps, npairs, N, K, dD, mn);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-28s N=%3d K=%3d mode=%d : KERNEL FAILED: %s\n", name, N, K, mode, cudaGetErrorString(e)); return 1; }
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double e_exact = 0, e_trunc = 0, e_rna = 0, e_fp32 = 0, ref_max = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s_exact = 0, s_tr = 0, s_rn = 0; float s32 = 0.f;
      for (int k = 0; k < K; ++k) {
        float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        s_exact += (double)a * b;
        s_tr += (double)tf32_trunc(a) * tf32_trunc(b);
        s_rn += (double)tf32_rna(a) * tf32_rna(b);
        s32 = fmaf(a, b, s32);
      }
      double d = D[(size_t)m * N + n];
      e_exact = fmax(e_exact, fabs(d - s_exact)); e_trunc = fmax(e_trunc, fabs(d - s_tr)); e_rna = fmax(e_rna, fabs(d - s_rn));
      e_fp32 = fmax(e_fp32, fabs((double)s32 - s_exact)); ref_max = fmax(ref_max, fabs(s_exact));
    }
  printf("[mn: layout %u lbo %u sbo %u kadv %u atom32 %d] ", mn.layout, mn.lbo, mn.sbo, mn.kadv, mn.atom32);
  printf("%-28s N=%3d K=%3d mode=%s : |D-exact| %.3e  |D-trunc model| %.3e  |D-rna model| %.3e  (fp32 fma chain vs exact %.3e, max|D| %.2f)\n",
         name, N, K, mode == 0 ? "1xTF32 raw" : "3xTF32    ", e_exact, e_trunc, e_rna, e_fp32, ref_max);
  for (int p = 0; p < npairs; ++p) { cudaFree(dA[p]); cudaFree(dB[p]); }
  cudaFree(dD);
  return 0;
}

int main() {
  int bad = 0;
  bad += run_case<0, 0>("A K-major,  B K-major", 128, 64, 0);
  // MN-major tf32: candidates for the BASE32B layout
  MnCfg cands[] = {
    {1, 4096, 512, 1024, 1}, {1, 4096, 1024, 1024, 1}, {1, 512, 4096, 1024, 1}, {1, 4096, 512, 512, 1},
    {1, 4096, 512, 1024, 0}, {2, 4096, 1024, 1024, 1}, {1, 1024, 4096, 1024, 1}, {1, 4096, 256, 1024, 1},
  };
  for (auto& c : cands) {
    bad += run_case<0, 1>("A K-major,  B MN-major", 128, 64, 0, c);
    bad += run_case<1, 0>("A MN-major, B K-major", 64, 32, 0, c);
  }
  printf("probe done, failures=%d\n", bad);
  return bad ? 1 : 0;
}

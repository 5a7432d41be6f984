// capi.cu -- error plumbing, per-layer entry points and the host-buffer block
// entry points of libsldm_sage.so (see include/sldm_sage.h).
#include "common.cuh"
#include <string.h>
#include <vector>
#include <atomic>
#include <mutex>

namespace sldm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) return SLDM_ENODEVICE;
  return SLDM_ECUDA;
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    cudaGetLastError();
    return 148;  // B200; only used to size grids
  }
  cached = sms;
  return cached;
}

}  // namespace sldm

using namespace sldm;

extern "C" int sldm_abi_version(void) { return SLDM_ABI_VERSION; }
extern "C" const char* sldm_last_error(void) { return g_err; }
extern "C" int64_t sldm_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int sldm_device_sm_count(int* out_sms) {
  SLDM_REQUIRE(out_sms != nullptr, SLDM_EINVAL, "sldm_device_sm_count: NULL");
  int dev = 0, sms = 0;
  SLDM_CUDA(cudaGetDevice(&dev));
  SLDM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  *out_sms = sms;
  return SLDM_OK;
}

// ------------------------------------------------------------ one layer fwd --
extern "C" int64_t sldm_sage_layer_fwd_workspace_bytes(int64_t N, int64_t E, int32_t Fin, int32_t Fout) {
  if (N < 0 || E < 0 || Fin < 0 || Fout < 0) return -1;
  return sldm_segment_workspace_bytes(N, E, Fin) + project_forward_ws_bytes(N, Fin, Fout);
}

extern "C" int sldm_sage_layer_forward(const float* x, int64_t N, int32_t Fin, int32_t Fout,
                                       const int32_t* csr, int64_t E,
                                       const float* W_l, const float* b_l, const float* W_r,
                                       const float* ln_w, const float* ln_b,
                                       float eps, float slope,
                                       float* out, float* agg, float* xhat_out, float* rstd_out,
                                       void* workspace, int64_t workspace_bytes,
                                       sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_sage_layer_forward: negative N or E");
  SLDM_REQUIRE(Fin >= 1 && Fout >= 1, SLDM_ESHAPE, "sldm_sage_layer_forward: Fin=%d Fout=%d", Fin, Fout);
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(x && csr && W_l && b_l && W_r && ln_w && ln_b && out && agg, SLDM_EINVAL,
               "sldm_sage_layer_forward: NULL pointer");
  const int64_t seg_ws = sldm_segment_workspace_bytes(N, E, Fin);
  const int64_t need = seg_ws + project_forward_ws_bytes(N, Fin, Fout);
  SLDM_REQUIRE(workspace != nullptr && workspace_bytes >= need, SLDM_EWORKSPACE,
               "sldm_sage_layer_forward: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  int rc = sldm_segment_reduce(x, N, Fin, csr, E, /*transpose=*/0, /*mean=*/1, nullptr, agg,
                               workspace, seg_ws, stream);
  if (rc) return rc;
  return project_forward_launch(agg, x, N, Fin, Fout, W_l, b_l, W_r, ln_w, ln_b, eps, slope,
                                out, xhat_out, rstd_out, static_cast<char*>(workspace) + seg_ws,
                                workspace_bytes - seg_ws, static_cast<cudaStream_t>(stream));
}

// ------------------------------------------------------------ one layer bwd --
extern "C" int64_t sldm_sage_layer_bwd_workspace_bytes(int64_t N, int64_t E, int32_t Fin, int32_t Fout) {
  if (N < 0 || E < 0 || Fin < 0 || Fout < 0) return -1;
  return sldm_segment_workspace_bytes(N, E, Fin) + layer_backward_ws_bytes(N, Fin, Fout);
}

extern "C" int sldm_sage_layer_backward(const float* dout, const float* x, const float* agg,
                                        const float* xhat, const float* rstd,
                                        int64_t N, int32_t Fin, int32_t Fout,
                                        const int32_t* csr, int64_t E,
                                        const float* W_l, const float* W_r,
                                        const float* ln_w, const float* ln_b, float slope,
                                        float* dx, float* dW_l, float* db_l, float* dW_r,
                                        float* dln_w, float* dln_b,
                                        float* dz, float* dagg, float* dxroot,
                                        void* workspace, int64_t workspace_bytes,
                                        sldm_stream_t stream) {
  return sldm_sage_layer_backward_stages(dout, x, agg, xhat, rstd, N, Fin, Fout, csr, E, W_l, W_r, ln_w, ln_b, slope,
                                         dx, dW_l, db_l, dW_r, dln_w, dln_b, dz, dagg, dxroot, workspace,
                                         workspace_bytes, stream, SLDM_BWD_STAGE_ALL);
}

static int layer_backward_impl(const float* dout, const float* x, const float* agg,
                               const float* xhat, const float* rstd,
                               int64_t N, int32_t Fin, int32_t Fout,
                               const int32_t* csr, int64_t E,
                               const float* W_l, const float* W_r,
                               const float* ln_w, const float* ln_b, float slope,
                               float* dx, float* dW_l, float* db_l, float* dW_r,
                               float* dln_w, float* dln_b,
                               float* dz, float* dagg, float* dxroot,
                               void* workspace, int64_t workspace_bytes,
                               sldm_stream_t stream, int32_t stages, bool bf16_feats) {
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_sage_layer_backward: negative N or E");
  SLDM_REQUIRE(Fin >= 1 && Fout >= 1, SLDM_ESHAPE, "sldm_sage_layer_backward: Fin=%d Fout=%d", Fin, Fout);
  SLDM_REQUIRE(dW_l && db_l && dW_r && dln_w && dln_b, SLDM_EINVAL, "sldm_sage_layer_backward: NULL gradient output");
  const bool need_dx = dx != nullptr;
  if (N > 0) {
    SLDM_REQUIRE(dout && x && agg && xhat && rstd && csr && W_l && W_r && ln_w && ln_b && dz, SLDM_EINVAL,
                 "sldm_sage_layer_backward: NULL pointer");
    SLDM_REQUIRE(!need_dx || (dagg && dxroot), SLDM_EINVAL, "sldm_sage_layer_backward: dx needs dagg and dxroot scratch");
  }
  const int64_t seg_ws = sldm_segment_workspace_bytes(N, E, Fin);
  const int64_t need = seg_ws + layer_backward_ws_bytes(N, Fin, Fout);
  SLDM_REQUIRE(N == 0 || (workspace != nullptr && workspace_bytes >= need), SLDM_EWORKSPACE,
               "sldm_sage_layer_backward: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int32_t* rowptr_dst = nullptr;
  if (N > 0) rowptr_dst = csr + csr_layout(N, E).off[SLDM_CSR_ROWPTR_DST];
  int rc = layer_backward_launch(dout, x, agg, xhat, rstd, N, Fin, Fout, rowptr_dst, W_l, W_r, ln_w, ln_b, slope,
                                 need_dx, dW_l, db_l, dW_r, dln_w, dln_b, dz, dagg, dxroot,
                                 static_cast<char*>(workspace) + seg_ws, workspace_bytes - seg_ws, s, stages, bf16_feats);
  if (rc || !need_dx || N == 0 || !(stages & SLDM_BWD_STAGE_GATHER)) return rc;
  // dx[j] = dxroot[j] + sum_{e: src[e]=j} dagg[dst[e]]   (dagg already divided by the count)
  return sldm_segment_reduce(dagg, N, Fin, csr, E, /*transpose=*/1, /*mean=*/0, dxroot, dx,
                             workspace, seg_ws, stream);
}

extern "C" int sldm_sage_layer_backward_stages(const float* dout, const float* x, const float* agg,
                                               const float* xhat, const float* rstd,
                                               int64_t N, int32_t Fin, int32_t Fout,
                                               const int32_t* csr, int64_t E,
                                               const float* W_l, const float* W_r,
                                               const float* ln_w, const float* ln_b, float slope,
                                               float* dx, float* dW_l, float* db_l, float* dW_r,
                                               float* dln_w, float* dln_b,
                                               float* dz, float* dagg, float* dxroot,
                                               void* workspace, int64_t workspace_bytes,
                                               sldm_stream_t stream, int32_t stages) {
  return layer_backward_impl(dout, x, agg, xhat, rstd, N, Fin, Fout, csr, E, W_l, W_r, ln_w, ln_b, slope, dx, dW_l, db_l,
                             dW_r, dln_w, dln_b, dz, dagg, dxroot, workspace, workspace_bytes, stream, stages, false);
}

// ------------------------------------------------------ bf16 feature storage --
extern "C" int sldm_sage_bf16_supported(int32_t Fin, int32_t Fout) {
  return (Fin % 64 == 0 && Fin >= 64 && Fin <= 128 && Fout % 32 == 0 && Fout >= 32 && Fout <= 128) ? 1 : 0;
}

extern "C" int sldm_sage_layer_forward_bf16(const void* x, int64_t N, int32_t Fin, int32_t Fout,
                                            const int32_t* csr, int64_t E,
                                            const float* W_l, const float* b_l, const float* W_r,
                                            const float* ln_w, const float* ln_b, float eps, float slope,
                                            void* out, void* agg, float* xhat_out, float* rstd_out,
                                            void* workspace, int64_t workspace_bytes, sldm_stream_t stream) {
  SLDM_REQUIRE(N >= 0 && E >= 0, SLDM_EINVAL, "sldm_sage_layer_forward_bf16: negative N or E");
  SLDM_REQUIRE(sldm_sage_bf16_supported(Fin, Fout), SLDM_EUNSUPPORTED,
               "sldm_sage_layer_forward_bf16: Fin=%d Fout=%d (needs Fin in {64,128}, Fout %% 32 == 0, Fout <= 128)", Fin, Fout);
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(x && csr && W_l && b_l && W_r && ln_w && ln_b && out && agg, SLDM_EINVAL,
               "sldm_sage_layer_forward_bf16: NULL pointer");
  SLDM_REQUIRE(project_forward_bf16_eligible(N, Fin, Fout, agg, x, out, xhat_out), SLDM_EUNSUPPORTED,
               "sldm_sage_layer_forward_bf16: buffers must be 16-byte aligned (and SLDM_DISABLE_TC unset)");
  const int64_t seg_ws = sldm_segment_workspace_bytes(N, E, Fin);
  const int64_t need = seg_ws + project_forward_ws_bytes(N, Fin, Fout);
  SLDM_REQUIRE(workspace != nullptr && workspace_bytes >= need, SLDM_EWORKSPACE,
               "sldm_sage_layer_forward_bf16: workspace %lld < %lld bytes", (long long)workspace_bytes, (long long)need);
  int rc = sldm_segment_mean_bf16(x, N, Fin, csr, E, agg, workspace, seg_ws, stream);
  if (rc) return rc;
  return project_forward_bf16_launch(agg, x, N, Fin, Fout, W_l, b_l, W_r, ln_w, ln_b, eps, slope, out, xhat_out,
                                     rstd_out, static_cast<char*>(workspace) + seg_ws, workspace_bytes - seg_ws,
                                     static_cast<cudaStream_t>(stream));
}

extern "C" int sldm_sage_project_forward_bf16(const void* agg, const void* x, int64_t N, int32_t Fin, int32_t Fout,
                                              const float* W_l, const float* b_l, const float* W_r,
                                              const float* ln_w, const float* ln_b, float eps, float slope,
                                              void* out, float* xhat_out, float* rstd_out,
                                              void* workspace, int64_t workspace_bytes, sldm_stream_t stream) {
  SLDM_REQUIRE(sldm_sage_bf16_supported(Fin, Fout), SLDM_EUNSUPPORTED, "sldm_sage_project_forward_bf16: Fin=%d Fout=%d", Fin, Fout);
  if (N == 0) return SLDM_OK;
  SLDM_REQUIRE(agg && x && W_l && b_l && W_r && ln_w && ln_b && out, SLDM_EINVAL, "sldm_sage_project_forward_bf16: NULL pointer");
  SLDM_REQUIRE(project_forward_bf16_eligible(N, Fin, Fout, agg, x, out, xhat_out), SLDM_EUNSUPPORTED,
               "sldm_sage_project_forward_bf16: buffers must be 16-byte aligned (and SLDM_DISABLE_TC unset)");
  return project_forward_bf16_launch(agg, x, N, Fin, Fout, W_l, b_l, W_r, ln_w, ln_b, eps, slope, out, xhat_out, rstd_out,
                                     workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int sldm_sage_layer_backward_bf16(const float* dout, const void* x, const void* agg,
                                             const float* xhat, const float* rstd,
                                             int64_t N, int32_t Fin, int32_t Fout,
                                             const int32_t* csr, int64_t E,
                                             const float* W_l, const float* W_r,
                                             const float* ln_w, const float* ln_b, float slope,
                                             float* dx, float* dW_l, float* db_l, float* dW_r,
                                             float* dln_w, float* dln_b,
                                             float* dz, float* dagg, float* dxroot,
                                             void* workspace, int64_t workspace_bytes,
                                             sldm_stream_t stream, int32_t stages) {
  SLDM_REQUIRE(sldm_sage_bf16_supported(Fin, Fout), SLDM_EUNSUPPORTED,
               "sldm_sage_layer_backward_bf16: Fin=%d Fout=%d (needs Fin in {64,128}, Fout %% 32 == 0, Fout <= 128)", Fin, Fout);
  return layer_backward_impl(dout, static_cast<const float*>(x), static_cast<const float*>(agg), xhat, rstd, N, Fin,
                             Fout, csr, E, W_l, W_r, ln_w, ln_b, slope, dx, dW_l, db_l, dW_r, dln_w, dln_b, dz, dagg,
                             dxroot, workspace, workspace_bytes, stream, stages, true);
}

// -------------------------------------------------- host buffers in and out --
namespace {
// Device memory of the host-buffer entry points comes from a small caching pool (per process, all devices): a block is
// handed back on destruction and reused by the next call of at least that size on the same device, so a caller that
// loops over sldm_sage_block_*_host pays cudaMalloc / cudaFree once, not ~25 times per call.  Blocks are only reused
// after the call that owned them has synchronised its stream.  SLDM_HOST_POOL_MB caps what is kept (default 4096).
struct DevPool {
  struct Block { void* p; size_t bytes; int dev; };
  std::mutex mu;
  std::vector<Block> free_blocks;
  size_t kept = 0;
  static size_t cap() {
    static size_t c = [] { const char* e = getenv("SLDM_HOST_POOL_MB"); return (size_t)(e ? atoll(e) : 4096) << 20; }();
    return c;
  }
  void* take(size_t bytes, int dev, size_t* real) {
    std::lock_guard<std::mutex> g(mu);
    int best = -1;
    for (int i = 0; i < (int)free_blocks.size(); ++i)
      if (free_blocks[i].dev == dev && free_blocks[i].bytes >= bytes && free_blocks[i].bytes <= 2 * bytes + (1 << 20) &&
          (best < 0 || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
    if (best < 0) return nullptr;
    void* p = free_blocks[best].p;
    *real = free_blocks[best].bytes;
    kept -= free_blocks[best].bytes;
    free_blocks.erase(free_blocks.begin() + best);
    return p;
  }
  void give(void* p, size_t bytes, int dev) {
    std::lock_guard<std::mutex> g(mu);
    if (kept + bytes > cap()) { cudaFree(p); return; }
    free_blocks.push_back({p, bytes, dev});
    kept += bytes;
  }
};
DevPool& pool() { static DevPool* p = new DevPool; return *p; }   // leaked on purpose: no CUDA calls at exit

thread_local cudaStream_t g_host_call_stream = nullptr;   // stream of the _host call running on this thread

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int dev = 0;
  cudaStream_t s = nullptr;
  int alloc(int64_t n) {
    if (n <= 0) n = 256;
    bytes = (size_t)((n + 255) / 256 * 256);
    s = g_host_call_stream;
    cudaGetDevice(&dev);
    size_t real = 0;
    p = pool().take(bytes, dev, &real);
    if (p) { bytes = real; return SLDM_OK; }
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { p = nullptr; return cuda_fail(e, "cudaMalloc"); }
    return SLDM_OK;
  }
  // (a block goes back only when nothing on the call's stream can still touch it: a no-op after the final sync of a
  //  successful call, a real wait on the error paths that return early)
  ~DevBuf() { if (p) { if (s) cudaStreamSynchronize(s); pool().give(p, bytes, dev); } }
  template <typename T> T* as() { return static_cast<T*>(p); }
};
}  // namespace

static int block_host_impl(const float* x_h, const int64_t* ei_h, int64_t N, int64_t E,
                           const int32_t* hdims, int32_t L, const float* const* params_h,
                           float eps, float slope, const float* dout_h,
                           float* out_h, float* dx_h, float* const* grads_h, bool train) {
  SLDM_REQUIRE(N >= 0 && E >= 0 && L >= 0, SLDM_EINVAL, "sage_block_host: negative N/E/L");
  SLDM_REQUIRE(hdims != nullptr, SLDM_EINVAL, "sage_block_host: hdims is NULL");
  for (int l = 0; l <= L; ++l) SLDM_REQUIRE(hdims[l] >= 1, SLDM_ESHAPE, "sage_block_host: hdims[%d]=%d", l, hdims[l]);
  SLDM_REQUIRE(N == 0 || (x_h && out_h), SLDM_EINVAL, "sage_block_host: NULL x/out");
  SLDM_REQUIRE(E == 0 || ei_h, SLDM_EINVAL, "sage_block_host: NULL edge_index");
  SLDM_REQUIRE(L == 0 || params_h, SLDM_EINVAL, "sage_block_host: NULL params");
  SLDM_REQUIRE(!train || (dout_h || N == 0), SLDM_EINVAL, "sage_block_host: NULL dout");
  SLDM_REQUIRE(!train || L == 0 || grads_h, SLDM_EINVAL, "sage_block_host: NULL grads");
  int ndev = 0;
  {
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      cudaGetLastError();
      set_error("sage_block_host: no CUDA device (there is no CPU fallback)");
      return SLDM_ENODEVICE;
    }
  }
  if (L == 0) {  // len(hdims) == 1: identity (src/models/blocks/sageblock.py:8 builds zero layers)
    if (N > 0 && out_h != x_h) memcpy(out_h, x_h, (size_t)N * hdims[0] * 4);
    if (train && dx_h && N > 0) memcpy(dx_h, dout_h, (size_t)N * hdims[0] * 4);
    return SLDM_OK;
  }
  cudaStream_t s = nullptr;
  SLDM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  struct StreamGuard { cudaStream_t s; ~StreamGuard() { g_host_call_stream = nullptr; cudaStreamDestroy(s); } } guard{s};
  g_host_call_stream = s;

  int Fmax = 0;
  for (int l = 0; l <= L; ++l) Fmax = hdims[l] > Fmax ? hdims[l] : Fmax;
  CsrLayout CL = csr_layout(N, E);
  int rc;
  DevBuf d_ei, d_csr, d_ws;
  if ((rc = d_ei.alloc(2 * E * 8))) return rc;
  if ((rc = d_csr.alloc(CL.off[SLDM_CSR_TOTAL] * 4))) return rc;
  int64_t ws_bytes = sldm_csr_workspace_bytes(N, E);
  for (int l = 0; l < L; ++l) {
    int64_t a = sldm_sage_layer_fwd_workspace_bytes(N, E, hdims[l], hdims[l + 1]);
    int64_t b = train ? sldm_sage_layer_bwd_workspace_bytes(N, E, hdims[l], hdims[l + 1]) : 0;
    ws_bytes = a > ws_bytes ? a : ws_bytes;
    ws_bytes = b > ws_bytes ? b : ws_bytes;
  }
  if ((rc = d_ws.alloc(ws_bytes))) return rc;
  if (E > 0) SLDM_CUDA(cudaMemcpyAsync(d_ei.p, ei_h, (size_t)2 * E * 8, cudaMemcpyHostToDevice, s));
  if ((rc = sldm_csr_build(d_ei.as<int64_t>(), E, N, d_csr.as<int32_t>(), d_ws.p, ws_bytes, s))) return rc;

  // activations: act[0] = x, act[l+1] = output of layer l; saved tensors per layer when training
  std::vector<DevBuf> act(L + 1), agg(L), xhat(train ? L : 0), rstd(train ? L : 0), par(5 * L);
  for (int l = 0; l <= L; ++l) if ((rc = act[l].alloc(N * (int64_t)hdims[l] * 4))) return rc;
  if (N > 0) SLDM_CUDA(cudaMemcpyAsync(act[0].p, x_h, (size_t)N * hdims[0] * 4, cudaMemcpyHostToDevice, s));
  for (int l = 0; l < L; ++l) {
    const int Fin = hdims[l], Fout = hdims[l + 1];
    const int64_t sizes[5] = {(int64_t)Fout * Fin, Fout, (int64_t)Fout * Fin, Fout, Fout};
    for (int k = 0; k < 5; ++k) {
      SLDM_REQUIRE(params_h[5 * l + k] != nullptr, SLDM_EINVAL, "sage_block_host: params[%d] is NULL", 5 * l + k);
      if ((rc = par[5 * l + k].alloc(sizes[k] * 4))) return rc;
      SLDM_CUDA(cudaMemcpyAsync(par[5 * l + k].p, params_h[5 * l + k], (size_t)sizes[k] * 4, cudaMemcpyHostToDevice, s));
    }
    if ((rc = agg[l].alloc(N * (int64_t)Fin * 4))) return rc;
    if (train) {
      if ((rc = xhat[l].alloc(N * (int64_t)Fout * 4))) return rc;
      if ((rc = rstd[l].alloc(N * 4))) return rc;
    }
    rc = sldm_sage_layer_forward(act[l].as<float>(), N, Fin, Fout, d_csr.as<int32_t>(), E,
                                 par[5 * l].as<float>(), par[5 * l + 1].as<float>(), par[5 * l + 2].as<float>(),
                                 par[5 * l + 3].as<float>(), par[5 * l + 4].as<float>(), eps, slope,
                                 act[l + 1].as<float>(), agg[l].as<float>(),
                                 train ? xhat[l].as<float>() : nullptr, train ? rstd[l].as<float>() : nullptr,
                                 d_ws.p, ws_bytes, s);
    if (rc) return rc;
  }
  if (N > 0) SLDM_CUDA(cudaMemcpyAsync(out_h, act[L].p, (size_t)N * hdims[L] * 4, cudaMemcpyDeviceToHost, s));

  if (train) {
    DevBuf g_cur, g_next, dz, dagg, dxroot;
    if ((rc = g_cur.alloc(N * (int64_t)Fmax * 4))) return rc;
    if ((rc = g_next.alloc(N * (int64_t)Fmax * 4))) return rc;
    if ((rc = dz.alloc(N * (int64_t)Fmax * 4))) return rc;
    if ((rc = dagg.alloc(N * (int64_t)Fmax * 4))) return rc;
    if ((rc = dxroot.alloc(N * (int64_t)Fmax * 4))) return rc;
    if (N > 0) SLDM_CUDA(cudaMemcpyAsync(g_cur.p, dout_h, (size_t)N * hdims[L] * 4, cudaMemcpyHostToDevice, s));
    std::vector<DevBuf> gpar(5 * L);
    for (int l = L - 1; l >= 0; --l) {
      const int Fin = hdims[l], Fout = hdims[l + 1];
      const int64_t sizes[5] = {(int64_t)Fout * Fin, Fout, (int64_t)Fout * Fin, Fout, Fout};
      for (int k = 0; k < 5; ++k) if ((rc = gpar[5 * l + k].alloc(sizes[k] * 4))) return rc;
      const bool need_dx = (l > 0) || (dx_h != nullptr);
      rc = sldm_sage_layer_backward(g_cur.as<float>(), act[l].as<float>(), agg[l].as<float>(),
                                    xhat[l].as<float>(), rstd[l].as<float>(), N, Fin, Fout,
                                    d_csr.as<int32_t>(), E, par[5 * l].as<float>(), par[5 * l + 2].as<float>(),
                                    par[5 * l + 3].as<float>(), par[5 * l + 4].as<float>(), slope,
                                    need_dx ? g_next.as<float>() : nullptr,
                                    gpar[5 * l].as<float>(), gpar[5 * l + 1].as<float>(), gpar[5 * l + 2].as<float>(),
                                    gpar[5 * l + 3].as<float>(), gpar[5 * l + 4].as<float>(),
                                    dz.as<float>(), dagg.as<float>(), dxroot.as<float>(), d_ws.p, ws_bytes, s);
      if (rc) return rc;
      for (int k = 0; k < 5; ++k) {
        SLDM_REQUIRE(grads_h[5 * l + k] != nullptr, SLDM_EINVAL, "sage_block_host: grads[%d] is NULL", 5 * l + k);
        SLDM_CUDA(cudaMemcpyAsync(grads_h[5 * l + k], gpar[5 * l + k].p, (size_t)sizes[k] * 4, cudaMemcpyDeviceToHost, s));
      }
      std::swap(g_cur.p, g_next.p);
    }
    if (dx_h && N > 0) SLDM_CUDA(cudaMemcpyAsync(dx_h, g_cur.p, (size_t)N * hdims[0] * 4, cudaMemcpyDeviceToHost, s));
    SLDM_CUDA(cudaStreamSynchronize(s));
    return SLDM_OK;
  }
  SLDM_CUDA(cudaStreamSynchronize(s));
  return SLDM_OK;
}

extern "C" int sldm_sage_block_forward_host(const float* x_h, const int64_t* edge_index_h,
                                            int64_t N, int64_t E,
                                            const int32_t* hdims, int32_t L,
                                            const float* const* params_h,
                                            float eps, float slope, float* out_h) {
  return block_host_impl(x_h, edge_index_h, N, E, hdims, L, params_h, eps, slope, nullptr, out_h,
                         nullptr, nullptr, false);
}

extern "C" int sldm_sage_block_train_host(const float* x_h, const int64_t* edge_index_h,
                                          int64_t N, int64_t E,
                                          const int32_t* hdims, int32_t L,
                                          const float* const* params_h,
                                          float eps, float slope, const float* dout_h,
                                          float* out_h, float* dx_h, float* const* grads_h) {
  return block_host_impl(x_h, edge_index_h, N, E, hdims, L, params_h, eps, slope, dout_h, out_h,
                         dx_h, grads_h, true);
}

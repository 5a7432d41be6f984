// tc_host.cu -- host helpers of the tensor path: CUtensorMap encoding through the
// driver entry point (no link-time dependency on libcuda).
#include "common.cuh"
#include "tc_common.cuh"

namespace sldm {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || p == nullptr) {
    cudaGetLastError();
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_2d_f32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                     uint32_t box_rows, uint32_t box_cols, int swizzle_atom32) {
  EncodeTiledFn fn = encode_fn();
  SLDM_REQUIRE(fn != nullptr, SLDM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
  SLDM_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (ld_elems * 4) % 16 == 0, SLDM_EINVAL,
               "tensor map: base / row pitch must be 16-byte aligned");
  SLDM_REQUIRE(box_cols * 4 <= 128 && box_rows <= 256 && box_rows >= 1, SLDM_EINVAL, "tensor map: bad box");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_elems * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_atom32 == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                  : (swizzle_atom32 == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                  : (swizzle_atom32 == 3 ? CU_TENSOR_MAP_SWIZZLE_NONE
                  : (swizzle_atom32 == 4 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B))),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SLDM_REQUIRE(r == CUDA_SUCCESS, SLDM_ECUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return SLDM_OK;
}

}  // namespace sldm

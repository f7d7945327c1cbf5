// Host-side TMA tensor-map construction.  The driver entry point is fetched through the runtime
// (cudaGetDriverEntryPoint) so the library never links libcuda directly.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace fz {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled(std::string* err) {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    if (err) *err = std::string("cuTensorMapEncodeTiled unavailable: ") + cudaGetErrorString(e);
    return nullptr;
  }
  fn = reinterpret_cast<PFN_encodeTiled>(p);
  return fn;
}

// Row-major bf16 matrix [rows][cols] with leading dimension ld (elements); box = box_cols x box_rows,
// 128-byte swizzle (box_cols * 2 bytes must be 128).  Out-of-bounds elements read as zero.
inline bool make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                              uint32_t box_cols, uint32_t box_rows, std::string* err) {
  PFN_encodeTiled enc = get_encode_tiled(err);
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    if (err) *err = "bf16 operand must be 16-byte aligned with a leading dimension that is a multiple of 8";
    return false;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

// Row-major matrix [rows][cols] of `esz`-byte elements, any swizzle mode (box_cols * esz must not exceed the swizzle span).
inline bool make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dt, uint32_t esz, const void* base, uint64_t rows, uint64_t cols,
                         uint64_t ld, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swizzle, std::string* err) {
  PFN_encodeTiled enc = get_encode_tiled(err);
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * esz) & 15) != 0) {
    if (err) *err = "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch";
    return false;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esz};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

// Row-major fp32 matrix [rows][cols] (leading dimension ld elements), box = box_cols x box_rows with
// 128-byte swizzle (box_cols must be 32).  Used as the destination of TMA reduce-add stores.
inline bool make_tmap_f32_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                             uint32_t box_cols, uint32_t box_rows, std::string* err) {
  PFN_encodeTiled enc = get_encode_tiled(err);
  if (!enc) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 4) & 15) != 0) {
    if (err) *err = "fp32 TMA target must be 16-byte aligned with a leading dimension that is a multiple of 4";
    return false;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(fp32) failed with CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

}  // namespace fz

// fa_tma_host.h -- host-side construction of TMA tensor maps without linking libcuda.
//
// cuTensorMapEncodeTiled is resolved through cudaGetDriverEntryPoint so the library only
// depends on libcudart.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace fa {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      fprintf(stderr, "fa_b200: cannot resolve cuTensorMapEncodeTiled (%s)\n", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

// Row-major [d2][d1][d0] tensor (d0 contiguous) of `esize`-byte elements; box = [1][box1][box0],
// out-of-bounds elements read as zero / are not written.  swizzle_mode: 0 none, 1 = 128 B (16-B
// chunks, box0 * esize must be 128), 2 = 128 B with 32-B atomicity (required by UMMA for MN-major
// 32-bit operands), 3 = 64 B (box0 * esize must be 64).  Returns 0 on success.
inline int make_tmap_3d(CUtensorMap* map, const void* base, int esize, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint32_t box0, uint32_t box1, int swizzle_mode = 1) {
  PFN_tmapEncodeTiled enc = get_tmap_encode();
  if (!enc) return -1;
  CUtensorMapDataType dt = esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {d0 * (uint64_t)esize, d0 * d1 * (uint64_t)esize};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_mode == 1   ? CU_TENSOR_MAP_SWIZZLE_128B
                   : swizzle_mode == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                   : swizzle_mode == 3 ? CU_TENSOR_MAP_SWIZZLE_64B
                                       : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "fa_b200: cuTensorMapEncodeTiled failed (%d) dims=[%llu,%llu,%llu] box=[%u,%u]\n", (int)r,
            (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1);
    return -2;
  }
  return 0;
}

}  // namespace fa

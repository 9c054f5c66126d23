// Host-side helpers shared by the C-ABI entry points.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/more4d_b200.h"

namespace m4d {

typedef __nv_bfloat16 bf16;

inline int cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return M4D_OK;
  fprintf(stderr, "[more4d_b200] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
  return M4D_ERR_CUDA;
}

#define M4D_CHECK_LAUNCH(what)                                  \
  do {                                                          \
    int _rc = ::m4d::cuda_ok(cudaGetLastError(), what);         \
    if (_rc != M4D_OK) return _rc;                              \
  } while (0)

#define M4D_REQUIRE(cond, code) \
  do {                          \
    if (!(cond)) return (code); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// cuTensorMapEncodeTiled is resolved through the runtime so the library has no link-time
// dependency on libcuda.so (it must load on a CPU-only box for the symbol tests).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// bf16 tensor map with 128-byte swizzle; dims/strides innermost first, strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous).  Out-of-bounds elements read as zero.
inline int make_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return M4D_ERR_NO_DEVICE;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      if (strides_bytes[i - 1] % 16 != 0) return M4D_ERR_ALIGN;
      gstr[i - 1] = strides_bytes[i - 1];
    }
  }
  if (!aligned16(base)) return M4D_ERR_ALIGN;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gdim, gstr,
                  bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[more4d_b200] cuTensorMapEncodeTiled failed: %d\n", static_cast<int>(r));
    return M4D_ERR_CUDA;
  }
  return M4D_OK;
}

inline int sm_count() {
  static int n = 0;
  if (n) return n;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;
  return n;
}

#ifdef M4D_DEV
// Development builds only (`python more4d_b200/build.py -DM4D_DEV`): kernel-variant selection for
// A/B measurements through m4d_dev_set_flags.  The product library has no global mutable state.
extern int g_dev_flags;
#endif

// 2-CTA (cta_group::2) GEMM, gemm2.cu
int gemm2_dispatch(const void* a, long long lda, const void* w, long long ldw, const void* bias, void* out,
                   long long ldo, int M, int N, int K, int epilogue, const float* residual, long long ldr,
                   const float* gate, long long gate_batch_stride, int rows_per_batch,
                   cudaStream_t stream);

}  // namespace m4d

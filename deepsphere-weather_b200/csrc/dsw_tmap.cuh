// Tensor-map (TMA descriptor) helper shared by the kernels that stage fp32 tiles with
// cp.async.bulk.tensor.  The driver entry point is fetched through the runtime, so the library has no
// link-time dependency on libcuda.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dsw {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// fp32 tensor of `rank` dims (innermost first), byte strides for dims 1..rank-1, un-swizzled box,
// zero fill out of bounds.  Returns false if the driver refuses (caller falls back).
inline bool encode_f32_map(CUtensorMap* out, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  {
    // cuTensorMapEncodeTiled is a driver call: it needs the device's primary context current on THIS thread.  A thread
    // that has only used the runtime API lazily (PyTorch's autograd worker on its first backward) has none bound yet.
    static thread_local int bound_device = -1;
    int dev = -1;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != bound_device) {
      // cudaSetDevice (CUDA >= 12) initialises the primary context and makes it current; unlike cudaFree(0) it touches no
      // stream, so it is legal while another thread's stream capture is in progress (CUDA graphs of a training step)
      (void)cudaSetDevice(dev);
      bound_device = dev;
    }
  }
  cuuint64_t d[5], s[4];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) d[i] = dims[i], b[i] = box[i], e[i] = 1;
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, b, e,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc == CUDA_ERROR_INVALID_CONTEXT || rc == CUDA_ERROR_NOT_INITIALIZED) {
    // A thread that has only used the runtime API lazily (PyTorch's autograd worker on its first backward) has no
    // driver context bound yet: bind the device's primary context and encode again.
    (void)cudaSetDevice([] { int d = 0; (void)cudaGetDevice(&d); return d; }());
    rc = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, s, b, e,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  return rc == CUDA_SUCCESS;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

}  // namespace dsw

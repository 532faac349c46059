// Microbenchmark: cost per TMA operation on one SM (all SMs busy): cp.async.bulk (UBLKCP) vs cp.async.bulk.tensor
// (UTMALDG) boxes of h rows x 256 B from an L2-resident [rows][64] fp32 array; `nthreads` threads issue one op each per
// round onto ONE shared mbarrier (as the chain kernel's issuer warp does), the round ends when all bytes have landed.
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool TENSOR>
__global__ void __launch_bounds__(128, 1) k_ops(const float* __restrict__ src, const __grid_constant__ CUtensorMap map, int rows_total,
                                                 int h, int nthreads, int rounds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  const int tid = threadIdx.x;
  if (tid == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (tid >= 32) return;
  const int per_cta = rows_total / gridDim.x;
  int row = blockIdx.x * per_cta + tid * h * 3;
  uint32_t ph = 0;
  const uint32_t bytes = (uint32_t)h * 256u;
  for (int r = 0; r < rounds; ++r) {
    if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes * nthreads) : "memory");
    __syncwarp();
    if (tid < nthreads) {
      if (TENSOR) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         s32(buf + (size_t)tid * bytes)),
                     "l"(&map), "r"(0), "r"(row), "r"(s32(bar))
                     : "memory");
      } else {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(buf + (size_t)tid * bytes)),
                     "l"(src + (size_t)row * 64), "r"(bytes), "r"(s32(bar))
                     : "memory");
      }
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                     s32(bar)),
                 "r"(ph)
                 : "memory");
    ph ^= 1;
    row += 97;
    if (row + 32 * h * 3 + h >= (blockIdx.x + 1) * per_cta) row = blockIdx.x * per_cta + tid * h * 3;
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows_total = 148 * 1200;  // 45 MB: L2-resident
  float* src;
  cudaMalloc(&src, (size_t)rows_total * 256);
  cudaMemset(src, 0, (size_t)rows_total * 256);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fp;
  cudaFuncSetAttribute(k_ops<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_ops<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  for (int h : {1, 2, 4, 8, 16})
    for (int nthreads : {1, 8, 32})
      for (int tensor = 0; tensor < 2; ++tensor) {
        CUtensorMap map;
        cuuint64_t dims[2] = {64, (cuuint64_t)rows_total}, strides[1] = {256};
        cuuint32_t box[2] = {64, (cuuint32_t)h}, es[2] = {1, 1};
        enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, src, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        const int rounds = 4000;
        float ms = 0;
        for (int rep = 0; rep < 2; ++rep) {
          cudaEventRecord(e0);
          if (tensor)
            k_ops<true><<<148, 128, 1024 + 32 * h * 256>>>(src, map, rows_total, h, nthreads, rounds);
          else
            k_ops<false><<<148, 128, 1024 + 32 * h * 256>>>(src, map, rows_total, h, nthreads, rounds);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&ms, e0, e1);
        }
        const double cyc_round = ms * 1e-3 * 1.9e9 / rounds;
        printf("%s h=%2d rows (%5d B)  ops/round %2d  %7.0f cycles/round  %6.1f cycles/op  %6.1f B/clk/SM  err=%d\n", tensor ? "tensor" : "bulk  ", h,
               h * 256, nthreads, cyc_round, cyc_round / nthreads, (double)nthreads * h * 256 / cyc_round, (int)cudaGetLastError());
      }
  return 0;
}

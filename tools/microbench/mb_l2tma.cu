// Microbenchmark: L2 -> shared memory bandwidth of cp.async.bulk (TMA) on every SM at once, as a function of the copy
// size and of the bytes kept in flight per SM.  The source region is small enough to stay L2-resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_l2tma mb_l2tma.cu ; ./mb_l2tma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One issuing thread per "lane group"; `inflight` copies of `chunk` bytes are kept outstanding per CTA.
__global__ void __launch_bounds__(128, 1) k_bulk(const uint8_t* __restrict__ src, size_t region, int chunk, int inflight, int iters,
                                                  unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* buf = smem + 1024;
  const int tid = threadIdx.x;
  if (tid < inflight) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bars + tid)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  const size_t per_cta = (region / gridDim.x) & ~(size_t)65535;
  const uint8_t* base = src + (size_t)blockIdx.x * per_cta;
  long long t0 = clock64();
  if (tid < inflight) {
    uint32_t ph = 0;
    size_t off = (size_t)tid * chunk;
    for (int i = 0; i < iters; ++i) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bars + tid)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       s32(buf + (size_t)tid * chunk)),
                   "l"(base + off), "r"(chunk), "r"(s32(bars + tid))
                   : "memory");
      asm volatile(
          "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
              s32(bars + tid)),
          "r"(ph)
          : "memory");
      ph ^= 1;
      off += (size_t)inflight * chunk;
      if (off + chunk > per_cta) off = (size_t)tid * chunk;
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (tid == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

int main() {
  const size_t region = (size_t)48 << 20;  // L2-resident
  uint8_t* src;
  cudaMalloc(&src, region);
  cudaMemset(src, 1, region);
  unsigned long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  const int chunks[] = {256, 512, 1024, 4096, 16384, 32768};
  for (int chunk : chunks)
    for (int inflight : {1, 2, 4, 8, 16, 32, 64}) {
      if ((size_t)chunk * inflight > 190 * 1024) continue;
      const int iters = (int)(((size_t)24 << 20) / ((size_t)chunk * inflight)) + 4;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k_bulk<<<148, 128, 1024 + chunk * inflight>>>(src, region, chunk, inflight, iters, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double bytes = 148.0 * chunk * inflight * iters;
      printf("chunk %6d B  inflight %3d (%7d B/SM)  %8.1f us  %7.2f TB/s  %6.1f B/clk/SM(@1.9GHz)  err=%d\n", chunk, inflight,
             chunk * inflight, ms * 1e3, bytes / ms / 1e9, bytes / 148 / (ms * 1e-3 * 1.9e9), (int)cudaGetLastError());
    }
  return 0;
}

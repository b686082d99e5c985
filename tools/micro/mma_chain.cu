// Microbenchmark: latency of back-to-back tcgen05.mma kind::tf32 (M=128, N, K=8) into 1/2/4 accumulators.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t sdesc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__global__ void k(int N, int nacc, int nmma, int a_from_tmem, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar; __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((float*)sm)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;"); }
  __syncthreads();
  if (threadIdx.x < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&slot))); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
  asm volatile("fence.proxy.async.shared::cta;"); asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  uint32_t tm = slot;
  uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
  if (threadIdx.x == 0) {
    uint64_t da = sdesc(s32(sm), 2048, 128), db = sdesc(s32(sm) + 32768, 1024, 128);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        uint32_t d = tm + 256 + (i % nacc) * 64, en = 1;
        if (a_from_tmem)
          asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;}" ::"r"(d), "r"(tm + (i % 8) * 8), "l"(db), "r"(idesc), "r"(en));
        else
          asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(en));
      }
      long long t1 = clock64();
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)));
      uint32_t ok = 0;
      while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(&bar)), "r"(rep & 1));
      long long t2 = clock64();
      out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[6];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  int Ns[] = {64, 128, 256};
  for (int at = 0; at < 2; ++at) for (int N : Ns) for (int nacc : {1, 2, 4}) {
    if (nacc * N > 256) continue;
    k<<<1, 128, 65536>>>(N, nacc, 48, at, d); cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    printf("A_from_%s N=%3d nacc=%d: issue %lld cyc, total %lld cyc for 48 MMAs -> %.1f cyc/MMA (%s)\n", at ? "tmem" : "smem", N, nacc, h[4], h[5], h[5] / 48.0, cudaGetErrorString(e));
  }
  return 0;
}

// How fast does one SM's TMA unit retire 1-D bulk copies (cp.async.bulk) of a given size?
// One CTA per SM, one producer thread, ring of NSLOT smem slots; consumers do nothing.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NLANES>
__global__ void k_tma(const char *src, size_t src_bytes, int copy_bytes, int copies_per_batch, int nbatch, long long *cycles) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar[2];
  const int lane = threadIdx.x;
  if (lane == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(s32(bar + i)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  const size_t batch_bytes = (size_t)copy_bytes * copies_per_batch;
  const char *base = src + ((size_t)blockIdx.x * 7919 * 4096) % (src_bytes - batch_bytes * (nbatch + 1));
  long long t0 = clock64();
  for (int b = 0; b < nbatch; ++b) {
    const int s = b & 1;
    if (b >= 2) {  // wait for the batch that used this slot
      uint32_t parity = ((b - 2) >> 1) & 1;
      asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar + s)), "r"(parity) : "memory");
    }
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(s32(bar + s)), "r"((uint32_t)batch_bytes) : "memory");
    __syncwarp();
    for (int c = lane; c < copies_per_batch; c += NLANES) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(s32(smem + s * batch_bytes + (size_t)c * copy_bytes)),
                   "l"(base + (size_t)b * batch_bytes + (size_t)c * copy_bytes), "r"((uint32_t)copy_bytes), "r"(s32(bar + s)) : "memory");
    }
  }
  for (int b = max(0, nbatch - 2); b < nbatch; ++b) {
    uint32_t parity = (b >> 1) & 1;
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(s32(bar + (b & 1))), "r"(parity) : "memory");
  }
  long long t1 = clock64();
  if (lane == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  size_t bytes = (size_t)2 << 30;
  char *src; CK(cudaMalloc(&src, bytes)); CK(cudaMemset(src, 1, bytes));
  long long *cyc; CK(cudaMallocManaged(&cyc, sizeof(long long) * sms));
  CK(cudaFuncSetAttribute(k_tma<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int batch = 32 * 1024;  // bytes per batch (per slot), 2 slots
  for (int grid : {1, sms}) {
    for (int cb : {256, 512, 1024, 2048, 4096, 8192, 16384, 32768}) {
      int per = batch / cb, nb = 400;
      k_tma<32><<<grid, 32, 2 * batch>>>(src, bytes, cb, per, nb, cyc);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      k_tma<32><<<grid, 32, 2 * batch>>>(src, bytes, cb, per, nb, cyc);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      double c = 0; for (int i = 0; i < grid; ++i) c += cyc[i]; c /= grid;
      printf("{\"test\": \"tma_bulk_1d\", \"grid\": %d, \"copy_bytes\": %d, \"cycles_per_copy\": %.1f, \"bytes_per_cycle_per_sm\": %.2f, \"GBps_total\": %.1f}\n",
             grid, cb, c / ((double)nb * per), (double)batch * nb / c, (double)batch * nb * grid / ms * 1e-6);
    }
  }
  return 0;
}

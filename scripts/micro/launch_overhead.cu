// Microbenchmark: event-timed duration of (nearly) empty kernels in a stream, alternating shared-memory
// configurations -- does a 227 KB dynamic-smem, TMEM-allocating persistent kernel pay a fixed cost when it
// follows a kernel with a different carve-out?
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(736, 1) big_kernel(int tmem, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint32_t tmem_base;
  if (tmem && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) smem[0] = 1;
  __syncthreads();
  if (tmem && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[0] = smem[0];
}
__global__ void small_kernel(float* out) { if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[1] = 2.f; }

int main() {
  float* out; cudaMalloc(&out, 64);
  cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaEvent_t e[64]; for (auto& x : e) cudaEventCreate(&x);
  for (int variant = 0; variant < 4; ++variant) {
    const size_t smem = (variant & 1) ? 227 * 1024 : 16 * 1024;
    const int tmem = variant >> 1;
    for (int rep = 0; rep < 3; ++rep) {
      // pattern: small, big, small, big, big, big
      cudaEventRecord(e[0]); small_kernel<<<1184, 256>>>(out);
      cudaEventRecord(e[1]); big_kernel<<<148, 736, smem>>>(tmem, out);
      cudaEventRecord(e[2]); small_kernel<<<1184, 256>>>(out);
      cudaEventRecord(e[3]); big_kernel<<<148, 736, smem>>>(tmem, out);
      cudaEventRecord(e[4]); big_kernel<<<148, 736, smem>>>(tmem, out);
      cudaEventRecord(e[5]); big_kernel<<<148, 736, smem>>>(tmem, out);
      cudaEventRecord(e[6]);
      cudaDeviceSynchronize();
      float t[6];
      for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], e[i], e[i + 1]);
      if (rep == 2)
        printf("smem=%3zu KB tmem=%d: small %.1f us | big-after-small %.1f | small-after-big %.1f | big-after-small %.1f | big-after-big %.1f | big-after-big %.1f\n",
               smem / 1024, tmem, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
    }
  }
  return 0;
}

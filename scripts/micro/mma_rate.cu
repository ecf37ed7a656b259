// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128) for kind::tf32 (K=8) and kind::f16 (K=16),
// SWIZZLE_128B K-major operands in shared memory, accumulator in TMEM.  One CTA per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc(int fmt, int m, int n) {
  return (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int KIND>  // 0 = tf32, 1 = bf16, 2 = tf32 with the A operand in tensor memory
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int iters, int distinct, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base;
  if (warp == 0) {
    // warp-uniform issue loop (descriptors in uniform registers), lane 0 issues; 16 MMAs per iteration
    const uint32_t leader = (tid == 0) ? 1u : 0u;
    const uint64_t da0 = umma_desc(smem_u32(smem)), db0 = umma_desc(smem_u32(smem + 64 * 1024));
    const uint32_t id = idesc(KIND == 1 ? 1 : 2, 128, n);
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const uint64_t da = da0 + 2 * (u & 3) * (distinct > 1 ? 1 : 0);
        const uint64_t db = db0 + 2 * (u & 3) * (distinct > 1 ? 1 : 0);
        if (KIND == 2) {
          const uint32_t ta = tmem_d + 256u + 8u * (u & 3) * (distinct > 1 ? 1 : 0);
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                       ::"r"(tmem_d), "r"(ta), "l"(db), "r"(id), "r"(1u), "r"(leader) : "memory");
        } else if (KIND == 0) {
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(1u), "r"(leader) : "memory");
        } else {
          asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                       ::"r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(1u), "r"(leader) : "memory");
        }
      }
    }
    if (tid == 0)
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long t2 = clock64();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  const size_t smem = 192 * 1024;
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 2048;
  for (int kind = 0; kind < 3; ++kind)
    for (int grid : {1, 148})
      for (int n : {16, 64, 128, 144, 256})
        for (int distinct : {1, 4}) {
          for (int rep = 0; rep < 2; ++rep) {
            if (kind == 0) rate_kernel<0><<<grid, 128, smem>>>(n, iters, distinct, out);
            else if (kind == 1) rate_kernel<1><<<grid, 128, smem>>>(n, iters, distinct, out);
            else rate_kernel<2><<<grid, 128, smem>>>(n, iters, distinct, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          }
          printf("kind=%s grid=%3d N=%3d distinct=%d: issue %.1f cyc/mma, complete %.1f cyc/mma\n", kind == 0 ? "tf32" : (kind == 1 ? "bf16" : "tf32-TS"),
                 grid, n, distinct, (double)out[0] / iters, (double)out[1] / iters);
        }
  return 0;
}

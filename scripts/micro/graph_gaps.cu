// Microbenchmark: cost of a kernel boundary inside a CUDA graph (chain of dependent, nearly empty kernels), with and
// without programmatic dependent launch (cudaLaunchAttributeProgrammaticStreamSerialization + griddepcontrol.wait).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__global__ void __launch_bounds__(736, 1) big_kernel(float* out, int pdl) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) smem[0] = 1;
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] += smem[0];
}
__global__ void small_kernel(float* out, int pdl) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0 && blockIdx.x == 0) out[1] += 2.f;
}

template <typename K>
static void launch(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, int pdl, float* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, out, pdl);
}

int main() {
  float* out; cudaMalloc(&out, 64); cudaMemset(out, 0, 64);
  cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int pattern = 0; pattern < 3; ++pattern) {       // 0: small only, 1: big only, 2: alternating
    for (int pdl = 0; pdl < 2; ++pdl) {
      const int n = 40;
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(s, cudaStreamCaptureModeGlobal);
      for (int i = 0; i < n; ++i) {
        const bool big = pattern == 1 || (pattern == 2 && (i & 1));
        if (big) launch(big_kernel, dim3(148), dim3(736), 200 * 1024, s, pdl, out);
        else launch(small_kernel, dim3(1184), dim3(256), 0, s, pdl, out);
      }
      cudaError_t e1 = cudaStreamEndCapture(s, &g);
      cudaError_t e2 = cudaGraphInstantiate(&ge, g, 0);
      for (int w = 0; w < 5; ++w) cudaGraphLaunch(ge, s);
      cudaStreamSynchronize(s);
      cudaEventRecord(a, s);
      for (int r = 0; r < 20; ++r) cudaGraphLaunch(ge, s);
      cudaEventRecord(b, s);
      cudaStreamSynchronize(s);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("pattern %d (%s) pdl=%d: %.2f us per kernel (graph of %d) capture=%d inst=%d err=%s\n", pattern,
             pattern == 0 ? "small" : pattern == 1 ? "big" : "alternating", pdl, ms * 1e3 / (20 * n), n, (int)e1, (int)e2,
             cudaGetErrorString(cudaGetLastError()));
      cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
  }
  return 0;
}

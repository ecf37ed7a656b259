// common.cuh -- shared host/device helpers for librgnn_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rgnn.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librgnn_b200 is written for sm_100a (B200) only"
#endif

namespace rgnn {

// ---- error handling ---------------------------------------------------------------
void set_last_cuda_error(cudaError_t err, const char* file, int line);
void count_launch(int n = 1);

#define RGNN_CUDA_CHECK(expr)                                  \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      ::rgnn::set_last_cuda_error(_e, __FILE__, __LINE__);     \
      return RGNN_ERR_CUDA;                                    \
    }                                                          \
  } while (0)

// Optional per-kernel timing (rgnn_profile_enable): a scope records CUDA events on the launch
// stream around the launches it encloses; rgnn_profile_collect sums them by name.
struct ProfileScope {
  ProfileScope(const char* name, cudaStream_t stream);
  ~ProfileScope();
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t start_;
  bool active_;
};
#define RGNN_PROFILE_CAT2(a, b) a##b
#define RGNN_PROFILE_CAT(a, b) RGNN_PROFILE_CAT2(a, b)
#define RGNN_PROFILE(name, stream) ::rgnn::ProfileScope RGNN_PROFILE_CAT(_prof_scope_, __LINE__)(name, stream)

// after a <<<>>> launch
#define RGNN_LAUNCH_CHECK()                                    \
  do {                                                         \
    ::rgnn::count_launch();                                    \
    cudaError_t _e = cudaGetLastError();                       \
    if (_e != cudaSuccess) {                                   \
      ::rgnn::set_last_cuda_error(_e, __FILE__, __LINE__);     \
      return RGNN_ERR_CUDA;                                    \
    }                                                          \
  } while (0)

#define RGNN_RETURN_IF_ERROR(expr)                             \
  do {                                                         \
    int _s = (expr);                                           \
    if (_s != RGNN_OK) return _s;                              \
  } while (0)

// ---- workspace bump allocator -----------------------------------------------------
constexpr size_t kAlign = 256;
inline size_t align_up(size_t v, size_t a = kAlign) { return (v + a - 1) / a * a; }

struct Arena {
  char* base;
  size_t size;
  size_t used;
  bool overflow;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n), used(0), overflow(false) {
    size_t mis = reinterpret_cast<uintptr_t>(base) % kAlign;
    if (mis) used = kAlign - mis;
  }
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (base == nullptr || used + bytes > size) {
      overflow = true;
      used += bytes;
      return nullptr;
    }
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

// Size-only arena: run the same take() sequence with a null base to learn the need.
struct SizeArena {
  size_t used = kAlign;  // slack for base misalignment
  template <typename T>
  T* take(size_t count) {
    used += align_up(count * sizeof(T));
    return nullptr;
  }
};

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
inline int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return 0;
  return dev;
}
constexpr int kMaxDevices = 64;
inline int sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device_index();
  int& c = cached[dev < kMaxDevices ? dev : 0];
  if (c == 0 || dev >= kMaxDevices) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    if (dev >= kMaxDevices) return v;
    c = v;
  }
  return c;
}

// One-time per-device opt-in to more than 48 KB of dynamic shared memory (cudaFuncSetAttribute applies to
// the current device only, so a process-wide flag would leave a second GPU unconfigured).
template <typename KernelT>
inline cudaError_t opt_in_dynamic_smem(KernelT kernel, bool (&done)[kMaxDevices], int bytes) {
  const int dev = current_device_index();
  if (dev < kMaxDevices && done[dev]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev < kMaxDevices) done[dev] = true;
  return e;
}

inline unsigned div_up(int64_t a, int64_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// ---- scans (scan.cu) ----------------------------------------------------------------
// Exclusive prefix sum of int32 `in[0..n)` into `out[0..n]` (out[n] = total).  `in` and
// `out` may alias when out == in is not needed; scratch must hold scan_scratch_ints(n).
size_t scan_scratch_ints(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t stream);
// 64-bit totals variant used for edge row pointers (n up to 2^31 rows, sums up to 2^62)
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, int64_t* scratch, cudaStream_t stream);

// ---- internal kernels shared between translation units ------------------------------
// y[m, n] = in(a)[m, k1 + k2] . w[n, k1 + k2]^T + bias (+ residual), where the input row is the
// concatenation [a1 | a2] (PyG's cat([x, m]) without materialising it) and in() optionally
// applies the previous layer's BatchNorm + ReLU to a1 on load:
//   a1' = relu?((a1 - mean[c]) * scale[c] + beta[c])
struct LinearArgs {
  const float* a1 = nullptr;   // [m, k1], row stride lda1
  int64_t lda1 = 0;
  int32_t k1 = 0;
  const int32_t* a1_rows = nullptr;  // optional gather: row r of the input is a1[a1_rows[r]]
  const int32_t* res_rows = nullptr; // optional gather for the residual rows
  const float* a2 = nullptr;   // optional second K-segment [m, k2]
  int64_t lda2 = 0;
  int32_t k2 = 0;
  const float* w = nullptr;    // [n, k1 + k2], row stride ldw (PyG Linear layout [out, in])
  int64_t ldw = 0;
  const float* bias = nullptr; // [n] or null
  const float* residual = nullptr;  // [m, n] row stride ldr, added to the result, or null
  int64_t ldr = 0;
  const float* res_mean = nullptr;  // optional normalisation of the residual on load (like a1)
  const float* res_scale = nullptr;
  const float* res_beta = nullptr;
  int32_t res_relu = 0;
  float* y = nullptr;          // [m, n], row stride ldy
  int64_t ldy = 0;
  int64_t m = 0;
  int32_t n = 0;
  const float* a1_mean = nullptr;   // [k1] BatchNorm-on-load parameters, all three or none
  const float* a1_scale = nullptr;
  const float* a1_beta = nullptr;
  int32_t relu_a1 = 0;         // apply ReLU to a1 on load (after the normalisation)
  int32_t relu_a2 = 0;
  const char* tag = "linear";  // profiling name
};
int launch_linear(const LinearArgs& args, cudaStream_t stream);

// Training-mode BatchNorm statistics of x [n, c] (row stride ldx): mean[c], scale[c] =
// weight / sqrt(var + eps), beta[c] = bias; optional running-stat update.  Deterministic
// (fixed-order fp64 partial sums).  scratch: bn_scratch_doubles(n, c) doubles.
size_t bn_scratch_doubles(int64_t n, int32_t c);
int bn_statistics(const float* x, int64_t ldx, int64_t n, int32_t c, const float* weight, const float* bias,
                  float eps, float momentum, float* running_mean, float* running_var, float* mean,
                  float* scale, float* beta, double* scratch, cudaStream_t stream);
// Same finalisation from column sums produced elsewhere (node_gemm epilogue), channel-major:
// partial[ch * P + p] = sum, partial[(c + ch) * P + p] = sum of squares of partition p of P.
int bn_finalize_partials(const double* partial, int64_t n_partials, int64_t n, int32_t c, const float* weight,
                         const float* bias, float eps, float momentum, float* running_mean, float* running_var,
                         float* mean, float* scale, float* beta, cudaStream_t stream);
// y = relu?((x - mean) * scale + beta)
// out_rows (optional): row r is written to y[out_rows[r]] (back to the caller's node order)
int bn_apply(const float* x, int64_t ldx, int64_t n, int32_t c, const float* mean, const float* scale,
             const float* beta, int32_t relu, float* y, int64_t ldy, cudaStream_t stream,
             const int32_t* out_rows = nullptr);

}  // namespace rgnn

// ---- device helpers -----------------------------------------------------------------
#ifdef __CUDACC__
namespace rgnn {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// order-preserving map double -> int64 (for atomicMin / atomicMax on coordinates)
__device__ __forceinline__ long long double_to_ordered(double d) {
  long long b = __double_as_longlong(d);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double ordered_to_double(long long b) {
  return __longlong_as_double(b >= 0 ? b : (b ^ 0x7fffffffffffffffLL));
}

}  // namespace rgnn
#endif

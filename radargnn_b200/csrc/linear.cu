// linear.cu -- dense node-level contractions: PyG `Linear` (y = x W^T + b) on a concatenated
// input [a1 | a2] with optional BatchNorm+ReLU applied to a1 on load, and the BatchNorm
// statistics / apply kernels (reference gnn/gnn_models.py:124-128, 137-178;
// gnn/mpnn_layers.py:89-90, 174-175).
//
// This is the fp32 CUDA-core (FFMA) path: exact fp32 products, fp32 accumulation, used for
// every shape.  Shared-memory tiled, 64x64 output tile per CTA, 4x4 outputs per thread.
#include "common.cuh"

namespace rgnn {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
linear_kernel(LinearArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, each a 4 x 4 block
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * BM;
  const int col0 = blockIdx.y * BN;
  const int K = p.k1 + p.k2;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: 64 rows x 16 k per tile, 4 elements per thread; k fastest (coalesced)
  const int lk = tid & 15, lr = tid >> 4;  // rows lr, lr + 16, lr + 32, lr + 48
  for (int k0 = 0; k0 < K; k0 += BK) {
    const int kg = k0 + lk;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = lr + i * 16;
      const int64_t row = row0 + r;
      float v = 0.f;
      if (row < p.m && kg < K) {
        if (kg < p.k1) {
          v = p.a1[(p.a1_rows != nullptr ? p.a1_rows[row] : row) * p.lda1 + kg];
          if (p.a1_mean != nullptr) v = (v - p.a1_mean[kg]) * p.a1_scale[kg] + p.a1_beta[kg];
          if (p.relu_a1) v = fmaxf(v, 0.f);
        } else {
          v = p.a2[row * p.lda2 + (kg - p.k1)];
          if (p.relu_a2) v = fmaxf(v, 0.f);
        }
      }
      As[lk][r] = v;
      const int col = col0 + r;
      Ws[lk][r] = (col < p.n && kg < K) ? p.w[static_cast<int64_t>(col) * p.ldw + kg] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= p.m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = col0 + tx * 4 + j;
      if (col >= p.n) continue;
      float v = acc[i][j];
      if (p.bias != nullptr) v += p.bias[col];
      if (p.residual != nullptr) {
        float r = p.residual[(p.res_rows != nullptr ? p.res_rows[row] : row) * p.ldr + col];
        if (p.res_mean != nullptr) r = (r - p.res_mean[col]) * p.res_scale[col] + p.res_beta[col];
        if (p.res_relu) r = fmaxf(r, 0.f);
        v += r;
      }
      p.y[row * p.ldy + col] = v;
    }
  }
}

// ---- BatchNorm (training mode) ------------------------------------------------------
constexpr int kBnRows = 512;  // rows per partial

// grid (ceil(n / kBnRows), ceil(c / 32)); block (32, 8): lanes over channels, 8 row groups
__global__ void __launch_bounds__(256)
bn_partial_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int c, double* __restrict__ partial) {
  __shared__ double ssum[8][33], ssq[8][33];
  const int ch = blockIdx.y * 32 + threadIdx.x;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kBnRows;
  const int64_t r1 = r0 + kBnRows < n ? r0 + kBnRows : n;
  double s = 0.0, q = 0.0;
  if (ch < c) {
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) {
      const double v = static_cast<double>(x[r * ldx + ch]);
      s += v;
      q += v * v;
    }
  }
  ssum[threadIdx.y][threadIdx.x] = s;
  ssq[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
#pragma unroll
    for (int g = 1; g < 8; ++g) { s += ssum[g][threadIdx.x]; q += ssq[g][threadIdx.x]; }
    partial[static_cast<int64_t>(ch) * gridDim.x + blockIdx.x] = s;          // channel-major, see bn_finalize_kernel
    partial[static_cast<int64_t>(c + ch) * gridDim.x + blockIdx.x] = q;
  }
}

// one block per channel: fixed-assignment strided partial sums + fixed-order tree => deterministic
__global__ void __launch_bounds__(128)
bn_finalize_kernel(const double* __restrict__ partial, int n_partials, int64_t n, int c,
                   const float* __restrict__ weight, const float* __restrict__ bias, float eps, float momentum,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean,
                   float* __restrict__ scale, float* __restrict__ beta) {
  __shared__ double ss[128], sq[128];
  const int ch = blockIdx.x;
  double s = 0.0, q = 0.0;
  for (int p = threadIdx.x; p < n_partials; p += 128) {
    s += partial[static_cast<int64_t>(ch) * n_partials + p];          // contiguous over p: coalesced
    q += partial[static_cast<int64_t>(c + ch) * n_partials + p];
  }
  ss[threadIdx.x] = s;
  sq[threadIdx.x] = q;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) { ss[threadIdx.x] += ss[threadIdx.x + o]; sq[threadIdx.x] += sq[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x != 0) return;
  s = ss[0]; q = sq[0];
  const double m = s / static_cast<double>(n);
  double var = q / static_cast<double>(n) - m * m;  // biased, used for the normalisation
  if (var < 0.0) var = 0.0;
  const float w = weight != nullptr ? weight[ch] : 1.f;
  const float b = bias != nullptr ? bias[ch] : 0.f;
  mean[ch] = static_cast<float>(m);
  scale[ch] = static_cast<float>(static_cast<double>(w) / sqrt(var + static_cast<double>(eps)));
  beta[ch] = b;
  if (running_mean != nullptr) running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(m);
  if (running_var != nullptr) {
    const double unbiased = n > 1 ? var * static_cast<double>(n) / static_cast<double>(n - 1) : var;
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int c, const float* __restrict__ mean,
                const float* __restrict__ scale, const float* __restrict__ beta, int relu,
                float* __restrict__ y, int64_t ldy, const int32_t* __restrict__ out_rows) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * c) return;
  const int64_t r = idx / c;
  const int ch = static_cast<int>(idx - r * c);
  float v = (x[r * ldx + ch] - mean[ch]) * scale[ch] + beta[ch];
  if (relu) v = fmaxf(v, 0.f);
  y[(out_rows != nullptr ? out_rows[r] : r) * ldy + ch] = v;
}

// 16-byte variant: c % 4 == 0, rows and parameter arrays 16-byte aligned; c / 4 consecutive threads per row
__global__ void __launch_bounds__(256)
bn_apply_vec4_kernel(const float* __restrict__ x, int64_t ldx, int64_t n, int c4, const float* __restrict__ mean,
                     const float* __restrict__ scale, const float* __restrict__ beta, int relu,
                     float* __restrict__ y, int64_t ldy, const int32_t* __restrict__ out_rows) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n * c4) return;
  const int64_t r = idx / c4;
  const int ch = static_cast<int>(idx - r * c4) << 2;
  const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + ch);
  const float4 mu = *reinterpret_cast<const float4*>(mean + ch);
  const float4 sc = *reinterpret_cast<const float4*>(scale + ch);
  const float4 be = *reinterpret_cast<const float4*>(beta + ch);
  float4 o = make_float4((v.x - mu.x) * sc.x + be.x, (v.y - mu.y) * sc.y + be.y, (v.z - mu.z) * sc.z + be.z,
                         (v.w - mu.w) * sc.w + be.w);
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  *reinterpret_cast<float4*>(y + (out_rows != nullptr ? static_cast<int64_t>(out_rows[r]) : r) * ldy + ch) = o;
}

constexpr int kSumBlocks = 592;  // 4 x 148 SMs

__global__ void __launch_bounds__(256)
sum_partial_kernel(const float* __restrict__ x, int64_t count, double* __restrict__ partial) {
  __shared__ double warp_sums[8];
  double s = 0.0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t vec = count >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < vec; i += stride) {
    const float4 v = x4[i];
    s += (static_cast<double>(v.x) + static_cast<double>(v.y)) + (static_cast<double>(v.z) + static_cast<double>(v.w));
  }
  for (int64_t i = (vec << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
    s += static_cast<double>(x[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += warp_sums[w];
    partial[blockIdx.x] = t;
  }
}

// one warp: lane l adds partials l, l + 32, ... in order, then a fixed butterfly => deterministic
__global__ void __launch_bounds__(32)
sum_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ result) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) t += partial[i];
  t = warp_sum(t);
  if (threadIdx.x == 0) *result = t;
}

}  // namespace

int launch_linear(const LinearArgs& args, cudaStream_t stream) {
  if (args.m <= 0 || args.n <= 0) return RGNN_OK;
  dim3 grid(div_up(args.m, BM), div_up(args.n, BN));
  RGNN_PROFILE(args.tag, stream);
  linear_kernel<<<grid, 256, 0, stream>>>(args);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

size_t bn_scratch_doubles(int64_t n, int32_t c) {
  const int64_t parts = (n + kBnRows - 1) / kBnRows;
  return static_cast<size_t>((parts > 0 ? parts : 1) * 2 * c);
}

int bn_statistics(const float* x, int64_t ldx, int64_t n, int32_t c, const float* weight, const float* bias,
                  float eps, float momentum, float* running_mean, float* running_var, float* mean,
                  float* scale, float* beta, double* scratch, cudaStream_t stream) {
  if (n <= 0 || c <= 0) return RGNN_OK;
  const int parts = static_cast<int>((n + kBnRows - 1) / kBnRows);
  dim3 grid(parts, div_up(c, 32)), block(32, 8);
  RGNN_PROFILE("bn_statistics", stream);
  bn_partial_kernel<<<grid, block, 0, stream>>>(x, ldx, n, c, scratch);
  RGNN_LAUNCH_CHECK();
  bn_finalize_kernel<<<c, 128, 0, stream>>>(scratch, parts, n, c, weight, bias, eps, momentum,
                                                         running_mean, running_var, mean, scale, beta);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int bn_finalize_partials(const double* partial, int64_t n_partials, int64_t n, int32_t c, const float* weight,
                         const float* bias, float eps, float momentum, float* running_mean, float* running_var,
                         float* mean, float* scale, float* beta, cudaStream_t stream) {
  if (n <= 0 || c <= 0) return RGNN_OK;
  RGNN_PROFILE("bn_statistics", stream);
  bn_finalize_kernel<<<c, 128, 0, stream>>>(partial, static_cast<int>(n_partials), n, c, weight, bias, eps,
                                                         momentum, running_mean, running_var, mean, scale, beta);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int bn_apply(const float* x, int64_t ldx, int64_t n, int32_t c, const float* mean, const float* scale,
             const float* beta, int32_t relu, float* y, int64_t ldy, cudaStream_t stream, const int32_t* out_rows) {
  if (n <= 0 || c <= 0) return RGNN_OK;
  RGNN_PROFILE("bn_apply", stream);
  auto al16 = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
  if (c % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && al16(x) && al16(y) && al16(mean) && al16(scale) && al16(beta))
    bn_apply_vec4_kernel<<<div_up(n * (c / 4), 256), 256, 0, stream>>>(x, ldx, n, c / 4, mean, scale, beta, relu, y, ldy, out_rows);
  else
    bn_apply_kernel<<<div_up(n * c, 256), 256, 0, stream>>>(x, ldx, n, c, mean, scale, beta, relu, y, ldy, out_rows);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // namespace rgnn

using namespace rgnn;

extern "C" {

int rgnn_linear_forward(const float* x, int64_t n, int32_t in_features, const float* weight, const float* bias,
                        int32_t out_features, int32_t relu_input, float* y, rgnn_stream_t stream) {
  if (n < 0 || in_features < 1 || out_features < 1) return RGNN_ERR_INVALID_ARGUMENT;
  if (n == 0) return RGNN_OK;
  if (x == nullptr || weight == nullptr || y == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  LinearArgs a;
  a.a1 = x; a.lda1 = in_features; a.k1 = in_features;
  a.w = weight; a.ldw = in_features; a.bias = bias;
  a.y = y; a.ldy = out_features; a.m = n; a.n = out_features;
  a.relu_a1 = relu_input;
  return launch_linear(a, static_cast<cudaStream_t>(stream));
}

int rgnn_affine_relu_forward(const float* x, int64_t n, int32_t channels, const float* mean, const float* scale,
                             const float* beta, int32_t apply_relu, float* out, rgnn_stream_t stream) {
  if (n < 0 || channels < 1) return RGNN_ERR_INVALID_ARGUMENT;
  if (n == 0) return RGNN_OK;
  if (x == nullptr || mean == nullptr || scale == nullptr || beta == nullptr || out == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  return bn_apply(x, channels, n, channels, mean, scale, beta, apply_relu, out, channels, static_cast<cudaStream_t>(stream));
}

size_t rgnn_sum_workspace_bytes(void) { return sizeof(double) * kSumBlocks + kAlign; }

int rgnn_sum_f32(const float* x, int64_t count, double* result, void* workspace, size_t workspace_bytes,
                 rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (count < 0 || result == nullptr || (count > 0 && x == nullptr)) return RGNN_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(x) % 16 != 0) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_sum_workspace_bytes()) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  Arena arena(workspace, workspace_bytes);
  double* partial = arena.take<double>(kSumBlocks);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  RGNN_PROFILE("loss_sum", stream);
  sum_partial_kernel<<<kSumBlocks, 256, 0, stream>>>(x, count, partial);
  RGNN_LAUNCH_CHECK();
  sum_final_kernel<<<1, 32, 0, stream>>>(partial, kSumBlocks, result);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

size_t rgnn_batchnorm_workspace_bytes(int64_t n, int32_t channels) {
  if (n < 0 || channels < 0) return 0;
  SizeArena a;
  a.take<double>(bn_scratch_doubles(n, channels));
  a.take<float>(static_cast<size_t>(channels) * 3);
  return a.used;
}

int rgnn_batchnorm_relu_forward(const float* x, int64_t n, int32_t channels, const float* weight,
                                const float* bias, float eps, float momentum, float* running_mean,
                                float* running_var, int32_t apply_relu, float* out, void* workspace,
                                size_t workspace_bytes, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n < 0 || channels < 1) return RGNN_ERR_INVALID_ARGUMENT;
  if (n == 0) return RGNN_OK;
  if (x == nullptr || out == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_batchnorm_workspace_bytes(n, channels)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  Arena arena(workspace, workspace_bytes);
  double* scratch = arena.take<double>(bn_scratch_doubles(n, channels));
  float* stats = arena.take<float>(static_cast<size_t>(channels) * 3);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  float* mean = stats, *scale = stats + channels, *beta = stats + 2 * channels;
  RGNN_RETURN_IF_ERROR(bn_statistics(x, channels, n, channels, weight, bias, eps, momentum, running_mean,
                                     running_var, mean, scale, beta, scratch, stream));
  return bn_apply(x, channels, n, channels, mean, scale, beta, apply_relu, out, channels, stream);
}

}  // extern "C"

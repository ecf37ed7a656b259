// conv_backward.cu -- the graph-specific half of the backward pass of MPNNConv / RadarPointGNNConv
// (autograd through reference gnn/mpnn_layers.py:86-101, 171-184, driven by gnn/trainer.py:228-231).
//
// With the factored message  m_e = A[t_e] + B[s_e] + W_e e_e + b  (conv.cu) the gradient of the aggregated
// messages gM [N, P] is routed back to
//     GA[n]   = sum over n's incoming edges of gm_e        (grad of the target term A = x W_t^T, and of b)
//     GB[n]   = sum over n's outgoing edges of gm_e        (grad of the source term B = x W_s^T)
//     dW_e    = sum_e gm_e^T e_e,      d e_e = gm_e W_e
// where gm_e is gM[t_e] on the channels whose max / min edge e wins (torch_scatter's arg-max routing), gM[t_e] for
// `add`, gM[t_e] / deg for `mean`.  Nothing is materialised per edge: one thread per (target node, 4 channels)
// recomputes the node's messages over its CSC segment (the same gather as the forward), finds the winning slot
// per channel and scatters.  The kernel also returns the aggregated messages M themselves (the forward keeps
// them on chip; dW_post = dy^T [x ; M] needs them).  The dense contractions of the backward
// (dy W_post, GA W_t, GB W_s, the weight gradients) are plain library GEMMs on the Python side.
//
// Ties between two edges on a channel (exact float equality) go to the first slot of the segment; sums into GB /
// d e / dW_e use atomics (fp32: the order, and with it the last bits, may vary between runs).
#include <math.h>

#include "common.cuh"

namespace rgnn {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int MODE>
__global__ void __launch_bounds__(kThreads)
conv_route_kernel(const float* __restrict__ a, const float* __restrict__ b, int pp, int p, const float* __restrict__ bias,
                  const float* __restrict__ w_e, int64_t ldwe, int de, const float* __restrict__ ea,
                  const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_src, int64_t n_nodes,
                  const float* __restrict__ gm, int64_t ldg, float* __restrict__ m_out, float* __restrict__ ga,
                  float* __restrict__ gb, float* __restrict__ dwe, float* __restrict__ dea) {
  extern __shared__ float sm[];   // [de][pp] W_e transposed, then [pp][de] block-local dW_e
  float* ws = sm;
  float* dws = sm + de * pp;
  for (int i = threadIdx.x; i < de * pp; i += blockDim.x) {
    const int d = i / pp, c = i - d * pp;
    ws[i] = c < p ? w_e[static_cast<int64_t>(c) * ldwe + d] : 0.f;
    dws[i] = 0.f;
  }
  __syncthreads();
  const int chunks = pp >> 2;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t node = gid / chunks;
  const int c0 = static_cast<int>(gid - node * chunks) << 2;
  if (node < n_nodes) {
    const int beg = csc_ptr[node], end = csc_ptr[node + 1], deg = end - beg;
    float g[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) g[j] = (c0 + j < p && deg > 0) ? gm[node * ldg + c0 + j] : 0.f;
    float base[4];   // A_n + b
#pragma unroll
    for (int j = 0; j < 4; ++j) base[j] = (c0 + j < p ? bias[c0 + j] : 0.f) + (a != nullptr ? a[node * pp + c0 + j] : 0.f);
    float best[4], acc[4] = {0.f, 0.f, 0.f, 0.f};
    int arg[4] = {beg, beg, beg, beg};
#pragma unroll
    for (int j = 0; j < 4; ++j) best[j] = MODE == RGNN_AGGR_MIN ? INFINITY : -INFINITY;
    // pass 1: messages of the segment -> aggregate (+ winning slot)
    for (int slot = beg; slot < end; ++slot) {
      const float4 bv = ld4(b + static_cast<int64_t>(csc_src[slot]) * pp + c0);
      float v[4] = {bv.x, bv.y, bv.z, bv.w};
      for (int d = 0; d < de; ++d) {
        const float e = ea[static_cast<int64_t>(slot) * de + d];
        const float4 w4 = ld4(ws + d * pp + c0);
        v[0] = fmaf(e, w4.x, v[0]); v[1] = fmaf(e, w4.y, v[1]); v[2] = fmaf(e, w4.z, v[2]); v[3] = fmaf(e, w4.w, v[3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == RGNN_AGGR_MAX) { if (v[j] > best[j]) { best[j] = v[j]; arg[j] = slot; } }
        else if (MODE == RGNN_AGGR_MIN) { if (v[j] < best[j]) { best[j] = v[j]; arg[j] = slot; } }
        else acc[j] += v[j];
      }
    }
    // aggregated messages (torch_scatter: an empty segment aggregates to 0)
    float mo[4] = {0.f, 0.f, 0.f, 0.f};
    if (deg > 0) {
      const float fd = static_cast<float>(deg);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN) mo[j] = base[j] + best[j];
        else if (MODE == RGNN_AGGR_ADD) mo[j] = fmaf(fd, base[j], acc[j]);
        else mo[j] = base[j] + acc[j] / fd;
      }
    }
    if (m_out != nullptr) *reinterpret_cast<float4*>(m_out + node * pp + c0) = make_float4(mo[0], mo[1], mo[2], mo[3]);
    // GA: every incoming edge carries the node's target term once
    const float ga_scale = MODE == RGNN_AGGR_ADD ? static_cast<float>(deg) : 1.f;
    *reinterpret_cast<float4*>(ga + node * pp + c0) = make_float4(g[0] * ga_scale, g[1] * ga_scale, g[2] * ga_scale, g[3] * ga_scale);
    // pass 2: scatter
    if (deg > 0 && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f || g[3] != 0.f)) {
      if (MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c0 + j >= p || g[j] == 0.f) continue;
          const int slot = arg[j];
          atomicAdd(gb + static_cast<int64_t>(csc_src[slot]) * pp + c0 + j, g[j]);
          for (int d = 0; d < de; ++d) {
            const float e = ea[static_cast<int64_t>(slot) * de + d];
            atomicAdd(dws + (c0 + j) * de + d, g[j] * e);
            if (dea != nullptr) atomicAdd(dea + static_cast<int64_t>(slot) * de + d, g[j] * ws[d * pp + c0 + j]);
          }
        }
      } else {
        const float sc = MODE == RGNN_AGGR_MEAN ? 1.f / static_cast<float>(deg) : 1.f;
        const float gs[4] = {g[0] * sc, g[1] * sc, g[2] * sc, g[3] * sc};
        for (int slot = beg; slot < end; ++slot) {
          float* gbrow = gb + static_cast<int64_t>(csc_src[slot]) * pp + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c0 + j < p) atomicAdd(gbrow + j, gs[j]);
          for (int d = 0; d < de; ++d) {
            const float e = ea[static_cast<int64_t>(slot) * de + d];
            float de_acc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (c0 + j < p) { atomicAdd(dws + (c0 + j) * de + d, gs[j] * e); de_acc = fmaf(gs[j], ws[d * pp + c0 + j], de_acc); }
            }
            if (dea != nullptr) atomicAdd(dea + static_cast<int64_t>(slot) * de + d, de_acc);
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p * de; i += blockDim.x)
    if (dws[i] != 0.f) atomicAdd(dwe + i, dws[i]);
}

}  // namespace
}  // namespace rgnn

using namespace rgnn;

extern "C" {

int rgnn_conv_backward_route(int32_t aggr, const float* a, const float* b, int32_t p, const float* bias, const float* w_e,
                             int64_t ldwe, int32_t de, const float* ea_csc, const int32_t* csc_ptr, const int32_t* csc_src,
                             int64_t n_nodes, int64_t n_edges, const float* grad_m, int64_t ld_grad, float* m_out, float* grad_a,
                             float* grad_b, float* grad_we, float* grad_ea_csc, rgnn_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n_nodes < 0 || n_edges < 0 || p < 1 || de < 0 || aggr < RGNN_AGGR_MAX || aggr > RGNN_AGGR_MIN) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_nodes == 0) return RGNN_OK;
  if (b == nullptr || bias == nullptr || csc_ptr == nullptr || grad_m == nullptr || grad_a == nullptr || grad_b == nullptr ||
      ld_grad < p)
    return RGNN_ERR_INVALID_ARGUMENT;
  if (de > 0 && (w_e == nullptr || grad_we == nullptr || (n_edges > 0 && ea_csc == nullptr))) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && csc_src == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  const int pp = (p + 3) & ~3;
  if (reinterpret_cast<uintptr_t>(b) % 16 != 0 || reinterpret_cast<uintptr_t>(grad_a) % 16 != 0 ||
      (m_out != nullptr && reinterpret_cast<uintptr_t>(m_out) % 16 != 0))
    return RGNN_ERR_INVALID_ARGUMENT;
  // grad_b, grad_we and grad_ea_csc are accumulated with atomics: start from zero
  RGNN_CUDA_CHECK(cudaMemsetAsync(grad_b, 0, sizeof(float) * n_nodes * pp, stream));
  if (de > 0) RGNN_CUDA_CHECK(cudaMemsetAsync(grad_we, 0, sizeof(float) * p * de, stream));
  if (grad_ea_csc != nullptr && n_edges > 0 && de > 0) RGNN_CUDA_CHECK(cudaMemsetAsync(grad_ea_csc, 0, sizeof(float) * n_edges * de, stream));
  const size_t smem = sizeof(float) * 2 * static_cast<size_t>(de) * pp;
  if (smem > 48 * 1024) return RGNN_ERR_UNSUPPORTED;
  const int64_t threads = n_nodes * (pp >> 2);
  const unsigned blocks = div_up(threads, kThreads);
  RGNN_PROFILE("conv_backward_route", stream);
#define RGNN_ROUTE(MODE_)                                                                                             \
  conv_route_kernel<MODE_><<<blocks, kThreads, smem, stream>>>(a, b, pp, p, bias, w_e, ldwe, de, ea_csc, csc_ptr, csc_src, \
                                                               n_nodes, grad_m, ld_grad, m_out, grad_a, grad_b, grad_we, grad_ea_csc)
  switch (aggr) {
    case RGNN_AGGR_MAX: RGNN_ROUTE(RGNN_AGGR_MAX); break;
    case RGNN_AGGR_MIN: RGNN_ROUTE(RGNN_AGGR_MIN); break;
    case RGNN_AGGR_ADD: RGNN_ROUTE(RGNN_AGGR_ADD); break;
    default: RGNN_ROUTE(RGNN_AGGR_MEAN); break;
  }
#undef RGNN_ROUTE
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

}  // extern "C"

// conv.cu -- one graph-convolution layer: MPNNConv / RadarPointGNNConv forward
// (reference gnn/mpnn_layers.py:86-101, 171-184; PyG propagate = index_select x2 -> cat ->
// Linear -> torch_scatter reduce at edge_index[1] -> cat -> Linear).
//
// Factored formulation (hard part 4).  The first Linear of pre_mlp is
// affine in its concatenated input, so with W_pre = [W_t | W_s | W_e] (column blocks for
// x_i = x[target], x_j = x[source], edge attributes)
//     m_e = W_t x_t + W_s x_s + W_e e_e + b
// and two node-level contractions A = x W_t^T, B = x W_s^T replace the E x P x P per-edge GEMM:
//     max:  M_n = A_n + b + max_e (B[s_e] + W_e e_e)           (0 when n has no incoming edge)
//     add:  M_n = deg_n (A_n + b) + sum_e (B[s_e] + W_e e_e)
//     mean: M_n = A_n + b + mean_e (B[s_e] + W_e e_e)          (0 when deg_n = 0)
// The edge kernel is then a gather + segmented reduce over the CSC view, nothing is
// materialised per edge.  With pre_layers > 1 (ReLU inside the message MLP) the per-edge
// activations U = A[t] + B[s] + W_e e + b are materialised and pushed through the remaining
// Linear layers before the segmented reduce.  An edge encoder (Linear De -> C) is folded into
// W_e and b.  Results differ from the reference only by fp32 rounding order.
#include <math.h>
#include <stdlib.h>

#include "conv.cuh"

namespace rgnn {
namespace {

constexpr int kAggThreads = 256;

// W_e (row-major [p, ldw], rows = output channels) -> shared [de][pp], zero padded
__device__ __forceinline__ void stage_edge_weights(const float* __restrict__ w_e, int64_t ldw, int p, int pp,
                                                   int de, float* __restrict__ smem) {
  for (int i = threadIdx.x; i < de * pp; i += blockDim.x) {
    const int d = i / pp, c = i - d * pp;
    smem[i] = c < p ? w_e[static_cast<int64_t>(c) * ldw + d] : 0.f;
  }
  __syncthreads();
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 fma4(float s, float4 w, float4 a) {
  return make_float4(fmaf(s, w.x, a.x), fmaf(s, w.y, a.y), fmaf(s, w.z, a.z), fmaf(s, w.w, a.w));
}
__device__ __forceinline__ float4 max4(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }
__device__ __forceinline__ float4 min4(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }

// One thread per (target node, 4-channel chunk).  MODE: rgnn_aggr, or -1 = write the per-edge
// activations U[slot] = A[t] + B[s] + W_e e + b instead of reducing (general path).
// DE > 0: edge-attribute width known at compile time -- the thread's 4 x DE edge weights live in
// registers and the loop carries no shared-memory traffic; DE == 0: runtime width, weights in smem.
template <int MODE, int DE>
__global__ void __launch_bounds__(kAggThreads)
edge_aggregate_kernel(const float* __restrict__ a, const float* __restrict__ b, int pp, int p,
                      const float* __restrict__ bias, const float* __restrict__ w_e, int64_t ldwe, int de,
                      const float* __restrict__ ea, const int32_t* __restrict__ csc_ptr,
                      const int32_t* __restrict__ csc_src, int64_t n_nodes, float* __restrict__ out,
                      IsolatedNodeTerm iso) {
  extern __shared__ float w_s[];  // [de][pp] (DE == 0 only)
  if (DE == 0) stage_edge_weights(w_e, ldwe, p, pp, de, w_s);
  // kSplit = 2 would let two adjacent lanes share a (node, chunk), one on the even slots and one on
  // the odd ones, combined with a shuffle.  Measured slower (0.66 vs 0.49 ms per step at the headline
  // size): the kernel is bound by instruction issue / L1 wavefronts, not by gathers in flight.
  constexpr int kSplit = 1;
  const int chunks = pp >> 2;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t node = gid / (chunks * kSplit);
  const bool live = node < n_nodes;
  const int rem = static_cast<int>(gid - node * (chunks * kSplit));
  const int c0 = (kSplit == 2 ? rem >> 1 : rem) << 2;
  const int part = kSplit == 2 ? (rem & 1) : 0;
  const int seg_beg = live ? csc_ptr[node] : 0, seg_end = live ? csc_ptr[node + 1] : 0;
  const int beg = seg_beg + part, end = seg_end;

  float4 wreg[DE > 0 ? DE : 1];
  if (DE > 0) {
#pragma unroll
    for (int d = 0; d < DE; ++d) {
      wreg[d].x = c0 + 0 < p ? w_e[static_cast<int64_t>(c0 + 0) * ldwe + d] : 0.f;
      wreg[d].y = c0 + 1 < p ? w_e[static_cast<int64_t>(c0 + 1) * ldwe + d] : 0.f;
      wreg[d].z = c0 + 2 < p ? w_e[static_cast<int64_t>(c0 + 2) * ldwe + d] : 0.f;
      wreg[d].w = c0 + 3 < p ? w_e[static_cast<int64_t>(c0 + 3) * ldwe + d] : 0.f;
    }
  }

  float4 base = make_float4(0.f, 0.f, 0.f, 0.f);  // A_n + b for this chunk
  if (a != nullptr && live) base = ld4(a + node * pp + c0);
  {
    float4 bb;
    bb.x = c0 + 0 < p ? bias[c0 + 0] : 0.f;
    bb.y = c0 + 1 < p ? bias[c0 + 1] : 0.f;
    bb.z = c0 + 2 < p ? bias[c0 + 2] : 0.f;
    bb.w = c0 + 3 < p ? bias[c0 + 3] : 0.f;
    base = add4(base, bb);
  }

  float4 acc;
  if (MODE == RGNN_AGGR_MAX) acc = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  else if (MODE == RGNN_AGGR_MIN) acc = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
  else acc = make_float4(0.f, 0.f, 0.f, 0.f);

  auto edge_term = [&](int slot, float4 v) {
    if (DE > 0) {
      const float* e = ea + static_cast<int64_t>(slot) * DE;
      if (DE == 2) {
        const float2 e2 = *reinterpret_cast<const float2*>(e);
        v = fma4(e2.x, wreg[0], v);
        v = fma4(e2.y, wreg[DE > 1 ? 1 : 0], v);
      } else if (DE == 4) {
        const float4 e4 = *reinterpret_cast<const float4*>(e);
        v = fma4(e4.x, wreg[0], v);
        v = fma4(e4.y, wreg[DE > 1 ? 1 : 0], v);
        v = fma4(e4.z, wreg[DE > 2 ? 2 : 0], v);
        v = fma4(e4.w, wreg[DE > 3 ? 3 : 0], v);
      } else {
#pragma unroll
        for (int d = 0; d < DE; ++d) v = fma4(e[d], wreg[d], v);
      }
    } else {
      const float* e = ea + static_cast<int64_t>(slot) * de;
      for (int d = 0; d < de; ++d) v = fma4(e[d], ld4(w_s + d * pp + c0), v);
    }
    return v;
  };

  const float* bcol = b + c0;
  int slot = beg;
  // 4 gathers in flight per thread
  for (; slot + 3 * kSplit < end; slot += 4 * kSplit) {
    const int s0 = csc_src[slot], s1 = csc_src[slot + kSplit], s2 = csc_src[slot + 2 * kSplit], s3 = csc_src[slot + 3 * kSplit];
    float4 v0 = ld4(bcol + static_cast<int64_t>(s0) * pp);
    float4 v1 = ld4(bcol + static_cast<int64_t>(s1) * pp);
    float4 v2 = ld4(bcol + static_cast<int64_t>(s2) * pp);
    float4 v3 = ld4(bcol + static_cast<int64_t>(s3) * pp);
    v0 = edge_term(slot, v0); v1 = edge_term(slot + kSplit, v1);
    v2 = edge_term(slot + 2 * kSplit, v2); v3 = edge_term(slot + 3 * kSplit, v3);
    if (MODE == RGNN_AGGR_MAX) acc = max4(acc, max4(max4(v0, v1), max4(v2, v3)));
    else if (MODE == RGNN_AGGR_MIN) acc = min4(acc, min4(min4(v0, v1), min4(v2, v3)));
    else if (MODE == -1) {
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(slot) * pp + c0) = add4(base, v0);
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(slot + 1) * pp + c0) = add4(base, v1);
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(slot + 2) * pp + c0) = add4(base, v2);
      *reinterpret_cast<float4*>(out + static_cast<int64_t>(slot + 3) * pp + c0) = add4(base, v3);
    } else {  // add / mean: fixed slot order
      acc = add4(acc, v0); acc = add4(acc, v1); acc = add4(acc, v2); acc = add4(acc, v3);
    }
  }
  for (; slot < end; slot += kSplit) {
    float4 v = ld4(bcol + static_cast<int64_t>(csc_src[slot]) * pp);
    v = edge_term(slot, v);
    if (MODE == RGNN_AGGR_MAX) acc = max4(acc, v);
    else if (MODE == RGNN_AGGR_MIN) acc = min4(acc, v);
    else if (MODE == -1) *reinterpret_cast<float4*>(out + static_cast<int64_t>(slot) * pp + c0) = add4(base, v);
    else acc = add4(acc, v);
  }
  if (MODE == -1) return;
  if (kSplit == 2) {  // combine the even-slot and odd-slot halves (fixed order: deterministic)
    float4 o;
    o.x = __shfl_xor_sync(0xffffffffu, acc.x, 1); o.y = __shfl_xor_sync(0xffffffffu, acc.y, 1);
    o.z = __shfl_xor_sync(0xffffffffu, acc.z, 1); o.w = __shfl_xor_sync(0xffffffffu, acc.w, 1);
    if (MODE == RGNN_AGGR_MAX) acc = max4(acc, o);
    else if (MODE == RGNN_AGGR_MIN) acc = min4(acc, o);
    else acc = part == 0 ? add4(acc, o) : add4(o, acc);
  }
  if (!live || part != 0) return;
  const int deg = seg_end - seg_beg;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);  // torch_scatter: empty segments aggregate to 0
  if (deg == 0 && iso.w_t != nullptr) {
    // Folded formulation (node_gemm.cu): the update weights carry W_m W_t for every node, so a
    // node without incoming edge cancels that term with M' = -W_t x_n instead of 0.
    float acc4[4] = {0.f, 0.f, 0.f, 0.f};
    const float* xr = iso.x + (iso.rows != nullptr ? static_cast<int64_t>(iso.rows[node]) : node) * iso.ldx;
    for (int i = 0; i < iso.c; ++i) {
      float xv = xr[i];
      if (iso.mean != nullptr) xv = (xv - iso.mean[i]) * iso.scale[i] + iso.beta[i];
      if (iso.relu) xv = fmaxf(xv, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c0 + j < p) acc4[j] = fmaf(iso.w_t[static_cast<int64_t>(c0 + j) * iso.ldw + i], xv, acc4[j]);
    }
    r = make_float4(-acc4[0], -acc4[1], -acc4[2], -acc4[3]);
  }
  if (deg > 0) {
    if (MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN) r = add4(base, acc);
    else if (MODE == RGNN_AGGR_ADD) {
      const float fd = static_cast<float>(deg);
      r = make_float4(fmaf(fd, base.x, acc.x), fmaf(fd, base.y, acc.y), fmaf(fd, base.z, acc.z), fmaf(fd, base.w, acc.w));
    } else {
      const float inv = 1.f / static_cast<float>(deg);
      r = make_float4(fmaf(acc.x, inv, base.x), fmaf(acc.y, inv, base.y), fmaf(acc.z, inv, base.z), fmaf(acc.w, inv, base.w));
    }
  }
  *reinterpret_cast<float4*>(out + node * pp + c0) = r;
}

// ---- split-layout aggregate (tensor-core path, ConvShape::split) ---------------------------------
// A QUARTER-WARP (8 lanes) per target node, lane q owns main channels 32 i + 4 q .. + 3 (i = 0..3), so
// one gather instruction of the warp reads one aligned 128-byte line of four different source rows
// (4 L1 wavefronts, all bytes used) and the per-slot bookkeeping -- broadcasting the slot's source
// index and edge attributes with shuffles, the 64-bit address -- is paid once per FOUR rows.  (One thread
// per (node, 4 channels), the layout above, pays ~8 wavefronts and ~34 instructions per slot plus a
// strided per-thread weight prologue as expensive as its main loop; a full warp per node pays 3 shuffles
// per slot per row, and shuffles occupy the same L1 data pipe as the gathers.)  The slot's source index
// and edge attributes are loaded once per 8 slots (lane = slot) and four slots = 16 gathers per lane are
// in flight.  The edge term runs on packed FFMA2, max / min on the 3-input FMNMX3.  The p - 128 tail
// channels live in their own narrow arrays and are processed slot-parallel (lane = slot) with a fixed
// butterfly over the 8 lanes at the end of the row.  One persistent CTA per SM walks a CONTIGUOUS range of
// (cell-sorted) nodes, 48 at a time: everything an SM gathers concurrently, and from one pass to the
// next, comes from the same few grid rows, so the L1 working set stays small (more, independent CTAs per
// SM measured slower: 3 -> 4 -> 5 resident 64-row blocks took 95 -> 116 -> 217 us, L1 thrashing).
constexpr int kSplitMain = 128;

__device__ __forceinline__ float4 fma4x2(float s, float4 w, float4 a) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __ffma2_rn(ss, make_float2(w.x, w.y), make_float2(a.x, a.y));
  const float2 hi = __ffma2_rn(ss, make_float2(w.z, w.w), make_float2(a.z, a.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}

template <int MODE>
__device__ __forceinline__ float4 combine4(float4 a, float4 b) {
  if (MODE == RGNN_AGGR_MAX) return max4(a, b);
  if (MODE == RGNN_AGGR_MIN) return min4(a, b);
  return add4(a, b);
}
// acc (op) v0 (op) v1, in this order (the sums must stay in slot order)
template <int MODE>
__device__ __forceinline__ float4 combine4x2(float4 a, float4 v0, float4 v1) {
  if (MODE == RGNN_AGGR_MAX)
    return make_float4(fmaxf(fmaxf(a.x, v0.x), v1.x), fmaxf(fmaxf(a.y, v0.y), v1.y), fmaxf(fmaxf(a.z, v0.z), v1.z), fmaxf(fmaxf(a.w, v0.w), v1.w));
  if (MODE == RGNN_AGGR_MIN)
    return make_float4(fminf(fminf(a.x, v0.x), v1.x), fminf(fminf(a.y, v0.y), v1.y), fminf(fminf(a.z, v0.z), v1.z), fminf(fminf(a.w, v0.w), v1.w));
  return add4(add4(a, v0), v1);
}

// kSplitThreads: 384 (12 warps, 168 registers) or 512 (16 warps, 128 registers): one persistent CTA per SM
template <int MODE, int DE, int kSplitThreads>
__global__ void __launch_bounds__(kSplitThreads, 1)
edge_aggregate_split_kernel(const float* __restrict__ bm, const float* __restrict__ bt, int p,
                            const float* __restrict__ bias, const float* __restrict__ w_e, int64_t ldwe,
                            const float* __restrict__ ea, const int32_t* __restrict__ csc_ptr,
                            const int32_t* __restrict__ csc_src, int n_nodes, int rows_per_cta,
                            float* __restrict__ out_m, float* __restrict__ out_t, IsolatedNodeTerm iso) {
  constexpr int kSplitPassRows = kSplitThreads / 32 * 4;  // rows a CTA works on at a time
  constexpr int kW = kSplitMain + 4;             // staged channels: main + one float4 of tail
  __shared__ __align__(16) float ws[DE][kW];     // W_e transposed: ws[d][channel], zero beyond p
  __shared__ __align__(16) float bs[kW];         // message bias, zero beyond p
  for (int i = threadIdx.x; i < DE * kW; i += blockDim.x) {
    const int d = i / kW, ch = i - d * kW;
    ws[d][ch] = ch < p ? w_e[static_cast<int64_t>(ch) * ldwe + d] : 0.f;
  }
  for (int i = threadIdx.x; i < kW; i += blockDim.x) bs[i] = i < p ? bias[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane & 7, quarter = lane >> 3;
  float4 w[DE][4], wt[DE];
#pragma unroll
  for (int d = 0; d < DE; ++d) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[d][i] = ld4(&ws[d][32 * i + 4 * q]);
    wt[d] = ld4(&ws[d][kSplitMain]);
  }
  constexpr float kInit = MODE == RGNN_AGGR_MAX ? -INFINITY : (MODE == RGNN_AGGR_MIN ? INFINITY : 0.f);
  const float4 init4 = make_float4(kInit, kInit, kInit, kInit);
  const float* bcol = bm + 4 * q;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(n_nodes, row_begin + rows_per_cta);
  const int kPasses = (max(row_end - row_begin, 0) + kSplitPassRows - 1) / kSplitPassRows;
  constexpr bool kOrderFree = MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN;
  constexpr int kHold = kSplitThreads > 384 ? 2 : 4;   // slots held per lane: a quarter holds 8 * kHold slots of its row
  const int row_base = row_begin + warp * 4 + quarter;

  // Software pipeline over the passes: the row pointers run two passes ahead and the first 32 slots
  // (lane q holds slots q, q + 8, q + 16, q + 24: source index, edge attributes) one pass ahead, so that
  // a pass starts its gathers without waiting for an index load.
  struct SlotRegs { int src[kHold]; float e[kHold][DE]; };
  auto load_ptr = [&](int pass, int& beg, int& deg) {
    const int row = row_base + pass * kSplitPassRows;
    beg = 0; deg = 0;
    if (pass < kPasses && row < row_end) { beg = csc_ptr[row]; deg = csc_ptr[row + 1] - beg; }
  };
  // source indices of the lane's slots: needed to issue the gathers, prefetched one pass ahead
  auto load_src = [&](int beg, int deg, int b, int (&src)[kHold]) {
#pragma unroll
    for (int h = 0; h < kHold; ++h) src[h] = (b + 8 * h + q < deg) ? csc_src[beg + b + 8 * h + q] : 0;
  };
  // edge attributes of the lane's slots: first used after the pass's first gathers have landed, so they are
  // loaded in the pass itself (zero for slots beyond the row's degree)
  auto load_attr = [&](int beg, int deg, int b, float (&e)[kHold][DE]) {
#pragma unroll
    for (int h = 0; h < kHold; ++h) {
      const bool on = b + 8 * h + q < deg;
      const float* ep = ea + static_cast<int64_t>(beg + b + 8 * h + q) * DE;
#pragma unroll
      for (int d = 0; d < DE; ++d) e[h][d] = on ? ep[d] : 0.f;
    }
  };
  int beg, deg, beg1, deg1, beg2, deg2;
  load_ptr(0, beg, deg);
  load_ptr(1, beg1, deg1);
  int pre_src[kHold];
  load_src(beg, deg, 0, pre_src);

  for (int it = 0; it < kPasses; ++it) {
    const int row = row_base + it * kSplitPassRows;
    const bool live = row < row_end;
    load_ptr(it + 2, beg2, deg2);
    SlotRegs cur;
#pragma unroll
    for (int h = 0; h < kHold; ++h) cur.src[h] = pre_src[h];
    load_attr(beg, deg, 0, cur.e);
    load_src(beg1, deg1, 0, pre_src);   // next pass (zeros past the last pass)
    const int nmax = __reduce_max_sync(0xffffffffu, deg);
    float4 acc[4] = {init4, init4, init4, init4};
    float4 tacc = init4;
    float4 v[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[u][i] = init4;
    for (int b = 0; b < nmax; b += 8 * kHold) {
      if (b > 0) { load_src(beg, deg, b, cur.src); load_attr(beg, deg, b, cur.e); }   // in-degree above 32: not prefetched
      // tail channels, slot-parallel: issue the narrow gathers now, use them after the first main group
      float4 tl[kHold];
#pragma unroll
      for (int h = 0; h < kHold; ++h)
        tl[h] = (b + 8 * h + q < deg) ? ld4(bt + static_cast<int64_t>(cur.src[h]) * 4) : init4;
#pragma unroll
      for (int hh = 0; hh < kHold; ++hh)
      for (int gg = 0; gg < 2; ++gg) {
        const int g = 2 * hh + gg;
        if (b + 4 * g >= nmax) break;   // warp-uniform (a later hh iteration breaks again right away)
        const int sreg = cur.src[hh];
        float ereg[DE];
#pragma unroll
        for (int d = 0; d < DE; ++d) ereg[d] = cur.e[hh][d];
        const int l0 = (4 * g) & 7;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int sidx = __shfl_sync(0xffffffffu, sreg, l0 + u, 8);
          const bool on = b + 4 * g + u < deg;
          const float* rp = bcol + static_cast<int64_t>(sidx) * kSplitMain;
          // max / min: a slot beyond the row's degree keeps the registers of an earlier slot of the same
          // row (its edge attributes are zero, so the value is re-submitted unchanged: harmless);
          // sums start every slot from zero
          if (kOrderFree) {
            if (on) {
#pragma unroll
              for (int i = 0; i < 4; ++i) v[u][i] = ld4(rp + 32 * i);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[u][i] = on ? ld4(rp + 32 * i) : init4;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int d = 0; d < DE; ++d) {
            const float ed = __shfl_sync(0xffffffffu, ereg[d], l0 + u, 8);   // 0 for slots beyond the row's degree
#pragma unroll
            for (int i = 0; i < 4; ++i) v[u][i] = fma4x2(ed, w[d][i], v[u][i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[i] = combine4x2<MODE>(acc[i], v[0][i], v[1][i]);
          acc[i] = combine4x2<MODE>(acc[i], v[2][i], v[3][i]);
        }
        if (g == 0) {
#pragma unroll
          for (int h = 0; h < kHold; ++h) {
            if (b + 8 * h + q < deg) {
              float4 t = tl[h];
#pragma unroll
              for (int d = 0; d < DE; ++d) t = fma4(cur.e[h][d], wt[d], t);
              tacc = combine4<MODE>(tacc, t);
            }
          }
        }
      }
    }
    const int deg_row = deg;
    beg = beg1; deg = deg1; beg1 = beg2; deg1 = deg2;
    // tail: fixed-order butterfly over the quarter's 8 lanes (= slots)
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      float4 other = tacc;
      other.x = __shfl_xor_sync(0xffffffffu, tacc.x, o);
      if (DE > 1) other.y = __shfl_xor_sync(0xffffffffu, tacc.y, o);
      if (DE > 2) other.z = __shfl_xor_sync(0xffffffffu, tacc.z, o);
      if (DE > 3) other.w = __shfl_xor_sync(0xffffffffu, tacc.w, o);
      tacc = combine4<MODE>(tacc, other);
    }
    if (!live) continue;
    float4 bmain[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) bmain[i] = ld4(&bs[32 * i + 4 * q]);
    const float4 btail = ld4(&bs[kSplitMain]);
    float4 r[4], rt = make_float4(0.f, 0.f, 0.f, 0.f);  // torch_scatter: empty segments aggregate to 0
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = rt;
    if (deg_row > 0) {
      if (MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN) {
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = add4(bmain[i], acc[i]);
        rt = add4(btail, tacc);
      } else if (MODE == RGNN_AGGR_ADD) {
        const float fd = static_cast<float>(deg_row);
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = fma4(fd, bmain[i], acc[i]);
        rt = fma4(fd, btail, tacc);
      } else {
        const float inv = 1.f / static_cast<float>(deg_row);
#pragma unroll
        for (int i = 0; i < 4; ++i) r[i] = fma4(inv, acc[i], bmain[i]);
        rt = fma4(inv, tacc, btail);
      }
    } else if (iso.w_t != nullptr) {
      // Folded formulation (node_gemm.cu): the update weights carry W_m W_t for every node, so a node
      // without incoming edge cancels that term with M' = -W_t x_n instead of 0 (rare: strided reads)
      float a16[4][4], t4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a16[i][j] = 0.f;
      const float* xr = iso.x + (iso.rows != nullptr ? static_cast<int64_t>(iso.rows[row]) : row) * iso.ldx;
      for (int c = 0; c < iso.c; ++c) {
        float xv = xr[c];
        if (iso.mean != nullptr) xv = (xv - iso.mean[c]) * iso.scale[c] + iso.beta[c];
        if (iso.relu) xv = fmaxf(xv, 0.f);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            a16[i][j] = fmaf(iso.w_t[static_cast<int64_t>(32 * i + 4 * q + j) * iso.ldw + c], xv, a16[i][j]);
        if (q == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (kSplitMain + j < p) t4[j] = fmaf(iso.w_t[static_cast<int64_t>(kSplitMain + j) * iso.ldw + c], xv, t4[j]);
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = make_float4(-a16[i][0], -a16[i][1], -a16[i][2], -a16[i][3]);
      rt = make_float4(-t4[0], -t4[1], -t4[2], -t4[3]);
    }
    // M' is written panel-major (tc_panel_offset): panel i of the row's 128-row tile, row slot row % 128, the
    // lane's 16-byte chunk at its swizzled position -- the eight lanes of the quarter fill one 128-byte line
    float* orow = out_m + (static_cast<int64_t>(row >> 7) * (kSplitMain / 32)) * 4096 + (row & 127) * 32 + ((q ^ (row & 7)) << 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(orow + i * 4096) = r[i];
    if (q == 0) *reinterpret_cast<float4*>(out_t + static_cast<int64_t>(row) * 4) = rt;
  }
}

template <int MODE>
int launch_edge_aggregate_split_mode(const float* bm, const float* bt, const ConvShape& s, const float* bias,
                                     const float* w_e, int64_t ldwe, const float* ea, const int32_t* csc_ptr,
                                     const int32_t* csc_src, int64_t n_nodes, float* out_m, float* out_t,
                                     cudaStream_t stream, const IsolatedNodeTerm& iso) {
  // one persistent CTA per SM over a contiguous node range (a multiple of the pass size)
  static int threads = 0;   // RGNN_AGG_THREADS = 384 | 512 (experiments)
  if (threads == 0) { const char* e = getenv("RGNN_AGG_THREADS"); threads = (e != nullptr && atoi(e) == 512) ? 512 : 384; }
  const int pass_rows = threads / 32 * 4;
  const int n = static_cast<int>(n_nodes);
  const int ctas = sm_count();
  // rows per CTA in units of one warp's 4 rows (not of whole passes): every SM gets the same share and the
  // last pass of a CTA is simply a short one (rounding up to whole passes left 9 of 148 SMs without work
  // at 100 k nodes)
  int rows_per_cta = static_cast<int>((n_nodes + ctas - 1) / ctas);
  rows_per_cta = (rows_per_cta + 3) / 4 * 4;
  (void)pass_rows;
  const unsigned blocks = div_up(n_nodes, rows_per_cta);
#define RGNN_SPLIT_CASE(DE_)                                                                                   \
  case DE_:                                                                                                    \
    if (threads == 512)                                                                                        \
      edge_aggregate_split_kernel<MODE, DE_, 512><<<blocks, 512, 0, stream>>>(bm, bt, s.p, bias, w_e, ldwe, ea, csc_ptr, \
                                                                              csc_src, n, rows_per_cta, out_m, out_t, iso); \
    else                                                                                                       \
      edge_aggregate_split_kernel<MODE, DE_, 384><<<blocks, 384, 0, stream>>>(bm, bt, s.p, bias, w_e, ldwe, ea, csc_ptr, \
                                                                              csc_src, n, rows_per_cta, out_m, out_t, iso); \
    break;
  switch (s.de) {
    RGNN_SPLIT_CASE(1) RGNN_SPLIT_CASE(2) RGNN_SPLIT_CASE(3) RGNN_SPLIT_CASE(4)
    default: return RGNN_ERR_UNSUPPORTED;
  }
#undef RGNN_SPLIT_CASE
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int launch_edge_aggregate_split(int aggr, const float* bm, const float* bt, const ConvShape& s, const float* bias,
                                const float* w_e, int64_t ldwe, const float* ea, const int32_t* csc_ptr,
                                const int32_t* csc_src, int64_t n_nodes, float* out_m, float* out_t,
                                cudaStream_t stream, const IsolatedNodeTerm& iso) {
  RGNN_PROFILE("edge_aggregate", stream);
  if (s.pm != kSplitMain || s.pt4 != 4) return RGNN_ERR_UNSUPPORTED;
  switch (aggr) {
    case RGNN_AGGR_MAX: return launch_edge_aggregate_split_mode<RGNN_AGGR_MAX>(bm, bt, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out_m, out_t, stream, iso);
    case RGNN_AGGR_MIN: return launch_edge_aggregate_split_mode<RGNN_AGGR_MIN>(bm, bt, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out_m, out_t, stream, iso);
    case RGNN_AGGR_ADD: return launch_edge_aggregate_split_mode<RGNN_AGGR_ADD>(bm, bt, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out_m, out_t, stream, iso);
    default: return launch_edge_aggregate_split_mode<RGNN_AGGR_MEAN>(bm, bt, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out_m, out_t, stream, iso);
  }
}

// general path: M[n] = reduce over the node's slots of U[slot] (already through pre_mlp)
template <int MODE>
__global__ void __launch_bounds__(kAggThreads)
segment_reduce_kernel(const float* __restrict__ u, int pp, const int32_t* __restrict__ csc_ptr, int64_t n_nodes,
                      float* __restrict__ out) {
  const int chunks = pp >> 2;
  const int64_t gid = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t node = gid / chunks;
  if (node >= n_nodes) return;
  const int c0 = static_cast<int>(gid - node * chunks) << 2;
  const int beg = csc_ptr[node], end = csc_ptr[node + 1];
  float4 acc;
  if (MODE == RGNN_AGGR_MAX) acc = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  else if (MODE == RGNN_AGGR_MIN) acc = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
  else acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int slot = beg; slot < end; ++slot) {
    const float4 v = ld4(u + static_cast<int64_t>(slot) * pp + c0);
    if (MODE == RGNN_AGGR_MAX) acc = max4(acc, v);
    else if (MODE == RGNN_AGGR_MIN) acc = min4(acc, v);
    else acc = add4(acc, v);
  }
  const int deg = end - beg;
  if (deg == 0) acc = make_float4(0.f, 0.f, 0.f, 0.f);
  else if (MODE == RGNN_AGGR_MEAN) {
    const float fd = static_cast<float>(deg);
    acc = make_float4(acc.x / fd, acc.y / fd, acc.z / fd, acc.w / fd);
  }
  *reinterpret_cast<float4*>(out + node * pp + c0) = acc;
}

// w_eff[p, de] = W_pre[:, off:off+C] . W_enc[C, de];  b_eff[p] = b_pre + W_pre[:, off:off+C] . b_enc
__global__ void fold_edge_encoder_kernel(const float* __restrict__ w_pre, int p, int off, int c,
                                         const float* __restrict__ w_enc, const float* __restrict__ b_enc, int de,
                                         const float* __restrict__ b_pre, float* __restrict__ w_eff,
                                         float* __restrict__ b_eff) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p * (de + 1)) return;
  const int row = idx / (de + 1), d = idx - row * (de + 1);
  const float* wr = w_pre + static_cast<int64_t>(row) * p + off;
  double acc = 0.0;
  if (d < de) {
    for (int j = 0; j < c; ++j) acc += static_cast<double>(wr[j]) * static_cast<double>(w_enc[j * de + d]);
    w_eff[row * de + d] = static_cast<float>(acc);
  } else {
    for (int j = 0; j < c; ++j) acc += static_cast<double>(wr[j]) * static_cast<double>(b_enc[j]);
    b_eff[row] = static_cast<float>(acc + static_cast<double>(b_pre[row]));
  }
}

__global__ void __launch_bounds__(256)
gather_edge_rows_kernel(const float* __restrict__ edge_attr, const int32_t* __restrict__ csc_eid, int64_t n_edges,
                        int de, float* __restrict__ ea_csc) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= n_edges * de) return;
  const int64_t slot = idx / de;
  const int d = static_cast<int>(idx - slot * de);
  ea_csc[idx] = edge_attr[static_cast<int64_t>(csc_eid[slot]) * de + d];
}

template <int MODE, int DE>
int launch_edge_aggregate_de(const float* a, const float* b, const ConvShape& s, const float* bias, const float* w_e,
                             int64_t ldwe, const float* ea, const int32_t* csc_ptr, const int32_t* csc_src,
                             int64_t n_nodes, float* out, cudaStream_t stream, const IsolatedNodeTerm& iso) {
  const int64_t threads = n_nodes * (s.pp >> 2);
  const size_t smem = DE == 0 ? sizeof(float) * s.de * s.pp : 0;
  if (smem > 48 * 1024)
    RGNN_CUDA_CHECK(cudaFuncSetAttribute(edge_aggregate_kernel<MODE, DE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem)));
  edge_aggregate_kernel<MODE, DE><<<div_up(threads, kAggThreads), kAggThreads, smem, stream>>>(
      a, b, s.pp, s.p, bias, w_e, ldwe, s.de, ea, csc_ptr, csc_src, n_nodes, out, iso);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

template <int MODE>
int launch_edge_aggregate(const float* a, const float* b, const ConvShape& s, const float* bias, const float* w_e,
                          int64_t ldwe, const float* ea, const int32_t* csc_ptr, const int32_t* csc_src,
                          int64_t n_nodes, float* out, cudaStream_t stream, IsolatedNodeTerm iso = IsolatedNodeTerm()) {
  RGNN_PROFILE("edge_aggregate", stream);
  // vector loads of the edge attributes need their natural alignment
  const bool ea_aligned = (reinterpret_cast<uintptr_t>(ea) % 16) == 0;
  switch (s.de) {
    case 1: return launch_edge_aggregate_de<MODE, 1>(a, b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out, stream, iso);
    case 2: if (ea_aligned) return launch_edge_aggregate_de<MODE, 2>(a, b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out, stream, iso); break;
    case 3: return launch_edge_aggregate_de<MODE, 3>(a, b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out, stream, iso);
    case 4: if (ea_aligned) return launch_edge_aggregate_de<MODE, 4>(a, b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out, stream, iso); break;
    default: break;
  }
  return launch_edge_aggregate_de<MODE, 0>(a, b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, out, stream, iso);
}

}  // namespace

// ---- tensor-core weight images -----------------------------------------------------------------
struct PackedLayout { size_t pre_off, post_off, fold_off, total; };

static PackedLayout packed_layout(const rgnn_conv_desc& d, const ConvShape& s) {
  PackedLayout pl{};
  size_t off = 0;
  pl.pre_off = off; off += align_up(sizeof(float) * tc_pack_floats(conv_pre_shape(s)));
  pl.post_off = off; off += align_up(sizeof(float) * tc_pack_floats(conv_post_shape(d, s)));
  pl.fold_off = off; off += align_up(sizeof(float) * static_cast<size_t>(s.c_out) * s.c);
  pl.total = off;
  return pl;
}

// W_s image for B = x W_s^T and the update image [W_x (+ W_m W_t) | W_m | (add: W_m W_t)]
static int pack_conv_weights(const rgnn_conv_desc& d, const ConvShape& s, float* wpack_pre, float* wpack_post,
                             float* w_fold, cudaStream_t stream) {
  const bool mpnn = d.conv_type == RGNN_CONV_MPNN;
  const int x_s_off = mpnn ? s.c : 0;
  TcWeightBlocks wb{};
  wb.count = 1;
  wb.block[0] = {d.pre_weight[0] + x_s_off, s.p, s.p, s.c, 0, 0};
  RGNN_RETURN_IF_ERROR(tc_pack_weights(wb, conv_pre_shape(s), wpack_pre, stream));
  const int64_t ldpost = s.c + s.p;
  int nb = 0;
  wb.block[nb++] = {d.post_weight[0], ldpost, s.c_out, s.c, 0, 0};
  int k_off = tc_seg_pad(s.c);  // segments start on 32-float panels
  if (s.split) {
    wb.block[nb++] = {d.post_weight[0] + s.c, ldpost, s.c_out, s.pm, k_off, 0};
    k_off += tc_seg_pad(s.pm);
    wb.block[nb++] = {d.post_weight[0] + s.c + s.pm, ldpost, s.c_out, s.p - s.pm, k_off, 0};
    k_off += tc_seg_pad(s.pt4);
  } else {
    wb.block[nb++] = {d.post_weight[0] + s.c, ldpost, s.c_out, s.p, k_off, 0};
    k_off += tc_seg_pad(s.pp);
  }
  if (mpnn) {
    RGNN_RETURN_IF_ERROR(tc_fold_weights(d.post_weight[0] + s.c, ldpost, d.pre_weight[0], s.p, s.c_out, s.p, s.c,
                                         w_fold, stream));
    if (d.aggr == RGNN_AGGR_ADD) {
      wb.block[nb++] = {w_fold, s.c, s.c_out, s.c, k_off, 0};   // deg * x segment
    } else {
      wb.block[nb++] = {w_fold, s.c, s.c_out, s.c, 0, 1};       // added onto W_x: (W_x + W_m W_t) x
    }
  }
  wb.count = nb;
  return tc_pack_weights(wb, conv_post_shape(d, s), wpack_post, stream);
}

int conv_shape(const rgnn_conv_desc& d, ConvShape* s) {
  if (d.in_channels < 1 || d.out_channels < 1 || d.edge_dim < 0) return RGNN_ERR_INVALID_ARGUMENT;
  if (d.pre_layers < 1 || d.pre_layers > RGNN_MAX_MLP_LAYERS || d.post_layers < 1 || d.post_layers > RGNN_MAX_MLP_LAYERS)
    return RGNN_ERR_INVALID_ARGUMENT;
  if (d.aggr < RGNN_AGGR_MAX || d.aggr > RGNN_AGGR_MIN) return RGNN_ERR_INVALID_ARGUMENT;
  s->c = d.in_channels;
  s->c_out = d.out_channels;
  s->de = d.edge_dim;
  if (d.conv_type == RGNN_CONV_MPNN) {
    s->de_eff = d.use_edge_encoder ? d.in_channels : d.edge_dim;
    s->p = 2 * d.in_channels + s->de_eff;
  } else if (d.conv_type == RGNN_CONV_RADAR_POINT_GNN) {
    if (d.use_edge_encoder) return RGNN_ERR_UNSUPPORTED;
    if (d.out_channels != d.in_channels) return RGNN_ERR_INVALID_ARGUMENT;  // residual update
    s->de_eff = d.edge_dim;
    s->p = d.in_channels + d.edge_dim;
  } else {
    return RGNN_ERR_INVALID_ARGUMENT;
  }
  s->pp = (s->p + 3) & ~3;
  s->general = d.pre_layers > 1;
  s->split = false; s->pm = 0; s->pt4 = 0;
  {
    // split message layout: main part = the node-feature columns of the message (2C or C), tail = the
    // edge-attribute columns; taken when the main rows are exactly one float4 per lane of a warp, the
    // edge weights fit the aggregate kernel's registers and both contractions run on the tensor cores
    const int main_cols = s->p - s->de_eff;
    if (!s->general && !d.use_edge_encoder && main_cols == 128 && s->de >= 1 && s->de <= 4) {
      ConvShape t = *s;
      t.split = true; t.pm = main_cols; t.pt4 = (s->p - main_cols + 3) & ~3;
      if (tc_gemm_supported(conv_pre_shape(t)) && tc_gemm_supported(conv_post_shape(d, t))) *s = t;
    }
  }
  for (int l = 0; l < d.pre_layers; ++l)
    if (d.pre_weight[l] == nullptr || d.pre_bias[l] == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  for (int l = 0; l < d.post_layers; ++l)
    if (d.post_weight[l] == nullptr || d.post_bias[l] == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (d.use_edge_encoder && (d.edge_encoder_weight == nullptr || d.edge_encoder_bias == nullptr))
    return RGNN_ERR_INVALID_ARGUMENT;
  return RGNN_OK;
}

int gather_edge_rows(const float* edge_attr, const int32_t* csc_eid, int64_t n_edges, int32_t de, float* ea_csc,
                     cudaStream_t stream) {
  if (n_edges == 0 || de == 0) return RGNN_OK;
  RGNN_PROFILE("gather_edge_rows", stream);
  gather_edge_rows_kernel<<<div_up(n_edges * de, 256), 256, 0, stream>>>(edge_attr, csc_eid, n_edges, de, ea_csc);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int conv_forward(const rgnn_conv_desc& d, const ConvShape& s, const ConvInput& in, int64_t n_nodes,
                 const int32_t* csc_ptr, const int32_t* csc_src, const int32_t* csc_eid, const float* edge_attr,
                 int64_t n_edges, float* out, const ConvWorkspace& w, cudaStream_t stream, int64_t* bn_partials) {
  if (bn_partials != nullptr) *bn_partials = 0;
  if (n_nodes == 0) return RGNN_OK;
  const bool mpnn = d.conv_type == RGNN_CONV_MPNN;
  const int x_s_off = mpnn ? s.c : 0;            // column block of x_j in W_pre
  const int e_off = mpnn ? 2 * s.c : s.c;        // column block of the edge attributes

  // edge attributes in CSC slot order
  const float* ea = edge_attr;
  if (csc_eid != nullptr && s.de > 0) {
    RGNN_RETURN_IF_ERROR(gather_edge_rows(edge_attr, csc_eid, n_edges, s.de, w.ea_csc, stream));
    ea = w.ea_csc;
  }

  // edge-term weights: W_e [p, de] with row stride ldwe, bias b [p]
  const float* w_e = d.pre_weight[0] + e_off;
  int64_t ldwe = s.p;
  const float* bias = d.pre_bias[0];
  if (d.use_edge_encoder) {
    fold_edge_encoder_kernel<<<div_up(s.p * (s.de + 1), 128), 128, 0, stream>>>(
        d.pre_weight[0], s.p, e_off, s.c, d.edge_encoder_weight, d.edge_encoder_bias, s.de, d.pre_bias[0],
        w.w_eff, w.b_eff);
    RGNN_LAUNCH_CHECK();
    w_e = w.w_eff; ldwe = s.de; bias = w.b_eff;
  }

  // ---- tensor-core path: B = x W_s^T, fold W_t through the update, fused BN column sums ---------
  const bool aligned = (in.ldx % 4 == 0) && (reinterpret_cast<uintptr_t>(in.x) % 16 == 0) &&
                       (in.mean == nullptr || reinterpret_cast<uintptr_t>(in.mean) % 16 == 0);
  if (s.split && !(w.tc_post && aligned)) return RGNN_ERR_UNSUPPORTED;  // split buffers only fit the tensor-core path
  if (w.tc_post && aligned) {
    // weight images: the caller's pre-packed buffer, or (weights may have changed since the last call)
    // repacked into the workspace on every forward
    const float* wpack_pre = w.wpack_pre;
    const float* wpack_post = w.wpack_post;
    const bool third_segment = mpnn && d.aggr == RGNN_AGGR_ADD;
    if (d.packed_weights != nullptr) {
      const PackedLayout pl = packed_layout(d, s);
      const char* base = static_cast<const char*>(d.packed_weights);
      wpack_pre = reinterpret_cast<const float*>(base + pl.pre_off);
      wpack_post = reinterpret_cast<const float*>(base + pl.post_off);
    } else {
      RGNN_RETURN_IF_ERROR(pack_conv_weights(d, s, w.wpack_pre, w.wpack_post, w.w_fold, stream));
    }

    TcGemmParams g1;
    g1.a1 = in.x; g1.lda1 = in.ldx; g1.k1 = s.c; g1.a1_rows = in.rows;
    g1.a1_mean = in.mean; g1.a1_scale = in.scale; g1.a1_beta = in.beta; g1.relu_a1 = in.relu;
    g1.wpack = wpack_pre; g1.n = s.p; g1.n_store = s.pp;
    g1.y = w.b; g1.ldy = s.pp; g1.m = n_nodes;
    if (s.split) { g1.ldy = s.pm; g1.y2 = w.bt; g1.ldy2 = s.pt4; g1.n_split = s.pm; g1.n_store = s.pm + s.pt4; }
    RGNN_RETURN_IF_ERROR(launch_tc_gemm(g1, "node_gemm_pre", stream));

    // M'_n = b + max_e (...) (mean: b + mean; add: deg b + sum): the W_t x_t term is folded into the
    // update weights.  max / min / mean carry it once per node with an incoming edge, so isolated
    // nodes cancel it (IsolatedNodeTerm); add carries it deg times (third K segment of the update).
    IsolatedNodeTerm iso;
    if (mpnn && d.aggr != RGNN_AGGR_ADD) {
      iso.w_t = d.pre_weight[0]; iso.ldw = s.p; iso.x = in.x; iso.ldx = in.ldx; iso.c = s.c; iso.rows = in.rows;
      iso.mean = in.mean; iso.scale = in.scale; iso.beta = in.beta; iso.relu = in.relu;
    }
    if (fused_layer_supported(d, s)) {
      // aggregate + node update in one kernel: M' stays in shared / tensor memory (fused_layer.cu)
      FusedLayerArgs f;
      f.aggr = d.aggr; f.bm = w.b; f.bt = w.bt; f.mt = w.mt; f.p = s.p; f.de = s.de;
      f.bias_msg = bias; f.w_e = w_e; f.ldwe = ldwe; f.ea = ea; f.csc_ptr = csc_ptr; f.csc_src = csc_src; f.iso = iso;
      f.x = in.x; f.ldx = in.ldx; f.x_rows = in.rows;
      f.x_mean = in.mean; f.x_scale = in.scale; f.x_beta = in.beta; f.relu_x = in.relu;
      f.wpack = wpack_post; f.w_tail = d.post_weight[0] + s.c + s.pm; f.ld_wtail = s.c + s.p;
      f.bias_post = d.post_bias[0]; f.c_out = s.c_out; f.y = out; f.ldy = s.c_out; f.n_nodes = n_nodes;
      f.status = w.tc_status;
      if (bn_partials != nullptr) { f.bn_partial = w.bn_partial; *bn_partials = fused_layer_partials(n_nodes); }
      return launch_fused_layer(f, stream);
    }
    if (s.split) {
      RGNN_RETURN_IF_ERROR(launch_edge_aggregate_split(d.aggr, w.b, w.bt, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes,
                                                       w.m, w.mt, stream, iso));
    } else
    switch (d.aggr) {
      case RGNN_AGGR_MAX: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MAX>(nullptr, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream, iso)); break;
      case RGNN_AGGR_MIN: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MIN>(nullptr, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream, iso)); break;
      case RGNN_AGGR_ADD: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_ADD>(nullptr, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream, iso)); break;
      default: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MEAN>(nullptr, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream, iso)); break;
    }

    float* first_out = d.post_layers == 1 ? out : w.t1;
    TcGemmParams g2;
    g2.a1 = in.x; g2.lda1 = in.ldx; g2.k1 = s.c; g2.a1_rows = in.rows;
    g2.a1_mean = in.mean; g2.a1_scale = in.scale; g2.a1_beta = in.beta; g2.relu_a1 = in.relu;
    g2.a2 = w.m; g2.lda2 = s.pp; g2.k2 = s.pp;
    if (s.split) { g2.lda2 = s.pm; g2.k2 = s.pm; g2.a2_panel_major = 1; g2.at = w.mt; g2.ldat = s.pt4; g2.kt = s.pt4; }
    if (third_segment) { g2.k3 = s.c; g2.csc_ptr = csc_ptr; g2.rowscale_mode = 2; }
    g2.wpack = wpack_post; g2.n = s.c_out; g2.n_store = s.c_out; g2.bias = d.post_bias[0];
    g2.y = first_out; g2.ldy = s.c_out; g2.m = n_nodes;
    if (!mpnn && d.post_layers == 1) {
      g2.residual = in.x; g2.ldr = in.ldx;
      g2.res_mean = in.mean; g2.res_scale = in.scale; g2.res_beta = in.beta; g2.res_relu = in.relu;
    }
    if (d.post_layers == 1 && bn_partials != nullptr) {
      g2.bn_partial = w.bn_partial;
      *bn_partials = tc_tiles(n_nodes);
    }
    RGNN_RETURN_IF_ERROR(launch_tc_gemm(g2, "node_gemm_post", stream));
    float* cur = first_out;
    for (int l = 1; l < d.post_layers; ++l) {
      const bool last = l == d.post_layers - 1;
      float* nxt = last ? out : (cur == w.t1 ? w.t2 : w.t1);
      LinearArgs lq;
      lq.a1 = cur; lq.lda1 = s.c_out; lq.k1 = s.c_out; lq.relu_a1 = 1;
      lq.w = d.post_weight[l]; lq.ldw = s.c_out; lq.bias = d.post_bias[l];
      lq.y = nxt; lq.ldy = s.c_out; lq.m = n_nodes; lq.n = s.c_out; lq.tag = "linear_post";
      if (!mpnn && last) {
        lq.residual = in.x; lq.ldr = in.ldx; lq.res_rows = in.rows;
        lq.res_mean = in.mean; lq.res_scale = in.scale; lq.res_beta = in.beta; lq.res_relu = in.relu;
      }
      RGNN_RETURN_IF_ERROR(launch_linear(lq, stream));
      cur = nxt;
    }
    return RGNN_OK;
  }

  // node-level halves of the first message Linear
  LinearArgs la;
  la.a1 = in.x; la.lda1 = in.ldx; la.k1 = s.c; la.a1_rows = in.rows;
  la.a1_mean = in.mean; la.a1_scale = in.scale; la.a1_beta = in.beta; la.relu_a1 = in.relu;
  la.ldw = s.p; la.m = n_nodes; la.n = s.p; la.ldy = s.pp;
  la.tag = "linear_pre_node";
  if (mpnn) {
    la.w = d.pre_weight[0]; la.y = w.a;
    RGNN_RETURN_IF_ERROR(launch_linear(la, stream));
  }
  la.w = d.pre_weight[0] + x_s_off; la.y = w.b;
  RGNN_RETURN_IF_ERROR(launch_linear(la, stream));

  const float* a_term = mpnn ? w.a : nullptr;
  if (!s.general) {
    switch (d.aggr) {
      case RGNN_AGGR_MAX: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MAX>(a_term, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream)); break;
      case RGNN_AGGR_MIN: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MIN>(a_term, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream)); break;
      case RGNN_AGGR_ADD: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_ADD>(a_term, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream)); break;
      default: RGNN_RETURN_IF_ERROR(launch_edge_aggregate<RGNN_AGGR_MEAN>(a_term, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.m, stream)); break;
    }
  } else {
    // per-edge activations through the remaining Linear layers, then the segmented reduce
    RGNN_RETURN_IF_ERROR(launch_edge_aggregate<-1>(a_term, w.b, s, bias, w_e, ldwe, ea, csc_ptr, csc_src, n_nodes, w.u1, stream));
    float* cur = w.u1;
    float* nxt = w.u2;
    for (int l = 1; l < d.pre_layers; ++l) {
      LinearArgs lu;
      lu.a1 = cur; lu.lda1 = s.pp; lu.k1 = s.p; lu.relu_a1 = 1;
      lu.w = d.pre_weight[l]; lu.ldw = s.p; lu.bias = d.pre_bias[l];
      lu.y = nxt; lu.ldy = s.pp; lu.m = n_edges; lu.n = s.p;
      lu.tag = "linear_pre_edge";
      RGNN_RETURN_IF_ERROR(launch_linear(lu, stream));
      float* t = cur; cur = nxt; nxt = t;
    }
    const int64_t threads = n_nodes * (s.pp >> 2);
    const unsigned blocks = div_up(threads, kAggThreads);
    switch (d.aggr) {
      case RGNN_AGGR_MAX: segment_reduce_kernel<RGNN_AGGR_MAX><<<blocks, kAggThreads, 0, stream>>>(cur, s.pp, csc_ptr, n_nodes, w.m); break;
      case RGNN_AGGR_MIN: segment_reduce_kernel<RGNN_AGGR_MIN><<<blocks, kAggThreads, 0, stream>>>(cur, s.pp, csc_ptr, n_nodes, w.m); break;
      case RGNN_AGGR_ADD: segment_reduce_kernel<RGNN_AGGR_ADD><<<blocks, kAggThreads, 0, stream>>>(cur, s.pp, csc_ptr, n_nodes, w.m); break;
      default: segment_reduce_kernel<RGNN_AGGR_MEAN><<<blocks, kAggThreads, 0, stream>>>(cur, s.pp, csc_ptr, n_nodes, w.m); break;
    }
    RGNN_LAUNCH_CHECK();
  }

  // node update: post_mlp([x ; M]) (+ x for RadarPointGNNConv)
  float* cur_out = d.post_layers == 1 ? out : w.t1;
  LinearArgs lp;
  lp.a1 = in.x; lp.lda1 = in.ldx; lp.k1 = s.c; lp.a1_rows = in.rows;
  lp.a1_mean = in.mean; lp.a1_scale = in.scale; lp.a1_beta = in.beta; lp.relu_a1 = in.relu;
  lp.a2 = w.m; lp.lda2 = s.pp; lp.k2 = s.p;
  lp.w = d.post_weight[0]; lp.ldw = s.c + s.p; lp.bias = d.post_bias[0];
  lp.y = cur_out; lp.ldy = s.c_out; lp.m = n_nodes; lp.n = s.c_out;
  lp.tag = "linear_post";
  if (!mpnn && d.post_layers == 1) {
    lp.residual = in.x; lp.ldr = in.ldx; lp.res_rows = in.rows;
    lp.res_mean = in.mean; lp.res_scale = in.scale; lp.res_beta = in.beta; lp.res_relu = in.relu;
  }
  RGNN_RETURN_IF_ERROR(launch_linear(lp, stream));
  for (int l = 1; l < d.post_layers; ++l) {
    const bool last = l == d.post_layers - 1;
    float* nxt = last ? out : (cur_out == w.t1 ? w.t2 : w.t1);
    LinearArgs lq;
    lq.a1 = cur_out; lq.lda1 = s.c_out; lq.k1 = s.c_out; lq.relu_a1 = 1;
    lq.w = d.post_weight[l]; lq.ldw = s.c_out; lq.bias = d.post_bias[l];
    lq.y = nxt; lq.ldy = s.c_out; lq.m = n_nodes; lq.n = s.c_out;
    if (!mpnn && last) {
      // residual is the (normalised) layer input, not the intermediate activation
      lq.residual = in.x; lq.ldr = in.ldx; lq.res_rows = in.rows;
      lq.res_mean = in.mean; lq.res_scale = in.scale; lq.res_beta = in.beta; lq.res_relu = in.relu;
    }
    RGNN_RETURN_IF_ERROR(launch_linear(lq, stream));
    cur_out = nxt;
  }
  return RGNN_OK;
}

}  // namespace rgnn

using namespace rgnn;

extern "C" {

size_t rgnn_conv_packed_bytes(const rgnn_conv_desc* desc) {
  if (desc == nullptr) return 0;
  rgnn_conv_desc d = *desc;
  static const float dummy = 0.f;
  for (int l = 0; l < RGNN_MAX_MLP_LAYERS; ++l) d.pre_weight[l] = d.pre_bias[l] = d.post_weight[l] = d.post_bias[l] = &dummy;
  d.edge_encoder_weight = d.edge_encoder_bias = &dummy;
  ConvShape s;
  if (conv_shape(d, &s) != RGNN_OK || s.general) return 0;
  if (!tc_gemm_supported(conv_pre_shape(s)) || !tc_gemm_supported(conv_post_shape(d, s))) return 0;
  return packed_layout(d, s).total;
}

int rgnn_conv_pack_weights(const rgnn_conv_desc* desc, void* packed, size_t packed_bytes, rgnn_stream_t stream) {
  if (desc == nullptr || packed == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  ConvShape s;
  RGNN_RETURN_IF_ERROR(conv_shape(*desc, &s));
  const size_t need = rgnn_conv_packed_bytes(desc);
  if (need == 0) return RGNN_ERR_UNSUPPORTED;
  if (packed_bytes < need || reinterpret_cast<uintptr_t>(packed) % 16 != 0) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  const PackedLayout pl = packed_layout(*desc, s);
  char* base = static_cast<char*>(packed);
  return pack_conv_weights(*desc, s, reinterpret_cast<float*>(base + pl.pre_off), reinterpret_cast<float*>(base + pl.post_off),
                           reinterpret_cast<float*>(base + pl.fold_off), static_cast<cudaStream_t>(stream));
}

size_t rgnn_conv_workspace_bytes(const rgnn_conv_desc* desc, int64_t n_nodes, int64_t n_edges) {
  ConvShape s;
  if (desc == nullptr || n_nodes < 0 || n_edges < 0) return 0;
  // sizing only needs the dimensions: tolerate null weight pointers here
  rgnn_conv_desc d = *desc;
  static const float dummy = 0.f;
  for (int l = 0; l < RGNN_MAX_MLP_LAYERS; ++l) {
    d.pre_weight[l] = d.pre_bias[l] = d.post_weight[l] = d.post_bias[l] = &dummy;
  }
  d.edge_encoder_weight = d.edge_encoder_bias = &dummy;
  if (conv_shape(d, &s) != RGNN_OK) return 0;
  SizeArena a;
  carve_conv_workspace(a, d, s, n_nodes, n_edges, true);
  return a.used;
}

int rgnn_conv_forward(const rgnn_conv_desc* desc, const float* x, int64_t n_nodes, const int32_t* csc_ptr,
                      const int32_t* csc_src, const int32_t* csc_eid, const float* edge_attr, int64_t n_edges,
                      float* out, void* workspace, size_t workspace_bytes, rgnn_stream_t stream) {
  if (desc == nullptr || n_nodes < 0 || n_edges < 0) return RGNN_ERR_INVALID_ARGUMENT;
  ConvShape s;
  RGNN_RETURN_IF_ERROR(conv_shape(*desc, &s));
  if (n_nodes == 0) return RGNN_OK;
  if (x == nullptr || out == nullptr || csc_ptr == nullptr) return RGNN_ERR_INVALID_ARGUMENT;
  if (n_edges > 0 && (csc_src == nullptr || (s.de > 0 && edge_attr == nullptr))) return RGNN_ERR_INVALID_ARGUMENT;
  if (workspace == nullptr || workspace_bytes < rgnn_conv_workspace_bytes(desc, n_nodes, n_edges)) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  Arena arena(workspace, workspace_bytes);
  ConvWorkspace w = carve_conv_workspace(arena, *desc, s, n_nodes, n_edges, true);
  if (arena.overflow) return RGNN_ERR_WORKSPACE_TOO_SMALL;
  ConvInput in;
  in.x = x; in.ldx = s.c;
  return conv_forward(*desc, s, in, n_nodes, csc_ptr, csc_src, csc_eid, edge_attr, n_edges, out, w,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"

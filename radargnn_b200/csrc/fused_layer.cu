// fused_layer.cu -- edge aggregate + node update of one MPNNConv layer in ONE kernel
// (reference gnn/mpnn_layers.py:86-92: propagate = gather -> message Linear -> scatter-reduce at
// edge_index[1], then update = post_mlp(cat([x, aggregated]))).
//
// The two-kernel path (conv.cu edge_aggregate_split_kernel -> node_gemm.cu) writes the aggregated
// messages M' [N, 132] to HBM and reads them back: 2 x 52.8 MB per layer at the headline size and a whole
// kernel (46 us) whose only job is to stream them.  Here M' never leaves the SM:
//
//   warps 0-11   AGGREGATE.  The segmented gather-reduce of edge_aggregate_split_kernel, unchanged in its
//                inner loop (a quarter-warp per target node, 16 x 16-byte gathers per lane in flight, packed
//                FFMA2 edge term, 3-input min / max): a warp takes units of 4 consecutive rows, 12 units =
//                48 rows are in progress per CTA.  A finished row (128 main channels) is written into a
//                shared-memory ring of 32-row quarters (16 KB each, the panel layout of the contraction:
//                per 32-float K block one [32 x 128 B] panel, 16-byte chunk c of row r at c ^ (r % 8)).
//   warps 12-15  TILE warps, thread = row of the current 128-row tile = TMEM lane:
//                  x row (global, BatchNorm + ReLU of the previous layer on load)   -> hi / lo -> TMEM
//                  tail channels of M' (the De edge-attribute columns; gather + reduce per row, in registers)
//                  M' quarter (ring, when its 8 units have arrived)                  -> hi / lo -> TMEM
//                  [warp 12, one lane] tcgen05.mma kind::tf32, A from TMEM, W resident in shared memory,
//                     3xTF32 with hi*hi and hi*lo as one MMA against [W_hi ; W_lo] (N = 2 np)
//                  epilogue: tcgen05.ld, + bias + rank-De update with the tail channels, transpose through
//                     shared memory, coalesced row stores of h, BatchNorm column sums (fixed order, fp64
//                     across tiles; ONE partial per CTA)
// The tile warps are far from busy (a tile costs them ~15 k cycles, the aggregate warps need ~29 k to
// produce it), so they run the steps of a tile one after the other and need no second accumulator or
// operand stage; the ring decouples the two sides by up to one tile.
//
// A CTA owns a CONTIGUOUS range of (cell-sorted) rows, like the stand-alone aggregate kernel: what an SM
// gathers concurrently comes from the same few grid rows (L1 hits).  Tiles start at the range start, not
// at multiples of 128.
//
// Supported: MPNNConv, C = 64 (message main part = 128 channels), De in 1..4, max / min / mean, one update
// Linear, c_out <= 64.  Everything else takes the two-kernel path.
#include <math.h>
#include <stdlib.h>

#include "conv.cuh"
#include "tc_common.cuh"

namespace rgnn {
namespace {

using namespace tc;

constexpr int kAggWarps = 12;
constexpr int kAggThreads = kAggWarps * 32;
constexpr int kTileThreads = 128;
constexpr int kFusedThreads = kAggThreads + kTileThreads;   // 512
constexpr int kPassRows = kAggWarps * 4;                    // rows in progress per CTA
constexpr int kRing = 4;                                    // 32-row quarters of M' in shared memory
constexpr int kMain = 128;                                  // main message channels (2 C)
constexpr int kC = 64;
constexpr int kQuarterFloats = 4 * 32 * 32;                 // 4 panels x 32 rows x 32 floats
constexpr int kKBlocks = (kC + kMain) / 32;                 // K blocks of the contraction: x (2) + M' main (4)
constexpr int kStageFloats = 32 * 36;                       // per tile warp: transpose buffer
constexpr int kHold = 4;                                    // slots held per lane: a quarter-warp holds 32 slots of its row
// register split (setmaxnreg): 3 aggregate warpgroups + 1 tile warpgroup share 4 x 128 registers per thread slot
constexpr int kAggRegs = 144, kTileRegs = 80;
constexpr int kLastTileFullDefault = 1;                     // see vbeg in the kernel
constexpr uint32_t kXCol = 128, kMCol = 256;                // TMEM columns: accumulator 0.., x stages, M' stages

struct FusedParams {
  const float* bm; const float* mt; int32_t p;   // B main [N, 128]; reduced tail channels of M' [N, 4]
  const float* bias_msg; const float* w_e; int64_t ldwe; const float* ea;
  const int32_t* csc_ptr; const int32_t* csc_src;
  IsolatedNodeTerm iso;
  const float* x; int64_t ldx; const int32_t* x_rows;
  const float* x_mean; const float* x_scale; const float* x_beta; int32_t relu_x;
  const float* wpack; const float* w_tail; int64_t ld_wtail; const float* bias_post;
  int32_t c_out, np;
  float* y; int64_t ldy;
  double* bn_partial; int32_t n_partials;
  int32_t* status;
  int32_t n_nodes, rows_per_cta, tiles_per_cta;
  long long* trace; int32_t trace_cta;   // debug timeline of one CTA (scripts/trace_fused.py)
  int32_t last_tile_full;                // tile layout: 1 = the partial tile comes first (RGNN_FUSED_LAST_TILE_FULL)
};

struct FusedSmem {
  float* w;        // [kKBlocks][hi np x 32 | lo np x 32], swizzled panels (tc_pack_weights image prefix)
  float* ring;     // [kRing][4 panels][32 rows][32 floats]
  float* stage;    // [4 tile warps][32][36]
  float* ws;       // [DE][kMain] edge weights of the main channels, transposed
  float* bs;       // [kMain] message bias
  float* wtail;    // [4 tail channels][64] update weights of the tail channels
  float* bias;     // [64] update bias
  float* bn;       // [3][64] BatchNorm-on-load of x
  float* colsum;   // [4 quarters][64]
  float* colsq;
  uint64_t* bar;   // full[kRing], (unused)[kRing], a_full, acc_full
  int32_t* consumed;  // [kRing] uses of a ring quarter the tile warp has drained so far (monotonic)
  uint32_t* tmem_base;
};

__host__ __device__ inline size_t fused_smem_floats(int np, int de) {
  return static_cast<size_t>(kKBlocks) * 2 * np * 32 + kRing * kQuarterFloats + 4 * kStageFloats + de * kMain + kMain +
         4 * 64 + 64 + 3 * 64 + 2 * 4 * 64 + 2 * (2 * kRing + 2) + kRing + 4;
}

__device__ __forceinline__ FusedSmem carve_fused(unsigned char* base, int np, int de) {
  FusedSmem s;
  float* f = reinterpret_cast<float*>(base);
  s.w = f; f += static_cast<size_t>(kKBlocks) * 2 * np * 32;
  s.ring = f; f += kRing * kQuarterFloats;
  s.stage = f; f += 4 * kStageFloats;
  s.ws = f; f += de * kMain;
  s.bs = f; f += kMain;
  s.wtail = f; f += 4 * 64;
  s.bias = f; f += 64;
  s.bn = f; f += 3 * 64;
  s.colsum = f; f += 4 * 64;
  s.colsq = f; f += 4 * 64;
  s.bar = reinterpret_cast<uint64_t*>(f); f += 2 * (2 * kRing + 2);
  s.consumed = reinterpret_cast<int32_t*>(f); f += kRing;
  s.tmem_base = reinterpret_cast<uint32_t*>(f);
  return s;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 fma4(float s, float4 w, float4 a) {
  return make_float4(fmaf(s, w.x, a.x), fmaf(s, w.y, a.y), fmaf(s, w.z, a.z), fmaf(s, w.w, a.w));
}
__device__ __forceinline__ float4 fma4x2(float s, float4 w, float4 a) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __ffma2_rn(ss, make_float2(w.x, w.y), make_float2(a.x, a.y));
  const float2 hi = __ffma2_rn(ss, make_float2(w.z, w.w), make_float2(a.z, a.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <int MODE>
__device__ __forceinline__ float4 combine4(float4 a, float4 b) {
  if (MODE == RGNN_AGGR_MAX) return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
  if (MODE == RGNN_AGGR_MIN) return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w));
  return add4(a, b);
}
// acc (op) v0 (op) v1, in this order (sums stay in slot order)
template <int MODE>
__device__ __forceinline__ float4 combine4x2(float4 a, float4 v0, float4 v1) {
  if (MODE == RGNN_AGGR_MAX)
    return make_float4(fmaxf(fmaxf(a.x, v0.x), v1.x), fmaxf(fmaxf(a.y, v0.y), v1.y), fmaxf(fmaxf(a.z, v0.z), v1.z), fmaxf(fmaxf(a.w, v0.w), v1.w));
  if (MODE == RGNN_AGGR_MIN)
    return make_float4(fminf(fminf(a.x, v0.x), v1.x), fminf(fminf(a.y, v0.y), v1.y), fminf(fminf(a.z, v0.z), v1.z), fminf(fminf(a.w, v0.w), v1.w));
  return add4(add4(a, v0), v1);
}

// A quarter of the ring goes back to the aggregate warps through a MONOTONIC counter, not an mbarrier: a warp
// owns 4 of every 12 units, so it skips every third quarter, and a parity wait cannot tell "two uses behind"
// from "done" (a warp racing ahead over empty rows at the end of a CTA's range overwrote live quarters).
__device__ __forceinline__ int ld_acquire_shared(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_shared(int32_t* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ bool wait_consumed(const int32_t* p, int need) {
  for (int it = 0; it < (1 << 24); ++it) {
    if (ld_acquire_shared(p) >= need) return true;
    __nanosleep(100);
  }
  return false;
}

// 32 fp32 values of this thread's row (8 x float4) -> hi / lo images -> 64 TMEM columns at `taddr`
__device__ __forceinline__ void store_panel_hi_lo(uint32_t taddr, const float4 (&v)[8]) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float t[16];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float4 x = v[h * 4 + jj];
      t[jj * 4 + 0] = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); t[jj * 4 + 1] = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
      t[jj * 4 + 2] = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); t[jj * 4 + 3] = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
    }
    tmem_st16(taddr + h * 16, t);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {   // lo parts = x - hi (exact)
      const float4 x = v[h * 4 + jj];
      t[jj * 4 + 0] = x.x - t[jj * 4 + 0]; t[jj * 4 + 1] = x.y - t[jj * 4 + 1];
      t[jj * 4 + 2] = x.z - t[jj * 4 + 2]; t[jj * 4 + 3] = x.w - t[jj * 4 + 3];
    }
    tmem_st16(taddr + 32 + h * 16, t);
  }
}

template <int MODE, int DE>
__global__ void __launch_bounds__(kFusedThreads, 1)
fused_layer_kernel(const __grid_constant__ FusedParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int np = p.np;
  const FusedSmem s = carve_fused(smem_raw, np, DE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint64_t* full = s.bar;               // [kRing] a quarter of M' is complete (8 units x 32 lanes)
  uint64_t* a_full = s.bar + 2 * kRing; // the tile's operands are in TMEM (128 threads)
  uint64_t* acc_full = a_full + 1;      // the tile's MMAs have completed (tcgen05.commit)

  // ---- setup: small tables -> shared memory, barriers, TMEM -------------------------------------------
  for (int i = tid; i < DE * kMain; i += kFusedThreads) {
    const int d = i / kMain, ch = i - d * kMain;
    s.ws[i] = p.w_e[static_cast<int64_t>(ch) * p.ldwe + d];
  }
  for (int i = tid; i < kMain; i += kFusedThreads) s.bs[i] = p.bias_msg[i];
  for (int i = tid; i < 4 * 64; i += kFusedThreads) {
    const int j = i >> 6, n = i & 63;
    s.wtail[i] = (kMain + j < p.p && n < p.c_out) ? p.w_tail[static_cast<int64_t>(n) * p.ld_wtail + j] : 0.f;
  }
  for (int i = tid; i < 64; i += kFusedThreads) {
    s.bias[i] = i < p.c_out ? p.bias_post[i] : 0.f;
    if (p.x_mean != nullptr) { s.bn[i] = p.x_mean[i]; s.bn[64 + i] = p.x_scale[i]; s.bn[128 + i] = p.x_beta[i]; }
  }
  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) { mbar_init(&full[i], 8 * 32); s.consumed[i] = 0; }
    mbar_init(a_full, kTileThreads);
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kAggWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *s.tmem_base;

  if (p.trace != nullptr && blockIdx.x == p.trace_cta && tid == 0) p.trace[1000] = clock64();
  const int row_begin = blockIdx.x * p.rows_per_cta;
  const int row_end = min(p.n_nodes, row_begin + p.rows_per_cta);
  // Tiles are laid so that the LAST one ends at row_end and the FIRST one is the partial one (its rows below row_begin
  // are dead): when the aggregate warps finish, the tile warps have a full tile's production time behind them and only
  // the last quarter's conversion, the MMAs and one epilogue remain exposed (a short last tile left ~1.5 tiles).
  const int vbeg = p.last_tile_full ? row_end - p.tiles_per_cta * kRows : row_begin;   // may be negative: every access checks row >= row_begin
  const int tiles = p.tiles_per_cta;
  bool timed_out = false;

  if (warp < kAggWarps) {
    // =========================== aggregate warps ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kAggRegs));
    const int q = lane & 7, quarter = lane >> 3;
    // edge weights of the lane's 16 channels: in registers for De <= 2 (32 registers); wider edge attributes would
    // spill at 144 registers (De = 4: 64 registers, 1 KB of spills, 2x the time per edge), so their weights are read
    // from shared memory one attribute at a time (4 LDS.128 per attribute and slot group)
    constexpr bool kWeightsInRegs = DE <= 2;
    float4 w[kWeightsInRegs ? DE : 1][4];
    if (kWeightsInRegs) {
#pragma unroll
      for (int d = 0; d < DE; ++d)
#pragma unroll
        for (int i = 0; i < 4; ++i) w[d][i] = ld4(&s.ws[d * kMain + 32 * i + 4 * q]);
    }
    constexpr float kInit = MODE == RGNN_AGGR_MAX ? -INFINITY : (MODE == RGNN_AGGR_MIN ? INFINITY : 0.f);
    const float4 init4 = make_float4(kInit, kInit, kInit, kInit);
    constexpr bool kOrderFree = MODE == RGNN_AGGR_MAX || MODE == RGNN_AGGR_MIN;
    const float* bcol = p.bm + 4 * q;
    const int units = tiles * 32;                                  // 4-row units of this CTA (whole tiles)
    const int kPasses = (units + kAggWarps - 1) / kAggWarps;
    const int row_base = vbeg + warp * 4 + quarter;

    // software pipeline over the passes: row pointers two passes ahead, the first 32 slot sources one pass ahead
    auto load_ptr = [&](int pass, int& beg, int& deg) {
      const int row = row_base + pass * kPassRows;
      beg = 0; deg = 0;
      if (pass < kPasses && row >= row_begin && row < row_end) { beg = p.csc_ptr[row]; deg = p.csc_ptr[row + 1] - beg; }
    };
    auto load_src = [&](int beg, int deg, int b, int (&src)[kHold]) {
#pragma unroll
      for (int h = 0; h < kHold; ++h) src[h] = (b + 8 * h + q < deg) ? p.csc_src[beg + b + 8 * h + q] : 0;
    };
    auto load_attr = [&](int beg, int deg, int b, float (&e)[kHold][DE]) {
#pragma unroll
      for (int h = 0; h < kHold; ++h) {
        const bool on = b + 8 * h + q < deg;
        const float* ep = p.ea + static_cast<int64_t>(beg + b + 8 * h + q) * DE;
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
      const int unit = warp + it * kAggWarps;
      long long* tr = (p.trace != nullptr && blockIdx.x == p.trace_cta && lane == 0 && it < 24) ? p.trace + 64 + warp * 48 + it * 2 : nullptr;
      if (tr != nullptr) tr[0] = clock64();
      const int row = row_base + it * kPassRows;
      const bool live = row >= row_begin && row < row_end;
      load_ptr(it + 2, beg2, deg2);
      int cur_src[kHold];
      float cur_e[kHold][DE];
#pragma unroll
      for (int h = 0; h < kHold; ++h) cur_src[h] = pre_src[h];
      load_attr(beg, deg, 0, cur_e);
      load_src(beg1, deg1, 0, pre_src);   // next pass (zeros past the last pass)
      const int nmax = __reduce_max_sync(0xffffffffu, deg);
      float4 acc[4] = {init4, init4, init4, init4};
      float4 v[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int i = 0; i < 4; ++i) v[u][i] = init4;
      for (int b = 0; b < nmax; b += 8 * kHold) {
        if (b > 0) { load_src(beg, deg, b, cur_src); load_attr(beg, deg, b, cur_e); }   // in-degree above 32: not prefetched
#pragma unroll
        for (int hh = 0; hh < kHold; ++hh)
        for (int gg = 0; gg < 2; ++gg) {
          const int g = 2 * hh + gg;
          if (b + 4 * g >= nmax) break;   // warp-uniform
          const int sreg = cur_src[hh];
          float ereg[DE];
#pragma unroll
          for (int d = 0; d < DE; ++d) ereg[d] = cur_e[hh][d];
          const int l0 = (4 * g) & 7;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int sidx = __shfl_sync(0xffffffffu, sreg, l0 + u, 8);
            const bool on = b + 4 * g + u < deg;
            const float* rp = bcol + static_cast<int64_t>(sidx) * kMain;
            // max / min: a slot beyond the row's degree keeps the registers of an earlier slot of the same row
            // (its edge attributes are zero: the value is re-submitted unchanged); sums start every slot from zero
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
          if (kWeightsInRegs) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
              for (int d = 0; d < DE; ++d) {
                const float ed = __shfl_sync(0xffffffffu, ereg[d], l0 + u, 8);   // 0 for slots beyond the row's degree
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][i] = fma4x2(ed, w[d][i], v[u][i]);
              }
            }
          } else {
#pragma unroll
            for (int d = 0; d < DE; ++d) {
              float4 wd[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) wd[i] = ld4(&s.ws[d * kMain + 32 * i + 4 * q]);
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float ed = __shfl_sync(0xffffffffu, ereg[d], l0 + u, 8);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][i] = fma4x2(ed, wd[i], v[u][i]);
              }
            }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[i] = combine4x2<MODE>(acc[i], v[0][i], v[1][i]);
            acc[i] = combine4x2<MODE>(acc[i], v[2][i], v[3][i]);
          }
        }
      }
      const int deg_row = deg;
      beg = beg1; deg = deg1; beg1 = beg2; deg1 = deg2;
      if (unit >= units) continue;   // warp-uniform: beyond the CTA's last tile
      float4 r[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);   // torch_scatter: empty segments aggregate to 0
      if (live) {
        if (deg_row > 0) {
          const float inv = 1.f / static_cast<float>(deg_row);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 bmain = ld4(&s.bs[32 * i + 4 * q]);
            if (MODE == RGNN_AGGR_MEAN) r[i] = fma4(inv, acc[i], bmain);
            else r[i] = add4(bmain, acc[i]);
          }
        } else if (p.iso.w_t != nullptr) {
          // the update weights carry W_m W_t for every node: a node without incoming edge cancels that term
          // with M' = -W_t x_n instead of 0 (rare: strided reads)
          float a16[4][4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) a16[i][j] = 0.f;
          const float* xr = p.iso.x + (p.iso.rows != nullptr ? static_cast<int64_t>(p.iso.rows[row]) : row) * p.iso.ldx;
          for (int c = 0; c < p.iso.c; ++c) {
            float xv = xr[c];
            if (p.iso.mean != nullptr) xv = (xv - p.iso.mean[c]) * p.iso.scale[c] + p.iso.beta[c];
            if (p.iso.relu) xv = fmaxf(xv, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j)
                a16[i][j] = fmaf(p.iso.w_t[static_cast<int64_t>(32 * i + 4 * q + j) * p.iso.ldw + c], xv, a16[i][j]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) r[i] = make_float4(-a16[i][0], -a16[i][1], -a16[i][2], -a16[i][3]);
        }
      }
      // the unit's four rows -> ring quarter (unit / 8) % kRing, once the tile warp has drained its previous use
      const int qi = unit >> 3, slot = qi & (kRing - 1), round = qi / kRing;
      const long long tw = tr != nullptr ? clock64() : 0;
      if (round >= 1 && !wait_consumed(&s.consumed[slot], round)) timed_out = true;
      if (tr != nullptr) tr[1] = clock64() - tw;
      const int rq = (4 * unit + quarter) & 31;
      float* dst = s.ring + slot * kQuarterFloats + rq * 32 + ((q ^ (rq & 7)) << 2);
#pragma unroll
      for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(dst + i * 1024) = r[i];
      mbar_arrive(&full[slot]);
    }
  } else {
    // =========================== tile warps ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTileRegs));
    const int qd = warp & 3;                      // TMEM lane quarter = 32-row quarter of the tile
    const int tt = tid - kAggThreads;             // 0..127 = row of the tile
    const uint32_t lane_base = tmem + (static_cast<uint32_t>(qd * 32) << 16);
    // resident update weights (overlaps the aggregate warps' first pass)
    {
      const int total16 = (kKBlocks * 2 * np * 32) >> 2;
      const uint32_t w_addr = smem_u32(s.w);
      for (int i = tt; i < total16; i += kTileThreads) cp_async16(w_addr + i * 16u, p.wpack + i * 4, 16);
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (MMA)
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    const uint32_t leader = lane == 0 ? 1u : 0u;
    const uint32_t idesc = umma_idesc_tf32(kRows, np), idesc2 = umma_idesc_tf32(kRows, 2 * np);
    const uint32_t w_block16 = (2u * static_cast<uint32_t>(np) * 128u) >> 4;
    const uint64_t dw0 = umma_desc(smem_u32(s.w));
    float* st = s.stage + qd * kStageFloats;
    const int n_blocks = np >> 4;
    double bn_sum = 0.0, bn_sq = 0.0;             // thread tt < c_out: running column sums of channel tt

    // the x rows of a tile are pulled into L2 two tiles ahead (a row = two 128-byte lines), so that the loads of
    // step (1) are L2 hits: a shared-memory staging buffer would take 32 KB away from the L1 the gathers live on
    auto prefetch_x = [&](int t) {
      const int row = vbeg + t * kRows + tt;
      if (t < tiles && row >= row_begin && row < row_end) {
        const float* xr = p.x + (p.x_rows != nullptr ? static_cast<int64_t>(p.x_rows[row]) : row) * p.ldx;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xr));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(xr + 32));
      }
    };
    prefetch_x(0);
    prefetch_x(1);

    // Tail channels of M' (the De edge-attribute columns of the message): reduced per row by tail_reduce_kernel
    // (below) before this kernel starts -- a per-row gather in the tile warps costs four dependent load rounds
    // of ~3 k cycles each behind the aggregate warps' gathers in the load-store queue.  One 16-byte load here,
    // fetched one tile ahead.
    const int pt = p.p - kMain;                   // tail channels (1..4)
    auto load_tail = [&](int t) {
      const int row = vbeg + t * kRows + tt;
      return (t < tiles && row >= row_begin && row < row_end) ? ld4(p.mt + static_cast<int64_t>(row) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 mt4 = load_tail(0);

    for (int t = 0; t < tiles; ++t) {
      const int row = vbeg + t * kRows + tt;
      const bool row_ok = row >= row_begin && row < row_end;
      const uint32_t par = static_cast<uint32_t>(t) & 1u;
      long long* tr = (p.trace != nullptr && blockIdx.x == p.trace_cta && tt == 0 && t < 8) ? p.trace + t * 8 : nullptr;
      if (tr != nullptr) tr[0] = clock64();
      // ---- (1) x row (L2-prefetched) -> BatchNorm + ReLU of the previous layer -> hi / lo -> TMEM ----
      {
        const float* xr = p.x + (row_ok ? (p.x_rows != nullptr ? static_cast<int64_t>(p.x_rows[row]) : row) : 0) * p.ldx;
#pragma unroll 1
        for (int pn = 0; pn < 2; ++pn) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = row_ok ? ld4(xr + pn * 32 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && (p.x_mean != nullptr || p.relu_x)) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 xv = v[j];
              if (p.x_mean != nullptr) {
                const float4 mu = ld4(s.bn + pn * 32 + 4 * j), sc = ld4(s.bn + 64 + pn * 32 + 4 * j), be = ld4(s.bn + 128 + pn * 32 + 4 * j);
                xv.x = (xv.x - mu.x) * sc.x + be.x; xv.y = (xv.y - mu.y) * sc.y + be.y;
                xv.z = (xv.z - mu.z) * sc.z + be.z; xv.w = (xv.w - mu.w) * sc.w + be.w;
              }
              if (p.relu_x) { xv.x = fmaxf(xv.x, 0.f); xv.y = fmaxf(xv.y, 0.f); xv.z = fmaxf(xv.z, 0.f); xv.w = fmaxf(xv.w, 0.f); }
              v[j] = xv;
            }
          }
          store_panel_hi_lo(lane_base + kXCol + pn * 64, v);
        }
      }
      prefetch_x(t + 2);
      if (tr != nullptr) tr[1] = clock64();
      // ---- (3) M' quarter: ring -> hi / lo -> TMEM ----
      if (tr != nullptr) tr[2] = clock64();
      if (!mbar_wait(&full[qd], par)) timed_out = true;
      if (tr != nullptr) tr[3] = clock64();
      {
        const float* src = s.ring + qd * kQuarterFloats + lane * 32;
        const int sw = lane & 7;
#pragma unroll
        for (int pn = 0; pn < 4; ++pn) {
          float4 v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = ld4(src + pn * 1024 + ((j ^ sw) << 2));
          store_panel_hi_lo(lane_base + kMCol + pn * 64, v);
        }
      }
      __syncwarp();
      if (lane == 0) st_release_shared(&s.consumed[qd], t + 1);   // the quarter is in registers / TMEM: the aggregate warps may refill it
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(a_full);
      if (tr != nullptr) tr[4] = clock64();
      // ---- (4) the tile's MMAs: [x | M'] (TMEM) . W^T (shared memory), issued by one lane of tile warp 0 ----
      if (qd == 0) {
        if (!mbar_wait(a_full, par)) timed_out = true;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int kc = 0; kc < kKBlocks; ++kc) {
          const uint32_t a_hi = tmem + (kc < 2 ? kXCol + static_cast<uint32_t>(kc) * 64u : kMCol + static_cast<uint32_t>(kc - 2) * 64u);
          const uint32_t a_lo = a_hi + 32u;
          const uint64_t w_hi = dw0 + static_cast<uint64_t>(static_cast<uint32_t>(kc) * w_block16);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            // D[:, 0:np] += A_hi W_hi^T, D[:, np:2np] += A_hi W_lo^T in one instruction, then D[:, 0:np] += A_lo W_hi^T
            umma_tf32_ts_pred(tmem, a_hi + 8 * jj, w_hi + 2 * jj, idesc2, (kc > 0 || jj > 0) ? 1u : 0u, leader);
            umma_tf32_ts_pred(tmem, a_lo + 8 * jj, w_hi + 2 * jj, idesc, 1u, leader);
          }
        }
        umma_commit_pred(acc_full, leader);
      }
      const float mt[4] = {mt4.x, mt4.y, mt4.z, mt4.w};
      mt4 = load_tail(t + 1);
      // ---- (5) epilogue ----
      if (tr != nullptr) tr[5] = clock64();
      if (!mbar_wait(acc_full, par)) timed_out = true;
      if (tr != nullptr) tr[6] = clock64();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int tile_row0 = vbeg + t * kRows + qd * 32;
      const int rows_first = max(0, row_begin - tile_row0);          // rows of the quarter below the CTA's range
      const int rows_valid = max(0, min(32, row_end - tile_row0));
      const int c4 = lane & 7, rsub = lane >> 3;
      for (int cd = 0; cd * 2 < n_blocks; ++cd) {
        float* strow = st + lane * 36;
        const int ncol = (cd * 2 + 1 < n_blocks) ? 32 : 16;
#pragma unroll 1
        for (int hb = 0; hb * 16 < ncol; ++hb) {
          const uint32_t taddr = lane_base + static_cast<uint32_t>(cd * 32 + hb * 16);
          uint32_t r[16], r2[16];
          tmem_ld16(taddr, r);
          tmem_ld16(taddr + np, r2);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int col = cd * 32 + hb * 16 + j4 * 4;
            float4 o = ld4(s.bias + col);   // update bias + (hi*hi + lo*hi) + hi*lo + rank-De update with the tail channels
            o.x += __uint_as_float(r[j4 * 4 + 0]) + __uint_as_float(r2[j4 * 4 + 0]);
            o.y += __uint_as_float(r[j4 * 4 + 1]) + __uint_as_float(r2[j4 * 4 + 1]);
            o.z += __uint_as_float(r[j4 * 4 + 2]) + __uint_as_float(r2[j4 * 4 + 2]);
            o.w += __uint_as_float(r[j4 * 4 + 3]) + __uint_as_float(r2[j4 * 4 + 3]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < pt) o = fma4(mt[j], ld4(s.wtail + j * 64 + col), o);
            *reinterpret_cast<float4*>(strow + hb * 16 + j4 * 4) = o;
          }
        }
        __syncwarp();
        const int col = cd * 32 + c4 * 4;
        float4 csum4 = make_float4(0.f, 0.f, 0.f, 0.f), csq4 = csum4;
        if (c4 * 4 < ncol && col < p.c_out) {
          float* dst = p.y + static_cast<int64_t>(tile_row0 + rsub) * p.ldy + col;
          const float* srow = st + rsub * 36 + c4 * 4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it * 4 + rsub >= rows_first && it * 4 + rsub < rows_valid) {
              const float4 v4 = ld4(srow + it * (4 * 36));
              *reinterpret_cast<float4*>(dst + static_cast<int64_t>(it) * 4 * p.ldy) = v4;
              csum4.x += v4.x; csum4.y += v4.y; csum4.z += v4.z; csum4.w += v4.w;
              csq4.x = fmaf(v4.x, v4.x, csq4.x); csq4.y = fmaf(v4.y, v4.y, csq4.y);
              csq4.z = fmaf(v4.z, v4.z, csq4.z); csq4.w = fmaf(v4.w, v4.w, csq4.w);
            }
          }
        }
        if (p.bn_partial != nullptr) {
          // column sums of the warp's 32 rows: the four row groups in a fixed order -> deterministic
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            csum4.x += __shfl_xor_sync(0xffffffffu, csum4.x, o); csum4.y += __shfl_xor_sync(0xffffffffu, csum4.y, o);
            csum4.z += __shfl_xor_sync(0xffffffffu, csum4.z, o); csum4.w += __shfl_xor_sync(0xffffffffu, csum4.w, o);
            csq4.x += __shfl_xor_sync(0xffffffffu, csq4.x, o); csq4.y += __shfl_xor_sync(0xffffffffu, csq4.y, o);
            csq4.z += __shfl_xor_sync(0xffffffffu, csq4.z, o); csq4.w += __shfl_xor_sync(0xffffffffu, csq4.w, o);
          }
          if (rsub == 0 && c4 * 4 < ncol) {
            *reinterpret_cast<float4*>(s.colsum + qd * 64 + col) = csum4;
            *reinterpret_cast<float4*>(s.colsq + qd * 64 + col) = csq4;
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // accumulator reads done before the next tile's MMAs
      if (tr != nullptr) tr[7] = clock64();
      if (p.bn_partial != nullptr) {
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the four quarters' column sums are in shared memory
        if (tt < p.c_out) {
          bn_sum += (static_cast<double>(s.colsum[tt]) + static_cast<double>(s.colsum[64 + tt])) +
                    (static_cast<double>(s.colsum[128 + tt]) + static_cast<double>(s.colsum[192 + tt]));
          bn_sq += (static_cast<double>(s.colsq[tt]) + static_cast<double>(s.colsq[64 + tt])) +
                   (static_cast<double>(s.colsq[128 + tt]) + static_cast<double>(s.colsq[192 + tt]));
        }
      }
    }
    if (p.bn_partial != nullptr && tt < p.c_out) {
      // channel-major: partial[ch * P + cta] (sums), partial[(c_out + ch) * P + cta] (squares)
      p.bn_partial[static_cast<int64_t>(tt) * p.n_partials + blockIdx.x] = bn_sum;
      p.bn_partial[static_cast<int64_t>(p.c_out + tt) * p.n_partials + blockIdx.x] = bn_sq;
    }
  }

  // a barrier that never completes is a kernel bug: fail loudly instead of handing back a partial result
  if (timed_out) {
    if (p.status != nullptr) atomicExch(p.status, RGNN_ERR_CUDA);
    __trap();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.trace != nullptr && blockIdx.x == p.trace_cta && tid == 0) p.trace[1001] = clock64();
  if (warp == kAggWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}


// ---- tail channels of the aggregated messages ------------------------------------------------------------
// M'_tail[n, 0:4] = b_tail + reduce over n's slots of (B_tail[src] + W_e,tail e)   (0 for an empty segment; the
// isolated-node cancel term -W_t x_n on the folded path).  8 lanes per row, lane = slot (coalesced slot loads,
// one 16-byte gather per lane), fixed-order butterfly: deterministic sums.  [N, 4] floats, 1.6 MB at the
// headline size -- read back by the fused kernel's tile warps with one load per row.
template <int MODE, int DE>
__global__ void __launch_bounds__(256)
tail_reduce_kernel(const float* __restrict__ bt, int p, const float* __restrict__ bias, const float* __restrict__ w_e,
                   int64_t ldwe, const float* __restrict__ ea, const int32_t* __restrict__ csc_ptr,
                   const int32_t* __restrict__ csc_src, int n_nodes, float* __restrict__ out_t, IsolatedNodeTerm iso) {
  // 4 lanes per row, lane q takes slots q, q + 4, q + 8, q + 12 of a batch of 16: four independent
  // index -> gather chains in flight per lane (the kernel is a chain of three dependent load rounds)
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = gid >> 2, q = threadIdx.x & 3;
  const bool live = row < n_nodes;
  constexpr float kInit = MODE == RGNN_AGGR_MAX ? -INFINITY : (MODE == RGNN_AGGR_MIN ? INFINITY : 0.f);
  int beg = 0, deg = 0;
  if (live) { beg = csc_ptr[row]; deg = csc_ptr[row + 1] - beg; }
  float4 we[DE];
#pragma unroll
  for (int d = 0; d < DE; ++d) {
    we[d].x = w_e[static_cast<int64_t>(kMain) * ldwe + d];
    we[d].y = kMain + 1 < p ? w_e[static_cast<int64_t>(kMain + 1) * ldwe + d] : 0.f;
    we[d].z = kMain + 2 < p ? w_e[static_cast<int64_t>(kMain + 2) * ldwe + d] : 0.f;
    we[d].w = kMain + 3 < p ? w_e[static_cast<int64_t>(kMain + 3) * ldwe + d] : 0.f;
  }
  float4 tacc = make_float4(kInit, kInit, kInit, kInit);
  for (int b = 0; b < deg; b += 16) {
    int src[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) src[j] = b + 4 * j + q < deg ? csc_src[beg + b + 4 * j + q] : 0;
    float4 t[4];
    float e[4][DE];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool on = b + 4 * j + q < deg;
      t[j] = on ? ld4(bt + static_cast<int64_t>(src[j]) * 4) : tacc;
#pragma unroll
      for (int d = 0; d < DE; ++d) e[j][d] = on ? ea[static_cast<int64_t>(beg + b + 4 * j + q) * DE + d] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (b + 4 * j + q < deg) {
        float4 tj = t[j];
#pragma unroll
        for (int d = 0; d < DE; ++d) tj = fma4(e[j][d], we[d], tj);
        tacc = combine4<MODE>(tacc, tj);
      }
    }
  }
#pragma unroll
  for (int o = 2; o > 0; o >>= 1) {   // fixed-order butterfly over the row's 4 lanes
    float4 other;
    other.x = __shfl_xor_sync(0xffffffffu, tacc.x, o); other.y = __shfl_xor_sync(0xffffffffu, tacc.y, o);
    other.z = __shfl_xor_sync(0xffffffffu, tacc.z, o); other.w = __shfl_xor_sync(0xffffffffu, tacc.w, o);
    tacc = combine4<MODE>(tacc, other);
  }
  if (!live || q != 0) return;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);   // torch_scatter: empty segments aggregate to 0
  if (deg > 0) {
    const float4 bte = make_float4(bias[kMain], kMain + 1 < p ? bias[kMain + 1] : 0.f, kMain + 2 < p ? bias[kMain + 2] : 0.f,
                                   kMain + 3 < p ? bias[kMain + 3] : 0.f);
    if (MODE == RGNN_AGGR_MEAN) r = fma4(1.f / static_cast<float>(deg), tacc, bte);
    else r = add4(bte, tacc);
  } else if (iso.w_t != nullptr) {
    const float* xr = iso.x + (iso.rows != nullptr ? static_cast<int64_t>(iso.rows[row]) : row) * iso.ldx;
    float t4[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < iso.c; ++c) {
      float xv = xr[c];
      if (iso.mean != nullptr) xv = (xv - iso.mean[c]) * iso.scale[c] + iso.beta[c];
      if (iso.relu) xv = fmaxf(xv, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (kMain + j < p) t4[j] = fmaf(iso.w_t[static_cast<int64_t>(kMain + j) * iso.ldw + c], xv, t4[j]);
    }
    r = make_float4(-t4[0], -t4[1], -t4[2], -t4[3]);
  }
  *reinterpret_cast<float4*>(out_t + static_cast<int64_t>(row) * 4) = r;
}

template <int MODE>
int launch_tail_mode(const FusedLayerArgs& a, cudaStream_t stream) {
  const int n = static_cast<int>(a.n_nodes);
  const unsigned blocks = div_up(static_cast<int64_t>(n) * 4, 256);
#define RGNN_TAIL_CASE(DE_)                                                                                         \
  case DE_:                                                                                                         \
    tail_reduce_kernel<MODE, DE_><<<blocks, 256, 0, stream>>>(a.bt, a.p, a.bias_msg, a.w_e, a.ldwe, a.ea, a.csc_ptr, \
                                                               a.csc_src, n, a.mt, a.iso);                          \
    break;
  switch (a.de) {
    RGNN_TAIL_CASE(1) RGNN_TAIL_CASE(2) RGNN_TAIL_CASE(3) RGNN_TAIL_CASE(4)
    default: return RGNN_ERR_UNSUPPORTED;
  }
#undef RGNN_TAIL_CASE
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

template <int MODE, int DE>
int launch_fused_instance(const FusedParams& p, int grid, cudaStream_t stream) {
  const size_t smem = sizeof(float) * fused_smem_floats(p.np, DE) + 1024;
  static bool configured[kMaxDevices] = {};
  RGNN_CUDA_CHECK(opt_in_dynamic_smem(fused_layer_kernel<MODE, DE>, configured, 227 * 1024));
  fused_layer_kernel<MODE, DE><<<grid, kFusedThreads, smem, stream>>>(p);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

template <int MODE>
int launch_fused_mode(const FusedParams& p, int de, int grid, cudaStream_t stream) {
  switch (de) {
    case 1: return launch_fused_instance<MODE, 1>(p, grid, stream);
    case 2: return launch_fused_instance<MODE, 2>(p, grid, stream);
    case 3: return launch_fused_instance<MODE, 3>(p, grid, stream);
    case 4: return launch_fused_instance<MODE, 4>(p, grid, stream);
    default: return RGNN_ERR_UNSUPPORTED;
  }
}

// rows per CTA: the same share for every SM, in units of one aggregate warp's 4 rows
int fused_rows_per_cta(int64_t n_nodes) {
  const int ctas = sm_count();
  int rows = static_cast<int>((n_nodes + ctas - 1) / ctas);
  rows = (rows + 3) / 4 * 4;
  return rows < 16 ? 16 : rows;
}

}  // namespace

bool fused_layer_supported(const rgnn_conv_desc& d, const ConvShape& s) {
  // read on every call (not cached): tests and scripts/check_fused.py switch between the two paths in one process
  const char* e = getenv("RGNN_DISABLE_FUSED_LAYER");
  if (e != nullptr && e[0] == '1') return false;
  return d.conv_type == RGNN_CONV_MPNN && s.split && s.c == kC && s.pm == kMain && s.pt4 == 4 && s.de >= 1 && s.de <= 4 &&
         d.post_layers == 1 && d.aggr != RGNN_AGGR_ADD && s.c_out >= 4 && s.c_out <= 64 && s.c_out % 4 == 0;
}

int64_t fused_layer_partials(int64_t n_nodes) {
  return div_up(n_nodes, fused_rows_per_cta(n_nodes));
}

// debug: device buffer of 1024 int64 receiving the timeline of CTA `cta` of the next launch
static long long* g_fused_trace = nullptr;
static int g_fused_trace_cta = 0;
extern "C" void rgnn_debug_trace_fused_layer(void* device_buffer, int cta) {
  g_fused_trace = static_cast<long long*>(device_buffer);
  g_fused_trace_cta = cta;
}

int launch_fused_layer(const FusedLayerArgs& a, cudaStream_t stream) {
  if (a.n_nodes <= 0) return RGNN_OK;
  if (a.n_nodes > 0x7ffffff0LL) return RGNN_ERR_INVALID_ARGUMENT;
  if (reinterpret_cast<uintptr_t>(a.x) % 16 != 0 || a.ldx % 4 != 0 || reinterpret_cast<uintptr_t>(a.y) % 16 != 0 || a.ldy % 4 != 0 ||
      reinterpret_cast<uintptr_t>(a.wpack) % 16 != 0)
    return RGNN_ERR_UNSUPPORTED;
  FusedParams p{};
  p.bm = a.bm; p.mt = a.mt; p.p = a.p; p.bias_msg = a.bias_msg; p.w_e = a.w_e; p.ldwe = a.ldwe; p.ea = a.ea;
  p.csc_ptr = a.csc_ptr; p.csc_src = a.csc_src; p.iso = a.iso;
  p.x = a.x; p.ldx = a.ldx; p.x_rows = a.x_rows; p.x_mean = a.x_mean; p.x_scale = a.x_scale; p.x_beta = a.x_beta; p.relu_x = a.relu_x;
  p.wpack = a.wpack; p.w_tail = a.w_tail; p.ld_wtail = a.ld_wtail; p.bias_post = a.bias_post;
  p.c_out = a.c_out; p.np = tc_padded_n(a.c_out);
  p.y = a.y; p.ldy = a.ldy; p.bn_partial = a.bn_partial; p.status = a.status;
  p.n_nodes = static_cast<int32_t>(a.n_nodes);
  p.rows_per_cta = fused_rows_per_cta(a.n_nodes);
  p.tiles_per_cta = (p.rows_per_cta + kRows - 1) / kRows;
  const int grid = static_cast<int>(div_up(a.n_nodes, p.rows_per_cta));
  p.n_partials = grid;
  p.trace = g_fused_trace; p.trace_cta = g_fused_trace_cta; g_fused_trace = nullptr;
  {
    const char* e = getenv("RGNN_FUSED_LAST_TILE_FULL");   // read per launch: scripts compare the two layouts in one process
    p.last_tile_full = (e != nullptr) ? (e[0] == '1' ? 1 : 0) : kLastTileFullDefault;
  }
  {
    RGNN_PROFILE("edge_tail_reduce", stream);
    int st = RGNN_ERR_UNSUPPORTED;
    switch (a.aggr) {
      case RGNN_AGGR_MAX: st = launch_tail_mode<RGNN_AGGR_MAX>(a, stream); break;
      case RGNN_AGGR_MIN: st = launch_tail_mode<RGNN_AGGR_MIN>(a, stream); break;
      case RGNN_AGGR_MEAN: st = launch_tail_mode<RGNN_AGGR_MEAN>(a, stream); break;
      default: break;
    }
    RGNN_RETURN_IF_ERROR(st);
  }
  RGNN_PROFILE("edge_update_fused", stream);
  switch (a.aggr) {
    case RGNN_AGGR_MAX: return launch_fused_mode<RGNN_AGGR_MAX>(p, a.de, grid, stream);
    case RGNN_AGGR_MIN: return launch_fused_mode<RGNN_AGGR_MIN>(p, a.de, grid, stream);
    case RGNN_AGGR_MEAN: return launch_fused_mode<RGNN_AGGR_MEAN>(p, a.de, grid, stream);
    default: return RGNN_ERR_UNSUPPORTED;
  }
}

}  // namespace rgnn

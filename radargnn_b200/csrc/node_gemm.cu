// node_gemm.cu -- the dense node-level contractions on the 5th-generation tensor cores.
//
//   y[m, n] = in(A)[m, K] . W[n, K]^T + bias (+ residual)        A = [a1 | a2 | a_tail | rowscale * a1]
//
// used for B = x W_s^T (first message Linear, source half) and for the node update
// post_mlp([x ; M]) of MPNNConv / RadarPointGNNConv (reference gnn/mpnn_layers.py:89-90, 98-99,
// 174-175, 181-182).  One persistent CTA per SM, 128-row tiles:
//   * tcgen05.mma.cta_group::1.kind::tf32, M = 128, N = padded output width, accumulator in TMEM;
//   * 3xTF32: every fp32 operand is split into hi (top 19 bits) + lo (remainder) and the product is
//     accumulated as hi*hi + lo*hi + hi*lo in fp32 -- ~2^-21 relative per product, i.e. fp32-grade
//     results (plain TF32 is ~1e-3 and would miss the reference's 1e-4 parity bar); for narrow outputs
//     hi*hi and hi*lo are one MMA with N = 2 np against [W_hi ; W_lo];
//   * the A operand needs a transform (BatchNorm+ReLU of the previous layer applied on load, hi/lo split),
//     so it cannot go from TMA straight to the MMA: raw fp32 panels (128 rows x 32 floats) are brought
//     into a shared-memory ring by the async engines (TMA tensor maps for row-major operands, bulk copies
//     for the panel-major M', 16-byte cp.async for a gathered operand), converter warps transform / split
//     them and write the hi and lo images into TENSOR MEMORY (tcgen05.st), and the MMAs read A from TMEM;
//   * W (hi and lo images, packed per weight version by pack_weights_kernel) stays resident in shared
//     memory for all tiles of the CTA;
//   * epilogue: tcgen05.ld (32 columns per warp) -> per-warp transpose buffer -> coalesced 128-byte row
//     stores with bias / residual, plus deterministic per-tile column sums for the BatchNorm statistics.
// DESIGN.md sections 3.5 / 3.6 hold the measurements behind every one of these choices.
#include <string.h>

#include "node_gemm.cuh"
#include "tc_common.cuh"

namespace rgnn {
namespace {

using namespace tc;   // kRows, kKc, kABufFloats and the inline-PTX helpers (tc_common.cuh)

// ---- weight packing ------------------------------------------------------------------------
// image = per 32-float K block one swizzled [np x 32] panel of hi parts directly followed by the panel of
// lo parts, so that a single B descriptor with N = 2 np covers [W_hi ; W_lo] of the block
// (np = padded width of one output-column chunk, rows chunks * np in total: chunk c of the image is the
// complete K-block sequence of output rows c * np .. c * np + np - 1)
__global__ void __launch_bounds__(256)
pack_weights_kernel(TcWeightBlocks blocks, int np, int chunks, int kp, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= chunks * np * kp) return;   // kp: multiple of 32 here
  const int nrow = idx / kp, kcol = idx - nrow * kp;
  float v = 0.f;
  for (int b = 0; b < blocks.count; ++b) {
    const TcWeightBlock& wb = blocks.block[b];
    if (nrow < wb.rows && kcol >= wb.k_offset && kcol < wb.k_offset + wb.cols) {
      const float t = wb.src[static_cast<int64_t>(nrow) * wb.ld + (kcol - wb.k_offset)];
      v = wb.accumulate ? v + t : t;
    }
  }
  float hi, lo;
  split_tf32(v, hi, lo);
  const int chunk = nrow / np, r = nrow - chunk * np;
  const int off = chunk * (2 * np * kp) + (kcol >> 5) * (2 * np * 32) + sw128_offset(r, (kcol & 31) >> 2) + (kcol & 3);
  out[off] = hi;
  out[np * 32 + off] = lo;
}

// w_fold[c_out, c] = W_m[c_out, p] . W_t[p, c]   (fp64 accumulate): the target-node half of the first
// message Linear folded through the update Linear
__global__ void __launch_bounds__(256)
fold_weights_kernel(const float* __restrict__ w_m, int64_t ld_m, const float* __restrict__ w_t, int64_t ld_t,
                    int c_out, int p, int c, float* __restrict__ w_fold) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= c_out * c) return;
  const int o = idx / c, i = idx - o * c;
  double acc = 0.0;
  for (int j = 0; j < p; ++j) acc += static_cast<double>(w_m[o * ld_m + j]) * static_cast<double>(w_t[j * ld_t + i]);
  w_fold[idx] = static_cast<float>(acc);
}

// ---- the GEMM ------------------------------------------------------------------------------
// K is laid out segment by segment, [a1 | a2 | a_tail | rowscale * a1], each padded to whole panels of
// 32 floats, so a panel (128 rows x 32 floats) lies in exactly one segment.
//
// The A operand lives in TENSOR MEMORY (tcgen05.mma with A from TMEM, W from shared memory):
//   global --cp.async--> raw fp32 panel ring in shared memory (up to 6 x 16 KB in flight per SM: enough
//   to cover DRAM latency at full bandwidth) --converter warps: BatchNorm + ReLU on load, row scale,
//   hi / lo split--> tcgen05.st into a TMEM A stage (64 columns: 32 hi + 32 lo) --> MMA.
// Shared memory then only holds W (resident) and the raw ring; there is no swizzled A stage, no
// generic->async proxy fence, and TMEM (512 columns) has room for the two accumulators plus 3-4 A stages.
//
// Warp roles (736 threads, one CTA per SM, persistent over 128-row tiles; a warp may only touch the TMEM
// lane quarter warp_id % 4, so every group of four consecutive warps covers all 128 lanes):
//   warps 0-1   loaders: cp.async 16-byte items of a panel into the ring slot (8 lanes per 128-byte row
//               segment, zero fill for rows / columns beyond the operand), completion via
//               cp.async.mbarrier.arrive on raw_full[slot]
//   warps 2-13  converters, three groups of four taking panels round-robin (the per-panel chain wait ->
//               LDS -> split -> tcgen05.st -> wait::st -> arrive is latency-bound, so panels overlap):
//               thread = row, 32 floats from the ring (swizzled: conflict-free), transform, split,
//               2 x 2 tcgen05.st.x16 -> a_full[stage]; the slot goes back to the loaders (raw_empty)
//   warps 14-21 epilogue: TMEM -> registers -> transpose buffer -> coalesced global stores (+ residual,
//               BatchNorm column sums); two warps per TMEM lane quarter, alternating 32-column blocks
//   warp  22    MMA issuer (one lane): per k-step hi*hi + lo*hi + hi*lo, tcgen05.commit -> a_empty / acc_full
constexpr int kMaxRaw = 6;
constexpr int kMaxAStages = 4;
constexpr int kLoaderWarps = 2, kConvGroups = 3, kConvWarps = 4 * kConvGroups;
constexpr int kLoaderThreads = kLoaderWarps * 32;
constexpr int kLoaderItems = 1024 / kLoaderThreads;   // 16-byte items of a panel per loader thread
constexpr int kConvGroupThreads = 128;            // one converter group = 4 warps, one per TMEM lane quarter
constexpr int kEpilogueThreads = 256;
constexpr int kEpiWarp0 = kLoaderWarps + kConvWarps;
constexpr int kProducerWarps = kEpiWarp0;         // warps in front of the epilogue warps
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kMmaWarp = kEpiWarp0 + kEpilogueThreads / 32;
constexpr int kGemmThreads = (kMmaWarp + 1) * 32;
constexpr int kMaxPanels = 32;
constexpr int kAStageCols = 64;                   // TMEM columns of one A stage: 32 hi + 32 lo

enum PanelFlags : int32_t { kPanelBn = 1, kPanelRelu = 2, kPanelRowScale = 4, kPanelGather = 8, kPanelBulk = 16, kPanelTmaA1 = 32, kPanelTmaAt = 64 };

struct PanelInfo {
  const float* base;   // segment base pointer (column 0 of the segment)
  int64_t ld;          // row stride of the segment, floats
  int32_t col0;        // first segment column of this panel
  int32_t valid;       // valid floats in this panel (multiple of 4, <= 32)
  int32_t flags;
  int32_t ksteps;      // 8-float MMA k-steps that carry data
};

struct SmemLayout {
  float* w;                        // per K block: [np x 32] hi panel, [np x 32] lo panel (swizzled)
  float* raw;                      // [raw_slots][128 x 32] fp32 panels as loaded (row r chunk c at c ^ (r % 8))
  float* col_sum;                  // [2 accumulators][4 quarters][np]
  float* col_sq;
  float* bias;                     // [np]
  float* bn;                       // [3][k1] mean | scale | beta of the a1 transform
  float* stage;                    // optional: 8 epilogue warps x [32][36] transpose buffers (coalesced stores)
  PanelInfo* panel;                // [kMaxPanels]
  uint64_t* bar;                   // raw_full[6], raw_empty[6], a_full[4], a_empty[4], acc_full[2], acc_empty[2], setup
  uint32_t* tmem_base;
};

constexpr int kStageFloats = 8 * 32 * 36;
constexpr int kBarCount = 2 * kMaxRaw + 2 * kMaxAStages + 5;

__device__ __forceinline__ SmemLayout carve_smem(unsigned char* base, int np, int kp32, int raw_slots, int k1, int staged) {
  SmemLayout s;
  float* f = reinterpret_cast<float*>(base);
  s.w = f; f += 2 * static_cast<size_t>(np) * kp32;
  s.raw = f; f += static_cast<size_t>(raw_slots) * kABufFloats;
  s.col_sum = f; f += 8 * np;
  s.col_sq = f; f += 8 * np;
  s.bias = f; f += np;
  s.bn = f; f += 3 * ((k1 + 3) & ~3);
  s.stage = nullptr;
  if (staged) { s.stage = f; f += kStageFloats; }
  s.panel = reinterpret_cast<PanelInfo*>(f); f += kMaxPanels * (sizeof(PanelInfo) / sizeof(float));
  s.bar = reinterpret_cast<uint64_t*>(f); f += 2 * kBarCount;
  s.tmem_base = reinterpret_cast<uint32_t*>(f);
  return s;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
node_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int np = p.np, kp = p.kp, kp32 = (p.kp + 31) & ~31;
  const int a_stages = p.a_stages, raw_slots = p.raw_slots;
  const SmemLayout s = carve_smem(smem_raw, np, kp32, raw_slots, p.k1, p.staged_epilogue);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int tmem_cols = 512;
  uint64_t* raw_full = s.bar;                          // [kMaxRaw]
  uint64_t* raw_empty = s.bar + kMaxRaw;               // [kMaxRaw]
  uint64_t* a_full = s.bar + 2 * kMaxRaw;              // [kMaxAStages]
  uint64_t* a_empty = a_full + kMaxAStages;            // [kMaxAStages]
  uint64_t* acc_full = a_empty + kMaxAStages;          // [2]
  uint64_t* acc_empty = acc_full + 2;                  // [2]

  uint64_t* setup_bar = acc_empty + 2;
  // Output-column chunks: a CTA works on ONE chunk of np columns (its slice of W stays resident) and on the
  // row tiles cta_c, cta_c + ctas_c, ... -- wide outputs / long K (d = 128 layers) do not fit shared memory
  // with all of W, the A panels are then re-streamed once per chunk (from L2)
  const int n_chunks = p.n_chunks;
  const int ck = static_cast<int>(blockIdx.x) % n_chunks, cta_c = static_cast<int>(blockIdx.x) / n_chunks;
  const int ctas_c = static_cast<int>(gridDim.x) / n_chunks;
  const int col0 = ck * np;                               // first output column of the chunk
  const int n_loc = min(np, p.n - col0);                  // valid / stored columns of the chunk
  const int n_store_loc = min(np, p.n_store - col0);
  const bool dual = p.dual != 0;   // one MMA per k-step for hi*hi and hi*lo: N = 2 np against [W_hi ; W_lo]
  const int acc_stride = dual ? 2 * np : np;   // TMEM columns of one accumulator

  // ---- one-time setup, part 1 (no global loads): barriers, TMEM, panel table ----------------------
  long long t_entry = 0;
  if (p.trace != nullptr && blockIdx.x == 0 && tid == 0) t_entry = clock64();
  if (p.trace != nullptr && tid == 0 && blockIdx.x < 192) {   // debug: per-CTA lifetime in ns (trace[256 + 2 b + {0, 1}])
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x] = static_cast<long long>(gt);
  }
  if (tid == 0) {
    for (int i = 0; i < kMaxRaw; ++i) { mbar_init(&raw_full[i], kLoaderThreads); mbar_init(&raw_empty[i], kConvGroupThreads); }
    for (int i = 0; i < kMaxAStages; ++i) { mbar_init(&a_full[i], kConvGroupThreads); mbar_init(&a_empty[i], 1); }
    mbar_init(&acc_full[0], 1); mbar_init(&acc_full[1], 1);
    mbar_init(&acc_empty[0], kEpilogueThreads); mbar_init(&acc_empty[1], kEpilogueThreads);
    mbar_init(setup_bar, kEpilogueThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s.tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int k1p = (p.k1 + 31) & ~31, k2p = (p.k2 + 31) & ~31, ktp = (p.kt + 31) & ~31;
  const int panels = (kp + kKc - 1) / kKc;
  if (tid < panels && tid < kMaxPanels) {
    // which segment panel `tid` lies in
    const int k0 = tid * kKc;
    PanelInfo pi;
    const int a1_flags = (p.a1_mean != nullptr ? kPanelBn : 0) | (p.relu_a1 ? kPanelRelu : 0) |
                         (p.a1_rows != nullptr ? kPanelGather : 0) | (p.a1_tma ? kPanelTmaA1 : 0);
    if (k0 < k1p) {
      pi.base = p.a1; pi.ld = p.lda1; pi.col0 = k0; pi.valid = min(32, p.k1 - k0); pi.flags = a1_flags;
    } else if (k0 < k1p + k2p) {
      pi.base = p.a2; pi.ld = p.lda2; pi.col0 = k0 - k1p; pi.valid = min(32, p.k2 - pi.col0);
      pi.flags = p.relu_a2 ? kPanelRelu : 0;
      if (p.a2_panel_major) { pi.flags |= kPanelBulk; pi.ld = p.k2 >> 5; }   // ld = panels per tile
    } else if (k0 < k1p + k2p + ktp) {
      pi.base = p.at; pi.ld = p.ldat; pi.col0 = k0 - k1p - k2p; pi.valid = min(32, p.kt - pi.col0);
      pi.flags = (p.relu_a2 ? kPanelRelu : 0) | (p.at_tma ? kPanelTmaAt : 0);
    } else {
      pi.base = p.a1; pi.ld = p.lda1; pi.col0 = k0 - k1p - k2p - ktp; pi.valid = min(32, p.k3 - pi.col0);
      pi.flags = a1_flags | kPanelRowScale;
    }
    pi.ksteps = (pi.valid + 7) >> 3;
    s.panel[tid] = pi;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *s.tmem_base;
  const uint32_t a_col0 = static_cast<uint32_t>(2 * acc_stride);   // A stages follow the two accumulators

  const int64_t n_tiles = (p.m + kRows - 1) / kRows;
  const int64_t my_tiles = cta_c < n_tiles ? (n_tiles - cta_c + ctas_c - 1) / ctas_c : 0;
  const int total = static_cast<int>(my_tiles) * panels;   // panels this CTA streams
  const int m32 = static_cast<int>(p.m);
  bool timed_out = false;
  // optional timeline of CTA 0 (debug): trace[role * 64 + tile * 2 + {0, 1}] = clock64
  long long* trace = (p.trace != nullptr && blockIdx.x == 0) ? p.trace : nullptr;
  if (trace != nullptr && tid == 0) { trace[3 * 64] = clock64(); trace[3 * 64 + 63] = t_entry; }

  if (warp < kLoaderWarps) {
    // =========================== loaders ===========================
    // warp w streams rows 64 w .. 64 w + 63 of every panel: item i of a lane = row 64 w + 4 i + lane / 8,
    // 16-byte chunk lane % 8 (a warp instruction covers four whole 128-byte row segments)
    const int c = lane & 7;
    const int rsub = warp * (kRows / kLoaderWarps) + (lane >> 3);
    const uint32_t raw_addr = smem_u32(s.raw);
    int tl = 0, pi = 0, slot = 0;
    uint32_t wrap = 0;   // how often the ring has wrapped
    int32_t rr[kLoaderItems];   // source rows of this thread's items (through the optional gather map), per tile; -1: none
    auto tile_rows = [&](int t) {
      const int row0 = static_cast<int>(cta_c + static_cast<int64_t>(t) * ctas_c) * kRows;
#pragma unroll
      for (int i = 0; i < kLoaderItems; ++i) {
        const int row = row0 + rsub + 4 * i;
        rr[i] = row < m32 ? (p.a1_rows != nullptr ? p.a1_rows[row] : row) : -1;
      }
    };
    if (total > 0) tile_rows(0);
    // L2 prefetch of a whole tile's operands with bulk prefetches (one instruction per contiguous region,
    // no shared memory, no registers): the ring only keeps 48 - 96 KB per SM in flight, DRAM latency under
    // load needs more, so the tiles two ahead are pulled into L2 and the ring's copies become L2 hits
    auto l2_prefetch_tile = [&](int t) {
      if (t >= my_tiles) return;
      const int64_t tile = cta_c + static_cast<int64_t>(t) * ctas_c;
      const int64_t row0 = tile * kRows;
      const uint32_t rows = static_cast<uint32_t>(p.m - row0 < kRows ? p.m - row0 : kRows);
      if (p.a1_rows == nullptr && (p.lda1 & 3) == 0) {
        const float* a = p.a1 + row0 * p.lda1;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(rows * static_cast<uint32_t>(p.lda1) * 4u) : "memory");
      }
      if (p.a2 != nullptr) {
        if (p.a2_panel_major) {
          const float* a = p.a2 + tile * (static_cast<int64_t>(p.k2 >> 5) << 12);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(static_cast<uint32_t>(p.k2 >> 5) * kABufFloats * 4u) : "memory");
        } else if ((p.lda2 & 3) == 0) {
          const float* a = p.a2 + row0 * p.lda2;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(rows * static_cast<uint32_t>(p.lda2) * 4u) : "memory");
        }
      }
      if (p.at != nullptr && (p.ldat & 3) == 0) {
        const float* a = p.at + row0 * p.ldat;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(rows * static_cast<uint32_t>(p.ldat) * 4u) : "memory");
      }
    };
    if (tid == 0) { l2_prefetch_tile(1); l2_prefetch_tile(2); }
    for (int g = 0; g < total; ++g) {
      if (tid == 0 && pi == 0) l2_prefetch_tile(tl + 3);
      if (trace != nullptr && tid == 0 && pi == 0 && tl < 32) trace[0 * 64 + tl * 2] = clock64();
      if (wrap >= 1u && !mbar_wait(&raw_empty[slot], (wrap - 1u) & 1u)) timed_out = true;
      const PanelInfo& info = s.panel[pi];
      const int row0 = static_cast<int>(cta_c + static_cast<int64_t>(tl) * ctas_c) * kRows;
      const bool col_ok = 4 * c < info.valid;
      const float* colp = info.base + info.col0 + 4 * c;
      const bool gather = (info.flags & kPanelGather) != 0;
      const int64_t ld = info.ld;
      const uint32_t slot_addr = raw_addr + static_cast<uint32_t>(slot) * (kABufFloats * 4u);
      if (info.flags & kPanelBulk) {
        // panel-major operand: the whole panel is one contiguous 16 KB block, already in the ring's
        // swizzled layout -> a single bulk copy that completes on the barrier's transaction count
        if (tid == 0) {
          const int64_t tile = cta_c + static_cast<int64_t>(tl) * ctas_c;
          const float* src = info.base + ((tile * ld + (info.col0 >> 5)) << 12);
          const uint32_t bar = smem_u32(&raw_full[slot]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kABufFloats * 4u) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(slot_addr), "l"(src), "r"(kABufFloats * 4u), "r"(bar) : "memory");
        } else {
          mbar_arrive(&raw_full[slot]);
        }
      } else if (info.flags & (kPanelTmaA1 | kPanelTmaAt)) {
        // row-major operand through the tensor-memory accelerator: one 2-D tile load per panel, written in
        // the ring's 128-byte-swizzled layout, rows / columns outside the tensor zero-filled
        if (tid == 0) {
          const uint64_t tmap = reinterpret_cast<uint64_t>((info.flags & kPanelTmaA1) ? &p.tm_a1 : &p.tm_at);
          const uint32_t bar = smem_u32(&raw_full[slot]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kABufFloats * 4u) : "memory");
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(slot_addr), "l"(tmap), "r"(info.col0), "r"(row0), "r"(bar) : "memory");
        } else {
          mbar_arrive(&raw_full[slot]);
        }
      } else {
#pragma unroll
      for (int i = 0; i < kLoaderItems; ++i) {
        const int rl = rsub + 4 * i;
        const bool ok = col_ok && rr[i] >= 0;
        const int64_t r = gather ? static_cast<int64_t>(rr[i]) : static_cast<int64_t>(row0 + rl);
        const float* src = ok ? colp + r * ld : p.a1;
        cp_async16(slot_addr + static_cast<uint32_t>(rl * 32 + ((c ^ (rl & 7)) << 2)) * 4u, src, ok ? 16u : 0u);
      }
      cp_async_arrive(&raw_full[slot]);
      }
      if (++slot == raw_slots) { slot = 0; ++wrap; }
      if (++pi == panels) {
        if (trace != nullptr && tid == 0 && tl < 32) trace[0 * 64 + tl * 2 + 1] = clock64();
        pi = 0; ++tl;
        if (tl < my_tiles) tile_rows(tl);
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < kEpiWarp0) {
    // =========================== converters ===========================
    const int grp = (warp - kLoaderWarps) >> 2;       // group g takes panels g, g + 3, g + 6, ...
    const int quarter = warp & 3;                     // TMEM lane quarter this warp may access
    const int rl = quarter * 32 + lane;               // row of the tile = TMEM lane
    const uint32_t lane_addr = tmem_d + (static_cast<uint32_t>(quarter * 32) << 16) + a_col0;
    const int k1r = (p.k1 + 3) & ~3;
    const float* rawrow = s.raw + rl * 32;
    const int sw = rl & 7;
    // A waiter may be at most one barrier phase ahead of the phase in flight (the parity test cannot tell
    // two phases apart) and panels do not complete in issue order (bulk copies vs. cp.async), so the host
    // makes the ring slot and TMEM stage counts multiples of the group count: a slot / stage is then always
    // used by the same group, which only waits for phase r after having consumed phase r - 1.
    const int n_groups = p.conv_groups;
    if (!mbar_wait(setup_bar, 0u)) timed_out = true;   // BatchNorm-on-load parameters are in shared memory
    // (tile, panel), ring slot and TMEM stage of panel g, advanced by kConvGroups panels per iteration
    int tl = 0, pi = grp, slot = grp, stg = grp;
    uint32_t raw_round = 0, a_round = 0;
    while (pi >= panels) { pi -= panels; ++tl; }
    while (slot >= raw_slots) { slot -= raw_slots; ++raw_round; }
    while (stg >= a_stages) { stg -= a_stages; ++a_round; }
    for (int g = grp; g < total && grp < n_groups; g += n_groups) {
      const PanelInfo& info = s.panel[pi];
      const int flags = info.flags & (kPanelBn | kPanelRelu | kPanelRowScale);
      if (!mbar_wait(&raw_full[slot], raw_round & 1u)) timed_out = true;
      const float* src = rawrow + static_cast<size_t>(slot) * kABufFloats;
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = *reinterpret_cast<const float4*>(src + ((j ^ sw) << 2));
      if (flags != 0) {
        // transform in place: BatchNorm + ReLU of the previous layer on load, row scale.  Rows and columns
        // beyond the operand were zero-filled by the loaders and must stay zero after the affine map.
        const int row = static_cast<int>(cta_c + static_cast<int64_t>(tl) * ctas_c) * kRows + rl;
        const bool row_ok = row < m32;
        float rs = 1.f;
        if ((flags & kPanelRowScale) && row_ok) {
          const int deg = p.csc_ptr[row + 1] - p.csc_ptr[row];
          rs = p.rowscale_mode == 2 ? static_cast<float>(deg) : (deg > 0 ? 1.f : 0.f);
        }
        if (!row_ok) rs = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 x = v[j];
          const int colc = info.col0 + 4 * j;
          const bool cv = 4 * j < info.valid;
          if ((flags & kPanelBn) && cv) {
            const float4 mu = *reinterpret_cast<const float4*>(s.bn + colc);
            const float4 sc = *reinterpret_cast<const float4*>(s.bn + k1r + colc);
            const float4 be = *reinterpret_cast<const float4*>(s.bn + 2 * k1r + colc);
            x.x = (x.x - mu.x) * sc.x + be.x; x.y = (x.y - mu.y) * sc.y + be.y;
            x.z = (x.z - mu.z) * sc.z + be.z; x.w = (x.w - mu.w) * sc.w + be.w;
          }
          if (flags & kPanelRelu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
          x.x *= rs; x.y *= rs; x.z *= rs; x.w *= rs;
          v[j] = x;
        }
      }
      // the MMAs of the previous use of this TMEM stage must have drained it
      if (a_round >= 1u && !mbar_wait(&a_empty[stg], (a_round - 1u) & 1u)) timed_out = true;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float t[16];
        const uint32_t col = static_cast<uint32_t>(stg * kAStageCols + h * 16);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {   // hi parts
          const float4 x = v[h * 4 + jj];
          t[jj * 4 + 0] = __uint_as_float(__float_as_uint(x.x) & 0xffffe000u); t[jj * 4 + 1] = __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
          t[jj * 4 + 2] = __uint_as_float(__float_as_uint(x.z) & 0xffffe000u); t[jj * 4 + 3] = __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        }
        tmem_st16(lane_addr + col, t);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {   // lo parts = x - hi (exact)
          const float4 x = v[h * 4 + jj];
          t[jj * 4 + 0] = x.x - t[jj * 4 + 0]; t[jj * 4 + 1] = x.y - t[jj * 4 + 1];
          t[jj * 4 + 2] = x.z - t[jj * 4 + 2]; t[jj * 4 + 3] = x.w - t[jj * 4 + 3];
        }
        tmem_st16(lane_addr + col + 32, t);
      }
      mbar_arrive(&raw_empty[slot]);   // the panel is in registers / TMEM: the slot can be refilled
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&a_full[stg]);
      pi += n_groups; while (pi >= panels) { pi -= panels; ++tl; }
      slot += n_groups; while (slot >= raw_slots) { slot -= raw_slots; ++raw_round; }
      stg += n_groups; while (stg >= a_stages) { stg -= a_stages; ++a_round; }
    }
  } else if (warp == kMmaWarp) {
    // =========================== MMA issuer ===========================
    // The whole warp runs this loop with uniform control flow and uniform operands; only the tcgen05
    // instructions themselves are predicated to lane 0.  Three MMAs per k-step (hi*hi, lo*hi, hi*lo).
    const uint32_t leader = lane == 0 ? 1u : 0u;
    const uint32_t idesc = umma_idesc_tf32(kRows, np);
    const uint32_t idesc2 = umma_idesc_tf32(kRows, dual ? 2 * np : np);
    const uint32_t w_block16 = (2u * static_cast<uint32_t>(np) * 128u) >> 4;  // one K block of W (hi + lo panels), 16-byte units
    const uint32_t w_lo16 = (static_cast<uint32_t>(np) * 128u) >> 4;          // hi panel -> lo panel
    const uint64_t dw0 = umma_desc(smem_u32(s.w));
    int stg = 0;
    uint32_t round = 0;
    if (!mbar_wait(setup_bar, 0u)) timed_out = true;   // W is resident
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int ab = static_cast<int>(tl & 1);
      if (tl >= 2 && !mbar_wait(&acc_empty[ab], static_cast<uint32_t>(((tl >> 1) - 1) & 1))) timed_out = true;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_addr = tmem_d + static_cast<uint32_t>(ab * acc_stride);
      if (trace != nullptr && lane == 0 && tl < 32) trace[1 * 64 + tl * 2] = clock64();
      for (int kc = 0; kc < panels; ++kc) {
        const int ksteps = s.panel[kc].ksteps;
        if (!mbar_wait(&a_full[stg], round & 1u)) timed_out = true;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = tmem_d + a_col0 + static_cast<uint32_t>(stg * kAStageCols);
        const uint32_t a_lo = a_hi + 32u;
        const uint64_t w_hi = dw0 + static_cast<uint64_t>(static_cast<uint32_t>(kc) * w_block16);
        const uint64_t w_lo = w_hi + w_lo16;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {  // a k-step is 8 TMEM columns of A and 32 bytes = 2 descriptor units of W
          if (jj < ksteps) {
            const uint32_t acc = (kc > 0 || jj > 0) ? 1u : 0u;
            if (dual) {
              // D[:, 0:np] += A_hi W_hi^T, D[:, np:2np] += A_hi W_lo^T in one instruction, then D[:, 0:np] += A_lo W_hi^T
              umma_tf32_ts_pred(d_addr, a_hi + 8 * jj, w_hi + 2 * jj, idesc2, acc, leader);
              umma_tf32_ts_pred(d_addr, a_lo + 8 * jj, w_hi + 2 * jj, idesc, 1u, leader);
            } else {
              umma_tf32_ts_pred(d_addr, a_hi + 8 * jj, w_hi + 2 * jj, idesc, acc, leader);
              umma_tf32_ts_pred(d_addr, a_lo + 8 * jj, w_hi + 2 * jj, idesc, 1u, leader);
              umma_tf32_ts_pred(d_addr, a_hi + 8 * jj, w_lo + 2 * jj, idesc, 1u, leader);
            }
          }
        }
        umma_commit_pred(&a_empty[stg], leader);   // arrives when the MMAs above have finished reading the stage
        if (++stg == a_stages) { stg = 0; ++round; }
      }
      umma_commit_pred(&acc_full[ab], leader);  // ... and when the whole tile's accumulator is complete
      if (trace != nullptr && lane == 0 && tl < 32) trace[1 * 64 + tl * 2 + 1] = clock64();
    }
  } else {
    // =========================== epilogue ===========================
    // one-time setup, part 2 (overlaps the first tile's loads): resident weights, bias, BatchNorm-on-load
    // parameters -> shared memory, published through setup_bar
    {
      const int et0 = tid - kProducerThreads;
      const int total16 = (2 * np * kp32) >> 2;
      const uint32_t w_addr = smem_u32(s.w);
      const float* wsrc = p.wpack + static_cast<size_t>(ck) * (2 * static_cast<size_t>(np) * kp32);
      for (int i = et0; i < total16; i += kEpilogueThreads) cp_async16(w_addr + i * 16u, wsrc + i * 4, 16);
      asm volatile("cp.async.commit_group;" ::: "memory");
      for (int i = et0; i < np; i += kEpilogueThreads) s.bias[i] = (p.bias != nullptr && i < n_loc) ? p.bias[col0 + i] : 0.f;
      if (p.a1_mean != nullptr) {
        const int k1r = (p.k1 + 3) & ~3;
        for (int i = et0; i < p.k1; i += kEpilogueThreads) {
          s.bn[i] = p.a1_mean[i]; s.bn[k1r + i] = p.a1_scale[i]; s.bn[2 * k1r + i] = p.a1_beta[i];
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // W: generic-proxy writes -> async proxy (MMA)
      mbar_arrive(setup_bar);
      if (!mbar_wait(setup_bar, 0u)) timed_out = true;
    }
    // thread = row (TMEM lane), 16 columns per tcgen05.ld; each lane stores 4 x 16 bytes of its own
    // row; BatchNorm column sums via a fixed-order butterfly over the warp's 32 rows.
    const int ew = warp - kProducerWarps;      // 0..7
    const int q = warp & 3, half = ew >> 2;    // TMEM lane quarter (hardware: warp id % 4), parity of the column blocks
    const int n_blocks = np >> 4;
    const bool vec = ((p.ldy & 3) == 0) && ((p.n_store & 3) == 0);
    const int et = tid - kProducerThreads;     // 0..255
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int ab = static_cast<int>(tl & 1);
      const int64_t tile = cta_c + tl * ctas_c;
      const int64_t row = tile * kRows + q * 32 + lane;
      const bool row_ok = row < p.m;
      float* yrow = p.y + (row_ok ? row : 0) * p.ldy;
      const float* rrow = p.residual != nullptr
          ? p.residual + (row_ok ? (p.a1_rows != nullptr ? static_cast<int64_t>(p.a1_rows[row]) : row) : 0) * p.ldr : nullptr;
      if (!mbar_wait(&acc_full[ab], static_cast<uint32_t>((tl >> 1) & 1))) timed_out = true;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (trace != nullptr && et == 0 && tl < 32) trace[2 * 64 + tl * 2] = clock64();
      float* csum = s.col_sum + (ab * 4 + q) * np;
      float* csq = s.col_sq + (ab * 4 + q) * np;
      if (s.stage != nullptr) {
        // Coalesced path (no residual / BatchNorm sums): 32 columns at a time through a per-warp
        // transpose buffer, so that a store instruction writes 4 rows x 128 contiguous bytes instead
        // of 32 rows x 16 bytes (8x fewer LSU wavefronts; the strided form bounded the kernel).
        // Padding columns need no masking: their W rows and bias entries are zero, so the accumulator
        // holds exact zeros there.  Everything that does not depend on the row is hoisted out of the
        // store loop -- the epilogue is a serial chain on two warps per scheduler, instruction count is
        // what bounds it.
        float* st = s.stage + ew * (32 * 36);
        const int n_dbl = (n_blocks + 1) >> 1;
        const int n_split = p.y2 != nullptr ? p.n_split : 0x7fffffff;
        const int64_t tile_row0 = tile * kRows + q * 32;
        const int rows_valid = static_cast<int>(p.m - tile_row0 < 32 ? p.m - tile_row0 : 32);
        const int c4 = lane & 7, rsub = lane >> 3;
        for (int cd = half; cd < n_dbl; cd += 2) {
          const bool etr = trace != nullptr && et == 0 && tl == 2 && (cd >> 1) < 5;   // debug: block timeline of one warp
          if (etr) trace[3 * 64 + 2 + (cd >> 1) * 4] = clock64();
          const uint32_t taddr = tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ab * acc_stride + cd * 32);
          float* strow = st + lane * 36;
          if (cd * 2 + 1 < n_blocks) {
            uint32_t r[32];
            tmem_ld32(taddr, r);
            if (dual) {   // second half of the accumulator: the hi * lo products
              uint32_t r2[32];
              tmem_ld32(taddr + np, r2);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4)
              *reinterpret_cast<uint4*>(strow + j4 * 4) = make_uint4(r[j4 * 4], r[j4 * 4 + 1], r[j4 * 4 + 2], r[j4 * 4 + 3]);
          } else {
            uint32_t r[16];
            tmem_ld16(taddr, r);
            if (dual) {
              uint32_t r2[16];
              tmem_ld16(taddr + np, r2);
#pragma unroll
              for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
            }
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
              *reinterpret_cast<uint4*>(strow + j4 * 4) = make_uint4(r[j4 * 4], r[j4 * 4 + 1], r[j4 * 4 + 2], r[j4 * 4 + 3]);
          }
          __syncwarp();
          if (etr) trace[3 * 64 + 3 + (cd >> 1) * 4] = clock64();
          const int col = cd * 32 + c4 * 4;
          float4 csum4 = make_float4(0.f, 0.f, 0.f, 0.f), csq4 = csum4;
          if (col < n_store_loc && col < np) {
            const float4 bv = *reinterpret_cast<const float4*>(s.bias + col);
            // split output: columns from n_split on go to the narrow tail array y2
            float* dst;
            int64_t ld;
            if (col < n_split) { ld = p.ldy; dst = p.y + (tile_row0 + rsub) * ld + col0 + col; }
            else { ld = p.ldy2; dst = p.y2 + (tile_row0 + rsub) * ld + (col - n_split); }
            const float* srow = st + rsub * 36 + c4 * 4;
            // residual (RadarPointGNNConv): the layer input, normalised on load like the A operand
            const bool has_res = p.residual != nullptr && col < n_loc;
            float4 rmu = make_float4(0.f, 0.f, 0.f, 0.f), rsc = make_float4(1.f, 1.f, 1.f, 1.f), rbe = rmu;
            if (has_res && p.res_mean != nullptr) {
              rmu = *reinterpret_cast<const float4*>(p.res_mean + col0 + col);
              rsc = *reinterpret_cast<const float4*>(p.res_scale + col0 + col);
              rbe = *reinterpret_cast<const float4*>(p.res_beta + col0 + col);
            }
            if (!has_res && p.bn_partial == nullptr) {
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                if (it * 4 + rsub < rows_valid) {
                  float4 v4 = *reinterpret_cast<const float4*>(srow + it * (4 * 36));
                  v4.x += bv.x; v4.y += bv.y; v4.z += bv.z; v4.w += bv.w;
                  *reinterpret_cast<float4*>(dst + it * 4 * ld) = v4;
                }
              }
            } else
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              if (it * 4 + rsub < rows_valid) {
                float4 v4 = *reinterpret_cast<const float4*>(srow + it * (4 * 36));
                v4.x += bv.x; v4.y += bv.y; v4.z += bv.z; v4.w += bv.w;
                if (has_res) {
                  const int64_t grow = tile_row0 + rsub + it * 4;
                  const int64_t rrow_i = p.a1_rows != nullptr ? static_cast<int64_t>(p.a1_rows[grow]) : grow;
                  float4 rv = *reinterpret_cast<const float4*>(p.residual + rrow_i * p.ldr + col0 + col);
                  if (p.res_mean != nullptr) {
                    rv.x = (rv.x - rmu.x) * rsc.x + rbe.x; rv.y = (rv.y - rmu.y) * rsc.y + rbe.y;
                    rv.z = (rv.z - rmu.z) * rsc.z + rbe.z; rv.w = (rv.w - rmu.w) * rsc.w + rbe.w;
                  }
                  if (p.res_relu) { rv.x = fmaxf(rv.x, 0.f); rv.y = fmaxf(rv.y, 0.f); rv.z = fmaxf(rv.z, 0.f); rv.w = fmaxf(rv.w, 0.f); }
                  v4.x += rv.x; v4.y += rv.y; v4.z += rv.z; v4.w += rv.w;
                }
                *reinterpret_cast<float4*>(dst + it * 4 * ld) = v4;
                csum4.x += v4.x; csum4.y += v4.y; csum4.z += v4.z; csum4.w += v4.w;
                csq4.x = fmaf(v4.x, v4.x, csq4.x); csq4.y = fmaf(v4.y, v4.y, csq4.y);
                csq4.z = fmaf(v4.z, v4.z, csq4.z); csq4.w = fmaf(v4.w, v4.w, csq4.w);
              }
            }
          }
          if (p.bn_partial != nullptr) {
            // column sums of the warp's 32 rows: the four row groups (lanes c4, c4 + 8, + 16, + 24) in a
            // fixed order -> deterministic
#pragma unroll
            for (int o = 8; o <= 16; o <<= 1) {
              csum4.x += __shfl_xor_sync(0xffffffffu, csum4.x, o); csum4.y += __shfl_xor_sync(0xffffffffu, csum4.y, o);
              csum4.z += __shfl_xor_sync(0xffffffffu, csum4.z, o); csum4.w += __shfl_xor_sync(0xffffffffu, csum4.w, o);
              csq4.x += __shfl_xor_sync(0xffffffffu, csq4.x, o); csq4.y += __shfl_xor_sync(0xffffffffu, csq4.y, o);
              csq4.z += __shfl_xor_sync(0xffffffffu, csq4.z, o); csq4.w += __shfl_xor_sync(0xffffffffu, csq4.w, o);
            }
            if (rsub == 0 && col < np) {
              *reinterpret_cast<float4*>(csum + col) = csum4;
              *reinterpret_cast<float4*>(csq + col) = csq4;
            }
          }
          if (etr) trace[3 * 64 + 4 + (cd >> 1) * 4] = clock64();
          __syncwarp();
        }
      } else
      for (int cb = half; cb < n_blocks; cb += 2) {
        uint32_t r[16];
        tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ab * acc_stride + cb * 16), r);
        if (dual) {   // second half of the accumulator: the hi * lo products
          uint32_t r2[16];
          tmem_ld16(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ab * acc_stride + np + cb * 16), r2);
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        float v[16];
        const bool full_block = cb * 16 + 16 <= n_loc;   // no padding columns inside this block
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 bv = *reinterpret_cast<const float4*>(s.bias + cb * 16 + j4 * 4);  // zero beyond n
          v[j4 * 4 + 0] = __uint_as_float(r[j4 * 4 + 0]) + bv.x;
          v[j4 * 4 + 1] = __uint_as_float(r[j4 * 4 + 1]) + bv.y;
          v[j4 * 4 + 2] = __uint_as_float(r[j4 * 4 + 2]) + bv.z;
          v[j4 * 4 + 3] = __uint_as_float(r[j4 * 4 + 3]) + bv.w;
        }
        if (!full_block) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cb * 16 + j >= n_loc) v[j] = 0.f;
        }
        if (rrow != nullptr && row_ok) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = cb * 16 + j;
            if (col < n_loc) {
              float rv = rrow[col0 + col];
              if (p.res_mean != nullptr) rv = (rv - p.res_mean[col0 + col]) * p.res_scale[col0 + col] + p.res_beta[col0 + col];
              if (p.res_relu) rv = fmaxf(rv, 0.f);
              v[j] += rv;
            }
          }
        }
        if (row_ok) {
          if (vec) {
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const int col = cb * 16 + j4 * 4;
              if (full_block || col < n_store_loc)
                *reinterpret_cast<float4*>(yrow + col0 + col) = make_float4(v[j4 * 4], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = cb * 16 + j;
              if (col < n_store_loc) yrow[col0 + col] = v[j];
            }
          }
        }
        if (p.bn_partial != nullptr) {
          float sv[16], sq[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { sv[j] = row_ok ? v[j] : 0.f; sq[j] = sv[j] * sv[j]; }
#pragma unroll
          for (int w = 8, bit = 16; w >= 1; w >>= 1, bit >>= 1) {
            // lanes with `bit` clear keep columns [0, w), the others keep [w, 2w)
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int j = 0; j < w; ++j) {
              const float send_s = upper ? sv[j] : sv[j + w], keep_s = upper ? sv[j + w] : sv[j];
              const float send_q = upper ? sq[j] : sq[j + w], keep_q = upper ? sq[j + w] : sq[j];
              sv[j] = keep_s + __shfl_xor_sync(0xffffffffu, send_s, bit);
              sq[j] = keep_q + __shfl_xor_sync(0xffffffffu, send_q, bit);
            }
          }
          sv[0] += __shfl_xor_sync(0xffffffffu, sv[0], 1);
          sq[0] += __shfl_xor_sync(0xffffffffu, sq[0], 1);
          if ((lane & 1) == 0) {
            const int col = cb * 16 + (((lane >> 4) & 1) << 3 | ((lane >> 3) & 1) << 2 | ((lane >> 2) & 1) << 1 | ((lane >> 1) & 1));
            csum[col] = sv[0];
            csq[col] = sq[0];
          }
        }
      }
      // accumulator drained: hand it back to the MMA issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&acc_empty[ab]);
      if (trace != nullptr && et == 0 && tl < 32) trace[2 * 64 + tl * 2 + 1] = clock64();
      if (p.bn_partial != nullptr) {
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the four quarters' column sums are in shared memory
        const float* cs = s.col_sum + ab * 4 * np;
        const float* cq = s.col_sq + ab * 4 * np;
        for (int c = et; c < n_loc; c += kEpilogueThreads) {
          const double sum = (static_cast<double>(cs[c]) + static_cast<double>(cs[np + c])) +
                             (static_cast<double>(cs[2 * np + c]) + static_cast<double>(cs[3 * np + c]));
          const double sq2 = (static_cast<double>(cq[c]) + static_cast<double>(cq[np + c])) +
                             (static_cast<double>(cq[2 * np + c]) + static_cast<double>(cq[3 * np + c]));
          // channel-major: partial[ch * T + tile] (sums), partial[(n + ch) * T + tile] (squares), so that the
          // finalisation reads every channel's partials contiguously
          p.bn_partial[static_cast<int64_t>(col0 + c) * n_tiles + tile] = sum;
          p.bn_partial[static_cast<int64_t>(p.n + col0 + c) * n_tiles + tile] = sq2;
        }
      }
    }
  }

  if (trace != nullptr && tid == 0) trace[3 * 64 + 1] = clock64();
  // a barrier that never completes is a kernel bug: fail loudly (sticky launch error at the caller's next
  // synchronisation) instead of handing back a partly computed result
  if (timed_out) {
    if (p.status != nullptr) atomicExch(p.status, RGNN_ERR_CUDA);
    __trap();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (p.trace != nullptr && tid == 0 && blockIdx.x < 192) {   // debug: all roles of the CTA are done
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.trace[256 + 2 * blockIdx.x + 1] = static_cast<long long>(gt);
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols) : "memory");
  }
}

size_t smem_bytes_for(int np, int kp, int raw_slots, int staged = 0) {
  const size_t kp32 = (static_cast<size_t>(kp) + 31) & ~static_cast<size_t>(31);
  return sizeof(float) * (2 * np * kp32 + static_cast<size_t>(raw_slots) * kABufFloats + 16 * np + np + 3 * kp32 +
                          2 * kBarCount + 4 + (staged ? kStageFloats : 0)) + kMaxPanels * sizeof(PanelInfo) + 64;
}

// hi*hi and hi*lo in one MMA (N = 2 np <= 256) when TMEM still holds two such accumulators + 2 A stages
bool pick_dual(int np) { return 2 * np <= 256 && 4 * np + 2 * kAStageCols <= 512; }

// TMEM: two accumulators of np (dual: 2 np) columns + A stages of 64 columns each
int pick_a_stages(int np) {
  const int st = (512 - (pick_dual(np) ? 4 : 2) * np) / kAStageCols;
  return st > kMaxAStages ? kMaxAStages : st;
}

// raw ring slots that fit beside W (0: the contraction does not fit this kernel)
int pick_raw_slots(int np, int kp, int staged) {
  if (pick_a_stages(np) < 1) return 0;
  if (kp > kMaxPanels * kKc) return 0;
  for (int sl = kMaxRaw; sl >= 2; --sl)
    if (smem_bytes_for(np, kp, sl, staged) <= 227 * 1024) return sl;
  return 0;
}

}  // namespace

// Output-column chunking: the fewest chunks whose slice of W (hi + lo images, np x kp each) fits shared memory
// next to the coalesced epilogue's transpose buffers and a raw ring of at least three panels; failing that,
// the fewest chunks that fit at all.  {0, 0}: the contraction does not fit this kernel.
TcChunking tc_chunking(const TcGemmShape& sh) {
  TcChunking c{0, 0};
  if (sh.n < 1) return c;
  const int kp = tc_padded_k(sh);
  for (int pass = 0; pass < 2; ++pass) {
    for (int nc = 1; nc <= 16; ++nc) {
      const int np = tc_padded_n((sh.n + nc - 1) / nc);
      if (np > 256) continue;
      const bool ok = pass == 0 ? pick_raw_slots(np, kp, 1) >= 3 : pick_raw_slots(np, kp, 0) >= 2;
      if (ok) { c.n_chunks = (sh.n + np - 1) / np; c.np = np; return c; }
    }
  }
  return c;
}

bool tc_gemm_supported(const TcGemmShape& sh) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RGNN_DISABLE_TCGEN05");
    enabled = (e != nullptr && e[0] == '1') ? 0 : 1;
  }
  if (!enabled) return false;
  if (sh.k1 < 4 || sh.k1 % 4 != 0 || sh.k2 % 4 != 0 || sh.n < 1) return false;
  if (sh.kt % 4 != 0) return false;
  return tc_chunking(sh).n_chunks > 0;
}

size_t tc_pack_floats(const TcGemmShape& sh) {
  const TcChunking c = tc_chunking(sh);
  return 2 * static_cast<size_t>(c.n_chunks) * c.np * tc_padded_k(sh);
}

int tc_fold_weights(const float* w_m, int64_t ld_m, const float* w_t, int64_t ld_t, int c_out, int p, int c,
                    float* w_fold, cudaStream_t stream) {
  RGNN_PROFILE("weight_prep", stream);
  fold_weights_kernel<<<div_up(c_out * c, 256), 256, 0, stream>>>(w_m, ld_m, w_t, ld_t, c_out, p, c, w_fold);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int tc_pack_weights(const TcWeightBlocks& blocks, const TcGemmShape& sh, float* wpack, cudaStream_t stream) {
  const TcChunking c = tc_chunking(sh);
  if (c.n_chunks < 1) return RGNN_ERR_UNSUPPORTED;
  const int kp = tc_padded_k(sh);
  RGNN_PROFILE("weight_prep", stream);
  pack_weights_kernel<<<div_up(static_cast<int64_t>(c.n_chunks) * c.np * kp, 256), 256, 0, stream>>>(blocks, c.np, c.n_chunks, kp, wpack);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

// debug: device buffer of 4 * 64 int64 receiving CTA 0's timeline of a later launch whose tag matches;
// "tag@k" skips the first k matching launches (k-th layer of a fused forward)
static long long* g_trace_buffer = nullptr;
static char g_trace_tag[64] = "";
static int g_trace_skip = 0;
extern "C" void rgnn_debug_trace_node_gemm(void* device_buffer, const char* tag) {
  g_trace_buffer = static_cast<long long*>(device_buffer);
  snprintf(g_trace_tag, sizeof(g_trace_tag), "%s", tag != nullptr ? tag : "");
  g_trace_skip = 0;
  char* at = strchr(g_trace_tag, '@');
  if (at != nullptr) { g_trace_skip = atoi(at + 1); *at = 0; }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// row-major [rows, cols] fp32 tensor (row stride ld floats), box = one panel: 32 floats x 128 rows
static bool encode_panel_map(CUtensorMap* tm, const float* base, int64_t rows, int32_t cols, int64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr || base == nullptr || rows <= 0 || cols <= 0) return false;
  if (reinterpret_cast<uintptr_t>(base) % 16 != 0 || (ld * 4) % 16 != 0) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {32, 128};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_tc_gemm(TcGemmParams p, const char* tag, cudaStream_t stream) {
  if (p.m <= 0) return RGNN_OK;
  if (g_trace_buffer != nullptr && strcmp(tag, g_trace_tag) == 0 && g_trace_skip-- <= 0) { p.trace = g_trace_buffer; g_trace_buffer = nullptr; }
  p.kp = tc_seg_pad(p.k1) + tc_seg_pad(p.k2) + tc_seg_pad(p.kt) + tc_seg_pad(p.k3);
  {
    TcGemmShape sh; sh.k1 = p.k1; sh.k2 = p.k2; sh.kt = p.kt; sh.k3 = p.k3; sh.n = p.n;
    const TcChunking c = tc_chunking(sh);
    if (c.n_chunks < 1) return RGNN_ERR_UNSUPPORTED;
    p.np = c.np; p.n_chunks = c.n_chunks;
  }
  if (p.n_chunks > 1 && p.y2 != nullptr) return RGNN_ERR_UNSUPPORTED;   // split outputs are single-chunk shapes
  if (p.n_store < p.n) p.n_store = p.n;
  p.a_stages = pick_a_stages(p.np);
  p.dual = pick_dual(p.np) ? 1 : 0;
  if (p.a2_panel_major && ((p.k2 & 31) != 0 || reinterpret_cast<uintptr_t>(p.a2) % 16 != 0)) return RGNN_ERR_INVALID_ARGUMENT;
  // coalesced (staged) epilogue whenever the rows allow 16-byte accesses and the transpose buffers still
  // leave room for the raw ring
  const bool res_ok = p.residual == nullptr ||
                      ((p.ldr & 3) == 0 && (p.n & 3) == 0 && reinterpret_cast<uintptr_t>(p.residual) % 16 == 0 &&
                       (p.res_mean == nullptr || (reinterpret_cast<uintptr_t>(p.res_mean) % 16 == 0 &&
                                                  reinterpret_cast<uintptr_t>(p.res_scale) % 16 == 0 &&
                                                  reinterpret_cast<uintptr_t>(p.res_beta) % 16 == 0)));
  p.staged_epilogue = (res_ok && (p.ldy & 3) == 0 && (p.n_store & 3) == 0 && reinterpret_cast<uintptr_t>(p.y) % 16 == 0 &&
                       pick_raw_slots(p.np, p.kp, 1) >= 2) ? 1 : 0;
  p.raw_slots = pick_raw_slots(p.np, p.kp, p.staged_epilogue);
  if (p.raw_slots < 2 || p.a_stages < 1) return RGNN_ERR_UNSUPPORTED;
  // converter groups: ring slots and TMEM stages in multiples of the group count (see the converter role)
  // Three groups when the ring and TMEM allow it (RGNN_GEMM_GROUPS forces 1 / 2 / 3 for experiments: two
  // groups over four ring slots instead of three over three measured the same on the update contraction).
  {
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("RGNN_GEMM_GROUPS"); forced = e != nullptr ? atoi(e) : 0; }
    int groups = (p.raw_slots >= 3 && p.a_stages >= 3) ? 3 : ((p.raw_slots >= 2 && p.a_stages >= 2) ? 2 : 1);
    if (forced >= 1 && forced <= groups) groups = forced;
    p.conv_groups = groups;
  }
  p.raw_slots = p.raw_slots / p.conv_groups * p.conv_groups;
  p.a_stages = p.a_stages / p.conv_groups * p.conv_groups;
  if (p.y2 != nullptr && (!p.staged_epilogue || (p.n_split & 3) != 0 || (p.ldy2 & 3) != 0)) return RGNN_ERR_UNSUPPORTED;
  // row-major operands go through TMA when they are plain (no gather map) and 16-byte aligned
  static int tma_enabled = -1;
  if (tma_enabled < 0) { const char* e = getenv("RGNN_DISABLE_TMA"); tma_enabled = (e != nullptr && e[0] == '1') ? 0 : 1; }
  memset(&p.tm_a1, 0, sizeof(p.tm_a1));
  memset(&p.tm_at, 0, sizeof(p.tm_at));
  p.a1_tma = (tma_enabled && p.a1_rows == nullptr && encode_panel_map(&p.tm_a1, p.a1, p.m, p.k1, p.lda1)) ? 1 : 0;
  p.at_tma = (tma_enabled && p.at != nullptr && p.kt > 0 && encode_panel_map(&p.tm_at, p.at, p.m, p.kt, p.ldat)) ? 1 : 0;
  const size_t smem = smem_bytes_for(p.np, p.kp, p.raw_slots, p.staged_epilogue);
  static bool configured[kMaxDevices] = {};
  RGNN_CUDA_CHECK(opt_in_dynamic_smem(node_gemm_kernel, configured, 227 * 1024));
  const int64_t tiles = (p.m + kRows - 1) / kRows;
  // CTA b works on chunk b % n_chunks: a multiple of n_chunks CTAs, at most one per SM
  int per_chunk = sm_count() / p.n_chunks;
  if (per_chunk < 1) per_chunk = 1;
  if (tiles < per_chunk) per_chunk = static_cast<int>(tiles);
  const int grid = per_chunk * p.n_chunks;
  RGNN_PROFILE(tag, stream);
  node_gemm_kernel<<<grid, kGemmThreads, smem, stream>>>(p);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int64_t tc_tiles(int64_t m) { return (m + kRows - 1) / kRows; }

}  // namespace rgnn

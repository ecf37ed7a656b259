// scan.cu -- exclusive prefix sums used by the cell list, the row pointers and the CSC build.
// Three-phase block scan (local scan + block totals, recursive scan of the totals, add).
#include "common.cuh"

namespace rgnn {
namespace {

constexpr int kScanThreads = 512;
constexpr int kItems = 4;
constexpr int kTile = kScanThreads * kItems;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const TIn* in, TOut* out, int64_t n, TOut* tile_sums) {
  // in / out may alias (in-place scan of the tile totals): no __restrict__ here
  // scans positions [0, n]; position n reads as 0 so that out[n] is the grand total
  __shared__ TOut warp_totals[kScanThreads / 32];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kTile + static_cast<int64_t>(threadIdx.x) * kItems;
  TOut v[kItems];
  TOut local = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t p = base + i;
    v[i] = (p < n) ? static_cast<TOut>(in[p]) : TOut(0);
    local += v[i];
  }
  // inclusive warp scan of the per-thread sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  TOut inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    TOut t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_totals[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    TOut w = (lane < kScanThreads / 32) ? warp_totals[lane] : TOut(0);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      TOut t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) warp_totals[lane] = w;  // inclusive
  }
  __syncthreads();
  TOut prefix = (warp > 0 ? warp_totals[warp - 1] : TOut(0)) + (inc - local);
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t p = base + i;
    if (p <= n) out[p] = prefix;
    prefix += v[i];
  }
  if (threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = warp_totals[kScanThreads / 32 - 1];
}

template <typename TOut>
__global__ void add_tile_offsets_kernel(TOut* __restrict__ out, int64_t n, const TOut* __restrict__ tile_offsets) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * kTile + threadIdx.x;
  const TOut off = tile_offsets[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    int64_t q = p + static_cast<int64_t>(i) * kScanThreads;
    if (q <= n) out[q] += off;
  }
}

template <typename TIn, typename TOut>
int scan_impl(const TIn* in, TOut* out, int64_t n, TOut* scratch, cudaStream_t stream) {
  const int64_t count = n + 1;
  const unsigned tiles = div_up(count, kTile);
  scan_tiles_kernel<TIn, TOut><<<tiles, kScanThreads, 0, stream>>>(in, out, n, scratch);
  RGNN_LAUNCH_CHECK();
  if (tiles > 1) {
    // scan the tile totals in place (exclusive) and add them back
    TOut* next = scratch + align_up(tiles + 1, 64);
    RGNN_RETURN_IF_ERROR((scan_impl<TOut, TOut>(scratch, scratch, tiles, next, stream)));
    add_tile_offsets_kernel<TOut><<<tiles, kScanThreads, 0, stream>>>(out, n, scratch);
    RGNN_LAUNCH_CHECK();
  }
  return RGNN_OK;
}

// ---- single-pass scan (decoupled look-back) -----------------------------------------------------------------
// One launch instead of three (tile scan, scan of the tile totals, add): a tile publishes its aggregate, then its
// inclusive prefix, in one 64-bit status word (flag in the high half); a later tile sums the words of its
// predecessors until it meets an inclusive one.  Tile ids are handed out by an atomic counter, so a tile only ever
// waits for tiles that have already started.  status[tiles] and the counter must be zero at launch.
constexpr unsigned long long kFlagAggregate = 1ull << 32, kFlagPrefix = 2ull << 32;

__global__ void __launch_bounds__(kScanThreads)
chained_scan_kernel(const int32_t* in, int32_t* out, int64_t n, unsigned long long* status, int32_t* counter) {
  __shared__ int32_t warp_totals[kScanThreads / 32];
  __shared__ int32_t tile_s, prefix_s;
  if (threadIdx.x == 0) tile_s = atomicAdd(counter, 1);
  __syncthreads();
  const int tile = tile_s;
  const int64_t base = static_cast<int64_t>(tile) * kTile + static_cast<int64_t>(threadIdx.x) * kItems;
  int32_t v[kItems], local = 0;
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t p = base + i;
    v[i] = p < n ? in[p] : 0;
    local += v[i];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int32_t inc = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_totals[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int32_t w = lane < kScanThreads / 32 ? warp_totals[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) warp_totals[lane] = w;  // inclusive
    const int32_t aggregate = __shfl_sync(0xffffffffu, w, kScanThreads / 32 - 1);
    volatile unsigned long long* st = status;
    if (lane == 0) st[tile] = (tile == 0 ? kFlagPrefix : kFlagAggregate) | static_cast<unsigned>(aggregate);
    int32_t exclusive = 0;
    for (int idx = tile - 1; idx >= 0; idx -= 32) {
      const int j = idx - lane;
      unsigned long long word = kFlagPrefix;   // lanes before tile 0: an inclusive prefix of 0 (never the first hit below unless real ones are all aggregates)
      if (j >= 0) {
        do { word = st[j]; } while ((word >> 32) == 0ull);
      }
      const unsigned is_prefix = __ballot_sync(0xffffffffu, (word >> 32) == 2ull);
      const int first = __ffs(is_prefix) - 1;                  // nearest predecessor holding an inclusive prefix
      int32_t val = (first < 0 || lane <= first) ? static_cast<int32_t>(word & 0xffffffffull) : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
      exclusive += val;
      if (first >= 0) break;
    }
    if (lane == 0) {
      if (tile > 0) st[tile] = kFlagPrefix | static_cast<unsigned>(exclusive + aggregate);
      prefix_s = exclusive;
    }
  }
  __syncthreads();
  int32_t prefix = prefix_s + (warp > 0 ? warp_totals[warp - 1] : 0) + (inc - local);
#pragma unroll
  for (int i = 0; i < kItems; ++i) {
    const int64_t p = base + i;
    if (p <= n) out[p] = prefix;
    prefix += v[i];
  }
}

}  // namespace

size_t scan_scratch_ints(int64_t n) {
  size_t total = 0;
  int64_t count = n + 1;
  while (true) {
    int64_t tiles = (count + kTile - 1) / kTile;
    total += align_up(tiles + 1, 64);
    if (tiles <= 1) break;
    count = tiles + 1;
  }
  // + the single-pass scan's status words (one 64-bit word per tile) and its tile counter
  return total + 64 + 2 * static_cast<size_t>((n + 1 + kTile - 1) / kTile) + 64;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t stream) {
  // totals fit 31 bits here (point / edge counts): the single-pass scan packs value and flag into one word
  const unsigned tiles = div_up(n + 1, kTile);
  if (in == out) return scan_impl<int32_t, int32_t>(in, out, n, scratch, stream);   // the chained kernel reads ahead of its writes
  unsigned long long* status = reinterpret_cast<unsigned long long*>(scratch + (reinterpret_cast<uintptr_t>(scratch) % 8 ? 1 : 0));
  int32_t* counter = reinterpret_cast<int32_t*>(status + tiles);
  RGNN_CUDA_CHECK(cudaMemsetAsync(status, 0, sizeof(unsigned long long) * tiles + sizeof(int32_t), stream));
  chained_scan_kernel<<<tiles, kScanThreads, 0, stream>>>(in, out, n, status, counter);
  RGNN_LAUNCH_CHECK();
  return RGNN_OK;
}

int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, int64_t* scratch, cudaStream_t stream) {
  return scan_impl<int32_t, int64_t>(in, out, n, scratch, stream);
}

}  // namespace rgnn

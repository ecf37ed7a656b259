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
  return total + 64;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* scratch, cudaStream_t stream) {
  return scan_impl<int32_t, int32_t>(in, out, n, scratch, stream);
}

int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, int64_t* scratch, cudaStream_t stream) {
  return scan_impl<int32_t, int64_t>(in, out, n, scratch, stream);
}

}  // namespace rgnn

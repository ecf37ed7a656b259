// node_gemm.cuh -- internal interface of the tcgen05 node contraction (node_gemm.cu).
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace rgnn {

// logical shape of one contraction: K = k1 + k2 + kt + k3 (kt: narrow tail columns of the second
// operand kept in their own array; k3 = 0 or k1: rowscale * a1), N outputs
struct TcGemmShape {
  int32_t k1 = 0, k2 = 0, kt = 0, k3 = 0, n = 0;
};

inline int tc_padded_n(int n) { return (n + 15) & ~15; }
// every K segment is padded to whole 32-float chunks (one 128-byte swizzled panel column)
inline int tc_seg_pad(int k) { return (k + 31) & ~31; }
inline int tc_padded_k(const TcGemmShape& s) { return tc_seg_pad(s.k1) + tc_seg_pad(s.k2) + tc_seg_pad(s.kt) + tc_seg_pad(s.k3); }

// one block of source weights copied into the packed image: rows x cols of a row-major matrix
// (row stride ld) placed at K offset k_offset
struct TcWeightBlock {
  const float* src;
  int64_t ld;
  int32_t rows, cols, k_offset;
  int32_t accumulate;  // add onto what earlier blocks put at these positions instead of replacing it
};
struct TcWeightBlocks {
  TcWeightBlock block[4];
  int32_t count;
};

struct TcGemmParams {
  const float* a1 = nullptr; int64_t lda1 = 0; int32_t k1 = 0;   // 16-byte aligned rows, k1 % 4 == 0
  const int32_t* a1_rows = nullptr;                              // optional gather map for a1 / residual rows
  const float* a2 = nullptr; int64_t lda2 = 0; int32_t k2 = 0;   // optional, k2 % 4 == 0
  int32_t a2_panel_major = 0;   // a2 is stored tile by tile as k2 / 32 swizzled [128 x 32] panels (tc_panel_offset): one
                                // contiguous 16 KB block per panel, fetched with a single bulk copy (k2 % 32 == 0)
  const float* at = nullptr; int64_t ldat = 0; int32_t kt = 0;   // optional narrow tail of a2 (kt % 4 == 0, kt <= ldat)
  int32_t k3 = 0;                                                // 0, or k1: last segment rowscale * a1
  const int32_t* csc_ptr = nullptr; int32_t rowscale_mode = 0;   // 1: in-degree > 0, 2: in-degree
  const float* a1_mean = nullptr; const float* a1_scale = nullptr; const float* a1_beta = nullptr;
  int32_t relu_a1 = 0, relu_a2 = 0;
  const float* wpack = nullptr;                                  // tc_pack_weights image
  int32_t n = 0, np = 0, n_chunks = 1, kp = 0, a_stages = 0, raw_slots = 0, staged_epilogue = 0, dual = 0, conv_groups = 0;
  const float* bias = nullptr;
  const float* residual = nullptr; int64_t ldr = 0;
  const float* res_mean = nullptr; const float* res_scale = nullptr; const float* res_beta = nullptr;
  int32_t res_relu = 0;
  float* y = nullptr; int64_t ldy = 0; int32_t n_store = 0;      // columns written (>= n; extras are zeros)
  float* y2 = nullptr; int64_t ldy2 = 0; int32_t n_split = 0;    // optional: columns >= n_split go to y2[:, col - n_split]
  double* bn_partial = nullptr;                                  // [2][n][tc_tiles(m)] column sums / squares per tile
  int64_t m = 0;
  int32_t* status = nullptr;                                     // device flag set if a barrier wait timed out
  long long* trace = nullptr;                                    // debug timeline of CTA 0 (node_gemm.cu)
  // TMA descriptors (filled by launch_tc_gemm): a1 / a_tail as 2-D row-major tensors, box = one panel
  // (32 floats x 128 rows, 128-byte swizzle, zero fill outside the tensor)
  int32_t a1_tma = 0, at_tma = 0;
  alignas(64) CUtensorMap tm_a1;
  alignas(64) CUtensorMap tm_at;
};

// float offset of element (row, col) of a panel-major operand with `panels` 32-float panels per 128-row tile
__host__ __device__ inline int64_t tc_panel_offset(int64_t row, int col, int panels) {
  const int rl = static_cast<int>(row & 127), c = (col & 31) >> 2;
  return ((row >> 7) * panels + (col >> 5)) * 4096 + rl * 32 + ((c ^ (rl & 7)) << 2) + (col & 3);
}
inline size_t tc_panel_major_floats(int64_t rows, int k) { return static_cast<size_t>((rows + 127) / 128) * 128 * k; }

// how the N output columns are split into chunks of np (padded) columns, one chunk per CTA (node_gemm.cu)
struct TcChunking { int32_t n_chunks, np; };
TcChunking tc_chunking(const TcGemmShape& sh);
bool tc_gemm_supported(const TcGemmShape& sh);
size_t tc_pack_floats(const TcGemmShape& sh);
int64_t tc_tiles(int64_t m);
int tc_fold_weights(const float* w_m, int64_t ld_m, const float* w_t, int64_t ld_t, int c_out, int p, int c,
                    float* w_fold, cudaStream_t stream);
int tc_pack_weights(const TcWeightBlocks& blocks, const TcGemmShape& sh, float* wpack, cudaStream_t stream);
int launch_tc_gemm(TcGemmParams p, const char* tag, cudaStream_t stream);

}  // namespace rgnn

// Warp-cooperative selection primitives shared by the fused encoder epilogue, the dense
// candidate kernel and the merge/select kernel.
#pragma once
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace qsae {

// Compact one survivor buffer `rb` (n <= kCandCap entries {float bits, column}, insertion order
// == ascending column) down to its k largest values, in place, all 32 lanes participating.
// Ties at the k-th value keep the earliest entries (lowest columns). Entry order is preserved.
// Returns the new count; *thr_out is the k-th largest value (a valid strict-greater threshold
// for later columns). Requires n > k.
__device__ __forceinline__ int warp_compact_row(uint2* rb, int n, int k, int lane, float* thr_out) {
  constexpr int PER_LANE = kCandCap / 32;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint32_t key[PER_LANE], col[PER_LANE];
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const int e = i * 32 + lane;
    if (e < n) {
      const uint2 t = rb[e];
      key[i] = float_to_key(__uint_as_float(t.x));
      col[i] = t.y;
    } else {
      key[i] = 0u;  // below every real key
      col[i] = 0u;
    }
  }
  // largest T with count(key >= T) >= k: bit-serial bisection, stops early once exactly k remain
  uint32_t T = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t probe = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < PER_LANE; ++i) c += (key[i] >= probe) ? 1 : 0;
    c = __reduce_add_sync(full, c);
    if (c >= k) T = probe;
    if (c == k) break;
  }
  int c_gt = 0;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) c_gt += (key[i] > T) ? 1 : 0;
  c_gt = __reduce_add_sync(full, c_gt);
  int eq_budget = k - c_gt;  // entries equal to T that may stay
  int out = 0;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) {
    const bool gt = key[i] > T;
    const bool eq = (key[i] == T) && (i * 32 + lane < n);
    const unsigned eq_b = __ballot_sync(full, eq);
    const bool keep = gt || (eq && (__popc(eq_b & lt_mask) < eq_budget));
    eq_budget = max(0, eq_budget - __popc(eq_b));
    const unsigned keep_b = __ballot_sync(full, keep);
    if (keep)
      rb[out + __popc(keep_b & lt_mask)] =
          make_uint2(__float_as_uint(key_to_float(key[i])), col[i]);
    out += __popc(keep_b);
  }
  *thr_out = key_to_float(T);
  return out;
}

// composite 64-bit sort key: larger == better (higher value, then lower column)
__device__ __forceinline__ uint64_t make_sort_key(float v, uint32_t col) {
  return (static_cast<uint64_t>(float_to_key(v)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - col);
}
__device__ __forceinline__ float sort_key_value(uint64_t k) {
  return key_to_float(static_cast<uint32_t>(k >> 32));
}
__device__ __forceinline__ uint32_t sort_key_col(uint64_t k) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(k);
}

}  // namespace qsae

// Warp-cooperative selection primitives shared by the fused encoder epilogue, the dense
// candidate kernel and the merge/select kernel.
#pragma once
#include "kernels.h"
#include "ptx_sm100.cuh"

namespace qsae {

// Cut one survivor buffer `rb` (n entries {float bits, column} in global memory, insertion order
// == ascending column) down to its k largest values, in place, all 32 lanes participating.
// Ties at the k-th value keep the earliest entries (lowest columns); entry order is preserved.
// Returns the new count (k when n > k); *thr_out is the k-th largest value. Bit-serial bisection
// over the monotone integer image of the floats, re-reading the (L1/L2 resident) buffer per bit:
// this is the rare overflow path, not the steady state.
static __device__ __noinline__ int warp_compact_row_generic(uint2* rb, int n, int k, int lane,
                                                     float* thr_out) {
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  if (n <= k) {
    *thr_out = -INFINITY;
    return n;
  }
  // largest T with count(key >= T) >= k  ==  the k-th largest key
  uint32_t T = 0u;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t probe = T | (1u << bit);
    int c = 0;
    for (int e = lane; e < n; e += 32) c += (float_to_key(__uint_as_float(rb[e].x)) >= probe) ? 1 : 0;
    c = __reduce_add_sync(full, c);
    if (c >= k) T = probe;
    if (c == k) break;
  }
  int c_gt = 0;
  for (int e = lane; e < n; e += 32) c_gt += (float_to_key(__uint_as_float(rb[e].x)) > T) ? 1 : 0;
  c_gt = __reduce_add_sync(full, c_gt);
  int eq_budget = k - c_gt;  // entries equal to T that may stay
  int out = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += 32) {
    const int e = base + lane;
    uint2 t = make_uint2(0u, 0u);
    if (e < n) t = rb[e];
    const uint32_t key = float_to_key(__uint_as_float(t.x));
    const bool gt = (e < n) && (key > T);
    const bool eq = (e < n) && (key == T);
    const unsigned eq_b = __ballot_sync(full, eq);
    const bool keep = gt || (eq && (__popc(eq_b & lt_mask) < eq_budget));
    eq_budget = max(0, eq_budget - __popc(eq_b));
    const unsigned keep_b = __ballot_sync(full, keep);
    if (keep) rb[out + __popc(keep_b & lt_mask)] = t;
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

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

// Block-cooperative radix select over n distinct 64-bit composite keys, key_at(e) for e in [0, n):
// returns T such that exactly min(k, n) keys satisfy key >= T. Most-significant-digit first, 8 bits
// per pass, every pass one sweep of all threads over the keys (shared-memory histogram); stops as soon
// as a digit bucket holds exactly the number of keys still wanted (distinct values: <= 4 passes over the
// value half, the column half only breaks ties). Any k (the warp bisections handle k <= kMaxK only).
// hist: 256 ints of shared memory; ctl: 3 ints. Must be called by every thread of the block.
template <typename KeyAt>
__device__ __forceinline__ uint64_t block_radix_select(KeyAt key_at, int n, int k, int* hist, int* ctl) {
  if (n <= k) return 0ull;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  uint64_t prefix = 0ull, mask = 0ull;
  int need = k;
#pragma unroll 1
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int b = threadIdx.x; b < 256; b += blockDim.x) hist[b] = 0;
    __syncthreads();
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const uint64_t key = key_at(e);
      if ((key & mask) == prefix) atomicAdd(&hist[static_cast<int>(key >> shift) & 0xFF], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      // lane l owns bins [8l, 8l + 8); walk from the top bin down to the one holding the need-th key
      int mine = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) mine += hist[lane * 8 + b];
      int incl = mine;  // inclusive suffix sum over lanes (lane 31 = top bins)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_down_sync(full, incl, o);
        if (lane + o < 32) incl += t;
      }
      const int above = incl - mine;
      if (above < need && need <= incl) {
        int cum = above;
#pragma unroll 1
        for (int b = 7; b >= 0; --b) {
          const int h = hist[lane * 8 + b];
          if (cum + h >= need) {
            ctl[0] = lane * 8 + b;
            ctl[1] = need - cum;
            ctl[2] = h;
            break;
          }
          cum += h;
        }
      }
    }
    __syncthreads();
    const int bin = ctl[0], bucket = ctl[2];
    need = ctl[1];
    prefix |= static_cast<uint64_t>(bin) << shift;
    mask |= 0xFFull << shift;
    __syncthreads();   // ctl / hist are rewritten by the next pass
    if (bucket == need) break;   // the whole bucket is wanted: the low bits of T stay zero
  }
  return prefix;
}

// Block-wide bitonic sort (descending) of sel[0, ksort), ksort a power of two, in shared memory.
__device__ __forceinline__ void block_bitonic_desc(uint64_t* sel, int ksort) {
  for (int size = 2; size <= ksort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (ksort >> 1); t += blockDim.x) {
        const int pos = ((t / stride) * (stride << 1)) + (t % stride);
        const int partner = pos + stride;
        const bool desc = (pos & size) == 0;
        const uint64_t a = sel[pos], b = sel[partner];
        if ((a < b) == desc) {
          sel[pos] = b;
          sel[partner] = a;
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace qsae

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
// per pass, every pass one sweep of all threads over the keys (shared-memory histogram). The first
// sweep finds the leading bits all keys share (survivors of one row sit in a narrow value range: sign
// and exponent are common) so that no pass piles every key onto one histogram bin; the select stops
// as soon as a digit bucket holds exactly the number of keys still wanted (distinct values: <= 3-4
// passes, the column half only breaks ties). Any k (the warp bisections handle k <= kMaxK only).
// hist: 256 ints of shared memory; ctl: 4 ints. Must be called by every thread of the block.
template <typename KeyAt>
__device__ __forceinline__ uint64_t block_radix_select(KeyAt key_at, int n, int k, int* hist, int* ctl) {
  if (n <= k) return 0ull;
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // ---- shared leading bits: AND / OR of all keys
  unsigned* u = reinterpret_cast<unsigned*>(hist);
  if (threadIdx.x == 0) { u[0] = 0xFFFFFFFFu; u[1] = 0xFFFFFFFFu; u[2] = 0u; u[3] = 0u; }
  __syncthreads();
  {
    uint64_t a = ~0ull, o = 0ull;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      const uint64_t key = key_at(e);
      a &= key;
      o |= key;
    }
    unsigned a_hi = static_cast<unsigned>(a >> 32), a_lo = static_cast<unsigned>(a);
    unsigned o_hi = static_cast<unsigned>(o >> 32), o_lo = static_cast<unsigned>(o);
    a_hi = __reduce_and_sync(full, a_hi); a_lo = __reduce_and_sync(full, a_lo);
    o_hi = __reduce_or_sync(full, o_hi); o_lo = __reduce_or_sync(full, o_lo);
    if (lane == 0) {
      atomicAnd(&u[0], a_hi); atomicAnd(&u[1], a_lo);
      atomicOr(&u[2], o_hi); atomicOr(&u[3], o_lo);
    }
  }
  __syncthreads();
  const uint64_t all_and = (static_cast<uint64_t>(u[0]) << 32) | u[1];
  const uint64_t all_or = (static_cast<uint64_t>(u[2]) << 32) | u[3];
  __syncthreads();   // hist is reused below
  const uint64_t diff = all_and ^ all_or;   // != 0: n > k >= 1 distinct keys
  const int top = 63 - __clzll(static_cast<long long>(diff));
  int shift = top - 7;
  if (shift < 0) shift = 0;
  // bits above the first digit are common to all keys
  uint64_t mask = (shift + 8 >= 64) ? 0ull : (~0ull << (shift + 8));
  uint64_t prefix = all_and & mask;
  int need = k;
#pragma unroll 1
  for (;;) {
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
    if (bucket == need || shift == 0) break;   // the whole bucket is wanted: the low bits of T stay zero
    shift = shift >= 8 ? shift - 8 : 0;        // a last partial digit re-reads a few decided bits (harmless)
  }
  return prefix;
}

// Block-wide bitonic sort (descending) of sel[0, ksort), ksort a power of two, in shared memory.
// Each warp owns a contiguous chunk of ksort / nwarps elements: every step whose exchange distance
// stays inside a chunk runs warp-locally (__syncwarp), only the long strides synchronise the block
// (ksort = 4096 on 32 warps: 15 block barriers instead of 78).
__device__ __forceinline__ void block_bitonic_desc(uint64_t* sel, int ksort) {
  const int nthreads = blockDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = nthreads >> 5;
  int chunk = ksort / nwarps;              // elements per warp (power of two); < 64: plain block steps only
  if (chunk < 64) chunk = 0;
  const int half = ksort >> 1;
  for (int size = 2; size <= ksort; size <<= 1) {
    int stride = size >> 1;
    for (; stride > 0 && 2 * stride > chunk; stride >>= 1) {   // block-level steps
      for (int t = threadIdx.x; t < half; t += nthreads) {
        const int pos = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
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
    if (stride > 0) {                                           // the remaining strides stay inside a warp's chunk
      const int t0 = warp * (chunk >> 1);
      for (; stride > 0; stride >>= 1) {
        for (int tt = lane; tt < (chunk >> 1); tt += 32) {
          const int t = t0 + tt;
          const int pos = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
          const int partner = pos + stride;
          const bool desc = (pos & size) == 0;
          const uint64_t a = sel[pos], b = sel[partner];
          if ((a < b) == desc) {
            sel[pos] = b;
            sel[partner] = a;
          }
        }
        __syncwarp();
      }
      __syncthreads();
    }
  }
}

// ---- register-resident variant for long sorts ------------------------------------------------------------
// A warp holds a chunk of 32 P consecutive elements, P per lane (element j * 32 + lane of the chunk): strides
// below 32 exchange through shuffles, strides 32 .. 16 P are register moves; only strides >= 32 P go through
// shared memory with a block barrier (ksort = 4096, 16 warps, P = 8: 10 block-level steps out of 78).
template <int P>
__device__ __forceinline__ void warp_bitonic_step(uint64_t (&k)[P], int lane, int stride, bool desc_uniform, int size,
                                                  bool use_uniform) {
  const unsigned full = 0xffffffffu;
  if (stride >= 32) {
    const int js = stride >> 5;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      if ((j & js) == 0) {
        const bool desc = use_uniform ? desc_uniform : (((j * 32) & size) == 0);
        const uint64_t a = k[j], b = k[j | js];
        const bool swap = (a < b) == desc;
        k[j] = swap ? b : a;
        k[j | js] = swap ? a : b;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const bool desc = use_uniform ? desc_uniform : ((((j * 32) | lane) & size) == 0);
      const bool lower = (lane & stride) == 0;
      const uint64_t mine = k[j];
      const uint64_t other = __shfl_xor_sync(full, mine, stride);
      const bool mine_big = mine > other;
      const bool take_max = (lower == desc);
      k[j] = (take_max == mine_big) ? mine : other;
    }
  }
}

template <int P>
__device__ __forceinline__ void block_bitonic_desc_regs(uint64_t* sel, int ksort) {
  constexpr int CH = 32 * P;
  const int nthreads = blockDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = nthreads >> 5;
  const int nchunks = ksort / CH;
  const int half = ksort >> 1;
  // ---- phase 1: every chunk fully sorted in registers, direction from its position (sizes 2 .. CH)
  for (int c = warp; c < nchunks; c += nwarps) {
    uint64_t k[P];
    const int base = c * CH;
#pragma unroll
    for (int j = 0; j < P; ++j) k[j] = sel[base + j * 32 + lane];
#pragma unroll
    for (int size = 2; size <= CH; size <<= 1) {
      const bool top = size == CH;
      const bool desc_u = (base & size) == 0;
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) warp_bitonic_step<P>(k, lane, stride, desc_u, size, top);
    }
#pragma unroll
    for (int j = 0; j < P; ++j) sel[base + j * 32 + lane] = k[j];
  }
  __syncthreads();
  // ---- phase 2: sizes 2 CH .. ksort: long strides in shared memory, the tail of every merge in registers
  for (int size = 2 * CH; size <= ksort; size <<= 1) {
    for (int stride = size >> 1; stride >= CH; stride >>= 1) {
      for (int t = threadIdx.x; t < half; t += nthreads) {
        const int pos = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));
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
    for (int c = warp; c < nchunks; c += nwarps) {
      uint64_t k[P];
      const int base = c * CH;
      const bool desc_u = (base & size) == 0;
#pragma unroll
      for (int j = 0; j < P; ++j) k[j] = sel[base + j * 32 + lane];
#pragma unroll
      for (int stride = CH >> 1; stride > 0; stride >>= 1) warp_bitonic_step<P>(k, lane, stride, desc_u, 0, true);
#pragma unroll
      for (int j = 0; j < P; ++j) sel[base + j * 32 + lane] = k[j];
    }
    __syncthreads();
  }
}

// Sort the `out` keys of sel[0, ksort) (ksort = next power of two >= out, the tail zero padded) descending and
// return the buffer that holds the result. A bitonic network is oblivious to padding: k = 2097 would pay for
// 4096 keys. When at least a quarter of the network would sort zeros and `scratch` is large enough, the keys are
// split by a second radix select into the top ksort / 2 and the remainder (padded to ITS next power of two), the
// two parts are sorted separately in scratch and already stand in order (2097 -> 2048 + 64: 2.4x less work).
// hist: 256 ints, ctl: 4 ints, pos: 2 ints of shared memory. Ends with a block barrier.
__device__ __forceinline__ const uint64_t* block_sort_selected(uint64_t* sel, int out, int ksort, uint64_t* scratch,
                                                               int scratch_cap, int* hist, int* ctl, int* pos) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int p2 = ksort >> 1;
  const int rest = out - p2;
  if (ksort >= 2048 && rest > 0 && rest <= (ksort >> 2)) {
    int r2 = 2;
    while (r2 < rest) r2 <<= 1;
    if (p2 + r2 <= scratch_cap) {
      const uint64_t T1 = block_radix_select([&](int e) { return sel[e]; }, out, p2, hist, ctl);
      if (threadIdx.x == 0) { pos[0] = 0; pos[1] = 0; }
      __syncthreads();
      for (int base = 0; base < out; base += blockDim.x) {
        const int e = base + threadIdx.x;
        const uint64_t key = (e < out) ? sel[e] : 0ull;
        const bool hi = (e < out) && (key >= T1);
        const bool lo = (e < out) && !hi;
        const unsigned bh = __ballot_sync(full, hi), bl = __ballot_sync(full, lo);
        int ph = 0, pl = 0;
        if (lane == 0) {
          if (bh) ph = atomicAdd(&pos[0], __popc(bh));
          if (bl) pl = atomicAdd(&pos[1], __popc(bl));
        }
        ph = __shfl_sync(full, ph, 0) + __popc(bh & lt_mask);
        pl = __shfl_sync(full, pl, 0) + __popc(bl & lt_mask);
        if (hi && ph < p2) scratch[ph] = key;
        if (lo && pl < r2) scratch[p2 + pl] = key;
      }
      for (int e = rest + threadIdx.x; e < r2; e += blockDim.x) scratch[p2 + e] = 0ull;
      __syncthreads();
      block_bitonic_desc_regs<8>(scratch, p2);
      if (r2 >= 256) block_bitonic_desc_regs<8>(scratch + p2, r2);
      else block_bitonic_desc(scratch + p2, r2);
      return scratch;
    }
  }
  if (ksort >= 2048) block_bitonic_desc_regs<8>(sel, ksort);
  else block_bitonic_desc(sel, ksort);
  return sel;
}

}  // namespace qsae

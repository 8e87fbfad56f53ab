// Warp-level int4 row decoder shared by the stand-alone decode kernel (decode.cu) and the merge kernels that
// decode a row right after selecting it (select_topk.cu): recon[d] = scale * sum_j v_j * dict[i_j, d] + bias[d]
// with dict = two's-complement nibbles (sae/binary.py:49-58 quantized_int_weights, :38 the matmul it replaces).
//
// Exact integer accumulation. The decoder is bound by instruction issue, not by memory (the packed dictionary
// is L2 resident at H = 32768), and int -> float conversion of every nibble was half of the issue slots. Instead
// the row's k values are converted ONCE to fixed point, v_j = round(v_j * 2^S) with S chosen from max_j |v_j| so
// that sum_j 15 |v_j| 2^S stays inside the accumulator, and every dictionary nibble contributes one integer
// multiply-add: acc[d] += u'_jd * vfix_j with u' = w + 8 in [0, 15] (w ^ 8 on the two's complement nibble). The
// bias of 8 leaves with one correction per row, acc[d] - 8 sum_j vfix_j, and a single int -> float conversion
// per output feature follows. The sum is exact in integers (order independent, deterministic); the only
// rounding is that of v_j to fixed point: 32-bit accumulators for k <= 128 (>= 18 fractional bits of max|v|),
// 64-bit accumulators above (WIDE: >= 30 fractional bits for every k <= 4096).
// WPL = 32-bit words (8 features each) per lane, read as one vector load: lane l owns the features
// [8 WPL l, 8 WPL (l + 1)) -- D <= 256 WPL. FULL: D == 256 WPL exactly, no per-word guards.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace qsae {

// 8-byte asynchronous copy global -> shared (LDGSTS): the gather of a whole batch of dictionary rows is issued
// before any of it is consumed
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int WPL, bool FULL, bool WIDE = false>
struct Int4RowDecoder {
  using acc_t = typename std::conditional<WIDE, long long, int>::type;
  acc_t acc[WPL][8];
  acc_t vsum;
  float to_fixed, from_fixed;
  bool bad;

  // amax: max |v| over the row's entries (already reduced over the warp); any_bad: a NaN / infinity among them
  __device__ __forceinline__ void begin(float amax, bool any_bad, int k) {
    int klog = 0;
    while ((1 << klog) < k) ++klog;
    const int e2 = max(static_cast<int>((__float_as_uint(amax) >> 23) & 0xFF) - 127, -100);  // amax < 2^(e2 + 1)
    // 32-bit: |v| 2^S < 2^(26 - klog), so 15 k |v| 2^S < 2^30. 64-bit: |v| 2^S < 2^30 (one int32 per value),
    // 15 * 4096 * 2^30 < 2^46.
    const int S = WIDE ? min(30 - 1 - e2, 120) : min(26 - klog - 1 - e2, 120);
    to_fixed = __uint_as_float(static_cast<uint32_t>(S + 127) << 23);
    from_fixed = __uint_as_float(static_cast<uint32_t>(127 - S) << 23);
    bad = any_bad;
#pragma unroll
    for (int c = 0; c < WPL; ++c)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c][j] = 0;
    vsum = 0;
  }

  // One chunk of up to 32 entries: lane l holds entry l (my_v, my_i; my_i < 0 = not in this dictionary).
  // m = entries in the chunk; SKIP: test a warp-uniform ownership mask (dictionary shards own ~1 / G of the
  // winners; the test costs the plain decoder 10 % of its issue slots, hence the template flag).
  template <bool SKIP>
  __device__ __forceinline__ void add_chunk(float my_v, int my_i, int m, const uint32_t* __restrict__ packed,
                                            int words_per_row, int lane) {
    const unsigned full = 0xffffffffu;
    const bool mine = my_i >= 0;
    const int my_f = mine ? __float2int_rn(my_v * to_fixed) : 0;   // entries outside contribute nothing: value 0, row 0
    my_i = mine ? my_i : 0;
    const unsigned owned = SKIP ? __ballot_sync(full, mine) : 0xffffffffu;
    const int w0 = lane * WPL;
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      if (SKIP && ((owned >> j) & 1u) == 0u) continue;
      const int vf = __shfl_sync(full, my_f, j);
      const int i = __shfl_sync(full, my_i, j);
      vsum += vf;
      const uint32_t* drow = packed + static_cast<size_t>(i) * words_per_row + w0;
      uint32_t word[WPL];
      if constexpr (FULL) {
        if constexpr (WPL == 1) {
          word[0] = __ldg(drow);
        } else if constexpr (WPL == 2) {
          const uint2 t = __ldg(reinterpret_cast<const uint2*>(drow));
          word[0] = t.x; word[1] = t.y;
        } else {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(drow));
          word[0] = t.x; word[1] = t.y; word[2] = t.z; word[3] = t.w;
        }
      } else {
#pragma unroll
        for (int c = 0; c < WPL; ++c) word[c] = (w0 + c < words_per_row) ? __ldg(drow + c) : 0x88888888u;
      }
#pragma unroll
      for (int c = 0; c < WPL; ++c) {
        const uint32_t bits = word[c] ^ 0x88888888u;               // biased nibbles u' = w + 8
        const uint32_t lo = bits & 0x0F0F0F0Fu;                     // features 0, 2, 4, 6 as bytes
        const uint32_t hi = (bits >> 4) & 0x0F0F0F0Fu;              // features 1, 3, 5, 7
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (WIDE) {
            acc[c][2 * q] += static_cast<long long>(static_cast<int>(__byte_perm(lo, 0u, 0x4440u + q))) * vf;
            acc[c][2 * q + 1] += static_cast<long long>(static_cast<int>(__byte_perm(hi, 0u, 0x4440u + q))) * vf;
          } else {
            acc[c][2 * q] += static_cast<int>(__byte_perm(lo, 0u, 0x4440u + q)) * vf;
            acc[c][2 * q + 1] += static_cast<int>(__byte_perm(hi, 0u, 0x4440u + q)) * vf;
          }
        }
      }
    }
  }

  // The same accumulation from dictionary rows that were staged in shared memory first (all of a batch's gathers
  // in flight at once instead of four at a time): row e of the batch at rows + e * 32 * WPL words; lane `first + e`
  // holds the batch's entry e (my_v; mine == false: the entry contributes nothing and its staged row is ignored).
  __device__ __forceinline__ void add_staged(float my_v, bool mine, int first, int m, const uint32_t* rows, int lane) {
    const unsigned full = 0xffffffffu;
    const int my_f = mine ? __float2int_rn(my_v * to_fixed) : 0;
#pragma unroll 4
    for (int e = 0; e < m; ++e) {
      const int vf = __shfl_sync(full, my_f, first + e);
      vsum += vf;
      uint32_t word[WPL];
      const uint32_t* r = rows + (e * 32 + lane) * WPL;
      if constexpr (WPL == 1) {
        word[0] = r[0];
      } else if constexpr (WPL == 2) {
        const uint2 t = *reinterpret_cast<const uint2*>(r);
        word[0] = t.x; word[1] = t.y;
      } else {
        const uint4 t = *reinterpret_cast<const uint4*>(r);
        word[0] = t.x; word[1] = t.y; word[2] = t.z; word[3] = t.w;
      }
#pragma unroll
      for (int c = 0; c < WPL; ++c) {
        const uint32_t bits = word[c] ^ 0x88888888u;
        const uint32_t lo = bits & 0x0F0F0F0Fu;
        const uint32_t hi = (bits >> 4) & 0x0F0F0F0Fu;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if constexpr (WIDE) {
            acc[c][2 * q] += static_cast<long long>(static_cast<int>(__byte_perm(lo, 0u, 0x4440u + q))) * vf;
            acc[c][2 * q + 1] += static_cast<long long>(static_cast<int>(__byte_perm(hi, 0u, 0x4440u + q))) * vf;
          } else {
            acc[c][2 * q] += static_cast<int>(__byte_perm(lo, 0u, 0x4440u + q)) * vf;
            acc[c][2 * q + 1] += static_cast<int>(__byte_perm(hi, 0u, 0x4440u + q)) * vf;
          }
        }
      }
    }
  }

  __device__ __forceinline__ void finish(float scale, const float* __restrict__ bias, float* __restrict__ recon_row,
                                         int D, int lane) const {
    const acc_t corr = 8 * vsum;
    const float qnan = __uint_as_float(0x7FC00000u);
    const int w0 = lane * WPL;
#pragma unroll
    for (int c = 0; c < WPL; ++c) {
      const int d = (w0 + c) * 8;
      if (FULL || d < D) {
        float o[8], bv[8];
        if (bias != nullptr) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + d));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + d + 4));
          bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
          bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) bv[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float sum = static_cast<float>(acc[c][q] - corr) * from_fixed;
          o[q] = bad ? qnan : (scale * sum + bv[q]);
        }
        float4* dst = reinterpret_cast<float4*>(recon_row + d);
        dst[0] = make_float4(o[0], o[1], o[2], o[3]);
        dst[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
};

}  // namespace qsae

// Sparse top-k decoders: recon[b, :] = scale * sum_j vals[b, j] * dict[idx[b, j], :] + bias.
//
// Replaces the dense latent.matmul(int_weights) of sae/binary.py:38 (and nn.Linear decode of
// sae/baseline.py:29): only the k selected dictionary rows are touched. One warp per token row;
// the 32 lanes read one dictionary row as a single contiguous, vectorised request (int4: 4 B per
// lane = 128 B per 256 features; fp32: 16 B per lane = 512 B per 128 features), so every gather
// is fully coalesced. The packed dictionaries are L2 resident at H = 32768; see DESIGN.md for the
// measured bound of each variant.
// Also: densify (sparse -> dense [B, H], sae/binary.py:96-99).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qsae {

namespace {

constexpr int kDecWarps = 4;

// NCH = number of 256-feature chunks a lane accumulates (D <= 256 * NCH)
//
// Exact integer accumulation. The kernel is bound by instruction issue, not by memory (the packed
// dictionary is L2 resident), and int -> float conversion of every nibble was half of the issue slots.
// Instead the row's k values are converted ONCE to fixed point, v_j = round(v_j * 2^S) with S chosen
// from max_j |v_j| so that sum_j 15 |v_j| 2^S < 2^30, and every dictionary nibble contributes one
// integer multiply-add: acc[d] += u'_jd * vfix_j with u' = w + 8 in [0, 15] (w ^ 8 on the two's
// complement nibble). The bias of 8 leaves with one correction per row, acc[d] - 8 sum_j vfix_j, and a
// single int -> float conversion per output feature follows. The sum is exact in integers (order
// independent, deterministic); the only rounding is that of v_j to >= 18 fractional bits of max|v|.
// WPL = 32-bit words (8 features each) per lane, read as one vector load: lane l owns the features
// [8 WPL l, 8 WPL (l + 1)) -- D <= 256 WPL. FULL: D == 256 WPL exactly, no per-word guards.
template <int WPL, bool FULL, bool RANGE>
__global__ void __launch_bounds__(kDecWarps * 32)
decode_int4_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int B, int k,
                   const uint32_t* __restrict__ packed, int H, int D, float scale,
                   const float* __restrict__ bias, float* __restrict__ recon, int idx_offset) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kDecWarps + warp;
  if (row >= B) return;
  const unsigned full = 0xffffffffu;
  const int words_per_row = D >> 3;
  const int w0 = lane * WPL;                      // first word of this lane
  const float* vrow = vals + static_cast<size_t>(row) * k;
  const int32_t* irow = idx + static_cast<size_t>(row) * k;

  // ---- scale of the row: max |v| over the entries that belong to this dictionary
  float amax = 0.f;
  bool bad = false;
  for (int e = lane; e < k; e += 32) {
    const int i = irow[e] - idx_offset;   // dictionary shards: only [idx_offset, idx_offset + H) is ours
    if (i >= 0 && i < H) {
      const float v = vrow[e];
      bad |= !(fabsf(v) <= 3.0e38f);      // NaN or infinity
      amax = fmaxf(amax, fabsf(v));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(full, amax, o));
  bad = __any_sync(full, bad);
  int klog = 0;
  while ((1 << klog) < k) ++klog;
  const int e2 = max(static_cast<int>((__float_as_uint(amax) >> 23) & 0xFF) - 127, -100);  // amax < 2^(e2 + 1)
  const int S = min(26 - klog - 1 - e2, 120);                                              // |v| 2^S < 2^(26 - klog)
  const float to_fixed = __uint_as_float(static_cast<uint32_t>(S + 127) << 23);
  const float from_fixed = __uint_as_float(static_cast<uint32_t>(127 - S) << 23);

  int acc[WPL][8];
#pragma unroll
  for (int c = 0; c < WPL; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0;
  int vsum = 0;

  for (int base = 0; base < k; base += 32) {
    const int e = base + lane;
    int my_i = (e < k) ? irow[e] - idx_offset : -1;
    const bool mine = my_i >= 0 && my_i < H;
    // entries outside this dictionary contribute nothing: value 0, row 0
    const int my_f = mine ? __float2int_rn(vrow[e] * to_fixed) : 0;
    my_i = mine ? my_i : 0;
    const int m = min(32, k - base);
    // dictionary shards own ~1 / G of the winners: skip the rest (warp-uniform test; RANGE is a template
    // flag because the test costs the plain decoder 10 % of its issue slots)
    const unsigned owned = RANGE ? __ballot_sync(full, mine) : 0xffffffffu;
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      if (RANGE && ((owned >> j) & 1u) == 0u) continue;
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
          acc[c][2 * q] += static_cast<int>(__byte_perm(lo, 0u, 0x4440u + q)) * vf;
          acc[c][2 * q + 1] += static_cast<int>(__byte_perm(hi, 0u, 0x4440u + q)) * vf;
        }
      }
    }
  }
  const int corr = 8 * vsum;
  const float qnan = __uint_as_float(0x7FC00000u);
#pragma unroll
  for (int c = 0; c < WPL; ++c) {
    const int d = (w0 + c) * 8;
    if (FULL || d < D) {
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float sum = static_cast<float>(acc[c][q] - corr) * from_fixed;
        o[q] = bad ? qnan : (scale * sum + (bias ? __ldg(bias + d + q) : 0.f));
      }
      float4* dst = reinterpret_cast<float4*>(recon + static_cast<size_t>(row) * D + d);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// generic row decoder over float4-sized groups: T = float (4 per 16 B) or int8 (4 per 4 B)
template <typename T, int NCH>
__global__ void __launch_bounds__(kDecWarps * 32)
decode_rows_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int B, int k,
                   const T* __restrict__ rows, int H, int D, float scale,
                   const float* __restrict__ bias, float* __restrict__ recon, int idx_offset) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kDecWarps + warp;
  if (row >= B) return;
  const unsigned full = 0xffffffffu;
  float4 acc[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int base = 0; base < k; base += 32) {
    const int e = base + lane;
    const float my_v = (e < k) ? vals[static_cast<size_t>(row) * k + e] : 0.f;
    // dictionary shards: only entries inside [idx_offset, idx_offset + H) belong to this dictionary
    const int my_i = (e < k) ? idx[static_cast<size_t>(row) * k + e] - idx_offset : -1;
    const int m = min(32, k - base);
#pragma unroll 4
    for (int j = 0; j < m; ++j) {
      const float v = __shfl_sync(full, my_v, j);
      const int i = __shfl_sync(full, my_i, j);
      if (i < 0 || i >= H) continue;
      const T* drow = rows + static_cast<size_t>(i) * D;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int d = (c * 32 + lane) * 4;
        if (d < D) {
          float4 w;
          if constexpr (sizeof(T) == 4) {
            w = __ldg(reinterpret_cast<const float4*>(drow + d));
          } else {
            const char4 q = __ldg(reinterpret_cast<const char4*>(drow + d));
            w = make_float4(q.x, q.y, q.z, q.w);
          }
          acc[c].x = fmaf(v, w.x, acc[c].x);
          acc[c].y = fmaf(v, w.y, acc[c].y);
          acc[c].z = fmaf(v, w.z, acc[c].z);
          acc[c].w = fmaf(v, w.w, acc[c].w);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int d = (c * 32 + lane) * 4;
    if (d < D) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + d));
      *reinterpret_cast<float4*>(recon + static_cast<size_t>(row) * D + d) =
          make_float4(scale * acc[c].x + b.x, scale * acc[c].y + b.y, scale * acc[c].z + b.z,
                      scale * acc[c].w + b.w);
    }
  }
}

__global__ void scatter_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx,
                               size_t total, int k, int H, float* __restrict__ dense) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int i = idx[e];
    if (i >= 0 && i < H) dense[(e / k) * static_cast<size_t>(H) + i] = vals[e];
  }
}

__global__ void pack_candidates_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, size_t n,
                                       uint2* __restrict__ out) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < n; e += stride)
    out[e] = make_uint2(__float_as_uint(vals[e]), static_cast<uint32_t>(idx[e]));
}

}  // namespace

const char* pack_candidates_launch(const float* vals, const int32_t* idx, size_t n, void* out, cudaStream_t stream) {
  if (n == 0) return nullptr;
  size_t g = (n + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  pack_candidates_kernel<<<static_cast<int>(g), 256, 0, stream>>>(vals, idx, n, reinterpret_cast<uint2*>(out));
  return cuda_err(cudaGetLastError());
}

const char* decode_int4_launch(const float* vals, const int32_t* idx, int B, int k,
                               const uint8_t* packed, int H, int D, float scale, const float* bias,
                               float* recon, int idx_offset, cudaStream_t stream, bool skip_unowned) {
  const int blocks = (B + kDecWarps - 1) / kDecWarps;
  const uint32_t* p32 = reinterpret_cast<const uint32_t*>(packed);
  const bool range = skip_unowned || idx_offset != 0;   // dictionary shards skip the winners other shards own
#define QSAE_DEC4(WPL, FULL)                                                                                          \
  do {                                                                                                                \
    if (range)                                                                                                        \
      decode_int4_kernel<WPL, FULL, true><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale,   \
                                                                                 bias, recon, idx_offset);            \
    else                                                                                                              \
      decode_int4_kernel<WPL, FULL, false><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale,  \
                                                                                  bias, recon, idx_offset);           \
  } while (0)
  const bool aligned = (reinterpret_cast<uintptr_t>(packed) & 15) == 0;
  if (D == 256) QSAE_DEC4(1, true);
  else if (D == 512 && aligned) QSAE_DEC4(2, true);
  else if (D == 1024 && aligned) QSAE_DEC4(4, true);
  else if (D <= 256) QSAE_DEC4(1, false);
  else if (D <= 512) QSAE_DEC4(2, false);
  else if (D <= 1024) QSAE_DEC4(4, false);
  else return "decode_int4: D must be <= 1024";
#undef QSAE_DEC4
  return cuda_err(cudaGetLastError());
}

template <typename T>
static const char* decode_rows_dispatch(const float* vals, const int32_t* idx, int B, int k,
                                        const T* rows, int H, int D, float scale, const float* bias,
                                        float* recon, int idx_offset, cudaStream_t stream) {
  const int blocks = (B + kDecWarps - 1) / kDecWarps;
  if (D <= 128)
    decode_rows_kernel<T, 1><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset);
  else if (D <= 256)
    decode_rows_kernel<T, 2><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset);
  else if (D <= 512)
    decode_rows_kernel<T, 4><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset);
  else if (D <= 1024)
    decode_rows_kernel<T, 8><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset);
  else
    return "decode_rows: D must be <= 1024";
  return cuda_err(cudaGetLastError());
}

const char* decode_int8_launch(const float* vals, const int32_t* idx, int B, int k,
                               const int8_t* rows, int H, int D, float scale, const float* bias,
                               float* recon, int idx_offset, cudaStream_t stream) {
  return decode_rows_dispatch<int8_t>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset, stream);
}

const char* decode_f32_launch(const float* vals, const int32_t* idx, int B, int k, const float* rows,
                              int H, int D, float scale, const float* bias, float* recon, int idx_offset,
                              cudaStream_t stream) {
  return decode_rows_dispatch<float>(vals, idx, B, k, rows, H, D, scale, bias, recon, idx_offset, stream);
}

const char* densify_launch(const float* vals, const int32_t* idx, int B, int k, int H, float* dense,
                           cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(dense, 0, static_cast<size_t>(B) * H * sizeof(float), stream);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  const size_t total = static_cast<size_t>(B) * k;
  if (total == 0) return nullptr;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  scatter_kernel<<<static_cast<int>(g), 256, 0, stream>>>(vals, idx, total, k, H, dense);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

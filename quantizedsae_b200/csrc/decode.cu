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

#include "decode_common.cuh"
#include "kernels.h"

namespace qsae {

namespace {

constexpr int kDecWarps = 4;

// decode_int4_kernel: one warp per token row around Int4RowDecoder (decode_common.cuh). WIDE (64-bit accumulators)
// is chosen by the launcher for k > 128, where the 32-bit fixed-point scale would leave fewer than 18 fractional
// bits of max |v|.
template <int WPL, bool FULL, bool RANGE, bool WIDE>
__global__ void __launch_bounds__(kDecWarps * 32)
decode_int4_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int B, int k,
                   const uint32_t* __restrict__ packed, int H, int D, float scale,
                   const float* __restrict__ bias, float* __restrict__ recon, int idx_offset) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kDecWarps + warp;
  if (row >= B) return;
  const unsigned full = 0xffffffffu;
  const int words_per_row = D >> 3;
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

  Int4RowDecoder<WPL, FULL, WIDE> dec;
  dec.begin(amax, bad, k);
  for (int base = 0; base < k; base += 32) {
    const int e = base + lane;
    int my_i = (e < k) ? irow[e] - idx_offset : -1;
    if (my_i >= H) my_i = -1;
    const float my_v = (my_i >= 0) ? vrow[e] : 0.f;
    dec.template add_chunk<RANGE>(my_v, my_i, min(32, k - base), packed, words_per_row, lane);
  }
  dec.finish(scale, bias, recon + static_cast<size_t>(row) * D, D, lane);
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
#define QSAE_DEC4K(WPL, FULL, RANGE, WIDE)                                                                             \
  decode_int4_kernel<WPL, FULL, RANGE, WIDE><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale, \
                                                                                    bias, recon, idx_offset)
#define QSAE_DEC4(WPL, FULL)                                  \
  do {                                                        \
    if (range && wide) QSAE_DEC4K(WPL, FULL, true, true);     \
    else if (range) QSAE_DEC4K(WPL, FULL, true, false);       \
    else if (wide) QSAE_DEC4K(WPL, FULL, false, true);        \
    else QSAE_DEC4K(WPL, FULL, false, false);                 \
  } while (0)
  const bool wide = k > 128;   // 64-bit accumulators keep >= 30 fractional bits of max |v| for any k
  const bool aligned = (reinterpret_cast<uintptr_t>(packed) & 15) == 0;
  if (D == 256) QSAE_DEC4(1, true);
  else if (D == 512 && aligned) QSAE_DEC4(2, true);
  else if (D == 1024 && aligned) QSAE_DEC4(4, true);
  else if (D <= 256) QSAE_DEC4(1, false);
  else if (D <= 512) QSAE_DEC4(2, false);
  else if (D <= 1024) QSAE_DEC4(4, false);
  else return "decode_int4: D must be <= 1024";
#undef QSAE_DEC4K
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

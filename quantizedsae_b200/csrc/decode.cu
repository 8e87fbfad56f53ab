// Sparse top-k decoders: recon[b, :] = scale * sum_j vals[b, j] * dict[idx[b, j], :] + bias.
//
// Replaces the dense latent.matmul(int_weights) of sae/binary.py:38 (and nn.Linear decode of
// sae/baseline.py:29): only the k selected dictionary rows are touched. One warp per token row;
// the 32 lanes read one dictionary row as a single contiguous, vectorised request (int4: 4 B per
// lane = 128 B per 256 features; fp32: 16 B per lane = 512 B per 128 features), so every gather
// is fully coalesced. Memory-bound by construction: the figure of merit is achieved GB/s.
// Also: densify (sparse -> dense [B, H], sae/binary.py:96-99).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qsae {

namespace {

constexpr int kDecWarps = 4;

// sign-extended nibble j (0..7) of a 32-bit word
__device__ __forceinline__ float nibble_f(uint32_t w, int j) {
  return static_cast<float>(static_cast<int32_t>(w << (28 - 4 * j)) >> 28);
}

// NCH = number of 256-feature chunks a lane accumulates (D <= 256 * NCH)
template <int NCH>
__global__ void __launch_bounds__(kDecWarps * 32)
decode_int4_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int B, int k,
                   const uint32_t* __restrict__ packed, int H, int D, float scale,
                   const float* __restrict__ bias, float* __restrict__ recon, int idx_offset) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * kDecWarps + warp;
  if (row >= B) return;
  const unsigned full = 0xffffffffu;
  const int words_per_row = D >> 3;
  float acc[NCH][8];
#pragma unroll
  for (int c = 0; c < NCH; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;

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
      if (i < 0 || i >= H) continue;  // warp-uniform
      const uint32_t* drow = packed + static_cast<size_t>(i) * words_per_row;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int w = c * 32 + lane;
        if (w < words_per_row) {
          const uint32_t bits = __ldg(drow + w);
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[c][q] = fmaf(v, nibble_f(bits, q), acc[c][q]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int d = (c * 32 + lane) * 8;
    if (d < D) {
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = scale * acc[c][q] + (bias ? __ldg(bias + d + q) : 0.f);
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
                               float* recon, int idx_offset, cudaStream_t stream) {
  const int blocks = (B + kDecWarps - 1) / kDecWarps;
  const uint32_t* p32 = reinterpret_cast<const uint32_t*>(packed);
  if (D <= 256)
    decode_int4_kernel<1><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale, bias, recon, idx_offset);
  else if (D <= 512)
    decode_int4_kernel<2><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale, bias, recon, idx_offset);
  else if (D <= 1024)
    decode_int4_kernel<4><<<blocks, kDecWarps * 32, 0, stream>>>(vals, idx, B, k, p32, H, D, scale, bias, recon, idx_offset);
  else
    return "decode_int4: D must be <= 1024";
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

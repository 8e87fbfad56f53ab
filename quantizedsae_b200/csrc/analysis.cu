// Sparse-native analysis consumers (scripts/analysis/dynamic_analysis.py:317-440 of the reference).
// The reference densifies the activity of every batch into a [B, H] boolean mask and forms the
// co-activation counts as a dense [H, B] x [B, H] matrix product (141 TFLOP per 65536-token batch at
// H = 32768) that it ships to the CPU chunk by chunk. Here the active sets stay sparse -- idx [B, cap] with
// ~k entries per row -- and the [H, H] int32 matrix stays resident in HBM (4.3 GB at H = 32768):
// k^2 atomic increments per token instead of H^2 multiply-adds per token.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qsae {

namespace {

__device__ __forceinline__ bool entry_active(const int32_t* idx, const float* vals, size_t e, int H) {
  const int i = idx[e];
  return i >= 0 && i < H && (vals == nullptr || vals[e] > 0.f);
}

// activation_counts[h] += number of rows in which h is active (dynamic_analysis.py:341)
__global__ void __launch_bounds__(256)
activation_counts_kernel(const int32_t* __restrict__ idx, const float* __restrict__ vals, size_t total, int H,
                         unsigned long long* __restrict__ counts) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride)
    if (entry_active(idx, vals, e, H)) atomicAdd(&counts[idx[e]], 1ull);
}

// coactivation[i, j] += 1 for every ordered pair (i, j) of latents active in the same row, i == j included
// (A^T A of the boolean activity matrix, dynamic_analysis.py:344-345). One warp per row: the row's active
// latents are compacted into shared memory, then lane j of the warp increments C[a_i, a_j] for each i.
constexpr int kCoWarps = 4;
constexpr int kCoCap = 2048;   // active latents per row held in shared memory

__global__ void __launch_bounds__(kCoWarps * 32)
coactivation_kernel(const int32_t* __restrict__ idx, const float* __restrict__ vals, int B, int cap, int H,
                    int32_t* __restrict__ cooc) {
  __shared__ int act[kCoWarps][kCoCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int row = blockIdx.x * kCoWarps + warp; row < B; row += gridDim.x * kCoWarps) {
    int n = 0;
    for (int base = 0; base < cap; base += 32) {
      const int e = base + lane;
      const bool a = e < cap && entry_active(idx, vals, static_cast<size_t>(row) * cap + e, H);
      const unsigned b = __ballot_sync(full, a);
      const int pos = n + __popc(b & lt_mask);
      if (a && pos < kCoCap) act[warp][pos] = idx[static_cast<size_t>(row) * cap + e];
      n += __popc(b);
    }
    n = min(n, kCoCap);   // the launcher rejects cap > kCoCap
    __syncwarp();
    for (int i = 0; i < n; ++i) {
      int32_t* crow = cooc + static_cast<size_t>(act[warp][i]) * H;
      for (int j = lane; j < n; j += 32) atomicAdd(&crow[act[warp][j]], 1);
    }
    __syncwarp();
  }
}

// *out += sum (a - b)^2 in fp64 (dynamic_analysis.py:96-100: mean squared reconstruction error numerator)
__global__ void __launch_bounds__(256)
sq_error_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, double* __restrict__ out) {
  __shared__ double s[256];
  double acc = 0.0;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
  const size_t n4 = vec ? n / 4 : 0;   // 16-byte loads, two in flight per array
  const float4* a4 = reinterpret_cast<const float4*>(a);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  for (size_t e = tid; e < n4; e += 2 * stride) {
    const float4 u0 = a4[e], v0 = b4[e];
    const bool two = e + stride < n4;
    const float4 u1 = two ? a4[e + stride] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 v1 = two ? b4[e + stride] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float d0 = u0.x - v0.x, d1 = u0.y - v0.y, d2 = u0.z - v0.z, d3 = u0.w - v0.w;   // exact differences would need fp64;
    const float d4 = u1.x - v1.x, d5 = u1.y - v1.y, d6 = u1.z - v1.z, d7 = u1.w - v1.w;   // the reference subtracts in fp32 too
    acc += static_cast<double>(d0) * d0 + static_cast<double>(d1) * d1 + static_cast<double>(d2) * d2 +
           static_cast<double>(d3) * d3 + static_cast<double>(d4) * d4 + static_cast<double>(d5) * d5 +
           static_cast<double>(d6) * d6 + static_cast<double>(d7) * d7;
  }
  for (size_t e = n4 * 4 + tid; e < n; e += stride) {
    const float d = a[e] - b[e];
    acc += static_cast<double>(d) * d;
  }
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(out, s[0]);
}


// Dense [B, H] latents -> per-row lists of the entries with (mode 0) v != 0 or (mode 1) v > thr, in ascending latent
// order: idx [B, cap] (-1 padded) and / or vals [B, cap] (0 padded) and / or pairs [B, cap, 2] (second component = latent,
// the format of qsae_decode_matryoshka_lists); cnt [B] counts every hit, stored or not. The entry the reference-shaped
// decoders take when a caller hands them the dense matrix (sae/binary.py:24, sae/quantized_matryoshka.py:47,99).
__global__ void __launch_bounds__(256)
compact_dense_kernel(const float* __restrict__ dense, int B, int H, int mode, float thr, int cap, int32_t* __restrict__ idx,
                     float* __restrict__ vals, int32_t* __restrict__ pairs, int32_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const bool vec = (H & 3) == 0 && (reinterpret_cast<uintptr_t>(dense) & 15) == 0;
  for (int row = warp; row < B; row += nwarps) {
    const float* r = dense + static_cast<size_t>(row) * H;
    int n = 0;
    const int step = vec ? 128 : 32;
    for (int base = 0; base < H; base += step) {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      int first, nv;
      if (vec) {
        first = base + lane * 4;
        nv = first < H ? 4 : 0;
        if (nv) {
          const float4 q = *reinterpret_cast<const float4*>(r + first);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
      } else {
        first = base + lane;
        nv = first < H ? 1 : 0;
        if (nv) v[0] = r[first];
      }
      bool hit[4];
      int mine = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        hit[c] = c < nv && (mode == 0 ? v[c] != 0.f : v[c] > thr);
        mine += hit[c] ? 1 : 0;
      }
      // exclusive prefix of the per-lane hit counts (ascending latent order = lane-major, then component)
      int incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int total = __shfl_sync(0xffffffffu, incl, 31);
      int pos = n + incl - mine;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (hit[c]) {
          if (pos < cap) {
            const size_t o = static_cast<size_t>(row) * cap + pos;
            if (idx) idx[o] = first + c;
            if (vals) vals[o] = v[c];
            if (pairs) { pairs[2 * o] = 0; pairs[2 * o + 1] = first + c; }
          }
          ++pos;
        }
      }
      n += total;
    }
    (void)lt;
    for (int e = min(n, cap) + lane; e < cap; e += 32) {       // padding
      const size_t o = static_cast<size_t>(row) * cap + e;
      if (idx) idx[o] = -1;
      if (vals) vals[o] = 0.f;
      if (pairs) { pairs[2 * o] = 0; pairs[2 * o + 1] = 0; }
    }
    if (lane == 0) cnt[row] = n;
  }
}

}  // namespace

const char* compact_dense_launch(const float* dense, int B, int H, int mode, float thr, int cap, int32_t* idx, float* vals,
                                 int32_t* pairs, int32_t* cnt, cudaStream_t stream) {
  if (B == 0) return nullptr;
  int blocks = (B + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  compact_dense_kernel<<<blocks, 256, 0, stream>>>(dense, B, H, mode, thr, cap, idx, vals, pairs, cnt);
  return cuda_err(cudaGetLastError());
}

const char* activation_counts_launch(const int32_t* idx, const float* vals, int B, int cap, int H,
                                     unsigned long long* counts, cudaStream_t stream) {
  const size_t total = static_cast<size_t>(B) * cap;
  if (total == 0) return nullptr;
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  activation_counts_kernel<<<static_cast<int>(g), 256, 0, stream>>>(idx, vals, total, H, counts);
  return cuda_err(cudaGetLastError());
}

const char* coactivation_launch(const int32_t* idx, const float* vals, int B, int cap, int H, int32_t* cooc,
                                cudaStream_t stream) {
  if (cap > kCoCap) return "coactivation: more than 2048 list entries per row";
  if (B == 0 || cap == 0) return nullptr;
  int blocks = (B + kCoWarps - 1) / kCoWarps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  coactivation_kernel<<<blocks, kCoWarps * 32, 0, stream>>>(idx, vals, B, cap, H, cooc);
  return cuda_err(cudaGetLastError());
}

const char* sq_error_launch(const float* a, const float* b, size_t n, double* out, cudaStream_t stream) {
  if (n == 0) return nullptr;
  size_t g = (n + 255) / 256;
  g = (n / 8 + 255) / 256 + 1;
  if (g > 148 * 8) g = 148 * 8;
  sq_error_kernel<<<static_cast<int>(g), 256, 0, stream>>>(a, b, n, out);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

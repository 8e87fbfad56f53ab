// Exact fp32 CUDA-core encoder for a handful of rows: z[r, h] = x[rows[r], :] . W[h, :] + b[h].
// Used for rows whose tensor-core selection could not be certified (qsae_encode_topk, exact=1)
// and as a GPU-side cross-check in the tests. Classic shared-memory tiling; not a throughput path.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qsae {

namespace {

constexpr int TH = 64;   // latents per block
constexpr int TR = 16;   // rows per block
constexpr int TK = 32;   // K slice

__global__ void __launch_bounds__(256)
encode_dense_kernel(const float* __restrict__ x, const int32_t* __restrict__ rows, int R,
                    const float* __restrict__ w, const float* __restrict__ bias, int H, int D, int act,
                    float* __restrict__ z) {
  __shared__ float ws[TH][TK + 1];
  __shared__ float xs[TR][TK + 1];
  const int h0 = blockIdx.x * TH, r0 = blockIdx.y * TR;
  const int tx = threadIdx.x & 63;   // latent within tile
  const int ty = threadIdx.x >> 6;   // 0..3 -> rows ty, ty+4, ty+8, ty+12
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < D; k0 += TK) {
    for (int i = threadIdx.x; i < TH * TK; i += 256) {
      const int hh = i / TK, kk = i % TK;
      ws[hh][kk] = (h0 + hh < H && k0 + kk < D) ? w[static_cast<size_t>(h0 + hh) * D + k0 + kk] : 0.f;
    }
    for (int i = threadIdx.x; i < TR * TK; i += 256) {
      const int rr = i / TK, kk = i % TK;
      float v = 0.f;
      if (r0 + rr < R && k0 + kk < D) {
        const int src = rows ? rows[r0 + rr] : r0 + rr;
        v = x[static_cast<size_t>(src) * D + k0 + kk];
      }
      xs[rr][kk] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float wv = ws[tx][kk];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(xs[ty + 4 * q][kk], wv, acc[q]);
    }
    __syncthreads();
  }
  const int h = h0 + tx;
  if (h < H) {
    const float b = bias ? bias[h] : 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = r0 + ty + 4 * q;
      if (r < R) {
        float v = acc[q] + b;
        if (act == 1) v = fmaxf(v, 0.f);
        z[static_cast<size_t>(r) * H + h] = v;
      }
    }
  }
}

}  // namespace

const char* encode_dense_launch(const float* x, const int32_t* rows, int R, const float* w,
                                const float* bias, int H, int D, int act, float* z,
                                cudaStream_t stream) {
  dim3 grid((H + TH - 1) / TH, (R + TR - 1) / TR);
  encode_dense_kernel<<<grid, 256, 0, stream>>>(x, rows, R, w, bias, H, D, act, z);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

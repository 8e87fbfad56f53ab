// Exact top-k of ONE row by one 256-thread block (CUDA-core dot products over the whole dictionary), shared by
// the stand-alone rescue kernel (rescue.cu) and the tail kernel of the prior path (select_topk.cu). Values are
// computed with the same per-lane FMA order and shuffle tree as the fp32 re-scoring in select_topk.cu, from fp32
// operands (exact mode) or from the bf16 operands the tensor cores saw (products exact in fp32).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"
#include "topk_common.cuh"

namespace qsae {

constexpr int kResThreads = 256;
constexpr int kResWarps = kResThreads / 32;
constexpr int kResSeg = 2048;                 // latents scored per pass (static smem stays < 48 KB)
constexpr int kResKeys = kResSeg + 256;       // survivors kept between passes (k_sel <= 224) + one pass
constexpr int kResSort = 256;                 // >= kMaxK, power of two

static __device__ __forceinline__ float4 bf16x4_to_float4(uint2 u) {
  float4 f;
  f.x = __uint_as_float(u.x << 16);
  f.y = __uint_as_float(u.x & 0xFFFF0000u);
  f.z = __uint_as_float(u.y << 16);
  f.w = __uint_as_float(u.y & 0xFFFF0000u);
  return f;
}


// Must be called by all kResThreads threads of the block; ends with the outputs of `row` written (no trailing
// barrier: callers synchronise before reusing shared memory or reading the outputs).
static __device__ __noinline__ void rescue_one_row(const RescueLaunch& p, int row) {
  __shared__ float4 xs[128];                       // one x row, D <= 512
  __shared__ float zseg[kResSeg];
  __shared__ uint64_t keys[kResKeys];
  __shared__ int s_cnt;
  __shared__ float s_thr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int D = p.D, H = p.H;
  {
    __syncthreads();
    for (int q = threadIdx.x; q < 128; q += kResThreads) {
      const int d = q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < D) {
        if (p.exact) v = *reinterpret_cast<const float4*>(p.x_f32 + static_cast<size_t>(row) * D + d);
        else v = bf16x4_to_float4(*reinterpret_cast<const uint2*>(p.x_bf16 + static_cast<size_t>(row) * D + d));
      }
      xs[q] = v;
    }
    if (threadIdx.x == 0) { s_cnt = 0; s_thr = -3.402823466e+38f; }
    __syncthreads();
    float4 xr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xr[c] = xs[c * 32 + lane];

    for (int seg0 = 0; seg0 < H; seg0 += kResSeg) {
      const int seg_len = min(kResSeg, H - seg0);
      // ---- score the segment: one latent per warp per trip
      for (int e = warp; e < seg_len; e += kResWarps) {
        const int h = seg0 + e;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int d = c * 128 + lane * 4;
          if (d < D) {
            float4 w;
            if (p.exact) w = __ldg(reinterpret_cast<const float4*>(p.w_f32 + static_cast<size_t>(h) * D + d));
            else w = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(p.w_bf16 + static_cast<size_t>(h) * D + d)));
            acc = fmaf(xr[c].x, w.x, acc); acc = fmaf(xr[c].y, w.y, acc);
            acc = fmaf(xr[c].z, w.z, acc); acc = fmaf(xr[c].w, w.w, acc);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
        float s = acc + __ldg(p.bias + h);
        if (p.act == 1) s = fmaxf(s, 0.f);
        if (lane == 0) zseg[e] = s;
      }
      __syncthreads();
      // ---- keep everything >= the running threshold (order is irrelevant: composite keys)
      const float thr = s_thr;
      for (int e = threadIdx.x; e < seg_len; e += kResThreads) {
        const float v = zseg[e];
        if (v >= thr) keys[atomicAdd(&s_cnt, 1)] = make_sort_key(v, static_cast<uint32_t>(seg0 + e));
      }
      __syncthreads();
      // ---- warp 0: cut to the k_sel largest composite keys; later segments only hold higher
      //      columns, so values equal to the new k-th value can no longer win
      if (warp == 0) {
        const int n = s_cnt;
        if (n > p.k_sel) {
          uint64_t T = 0ull;
#pragma unroll 1
          for (int bit = 63; bit >= 0; --bit) {
            const uint64_t probe = T | (1ull << bit);
            int c = 0;
            for (int e = lane; e < n; e += 32) c += (keys[e] >= probe) ? 1 : 0;
            c = __reduce_add_sync(full, c);
            if (c >= p.k_sel) T = probe;
            if (c == p.k_sel) break;
          }
          int out = 0;
          for (int base = 0; base < n; base += 32) {
            const int e = base + lane;
            const uint64_t key = (e < n) ? keys[e] : 0ull;
            const bool keep = (e < n) && (key >= T);
            const unsigned b = __ballot_sync(full, keep);
            __syncwarp();
            if (keep) keys[out + __popc(b & lt_mask)] = key;
            out += __popc(b);
          }
          // the k_sel-th largest value: minimum over the kept keys
          uint64_t mn = ~0ull;
          __syncwarp();
          for (int e = lane; e < out; e += 32) mn = min(mn, keys[e]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(full, mn, o));
          if (lane == 0) {
            s_cnt = out;
            s_thr = key_to_float(float_to_key(sort_key_value(mn)) + 1u);
          }
        }
      }
      __syncthreads();
    }

    // ---- sort the survivors (<= k_sel <= kResSort) and emit
    const int n = s_cnt;
    for (int e = n + threadIdx.x; e < kResSort; e += kResThreads) keys[e] = 0ull;
    __syncthreads();
    for (int size = 2; size <= kResSort; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = threadIdx.x; t < (kResSort >> 1); t += kResThreads) {
          const int pos = ((t / stride) * (stride << 1)) + (t % stride);
          const int partner = pos + stride;
          const bool desc = (pos & size) == 0;
          const uint64_t a = keys[pos], b = keys[partner];
          if ((a < b) == desc) { keys[pos] = b; keys[partner] = a; }
        }
        __syncthreads();
      }
    }
    for (int j = threadIdx.x; j < p.k_out; j += kResThreads) {
      const bool valid = j < n;
      p.out_vals[static_cast<size_t>(row) * p.k_out + j] = valid ? sort_key_value(keys[j]) : 0.f;
      p.out_idx[static_cast<size_t>(row) * p.k_out + j] = valid ? static_cast<int32_t>(sort_key_col(keys[j])) : -1;
    }
    if (p.out_flags != nullptr && threadIdx.x == 0) p.out_flags[row] = 0;  // exact by construction
  }
}

}  // namespace qsae

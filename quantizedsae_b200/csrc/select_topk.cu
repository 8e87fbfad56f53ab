// Merge the per-sub-stream survivor lists of one row into the final ordered top-k
// (value desc, column asc), optionally re-scoring the survivors exactly in fp32 first.
// One warp per row. Also: survivor generation from a dense [R, H] matrix (qsae_topk_dense).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"
#include "topk_common.cuh"

namespace qsae {

namespace {

constexpr int kSelWarps = 4;

// shared memory per warp: n_max gathered keys + ksort selected keys (+ ksort bf16 scores if exact)
__global__ void __launch_bounds__(kSelWarps * 32)
select_topk_kernel(SelectLaunch p, int n_max, int ksort) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kSelWarps + warp;
  if (row >= p.B) return;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint64_t* keys = reinterpret_cast<uint64_t*>(sel_smem) + static_cast<size_t>(warp) * (n_max + ksort);
  uint64_t* sel = keys + n_max;

  // ---- gather
  int n = 0;
  const uint2* cand = reinterpret_cast<const uint2*>(p.cand);
  for (int s = 0; s < p.nsub; ++s) {
    const size_t slot = static_cast<size_t>(row) * p.nsub + s;
    const int c = min(p.cand_cnt[slot], kCandCap);
    const uint2* src = cand + slot * kCandCap;
    for (int e = lane; e < c; e += 32) {
      const uint2 t = src[e];
      keys[n + e] = make_sort_key(__uint_as_float(t.x), t.y);
    }
    n += c;
  }
  __syncwarp();

  // ---- k_sel largest composite keys (unique, so the count lands on k_sel exactly)
  const int k_sel = min(p.k_sel, n);
  uint64_t T = 0ull;
  if (n > k_sel) {
#pragma unroll 1
    for (int bit = 63; bit >= 0; --bit) {
      const uint64_t probe = T | (1ull << bit);
      int c = 0;
      for (int e = lane; e < n; e += 32) c += (keys[e] >= probe) ? 1 : 0;
      c = __reduce_add_sync(full, c);
      if (c >= k_sel) T = probe;
      if (c == k_sel) break;
    }
  }
  int out = 0;
  for (int base = 0; base < n; base += 32) {
    const int e = base + lane;
    const uint64_t key = (e < n) ? keys[e] : 0ull;
    const bool keep = (e < n) && (key >= T);
    const unsigned b = __ballot_sync(full, keep);
    if (keep) sel[out + __popc(b & lt_mask)] = key;
    out += __popc(b);
  }
  for (int e = out + lane; e < ksort; e += 32) sel[e] = 0ull;
  __syncwarp();

  // ---- optional exact fp32 re-scoring of the selected survivors
  float worst_bf16 = INFINITY;  // weakest tensor-core score among the selected
  float max_dev = 0.f;          // largest |fp32 - tensor-core| seen on this row
  if (p.exact) {
    const int D = p.D;
    float4 xr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int d = c * 128 + lane * 4;
      xr[c] = (d < D) ? *reinterpret_cast<const float4*>(p.x_f32 + static_cast<size_t>(row) * D + d)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int j = 0; j < out; ++j) {
      const uint64_t key = sel[j];
      const uint32_t col = sort_key_col(key);
      const float* wrow = p.w_f32 + static_cast<size_t>(col) * D;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + d));
          acc = fmaf(xr[c].x, w.x, acc);
          acc = fmaf(xr[c].y, w.y, acc);
          acc = fmaf(xr[c].z, w.z, acc);
          acc = fmaf(xr[c].w, w.w, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
      float s = acc + __ldg(p.bias + col);
      if (p.act == 1) s = fmaxf(s, 0.f);
      const float old = sort_key_value(key);
      worst_bf16 = fminf(worst_bf16, old);
      max_dev = fmaxf(max_dev, fabsf(s - old));
      if (lane == 0) sel[j] = make_sort_key(s, col);
    }
    __syncwarp();
  }

  // ---- bitonic sort of sel[0, ksort) descending
  for (int size = 2; size <= ksort; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = lane; t < (ksort >> 1); t += 32) {
        const int pos = ((t / stride) * (stride << 1)) + (t % stride);
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
  }

  // ---- emit
  for (int j = lane; j < p.k_out; j += 32) {
    const uint64_t key = sel[j];
    const bool valid = j < out;
    p.out_vals[static_cast<size_t>(row) * p.k_out + j] = valid ? sort_key_value(key) : 0.f;
    p.out_idx[static_cast<size_t>(row) * p.k_out + j] = valid ? static_cast<int32_t>(sort_key_col(key)) : -1;
  }
  if (p.out_flags != nullptr && lane == 0) {
    int flag = 0;
    if (p.exact && n > k_sel && p.k_out <= out) {
      // every dropped candidate scored <= worst_bf16 on the tensor cores; the selection is
      // certified when even 4x the largest observed rounding deviation cannot lift one of
      // them over the exact k-th value
      const float kth = sort_key_value(sel[p.k_out - 1]);
      if (!(worst_bf16 + 4.f * max_dev < kth)) flag = 1;
    }
    p.out_flags[row] = flag;
  }
}

// Streaming survivors from a dense row: same threshold + compaction scheme as the fused
// epilogue, one warp per row, lanes over 32 consecutive columns.
__global__ void __launch_bounds__(kSelWarps * 32)
dense_candidates_kernel(const float* __restrict__ z, int R, int H, int k, uint2* cand, int* cand_cnt) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kSelWarps + warp;
  if (row >= R) return;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint2* rb = cand + static_cast<size_t>(row) * kCandCap;
  const float* zr = z + static_cast<size_t>(row) * H;
  int cnt = 0;
  float thr = -INFINITY;
  for (int base = 0; base < H; base += 32) {
    const int c = base + lane;
    const float v = (c < H) ? zr[c] : -INFINITY;
    const bool keep = v > thr;
    const unsigned b = __ballot_sync(full, keep);
    if (keep) rb[cnt + __popc(b & lt_mask)] = make_uint2(__float_as_uint(v), static_cast<uint32_t>(c));
    cnt += __popc(b);
    if (cnt > kCandCap - 32) {
      __syncwarp();
      cnt = warp_compact_row(rb, cnt, k, lane, &thr);
      __syncwarp();
    }
  }
  if (cnt > k) {
    __syncwarp();
    cnt = warp_compact_row(rb, cnt, k, lane, &thr);
  }
  if (lane == 0) cand_cnt[row] = cnt;
}

int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace

const char* select_topk_launch(const SelectLaunch& p, cudaStream_t stream) {
  const int ksort = next_pow2(p.k_sel < 2 ? 2 : p.k_sel);
  const int n_max = p.nsub * (p.k_sel < kCandCap ? p.k_sel : kCandCap);
  const size_t smem = static_cast<size_t>(kSelWarps) * (n_max + ksort) * sizeof(uint64_t);
  if (smem > 200 * 1024) return "select_topk: too many survivors per row for shared memory";
  static size_t smem_attr = 0;
  if (smem > 48 * 1024 && smem > smem_attr) {
    cudaError_t e = cudaFuncSetAttribute(select_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    smem_attr = smem;
  }
  const int blocks = (p.B + kSelWarps - 1) / kSelWarps;
  select_topk_kernel<<<blocks, kSelWarps * 32, smem, stream>>>(p, n_max, ksort);
  return cuda_err(cudaGetLastError());
}

const char* dense_candidates_launch(const float* z, int R, int H, int k, void* cand, int* cand_cnt,
                                    cudaStream_t stream) {
  const int blocks = (R + kSelWarps - 1) / kSelWarps;
  dense_candidates_kernel<<<blocks, kSelWarps * 32, 0, stream>>>(z, R, H, k,
                                                                 reinterpret_cast<uint2*>(cand), cand_cnt);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

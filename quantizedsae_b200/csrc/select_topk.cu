// Merge the per-sub-stream survivor lists of one row into the final ordered top-k
// (value desc, column asc), optionally re-scoring the survivors exactly in fp32 first.
// One warp per row. Also: survivor generation from a dense [R, H] matrix (qsae_topk_dense).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include "decode_common.cuh"
#include "kernels.h"
#include "rescue_common.cuh"
#include "topk_common.cuh"

namespace qsae {

namespace {

constexpr int kMaxSelWarps = 4;   // dense_candidates_kernel: rows per block
constexpr int kSelThreads = 256;  // select_topk_kernel: one block per row

// index of survivor list (row, s); cand_cnt == nullptr means every list holds exactly `cap` entries
__device__ __forceinline__ size_t list_slot(const SelectLaunch& p, int row, int s) {
  const long long rs = p.row_stride ? p.row_stride : p.nsub;
  const long long ss = p.sub_stride ? p.sub_stride : 1;
  return static_cast<size_t>(row * rs + s * ss);
}
__device__ __forceinline__ int list_count(const SelectLaunch& p, size_t slot) {
  return p.cand_cnt ? min(p.cand_cnt[slot], p.cap) : p.cap;
}
// first entry of list (row, s): local layout, or the owner's buffer in peer memory
__device__ __forceinline__ const uint2* list_ptr(const SelectLaunch& p, int row, int s) {
  if (p.list_bases != nullptr)
    return reinterpret_cast<const uint2*>(p.list_bases[s]) + static_cast<size_t>(row) * p.cap;
  return reinterpret_cast<const uint2*>(p.cand) + list_slot(p, row, s) * p.cap;
}

__device__ __forceinline__ void atomic_min_float_key(unsigned* addr, float v) { atomicMin(addr, float_to_key(v)); }
__device__ __forceinline__ void atomic_max_float_key(unsigned* addr, float v) { atomicMax(addr, float_to_key(v)); }

// One block per row. Warps gather the row's sub-streams in parallel (batched, coalesced loads,
// filtered by the row-level bound), the block radix-selects the k_sel largest composite keys, all
// warps re-score them in fp32 when asked to, then the block sorts and emits. Any k_sel that fits
// shared memory: n_max gathered keys + ksort selected keys.
// INLINE: the caller recomputes a failing row itself (tail kernel): nothing is appended to the rescue list and the
// return value says whether the row needs the exact recomputation (uniform over the block).
template <int THREADS, bool INLINE = false>
__device__ __forceinline__ bool select_row_block(const SelectLaunch& p, int row, int n_max, int ksort,
                                                 uint8_t* sel_smem) {
  __shared__ int s_n, s_out, s_ovf, s_res;
  __shared__ unsigned long long s_listmin[32];   // smallest composite key of each list (incomplete-list check)
  __shared__ int s_hist[256];
  __shared__ int s_ctl[4];
  __shared__ unsigned s_worst_key, s_dev_key;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int nwarps = THREADS / 32;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint64_t* keys = reinterpret_cast<uint64_t*>(sel_smem);
  uint64_t* sel = keys + n_max;

  if (threadIdx.x == 0) {
    s_n = 0;
    s_out = 0;
    s_ovf = 0;
    s_worst_key = 0xFFFFFFFFu;
    s_dev_key = float_to_key(0.f);
  }
  // row-level bound: every sub-stream's bound is a lower bound of the row's k_sel-th value
  float thr_row = -INFINITY;
  if (p.cand_thr != nullptr)
    for (int s = 0; s < p.nsub; ++s)
      thr_row = fmaxf(thr_row, p.cand_thr[list_slot(p, row, s)]);
  __syncthreads();

  // ---- gather survivors >= thr_row (order is irrelevant: the composite keys are unique)
  constexpr int BATCH = 8;
  for (int s = warp; s < p.nsub; s += nwarps) {
    const size_t slot = list_slot(p, row, s);
    const int c = list_count(p, slot);
    // a threshold-only sweep (no in-kernel cut) stops collecting when a list fills up
    if (lane == 0 && p.cand_cnt != nullptr && p.cand_cnt[slot] > p.cap - 32) s_ovf = 1;
    const uint2* src = list_ptr(p, row, s);
    const uint32_t col_add = static_cast<uint32_t>(s) * static_cast<uint32_t>(p.sub_col_offset);
    uint64_t lmin = ~0ull;
    for (int base = 0; base < c; base += 32 * BATCH) {
      uint2 t[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int e = base + i * 32 + lane;
        t[i] = (e < c) ? src[e] : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        const int e = base + i * 32 + lane;
        const bool keep = (e < c) && (__uint_as_float(t[i].x) >= thr_row);
        const unsigned b = __ballot_sync(full, keep);
        if (b != 0u) {
          int pos = 0;
          if (lane == 0) pos = atomicAdd(&s_n, __popc(b));
          pos = __shfl_sync(full, pos, 0) + __popc(b & lt_mask);
          if (keep) {
            const uint64_t key = make_sort_key(__uint_as_float(t[i].x), t[i].y + col_add);
            keys[pos] = key;
            lmin = min(lmin, key);
          }
        }
      }
    }
    if (p.incomplete != nullptr && s < 32) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmin = min(lmin, __shfl_xor_sync(full, lmin, o));
      if (lane == 0) s_listmin[s] = lmin;
    }
  }
  __syncthreads();
  const int n = s_n;
  const int k_sel = min(p.k_sel, n);

  // ---- prior mode: the threshold was only probably valid. Fewer than k_sel survivors means it
  //      was too high for this row, a full list that it was far too low: hand the row to the exact
  //      rescue kernel.
  if (p.check_count != 0 && (n < p.k_sel || s_ovf != 0)) {
    if (threadIdx.x == 0) {
      if (!INLINE) p.rescue_rows[atomicAdd(p.rescue_count, 1)] = row;
      if (p.out_flags != nullptr) p.out_flags[row] = 2;
    }
    return true;
  }

  // ---- the k_sel largest composite keys -> sel[0, out)
  const uint64_t T = block_radix_select([&](int e) { return keys[e]; }, n, k_sel, s_hist, s_ctl);
  for (int base = 0; base < n; base += THREADS) {
    const int e = base + threadIdx.x;
    const uint64_t key = (e < n) ? keys[e] : 0ull;
    const bool keep = (e < n) && (key >= T);
    const unsigned b = __ballot_sync(full, keep);
    if (b != 0u) {
      int pos = 0;
      if (lane == 0) pos = atomicAdd(&s_out, __popc(b));
      pos = __shfl_sync(full, pos, 0) + __popc(b & lt_mask);
      if (keep && pos < ksort) sel[pos] = key;
    }
  }
  __syncthreads();
  const int out = min(s_out, ksort);
  for (int e = out + threadIdx.x; e < ksort; e += THREADS) sel[e] = 0ull;
  // truncated lists: a list whose every entry was selected may be hiding further winners (conservative: also
  // fires when the list's last entry is exactly the k-th)
  if (p.incomplete != nullptr && threadIdx.x < p.nsub && threadIdx.x < 32 && s_listmin[threadIdx.x] >= T &&
      s_listmin[threadIdx.x] != ~0ull)
    atomicExch(p.incomplete, 1);
  __syncthreads();

  // ---- optional exact fp32 re-scoring, candidates spread over the warps
  if (p.exact) {
    const int D = p.D;
    float4 xr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int d = c * 128 + lane * 4;
      xr[c] = (d < D) ? *reinterpret_cast<const float4*>(p.x_f32 + static_cast<size_t>(row) * D + d)
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float worst = INFINITY, dev = 0.f;
#pragma unroll 1
    for (int j = warp * 2; j < out; j += nwarps * 2) {  // two candidates per trip
      const bool has1 = (j + 1) < out;
      const uint64_t key0 = sel[j];
      const uint64_t key1 = has1 ? sel[j + 1] : key0;
      const uint32_t col0 = sort_key_col(key0), col1 = sort_key_col(key1);
      const float* w0 = p.w_f32 + static_cast<size_t>(col0) * D;
      const float* w1 = p.w_f32 + static_cast<size_t>(col1) * D;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(w0 + d));
          const float4 v = __ldg(reinterpret_cast<const float4*>(w1 + d));
          a0 = fmaf(xr[c].x, u.x, a0); a0 = fmaf(xr[c].y, u.y, a0);
          a0 = fmaf(xr[c].z, u.z, a0); a0 = fmaf(xr[c].w, u.w, a0);
          a1 = fmaf(xr[c].x, v.x, a1); a1 = fmaf(xr[c].y, v.y, a1);
          a1 = fmaf(xr[c].z, v.z, a1); a1 = fmaf(xr[c].w, v.w, a1);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(full, a0, o);
        a1 += __shfl_xor_sync(full, a1, o);
      }
      float s0 = a0 + __ldg(p.bias + col0);
      float s1 = a1 + __ldg(p.bias + col1);
      if (p.act == 1) { s0 = fmaxf(s0, 0.f); s1 = fmaxf(s1, 0.f); }
      const float old0 = sort_key_value(key0), old1 = sort_key_value(key1);
      worst = fminf(worst, fminf(old0, old1));
      dev = fmaxf(dev, fmaxf(fabsf(s0 - old0), fabsf(s1 - old1)));
      if (lane == 0) {
        sel[j] = make_sort_key(s0, col0);
        if (has1) sel[j + 1] = make_sort_key(s1, col1);
      }
    }
    if (lane == 0) {
      atomic_min_float_key(&s_worst_key, worst);
      atomic_max_float_key(&s_dev_key, dev);
    }
  }
  __syncthreads();

  // ---- sort (the gathered keys are no longer needed: their buffer is the scratch of the two-part sort)
  __shared__ int s_pos[2];
  // (fast mode, on request: the winners as a set -- the decoders and the dense latent do not depend on the order)
  const uint64_t* sorted = (p.unsorted != 0 && !p.exact) ? sel
                                                         : block_sort_selected(sel, out, ksort, keys, n_max, s_hist, s_ctl, s_pos);

  // ---- emit
  for (int j = threadIdx.x; j < p.k_out; j += THREADS) {
    const uint64_t key = sorted[j];
    const bool valid = j < out;
    p.out_vals[static_cast<size_t>(row) * p.k_out + j] = valid ? sort_key_value(key) : 0.f;
    p.out_idx[static_cast<size_t>(row) * p.k_out + j] = valid ? static_cast<int32_t>(sort_key_col(key)) : -1;
  }
  if (threadIdx.x == 0) {
    int flag = 0;
    if (p.exact && n > k_sel && p.k_out <= out) {
      // every dropped candidate scored <= worst_bf16 on the tensor cores; the selection is
      // certified when even 4x the largest observed rounding deviation cannot lift one of
      // them over the exact k-th value
      const float worst_bf16 = key_to_float(s_worst_key);
      const float max_dev = key_to_float(s_dev_key);
      const float kth = sort_key_value(sorted[p.k_out - 1]);
      if (!(worst_bf16 + 4.f * max_dev < kth)) flag = 1;
    }
    // an uncertified row is recomputed exactly when a rescue pass follows (it then clears the flag)
    if (flag != 0 && p.rescue_count != nullptr && !INLINE) p.rescue_rows[atomicAdd(p.rescue_count, 1)] = row;
    if (p.out_flags != nullptr) p.out_flags[row] = flag;
    s_res = (flag != 0 && p.rescue_count != nullptr) ? 1 : 0;
  }
  if (INLINE) {
    __syncthreads();
    return s_res != 0;
  }
  return false;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS >= 512 ? 2 : 1)
select_topk_kernel(SelectLaunch p, int n_max, int ksort) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  select_row_block<THREADS>(p, blockIdx.x, n_max, ksort, sel_smem);
}

// the same merge for an explicit (device-side) list of rows: persistent small grid
__global__ void __launch_bounds__(kSelThreads)
select_topk_list_kernel(SelectLaunch p, int n_max, int ksort, const int* count, const int32_t* rows) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  const int n = min(*count, p.B);
  for (int li = blockIdx.x; li < n; li += gridDim.x) {
    select_row_block<kSelThreads>(p, rows[li], n_max, ksort, sel_smem);
    __syncthreads();
  }
}

// Top-k of one dense row z[0, H) (any k): radix select on the composite keys built on the fly, then
// sort in shared memory. Leaves the ordered keys in sel[0, ksort) (zero padded); returns their count.
template <int THREADS, bool BYPASS_L1 = false>
__device__ __forceinline__ int dense_row_topk(const float* zptr, int H, int k, int ksort, uint64_t* sel, int* hist,
                                              int* ctl, int* s_out) {
  // BYPASS_L1: the row was written by other blocks of this launch (L1 is not coherent across SMs)
  auto z = [&](int e) { return BYPASS_L1 ? __ldcg(zptr + e) : zptr[e]; };
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  if (threadIdx.x == 0) *s_out = 0;
  const uint64_t T = block_radix_select([&](int e) { return make_sort_key(z(e), static_cast<uint32_t>(e)); }, H, k,
                                        hist, ctl);
  __syncthreads();
  for (int base = 0; base < H; base += THREADS) {
    const int e = base + threadIdx.x;
    const uint64_t key = (e < H) ? make_sort_key(z(e), static_cast<uint32_t>(e)) : 0ull;
    const bool keep = (e < H) && (key >= T);
    const unsigned b = __ballot_sync(full, keep);
    if (b != 0u) {
      int pos = 0;
      if (lane == 0) pos = atomicAdd(s_out, __popc(b));
      pos = __shfl_sync(full, pos, 0) + __popc(b & lt_mask);
      if (keep && pos < ksort) sel[pos] = key;
    }
  }
  __syncthreads();
  const int out = min(*s_out, ksort);
  for (int e = out + threadIdx.x; e < ksort; e += THREADS) sel[e] = 0ull;
  __syncthreads();
  if (ksort >= 2048) block_bitonic_desc_regs<8>(sel, ksort);
  else block_bitonic_desc(sel, ksort);
  return out;
}

__device__ __forceinline__ void emit_sorted(const uint64_t* sel, int out, int row, int k_out, float* out_vals,
                                            int32_t* out_idx) {
  for (int j = threadIdx.x; j < k_out; j += blockDim.x) {
    const uint64_t key = sel[j];
    const bool valid = j < out;
    out_vals[static_cast<size_t>(row) * k_out + j] = valid ? sort_key_value(key) : 0.f;
    out_idx[static_cast<size_t>(row) * k_out + j] = valid ? static_cast<int32_t>(sort_key_col(key)) : -1;
  }
}

// Decode of one row by the calling warp from the (values, indices) this block has just written to global memory
// (callers synchronise the block first).
__device__ __forceinline__ void decode_row_from_outputs(const SelectLaunch& p, int row, int lane) {
  const unsigned full = 0xffffffffu;
  const int k = p.k_out;
  const volatile float* vrow = p.out_vals + static_cast<size_t>(row) * k;
  const volatile int32_t* irow = p.out_idx + static_cast<size_t>(row) * k;
  float amax = 0.f;
  bool bad = false;
  for (int e = lane; e < k; e += 32) {
    if (irow[e] >= 0) {
      const float v = vrow[e];
      bad |= !(fabsf(v) <= 3.0e38f);
      amax = fmaxf(amax, fabsf(v));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(full, amax, o));
  bad = __any_sync(full, bad);
  Int4RowDecoder<2, true, false> dec;
  dec.begin(amax, bad, k);
  for (int base = 0; base < k; base += 32) {
    const int e = base + lane;
    const int my_i = (e < k) ? irow[e] : -1;
    const float my_v = (my_i >= 0) ? vrow[e] : 0.f;
    dec.template add_chunk<false>(my_v, my_i, min(32, k - base), p.dec_packed, 64, lane);
  }
  dec.finish(p.dec_scale, p.dec_bias, p.dec_recon + static_cast<size_t>(row) * 512, 512, lane);
}

// Tail of the prior path in one launch: see select_tail_launch in kernels.h.
__global__ void __launch_bounds__(kSelThreads)
select_tail_kernel(SelectLaunch p, RescueLaunch r, int n_max, int ksort, const int* ovf_count, const int32_t* ovf_rows) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows listed by the merge kernel for the exact recomputation; this kernel appends nothing to that list
  const int n_res = min(*r.rescue_count, p.B);
  const int n_ovf = min(*ovf_count, p.B);
  for (int li = blockIdx.x; li < n_ovf; li += gridDim.x) {
    const int row = ovf_rows[li];
    const bool redo = select_row_block<kSelThreads, true>(p, row, n_max, ksort, sel_smem);
    __syncthreads();
    if (redo) {
      rescue_one_row(r, row);
      __syncthreads();
    }
    if (p.dec_kind == 1 && warp == 0) decode_row_from_outputs(p, row, lane);
    __syncthreads();
  }
  // ---- rows whose prior failed the count check. The first kRescueSlots of them are recomputed by the whole grid:
  //      every block scores its slice of the dictionary for the row (CUDA cores, the arithmetic of rescue_one_row,
  //      four latents in flight per warp) into a dense scratch row, and the block that finishes last selects, emits
  //      and decodes. (A single block per row, as for the remaining rows below, takes milliseconds per row: its
  //      33 MB dictionary sweep is bound by the latency of one load chain per warp.)
  const int n_fast = (r.z_scratch != nullptr) ? min(n_res, kRescueSlots) : 0;
  if (n_fast > 0) {
    __shared__ float4 xs[128];
    __shared__ int s_last;
    const unsigned full = 0xffffffffu;
    const int D = r.D, H = r.H;
    const int per = (H + gridDim.x - 1) / gridDim.x;
    const int h0 = blockIdx.x * per, h1 = min(H, h0 + per);
    uint64_t* sel = reinterpret_cast<uint64_t*>(sel_smem);
    int* hist = reinterpret_cast<int*>(sel_smem + static_cast<size_t>(ksort) * 8);    // 256 + 4 + 1 ints behind the sort buffer
    for (int li = 0; li < n_fast; ++li) {
      const int row = r.rescue_rows[li];
      float* z = r.z_scratch + static_cast<size_t>(li) * H;
      __syncthreads();
      for (int q = threadIdx.x; q < 128; q += kSelThreads) {
        const int d = q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (d < D) {
          if (r.exact) v = *reinterpret_cast<const float4*>(r.x_f32 + static_cast<size_t>(row) * D + d);
          else v = bf16x4_to_float4(*reinterpret_cast<const uint2*>(r.x_bf16 + static_cast<size_t>(row) * D + d));
        }
        xs[q] = v;
      }
      __syncthreads();
      float4 xr[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) xr[c] = xs[c * 32 + lane];
      for (int hb = h0 + warp * 4; hb < h1; hb += (kSelThreads / 32) * 4) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float4 w[4][4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int h = min(hb + t, h1 - 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int d = c * 128 + lane * 4;
            w[t][c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (d < D) {
              if (r.exact) w[t][c] = __ldg(reinterpret_cast<const float4*>(r.w_f32 + static_cast<size_t>(h) * D + d));
              else w[t][c] = bf16x4_to_float4(__ldg(reinterpret_cast<const uint2*>(r.w_bf16 + static_cast<size_t>(h) * D + d)));
            }
          }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[t] = fmaf(xr[c].x, w[t][c].x, acc[t]); acc[t] = fmaf(xr[c].y, w[t][c].y, acc[t]);
            acc[t] = fmaf(xr[c].z, w[t][c].z, acc[t]); acc[t] = fmaf(xr[c].w, w[t][c].w, acc[t]);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int t = 0; t < 4; ++t) acc[t] += __shfl_xor_sync(full, acc[t], o);
        }
        if (lane < 4 && hb + lane < h1) {
          float sc = (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]) + __ldg(r.bias + hb + lane);
          if (r.act == 1) sc = fmaxf(sc, 0.f);
          z[hb + lane] = sc;
        }
      }
      __threadfence();
      __syncthreads();
      if (threadIdx.x == 0) s_last = (atomicAdd(&r.slot_done[li], 1) == static_cast<int>(gridDim.x) - 1) ? 1 : 0;
      __syncthreads();
      if (s_last) {   // block-uniform: every other block's slice is visible (their fence precedes their ticket)
        __threadfence();
        const int ksort_out = ksort;
        const int out = dense_row_topk<kSelThreads, true>(z, H, r.k_out, ksort_out, sel, hist, hist + 256, hist + 260);
        emit_sorted(sel, out, row, r.k_out, r.out_vals, r.out_idx);
        if (r.out_flags != nullptr && threadIdx.x == 0) r.out_flags[row] = 0;  // exact by construction
        __syncthreads();
        if (p.dec_kind == 1 && warp == 0) decode_row_from_outputs(p, row, lane);
      }
    }
    __syncthreads();
  }
  for (int li = n_fast + blockIdx.x; li < n_res; li += gridDim.x) {
    const int row = r.rescue_rows[li];
    rescue_one_row(r, row);
    __syncthreads();
    if (p.dec_kind == 1 && warp == 0) decode_row_from_outputs(p, row, lane);
    __syncthreads();
  }
}

constexpr int kLargeThreads = 512;   // long sorts (ksort >= kLongSort): 16 warps, register-resident bitonic chunks
constexpr int kLongSort = 2048;

// dense [R, H] matrix -> ordered top-k per row, any k (block per row)
__global__ void __launch_bounds__(kLargeThreads)
select_dense_kernel(const float* __restrict__ z, int R, int H, int k, int ksort, float* out_vals, int32_t* out_idx) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  __shared__ int s_hist[256];
  __shared__ int s_ctl[4];
  __shared__ int s_out;
  uint64_t* sel = reinterpret_cast<uint64_t*>(sel_smem);
  for (int row = blockIdx.x; row < R; row += gridDim.x) {
    const int out = dense_row_topk<kLargeThreads>(z + static_cast<size_t>(row) * H, H, k, ksort, sel, s_hist, s_ctl, &s_out);
    emit_sorted(sel, out, row, k, out_vals, out_idx);
    __syncthreads();
  }
}

// Exact recomputation of listed rows for any k (the large-k counterpart of rescue.cu): dense
// pre-activations of the row into this block's scratch line (CUDA cores, same per-lane FMA order and
// shuffle tree as the fp32 re-scoring), then the dense-row top-k above.
__global__ void __launch_bounds__(kLargeThreads)
rescue_large_kernel(RescueLaunch p, float* scratch, int ksort) {
  extern __shared__ __align__(16) uint8_t sel_smem[];
  __shared__ float4 xs[128];
  __shared__ int s_hist[256];
  __shared__ int s_ctl[4];
  __shared__ int s_out;
  uint64_t* sel = reinterpret_cast<uint64_t*>(sel_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned full = 0xffffffffu;
  const int D = p.D, H = p.H;
  const int count = min(*p.rescue_count, p.B);
  float* zrow = scratch + static_cast<size_t>(blockIdx.x) * H;
  for (int li = blockIdx.x; li < count; li += gridDim.x) {
    const int row = p.rescue_rows[li];
    __syncthreads();
    for (int q = threadIdx.x; q < 128; q += kLargeThreads) {
      const int d = q * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < D) {
        if (p.exact) {
          v = *reinterpret_cast<const float4*>(p.x_f32 + static_cast<size_t>(row) * D + d);
        } else {
          const uint2 u = *reinterpret_cast<const uint2*>(p.x_bf16 + static_cast<size_t>(row) * D + d);
          v = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                          __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
        }
      }
      xs[q] = v;
    }
    __syncthreads();
    float4 xr[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) xr[c] = xs[c * 32 + lane];
    for (int h = warp; h < H; h += kLargeThreads / 32) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          float4 w;
          if (p.exact) {
            w = __ldg(reinterpret_cast<const float4*>(p.w_f32 + static_cast<size_t>(h) * D + d));
          } else {
            const uint2 u = __ldg(reinterpret_cast<const uint2*>(p.w_bf16 + static_cast<size_t>(h) * D + d));
            w = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u),
                            __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xFFFF0000u));
          }
          acc = fmaf(xr[c].x, w.x, acc); acc = fmaf(xr[c].y, w.y, acc);
          acc = fmaf(xr[c].z, w.z, acc); acc = fmaf(xr[c].w, w.w, acc);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
      float s = acc + __ldg(p.bias + h);
      if (p.act == 1) s = fmaxf(s, 0.f);
      if (lane == 0) zrow[h] = s;
    }
    __syncthreads();
    const int out = dense_row_topk<kLargeThreads>(zrow, H, p.k_out, ksort, sel, s_hist, s_ctl, &s_out);
    emit_sorted(sel, out, row, p.k_out, p.out_vals, p.out_idx);
    if (p.out_flags != nullptr && threadIdx.x == 0) p.out_flags[row] = 0;  // exact by construction
  }
}

// ------------------------------------------------------------------------------------------
// Warp-per-row merge for the prior-threshold path, where a row has only a few hundred survivors:
// gather into a per-warp staging area, keep the composite keys in registers, bisect there.
// No block-level synchronisation at all. Rows that do not fit the staging area are listed for
// the block-per-row kernel above.
// ------------------------------------------------------------------------------------------
constexpr int kSmallWarps = 4;

// the k_sel largest of the n staged composite keys -> sel[0, k_sel) (unordered); R keys per lane.
// Bisection on the 32-bit value keys only, starting below the bits all survivors share (they come
// from a narrow tail of the row's distribution) and stopping as soon as a probe isolates exactly
// k_sel entries: ~10-15 rounds of R compares instead of 64 rounds over 64-bit keys. Ties at the k-th
// value (rare) are resolved by a second bisection over the column halves of the tied keys.
template <int R>
__device__ __forceinline__ int small_select(uint64_t* stage, int n, int k_sel, int lane) {
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint64_t key[R];
  uint32_t vk[R];
#pragma unroll
  for (int i = 0; i < R; ++i) {
    key[i] = (lane + 32 * i < n) ? stage[lane + 32 * i] : 0ull;
    vk[i] = static_cast<uint32_t>(key[i] >> 32);
  }
  __syncwarp();
  uint32_t T = 0u;        // keep: vk > T, or vk == T and low half >= L
  uint32_t L = 0u;
  bool strict_only = false;
  if (n > k_sel) {
    uint32_t a = 0xFFFFFFFFu, o = 0u;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      if (lane + 32 * i < n) { a &= vk[i]; o |= vk[i]; }
    }
    a = __reduce_and_sync(full, a);
    o = __reduce_or_sync(full, o);
    const uint32_t diff = a ^ o;
    int c_ge = n;         // count(vk >= T) for the current T
    T = a;                // all shared leading bits; the differing bits start cleared
    if (diff != 0u) {
      const int top = 31 - __clz(diff);
      T = a & ~((2u << top) - 1u);
#pragma unroll 1
      for (int bit = top; bit >= 0; --bit) {
        const uint32_t probe = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) c += (vk[i] >= probe) ? 1 : 0;   // padding keys are 0 < probe
        c = __reduce_add_sync(full, c);
        if (c >= k_sel) { T = probe; c_ge = c; }
        if (c == k_sel) break;
      }
    }
    if (c_ge > k_sel) {
      // T is the k-th largest value and several entries carry it: keep every vk > T and the
      // (k_sel - count(vk > T)) tied entries with the largest low halves (lowest columns)
      int c_gt = 0;
#pragma unroll
      for (int i = 0; i < R; ++i) c_gt += (vk[i] > T) ? 1 : 0;
      c_gt = __reduce_add_sync(full, c_gt);
      const int need = k_sel - c_gt;
#pragma unroll 1
      for (int bit = 31; bit >= 0; --bit) {
        const uint32_t probe = L | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < R; ++i) c += (vk[i] == T && static_cast<uint32_t>(key[i]) >= probe) ? 1 : 0;
        c = __reduce_add_sync(full, c);
        if (c >= need) L = probe;
        if (c == need) break;
      }
    } else {
      strict_only = true;  // vk >= T is exactly the k_sel largest
    }
  } else {
    strict_only = true;    // keep everything
  }
  int out = 0;  // the staging area is free again: the keys live in registers
#pragma unroll
  for (int i = 0; i < R; ++i) {
    const bool valid = lane + 32 * i < n;
    const bool keep = valid && (strict_only ? (vk[i] >= T)
                                            : (vk[i] > T || (vk[i] == T && static_cast<uint32_t>(key[i]) >= L)));
    const unsigned b = __ballot_sync(full, keep);
    if (keep) stage[out + __popc(b & lt_mask)] = key[i];
    out += __popc(b);
  }
  return out;
}

// Bitonic sort (descending) of 32 * P composite keys held P per lane, element e = j * 32 + lane:
// strides below 32 exchange through shuffles, larger strides are register moves.
template <int P, typename K>
__device__ __forceinline__ void warp_bitonic_desc(K (&k)[P], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int size = 2; size <= 32 * P; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride >= 32) {
        const int js = stride >> 5;
#pragma unroll
        for (int j = 0; j < P; ++j) {
          if ((j & js) == 0) {
            const bool desc = ((j * 32) & size) == 0;   // size >= 64 here: depends on j only
            const K a = k[j], b = k[j | js];
            const bool swap = (a < b) == desc;
            k[j] = swap ? b : a;
            k[j | js] = swap ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < P; ++j) {
          const bool desc = (((j * 32) | lane) & size) == 0;
          const bool lower = (lane & stride) == 0;
          const K mine = k[j];
          const K other = __shfl_xor_sync(full, mine, stride);
          const bool mine_big = mine > other;
          const bool take_max = (lower == desc);
          k[j] = (take_max == mine_big) ? mine : other;
        }
      }
    }
  }
}

// TAIL: 32 P + 1 keys (k_sel = 65, the reference's default k at H = 32768, would otherwise pay for the 128-key
// network: 28 compare-exchange steps over four keys per lane instead of 21 over two). The smallest key is moved out
// first -- it is the last element of the order whatever the rest looks like -- and the other 32 P are sorted.
template <int P, bool TAIL = false>
__device__ __forceinline__ void sort_and_emit(const SelectLaunch& p, int row, uint64_t* stage, int stage_rows, int out, int n,
                                              int k_sel, float worst_bf16, float max_dev, int lane) {
  const unsigned full = 0xffffffffu;
  const uint64_t* sel = stage;
  uint64_t k[P];
#pragma unroll
  for (int j = 0; j < P; ++j) k[j] = sel[j * 32 + lane];
  uint64_t tail = 0ull;
  if constexpr (TAIL) {
    tail = sel[32 * P];                      // zero when the row has at most 32 P entries
    uint64_t mn = k[0];
#pragma unroll
    for (int j = 1; j < P; ++j) mn = (k[j] < mn) ? k[j] : mn;
    const uint32_t hmin = __reduce_min_sync(full, static_cast<uint32_t>(mn >> 32));
    const uint32_t lmin = __reduce_min_sync(full, static_cast<uint32_t>(mn >> 32) == hmin ? static_cast<uint32_t>(mn) : 0xFFFFFFFFu);
    const uint64_t wmin = (static_cast<uint64_t>(hmin) << 32) | lmin;
    if (wmin < tail) {                       // warp-uniform; the keys of a full row are distinct (distinct columns)
#pragma unroll
      for (int j = 0; j < P; ++j) k[j] = (k[j] == wmin) ? tail : k[j];
      tail = wmin;
    }
  }
  warp_bitonic_desc<P>(k, lane);
  auto finish_row = [&](uint64_t kth_key) {  // by the thread that holds output position k_out - 1
    int flag = 0;
    if (p.exact && n > k_sel && p.k_out <= out) {
      // every dropped candidate scored <= worst_bf16 on the tensor cores; certified when even 4x the
      // largest observed rounding deviation cannot lift one of them over the exact k-th value
      const float kth = sort_key_value(kth_key);
      if (!(worst_bf16 + 4.f * max_dev < kth)) flag = 1;
    }
    // an uncertified row is recomputed exactly when a rescue pass follows (it then clears the flag)
    if (flag != 0 && p.rescue_count != nullptr) p.rescue_rows[atomicAdd(p.rescue_count, 1)] = row;
    if (p.out_flags != nullptr) p.out_flags[row] = flag;
  };
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int e = j * 32 + lane;
    if (e < p.k_out) {
      const bool valid = e < out;
      p.out_vals[static_cast<size_t>(row) * p.k_out + e] = valid ? sort_key_value(k[j]) : 0.f;
      p.out_idx[static_cast<size_t>(row) * p.k_out + e] = valid ? static_cast<int32_t>(sort_key_col(k[j])) : -1;
    }
    if (e == p.k_out - 1) finish_row(k[j]);
  }
  const bool tail_out = TAIL && 32 * P < p.k_out;    // the tail is output position 32 P
  const bool tail_valid = tail_out && 32 * P < out;
  if (tail_out && lane == 0) {
    p.out_vals[static_cast<size_t>(row) * p.k_out + 32 * P] = tail_valid ? sort_key_value(tail) : 0.f;
    p.out_idx[static_cast<size_t>(row) * p.k_out + 32 * P] = tail_valid ? static_cast<int32_t>(sort_key_col(tail)) : -1;
    if (32 * P == p.k_out - 1) finish_row(tail);
  }
  if (p.dec_kind == 1) {
    // fused decode (sae/binary.py:38 restricted to the k winners), straight from the sorted registers. A row that
    // was just listed for the exact recomputation is decoded again by the tail kernel.
    float amax = 0.f;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < P; ++j) {
      const int e = j * 32 + lane;
      if (e < p.k_out && e < out) {
        const float v = sort_key_value(k[j]);
        bad |= !(fabsf(v) <= 3.0e38f);
        amax = fmaxf(amax, fabsf(v));
      }
    }
    if (tail_valid) {
      const float v = sort_key_value(tail);
      bad |= !(fabsf(v) <= 3.0e38f);
      amax = fmaxf(amax, fabsf(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(full, amax, o));
    bad = __any_sync(full, bad);
    Int4RowDecoder<2, true, false> dec;
    dec.begin(amax, bad, p.k_out);
    // (Staging the gathered rows through the warp's shared-memory area with cp.async, a whole batch in flight at
    // once, was measured and is slower: 41.5 vs 38.9 us for the stage at B = 4096 -- LDGSTS issues at 8 cycles per
    // instruction and the kernel is not bound by the latency of these L2-resident rows.)
#pragma unroll
    for (int j = 0; j < P; ++j) {
      if (j * 32 < p.k_out) {
        const int e = j * 32 + lane;
        const bool valid = e < p.k_out && e < out;
        dec.template add_chunk<false>(valid ? sort_key_value(k[j]) : 0.f, valid ? static_cast<int>(sort_key_col(k[j])) : -1,
                                      min(32, p.k_out - j * 32), p.dec_packed, 64, lane);
      }
    }
    if (tail_out) {
      const bool valid = tail_valid && lane == 0;
      dec.template add_chunk<false>(valid ? sort_key_value(tail) : 0.f, valid ? static_cast<int>(sort_key_col(tail)) : -1, 1,
                                    p.dec_packed, 64, lane);
    }
    dec.finish(p.dec_scale, p.dec_bias, p.dec_recon + static_cast<size_t>(row) * 512, 512, lane);
  }
}

// One row: gather -> select -> (re-score) -> sort -> emit. R = survivor keys per lane (capacity 32 * R).
template <int R>
__device__ __forceinline__ void small_row(const SelectLaunch& p, int row, int ksort, uint64_t* stage, int* ovf_count,
                                          int32_t* ovf_rows, int lane) {
  const unsigned full = 0xffffffffu;
  // ---- counts and offsets of the row's sub-streams (nsub <= 32)
  const int my_c = (lane < p.nsub) ? list_count(p, list_slot(p, row, lane)) : 0;
  int incl = my_c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += t;
  }
  const int n_all = __shfl_sync(full, incl, 31);
  if (p.check_count != 0 && n_all < p.k_sel) {  // prior threshold too high for this row
    if (lane == 0) {
      p.rescue_rows[atomicAdd(p.rescue_count, 1)] = row;
      if (p.out_flags != nullptr) p.out_flags[row] = 2;
    }
    return;
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  int n = n_all;
  if (n_all <= 32 * R) {
    // ---- gather (every survivor already passed the row's threshold in the sweep). Lists are short (a few dozen
    //      entries): the first 32 entries of four lists are loaded before any is stored, so a row costs
    //      nsub / 4 dependent L2 round trips instead of nsub; longer lists finish in a second loop.
    for (int s0 = 0; s0 < p.nsub; s0 += 4) {
      uint2 t[4];
      int cs[4], offs[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int s = min(s0 + q, p.nsub - 1);
        cs[q] = (s0 + q < p.nsub) ? __shfl_sync(full, my_c, s) : 0;
        offs[q] = __shfl_sync(full, incl, s) - __shfl_sync(full, my_c, s);
        t[q] = make_uint2(0u, 0u);
        if (lane < cs[q]) t[q] = list_ptr(p, row, s)[lane];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t col_add = static_cast<uint32_t>(s0 + q) * static_cast<uint32_t>(p.sub_col_offset);
        if (lane < cs[q]) stage[offs[q] + lane] = make_sort_key(__uint_as_float(t[q].x), t[q].y + col_add);
      }
    }
    for (int s = 0; s < p.nsub; ++s) {
      const int c = __shfl_sync(full, my_c, s);
      if (c <= 32) continue;
      const int off = __shfl_sync(full, incl, s) - c;
      const uint2* src = list_ptr(p, row, s);
      const uint32_t col_add = static_cast<uint32_t>(s) * static_cast<uint32_t>(p.sub_col_offset);
#pragma unroll 4
      for (int e = 32 + lane; e < c; e += 32) {
        const uint2 t = src[e];
        stage[off + e] = make_sort_key(__uint_as_float(t.x), t.y + col_add);
      }
    }
    __syncwarp();
  } else {
    // ---- more survivors than the registers hold (a few per cent of the rows: the survivor count of a prior
    //      threshold is negative-binomial). Two passes instead of a second kernel tier: the k_sel-th largest value
    //      of the FIRST 32 R survivors is a lower bound of the row's k_sel-th largest, and only about
    //      k_sel * n / (32 R) survivors reach it.
    int filled = 0;
    for (int s = 0; s < p.nsub && filled < 32 * R; ++s) {
      const int c = min(__shfl_sync(full, my_c, s), 32 * R - filled);
      const uint2* src = list_ptr(p, row, s);
      const uint32_t col_add = static_cast<uint32_t>(s) * static_cast<uint32_t>(p.sub_col_offset);
#pragma unroll 4
      for (int e = lane; e < c; e += 32) {
        const uint2 t = src[e];
        stage[filled + e] = make_sort_key(__uint_as_float(t.x), t.y + col_add);
      }
      filled += c;
    }
    __syncwarp();
    const int kept = small_select<R>(stage, 32 * R, p.k_sel, lane);
    __syncwarp();
    uint32_t tmin = 0xFFFFFFFFu;
    for (int e = lane; e < kept; e += 32) tmin = min(tmin, static_cast<uint32_t>(stage[e] >> 32));
    tmin = __reduce_min_sync(full, tmin);
    __syncwarp();
    int pos = 0;
    for (int s = 0; s < p.nsub; ++s) {
      const int c = __shfl_sync(full, my_c, s);
      const uint2* src = list_ptr(p, row, s);
      const uint32_t col_add = static_cast<uint32_t>(s) * static_cast<uint32_t>(p.sub_col_offset);
#pragma unroll 2
      for (int base = 0; base < c; base += 32) {
        const int e = base + lane;
        uint2 t = make_uint2(0u, 0u);
        if (e < c) t = src[e];
        const bool keep = (e < c) && (float_to_key(__uint_as_float(t.x)) >= tmin);
        const unsigned b = __ballot_sync(full, keep);
        const int at = pos + __popc(b & lt_mask);
        if (keep && at < 32 * R) stage[at] = make_sort_key(__uint_as_float(t.x), t.y + col_add);
        pos += __popc(b);
      }
    }
    __syncwarp();
    if (pos > 32 * R) {   // floods of equal values: block-per-row kernel
      if (lane == 0) ovf_rows[atomicAdd(ovf_count, 1)] = row;
      return;
    }
    n = pos;
  }

  const int k_sel = min(p.k_sel, n);
  const int out = small_select<R>(stage, n, k_sel, lane);
  n = n_all;   // candidates were dropped iff the row had more than k_sel survivors in total
  uint64_t* sel = stage;
  for (int e = out + lane; e < ksort; e += 32) sel[e] = 0ull;
  __syncwarp();

  // ---- optional exact fp32 re-scoring
  float worst_bf16 = INFINITY, max_dev = 0.f;
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
    for (int j = 0; j < out; j += 2) {
      const bool has1 = (j + 1) < out;
      const uint64_t key0 = sel[j];
      const uint64_t key1 = has1 ? sel[j + 1] : key0;
      const uint32_t col0 = sort_key_col(key0), col1 = sort_key_col(key1);
      const float* w0 = p.w_f32 + static_cast<size_t>(col0) * D;
      const float* w1 = p.w_f32 + static_cast<size_t>(col1) * D;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int d = c * 128 + lane * 4;
        if (d < D) {
          const float4 u = __ldg(reinterpret_cast<const float4*>(w0 + d));
          const float4 v = __ldg(reinterpret_cast<const float4*>(w1 + d));
          a0 = fmaf(xr[c].x, u.x, a0); a0 = fmaf(xr[c].y, u.y, a0);
          a0 = fmaf(xr[c].z, u.z, a0); a0 = fmaf(xr[c].w, u.w, a0);
          a1 = fmaf(xr[c].x, v.x, a1); a1 = fmaf(xr[c].y, v.y, a1);
          a1 = fmaf(xr[c].z, v.z, a1); a1 = fmaf(xr[c].w, v.w, a1);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(full, a0, o);
        a1 += __shfl_xor_sync(full, a1, o);
      }
      float s0 = a0 + __ldg(p.bias + col0);
      float s1 = a1 + __ldg(p.bias + col1);
      if (p.act == 1) { s0 = fmaxf(s0, 0.f); s1 = fmaxf(s1, 0.f); }
      const float old0 = sort_key_value(key0), old1 = sort_key_value(key1);
      worst_bf16 = fminf(worst_bf16, fminf(old0, old1));
      max_dev = fmaxf(max_dev, fmaxf(fabsf(s0 - old0), fabsf(s1 - old1)));
      __syncwarp();
      if (lane == 0) {
        sel[j] = make_sort_key(s0, col0);
        if (has1) sel[j + 1] = make_sort_key(s1, col1);
      }
    }
    __syncwarp();
  }

  // ---- sort in registers and emit (value desc, column asc)
  constexpr int kStageRows = R;   // 256-byte dictionary rows that fit the warp's 32 R x 8-byte staging area
  if (ksort <= 32) sort_and_emit<1>(p, row, sel, kStageRows, out, n, k_sel, worst_bf16, max_dev, lane);
  else if (ksort <= 64) sort_and_emit<2>(p, row, sel, kStageRows, out, n, k_sel, worst_bf16, max_dev, lane);
  else if (p.k_sel == 65) sort_and_emit<2, true>(p, row, sel, kStageRows, out, n, k_sel, worst_bf16, max_dev, lane);
  else if (ksort <= 128) sort_and_emit<4>(p, row, sel, kStageRows, out, n, k_sel, worst_bf16, max_dev, lane);
  else sort_and_emit<8>(p, row, sel, kStageRows, out, n, k_sel, worst_bf16, max_dev, lane);
  __syncwarp();
}

// rows == nullptr: one warp per row of the batch; otherwise a persistent grid over rows[0, *count)
template <int R>
__global__ void __launch_bounds__(kSmallWarps * 32, R <= 16 ? 8 : 1)
select_small_kernel(SelectLaunch p, int ksort, const int* count, const int32_t* rows, int* ovf_count,
                    int32_t* ovf_rows) {
  __shared__ uint64_t stage_all[kSmallWarps][32 * R];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (rows == nullptr) {
    const int row = blockIdx.x * kSmallWarps + warp;
    if (row < p.B) small_row<R>(p, row, ksort, stage_all[warp], ovf_count, ovf_rows, lane);
  } else {
    const int n_list = min(*count, p.B);
    for (int li = blockIdx.x * kSmallWarps + warp; li < n_list; li += gridDim.x * kSmallWarps)
      small_row<R>(p, rows[li], ksort, stage_all[warp], ovf_count, ovf_rows, lane);
  }
}

// prior[row] = m-th largest of the row's n = nsub * kTopM pre-pass values (n <= 512, P = values per lane,
// m <= 32): one warp per row sorts the ordered-integer images of the floats in registers (bitonic network,
// shuffles below stride 32) and lane m - 1 holds the answer.
template <int P>
__global__ void __launch_bounds__(256)
prior_from_top_kernel(const float* __restrict__ top, int B, int n, int m, float* __restrict__ prior) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= B) return;
  const float* src = top + static_cast<size_t>(row) * n;
  uint32_t key[P];
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int slot = lane + 32 * i;
    key[i] = (slot < n) ? float_to_key(src[slot]) : 0u;   // padding sorts last
  }
  // m-th largest by m rounds of "take the maximum, retire every key equal to it" (m <= 32 rounds of ~3 P instructions
  // instead of a bitonic network over 32 P keys). Equal keys retire together, so with ties the result is at most the
  // true m-th largest: still a valid lower bound (the sweep's result never depends on it).
  uint32_t T = 0u;
#pragma unroll 1
  for (int round = 0; round < m; ++round) {
    uint32_t mx = key[0];
#pragma unroll
    for (int i = 1; i < P; ++i) mx = max(mx, key[i]);
    T = __reduce_max_sync(0xffffffffu, mx);
#pragma unroll
    for (int i = 0; i < P; ++i) key[i] = (key[i] == T) ? 0u : key[i];
  }
  if (lane == 0) prior[row] = key_to_float(T);
}

// Streaming survivors from a dense row, one warp per row, lanes over 32 consecutive columns.
// Lane l only ever sees columns == l (mod 32): with m = ceil(k/32) running maxima per lane there
// are 32*m seen values >= min over lanes of the m-th largest, the same class bound as in the
// fused epilogue (here m <= 7 covers kMaxK).
constexpr int kDenseTop = (kMaxK + 31) / 32;

__global__ void __launch_bounds__(kMaxSelWarps * 32)
dense_candidates_kernel(const float* __restrict__ z, int R, int H, int k, uint2* cand, int* cand_cnt) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * kMaxSelWarps + warp;
  if (row >= R) return;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  uint2* rb = cand + static_cast<size_t>(row) * kDenseCap;
  const float* zr = z + static_cast<size_t>(row) * H;
  const int m = (k + 31) / 32;
  float top[kDenseTop];
#pragma unroll
  for (int i = 0; i < kDenseTop; ++i) top[i] = -INFINITY;
  int cnt = 0;
  float thr_ge = -INFINITY, thr_gt = -INFINITY;
  for (int base = 0; base < H; base += 32) {
    const int c = base + lane;
    const float v = (c < H) ? zr[c] : -INFINITY;
    // insert v into this lane's sorted top-m (descending)
    float carry = v;
#pragma unroll
    for (int i = 0; i < kDenseTop; ++i) {
      if (i < m) {
        const float hi = fmaxf(top[i], carry);
        carry = fminf(top[i], carry);
        top[i] = hi;
      }
    }
    float mine = top[0];
#pragma unroll
    for (int i = 1; i < kDenseTop; ++i)
      if (i < m) mine = top[i];
    float bound = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bound = fminf(bound, __shfl_xor_sync(full, bound, o));
    thr_ge = fmaxf(thr_ge, bound);
    const bool keep = (v >= thr_ge) && (v > thr_gt);
    const unsigned b = __ballot_sync(full, keep);
    if (keep) rb[cnt + __popc(b & lt_mask)] = make_uint2(__float_as_uint(v), static_cast<uint32_t>(c));
    cnt += __popc(b);
    if (cnt > kDenseCap - 32) {
      __syncwarp();
      cnt = warp_compact_row_generic(rb, cnt, k, lane, &thr_gt);
      __syncwarp();
    }
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
  const int n_max = p.nsub * p.cap;
  const size_t smem = static_cast<size_t>(n_max + ksort) * sizeof(uint64_t);
  const size_t budget = kSelectSmemBudget;
  if (smem > budget) return "select_topk: too many survivors per row for shared memory";
  const bool large = ksort >= kLongSort;   // long sorts: 16 warps per row
  static bool attr_set[2] = {false, false};
  if (smem > 48 * 1024 && !attr_set[large]) {
    cudaError_t e = large ? cudaFuncSetAttribute(select_topk_kernel<kLargeThreads>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget))
                          : cudaFuncSetAttribute(select_topk_kernel<kSelThreads>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set[large] = true;
  }
  if (large) select_topk_kernel<kLargeThreads><<<p.B, kLargeThreads, smem, stream>>>(p, n_max, ksort);
  else select_topk_kernel<kSelThreads><<<p.B, kSelThreads, smem, stream>>>(p, n_max, ksort);
  return cuda_err(cudaGetLastError());
}

const char* select_dense_launch(const float* z, int R, int H, int k, int num_sms, float* out_vals, int32_t* out_idx,
                                cudaStream_t stream) {
  const int ksort = next_pow2(k < 2 ? 2 : k);
  const size_t smem = static_cast<size_t>(ksort) * sizeof(uint64_t);
  if (smem > kSelectSmemBudget) return "select_dense: k too large for shared memory";
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(select_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSelectSmemBudget));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  const int blocks = R < num_sms * 2 ? R : num_sms * 2;
  select_dense_kernel<<<blocks, kLargeThreads, smem, stream>>>(z, R, H, k, ksort, out_vals, out_idx);
  return cuda_err(cudaGetLastError());
}

int rescue_large_blocks(int H, int num_sms) {
  // one scratch line of H floats per block, at most 256 MB in total
  long long b = (256ll << 20) / (static_cast<long long>(H) * 4);
  if (b > num_sms) b = num_sms;
  if (b < 8) b = 8;
  return static_cast<int>(b);
}
size_t rescue_large_scratch_bytes(int H, int num_sms) {
  return static_cast<size_t>(rescue_large_blocks(H, num_sms)) * H * sizeof(float);
}

const char* rescue_large_launch(const RescueLaunch& p, void* scratch, int num_sms, cudaStream_t stream) {
  const int ksort = next_pow2(p.k_out < 2 ? 2 : p.k_out);
  const size_t smem = static_cast<size_t>(ksort) * sizeof(uint64_t);
  if (smem > kSelectSmemBudget) return "rescue_large: k too large for shared memory";
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(rescue_large_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kSelectSmemBudget));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  rescue_large_kernel<<<rescue_large_blocks(p.H, num_sms), kLargeThreads, smem, stream>>>(p, static_cast<float*>(scratch), ksort);
  return cuda_err(cudaGetLastError());
}

const char* select_small_launch(const SelectLaunch& p, int tier, const int* count, const int32_t* rows, int num_sms,
                                int* ovf_count, int32_t* ovf_rows, cudaStream_t stream) {
  if (p.nsub > 32) return "select_small: at most 32 sub-streams per row";
  const int ksort = next_pow2(p.k_sel < 32 ? 32 : p.k_sel);
  if (ksort > 256) return "select_small: k too large";
  const int blocks = rows ? num_sms * 4 : (p.B + kSmallWarps - 1) / kSmallWarps;
  switch (tier) {
    case 8: select_small_kernel<8><<<blocks, kSmallWarps * 32, 0, stream>>>(p, ksort, count, rows, ovf_count, ovf_rows); break;
    case 12: select_small_kernel<12><<<blocks, kSmallWarps * 32, 0, stream>>>(p, ksort, count, rows, ovf_count, ovf_rows); break;
    case 16: select_small_kernel<16><<<blocks, kSmallWarps * 32, 0, stream>>>(p, ksort, count, rows, ovf_count, ovf_rows); break;
    case 32: select_small_kernel<32><<<blocks, kSmallWarps * 32, 0, stream>>>(p, ksort, count, rows, ovf_count, ovf_rows); break;
    default: return "select_small: tier must be 8, 12, 16 or 32 keys per lane";
  }
  return cuda_err(cudaGetLastError());
}

const char* select_topk_list_launch(const SelectLaunch& p, const int* count, const int32_t* rows, int num_sms,
                                    cudaStream_t stream) {
  const int ksort = next_pow2(p.k_sel < 2 ? 2 : p.k_sel);
  const int n_max = p.nsub * p.cap;
  const size_t smem = static_cast<size_t>(n_max + ksort) * sizeof(uint64_t);
  const size_t budget = 200 * 1024;
  if (smem > budget) return "select_topk: too many survivors per row for shared memory";
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(select_topk_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(budget));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  // persistent grid, as many blocks as fit an SM's shared memory (at most 6); returns at once when the list is empty
  int per_sm = static_cast<int>((200 * 1024) / (smem + 2048));
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  select_topk_list_kernel<<<num_sms * per_sm, kSelThreads, smem, stream>>>(p, n_max, ksort, count, rows);
  return cuda_err(cudaGetLastError());
}

const char* select_tail_launch(const SelectLaunch& p, const RescueLaunch& r, const int* ovf_count, const int32_t* ovf_rows,
                               int num_sms, cudaStream_t stream) {
  const int ksort = next_pow2(p.k_sel < 2 ? 2 : p.k_sel);
  const int n_max = p.nsub * p.cap;
  const size_t smem = static_cast<size_t>(n_max + ksort) * sizeof(uint64_t);
  const size_t budget = 160 * 1024;   // next to ~31 KB of static shared memory (rescue_one_row + the block select)
  if (smem > budget) return "select_tail: too many survivors per row for shared memory";
  if (kResThreads != kSelThreads) return "select_tail: block size mismatch";
  static bool attr_set = false;
  if (smem > 16 * 1024 && !attr_set) {
    cudaError_t e = cudaFuncSetAttribute(select_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(budget));
    if (e != cudaSuccess) return cudaGetErrorString(e);
    attr_set = true;
  }
  select_tail_kernel<<<num_sms, kSelThreads, smem, stream>>>(p, r, n_max, ksort, ovf_count, ovf_rows);
  return cuda_err(cudaGetLastError());
}

const char* prior_from_top_launch(const float* top, int B, int nsub, int m, float* prior, cudaStream_t stream) {
  const int n = nsub * kTopM;
  const int blocks = (B + 7) / 8;
  if (n <= 32) prior_from_top_kernel<1><<<blocks, 256, 0, stream>>>(top, B, n, m, prior);
  else if (n <= 64) prior_from_top_kernel<2><<<blocks, 256, 0, stream>>>(top, B, n, m, prior);
  else if (n <= 128) prior_from_top_kernel<4><<<blocks, 256, 0, stream>>>(top, B, n, m, prior);
  else if (n <= 256) prior_from_top_kernel<8><<<blocks, 256, 0, stream>>>(top, B, n, m, prior);
  else if (n <= 512) prior_from_top_kernel<16><<<blocks, 256, 0, stream>>>(top, B, n, m, prior);
  else return "prior_from_top: too many sub-streams";
  return cuda_err(cudaGetLastError());
}

const char* dense_candidates_launch(const float* z, int R, int H, int k, void* cand, int* cand_cnt,
                                    cudaStream_t stream) {
  const int blocks = (R + kMaxSelWarps - 1) / kMaxSelWarps;
  dense_candidates_kernel<<<blocks, kMaxSelWarps * 32, 0, stream>>>(z, R, H, k,
                                                                    reinterpret_cast<uint2*>(cand), cand_cnt);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

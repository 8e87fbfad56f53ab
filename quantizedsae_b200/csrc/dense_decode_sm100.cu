// Dense decoder GEMM for sm_100a:  C[B, N] = A[B, K] * Bw[N, K]^T,  K = hidden_dim (large),
// N = input_dim (<= 512), bf16 operands, fp32 accumulation in TMEM.
//
// t_sae (sae/ternary.py:116-122, :41-52): recon = h @ T^T with dense ReLU latents h [B, H] and the
// exact-ternary decoder T = sign(W) * (|W| >= 0.5), W = decoder.weight [D, H] -- already K-major
// for the UMMA B operand. One CTA owns 128 rows and the whole N (one or two 256-column TMEM
// accumulators); both operands stream through a TMA/mbarrier ring over K. K is split over
// gridDim.x CTAs that write fp32 partial tiles, reduced in a fixed order by a second kernel
// (deterministic, unlike atomics). An optional second pass re-runs the K range with a second A
// operand into the same accumulator: A = hi + lo bf16 split of an fp32 matrix gives ~2^-17
// relative operand error instead of 2^-9 (exact mode).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tmap_cache.cuh"

namespace qsae {

namespace {

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;
constexpr int kABytes = BM * BK * 2;   // 16 KiB
constexpr int kBBytes = BN * BK * 2;   // 32 KiB per N tile
constexpr int kThreads = 256;          // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue

// PAIR (cta_group::2): two CTAs (256 rows) execute one MMA; each keeps its own A tile and HALF of
// every B (= T) tile -- 128 of the 256 output columns of each N tile -- so a stage is 48 KiB instead
// of 80 KiB: four stages instead of two, and 40 % less operand traffic per flop (the single-CTA
// kernel pulls 80 KB per 1024 MMA cycles, above the L2 slice throughput of the part).
template <int N_TILES, bool PAIR>
struct Cfg {
  static constexpr int kBTileBytes = PAIR ? kBBytes / 2 : kBBytes;
  static constexpr int kStageBytes = kABytes + N_TILES * kBTileBytes;
  static constexpr int kStages = PAIR ? 4 : ((N_TILES == 1) ? 4 : 2);
  static constexpr int kTmemCols = N_TILES * BN;  // 256 or 512
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kSmem = kBarOff + 8 * (2 * kStages + 1) + 16;
};

struct DecodeLaunch {
  int B, K, N;
  int k_chunks;          // ceil(K / 64)
  int chunks_per_split;
  int n_passes;          // 1, or 2 (hi then lo)
  float* partial;        // [splits][B][N]
};

template <int N_TILES, bool PAIR>
__global__ void __launch_bounds__(kThreads, 1)
dense_decode_kernel(const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1,
                    const __grid_constant__ CUtensorMap tmap_b, DecodeLaunch p) {
  using C = Cfg<N_TILES, PAIR>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* acc_full = bars + 2 * C::kStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pairs sit next to each other along grid.x (two row blocks); otherwise x = K split, y = row block
  const int split = PAIR ? blockIdx.y : blockIdx.x;
  const int m0 = (PAIR ? blockIdx.x : blockIdx.y) * BM;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0u;
  const int kc_begin = split * p.chunks_per_split;
  const int kc_end = min(p.k_chunks, kc_begin + p.chunks_per_split);
  const int n_chunks = max(0, kc_end - kc_begin) * p.n_passes;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0u) {
      printf("qsae: dynamic shared memory is not 1024-byte aligned\n");
      __trap();
    }
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  if (warp == 2) {
    if constexpr (PAIR) tmem_alloc_pair<C::kTmemCols>(tmem_ptr);
    else tmem_alloc<C::kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0 && n_chunks > 0) {
      tma_prefetch_desc(&tmap_a0);
      tma_prefetch_desc(&tmap_a1);
      tma_prefetch_desc(&tmap_b);
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < p.n_passes; ++pass) {
        const CUtensorMap* ta = pass == 0 ? &tmap_a0 : &tmap_a1;
        for (int kc = kc_begin; kc < kc_end; ++kc) {
          mbar_wait(&empty[stage], phase ^ 1u);
          uint8_t* st = smem + stage * C::kStageBytes;
          if constexpr (PAIR) {
            // both CTAs' loads are credited to the leader's barrier (2 x 48 KiB per stage)
            const uint32_t full_leader = mapa_u32(smem_u32(&full[stage]), 0);
            if (leader) mbar_arrive_expect_tx(&full[stage], 2 * C::kStageBytes);
            tma_load_2d_pair(st, ta, full_leader, kc * BK, m0, kPolicyEvictFirst);
#pragma unroll
            for (int nt = 0; nt < N_TILES; ++nt)
              tma_load_2d_pair(st + kABytes + nt * C::kBTileBytes, &tmap_b, full_leader, kc * BK,
                               nt * BN + static_cast<int>(cta_rank) * (BN / 2), kPolicyEvictLast);
          } else {
            mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
            tma_load_2d(st, ta, &full[stage], kc * BK, m0, kPolicyEvictFirst);
#pragma unroll
            for (int nt = 0; nt < N_TILES; ++nt)
              tma_load_2d(st + kABytes + nt * kBBytes, &tmap_b, &full[stage], kc * BK, nt * BN, kPolicyEvictLast);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_chunks > 0 && (!PAIR || leader)) {
      constexpr uint32_t idesc = umma_idesc_bf16_f32(PAIR ? 2 * BM : BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_chunks; ++i) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + stage * C::kStageBytes);
        const uint64_t a_desc = umma_desc_kmajor_sw128(st);
#pragma unroll
        for (int nt = 0; nt < N_TILES; ++nt) {
          const uint64_t b_desc = umma_desc_kmajor_sw128(st + kABytes + nt * C::kBTileBytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            if constexpr (PAIR)
              umma_f16_ss_pair(tmem_base + nt * BN, a_desc + ks * 2, b_desc + ks * 2, idesc, (i | ks) != 0 ? 1u : 0u);
            else
              umma_f16_ss(tmem_base + nt * BN, a_desc + ks * 2, b_desc + ks * 2, idesc, (i | ks) != 0 ? 1u : 0u);
          }
        }
        if constexpr (PAIR) {
          umma_commit_pair(&empty[stage], 0x3);
          if (i == n_chunks - 1) umma_commit_pair(acc_full, 0x3);
        } else {
          umma_commit(&empty[stage]);
          if (i == n_chunks - 1) umma_commit(acc_full);
        }
        if (++stage == C::kStages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp - 4;
    const int row = m0 + quad * 32 + lane;
    float* dst = p.partial + (static_cast<size_t>(split) * p.B + row) * p.N;
    if (n_chunks > 0) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
    }
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    for (int c0 = 0; c0 < p.N; c0 += 32) {
      uint32_t r[32];
      if (n_chunks > 0) {
        tmem_ld_32x32b_x32(lane_taddr + c0, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (row < p.B) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (c0 + j < p.N)   // N % 4 == 0
            *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tc_fence_after();
  if (warp == 2) {
    if constexpr (PAIR) tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int splits, size_t n,
                                       const float* __restrict__ bias, const float* __restrict__ prev, int N,
                                       float* __restrict__ out) {
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
    float4 acc = reinterpret_cast<const float4*>(partial)[i];
    for (int s = 1; s < splits; ++s) {
      const float4 v = reinterpret_cast<const float4*>(partial + static_cast<size_t>(s) * n)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (bias != nullptr) {
      const int d = static_cast<int>((i * 4) % N);
      acc.x += bias[d]; acc.y += bias[d + 1]; acc.z += bias[d + 2]; acc.w += bias[d + 3];
    }
    if (prev != nullptr) {   // cumulative outputs (q_sae levels): out = prev + this range
      const float4 q = reinterpret_cast<const float4*>(prev)[i];
      acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_tmap(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return false;
    enc = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  const TmapKey key{base, static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
                    static_cast<unsigned long long>(ld) * 2, static_cast<unsigned int>(BK), static_cast<unsigned int>(box_rows),
                    static_cast<int>(CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), static_cast<int>(CU_TENSOR_MAP_SWIZZLE_128B),
                    tmap_current_device()};
  EncodeTiledFn fn = enc;
  return tmap_cache().get(key, map, [&](CUtensorMap* m) {
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};   // row pitch in bytes (a K sub-range keeps the full pitch)
    cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
  });
}

template <int N_TILES, bool PAIR>
cudaError_t launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const DecodeLaunch& p,
                   int splits, cudaStream_t stream) {
  using C = Cfg<N_TILES, PAIR>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(dense_decode_kernel<N_TILES, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int m_tiles = (p.B + BM - 1) / BM;
  if constexpr (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((m_tiles + 1) / 2 * 2, splits);   // whole pairs along x; a padding CTA multiplies zero rows
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, dense_decode_kernel<N_TILES, PAIR>, a0, a1, b, p);
  } else {
    dim3 grid(splits, m_tiles);
    dense_decode_kernel<N_TILES, PAIR><<<grid, kThreads, C::kSmem, stream>>>(a0, a1, b, p);
    return cudaGetLastError();
  }
}

}  // namespace

int dense_decode_pick_splits(int B, int K, int num_sms) {
  // cost in units of one K chunk of MMA work per CTA: waves x (chunks per CTA + fixed prologue /
  // accumulator drain) + the partial-tile round trip through HBM, which grows with the split count
  const int m_tiles = (B + BM - 1) / BM;
  const int k_chunks = (K + BK - 1) / BK;
  int best_s = 1;
  double best = 1e30;
  for (int s = 1; s <= 16; ++s) {
    if (s > 1 && k_chunks / s < 8) break;
    const long units = static_cast<long>(m_tiles) * s;
    const long waves = (units + num_sms - 1) / num_sms;
    const int cps = (k_chunks + s - 1) / s;
    const double cost = static_cast<double>(waves) * (cps + 12) + (s > 1 ? 0.11 * s * m_tiles : 0.0);
    if (cost < best * 0.97) { best = cost; best_s = s; }
  }
  return best_s;
}

size_t dense_decode_workspace_bytes(int B, int K, int N, int num_sms) {
  return static_cast<size_t>(dense_decode_pick_splits(B, K, num_sms)) * B * N * sizeof(float);
}

const char* dense_decode_launch(const uint16_t* a_hi, const uint16_t* a_lo, int lda, const uint16_t* b_t, int ldb, int B,
                                int K, int N, const float* bias, const float* prev, float* out, void* workspace,
                                int num_sms, cudaStream_t stream) {
  if (N > 512 || (N % 4) != 0) return "dense_decode: N must be a multiple of 4, <= 512";
  if ((lda % 8) != 0 || (ldb % 8) != 0) return "dense_decode: row pitches must be multiples of 8 elements";
  CUtensorMap ta0, ta1, tb;
  if (!make_tmap(&ta0, a_hi, B, K, lda, BM)) return "cuTensorMapEncodeTiled(A) failed";
  if (!make_tmap(&ta1, a_lo ? a_lo : a_hi, B, K, lda, BM)) return "cuTensorMapEncodeTiled(A lo) failed";
  // CTA pairs need two row blocks; QSAE_DECODE_PAIR=0/1 overrides (tests and tuning experiments)
  bool pair = B > BM;
  if (tuning().decode_pair >= 0) pair = pair && tuning().decode_pair != 0;
  if (!make_tmap(&tb, b_t, N, K, ldb, pair ? BN / 2 : BN)) return "cuTensorMapEncodeTiled(B) failed";
  DecodeLaunch p;
  p.B = B; p.K = K; p.N = N;
  p.k_chunks = (K + BK - 1) / BK;
  const int splits = dense_decode_pick_splits(B, K, num_sms);
  p.chunks_per_split = (p.k_chunks + splits - 1) / splits;
  p.n_passes = a_lo ? 2 : 1;
  p.partial = static_cast<float*>(workspace);
  cudaError_t e;
  if (pair) e = (N <= 256) ? launch<1, true>(ta0, ta1, tb, p, splits, stream) : launch<2, true>(ta0, ta1, tb, p, splits, stream);
  else e = (N <= 256) ? launch<1, false>(ta0, ta1, tb, p, splits, stream) : launch<2, false>(ta0, ta1, tb, p, splits, stream);
  if (e != cudaSuccess) return cudaGetErrorString(e);
  const size_t n = static_cast<size_t>(B) * N;
  size_t g = (n / 4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  count_launches(1);
  reduce_partials_kernel<<<static_cast<int>(g), 256, 0, stream>>>(p.partial, splits, n, bias, prev, N, out);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

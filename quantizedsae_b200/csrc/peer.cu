// Peer-memory exchange for the dictionary-sharded forward (one process per GPU, NVLink / NVSwitch):
// instead of an NCCL all-gather of the candidate lists followed by the merge, and a reduce-scatter after the
// decode, every rank publishes its lists / partial reconstructions in a cudaMalloc'ed buffer that its peers
// map through CUDA IPC; the merge kernel gathers the lists of all shards straight from peer memory (P2P loads
// over NVLink, fused with the selection) and the final kernel sums the peers' partial rows it owns.
// Synchronisation: per (phase, producer) sequence flags living in the consumer's buffer, written remotely by a
// one-block signal kernel after the producing kernel (stream order), polled by a one-block wait kernel before the
// consuming kernel. The wait is bounded (QSAE_PEER_TIMEOUT_MS, default 20 s): on timeout it raises a device flag and
// returns, so a lost peer can never hang the GPU; the host side (sharded.PeerExchange) reads the flag after every
// forward and raises.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace qsae {

namespace {

__global__ void peer_signal_kernel(unsigned* const* __restrict__ targets, int n, unsigned value) {
  const int g = threadIdx.x;
  if (g >= n) return;
  __threadfence_system();
  *reinterpret_cast<volatile unsigned*>(targets[g]) = value;
  __threadfence_system();
}

__global__ void peer_wait_kernel(const unsigned* __restrict__ flags, int n, unsigned value, int* __restrict__ timed_out,
                                 long long timeout_cycles) {
  const int g = threadIdx.x;
  if (g >= n) return;
  const volatile unsigned* f = flags + g;
  const long long t0 = clock64();
  // flags only grow (sequence numbers); signed difference tolerates wrap-around
  while (static_cast<int>(*f - value) < 0) {
    if (clock64() - t0 > timeout_cycles) {
      atomicExch(timed_out, 1);
      break;
    }
    __nanosleep(200);
  }
  __threadfence_system();
}

// out[r, :] = sum_g partial_g[row_begin + r, :], g in fixed order (deterministic); float4 lanes
__global__ void __launch_bounds__(256)
reduce_partials_peer_kernel(const float* const* __restrict__ bases, int n, int row_begin, int rows, int D,
                            float* __restrict__ out) {
  const size_t total4 = static_cast<size_t>(rows) * D / 4;
  const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
  const size_t off4 = static_cast<size_t>(row_begin) * D / 4;
  for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total4; e += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int g = 0; g < n; ++g) {
      const float4 v = reinterpret_cast<const float4*>(bases[g])[off4 + e];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[e] = acc;
  }
}

}  // namespace

const char* peer_signal_launch(unsigned* const* targets, int n, unsigned value, cudaStream_t stream) {
  if (n < 1 || n > 32) return "peer_signal: 1 <= n <= 32";
  peer_signal_kernel<<<1, 32, 0, stream>>>(targets, n, value);
  return cuda_err(cudaGetLastError());
}

const char* peer_wait_launch(const unsigned* flags, int n, unsigned value, int* timed_out, cudaStream_t stream) {
  if (n < 1 || n > 32) return "peer_wait: 1 <= n <= 32";
  const long long cycles = static_cast<long long>(tuning().peer_timeout_ms > 0 ? tuning().peer_timeout_ms : 20000) * 2000000ll;   // ~2 GHz
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, n, value, timed_out, cycles);
  return cuda_err(cudaGetLastError());
}

const char* reduce_partials_peer_launch(const float* const* bases, int n, int row_begin, int rows, int D, float* out,
                                        cudaStream_t stream) {
  if ((D % 4) != 0) return "reduce_partials_peer: D must be a multiple of 4";
  if (rows <= 0) return nullptr;
  size_t g = (static_cast<size_t>(rows) * D / 4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  reduce_partials_peer_kernel<<<static_cast<int>(g), 256, 0, stream>>>(bases, n, row_begin, rows, D, out);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

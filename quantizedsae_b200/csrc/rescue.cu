// Exact per-row top-k for the rows whose prior threshold failed the merge kernel's count check
// (probability ~1e-7 per row by construction, see qsae_api.cu:choose_prior_rank). CUDA-core dot
// products over the whole dictionary for just those rows; persistent small grid that returns at
// once when the list is empty. Values are computed with the same per-lane FMA order and shuffle
// tree as the fp32 re-scoring in select_topk.cu, from fp32 operands (exact mode) or from the
// bf16 operands the tensor cores saw (products exact in fp32).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "kernels.h"
#include "rescue_common.cuh"
#include "topk_common.cuh"

namespace qsae {

namespace {

__global__ void __launch_bounds__(kResThreads)
rescue_rows_kernel(RescueLaunch p) {
  const int count = min(*p.rescue_count, p.B);
  for (int li = blockIdx.x; li < count; li += gridDim.x) {
    __syncthreads();
    rescue_one_row(p, p.rescue_rows[li]);
  }
}

}  // namespace

const char* rescue_rows_launch(const RescueLaunch& p, int num_sms, cudaStream_t stream) {
  rescue_rows_kernel<<<num_sms, kResThreads, 0, stream>>>(p);
  return cuda_err(cudaGetLastError());
}

}  // namespace qsae

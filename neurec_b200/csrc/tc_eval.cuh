// Interface between the tensor-core candidate pass (tc_eval.cu) and the evaluator (evaluator.cu).
#pragma once
#include "common.cuh"

namespace nrc {
namespace tc {

// Runs bf16 conversion + the tcgen05 candidate kernel on `st`.  On return (asynchronously)
// *cand points at [num_eval, cap] ascending candidate item ids and *cand_cnt at [num_eval]
// counts (count > cap = overflow).  Buffers are library-owned and reused between calls.
int run_candidates(const float* U, const float* V, int D, int N, const int32_t* users, int num_eval,
                   const int64_t* train_ptr, const int32_t* train_idx, int K, int cap,
                   const int32_t** cand, const int32_t** cand_cnt, cudaStream_t st);

}  // namespace tc
}  // namespace nrc

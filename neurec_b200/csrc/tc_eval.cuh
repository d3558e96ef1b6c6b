// Interface between the tensor-core candidate pass (tc_eval.cu) and the evaluator (evaluator.cu).
#pragma once
#include "common.cuh"

namespace nrc {
namespace tc {

// Runs bf16 conversion + the tcgen05 candidate kernel on `st`.  LQ = rank of the running
// threshold: every unmasked item whose score may exceed the LQ-th best score of the items before
// it is reported (LQ = min(2*top_k, N) covers every element that can enter the reference's heap).  On return (asynchronously)
// *cand points at [num_eval, cap] ascending candidate item ids and *cand_cnt at [num_eval]
// counts (count > cap = overflow).  Buffers are library-owned and reused between calls.
int run_candidates(const float* U, const float* V, int D, int N, const int32_t* users, int num_eval,
                   const int64_t* train_ptr, const int32_t* train_idx, int LQ, int cap,
                   const int32_t** cand, const int32_t** cand_cnt, cudaStream_t st);

}  // namespace tc
}  // namespace nrc

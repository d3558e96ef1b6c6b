// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// C-ABI shim around the reference's own, unmodified C++ headers so that tests and
// bench.py's cpu_baseline / --impl reference legs can call the REAL reference code.
// The headers are compiled from where they lie under /root/reference (see Makefile,
// REF=...); nothing from the reference is copied into this repository.  The resulting
// library goes to oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
//
// Wrapped reference entry points:
//   cpp_evaluate_matrix   evaluator/backend/cpp/include/evaluate.h:53-72
//   arg_top_k_2d          util/cython/include/arg_topk.h:27-45
#include <vector>
#include <unordered_set>
#include <cstdint>

#include "evaluate.h"   // reference header (-I$(REF)/evaluator/backend/cpp/include)
#include "arg_topk.h"   // reference header (-I$(REF)/util/cython/include)

extern "C" {

// test_items are passed as CSR (indptr[B+1], indices) and converted to the
// vector<unordered_set<int>> the reference wants -- the same conversion Cython does at
// cpp_evaluator.pyx:32.  `marshal_only` != 0 stops after that conversion (lets the
// bench separate marshalling from evaluation time).
void ref_cpp_evaluate_matrix(float* rating_matrix, int rating_len, int num_users,
                             const int64_t* test_indptr, const int32_t* test_indices,
                             const int32_t* metric, int metric_num, int top_k,
                             int thread_num, float* results) {
    std::vector<std::unordered_set<int>> test_items(num_users);
    for (int u = 0; u < num_users; ++u)
        for (int64_t p = test_indptr[u]; p < test_indptr[u + 1]; ++p)
            test_items[u].insert(test_indices[p]);
    std::vector<int> metric_vec(metric, metric + metric_num);
    cpp_evaluate_matrix(rating_matrix, rating_len, test_items, metric_vec, top_k,
                        thread_num, results);
}

void ref_arg_top_k_2d(float* ratings, int rating_len, int rows_num, int top_k,
                      int thread_num, int* results) {
    arg_top_k_2d(ratings, rating_len, rows_num, top_k, thread_num, results);
}

}  // extern "C"

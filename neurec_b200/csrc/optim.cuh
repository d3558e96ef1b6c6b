// Multi-tensor optimizer launch: every variable of a model updated by ONE kernel.
#pragma once
#include "common.cuh"

namespace nrc {

constexpr int kMaxOptSegs = 12;

struct OptSeg {
    float* var;
    float* grad;
    float* s0;
    float* s1;
    const int32_t* touched;  // per-row stamps (IndexedSlices variables) or nullptr
    int64_t elems;           // rows * dim
    int64_t begin;           // first global element index of this segment
    int dim;
    int dense_var;           // 1: gradient was a dense tensor (tf.layers.dense kernel / bias)
};

struct OptLaunch {
    OptSeg seg[kMaxOptSegs];
    int nseg;
    int kind;
    float h[4];
    int64_t total;
};

int opt_launch_init(OptLaunch& L, int opt_kind, const float* hyper_host);
int opt_launch_add(OptLaunch& L, float* var, float* grad, float* s0, float* s1,
                   const int32_t* touched, int64_t rows, int dim, int dense_var);
int opt_launch_run(const OptLaunch& L, int32_t stamp, cudaStream_t st);

}  // namespace nrc

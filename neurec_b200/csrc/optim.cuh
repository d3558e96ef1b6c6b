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

// One element of one variable: TensorFlow-1.12 update rules (formulas in optim.cu's header comment).
// h = hyper-parameters {lr or Adam's lr_t, beta1|rho|momentum, beta2|momentum, eps}.
__device__ __forceinline__ void opt_update(int kind, int dense_var, bool touched, float h0,
                                           float h1, float h2, float h3, float& var, float g,
                                           float& s0, float& s1) {
    switch (kind) {
        case NRC_OPT_GD:
            var = __fsub_rn(var, __fmul_rn(g, h0));
            break;
        case NRC_OPT_ADAM: {
            const float omb1 = __fsub_rn(1.0f, h1), omb2 = __fsub_rn(1.0f, h2);
            if (dense_var) {
                s0 = __fadd_rn(s0, __fmul_rn(__fsub_rn(g, s0), omb1));
                s1 = __fadd_rn(s1, __fmul_rn(__fsub_rn(__fmul_rn(g, g), s1), omb2));
                var = __fsub_rn(var, __fdiv_rn(__fmul_rn(s0, h0), __fadd_rn(__fsqrt_rn(s1), h3)));
            } else {
                s0 = __fadd_rn(__fmul_rn(s0, h1), __fmul_rn(g, omb1));
                s1 = __fadd_rn(__fmul_rn(s1, h2), __fmul_rn(__fmul_rn(g, g), omb2));
                var = __fsub_rn(var, __fdiv_rn(__fmul_rn(h0, s0), __fadd_rn(__fsqrt_rn(s1), h3)));
            }
            break;
        }
        case NRC_OPT_ADAGRAD:
            if (touched) {
                s0 = __fadd_rn(s0, __fmul_rn(g, g));
                var = __fsub_rn(var, __fmul_rn(__fmul_rn(h0, g), __fdiv_rn(1.0f, __fsqrt_rn(s0))));
            }
            break;
        case NRC_OPT_RMSPROP:
            if (touched) {  // h = {lr, rho, momentum, eps}
                if (dense_var) {   // ApplyRMSProp functor
                    s0 = __fadd_rn(s0, __fmul_rn(__fsub_rn(__fmul_rn(g, g), s0), __fsub_rn(1.0f, h1)));
                    s1 = __fadd_rn(__fmul_rn(s1, h2),
                                   __fmul_rn(__fmul_rn(h0, g), __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(s0, h3)))));
                } else {           // SparseApplyRMSProp: ms*rho + g*g*(1-rho); mom*mu + rsqrt(ms+eps)*lr*g
                    s0 = __fadd_rn(__fmul_rn(s0, h1), __fmul_rn(__fmul_rn(g, g), __fsub_rn(1.0f, h1)));
                    s1 = __fadd_rn(__fmul_rn(s1, h2),
                                   __fmul_rn(__fmul_rn(__fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(s0, h3))), h0), g));
                }
                var = __fsub_rn(var, s1);
            }
            break;
        default:  // NRC_OPT_MOMENTUM  h = {lr, momentum}
            if (touched) {
                s0 = __fadd_rn(__fmul_rn(s0, h1), g);
                var = __fsub_rn(var, __fmul_rn(s0, h0));
            }
            break;
    }
}

int opt_launch_init(OptLaunch& L, int opt_kind, const float* hyper_host);
int opt_launch_add(OptLaunch& L, float* var, float* grad, float* s0, float* s1,
                   const int32_t* touched, int64_t rows, int dim, int dense_var);
int opt_launch_run(const OptLaunch& L, int32_t stamp, cudaStream_t st);

}  // namespace nrc

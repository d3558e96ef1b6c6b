// Shapes shared by the NCF kernels (ncf.cu, ncf_epoch.cu).
#pragma once
#include "common.cuh"

namespace nrc {

constexpr int kNcfMaxLayers = 4;
constexpr int kNcfThreads = 128;   // threads per sample CTA
constexpr int kNcfWarps = 8;       // warps per CTA of the score kernel
constexpr int kWgradSlices = 16;   // batch slices of the weight-gradient kernel

struct NcfDev {
    int mf_dim, mlp_dim, n_layers, n_towers;
    int in_dim[kNcfMaxLayers], out_dim[kNcfMaxLayers];
    int w_off[kNcfMaxLayers], b_off[kNcfMaxLayers];      // offsets in the packed dense buffer
    int sw_off[kNcfMaxLayers], sb_off[kNcfMaxLayers];    // offsets in the padded smem copy (scores)
    int a_off[kNcfMaxLayers + 1];                        // activation offsets (a_0 = input)
    int tower_size, s_tower_size, act_size;
};

struct NcfPtrs {
    const float* mf_user; const float* mf_item; const float* mlp_user; const float* mlp_item;
    const float* dense;
    float* g_mf_user; float* g_mf_item; float* g_mlp_user; float* g_mlp_item; float* g_dense;
    int32_t* t_user; int32_t* t_item;
};

int ncf_make(NcfDev& S, const nrc_ncf_shape* sh);

}  // namespace nrc

// NCF family (MLP, NeuMF = GMF + MLP): fused gather -> MLP tower forward -> loss -> backward ->
// gradient accumulation, and the score-all-items predict.
//
// Replaces (reference paths):
//   model/general_recommender/NeuMF.py:69-104   _create_inference / _create_loss
//   model/general_recommender/MLP.py:57-87      same for the MLP-only model
//   model/general_recommender/NeuMF.py:158-168  predict: one forward over ALL items per user
//   util/learner.py:18-41                       pairwise / pointwise losses
// Third-party arithmetic restated (tensorflow==1.12.3, not vendored): tf.layers.dense =
// relu(x . kernel + bias) with kernel [in, out]; ReluGrad passes dy where the OUTPUT is > 0;
// prediction = reduce_sum(concat(mf_vector, mlp_vector)) -- there is no output layer
// (NeuMF.py:85).  In pairwise NeuMF tf.layers.dense is re-instantiated for the negative tower,
// so the two towers have DIFFERENT weights (NeuMF.py:81-82, 90-92; n_towers = 2); MLP.py shares
// its Dense objects (n_towers = 1).
//
// Kernel shape.  CTA = 8 warps; one warp runs one sample's towers forward and backward with the
// dense weights in shared memory (row stride out+1: conflict-free both by output column for the
// forward pass and by input row for the backward pass).  Per-sample weight gradients are NOT
// accumulated with atomics: every group of 8 samples leaves its activations and deltas in
// shared memory and the 256 threads each own a fixed set of (k, j) weight-gradient entries in
// REGISTERS, adding a_l[s][k] * delta_l[s][j] over the group; one RED.ADD per entry per CTA at
// the end.  Embedding-row gradients go straight to the dense accumulators (duplicates sum).
#include "common.cuh"
#include "optim.cuh"

namespace nrc {

constexpr int kNcfMaxLayers = 4;
constexpr int kNcfEntries = 16;   // weight-gradient entries per thread and layer
constexpr int kNcfWarps = 8;

struct NcfDev {
    int mf_dim, mlp_dim, n_layers, n_towers;
    int in_dim[kNcfMaxLayers], out_dim[kNcfMaxLayers];
    int w_off[kNcfMaxLayers], b_off[kNcfMaxLayers];      // offsets in the packed dense buffer
    int sw_off[kNcfMaxLayers], sb_off[kNcfMaxLayers];    // offsets in the padded smem copy
    int a_off[kNcfMaxLayers + 1];                        // activation offsets (a_0 = input)
    int tower_size, s_tower_size, act_size;
};

struct NcfPtrs {
    const float* mf_user; const float* mf_item; const float* mlp_user; const float* mlp_item;
    const float* dense;
    float* g_mf_user; float* g_mf_item; float* g_mlp_user; float* g_mlp_item; float* g_dense;
    int32_t* t_user; int32_t* t_item;
};

static int ncf_make(NcfDev& S, const nrc_ncf_shape* sh) {
    NRC_REQUIRE(sh != nullptr, NRC_E_VALUE, "shape is NULL");
    NRC_REQUIRE(sh->n_layers >= 0 && sh->n_layers <= kNcfMaxLayers, NRC_E_LIMIT,
                "n_layers %d outside [0, %d]", sh->n_layers, kNcfMaxLayers);
    NRC_REQUIRE(sh->n_towers == 1 || sh->n_towers == 2, NRC_E_VALUE, "n_towers must be 1 or 2");
    NRC_REQUIRE(sh->mf_dim >= 0 && sh->mlp_dim >= 0 && (sh->mf_dim > 0 || sh->n_layers > 0),
                NRC_E_VALUE, "model has neither an MF nor an MLP part");
    NRC_REQUIRE(sh->n_layers == 0 || sh->mlp_dim > 0, NRC_E_VALUE, "mlp_dim must be > 0");
    S.mf_dim = sh->mf_dim; S.mlp_dim = sh->n_layers ? sh->mlp_dim : 0;
    S.n_layers = sh->n_layers; S.n_towers = sh->n_towers;
    int in = 2 * S.mlp_dim, off = 0, soff = 0, aoff = 0;
    S.a_off[0] = 0; aoff = in;
    for (int l = 0; l < kNcfMaxLayers; ++l) {
        if (l >= S.n_layers) { S.in_dim[l] = S.out_dim[l] = 0; S.w_off[l] = S.b_off[l] = S.sw_off[l] = S.sb_off[l] = 0; S.a_off[l + 1] = aoff; continue; }
        const int out = sh->layers[l];
        NRC_REQUIRE(out > 0 && out <= 256, NRC_E_LIMIT, "layer width %d outside [1, 256]", out);
        const int stride = 256 / out;
        NRC_REQUIRE((in + stride - 1) / stride <= kNcfEntries, NRC_E_LIMIT,
                    "dense layer %d (%d -> %d) is too large for this build (in*out <= ~4096)", l, in, out);
        S.in_dim[l] = in; S.out_dim[l] = out;
        S.w_off[l] = off; off += in * out; S.b_off[l] = off; off += out;
        S.sw_off[l] = soff; soff += in * (out + 1); S.sb_off[l] = soff; soff += out;
        S.a_off[l + 1] = aoff; aoff += out;
        in = out;
    }
    S.tower_size = off; S.s_tower_size = soff; S.act_size = aoff;
    return NRC_OK;
}

static size_t ncf_smem_bytes(const NcfDev& S, int passes) {
    // weights (all towers) + per-warp, per-pass activations and deltas
    return ((size_t)S.n_towers * S.s_tower_size + (size_t)kNcfWarps * passes * 2 * S.act_size) * 4;
}

// One tower forward for the warp's sample.  act: this pass' activation buffer (a_0 filled).
__device__ __forceinline__ float ncf_tower_forward(const NcfDev& S, const float* sW, float* act,
                                                   int lane) {
    for (int l = 0; l < S.n_layers; ++l) {
        const int in = S.in_dim[l], out = S.out_dim[l];
        const float* W = sW + S.sw_off[l];
        const float* B = sW + S.sb_off[l];
        const float* a = act + S.a_off[l];
        float* o = act + S.a_off[l + 1];
        // lane owns outputs lane, lane+32, ... (<= 8): one broadcast read of a[k] feeds up to 8
        // independent FMA chains
        float acc[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) acc[r] = (lane + 32 * r < out) ? B[lane + 32 * r] : 0.0f;
        const int nr = (out + 31) >> 5;
        if (nr == 1) {
            const bool ok = lane < out;
            const float* w = W + (ok ? lane : 0);
#pragma unroll 8
            for (int k = 0; k < in; ++k) acc[0] = fmaf(a[k], w[k * (out + 1)], acc[0]);
        } else if (nr == 2) {
            const bool ok1 = lane + 32 < out;
            const float* w0 = W + lane;
            const float* w1 = W + (ok1 ? lane + 32 : lane);
#pragma unroll 8
            for (int k = 0; k < in; ++k) {
                const float ak = a[k];
                acc[0] = fmaf(ak, w0[k * (out + 1)], acc[0]);
                acc[1] = fmaf(ak, w1[k * (out + 1)], acc[1]);
            }
        } else {
            for (int k = 0; k < in; ++k) {
                const float ak = a[k];
#pragma unroll
                for (int r = 0; r < 8; ++r)
                    if (lane + 32 * r < out) acc[r] = fmaf(ak, W[k * (out + 1) + lane + 32 * r], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
            if (lane + 32 * r < out) o[lane + 32 * r] = fmaxf(acc[r], 0.0f);  // tf.nn.relu
        __syncwarp();
    }
    float s = 0.0f;
    if (S.n_layers > 0) {
        const float* o = act + S.a_off[S.n_layers];
        for (int j = lane; j < S.out_dim[S.n_layers - 1]; j += kWarp) s += o[j];
    }
    return warp_sum(s);
}

// Backward through one tower: fills delta buffers (same layout as act; delta of a_0 = gradient
// w.r.t. the concatenated MLP embeddings).  g = dLoss/dPrediction of this pass.
__device__ __forceinline__ void ncf_tower_backward(const NcfDev& S, const float* sW,
                                                   const float* act, float* del, float g, int lane) {
    if (S.n_layers == 0) return;
    {
        const int L = S.n_layers;
        const float* o = act + S.a_off[L];
        float* d = del + S.a_off[L];
        for (int j = lane; j < S.out_dim[L - 1]; j += kWarp) d[j] = (o[j] > 0.0f) ? g : 0.0f;
        __syncwarp();
    }
    for (int l = S.n_layers - 1; l >= 0; --l) {
        const int in = S.in_dim[l], out = S.out_dim[l];
        const float* W = sW + S.sw_off[l];
        const float* d = del + S.a_off[l + 1];
        const float* a = act + S.a_off[l];
        float* dp = del + S.a_off[l];
        // lane owns input rows lane, lane+32, ... : one broadcast read of d[j] feeds them all
        const int nr = (in + 31) >> 5;
        if (nr <= 2) {
            const bool ok1 = lane + 32 < in;
            const float* w0 = W + (size_t)((lane < in) ? lane : 0) * (out + 1);
            const float* w1 = W + (size_t)(ok1 ? lane + 32 : 0) * (out + 1);
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll 8
            for (int j = 0; j < out; ++j) {
                const float dj = d[j];
                s0 = fmaf(w0[j], dj, s0);
                s1 = fmaf(w1[j], dj, s1);
            }
            if (lane < in) dp[lane] = (l > 0) ? ((a[lane] > 0.0f) ? s0 : 0.0f) : s0;
            if (ok1) dp[lane + 32] = (l > 0) ? ((a[lane + 32] > 0.0f) ? s1 : 0.0f) : s1;
        } else {
            for (int k = lane; k < in; k += kWarp) {
                float s = 0.0f;
                for (int j = 0; j < out; ++j) s = fmaf(W[k * (out + 1) + j], d[j], s);
                dp[k] = (l > 0) ? ((a[k] > 0.0f) ? s : 0.0f) : s;
            }
        }
        __syncwarp();
    }
}

// kind: 0 = pointwise (labels), 1 = pairwise.
template <int NT>
__global__ void __launch_bounds__(kNcfWarps * 32)
ncf_grad_kernel(const NcfDev S, const NcfPtrs P, const int32_t* __restrict__ users,
                const int32_t* __restrict__ items, const void* __restrict__ third, int64_t batch,
                int pairwise, int loss_kind, float reg_mf, float reg_mlp, int32_t stamp,
                float* __restrict__ loss) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    const int passes = pairwise ? 2 : 1;
    float* sW = sm;                                           // [n_towers][s_tower_size]
    float* sAct = sW + S.n_towers * S.s_tower_size;           // [warp][pass][act_size]
    float* sDel = sAct + kNcfWarps * passes * S.act_size;     // [warp][pass][act_size]

    // stage dense weights with the padded row stride
    for (int t = 0; t < S.n_towers; ++t)
        for (int l = 0; l < S.n_layers; ++l) {
            const int in = S.in_dim[l], out = S.out_dim[l];
            const float* gW = P.dense + (size_t)t * S.tower_size + S.w_off[l];
            float* dW = sW + t * S.s_tower_size + S.sw_off[l];
            // warp w copies rows w, w+8, ...: coalesced, no integer division, loads pipeline
#pragma unroll 4
            for (int k = warp; k < in; k += kNcfWarps)
                for (int j = lane; j < out; j += kWarp) dW[k * (out + 1) + j] = __ldg(gW + k * out + j);
            const float* gB = P.dense + (size_t)t * S.tower_size + S.b_off[l];
            float* dB = sW + t * S.s_tower_size + S.sb_off[l];
            for (int e = tid; e < out; e += blockDim.x) dB[e] = gB[e];
        }
    __syncthreads();

    float accW[NT][kNcfMaxLayers][kNcfEntries];
    float accB[NT][kNcfMaxLayers];
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int l = 0; l < kNcfMaxLayers; ++l) {
            accB[t][l] = 0.0f;
#pragma unroll
            for (int m = 0; m < kNcfEntries; ++m) accW[t][l][m] = 0.0f;
        }

    const float inv_b = 1.0f / (float)batch;
    float loss_acc = 0.0f;
    const int64_t n_groups = (batch + kNcfWarps - 1) / kNcfWarps;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int64_t b = grp * kNcfWarps + warp;
        const bool live = b < batch;
        float* act0 = sAct + (warp * passes) * S.act_size;
        float* del0 = sDel + (warp * passes) * S.act_size;
        if (live) {
            const int u = users[b];
            const int it[2] = {items[b], pairwise ? reinterpret_cast<const int32_t*>(third)[b] : 0};
            float yhat[2] = {0.0f, 0.0f};
            for (int p = 0; p < passes; ++p) {
                float* act = act0 + p * S.act_size;
                float mf = 0.0f;
                for (int k = lane; k < S.mf_dim; k += kWarp)
                    mf = fmaf(P.mf_user[(size_t)u * S.mf_dim + k], P.mf_item[(size_t)it[p] * S.mf_dim + k], mf);
                mf = warp_sum(mf);
                for (int k = lane; k < S.mlp_dim; k += kWarp) {
                    act[k] = P.mlp_user[(size_t)u * S.mlp_dim + k];
                    act[S.mlp_dim + k] = P.mlp_item[(size_t)it[p] * S.mlp_dim + k];
                }
                __syncwarp();
                const int tower = (p == 1 && S.n_towers == 2) ? 1 : 0;
                yhat[p] = mf + ncf_tower_forward(S, sW + tower * S.s_tower_size, act, lane);
            }
            float l, g;
            if (pairwise) {
                const float x = yhat[0] - yhat[1];  // NeuMF.py:92 result = output - output_neg
                if (loss_kind == NRC_LOSS_BPR) { l = (x >= 0.f) ? log1pf(expf(-x)) : (-x + log1pf(expf(x))); g = -1.0f / (1.0f + expf(x)); }
                else if (loss_kind == NRC_LOSS_HINGE) { const float t = x + 1.0f; l = fmaxf(t, 0.f); g = (t > 0.f) ? 1.f : 0.f; }
                else { const float t = 1.0f - x; l = t * t; g = -2.0f * t; }
            } else {
                const float x = yhat[0], z = reinterpret_cast<const float*>(third)[b];
                if (loss_kind == NRC_LOSS_CROSS_ENTROPY) {
                    const float e = expf(-fabsf(x));
                    l = (fmaxf(x, 0.f) - x * z + log1pf(e)) * inv_b;
                    const float s = (x >= 0.f) ? 1.f / (1.f + e) : e / (1.f + e);
                    g = (s - z) * inv_b;
                } else { const float t = z - x; l = t * t; g = -2.0f * t; }
            }
            // regularisers (NeuMF.py:94-100): reg_mf*l2(p1,q2,q1) + reg_mlp*l2(m1,n2,n1)
            float sq_mf = 0.f, sq_mlp = 0.f;
            for (int p = 0; p < passes; ++p) {
                const float gp = (p == 0) ? g : -g;
                const int tower = (p == 1 && S.n_towers == 2) ? 1 : 0;
                float* act = act0 + p * S.act_size;
                float* del = del0 + p * S.act_size;
                ncf_tower_backward(S, sW + tower * S.s_tower_size, act, del, gp, lane);
                for (int k = lane; k < S.mf_dim; k += kWarp) {
                    const float pu = P.mf_user[(size_t)u * S.mf_dim + k];
                    const float qi = P.mf_item[(size_t)it[p] * S.mf_dim + k];
                    atomicAdd(P.g_mf_user + (size_t)u * S.mf_dim + k, gp * qi + (p == 0 ? reg_mf * pu : 0.f));
                    atomicAdd(P.g_mf_item + (size_t)it[p] * S.mf_dim + k, gp * pu + reg_mf * qi);
                    sq_mf += qi * qi + (p == 0 ? pu * pu : 0.f);
                }
                for (int k = lane; k < S.mlp_dim; k += kWarp) {
                    const float mu = act[k], mi = act[S.mlp_dim + k];
                    atomicAdd(P.g_mlp_user + (size_t)u * S.mlp_dim + k, del[k] + (p == 0 ? reg_mlp * mu : 0.f));
                    atomicAdd(P.g_mlp_item + (size_t)it[p] * S.mlp_dim + k, del[S.mlp_dim + k] + reg_mlp * mi);
                    sq_mlp += mi * mi + (p == 0 ? mu * mu : 0.f);
                }
                if (lane == 0) P.t_item[it[p]] = stamp;
            }
            if (lane == 0) P.t_user[u] = stamp;
            if (reg_mf != 0.f) l += reg_mf * 0.5f * warp_sum(sq_mf);
            if (reg_mlp != 0.f) l += reg_mlp * 0.5f * warp_sum(sq_mlp);
            loss_acc += l;
        } else {
            // dead warp: zero its buffers so the group reduction below adds nothing
            for (int e = lane; e < passes * S.act_size; e += kWarp) { del0[e] = 0.0f; act0[e] = 0.0f; }
        }
        __syncthreads();
        // weight / bias gradients of this group, accumulated in registers
#pragma unroll
        for (int l = 0; l < kNcfMaxLayers; ++l) {
            if (l < S.n_layers) {
                const int in = S.in_dim[l], out = S.out_dim[l];
                const int stride = 256 / out;
                const int j = tid % out, k0 = tid / out;
                if (k0 < stride) {
                    for (int w = 0; w < kNcfWarps; ++w)
                        for (int p = 0; p < passes; ++p) {
                            const int tower = (NT == 2 && p == 1) ? 1 : 0;
                            const float* a = sAct + (w * passes + p) * S.act_size + S.a_off[l];
                            const float dj = sDel[(w * passes + p) * S.act_size + S.a_off[l + 1] + j];
#pragma unroll
                            for (int m = 0; m < kNcfEntries; ++m) {
                                const int k = k0 + m * stride;
                                if (k < in) {
                                    if (tower == 0) accW[0][l][m] = fmaf(a[k], dj, accW[0][l][m]);
                                    else accW[NT - 1][l][m] = fmaf(a[k], dj, accW[NT - 1][l][m]);
                                }
                            }
                            if (k0 == 0) {
                                if (tower == 0) accB[0][l] += dj; else accB[NT - 1][l] += dj;
                            }
                        }
                }
            }
        }
        __syncthreads();
    }
    // flush the register accumulators: one RED per owned entry
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int l = 0; l < kNcfMaxLayers; ++l) {
            if (l < S.n_layers) {
                const int in = S.in_dim[l], out = S.out_dim[l];
                const int stride = 256 / out;
                const int j = tid % out, k0 = tid / out;
                if (k0 < stride) {
                    float* gW = P.g_dense + (size_t)t * S.tower_size + S.w_off[l];
#pragma unroll
                    for (int m = 0; m < kNcfEntries; ++m) {
                        const int k = k0 + m * stride;
                        if (k < in && accW[t][l][m] != 0.0f) atomicAdd(gW + k * out + j, accW[t][l][m]);
                    }
                    if (k0 == 0 && accB[t][l] != 0.0f)
                        atomicAdd(P.g_dense + (size_t)t * S.tower_size + S.b_off[l] + j, accB[t][l]);
                }
            }
        }
    if (lane == 0 && loss && loss_acc != 0.0f) atomicAdd(loss, loss_acc);
}

// predict for ALL items (NeuMF.py:163-168 / MLP.py:136-140): scores[b, i] = tower-0 forward.
// One warp per (user, item); CTA = 8 warps over 8 consecutive items of one user.
__global__ void __launch_bounds__(kNcfWarps * 32)
ncf_scores_kernel(const NcfDev S, const NcfPtrs P, const int32_t* __restrict__ users, int n_users,
                  int num_items, float* __restrict__ scores) {
    extern __shared__ __align__(16) float sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    float* sW = sm;
    float* sAct = sW + S.s_tower_size;
    for (int l = 0; l < S.n_layers; ++l) {
        const int in = S.in_dim[l], out = S.out_dim[l];
#pragma unroll 4
        for (int k = warp; k < in; k += kNcfWarps)
            for (int j = lane; j < out; j += kWarp)
                sW[S.sw_off[l] + k * (out + 1) + j] = __ldg(P.dense + S.w_off[l] + k * out + j);
        for (int e = tid; e < out; e += blockDim.x) sW[S.sb_off[l] + e] = P.dense[S.b_off[l] + e];
    }
    __syncthreads();
    float* act = sAct + warp * S.act_size;
    const int64_t total = (int64_t)n_users * num_items;
    for (int64_t e = (int64_t)blockIdx.x * kNcfWarps + warp; e < total; e += (int64_t)gridDim.x * kNcfWarps) {
        const int b = (int)(e / num_items), i = (int)(e - (int64_t)b * num_items);
        const int u = users[b];
        float mf = 0.0f;
        for (int k = lane; k < S.mf_dim; k += kWarp)
            mf = fmaf(P.mf_user[(size_t)u * S.mf_dim + k], P.mf_item[(size_t)i * S.mf_dim + k], mf);
        mf = warp_sum(mf);
        for (int k = lane; k < S.mlp_dim; k += kWarp) {
            act[k] = P.mlp_user[(size_t)u * S.mlp_dim + k];
            act[S.mlp_dim + k] = P.mlp_item[(size_t)i * S.mlp_dim + k];
        }
        __syncwarp();
        const float y = mf + ncf_tower_forward(S, sW, act, lane);
        if (lane == 0) scores[e] = y;
        __syncwarp();
    }
}

static int ncf_launch_grad(const nrc_ncf_shape* shape, const NcfPtrs& P, const int32_t* users,
                           const int32_t* items, const void* third, int64_t batch, int pairwise,
                           int loss_kind, float reg_mf, float reg_mlp, int32_t stamp, float* loss,
                           cudaStream_t st) {
    NcfDev S;
    int rc = ncf_make(S, shape);
    if (rc) return rc;
    if (pairwise)
        NRC_REQUIRE(loss_kind == NRC_LOSS_BPR || loss_kind == NRC_LOSS_HINGE || loss_kind == NRC_LOSS_SQUARE,
                    NRC_E_VALUE, "please choose a suitable loss function");
    else
        NRC_REQUIRE(loss_kind == NRC_LOSS_CROSS_ENTROPY || loss_kind == NRC_LOSS_SQUARE, NRC_E_VALUE,
                    "please choose a suitable loss function");
    if (batch <= 0) return NRC_OK;
    const int passes = pairwise ? 2 : 1;
    const size_t smem = ncf_smem_bytes(S, passes);
    NRC_REQUIRE(smem <= 200 * 1024, NRC_E_LIMIT, "NCF tower needs %zu B of shared memory", smem);
    int64_t groups = (batch + kNcfWarps - 1) / kNcfWarps;
    int64_t cap = (int64_t)sm_count() * 2;
    const int grid = (int)(groups < cap ? groups : cap);
    static bool attr_done = false;  // not inside a stream capture: set once, on first use
    if (!attr_done) {
        NRC_CUDA_CHECK(cudaFuncSetAttribute(ncf_grad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        NRC_CUDA_CHECK(cudaFuncSetAttribute(ncf_grad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done = true;
    }
    if (S.n_towers == 2 && pairwise) {
        ncf_grad_kernel<2><<<grid, kNcfWarps * 32, smem, st>>>(S, P, users, items, third, batch, pairwise,
                                                              loss_kind, reg_mf, reg_mlp, stamp, loss);
    } else {
        ncf_grad_kernel<1><<<grid, kNcfWarps * 32, smem, st>>>(S, P, users, items, third, batch, pairwise,
                                                              loss_kind, reg_mf, reg_mlp, stamp, loss);
    }
    NRC_CUDA_CHECK(cudaGetLastError());
    return NRC_OK;
}

}  // namespace nrc

using namespace nrc;

extern "C" int nrc_ncf_dense_size(const nrc_ncf_shape* shape) {
    NcfDev S;
    if (ncf_make(S, shape)) return NRC_E_VALUE;
    return S.tower_size * S.n_towers;
}

extern "C" int nrc_ncf_grad(const nrc_ncf_shape* shape, const float* mf_user, const float* mf_item,
                            const float* mlp_user, const float* mlp_item, const float* dense,
                            const int32_t* users, const int32_t* items, const void* third,
                            int64_t batch, int32_t pairwise, int32_t loss_kind, float reg_mf,
                            float reg_mlp, float* g_mf_user, float* g_mf_item, float* g_mlp_user,
                            float* g_mlp_item, float* g_dense, int32_t* touched_user,
                            int32_t* touched_item, int32_t stamp, float* loss, void* stream) {
    NcfPtrs P{mf_user, mf_item, mlp_user, mlp_item, dense, g_mf_user, g_mf_item, g_mlp_user,
              g_mlp_item, g_dense, touched_user, touched_item};
    return ncf_launch_grad(shape, P, users, items, third, batch, pairwise, loss_kind, reg_mf,
                           reg_mlp, stamp, loss, as_stream(stream));
}

extern "C" int nrc_ncf_scores(const nrc_ncf_shape* shape, const float* mf_user, const float* mf_item,
                              const float* mlp_user, const float* mlp_item, const float* dense,
                              const int32_t* users, int32_t n_users, int32_t num_items,
                              float* scores, void* stream) {
    NcfDev S;
    int rc = ncf_make(S, shape);
    if (rc) return rc;
    if (n_users <= 0) return NRC_OK;
    S.n_towers = 1;
    NcfPtrs P{mf_user, mf_item, mlp_user, mlp_item, dense, nullptr, nullptr, nullptr, nullptr,
              nullptr, nullptr, nullptr};
    const size_t smem = ((size_t)S.s_tower_size + (size_t)kNcfWarps * S.act_size) * 4;
    NRC_REQUIRE(smem <= 200 * 1024, NRC_E_LIMIT, "NCF tower needs %zu B of shared memory", smem);
    NRC_CUDA_CHECK(cudaFuncSetAttribute(ncf_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int64_t total = (int64_t)n_users * num_items;
    int64_t blocks = (total + kNcfWarps - 1) / kNcfWarps;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    ncf_scores_kernel<<<(unsigned)blocks, kNcfWarps * 32, smem, as_stream(stream)>>>(S, P, users, n_users,
                                                                                    num_items, scores);
    NRC_CUDA_CHECK(cudaGetLastError());
    return NRC_OK;
}

extern "C" int nrc_ncf_train_epoch(const nrc_ncf_shape* shape, float* mf_user, float* mf_item,
                                   float* mlp_user, float* mlp_item, float* dense,
                                   const int32_t* users, const int32_t* items, const void* third,
                                   int64_t n, int32_t batch_size, int32_t pairwise, int32_t loss_kind,
                                   float reg_mf, float reg_mlp, int32_t opt_kind,
                                   const float* lr_t_host, const float* hyper_host,
                                   float* const* grads, float* const* slot0, float* const* slot1,
                                   int32_t* touched_user, int32_t* touched_item, int32_t first_stamp,
                                   float* step_loss, void* stream) {
    NRC_REQUIRE(batch_size > 0, NRC_E_VALUE, "batch_size should be a positive integeral value");
    NcfDev S;
    int rc = ncf_make(S, shape);
    if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int64_t steps = (n + batch_size - 1) / batch_size;
    if (steps == 0) return NRC_OK;
    NRC_CUDA_CHECK(cudaMemsetAsync(step_loss, 0, (size_t)steps * sizeof(float), st));
    float hyper[4] = {hyper_host[0], hyper_host[1], hyper_host[2], hyper_host[3]};
    float* vars[5] = {mf_user, mf_item, mlp_user, mlp_item, dense};
    const int64_t rows[5] = {shape->num_users, shape->num_items, shape->num_users, shape->num_items, 1};
    const int dims[5] = {S.mf_dim, S.mf_dim, S.mlp_dim, S.mlp_dim, S.tower_size * S.n_towers};
    const int32_t* tch[5] = {touched_user, touched_item, touched_user, touched_item, nullptr};
    NcfPtrs P{mf_user, mf_item, mlp_user, mlp_item, dense, grads[0], grads[1], grads[2], grads[3],
              grads[4], touched_user, touched_item};
    for (int64_t s = 0; s < steps; ++s) {
        const int64_t off = s * batch_size;
        const int64_t bs = (n - off < batch_size) ? (n - off) : batch_size;
        const int32_t stamp = first_stamp + (int32_t)s;
        const void* th = pairwise ? (const void*)(reinterpret_cast<const int32_t*>(third) + off)
                                  : (const void*)(reinterpret_cast<const float*>(third) + off);
        rc = ncf_launch_grad(shape, P, users + off, items + off, th, bs, pairwise, loss_kind, reg_mf,
                             reg_mlp, stamp, step_loss + s, st);
        if (rc) return rc;
        if (opt_kind == NRC_OPT_ADAM) hyper[0] = lr_t_host[s];
        OptLaunch L;
        rc = opt_launch_init(L, opt_kind, hyper);
        if (rc) return rc;
        for (int v = 0; v < 5; ++v) {
            if (dims[v] == 0) continue;
            rc = opt_launch_add(L, vars[v], grads[v], slot0 ? slot0[v] : nullptr,
                                slot1 ? slot1[v] : nullptr, tch[v], rows[v], dims[v], v == 4);
            if (rc) return rc;
        }
        rc = opt_launch_run(L, stamp, st);
        if (rc) return rc;
    }
    return NRC_OK;
}

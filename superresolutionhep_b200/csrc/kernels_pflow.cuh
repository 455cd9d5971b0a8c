// fp32 CUDA-core kernels of the particle-flow forward (SAPF, h = 64, 4 heads of 16): the pieces
// that the generic fp32 kernels of kernels_f32.cuh (GEMM with fused epilogue, LayerNorm + adaLN
// modulate, varlen self-attention) do not cover.  The path is HBM/latency-bound (SURVEY 8d:
// ~0.2 MFLOP per cell), so everything stays fp32 and the cardinality argmax is exact.
#pragma once
#include "../../include/pflow.h"
#include "common.cuh"

namespace srhep {

constexpr int kPfH = 64;          // h_dim
constexpr int kPfMaxP = 8;        // max_particles supported

// ------------------------------------------------------------------------------------
// 1. Encoder cell initialisation (pflow/models/encoder.py:42-52):
//    cat[e, eta, cosphi, sinphi, Embedding(3, E)[layer]] -> Linear(4+E -> 64) -> LeakyReLU -> Linear(64 -> 64)
// ------------------------------------------------------------------------------------
struct PfCellInitParams {
    PflowCells in; int M; int emb_dim;
    const float* table;                 // [3, emb_dim]
    const float* w0; const float* b0;   // [64, 4 + emb_dim]
    const float* w2; const float* b2;   // [64, 64]
    float* out;                         // [M, 64]
};

__global__ void __launch_bounds__(256) pf_cell_init_kernel(PfCellInitParams p) {
    constexpr int RB = 32;                                   // rows per iteration
    __shared__ float w2t[kPfH][kPfH + 1];                    // w2 transposed: [k][col]
    __shared__ float w0s[kPfH][12 + 1];
    __shared__ float hid[RB][kPfH];
    __shared__ float xin[RB][12];
    const int tid = threadIdx.x, col = tid & 63, rg = tid >> 6;   // 4 row groups x 64 columns
    const int din = 4 + p.emb_dim;
    for (int i = tid; i < kPfH * kPfH; i += 256) w2t[i & 63][i >> 6] = p.w2[i];          // w2[col][k] -> w2t[k][col]
    for (int i = tid; i < kPfH * din; i += 256) w0s[i / din][i % din] = p.w0[i];
    const float b0 = p.b0[col], b2 = p.b2[col];
    for (int r0 = blockIdx.x * RB; r0 < p.M; r0 += gridDim.x * RB) {
        __syncthreads();
        for (int i = tid; i < RB * din; i += 256) {
            const int r = i / din, k = i % din, row = r0 + r;
            float v = 0.f;
            if (row < p.M) {
                if (k == 0) v = p.in.e[row];
                else if (k == 1) v = p.in.eta[row];
                else if (k == 2) v = p.in.cosphi[row];
                else if (k == 3) v = p.in.sinphi[row];
                else v = p.table[p.in.layer[row] * p.emb_dim + (k - 4)];
            }
            xin[r][k] = v;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RB / 4; ++i) {
            const int r = rg * (RB / 4) + i;
            float a = b0;
            for (int k = 0; k < din; ++k) a = fmaf(w0s[col][k], xin[r][k], a);
            hid[r][col] = leaky_relu(a);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RB / 4; ++i) {
            const int r = rg * (RB / 4) + i, row = r0 + r;
            float a = b2;
#pragma unroll 16
            for (int k = 0; k < kPfH; ++k) a = fmaf(w2t[k][col], hid[r][k], a);
            if (row < p.M) p.out[(size_t)row * kPfH + col] = a;
        }
    }
}

// ------------------------------------------------------------------------------------
// 2. Masked mean over the cells of an event (encoder.py:54-55, cardinality_predictor.py:18-19,
//    kinematics_predictor.py:119-120) and SiLU of it (input of every adaLN Linear).
//    One block per event.  n = 0 gives 0/0 = NaN, as in the reference.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pf_event_mean_kernel(const float* __restrict__ x, const int* __restrict__ cu, float* mean, float* silu_mean) {
    __shared__ float part[4][kPfH];
    const int e = blockIdx.x, col = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int r0 = cu[e], r1 = cu[e + 1];
    float s = 0.f;
    for (int r = r0 + rg; r < r1; r += 4) s += x[(size_t)r * kPfH + col];
    part[rg][col] = s;
    __syncthreads();
    if (rg == 0) {
        const float m = (part[0][col] + part[1][col] + part[2][col] + part[3][col]) / (float)(r1 - r0);
        mean[(size_t)e * kPfH + col] = m;
        silu_mean[(size_t)e * kPfH + col] = silu(m);
    }
}

// ------------------------------------------------------------------------------------
// 3. Cardinality head (cardinality_predictor.py:17-22 + models/dense.py:49-83):
//    [LN -> Linear -> LeakyReLU] x n_hidden -> Linear -> logits; argmax; part_mask = arange(P) < n_pred
//    (model_pf.py:65-67).  One block of 128 threads per event.
// ------------------------------------------------------------------------------------
struct PfCardParams {
    const float* g;                         // [B, 64] masked mean of the encoded cells
    int n_hidden; int width[PFLOW_MAX_CARD_HIDDEN + 2];      // 64, hidden..., card_out
    const float* w[PFLOW_MAX_CARD_HIDDEN + 1]; const float* b[PFLOW_MAX_CARD_HIDDEN + 1];
    float* logits; int* n_pred;             // [B, card_out], [B]
    const uint8_t* part_mask_in;            // training mode: given mask (B, P); null: from the argmax
    uint8_t* part_mask; int P;              // [B, P]
};

__global__ void __launch_bounds__(128) pf_cardinality_kernel(PfCardParams p) {
    __shared__ float va[128], vb[128], red[8];
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* cur = va; float* nxt = vb;
    if (tid < p.width[0]) cur[tid] = p.g[(size_t)e * p.width[0] + tid];
    __syncthreads();
    for (int l = 0; l <= p.n_hidden; ++l) {
        const int win = p.width[l], wout = p.width[l + 1];
        const bool hidden = l < p.n_hidden;
        float mean = 0.f, rstd = 1.f;
        if (hidden) {                                        // non-affine LayerNorm in front of every hidden Linear
            float v = tid < win ? cur[tid] : 0.f;
            float s = warp_sum(v);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            mean = (red[0] + red[1] + red[2] + red[3]) / (float)win;
            __syncthreads();
            const float d = tid < win ? v - mean : 0.f;
            s = warp_sum(d * d);
            if (lane == 0) red[warp] = s;
            __syncthreads();
            rstd = 1.0f / sqrtf((red[0] + red[1] + red[2] + red[3]) / (float)win + kLnEps);
            if (tid < win) cur[tid] = (v - mean) * rstd;
            __syncthreads();
        }
        if (tid < wout) {
            const float* wr = p.w[l] + (size_t)tid * win;
            float a = p.b[l][tid];
            for (int k = 0; k < win; ++k) a = fmaf(__ldg(wr + k), cur[k], a);
            nxt[tid] = hidden ? leaky_relu(a) : a;
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
    }
    const int nout = p.width[p.n_hidden + 1];
    if (tid < nout) p.logits[(size_t)e * nout + tid] = cur[tid];
    if (tid == 0) {
        int best = 0; float bv = cur[0];
        for (int i = 1; i < nout; ++i) if (cur[i] > bv) { bv = cur[i]; best = i; }      // first maximum, like torch.argmax
        if (p.n_pred) p.n_pred[e] = best;
        for (int j = 0; j < p.P; ++j) p.part_mask[(size_t)e * p.P + j] = p.part_mask_in ? p.part_mask_in[(size_t)e * p.P + j] : (uint8_t)(j < best);
    }
}

// rows [B * P, 64] <- the P initial particle embeddings (kinematics_predictor.py:84-91), same for every event
__global__ void pf_bcast_particles_kernel(const float* __restrict__ pe, float* out, int n_rows, int P) {
    const size_t n = (size_t)n_rows * kPfH;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = pe[((i / kPfH) % P) * kPfH + (i % kPfH)];
}

// ------------------------------------------------------------------------------------
// 4. Particle -> cell cross-attention of the decoder (models/attention.py:135-221 with
//    q = particles, k = v = modulated LN(cells); masks pad_q | pad_k).  One block per event,
//    one warp per (particle, head); lanes stride over the event's cells with a private online
//    softmax that is merged across the warp at the end.  A masked particle (p >= n_pred) gets a
//    zero row (fully masked softmax -> masked_fill(0)), so that linear_out returns its bias.
// ------------------------------------------------------------------------------------
struct PfCrossParams {
    const float* q;      // [B * P, 64] projected queries
    const float* kv;     // [T, 128] projected keys | values
    const int* cu; const uint8_t* part_mask; int P;
    float* out;          // [B * P, 64]
};

template <int HD>
__global__ void __launch_bounds__(512) pf_cross_attn_kernel(PfCrossParams p) {
    const int e = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int heads = kPfH / HD;
    const int r0 = p.cu[e], r1 = p.cu[e + 1];
    for (int pair = warp; pair < p.P * heads; pair += blockDim.x >> 5) {
        const int pi = pair / heads, head = pair % heads;
        float* op = p.out + ((size_t)e * p.P + pi) * kPfH + head * HD;
        if (!p.part_mask[(size_t)e * p.P + pi]) {
            if (lane < HD) op[lane] = 0.f;
            continue;
        }
        float q[HD], acc[HD];
        const float inv = 1.0f / sqrtf((float)HD);
#pragma unroll
        for (int d = 0; d < HD; ++d) { q[d] = p.q[((size_t)e * p.P + pi) * kPfH + head * HD + d] * inv; acc[d] = 0.f; }
        float m = -INFINITY, l = 0.f;
        for (int r = r0 + lane; r < r1; r += 32) {
            const float* kr = p.kv + (size_t)r * (2 * kPfH) + head * HD;
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
                s = fmaf(q[d], k4.x, s); s = fmaf(q[d + 1], k4.y, s); s = fmaf(q[d + 2], k4.z, s); s = fmaf(q[d + 3], k4.w, s);
            }
            const float mn = fmaxf(m, s);
            const float corr = expf(m - mn), pw = expf(s - mn);
            l = l * corr + pw;
#pragma unroll
            for (int d = 0; d < HD; d += 4) {
                const float4 v4 = *reinterpret_cast<const float4*>(kr + kPfH + d);
                acc[d] = fmaf(pw, v4.x, acc[d] * corr); acc[d + 1] = fmaf(pw, v4.y, acc[d + 1] * corr);
                acc[d + 2] = fmaf(pw, v4.z, acc[d + 2] * corr); acc[d + 3] = fmaf(pw, v4.w, acc[d + 3] * corr);
            }
            m = mn;
        }
        const float mall = warp_max(m);
        const float sc = m == -INFINITY ? 0.f : expf(m - mall);
        l = warp_sum(l * sc);
        const float il = l > 0.f ? 1.f / l : 0.f;
#pragma unroll
        for (int d = 0; d < HD; ++d) {
            const float a = warp_sum(acc[d] * sc);
            if (lane == d) op[d] = a * il;
        }
    }
}

// ------------------------------------------------------------------------------------
// 5. AttnKinematicNet (kinematics_predictor.py:24-57): scores = (W_q p)(W_k c)^T / sqrt(64), softmax
//    over the VALID PARTICLES of each cell, energy-weighted eta / phi / E per particle, pT = E / cosh(eta),
//    VarTransformation.forward.  One block per event; inc weights written particle-major (P, T).
// ------------------------------------------------------------------------------------
struct PfKinParams {
    const float* qp;     // [B * P, 64]
    const float* kp;     // [T, 64]
    const int* cu; const uint8_t* part_mask; int P; int T;
    const float* e_raw; const float* eta_raw; const float* phi;
    PflowVarTransform tr[3];      // pt, eta, e
    float* inc;          // [P, T]
    float* kin;          // [B, P, 4]
};

__device__ __forceinline__ float pf_var_forward(const PflowVarTransform& t, float x) {
    if (t.trans == PFLOW_TRANS_POW) x = powf(x, t.m);
    else if (t.trans == PFLOW_TRANS_POW_SIGNED) x = (x >= 0.f ? 1.f : -1.f) * powf(fabsf(x), t.m);
    if (t.scale == PFLOW_SCALE_MINMAX) x = (x - t.min) / (t.max - t.min) * (t.hi - t.lo) + t.lo;
    else if (t.scale == PFLOW_SCALE_STANDARD) x = (x - t.mean) / t.std;
    return x;
}

__global__ void __launch_bounds__(256) pf_kin_kernel(PfKinParams p) {
    __shared__ float qs[kPfMaxP][kPfH];
    __shared__ float red[8][3 * kPfMaxP];
    const int e = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = p.cu[e], r1 = p.cu[e + 1];
    for (int i = tid; i < p.P * kPfH; i += 256) qs[i / kPfH][i % kPfH] = p.qp[(size_t)e * p.P * kPfH + i];
    __syncthreads();
    bool pm[kPfMaxP];
#pragma unroll
    for (int j = 0; j < kPfMaxP; ++j) pm[j] = j < p.P && p.part_mask[(size_t)e * p.P + j];
    float sa[kPfMaxP], sb[kPfMaxP], sc[kPfMaxP];
#pragma unroll
    for (int j = 0; j < kPfMaxP; ++j) { sa[j] = 0.f; sb[j] = 0.f; sc[j] = 0.f; }
    const float inv = 1.0f / sqrtf((float)kPfH);
    for (int r = r0 + tid; r < r1; r += 256) {
        float s[kPfMaxP];
#pragma unroll
        for (int j = 0; j < kPfMaxP; ++j) s[j] = 0.f;
        const float* kr = p.kp + (size_t)r * kPfH;
#pragma unroll 4
        for (int d = 0; d < kPfH; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
            for (int j = 0; j < kPfMaxP; ++j)
                if (j < p.P) { s[j] = fmaf(qs[j][d], k4.x, s[j]); s[j] = fmaf(qs[j][d + 1], k4.y, s[j]); s[j] = fmaf(qs[j][d + 2], k4.z, s[j]); s[j] = fmaf(qs[j][d + 3], k4.w, s[j]); }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < kPfMaxP; ++j) if (pm[j]) { s[j] *= inv; mx = fmaxf(mx, s[j]); }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < kPfMaxP; ++j) { s[j] = pm[j] ? expf(s[j] - mx) : 0.f; den += s[j]; }
        const float iden = den > 0.f ? 1.f / den : 0.f;      // no valid particle: all-masked softmax -> 0
        const float er = p.e_raw[r], et = p.eta_raw[r], ph = p.phi[r];
#pragma unroll
        for (int j = 0; j < kPfMaxP; ++j)
            if (j < p.P) {
                const float w = s[j] * iden;
                p.inc[(size_t)j * p.T + r] = w;
                const float ei = w * er;
                sa[j] += ei; sb[j] = fmaf(ei, et, sb[j]); sc[j] = fmaf(ei, ph, sc[j]);
            }
    }
#pragma unroll
    for (int j = 0; j < kPfMaxP; ++j) {
        const float a = warp_sum(sa[j]), b = warp_sum(sb[j]), c = warp_sum(sc[j]);
        if (lane == 0) { red[warp][3 * j] = a; red[warp][3 * j + 1] = b; red[warp][3 * j + 2] = c; }
    }
    __syncthreads();
    if (tid < p.P) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[w][3 * tid]; b += red[w][3 * tid + 1]; c += red[w][3 * tid + 2]; }
        const float den = a + (a == 0.f ? 1.f : 0.f);        // e_raw_inc.sum == 0 -> + 1 (kinematics_predictor.py:41-42)
        const float eta = b / den, phi = c / den;
        const float pt = a / coshf(eta);
        float* o = p.kin + ((size_t)e * p.P + tid) * 4;
        o[0] = pf_var_forward(p.tr[0], pt); o[1] = pf_var_forward(p.tr[1], eta); o[2] = phi; o[3] = pf_var_forward(p.tr[2], a);
    }
}

}  // namespace srhep

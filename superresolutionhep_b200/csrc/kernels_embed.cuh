// Per-cell embedding on tensor cores (models/flow_model.py:192-215, models/dense.py:49-83): for every 128-row tile
//
//   hidden  = LeakyReLU(Linear1(LN(cat[x_cell, time_emb])))   three nets (etaphi 3 in, e_proxy 1 in, noisy input 1 in), 64 wide each,
//             evaluated per cell from per-event pieces (W_c . time_emb, its mean / variance: event_prep_kernel) and 1-3 MACs per unit,
//             written straight into shared memory as the 16-bit A operand (one 64-column k-block per net)
//   out     = LeakyReLU(hidden . W2^T + b2)                    block-diagonal: three tcgen05 MMAs groups (N = 32, 32, 64) into 128 TMEM columns
//   tok     = [etaphi(32) | layer_out[event][layer](32) | e_proxy_emb(31) | e_proxy(1) | noisy(64)]   -> fp32 cond part + 16-bit feat_0 operand
//
// The second Linear is 85 % of the embedding FLOPs and ran at a few per cent of the FMA peak as a CUDA-core loop; here it costs ~300 tensor
// cycles per tile.  Operands are fp16 whatever the transformer's operand format (hidden units sit behind a LayerNorm, weights are small).
//   warp 0: weight load + MMA issuer + TMEM allocator   warps 1-4: one thread = one cell (TMEM lane = row)
#pragma once
#include "kernels_chain.cuh"

namespace srhep {

constexpr int kEmbThreads = 160;
constexpr int kEmbEv = 8;                           // events per tile whose per-event pieces are staged in shared memory
constexpr int kEmbEvFloats = 292;                   // ev_a (3 x 64) | ev_stats (2) | pad (2) | layer_out (3 x 32)
constexpr uint32_t kEmbOffW = 49152;                // after the 3 k-blocks of A
constexpr uint32_t kEmbOffEv = kEmbOffW + 16384;
constexpr uint32_t kEmbOffBars = kEmbOffEv + kEmbEv * kEmbEvFloats * 4;
constexpr size_t kEmbSmemBytes = ((kEmbOffBars + 64 + 1023) / 1024) * 1024;

struct EmbedTcParams {
    int M; int row0; int lp_fp16;
    const float* eta; const float* cosphi; const float* sinphi; const float* e_proxy; const int* layer;   // global rows
    StageRef stage;                                   // x_in: pass-local rows
    const int* row_event;                             // pass-local rows -> global event
    const float* ev_a; const float* ev_stats; const float* layer_out;     // per event: [3][64], [2], [3][32]
    const uint8_t* w_img;                             // [128 rows x 128 B] fp16, 128B-swizzled: row g = second-layer weights of GEMM column g over ITS net's 64 hidden units
    float r1[192], b1[192], w0[192], w1[64], w2[64];  // first Linear: time-embedding row sums, bias, weights of the per-cell inputs (w1, w2: etaphi only)
    float b2[128];                                    // second Linear bias in GEMM column order (column 63 is padding)
    float te;                                         // t_emb as float
    float* tok_feat; int ld;                          // [M, 160] fp32 (only the 96 cond columns are written: the head reads them)
    void* tok_lp; int ld_lp;                          // [M, 192] 16-bit feat_0 operand (columns 160.. stay zero)
};

__device__ __forceinline__ void emb_store_lp(void* base, size_t off, const float (&v)[32], int fp16) {
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) pk[i] = pack16(v[2 * i], v[2 * i + 1], fp16);
    uint16_t* d = reinterpret_cast<uint16_t*>(base) + off;
    stg256(d, &pk[0]); stg256(d + 16, &pk[8]);
}

__global__ void __launch_bounds__(kEmbThreads, 3) embed_tc_kernel(const __grid_constant__ EmbedTcParams p) {
    extern __shared__ __align__(1024) uint8_t emb_smem[];
    uint8_t* smem = emb_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_w = smem + kEmbOffW;
    float* s_ev = reinterpret_cast<float*>(smem + kEmbOffEv);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kEmbOffBars);
    uint64_t* w_full = bars;        // weights landed
    uint64_t* a_ready = bars + 1;   // cell threads -> MMA: A operand written
    uint64_t* acc_full = bars + 2;  // MMA -> cell threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + 127) / 128;
    constexpr uint32_t kTmemCols = 128;

    if (warp == 0 && lane == 0) {
        mbar_init(w_full, 1); mbar_init(a_ready, 4); mbar_init(acc_full, 1);
        mbar_fence_init();
        mbar_expect_tx(w_full, 16384);
        bulk_load(s_w, p.w_img, 16384, w_full);
    }
    if (warp == 0) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        const uint32_t id32 = umma_idesc_16(128, 32, 1), id64 = umma_idesc_16(128, 64, 1);
        mbar_wait(w_full, 0);
        uint32_t it = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
            mbar_wait(a_ready, it & 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a = smem_u32(smem), w = smem_u32(s_w);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, umma_desc_sw128(a + k * 32), umma_desc_sw128(w + k * 32), id32, (uint32_t)(k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 32, umma_desc_sw128(a + 16384 + k * 32), umma_desc_sw128(w + 4096 + k * 32), id32, (uint32_t)(k != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 64, umma_desc_sw128(a + 32768 + k * 32), umma_desc_sw128(w + 8192 + k * 32), id64, (uint32_t)(k != 0));
                tc_commit(acc_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int rt = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t a_sh = smem_u32(smem);
        const float* x_in = load_stage(p.stage).x_in;
        const int et = threadIdx.x - 32;                                   // 0..127 among the cell threads
        uint32_t it = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++it) {
            const int row = t * 128 + rt;
            const bool valid = row < p.M;
            const int ev0 = p.row_event[t * 128];
            const int ne = p.row_event[min(t * 128 + 127, p.M - 1)] - ev0 + 1;
            const bool staged = ne <= kEmbEv;
            const int evt = p.row_event[min(row, p.M - 1)];
            // ---- stage the per-event pieces of this tile (the previous tile's readers are behind the barrier at its end)
            if (staged) {
                for (int i = et; i < ne * kEmbEvFloats; i += 128) {
                    const int e = i / kEmbEvFloats, k = i % kEmbEvFloats;
                    const size_t ge = (size_t)(ev0 + e);
                    float v = 0.f;
                    if (k < 192) v = p.ev_a[ge * 192 + k];
                    else if (k < 194) v = p.ev_stats[ge * 2 + (k - 192)];
                    else if (k >= 196) v = p.layer_out[ge * 96 + (k - 196)];
                    s_ev[i] = v;
                }
            }
            const size_t grow = (size_t)p.row0 + min(row, p.M - 1);
            const float a0 = p.eta[grow], a1 = p.cosphi[grow], a2 = p.sinphi[grow], pr = p.e_proxy[grow], xt = x_in[min(row, p.M - 1)];
            const int lay = p.layer[grow];
            named_bar_sync(1, 128);
            const float* ea = staged ? s_ev + (evt - ev0) * kEmbEvFloats : p.ev_a + (size_t)evt * 192;
            const float* es = staged ? ea + 192 : p.ev_stats + (size_t)evt * 2;
            const float* lo = staged ? ea + 196 : p.layer_out + (size_t)evt * 96;
            const float mt = es[0], vt = es[1], te = p.te;
            // LayerNorm statistics of cat[x_cell, time_emb] from the per-event statistics of time_emb (kernels_f32.cuh: embed_tokens_kernel)
            float mu[3], rs[3];
            {
                const float nf = 3.f + te;
                mu[0] = (a0 + a1 + a2 + te * mt) / nf;
                const float d = mt - mu[0];
                rs[0] = 1.0f / sqrtf((vt + te * d * d + (a0 - mu[0]) * (a0 - mu[0]) + (a1 - mu[0]) * (a1 - mu[0]) + (a2 - mu[0]) * (a2 - mu[0])) / nf + kLnEps);
            }
#pragma unroll
            for (int n = 1; n < 3; ++n) {
                const float a = n == 1 ? pr : xt, nf = 1.f + te;
                mu[n] = (a + te * mt) / nf;
                const float d = mt - mu[n];
                rs[n] = 1.0f / sqrtf((vt + te * d * d + (a - mu[n]) * (a - mu[n])) / nf + kLnEps);
            }
            // ---- hidden units -> A operand (k-block n = net n)
#pragma unroll
            for (int n = 0; n < 3; ++n) {
                const float xin = n == 0 ? a0 - mu[0] : (n == 1 ? pr - mu[1] : xt - mu[2]);
                const float y1 = a1 - mu[0], y2 = a2 - mu[0];
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int k = n * 64 + half * 32 + j;
                        float acc = fmaf(-mu[n], p.r1[k], ea[k]);
                        acc = fmaf(p.w0[k], xin, acc);
                        if (n == 0) { acc = fmaf(p.w1[half * 32 + j], y1, acc); acc = fmaf(p.w2[half * 32 + j], y2, acc); }
                        v[j] = leaky_relu(fmaf(rs[n], acc, p.b1[k]));
                    }
                    chain_store_a(a_sh, rt, n * 64 + half * 32, v, 1);
                }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
            // ---- second Linear done on the tensor core: bias, LeakyReLU, assemble the token row, store
            mbar_wait(acc_full, it & 1);
            tc_fence_after();
            const size_t ro = (size_t)min(row, p.M - 1);
#pragma unroll 1
            for (int c = 0; c < 5; ++c) {                                  // token chunks of 32 columns: etaphi | layer | proxy + raw | noisy lo | noisy hi
                float v[32];
                if (c == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = lo[lay * 32 + j];
                } else {
                    const int g = c == 0 ? 0 : c - 1;                      // GEMM chunk
                    uint32_t r[32];
                    tmem_ld32(t_lane + g * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = leaky_relu(__uint_as_float(r[j]) + p.b2[g * 32 + j]);
                    if (c == 2) v[31] = pr;                                // GEMM column 63 is padding: the raw e_proxy sits there
                }
                if (valid) {
                    if (c < 3) {
                        float* d = p.tok_feat + ro * p.ld + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) stg256(d + j, reinterpret_cast<const uint32_t*>(&v[j]));
                    }
                    emb_store_lp(p.tok_lp, ro * p.ld_lp + c * 32, v, p.lp_fp16);
                }
            }
            tc_fence_before();
            named_bar_sync(1, 128);                                        // TMEM drained and the staged event rows no longer read: next tile may overwrite both
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// context = cat[time_emb, masked mean of cond_feat] (models/flow_model.py:210-222) and SiLU(context), reading the cond columns of tok_feat
struct ContextRowsParams {
    const float* temb; const float* tok_feat; int ld; int row0;
    const int* cu_seqlens;          // global
    float* ctx; float* silu_ctx;    // [B, t_emb + cond]
    int t_emb, cond, e0;
};
__global__ void __launch_bounds__(384) context_rows_kernel(ContextRowsParams p) {
    __shared__ float part[4][96];
    const int e = p.e0 + blockIdx.x;
    const int r0 = p.cu_seqlens[e] - p.row0, r1 = p.cu_seqlens[e + 1] - p.row0;
    const int col = threadIdx.x % 96, rg = threadIdx.x / 96;
    float s = 0.f;
    if (col < p.cond)
        for (int r = r0 + rg; r < r1; r += 4) s += p.tok_feat[(size_t)r * p.ld + col];
    part[rg][col] = s;
    __syncthreads();
    const int width = p.t_emb + p.cond;
    for (int c = threadIdx.x; c < width; c += blockDim.x) {
        float v;
        if (c < p.t_emb) v = p.temb[(size_t)e * p.t_emb + c];
        else { const int k = c - p.t_emb; v = r1 > r0 ? (part[0][k] + part[1][k] + part[2][k] + part[3][k]) / (float)(r1 - r0) : 0.f; }
        p.ctx[(size_t)e * width + c] = v;
        p.silu_ctx[(size_t)e * width + c] = silu(v);
    }
}

}  // namespace srhep

// Fused velocity head for sm_100a (models/flow_model.py:258 `v_t_pred_net`, models/dense.py:49-83, and
// the fixed-grid ODE update of torchdiffeq): for every 128-row tile, ONE kernel runs
//
//   G1  h1 = hin . W1^T            (K = 512, N = 128; hin = LN(cat[mod(norm_v_t(...)), ctx]) from head_prep, 16 bit)
//   E1  LeakyReLU(h1 + b1) -> LayerNorm(128)                    -> A2 (shared memory, 16 bit)
//   G2  h2 = A2 . W2^T             (K = 128, N = 64)
//   E2  LeakyReLU(h2 + b2) -> LayerNorm(64)                     -> A3 (shared memory, 16 bit)
//   G3  h3 = A3 . W3^T             (K = 64,  N = 32)
//   E3  LeakyReLU(h3 + b3) [-> LayerNorm(32)] -> v = w4 . h3 + b4 -> out = base + coef * v  (solution[j+1])
//
// all three GEMMs on tcgen05 with TMEM accumulators (128 + 128 + 64 + 32 columns: the first one is
// double-buffered so that G1 of the next tile and its TMA traffic run under the epilogues of this one).
// hin and W1 stream through a 3-slot TMA ring; W2 / W3 stay resident in shared memory; biases and the
// last Linear travel by value in the constant bank.  Per cell the kernel reads 1 KB and writes 4-8 B.
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: epilogue, one thread = one row
#pragma once
#include "kernels_chain.cuh"

namespace srhep {

constexpr int kHeadThreads = 192;
constexpr int kHeadSlots = 3;
constexpr uint32_t kHeadSlotBytes = 32768;           // hin k-block (128 rows x 128 B) | W1 k-block (128 rows x 128 B)
constexpr uint32_t kHeadOffA2 = kHeadSlots * kHeadSlotBytes;        // 2 k-blocks
constexpr uint32_t kHeadOffA3 = kHeadOffA2 + 32768;                 // 1 k-block
constexpr uint32_t kHeadOffW2 = kHeadOffA3 + 16384;                 // 2 k-blocks x (64 rows x 128 B)
constexpr uint32_t kHeadOffW3 = kHeadOffW2 + 16384;                 // 32 rows x 128 B
constexpr uint32_t kHeadOffBars = kHeadOffW3 + 4096;
constexpr size_t kHeadSmemBytes = kHeadOffBars + 256;
constexpr int kHeadK1 = 512, kHeadH1 = 128, kHeadH2 = 64, kHeadH3 = 32;
// split (fp32-grade) mode: only the first GEMM (K = 512, 90 % of the head's FLOPs) runs here, on (hi, lo) fp16 planes of hin and W1
// (slot = hin hi | W1 hi | hin lo | W1 lo); LeakyReLU(h1 + b1) leaves as fp32 rows for the CUDA-core tail (head_tail_kernel)
constexpr uint32_t kHeadSlotBytesSplit = 65536;
constexpr size_t kHeadSmemBytesSplit = kHeadSlots * kHeadSlotBytesSplit + 256;

struct HeadChainParams {
    int M; int fp16; int final_ln;
    const uint8_t* w1;           // [8 k-blocks][128 rows x 128 B] pre-swizzled
    const uint8_t* w2;           // [2 k-blocks][64 rows x 128 B]
    const uint8_t* w3;           // [32 rows x 128 B]
    float b1[kHeadH1], b2[kHeadH2], b3[kHeadH3], w4[kHeadH3]; float b4;
    StageRef stage;              // ODE stage: out / base / vout are pass-local rows
    const uint8_t* w1_lo;        // split mode: low plane of W1 (images hold W1 * 2^s), 2^-s, and the fp32 output rows [M, 128]
    float ws1; float* h1out;
    Extents ext;                 // checked in the -DSRHEP_BOUNDS build only
};

__device__ __forceinline__ void ln_inplace_128(float (&v)[128]) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 128; ++j) s += v[j];
    const float mean = s * (1.0f / 128.f);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 128; ++j) { const float d = v[j] - mean; q = fmaf(d, d, q); }
    const float rstd = rsqrtf(q * (1.0f / 128.f) + kLnEps);
#pragma unroll
    for (int j = 0; j < 128; ++j) v[j] = (v[j] - mean) * rstd;
}

template <bool kSplit = false>
__global__ void __launch_bounds__(kHeadThreads, 1) head_chain_kernel(const __grid_constant__ CUtensorMap tmap_hin, const __grid_constant__ CUtensorMap tmap_hin_lo,
                                                                     const __grid_constant__ HeadChainParams p) {
    constexpr uint32_t kSlot = kSplit ? kHeadSlotBytesSplit : kHeadSlotBytes;
    extern __shared__ __align__(1024) uint8_t head_smem[];
    uint8_t* smem = head_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_a2 = smem + kHeadOffA2;
    uint8_t* s_a3 = smem + kHeadOffA3;
    uint8_t* s_w2 = smem + kHeadOffW2;
    uint8_t* s_w3 = smem + kHeadOffW3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (kSplit ? kHeadSlots * kHeadSlotBytesSplit : kHeadOffBars));
    uint64_t* full = bars;               // [3] TMA -> MMA
    uint64_t* empty = bars + 3;          // [3] MMA -> TMA
    uint64_t* w23_full = bars + 6;
    uint64_t* acc1_full = bars + 7;      // [2] MMA -> epilogue
    uint64_t* acc1_empty = bars + 9;     // [2] epilogue -> MMA
    uint64_t* a2_ready = bars + 11;      // epilogue -> MMA
    uint64_t* acc2_full = bars + 12;
    uint64_t* a3_ready = bars + 13;
    uint64_t* acc3_full = bars + 14;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + 127) / 128;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kColAcc2 = 256, kColAcc3 = 320;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_hin);
        if (kSplit) prefetch_tmap(&tmap_hin_lo);
        for (int i = 0; i < kHeadSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(w23_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 4); }
        mbar_init(a2_ready, 4); mbar_init(acc2_full, 1); mbar_init(a3_ready, 4); mbar_init(acc3_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            if (!kSplit) {
            mbar_expect_tx(w23_full, 16384 + 4096);
            bulk_load(s_w2, p.w2, 16384, w23_full);
            bulk_load(s_w3, p.w3, 4096, w23_full);
            }
            uint32_t it = 0;
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) {
                for (int kb = 0; kb < kHeadK1 / 64; ++kb, ++it) {
                    const uint32_t s = it % kHeadSlots, ph = (it / kHeadSlots) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], kSlot);
                    tma_load_2d(smem + s * kSlot, &tmap_hin, &full[s], kb * 64, t * 128);
                    bulk_load(smem + s * kSlot + 16384, p.w1 + (size_t)kb * 16384, 16384, &full[s]);
                    if (kSplit) {
                        tma_load_2d(smem + s * kSlot + 32768, &tmap_hin_lo, &full[s], kb * 64, t * 128);
                        bulk_load(smem + s * kSlot + 49152, p.w1_lo + (size_t)kb * 16384, 16384, &full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc1 = umma_idesc_16(128, kHeadH1, p.fp16), idesc2 = umma_idesc_16(128, kHeadH2, p.fp16), idesc3 = umma_idesc_16(128, kHeadH3, p.fp16);
        uint32_t it = 0;
        auto issue_g1 = [&](uint32_t j) {                      // j = CTA-local tile counter
            const uint32_t buf = j & 1, use = j >> 1;
            mbar_wait(&acc1_empty[buf], (use & 1) ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < kHeadK1 / 64; ++kb, ++it) {
                const uint32_t s = it % kHeadSlots, ph = (it / kHeadSlots) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + s * kSlot), b_addr = a_addr + 16384;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        umma_bf16(tmem_base + buf * 128, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1, (uint32_t)((kb | k) != 0));
                        if (kSplit) {
                            umma_bf16(tmem_base + buf * 128, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + 32768 + k * 32), idesc1, 1u);
                            umma_bf16(tmem_base + buf * 128, umma_desc_sw128(a_addr + 32768 + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1, 1u);
                        }
                    }
                    tc_commit(&empty[s]);
                    if (kb == kHeadK1 / 64 - 1) tc_commit(&acc1_full[buf]);
                }
                __syncwarp();
            }
        };
        if (!kSplit) mbar_wait(w23_full, 0);
        uint32_t j = 0;
        if (blockIdx.x < m_tiles) issue_g1(0);
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++j) {
            if (t + (int)gridDim.x < m_tiles) issue_g1(j + 1);       // next tile's big GEMM runs under this tile's epilogues
            if (kSplit) continue;                                     // the tail runs on the CUDA cores
            mbar_wait(a2_ready, j & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_bf16(tmem_base + kColAcc2, umma_desc_sw128(smem_u32(s_a2) + (k >> 2) * 16384 + (k & 3) * 32),
                              umma_desc_sw128(smem_u32(s_w2) + (k >> 2) * 8192 + (k & 3) * 32), idesc2, (uint32_t)(k != 0));
                tc_commit(acc2_full);
            }
            __syncwarp();
            mbar_wait(a3_ready, j & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + kColAcc3, umma_desc_sw128(smem_u32(s_a3) + k * 32), umma_desc_sw128(smem_u32(s_w3) + k * 32), idesc3, (uint32_t)(k != 0));
                tc_commit(acc3_full);
            }
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int rt = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const StageParams st = load_stage(p.stage);
        const int fp16 = p.fp16;
        uint32_t j = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++j) {
            const int row = t * 128 + rt;
            const bool valid = row < p.M;
            SRHEP_CHECK(p.M <= p.ext.rows_cap && t * 128 + 127 < ((p.ext.rows_cap + 127) & ~127));
            const uint32_t buf = j & 1;
            if constexpr (kSplit) {                                   // E1': LeakyReLU(h1 + b1) -> fp32 rows
                mbar_wait(&acc1_full[buf], (j >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_lane + buf * 128 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(leaky_relu(fmaf(__uint_as_float(r[i]), p.ws1, p.b1[c * 32 + i])));
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 32; i += 8) stg256(p.h1out + (size_t)row * kHeadH1 + c * 32 + i, &r[i]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);
                continue;
            }
            // ---------------------------------------------------------------- E1
            {
                mbar_wait(&acc1_full[buf], (j >> 1) & 1);
                tc_fence_after();
                float v[128];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_lane + buf * 128 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[c * 32 + i] = leaky_relu(__uint_as_float(r[i]) + p.b1[c * 32 + i]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc1_empty[buf]);            // accumulator drained: the MMA warp may start the tile after next
                ln_inplace_128(v);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = v[c * 32 + i];
                    chain_store_a(smem_u32(s_a2), rt, c * 32, w, fp16);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(a2_ready);
            }
            // ---------------------------------------------------------------- E2
            {
                mbar_wait(acc2_full, j & 1);
                tc_fence_after();
                float v[64];
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_lane + kColAcc2 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) { v[c * 32 + i] = leaky_relu(__uint_as_float(r[i]) + p.b2[c * 32 + i]); s += v[c * 32 + i]; }
                }
                const float mean = s * (1.0f / 64.f);
                float qq = 0.f;
#pragma unroll
                for (int i = 0; i < 64; ++i) { const float d = v[i] - mean; qq = fmaf(d, d, qq); }
                const float rstd = rsqrtf(qq * (1.0f / 64.f) + kLnEps);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    float w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = (v[c * 32 + i] - mean) * rstd;
                    chain_store_a(smem_u32(s_a3), rt, c * 32, w, fp16);
                }
                fence_async_smem();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(a3_ready);
            }
            // ---------------------------------------------------------------- E3 + ODE update
            {
                mbar_wait(acc3_full, j & 1);
                tc_fence_after();
                uint32_t r[32];
                tmem_ld32(t_lane + kColAcc3, r);
                tmem_ld_wait();
                tc_fence_before();
                float v[32];
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) { v[i] = leaky_relu(__uint_as_float(r[i]) + p.b3[i]); s += v[i]; }
                if (p.final_ln) {
                    const float mean = s * (1.0f / 32.f);
                    float qq = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; qq = fmaf(d, d, qq); }
                    const float rstd = rsqrtf(qq * (1.0f / 32.f) + kLnEps);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd;
                }
                float vel = p.b4;
#pragma unroll
                for (int i = 0; i < 32; ++i) vel = fmaf(p.w4[i], v[i], vel);
                if (valid) {
                    if (st.vout) st.vout[row] = vel;
                    if (st.out) st.out[row] = fmaf(st.coef, vel, st.base[row]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace srhep

// Varlen self-attention on tensor cores, second generation (head_dim 64).
//
// One work item = (event, 128-query tile); blockIdx.y = head.  Keys/values of the item's event are
// streamed in tiles of 64 through a 4-stage TMA ring.  Per key tile:
//     S = Q K^T                 tcgen05, 128 x 64 fp32 in TMEM, DOUBLE-BUFFERED: the MMA warp issues S of
//                               tile j+1 before it waits for P of tile j, so the tensor core is never idle
//                               while the softmax warps work
//     P = exp2(S c - m)         ONE pass: each softmax thread holds its whole row of the tile (64 values) in
//                               registers (one TMEM read per element instead of two), fp32, lazy running max
//     O += P V                  tcgen05, P staged in shared memory as the 16-bit A operand (double-buffered),
//                               V consumed MN-major straight from its TMA tile
// Padded cells never enter (models/attention.py:238-265 + models/utils.py:23-34 restricted to real rows; keys
// past the event's end inside the last tile are masked to -inf in registers).
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: softmax + epilogue (thread = query row)
#pragma once
#include "kernels_bf16.cuh"

namespace srhep {

constexpr int kAtt2Threads = 192;
constexpr int kAtt2Stages = 4;
constexpr int kAtt2KvTile = 64;
#ifndef SRHEP_POLY_MASK
#define SRHEP_POLY_MASK 0x00
#endif
constexpr int kPolyMask = SRHEP_POLY_MASK;          // which of the 8 groups of 4 scores per 32 use the polynomial exp2 for their second pair (0xAA: every other group = 25 %).  Default 0: at the 1 000 W power cap the extra FMA instructions cost more than the SFU time they save (0x00 10 340, 0x22 10 315, 0xAA 10 270, 0xFF 10 230 events/s on one box)
constexpr uint32_t kAtt2OffKv = 16384;                                   // after Q
constexpr uint32_t kAtt2OffP = kAtt2OffKv + kAtt2Stages * 16384;         // 2 x 16 KB
constexpr uint32_t kAtt2OffBars = kAtt2OffP + 2 * 16384;
constexpr size_t kAtt2SmemBytes = kAtt2OffBars + 256;                     // 2 CTAs per SM

#ifdef SRHEP_TIMELINE
#define ATT_STAMP(item, k) do { if (p.dbg && blockIdx.x == 0 && blockIdx.y == 0 && (item) < 4 && (k) < 64) p.dbg[(item) * 64 + (k)] = clock64(); } while (0)
#else
#define ATT_STAMP(item, k) do { } while (0)
#endif


// One key tile of one query row: scores (64, or 32 when the second half of a ragged tile is all padding) -> P = exp2(S c - m)
// in the 16-bit A-operand layout, running sum and lazy running maximum.  kHalf2 = false writes zeros for the second half.
template <bool kFp16, bool kHalf2>
__device__ __forceinline__ void att2_softmax_tile(uint32_t t_s, uint64_t* s_empty_bar, int lane, int kv_valid, int j, float scale_log2, float& m_ref, float& l,
                                                  float& corr, bool& rescale, uint8_t* prow, int row, uint64_t* pv_bar, uint32_t pv_par) {
    constexpr int fp16 = kFp16 ? 1 : 0;
    uint32_t r0[32], r1[kHalf2 ? 32 : 1];
    tmem_ld32(t_s, r0);
    if (kHalf2) tmem_ld32(t_s + 32, reinterpret_cast<uint32_t (&)[32]>(r1));
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(s_empty_bar);                   // the scores live in registers now: S(it + 2) may be issued
    if (kv_valid < kAtt2KvTile) {                              // ragged last tile of the event: keys past its end count as -inf
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (i >= kv_valid) r0[i] = 0xff800000u;
            if (kHalf2 && 32 + i >= kv_valid) r1[i] = 0xff800000u;
        }
    }
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};       // four independent chains
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (kHalf2) mx4[u] = fmaxf(mx4[u], fmaxf(__uint_as_float(r0[i + u]), __uint_as_float(r1[i + u])));
            else mx4[u] = fmaxf(mx4[u], __uint_as_float(r0[i + u]));
        }
    }
    const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale_log2;
    if (mx > m_ref + 8.f) {                                    // first tile: m_ref = -inf
        if (j > 0) { corr = fast_exp2(m_ref - mx); rescale = true; }
        m_ref = mx;
    }
    uint32_t pk[32];
    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nm2 = pack_f32x2(-m_ref, -m_ref);
    uint64_t l2a = 0ull, l2b = 0ull;                            // two packed running sums = four independent chains
#pragma unroll
    for (int i = 0; i < 32; i += 4) {                          // exp2(-inf) = 0 takes care of the masked keys
        const uint64_t t0 = ffma2(pack_f32x2(__uint_as_float(r0[i]), __uint_as_float(r0[i + 1])), sc2, nm2);
        const uint64_t t1 = ffma2(pack_f32x2(__uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3])), sc2, nm2);
        const float e0 = fast_exp2(f32x2_lo(t0)), e1 = fast_exp2(f32x2_hi(t0));
        float e2, e3;
        if (kPolyMask & (1 << ((i >> 2) & 7))) exp2_poly_x2(t1, e2, e3);     // part of the exponentials leave the SFU for the FMA pipe
        else { e2 = fast_exp2(f32x2_lo(t1)); e3 = fast_exp2(f32x2_hi(t1)); }
        l2a = fadd2(l2a, pack_f32x2(e0, e1)); l2b = fadd2(l2b, pack_f32x2(e2, e3));
        pk[i >> 1] = pack16(e0, e1, fp16); pk[(i >> 1) + 1] = pack16(e2, e3, fp16);
    }
    if (kHalf2) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const uint64_t t0 = ffma2(pack_f32x2(__uint_as_float(r1[i]), __uint_as_float(r1[i + 1])), sc2, nm2);
            const uint64_t t1 = ffma2(pack_f32x2(__uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3])), sc2, nm2);
            const float e0 = fast_exp2(f32x2_lo(t0)), e1 = fast_exp2(f32x2_hi(t0));
            float e2, e3;
            if (kPolyMask & (1 << ((i >> 2) & 7))) exp2_poly_x2(t1, e2, e3);
            else { e2 = fast_exp2(f32x2_lo(t1)); e3 = fast_exp2(f32x2_hi(t1)); }
            l2a = fadd2(l2a, pack_f32x2(e0, e1)); l2b = fadd2(l2b, pack_f32x2(e2, e3));
            pk[16 + (i >> 1)] = pack16(e0, e1, fp16); pk[17 + (i >> 1)] = pack16(e2, e3, fp16);
        }
    }
    l = fmaf(l, corr, (f32x2_lo(l2a) + f32x2_hi(l2a)) + (f32x2_lo(l2b) + f32x2_hi(l2b)));
    if (pv_bar) mbar_wait(pv_bar, pv_par);                      // the P buffer is free once PV of tile it - 2 retired
#pragma unroll
    for (int g = 0; g < (kHalf2 ? 8 : 4); ++g)
        *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
    if (!kHalf2) {
#pragma unroll
        for (int g = 4; g < 8; ++g) *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);   // exp2(-inf) = 0 in either 16-bit format
    }
}

template <bool kFp16>
__global__ void __launch_bounds__(kAtt2Threads, 2) attn2_bf16_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                                                                     AttnBf16Params p) {
    extern __shared__ __align__(1024) uint8_t attn2_smem[];
    uint8_t* smem = attn2_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_q = smem;
    uint8_t* s_kv = smem + kAtt2OffKv;                 // stage s: K (64 keys x 128 B) at +0, V at +8192
    uint8_t* s_p = smem + kAtt2OffP;                   // 2 buffers of [128 rows x 128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAtt2OffBars);
    uint64_t* q_full = bars;             // TMA -> MMA
    uint64_t* q_empty = bars + 1;        // MMA -> TMA   (last QK^T of the item retired)
    uint64_t* kv_full = bars + 2;        // [4]
    uint64_t* kv_empty = bars + 6;       // [4]           (PV of the tile retired)
    uint64_t* s_full = bars + 10;        // [2] MMA -> softmax
    uint64_t* s_empty = bars + 12;       // [2] softmax -> MMA (S copied to registers)
    uint64_t* p_full = bars + 14;        // [2] softmax -> MMA (P in smem, O rescaled if needed)
    uint64_t* pv_done = bars + 16;       // [2] MMA -> softmax (PV retired: P buffer free, O readable)
    uint64_t* o_empty = bars + 18;       // epilogue -> MMA (O read out)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    constexpr uint32_t kTmemCols = 256;
    constexpr uint32_t kColO = 128;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_kv);
        mbar_init(q_full, 1); mbar_init(q_empty, 1);
        for (int i = 0; i < kAtt2Stages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
        mbar_init(o_empty, 4);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, item_i = 0;
            for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
                const AttnItem a = p.items[w];
                mbar_wait(q_empty, (item_i & 1) ^ 1);
                ATT_STAMP(item_i, 0);
                mbar_expect_tx(q_full, 16384);
                tma_load_2d(s_q, &tmap_q, q_full, head * 64, a.q_row);
                const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
                for (int j = 0; j < n_kv; ++j, ++it) {
                    const uint32_t s = it % kAtt2Stages, ph = (it / kAtt2Stages) & 1;
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    ATT_STAMP(item_i, 1 + j);
                    mbar_expect_tx(&kv_full[s], 16384);
                    tma_load_2d(s_kv + s * 16384, &tmap_kv, &kv_full[s], p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                    tma_load_2d(s_kv + s * 16384 + 8192, &tmap_kv, &kv_full[s], 2 * p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc_s = umma_idesc_16(128, 64, kFp16 ? 1 : 0);        // S = Q K^T
        const uint32_t idesc_o = umma_idesc_16(128, 64, kFp16 ? 1 : 0) | (1u << 16);  // O = P V (V MN-major)
        uint32_t it = 0, item_i = 0;
        auto issue_s = [&](uint32_t t, bool last_of_item) {                    // t = global key-tile counter
            const uint32_t s = t % kAtt2Stages, ph = (t / kAtt2Stages) & 1, b = t & 1;
            mbar_wait(&kv_full[s], ph);
            mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t qa = smem_u32(s_q), ka = smem_u32(s_kv + s * 16384);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16(tmem_base + b * 64, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + k * 32), idesc_s, (uint32_t)(k != 0));
                tc_commit(&s_full[b]);
                if (last_of_item) tc_commit(q_empty);
            }
            __syncwarp();
        };
        for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
            const AttnItem a = p.items[w];
            const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
            mbar_wait(q_full, item_i & 1);
            if (lane == 0) ATT_STAMP(item_i, 16);
            issue_s(it, n_kv == 1);
            if (lane == 0) ATT_STAMP(item_i, 17);
            for (int j = 0; j < n_kv; ++j, ++it) {
                if (j + 1 < n_kv) issue_s(it + 1, j + 2 == n_kv);          // next tile's scores run under this tile's softmax
                const uint32_t s = it % kAtt2Stages, b = it & 1;
                mbar_wait(&p_full[b], (it >> 1) & 1);
                if (lane == 0) ATT_STAMP(item_i, 24 + j);
                if (j == 0) mbar_wait(o_empty, (item_i & 1) ^ 1);          // previous item's O has been read out
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t pa = smem_u32(s_p + b * 16384), va = smem_u32(s_kv + s * 16384 + 8192);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + kColO, umma_desc_sw128(pa + k * 32), umma_desc_mn_sw128(va + k * 2048), idesc_o, (uint32_t)((j | k) != 0));
                    tc_commit(&kv_empty[s]);
                    tc_commit(&pv_done[b]);
                }
                __syncwarp();
            }
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        constexpr int fp16 = kFp16 ? 1 : 0;
        uint32_t it = 0, item_i = 0;
        // the descriptor of the NEXT item is fetched while this one is processed: its L2 round trip sat on the critical path of every item (10 % of the softmax warps' time)
        AttnItem a_next = (int)blockIdx.x < p.n_items ? p.items[blockIdx.x] : AttnItem{0, 0, 0, 0};
        for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
            const AttnItem a = a_next;
            if (w + (int)gridDim.x < p.n_items) a_next = p.items[w + gridDim.x];
            const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
            float m_ref = -INFINITY, l = 0.f;
            // query rows past the event's end (the last 128-row tile of an event is ragged): a warp whose 32 rows are all padding
            // keeps the barrier protocol going but does none of the arithmetic (15 % of the softmax work on single_e shapes)
            const bool wact = q * 32 < a.q_len;
            if (!wact) {
                for (int j = 0; j < n_kv; ++j, ++it) {
                    const uint32_t b = it & 1;
                    mbar_wait(&s_full[b], (it >> 1) & 1);
                    if (lane == 0) mbar_arrive(&s_empty[b]);
                    // the p_full phase of tile it - 2 must be over before this warp arrives for tile it
                    if (it >= 2) mbar_wait(&pv_done[b], ((it >> 1) - 1) & 1);
                    if (lane == 0) mbar_arrive(&p_full[b]);
                }
                mbar_wait(&pv_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
                if (lane == 0) mbar_arrive(o_empty);
                continue;
            }
            for (int j = 0; j < n_kv; ++j, ++it) {
                const uint32_t b = it & 1;
                const int kv_valid = min(kAtt2KvTile, a.k_len - j * kAtt2KvTile);
                const bool half2 = kv_valid > 32;                            // ragged last key tile with at most 32 keys: its second half is all padding
                mbar_wait(&s_full[b], (it >> 1) & 1);
                if (warp == 2 && lane == 0) ATT_STAMP(item_i, 32 + j);
                tc_fence_after();
                float corr = 1.f;
                bool rescale = false;
                // the P buffer is free once PV of tile it - 2 retired (waited for inside, right before the stores)
                uint64_t* pv_bar = it >= 2 ? &pv_done[b] : nullptr; const uint32_t pv_par = ((it >> 1) - 1) & 1;
                if (half2) att2_softmax_tile<kFp16, true>(t_lane + b * 64, &s_empty[b], lane, kv_valid, j, p.scale_log2, m_ref, l, corr, rescale, s_p + b * 16384 + row * 128, row, pv_bar, pv_par);
                else att2_softmax_tile<kFp16, false>(t_lane + b * 64, &s_empty[b], lane, kv_valid, j, p.scale_log2, m_ref, l, corr, rescale, s_p + b * 16384 + row * 128, row, pv_bar, pv_par);
                if (__any_sync(0xffffffffu, rescale)) {                    // rare: raise the reference maximum; O must be quiescent (PV of tile it - 1 retired)
                    mbar_wait(&pv_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(t_lane + kColO + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
                        tmem_st32(t_lane + kColO + c0, r);
                    }
                    tmem_st_wait();
                }
                fence_async_smem();                                         // P stores -> visible to the tensor core proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
                if (warp == 2 && lane == 0) ATT_STAMP(item_i, 40 + j);
            }
            // epilogue: O / l -> 16 bit -> global
            mbar_wait(&pv_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
            if (warp == 2 && lane == 0) ATT_STAMP(item_i, 48);
            tc_fence_after();
            const float inv = l > 0.f ? 1.f / l : 0.f;
            const bool valid = row < a.q_len;
            __nv_bfloat16* orow = p.out + (size_t)(a.q_row + row) * p.ldo + head * 64;
            uint32_t o0[32], o1[32];
            tmem_ld32(t_lane + kColO, o0);
            tmem_ld32(t_lane + kColO + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            if (valid) {
                uint32_t pk[32];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    pk[i] = pack16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv, fp16);
                    pk[16 + i] = pack16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv, fp16);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) stg256(orow + 16 * g, &pk[8 * g]);
            }
            if (warp == 2 && lane == 0) ATT_STAMP(item_i, 49);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace srhep

// Varlen self-attention on tensor cores, third generation (head_dim 64).
//
// Same work decomposition and arithmetic as kernels_attn.cuh (one work item = (event, 128-query tile), blockIdx.y = head,
// 64-key tiles, S double-buffered in TMEM, one-pass fp32 softmax with a lazy running maximum, exp2 with 1/sqrt(hd) folded
// in; models/attention.py:238-265 + models/utils.py:23-34 restricted to real rows).  What changed, from the CTA timeline
// of the second generation (profiles/r02_attn_timeline.txt: 1 500 cycles per key tile in steady state, but 12 300 per
// 6-tile item: the Q tile of the next item was requested only when the last QK^T of the current one had retired and took
// 3 600 - 4 200 cycles to arrive, with the first K tile queued behind it):
//   * P never touches shared memory: the softmax warps write it as the 16-bit A operand straight into TENSOR MEMORY
//     (tcgen05.st, 32 columns per 64-key tile, double-buffered) and O += P V reads A from TMEM
//     (tcgen05.mma ... [d], [a_tmem], b_desc).  That removes 16 KB of st.shared + the async-proxy fence per tile and
//     frees 32 KB of shared memory, which pays for
//   * a DOUBLE-BUFFERED Q tile and a 5-stage K/V ring: the producer runs ahead ACROSS work items (the Q and the first
//     K/V tiles of item i + 1 are in flight while item i is still in its softmax), and
//   * ragged key tiles execute only what they hold: S = Q K^T with N = kv_valid rounded up to 16, O += P V over
//     ceil(kv_valid / 16) k-steps.
// With kSplit the same kernel is the fp32-grade attention of the 'highest' path: every 16-bit operand is a PAIR of fp16
// planes (x = hi + lo, |lo| <= 2^-11 |hi|) and each product runs as hi.hi + hi.lo + lo.hi on the same fp32 accumulator
// (see kernels_split.cuh); one CTA per SM then (twice the shared memory and 320 TMEM columns).
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-5: softmax + epilogue (thread = query row)
#pragma once
#include "kernels_attn.cuh"

namespace srhep {

constexpr int kAtt3Threads = 192;
template <bool kSplit> struct Att3Cfg {
    static constexpr int kPlanes = kSplit ? 2 : 1;
    static constexpr int kStages = kSplit ? 4 : 5;
    static constexpr uint32_t kQBytes = 16384 * kPlanes;                  // one Q buffer (hi [| lo])
    static constexpr uint32_t kStageBytes = 16384 * kPlanes;              // K hi | V hi [| K lo | V lo]
    static constexpr uint32_t kOffKv = 2 * kQBytes;
    static constexpr uint32_t kOffBars = kOffKv + kStages * kStageBytes;
    static constexpr size_t kSmemBytes = kOffBars + 256;
    static constexpr uint32_t kTmemCols = kSplit ? 512 : 256;
    static constexpr uint32_t kColO = 128;
    static constexpr uint32_t kColP = 192;                                // buffer b at + b * 32 * kPlanes (hi, then lo)
};

// the A operand of an MMA taken from tensor memory (K-major, one row per lane, two 16-bit values per 32-bit column)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
          "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

// One key tile of one query row: scores (64, or 32 when the second half of a ragged tile is all padding) -> P = exp2(S c - m)
// as the 16-bit A operand in tensor memory, running sum and lazy running maximum.
template <bool kFp16, bool kSplit, bool kHalf2>
__device__ __forceinline__ void att3_softmax_tile(uint32_t t_s, uint32_t t_p, uint64_t* s_empty_bar, int lane, int kv_valid, int j, float scale_log2, float& m_ref, float& l,
                                                  float& corr, bool& rescale, uint64_t* pv_bar, uint32_t pv_par) {
    constexpr int fp16 = kFp16 ? 1 : 0;
    constexpr int NP = kHalf2 ? 32 : 16;
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
    uint32_t pk[NP], pl[kSplit ? NP : 1];
    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nm2 = pack_f32x2(-m_ref, -m_ref);
    uint64_t l2a = 0ull, l2b = 0ull;                            // two packed running sums = four independent chains
    auto half = [&](const uint32_t (&r)[32], int o) {
#pragma unroll
        for (int i = 0; i < 32; i += 4) {                      // exp2(-inf) = 0 takes care of the masked keys
            const uint64_t t0 = ffma2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), sc2, nm2);
            const uint64_t t1 = ffma2(pack_f32x2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2, nm2);
            const float e0 = fast_exp2(f32x2_lo(t0)), e1 = fast_exp2(f32x2_hi(t0));
            const float e2 = fast_exp2(f32x2_lo(t1)), e3 = fast_exp2(f32x2_hi(t1));
            l2a = fadd2(l2a, pack_f32x2(e0, e1)); l2b = fadd2(l2b, pack_f32x2(e2, e3));
            if (kSplit) { split16(e0, e1, pk[o + (i >> 1)], pl[kSplit ? o + (i >> 1) : 0]); split16(e2, e3, pk[o + (i >> 1) + 1], pl[kSplit ? o + (i >> 1) + 1 : 0]); }
            else { pk[o + (i >> 1)] = pack16(e0, e1, fp16); pk[o + (i >> 1) + 1] = pack16(e2, e3, fp16); }
        }
    };
    half(r0, 0);
    if (kHalf2) half(reinterpret_cast<const uint32_t (&)[32]>(r1), 16);
    l = fmaf(l, corr, (f32x2_lo(l2a) + f32x2_hi(l2a)) + (f32x2_lo(l2b) + f32x2_hi(l2b)));
    if (pv_bar) { mbar_wait(pv_bar, pv_par); tc_fence_after(); }      // the P buffer is free once PV of tile it - 2 retired
    if (kHalf2) {
        tmem_st32(t_p, reinterpret_cast<const uint32_t (&)[32]>(pk));
        if (kSplit) tmem_st32(t_p + 32, reinterpret_cast<const uint32_t (&)[32]>(pl));
    } else {
        tmem_st16(t_p, pk);
        if (kSplit) tmem_st16(t_p + 32, pl);
    }
}

template <bool kFp16, bool kSplit>
__global__ void __launch_bounds__(kAtt3Threads, kSplit ? 1 : 2) attn3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                                                                             const __grid_constant__ CUtensorMap tmap_q_lo, const __grid_constant__ CUtensorMap tmap_kv_lo,
                                                                             AttnBf16Params p) {
    using Cfg = Att3Cfg<kSplit>;
    static_assert(!kSplit || kFp16, "the split (fp32-grade) mode runs on fp16 planes");
    extern __shared__ __align__(1024) uint8_t attn3_smem[];
    uint8_t* smem = attn3_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_q = smem;                               // buffer i: hi at + i * kQBytes, lo 16 KB later
    uint8_t* s_kv = smem + Cfg::kOffKv;                // stage s: K hi at +0, V hi at +8192 [, K lo at +16384, V lo at +24576]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kOffBars);
    uint64_t* q_full = bars;             // [2] TMA -> MMA
    uint64_t* q_empty = bars + 2;        // [2] MMA -> TMA   (last QK^T of the item retired)
    uint64_t* kv_full = bars + 4;        // [stages]
    uint64_t* kv_empty = bars + 9;       // [stages]         (PV of the tile retired)
    uint64_t* s_full = bars + 14;        // [2] MMA -> softmax
    uint64_t* s_empty = bars + 16;       // [2] softmax -> MMA (S copied to registers)
    uint64_t* p_full = bars + 18;        // [2] softmax -> MMA (P in tensor memory, O rescaled if needed)
    uint64_t* pv_done = bars + 20;       // [2] MMA -> softmax (PV retired: P buffer free, O readable)
    uint64_t* o_empty = bars + 22;       // epilogue -> MMA (O read out)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 23);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    constexpr int kStages = Cfg::kStages;
    constexpr uint32_t kPCols = 32 * Cfg::kPlanes;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q); prefetch_tmap(&tmap_kv);
        if (kSplit) { prefetch_tmap(&tmap_q_lo); prefetch_tmap(&tmap_kv_lo); }
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
        for (int i = 0; i < kStages; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); mbar_init(&p_full[i], 4); mbar_init(&pv_done[i], 1); }
        mbar_init(o_empty, 4);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0, item_i = 0;
            for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
                const AttnItem a = p.items[w];
                const uint32_t qb = item_i & 1;
                mbar_wait(&q_empty[qb], ((item_i >> 1) & 1) ^ 1);      // the buffer of item i - 2: free long ago, the producer runs ahead across items
                ATT_STAMP(item_i, 0);
                mbar_expect_tx(&q_full[qb], Cfg::kQBytes);
                tma_load_2d(s_q + qb * Cfg::kQBytes, &tmap_q, &q_full[qb], head * 64, a.q_row);
                if (kSplit) tma_load_2d(s_q + qb * Cfg::kQBytes + 16384, &tmap_q_lo, &q_full[qb], head * 64, a.q_row);
                const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
                for (int j = 0; j < n_kv; ++j, ++it) {
                    const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                    mbar_wait(&kv_empty[s], ph ^ 1);
                    ATT_STAMP(item_i, 1 + j);
                    uint8_t* st = s_kv + s * Cfg::kStageBytes;
                    mbar_expect_tx(&kv_full[s], Cfg::kStageBytes);
                    tma_load_2d(st, &tmap_kv, &kv_full[s], p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                    tma_load_2d(st + 8192, &tmap_kv, &kv_full[s], 2 * p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                    if (kSplit) {
                        tma_load_2d(st + 16384, &tmap_kv_lo, &kv_full[s], p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                        tma_load_2d(st + 24576, &tmap_kv_lo, &kv_full[s], 2 * p.h_dim + head * 64, a.k_row + j * kAtt2KvTile);
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr int fp16 = kFp16 ? 1 : 0;
        uint32_t it = 0, item_i = 0;
        auto issue_s = [&](uint32_t t, uint32_t qb, int kv_valid, bool last_of_item) {        // t = global key-tile counter
            const uint32_t s = t % kStages, ph = (t / kStages) & 1, b = t & 1;
            mbar_wait(&kv_full[s], ph);
            mbar_wait(&s_empty[b], ((t >> 1) & 1) ^ 1);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t idesc_s = umma_idesc_16(128, (kv_valid + 15) & ~15, fp16);      // a ragged tile computes only the key columns it holds (N % 16 == 0)
                const uint32_t qa = smem_u32(s_q + qb * Cfg::kQBytes), ka = smem_u32(s_kv + s * Cfg::kStageBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_bf16(tmem_base + b * 64, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + k * 32), idesc_s, (uint32_t)(k != 0));
                    if (kSplit) {
                        umma_bf16(tmem_base + b * 64, umma_desc_sw128(qa + k * 32), umma_desc_sw128(ka + 16384 + k * 32), idesc_s, 1u);
                        umma_bf16(tmem_base + b * 64, umma_desc_sw128(qa + 16384 + k * 32), umma_desc_sw128(ka + k * 32), idesc_s, 1u);
                    }
                }
                tc_commit(&s_full[b]);
                if (last_of_item) tc_commit(&q_empty[qb]);
            }
            __syncwarp();
        };
        const uint32_t idesc_o = umma_idesc_16(128, 64, fp16) | (1u << 16);                    // O = P V (V MN-major)
        // The scores of a tile are issued one tile ahead of its softmax, ACROSS work items: S(first tile of item i + 1) goes out before
        // P(last tile of item i) is awaited (its Q tile landed long ago), so the softmax warps find it ready after their epilogue.
        bool first_issued = false;
        AttnItem a_nx = (int)blockIdx.x < p.n_items ? p.items[blockIdx.x] : AttnItem{0, 0, 0, 0};
        for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
            const AttnItem a = a_nx;
            const bool has_next = w + (int)gridDim.x < p.n_items;
            if (has_next) a_nx = p.items[w + gridDim.x];
            const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
            const uint32_t qb = item_i & 1;
            auto valid_of = [&](int j) { return min(kAtt2KvTile, a.k_len - j * kAtt2KvTile); };
            if (!first_issued) {
                mbar_wait(&q_full[qb], (item_i >> 1) & 1);
                if (lane == 0) ATT_STAMP(item_i, 16);
                issue_s(it, qb, valid_of(0), n_kv == 1);
                if (lane == 0) ATT_STAMP(item_i, 17);
            }
            first_issued = false;
            for (int j = 0; j < n_kv; ++j, ++it) {
                if (j + 1 < n_kv) issue_s(it + 1, qb, valid_of(j + 1), j + 2 == n_kv);         // next tile's scores run under this tile's softmax
                else if (has_next) {                                                           // ... and so do the first scores of the next item
                    const int nkv2 = (a_nx.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
                    mbar_wait(&q_full[qb ^ 1], ((item_i + 1) >> 1) & 1);
                    if (lane == 0) ATT_STAMP(item_i + 1, 16);
                    issue_s(it + 1, qb ^ 1, min(kAtt2KvTile, a_nx.k_len), nkv2 == 1);
                    if (lane == 0) ATT_STAMP(item_i + 1, 17);
                    first_issued = true;
                }
                const uint32_t s = it % kStages, b = it & 1;
                mbar_wait(&p_full[b], (it >> 1) & 1);
                if (lane == 0) ATT_STAMP(item_i, 24 + j);
                if (j == 0) mbar_wait(o_empty, (item_i & 1) ^ 1);          // previous item's O has been read out
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t pa = tmem_base + Cfg::kColP + b * kPCols, va = smem_u32(s_kv + s * Cfg::kStageBytes + 8192);
                    const int nk = (valid_of(j) + 15) >> 4;                 // 16 keys per k-step; P is zero past kv_valid inside the last one
                    for (int k = 0; k < nk; ++k) {
                        umma_ts(tmem_base + Cfg::kColO, pa + k * 8, umma_desc_mn_sw128(va + k * 2048), idesc_o, (uint32_t)((j | k) != 0));
                        if (kSplit) {
                            umma_ts(tmem_base + Cfg::kColO, pa + k * 8, umma_desc_mn_sw128(va + 16384 + k * 2048), idesc_o, 1u);
                            umma_ts(tmem_base + Cfg::kColO, pa + 32 + k * 8, umma_desc_mn_sw128(va + k * 2048), idesc_o, 1u);
                        }
                    }
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
        // The epilogue of an item (O / l -> 16 bit -> global) is DEFERRED until the first key tile of the next item has been through the
        // softmax: the last O += P V of the item retires meanwhile instead of being waited for with the SFU idle.  (The first tile of an item
        // never touches O; the MMA warp holds the next item's first P V back until o_empty, which the deferred epilogue signals.)
        bool pend = false; float pend_inv = 0.f; int pend_qrow = 0, pend_qlen = 0; uint32_t pend_it = 0;
        auto epilogue = [&]() {
            mbar_wait(&pv_done[(pend_it - 1) & 1], ((pend_it - 1) >> 1) & 1);
            if (warp == 2 && lane == 0) ATT_STAMP(item_i, 48);
            tc_fence_after();
            const bool valid = row < pend_qlen;
            __nv_bfloat16* orow = p.out + (size_t)(pend_qrow + row) * p.ldo + head * 64;
            uint32_t o0[32], o1[32];
            tmem_ld32(t_lane + Cfg::kColO, o0);
            tmem_ld32(t_lane + Cfg::kColO + 32, o1);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_empty);
            const float inv = pend_inv;
            if (valid) {
                uint32_t pk[32];
                if constexpr (kSplit) {
                    uint32_t pl[32];
                    __nv_bfloat16* lrow = p.out_lo + (size_t)(pend_qrow + row) * p.ldo + head * 64;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        split16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv, pk[i], pl[i]);
                        split16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv, pk[16 + i], pl[16 + i]);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) { stg256(orow + 16 * g, &pk[8 * g]); stg256(lrow + 16 * g, &pl[8 * g]); }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        pk[i] = pack16(__uint_as_float(o0[2 * i]) * inv, __uint_as_float(o0[2 * i + 1]) * inv, fp16);
                        pk[16 + i] = pack16(__uint_as_float(o1[2 * i]) * inv, __uint_as_float(o1[2 * i + 1]) * inv, fp16);
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) stg256(orow + 16 * g, &pk[8 * g]);
                }
            }
            if (warp == 2 && lane == 0) ATT_STAMP(item_i, 49);
            pend = false;
        };
        // the descriptor of the NEXT item is fetched while this one is processed
        AttnItem a_next = (int)blockIdx.x < p.n_items ? p.items[blockIdx.x] : AttnItem{0, 0, 0, 0};
        for (int w = blockIdx.x; w < p.n_items; w += gridDim.x, ++item_i) {
            const AttnItem a = a_next;
            if (w + (int)gridDim.x < p.n_items) a_next = p.items[w + gridDim.x];
            const int n_kv = (a.k_len + kAtt2KvTile - 1) / kAtt2KvTile;
            float m_ref = -INFINITY, l = 0.f;
            // query rows past the event's end (the last 128-row tile of an event is ragged): a warp whose 32 rows are all padding
            // keeps the barrier protocol going but does none of the arithmetic (its P rows stay whatever they were: rows are independent)
            const bool wact = q * 32 < a.q_len;
            if (!wact) {
                if (pend) epilogue();
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
                uint64_t* pv_bar = it >= 2 ? &pv_done[b] : nullptr; const uint32_t pv_par = ((it >> 1) - 1) & 1;
                const uint32_t t_p = t_lane + Cfg::kColP + b * kPCols;
                if (half2) att3_softmax_tile<kFp16, kSplit, true>(t_lane + b * 64, t_p, &s_empty[b], lane, kv_valid, j, p.scale_log2, m_ref, l, corr, rescale, pv_bar, pv_par);
                else att3_softmax_tile<kFp16, kSplit, false>(t_lane + b * 64, t_p, &s_empty[b], lane, kv_valid, j, p.scale_log2, m_ref, l, corr, rescale, pv_bar, pv_par);
                if (__any_sync(0xffffffffu, rescale)) {                    // rare: raise the reference maximum; O must be quiescent (PV of tile it - 1 retired)
                    mbar_wait(&pv_done[(it - 1) & 1], ((it - 1) >> 1) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = 0; c0 < 64; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(t_lane + Cfg::kColO + c0, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
                        tmem_st32(t_lane + Cfg::kColO + c0, r);
                    }
                }
                tmem_st_wait();                                             // P (and a rescaled O) are in tensor memory
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[b]);
                if (warp == 2 && lane == 0) ATT_STAMP(item_i, 40 + j);
                if (j == 0 && pend) epilogue();                            // the previous item's output, its last P V retired by now
            }
            pend = true; pend_inv = l > 0.f ? 1.f / l : 0.f; pend_qrow = a.q_row; pend_qlen = a.q_len; pend_it = it;
        }
        if (pend) epilogue();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, Cfg::kTmemCols); }
}

}  // namespace srhep

// Velocity head in ONE kernel for sm_100a (round 2): the head preparation (models/flow_model.py:241-245 and the first
// LayerNorm of v_t_pred_net, models/dense.py:62) no longer writes its 512-wide 16-bit operand row to HBM for head_chain_kernel
// to read back: eight preparation warps of the CTA compute
//
//   hin[row] = LN_512( cat[ modulate(norm_v_t(cat[final_norm(x), cond_feat])), context ] )
//
// for the 128 rows of a tile and write them, as the K-major 128-byte-swizzled A operand, STRAIGHT into shared memory; the
// three tcgen05 GEMMs, their LayerNorm epilogues and the ODE update follow as in kernels_head.cuh.  Per cell the head now
// reads 1 KB (residual) + 384 B (cond_feat) and writes 4-8 B; the 1 KB store and the 1 KB load of `hin` are gone
// (2.1 GB per evaluation of 4096 single_e events) and so is one launch.
//
//   warp 0: producer (W1 streamed from L2 through a ring of (k-block, 64-row half) slots; W2 / W3 once)
//   warp 1: TMEM allocator + MMA issuer          warps 2-5: epilogues E1-E3 + ODE update, one thread = one row
//   warps 6-13: head preparation, one warp = one row (two rows interleaved), rows w, w + 8, ... of the tile
//
// Row math of the preparation warps: every lane owns groups of 4 consecutive columns (h -> 2 pieces per lane, cond -> 1 for
// lanes 0-23, ctx -> 1 + 1 for lanes 0-7).  The three chained LayerNorms take ONE warp reduction round each (sum and sum of
// squares together; the cond_feat sums ride along in the first round, the event's context sums are computed when the event
// changes), two rows share every round, and norm_v_t's affine and the adaLN modulation are combined per event into one
// multiply-add  P = w (1 + scale),  Q = b (1 + scale) + shift.  The same row math backs head_prep_v5_kernel, the stand-alone
// variant (diagnostic switch SRHEP_NO_HEAD_FUSED=1: head_prep_v5 + head_chain_kernel).
#pragma once
#include "kernels_f32.cuh"
#include "kernels_head.cuh"

namespace srhep {

struct HeadPrepLane {
    float4 fw0, fw1, fb0, fb1;          // final_norm weight / bias pieces (columns 4 lane .., 128 + 4 lane ..)
    float4 P0, P1, Pc, Q0, Q1, Qc;      // per event: nv_w (1 + scale),  nv_b (1 + scale) + shift   (c: cond columns 256 + 4 lane .., lanes 0-23, else 0)
    float4 x0, x1;                      // the event's context pieces (x1: lanes 0-7, else 0)
    float cs, cq;                       // sum / sum of squares of the event's 160 context values
    int ev;
};

__device__ __forceinline__ float4 hp_ld4(const float* q) { return *reinterpret_cast<const float4*>(q); }
__device__ __forceinline__ float4 hp_mul4(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 hp_fma4(float4 a, float4 b, float4 c) { return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)); }
// packed accumulation of a piece into (sum pair, sum-of-squares pair)
__device__ __forceinline__ void hp_acc(float4 v, uint64_t& s, uint64_t& q) {
    const uint64_t lo = pack_f32x2(v.x, v.y), hi = pack_f32x2(v.z, v.w);
    s = fadd2(s, lo); s = fadd2(s, hi);
    q = ffma2(lo, lo, q); q = ffma2(hi, hi, q);
}
// (v r + nmr) w + b on packed pairs
__device__ __forceinline__ float4 hp_aff4(float4 v, uint64_t rr, uint64_t nm, float4 w, float4 b) {
    const uint64_t y0 = ffma2(ffma2(pack_f32x2(v.x, v.y), rr, nm), pack_f32x2(w.x, w.y), pack_f32x2(b.x, b.y));
    const uint64_t y1 = ffma2(ffma2(pack_f32x2(v.z, v.w), rr, nm), pack_f32x2(w.z, w.w), pack_f32x2(b.z, b.w));
    return make_float4(f32x2_lo(y0), f32x2_hi(y0), f32x2_lo(y1), f32x2_hi(y1));
}
__device__ __forceinline__ float4 hp_nrm4(float4 v, uint64_t rr, uint64_t nm) {
    const uint64_t y0 = ffma2(pack_f32x2(v.x, v.y), rr, nm), y1 = ffma2(pack_f32x2(v.z, v.w), rr, nm);
    return make_float4(f32x2_lo(y0), f32x2_hi(y0), f32x2_lo(y1), f32x2_hi(y1));
}

__device__ __forceinline__ void hp_init(HeadPrepLane& L, const HeadPrepParams& p, int lane, bool fn_in_regs = true) {
    const int c0 = lane * 4, c1 = 128 + lane * 4;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (fn_in_regs) { L.fw0 = hp_ld4(p.fn_w + c0); L.fw1 = hp_ld4(p.fn_w + c1); L.fb0 = hp_ld4(p.fn_b + c0); L.fb1 = hp_ld4(p.fn_b + c1); }
    else { L.fw0 = L.fw1 = L.fb0 = L.fb1 = z4; }
    L.P0 = L.P1 = L.Pc = L.Q0 = L.Q1 = L.Qc = L.x0 = L.x1 = z4;
    L.cs = L.cq = 0.f; L.ev = -1;
}
// the event changes (warp-uniform): context pieces and their sums, combined norm_v_t affine + modulation
__device__ __forceinline__ void hp_event(HeadPrepLane& L, const HeadPrepParams& p, int lane, int ev) {
    const int c0 = lane * 4, c1 = 128 + lane * 4;
    const bool hc = lane < 24, hx1 = lane < 8;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f), one4 = make_float4(1.f, 1.f, 1.f, 1.f);
    const float* cx = p.ctx + (size_t)ev * 160;
    const float* sh = p.shift + (size_t)ev * p.ld_mod; const float* sc = p.scale + (size_t)ev * p.ld_mod;
    L.x0 = hp_ld4(cx + c0); L.x1 = hx1 ? hp_ld4(cx + c1) : z4;
    float4 s = hp_fma4(hp_ld4(sc + c0), one4, one4);
    L.P0 = hp_mul4(hp_ld4(p.nv_w + c0), s); L.Q0 = hp_fma4(hp_ld4(p.nv_b + c0), s, hp_ld4(sh + c0));
    s = hp_fma4(hp_ld4(sc + c1), one4, one4);
    L.P1 = hp_mul4(hp_ld4(p.nv_w + c1), s); L.Q1 = hp_fma4(hp_ld4(p.nv_b + c1), s, hp_ld4(sh + c1));
    if (hc) {
        s = hp_fma4(hp_ld4(sc + 256 + c0), one4, one4);
        L.Pc = hp_mul4(hp_ld4(p.nv_w + 256 + c0), s); L.Qc = hp_fma4(hp_ld4(p.nv_b + 256 + c0), s, hp_ld4(sh + 256 + c0));
    } else { L.Pc = z4; L.Qc = z4; }
    uint64_t s2 = 0ull, q2 = 0ull;
    hp_acc(L.x0, s2, q2); hp_acc(L.x1, s2, q2);
    float a = f32x2_lo(s2) + f32x2_hi(s2), b = f32x2_lo(q2) + f32x2_hi(q2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    L.cs = a; L.cq = b; L.ev = ev;
}

// Sum of V values over the 32 lanes, every lane ends with all V totals.  Instead of V independent butterflies (5 V dependent
// shuffle + add pairs) the values are TRANSPOSED while they are reduced: at each of the first log2(V) levels a lane keeps one half of
// its values and sends the other half to its partner, so that V / 2, V / 4, ... shuffles do the work; the last levels reduce the one
// value left, and V independent broadcasts hand the totals back (V = 8: 17 shuffles of which 8 are independent, instead of 40).
template <int V>
__device__ __forceinline__ void hp_reduce(float (&v)[V], int lane) {
    static_assert(V == 2 || V == 4 || V == 8, "hp_reduce");
    if constexpr (V == 2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { v[0] += __shfl_xor_sync(0xffffffffu, v[0], o); v[1] += __shfl_xor_sync(0xffffffffu, v[1], o); }
    } else {
        int o = 16;
#pragma unroll
        for (int cnt = V; cnt > 1; cnt >>= 1, o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int k = 0; k < cnt / 2; ++k) {
                const float send = up ? v[k] : v[k + cnt / 2];
                const float keep = up ? v[k + cnt / 2] : v[k];
                v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
#pragma unroll
        for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
        // lane bits 4, 3 (, 2) say which value this lane holds: value i sits in lane (i's bits, most significant first) << (V == 8 ? 2 : 3)
        const float mine = v[0];
#pragma unroll
        for (int i = 0; i < V; ++i) v[i] = __shfl_sync(0xffffffffu, mine, V == 8 ? (i << 2) : (i << 3));
    }
}

// NR rows of the same event at once.  a0 / a1: the residual pieces, cc: the cond_feat piece (zero for lanes >= 24).
// sink(r, piece, float4) receives the finished pieces of row r: columns hp_piece_col(piece, lane) .. + 3 of the 512-wide operand row.
// fn_sh != 0: final_norm weight | bias (2 x 256 floats) are read from that shared-memory address instead of L.fw* / L.fb* (the fused
// kernel's preparation warps: 16 registers less, no spills under their 128-register budget).
template <int NR, typename Sink>
__device__ __forceinline__ void hp_rows(const HeadPrepLane& L, float4 (&a0)[NR], float4 (&a1)[NR], float4 (&cc)[NR], int lane, Sink&& sink, uint32_t fn_sh = 0) {
    const bool hc = lane < 24, hx1 = lane < 8;
    float4 fw0, fw1, fb0, fb1;
    if (fn_sh) { fw0 = lds_f4(fn_sh + lane * 16); fw1 = lds_f4(fn_sh + 512 + lane * 16); fb0 = lds_f4(fn_sh + 1024 + lane * 16); fb1 = lds_f4(fn_sh + 1536 + lane * 16); }
    else { fw0 = L.fw0; fw1 = L.fw1; fb0 = L.fb0; fb1 = L.fb1; }
    // round 1: final_norm statistics over h = 256; the cond_feat sums ride along
    float v1[NR * 4];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        uint64_t s = 0ull, q = 0ull, cs = 0ull, cq = 0ull;
        hp_acc(a0[r], s, q); hp_acc(a1[r], s, q); hp_acc(cc[r], cs, cq);
        v1[r * 4] = f32x2_lo(s) + f32x2_hi(s); v1[r * 4 + 1] = f32x2_lo(q) + f32x2_hi(q); v1[r * 4 + 2] = f32x2_lo(cs) + f32x2_hi(cs); v1[r * 4 + 3] = f32x2_lo(cq) + f32x2_hi(cq);
    }
    hp_reduce<NR * 4>(v1, lane);
    float v[NR][4];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) v[r][k] = v1[r * 4 + k];
    float w2[NR * 2];
    float w[NR][2];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float mean = v[r][0] * (1.0f / 256.f);
        const float rstd = rsqrtf(fmaxf(v[r][1] * (1.0f / 256.f) - mean * mean, 0.f) + kLnEps);
        const uint64_t rr = pack_f32x2(rstd, rstd), nm = pack_f32x2(-mean * rstd, -mean * rstd);
        a0[r] = hp_aff4(a0[r], rr, nm, fw0, fb0);
        a1[r] = hp_aff4(a1[r], rr, nm, fw1, fb1);
        // round 2: norm_v_t statistics over h + cond = 352
        uint64_t s = 0ull, q = 0ull;
        hp_acc(a0[r], s, q); hp_acc(a1[r], s, q);
        w2[r * 2] = f32x2_lo(s) + f32x2_hi(s); w2[r * 2 + 1] = f32x2_lo(q) + f32x2_hi(q);
    }
    hp_reduce<NR * 2>(w2, lane);
#pragma unroll
    for (int r = 0; r < NR; ++r) { w[r][0] = w2[r * 2]; w[r][1] = w2[r * 2 + 1]; }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float mean = (w[r][0] + v[r][2]) * (1.0f / 352.f);
        const float rstd = rsqrtf(fmaxf((w[r][1] + v[r][3]) * (1.0f / 352.f) - mean * mean, 0.f) + kLnEps);
        const uint64_t rr = pack_f32x2(rstd, rstd), nm = pack_f32x2(-mean * rstd, -mean * rstd);
        a0[r] = hp_aff4(a0[r], rr, nm, L.P0, L.Q0);
        a1[r] = hp_aff4(a1[r], rr, nm, L.P1, L.Q1);
        cc[r] = hp_aff4(cc[r], rr, nm, L.Pc, L.Qc);          // lanes >= 24: P = Q = 0 -> 0
        // round 3: LayerNorm (no affine) over v_in + ctx = 512; the context sums are the event's
        uint64_t s = 0ull, q = 0ull;
        hp_acc(a0[r], s, q); hp_acc(a1[r], s, q); hp_acc(cc[r], s, q);
        w2[r * 2] = f32x2_lo(s) + f32x2_hi(s); w2[r * 2 + 1] = f32x2_lo(q) + f32x2_hi(q);
    }
    hp_reduce<NR * 2>(w2, lane);
#pragma unroll
    for (int r = 0; r < NR; ++r) { w[r][0] = w2[r * 2]; w[r][1] = w2[r * 2 + 1]; }
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const float mean = (w[r][0] + L.cs) * (1.0f / 512.f);
        const float rstd = rsqrtf(fmaxf((w[r][1] + L.cq) * (1.0f / 512.f) - mean * mean, 0.f) + kLnEps);
        const uint64_t rr = pack_f32x2(rstd, rstd), nm = pack_f32x2(-mean * rstd, -mean * rstd);
        sink(r, 0, hp_nrm4(a0[r], rr, nm));
        sink(r, 1, hp_nrm4(a1[r], rr, nm));
        if (hc) sink(r, 2, hp_nrm4(cc[r], rr, nm));
        sink(r, 3, hp_nrm4(L.x0, rr, nm));
        if (hx1) sink(r, 4, hp_nrm4(L.x1, rr, nm));
    }
}

// first column of a lane's piece 0..4: residual (two), cond_feat (lanes 0-23), context (two, the second for lanes 0-7)
__device__ __forceinline__ int hp_piece_col(int piece, int lane) { return (piece == 0 ? 0 : piece == 1 ? 128 : piece == 2 ? 256 : piece == 3 ? 352 : 480) + lane * 4; }
__device__ __forceinline__ uint2 hp_pack4(float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 w; w.x = *reinterpret_cast<const uint32_t*>(&a); w.y = *reinterpret_cast<const uint32_t*>(&b);
    return w;
}

// One pair of rows (a, b) of a warp: same event -> both rows share the three reduction rounds; else one after the other.
template <typename Sink>
__device__ __forceinline__ void hp_pair(HeadPrepLane& L, const HeadPrepParams& p, int lane, bool two, const int (&ev)[2],
                                        float4 (&a0)[2], float4 (&a1)[2], float4 (&cc)[2], Sink&& sink, uint32_t fn_sh = 0) {
    if (ev[0] != L.ev) hp_event(L, p, lane, ev[0]);
    if (two && ev[1] == ev[0]) { hp_rows<2>(L, a0, a1, cc, lane, sink, fn_sh); return; }
    {
        float4 b0[1] = {a0[0]}, b1[1] = {a1[0]}, bc[1] = {cc[0]};
        hp_rows<1>(L, b0, b1, bc, lane, [&](int, int piece, float4 v) { sink(0, piece, v); }, fn_sh);
    }
    if (two) {
        hp_event(L, p, lane, ev[1]);
        float4 b0[1] = {a0[1]}, b1[1] = {a1[1]}, bc[1] = {cc[1]};
        hp_rows<1>(L, b0, b1, bc, lane, [&](int, int piece, float4 v) { sink(1, piece, v); }, fn_sh);
    }
}

// Stand-alone head preparation (16-bit operand rows to global memory) on the same row math: 64 rows per block, warp w takes rows
// w + 8 i; the pairs are (i, i + 1) = rows 8 apart, the next pair's loads are in flight while this pair's rounds run.
__global__ void __launch_bounds__(256, 2) head_prep_v5_kernel(const __grid_constant__ HeadPrepParams p, __half* hin, int ldh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool hc = lane < 24;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    HeadPrepLane L;
    hp_init(L, p, lane);
    const int c0 = lane * 4, c1 = 128 + lane * 4;
    const int row_first = blockIdx.x * 64 + warp;
    float4 na0[2], na1[2], ncc[2]; int nev[2] = {0, 0};
    auto load_pair = [&](int k) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int row = row_first + 16 * k + 8 * r;
            if (row < p.M) {
                na0[r] = hp_ld4(p.x + xblk_index(row, c0)); na1[r] = hp_ld4(p.x + xblk_index(row, c1));
                ncc[r] = hc ? hp_ld4(p.tok_feat + (size_t)row * p.ldt + c0) : z4;
                nev[r] = p.row_event[row];
            } else { na0[r] = na1[r] = ncc[r] = z4; }
        }
    };
    load_pair(0);
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const int row = row_first + 16 * k;
        if (row >= p.M) break;
        float4 a0[2] = {na0[0], na0[1]}, a1[2] = {na1[0], na1[1]}, cc[2] = {ncc[0], ncc[1]};
        const int ev[2] = {nev[0], nev[1]};
        if (k + 1 < 4) load_pair(k + 1);
        SRHEP_CHECK(row < p.ext.rows_cap && ev[0] >= 0 && ev[0] < p.ext.n_events);
        hp_pair(L, p, lane, row + 8 < p.M, ev, a0, a1, cc, [&](int r, int piece, float4 v) {
            *reinterpret_cast<uint2*>(hin + (size_t)(row + 8 * r) * ldh + hp_piece_col(piece, lane)) = hp_pack4(v);
        });
    }
}

// ------------------------------------------------------------------------------------------------------------------------
constexpr int kHFThreads = 448;
constexpr int kHFPrepWarp0 = 6, kHFPrepWarps = 8;
constexpr int kHFSlots = 5;
constexpr uint32_t kHFSlotBytes = 8192;                              // W1 (k-block, 64-row half): 64 rows x 128 B
constexpr uint32_t kHFABytes = 131072;                               // 8 k-blocks x (128 rows x 128 B)
constexpr uint32_t kHFOffRing = kHFABytes;
constexpr uint32_t kHFOffA2 = kHFOffRing + kHFSlots * kHFSlotBytes;  // 2 k-blocks; A3 (1 k-block) reuses its first half once G2 has retired
constexpr uint32_t kHFOffW2 = kHFOffA2 + 32768;
constexpr uint32_t kHFOffW3 = kHFOffW2 + 16384;
constexpr uint32_t kHFOffBars = kHFOffW3 + 4096;
constexpr uint32_t kHFOffFn = kHFOffBars + 256;                      // final_norm weight | bias: 2 x 256 floats
constexpr uint32_t kHFOffB1 = kHFOffFn + 2048;                       // b1 (128 floats) for E1's rolled loops
constexpr size_t kHFSmemBytes = kHFOffB1 + 512;
static_assert(kHFOffA2 % 1024 == 0 && kHFSmemBytes <= 232448, "head_fused shared-memory layout");

struct HeadFusedParams {
    HeadChainParams c;           // M, weights, biases, stage (tensor maps unused)
    HeadPrepParams q;            // residual, cond_feat, norms, adaLN rows, context (final_tap unused)
    int diag;                    // SRHEP_HF_DIAG, results wrong on purpose: bit 0 = the preparation warps do no work, bit 1 = no W1 streaming / first GEMM
};

__global__ void __launch_bounds__(kHFThreads, 1) head_fused_kernel(const __grid_constant__ HeadFusedParams pp) {
    const HeadChainParams& p = pp.c;
    extern __shared__ __align__(1024) uint8_t hf_smem[];
    uint8_t* smem = hf_smem;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* s_ring = smem + kHFOffRing;
    uint8_t* s_a2 = smem + kHFOffA2;
    uint8_t* s_a3 = s_a2;
    uint8_t* s_w2 = smem + kHFOffW2;
    uint8_t* s_w3 = smem + kHFOffW3;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHFOffBars);
    uint64_t* full = bars;               // [5] producer -> MMA
    uint64_t* empty = bars + 5;          // [5] MMA -> producer
    uint64_t* w23_full = bars + 10;
    uint64_t* acc1_full = bars + 11;     // [2] MMA -> epilogue
    uint64_t* acc1_empty = bars + 13;    // [2] epilogue -> MMA
    uint64_t* a2_ready = bars + 15;      // epilogue -> MMA
    uint64_t* acc2_full = bars + 16;
    uint64_t* a3_ready = bars + 17;
    uint64_t* acc3_full = bars + 18;
    uint64_t* a_ready = bars + 19;       // preparation warps -> MMA: the tile's operand rows are in shared memory
    uint64_t* a_free = bars + 20;        // MMA -> preparation warps: the tile's first GEMM has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + 127) / 128;
    constexpr uint32_t kTmemCols = 512;
    constexpr uint32_t kColAcc2 = 256, kColAcc3 = 320;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kHFSlots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(w23_full, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&acc1_full[i], 1); mbar_init(&acc1_empty[i], 4); }
        mbar_init(a2_ready, 4); mbar_init(acc2_full, 1); mbar_init(a3_ready, 4); mbar_init(acc3_full, 1);
        mbar_init(a_ready, kHFPrepWarps); mbar_init(a_free, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + 128) {                // final_norm weight | bias -> shared memory for the preparation warps
        const int i = threadIdx.x - 64;
        reinterpret_cast<float4*>(smem + kHFOffFn)[i] = hp_ld4((i < 64 ? pp.q.fn_w : pp.q.fn_b - 256) + i * 4);
        reinterpret_cast<float*>(smem + kHFOffB1)[i] = p.b1[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(w23_full, 16384 + 4096);
            bulk_load(s_w2, p.w2, 16384, w23_full);
            bulk_load(s_w3, p.w3, 4096, w23_full);
            // The preparation warps keep one pair of rows of look-ahead in registers, which covers an L2 hit, not an HBM miss: the
            // tile's residual block (128 KB contiguous in the blocked layout) and its cond_feat rows are pulled into L2 one tile ahead.
            auto prefetch_tile = [&](int t) {
                const HeadPrepParams& q = pp.q;
                const float* xt = q.x + (size_t)t * 128 * 256;
#pragma unroll
                for (int i = 0; i < 4; ++i) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xt + i * 8192), "r"(32768u) : "memory");
                const int rows = min(128, q.M - t * 128);
                const float* ft = q.tok_feat + (size_t)t * 128 * q.ldt;                                          // row by row: only the cond_feat columns (384 of a row's 640 bytes) are read
                for (int r = 0; r < rows; ++r) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ft + (size_t)r * q.ldt), "r"(384u) : "memory");
            };
            if ((int)blockIdx.x < m_tiles) prefetch_tile(blockIdx.x);
            uint32_t it = 0;
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x) {
                if (t + (int)gridDim.x < m_tiles) prefetch_tile(t + gridDim.x);
                for (int sl = 0; sl < 16; ++sl, ++it) {                // (k-block, half) = (sl >> 1, sl & 1): contiguous in the W1 image
                    const uint32_t s = it % kHFSlots, ph = (it / kHFSlots) & 1;
                    if (pp.diag & 2) continue;
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], kHFSlotBytes);
                    bulk_load(s_ring + s * kHFSlotBytes, p.w1 + (size_t)sl * kHFSlotBytes, kHFSlotBytes, &full[s]);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc1 = umma_idesc_16(128, 64, 1), idesc2 = umma_idesc_16(128, kHeadH2, 1), idesc3 = umma_idesc_16(128, kHeadH3, 1);
        uint32_t it = 0, j = 0;
        mbar_wait(w23_full, 0);
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++j) {
            const uint32_t buf = j & 1, use = j >> 1;
            mbar_wait(&acc1_empty[buf], (use & 1) ^ 1);
            mbar_wait(a_ready, j & 1);
            tc_fence_after();
            for (int sl = 0; sl < 16; ++sl, ++it) {
                const uint32_t s = it % kHFSlots, ph = (it / kHFSlots) & 1;
                if (!(pp.diag & 2)) mbar_wait(&full[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + (sl >> 1) * 16384), b_addr = smem_u32(s_ring + s * kHFSlotBytes);
                    if (!(pp.diag & 2)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + buf * 128 + (sl & 1) * 64, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc1, (uint32_t)(((sl >> 1) | k) != 0));
                    tc_commit(&empty[s]);
                    }
                    if (sl == 15) { tc_commit(a_free); tc_commit(&acc1_full[buf]); }
                }
                __syncwarp();
            }
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
    } else if (warp < kHFPrepWarp0) {
        const int q = warp & 3;
        const int rt = q * 32 + lane;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const StageParams st = load_stage(p.stage);
        uint32_t j = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++j) {
            const int row = t * 128 + rt;
            const bool valid = row < p.M;
            SRHEP_CHECK(p.M <= p.ext.rows_cap);
            const uint32_t buf = j & 1;
            // ---------------------------------------------------------------- E1: LeakyReLU(h1 + b1) -> LayerNorm(128) -> A2.  Two reads of the
            // accumulator (sum and sum of squares, then the output) instead of 128 live registers: the CTA carries 14 warps
            {
                mbar_wait(&acc1_full[buf], (j >> 1) & 1);
                tc_fence_after();
                const float* sb1 = reinterpret_cast<const float*>(smem + kHFOffB1);
                float s = 0.f, qq = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_lane + buf * 128 + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float y = leaky_relu(__uint_as_float(r[i]) + sb1[c * 32 + i]); s += y; qq = fmaf(y, y, qq); }
                }
                const float mean = s * (1.0f / 128.f);
                const float rstd = rsqrtf(fmaxf(qq * (1.0f / 128.f) - mean * mean, 0.f) + kLnEps);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_lane + buf * 128 + c * 32, r);
                    tmem_ld_wait();
                    float w[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) w[i] = (leaky_relu(__uint_as_float(r[i]) + sb1[c * 32 + i]) - mean) * rstd;
                    chain_store_a(smem_u32(s_a2), rt, c * 32, w, 1);
                }
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) { mbar_arrive(&acc1_empty[buf]); mbar_arrive(a2_ready); }
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
                    chain_store_a(smem_u32(s_a3), rt, c * 32, w, 1);          // A3 overwrites A2: acc2_full says G2 has read it
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
    } else {
        // ---------------------------------------------------------------- head preparation: rows pw + 8 i of the tile, pairs (i, i + 1)
        const HeadPrepParams& q = pp.q;
        const int pw = warp - kHFPrepWarp0;
        const bool hc = lane < 24;
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const int c0 = lane * 4, c1 = 128 + lane * 4;
        const uint32_t a_sh = smem_u32(smem);
        HeadPrepLane L;
        hp_init(L, q, lane, false);
        const uint32_t fn_sh = a_sh + kHFOffFn;
        // a warp's rows are pw + 8 i: row & 7 = pw & 7 for all of them, so the swizzled position of a lane's piece inside its row never changes
        uint32_t poff[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) { const int col = hp_piece_col(i, lane); poff[i] = (uint32_t)((col >> 6) * 16384 + ((((col & 63) >> 3) ^ (pw & 7)) << 4) + (col & 7) * 2); }
        float4 na0[2], na1[2], ncc[2]; int nev[2] = {0, 0};
        auto load_pair = [&](int t, int k) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = t * 128 + pw + 16 * k + 8 * r;
                if (pp.diag & 4) { na0[r] = na1[r] = make_float4((float)lane, 1.f, 2.f, (float)row); ncc[r] = z4; nev[r] = 0; }      // no row loads
                else if (row < q.M) {
                    na0[r] = hp_ld4(q.x + xblk_index(row, c0)); na1[r] = hp_ld4(q.x + xblk_index(row, c1));
                    ncc[r] = hc ? hp_ld4(q.tok_feat + (size_t)row * q.ldt + c0) : z4;
                    nev[r] = q.row_event[row];
                } else { na0[r] = na1[r] = ncc[r] = z4; }
            }
        };
        uint32_t j = 0;
        if ((int)blockIdx.x < m_tiles) load_pair(blockIdx.x, 0);
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++j) {
            bool wait_free = j > 0;                                    // the previous tile's first GEMM still reads the operand buffer
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const int rt = pw + 16 * k, row = t * 128 + rt;
                float4 a0[2] = {na0[0], na0[1]}, a1[2] = {na1[0], na1[1]}, cc[2] = {ncc[0], ncc[1]};
                const int ev[2] = {nev[0], nev[1]};
                if (!(pp.diag & 1)) {
                if (k + 1 < 8) load_pair(t, k + 1);
                else if (t + (int)gridDim.x < m_tiles) load_pair(t + gridDim.x, 0);
                }
                if (row >= q.M || (pp.diag & 1)) continue;             // warp-uniform; rows past the end keep whatever the buffer holds (row-local, never stored)
                SRHEP_CHECK(row < q.ext.rows_cap && ev[0] >= 0 && ev[0] < q.ext.n_events);
                hp_pair(L, q, lane, row + 8 < q.M, ev, a0, a1, cc, [&](int r, int piece, float4 v) {
                    if (wait_free) { mbar_wait(a_free, (j - 1) & 1); wait_free = false; }
                    const uint2 w = hp_pack4(v);
                    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a_sh + poff[piece] + (uint32_t)((rt + 8 * r) * 128)), "r"(w.x), "r"(w.y) : "memory");
                }, fn_sh);
            }
            if (wait_free) mbar_wait(a_free, (j - 1) & 1);             // a warp without rows in this tile still keeps its phase count
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_ready);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace srhep

// Fused DiT-layer chain for sm_100a: everything between two attention calls that is local to a
// cell (row), in ONE kernel per layer (models/diffusion_transformer.py:30-53, models/dense.py:49-83,
// models/attention.py:112-123):
//
//   stage 0  out-projection of the attention output      x1 = x + gate_msa * (A Wo^T + bo)
//            LN2(x1) * (1 + scale_mlp) + shift_mlp, then the Dense's own non-affine LN  -> A
//   stage 1  MLP first Linear + LeakyReLU                                              -> A
//   stage 2  MLP second Linear + LeakyReLU              x2 = x1 + gate_mlp * leaky(A W2^T + b2)
//            LN1_{l+1}(x2) * (1 + scale_msa) + shift_msa  (next layer's attention input) -> A
//   stage 3-5  Q, K, V projections of the NEXT layer                                   -> global (16 bit)
//
// A CTA owns one 128-row tile at a time.  The A operand (128 x 256, 16 bit, 64 KB, K-major with
// 128-byte swizzle) is loaded once by TMA (the attention output) and then rewritten IN PLACE by the
// epilogue of each stage; the six 256 x 256 weight matrices are streamed from L2 through a ring of
// 16 KB slots (128 output rows x 64 k) by cp.async.bulk; the fp32 accumulator (128 lanes x 256
// columns of TMEM) doubles as the parking place of the finished residual row between the
// LayerNorm passes.  The activations of the tile never leave the SM between stages: per cell
// and layer the kernel reads 512 B (attention output) + 1 KB (residual), writes 1 KB (residual) +
// 1.5 KB (q|k|v) for 786 432 FLOP.  Two CTAs share an SM (2 x 112 KB of shared memory, 2 x 256
// TMEM columns) so that one CTA's epilogue runs under the other's MMAs.
//   warp 0: TMA producer   warp 1: TMEM allocator + MMA issuer   warps 2-9: epilogue (TMEM lane
//   quarter = warp & 3, column half = (warp - 2) >> 2; one thread = half a row)
#pragma once
#include "kernels_bf16.cuh"

namespace srhep {

constexpr int kChainThreads = 320;
constexpr int kChainSlots = 3;
constexpr uint32_t kChainSlotBytes = 16384;          // 128 weight rows x 128 B
constexpr uint32_t kChainABytes = 65536;             // 4 k-blocks x (128 rows x 128 B)
constexpr size_t kChainSmemBytes = kChainABytes + kChainSlots * kChainSlotBytes + 256;
// split (fp32-grade) mode: every 16-bit operand is a pair of fp16 planes (hi, lo), each product runs as hi.hi + hi.lo + lo.hi
// on the same fp32 accumulator (kernels_attn3.cuh: split16); the A buffer and the weight slots double, one CTA per SM
constexpr size_t kChainSmemBytesSplit = 2 * kChainABytes + kChainSlots * 2 * kChainSlotBytes + 256;
constexpr int kChainH = 256;
constexpr int kChainStageEv = 16;                    // events per tile whose adaLN rows are staged in shared memory (3 KB each, after the 4 KB statistics scratch)

struct ChainParams {
    int M;                       // rows of this pass
    int n_stages;                // 6, or 3 for the last layer (no next attention)
    int fp16;                    // 16-bit operand format: 0 = bf16, 1 = fp16
    const int* row_event;        // [M] global event id of each row
    float* x;                    // fp32 residual stream in the BLOCKED layout (common.cuh: xblk_index), updated in place
    const uint8_t* w[6];         // pre-swizzled weight images [4 k-blocks][256 rows x 128 B]
    const uint8_t* w_lo[6];      // split mode: the low planes of the same weights
    float wscale[6];             // split mode: the images hold W * 2^s (s per matrix, so that the low plane stays out of the fp16 subnormals); this is 2^-s, applied to the accumulator
    float cst[6][256];           // per-column constants BY VALUE (constant bank, no LSU traffic): biases of stages 0-5
    const float* gate_msa;       // per-event rows (stride ld_mod floats)
    const float* gate_mlp;
    // LayerNorm affine and adaLN modulation combined per event by modpq_kernel:  (LN(x) w + b)(1 + scale) + shift = LN(x) P + Q  with
    // P = w (1 + scale), Q = b (1 + scale) + shift  (stride ld_pq floats): the LayerNorm-modulate passes run on packed f32x2 instructions only
    const float* p_mlp; const float* q_mlp;               // norm2 / (scale_mlp, shift_mlp) of this layer
    const float* p_nxt; const float* q_nxt;               // norm1 / (scale_msa, shift_msa) of the next layer
    int ld_mod, ld_pq;
    const float* row_bias; int ld_row_bias;   // first-layer mode only: per-event bias rows of feat_0 (its context part)
    void* qkv;                   // [M, 768] 16-bit q|k|v of the next layer
    void* qkv_lo;                // split mode: its low plane
    long long* dbg;              // optional timeline of CTA 0 (clock64 stamps, 32 per tile, first 8 tiles); null in production
    Extents ext;                 // checked in the -DSRHEP_BOUNDS build only
    int a_early;                 // release the A buffer k-block by k-block (1 in production; 0 = after the tile's last MMA, for A/B runs)
    int ln_direct;               // stage 2: LayerNorm output straight to A (1 in production; 0 = parked in TMEM and copied in a third pass, for A/B runs)
    int a_pf;                    // weight-slot index of a tile at which the producer pulls the NEXT tile's A operand into L2 (-1 = never).  The L2 turns
                                 // over every ~70 k cycles under this kernel's traffic: a whole tile period ahead is too early, the last stages are not
};

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// four consecutive values of a per-event adaLN row: from the copy staged in shared memory (the usual case) or from global memory
__device__ __forceinline__ float4 par_f4(bool staged, uint32_t sm_addr, const float* g) { return staged ? lds_f4(sm_addr) : ldg128_stream(g); }

__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float sum_f32x2(uint64_t v) { return f32x2_lo(v) + f32x2_hi(v); }

// acc + bias (+ LeakyReLU), gated into the residual: returns the new residual chunk in r[] (as bits).  Everything whose operands
// already sit in registers runs as packed f32x2 instructions (two columns per FFMA2 / FADD2): at the power cap the epilogue's
// instruction count is what the chain pays for.  s1 / s2 are PAIRS of partial sums (even / odd columns).
template <bool kAct, bool kScaled = false>
__device__ __forceinline__ void chain_resid_chunk(uint32_t (&r)[32], const float (&xr)[32], const float* bias /*constant bank*/,
                                                  bool staged, uint32_t gate_sm, const float* __restrict__ gate, uint64_t& s1, uint64_t& s2, float ws = 1.f) {
    const uint64_t slope2 = pack_f32x2(kLeaky, kLeaky);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 g4 = par_f4(staged, gate_sm + j * 4, gate + j);
        const uint64_t gg[2] = {pack_f32x2(g4.x, g4.y), pack_f32x2(g4.z, g4.w)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float w0, w1;
            if (kScaled) { w0 = fmaf(__uint_as_float(r[j + 2 * u]), ws, bias[j + 2 * u]); w1 = fmaf(__uint_as_float(r[j + 2 * u + 1]), ws, bias[j + 2 * u + 1]); }
            else { w0 = __uint_as_float(r[j + 2 * u]) + bias[j + 2 * u]; w1 = __uint_as_float(r[j + 2 * u + 1]) + bias[j + 2 * u + 1]; }
            uint64_t w = pack_f32x2(w0, w1);
            if (kAct) { const uint64_t lk = fmul2(w, slope2); w = pack_f32x2(fmaxf(w0, f32x2_lo(lk)), fmaxf(w1, f32x2_hi(lk))); }   // LeakyReLU = max(x, 0.01 x)
            w = ffma2(gg[u], w, pack_f32x2(xr[j + 2 * u], xr[j + 2 * u + 1]));
            s1 = fadd2(s1, w); s2 = ffma2(w, w, s2);
            r[j + 2 * u] = (uint32_t)w; r[j + 2 * u + 1] = (uint32_t)(w >> 32);
        }
    }
}

// first-layer mode: leaky(acc + bias + per-event bias) IS the residual row (feat_0, models/flow_model.py:224-228)
template <bool kScaled = false>
__device__ __forceinline__ void chain_first_chunk(uint32_t (&r)[32], const float* /*bias: zero in this mode*/, bool staged, uint32_t rb_sm, const float* __restrict__ rb,
                                                  uint64_t& s1, uint64_t& s2, float ws = 1.f) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 b4 = par_f4(staged, rb_sm + j * 4, rb + j);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};            // feat_0's own bias is inside the per-event rows (gemm of the context part): no per-column constant to add
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
            const float w0 = leaky_relu(kScaled ? fmaf(__uint_as_float(r[j + u]), ws, bb[u]) : __uint_as_float(r[j + u]) + bb[u]);
            const float w1 = leaky_relu(kScaled ? fmaf(__uint_as_float(r[j + u + 1]), ws, bb[u + 1]) : __uint_as_float(r[j + u + 1]) + bb[u + 1]);
            const uint64_t w = pack_f32x2(w0, w1);
            s1 = fadd2(s1, w); s2 = ffma2(w, w, s2);
            r[j + u] = __float_as_uint(w0); r[j + u + 1] = __float_as_uint(w1);
        }
    }
}

// r[] (bits of the residual row chunk) -> LN(r) P + Q in place  (= (LN(r) w + b)(1 + scale) + shift, see ChainParams); accumulates sum /
// sum of squares of the result as PAIRS of partial sums.  rs2 = (rstd, rstd), nm2 = (-mean rstd, -mean rstd).  Everything is packed
// f32x2: 2.5 instructions per column (it was 5.5 with the LayerNorm weights as scalar constant-bank operands).
__device__ __forceinline__ void chain_ln_mod_chunk(uint32_t (&r)[32], uint64_t rs2, uint64_t nm2, bool staged, uint32_t p_sm, uint32_t q_sm,
                                                   const float* __restrict__ pg, const float* __restrict__ qg, uint64_t& t1, uint64_t& t2) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        const float4 p4 = par_f4(staged, p_sm + j * 4, pg + j), q4 = par_f4(staged, q_sm + j * 4, qg + j);
        const uint64_t pp[2] = {pack_f32x2(p4.x, p4.y), pack_f32x2(p4.z, p4.w)}, qq[2] = {pack_f32x2(q4.x, q4.y), pack_f32x2(q4.z, q4.w)};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const uint64_t xh = ffma2(pack_f32x2(__uint_as_float(r[j + 2 * u]), __uint_as_float(r[j + 2 * u + 1])), rs2, nm2);
            const uint64_t y = ffma2(xh, pp[u], qq[u]);
            t1 = fadd2(t1, y); t2 = ffma2(y, y, t2);
            r[j + 2 * u] = (uint32_t)y; r[j + 2 * u + 1] = (uint32_t)(y >> 32);
        }
    }
}

// 32 consecutive values of row `rt` -> 16-bit, into the K-major 128B-swizzled A buffer (32-bit shared address
// `a_addr`) at columns [col0, col0 + 32)
__device__ __forceinline__ void chain_store_a(uint32_t a_addr, int rt, int col0, const float (&v)[32], int fp16) {
    const uint32_t arow = a_addr + (uint32_t)((col0 >> 6) * 16384 + rt * 128);
    const int cb = (col0 & 63) >> 3;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(arow + (uint32_t)(((cb + g) ^ (rt & 7)) << 4)),
                     "r"(pack16(v[8 * g], v[8 * g + 1], fp16)), "r"(pack16(v[8 * g + 2], v[8 * g + 3], fp16)),
                     "r"(pack16(v[8 * g + 4], v[8 * g + 5], fp16)), "r"(pack16(v[8 * g + 6], v[8 * g + 7], fp16)) : "memory");
}
// the same for the split mode: hi plane at a_addr, lo plane kChainABytes later
__device__ __forceinline__ void chain_store_a_split(uint32_t a_addr, int rt, int col0, const float (&v)[32]) {
    const uint32_t arow = a_addr + (uint32_t)((col0 >> 6) * 16384 + rt * 128);
    const int cb = (col0 & 63) >> 3;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) split16(v[8 * g + 2 * u], v[8 * g + 2 * u + 1], hi[u], lo[u]);
        const uint32_t a = arow + (uint32_t)(((cb + g) ^ (rt & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + kChainABytes), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
}
// 16 already packed pairs (32 consecutive columns of row `rt`) -> the A buffer
__device__ __forceinline__ void chain_store_a_packed(uint32_t a_addr, int rt, int col0, const uint32_t* pk) {
    const uint32_t arow = a_addr + (uint32_t)((col0 >> 6) * 16384 + rt * 128);
    const int cb = (col0 & 63) >> 3;
#pragma unroll
    for (int g = 0; g < 4; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(arow + (uint32_t)(((cb + g) ^ (rt & 7)) << 4)),
                     "r"(pk[4 * g]), "r"(pk[4 * g + 1]), "r"(pk[4 * g + 2]), "r"(pk[4 * g + 3]) : "memory");
}
__device__ __forceinline__ void sts_f2(uint32_t addr, float a, float b) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory"); }
__device__ __forceinline__ float2 lds_f2(uint32_t addr) { float2 v; asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory"); return v; }

// Row-per-lane registers -> line-per-4-lanes registers.  In: a[32] = this lane's 128 bytes (one full line: 64 16-bit columns
// of its row) as four 32-byte pieces.  Out: a[8 i .. 8 i + 7] = piece (lane & 3) of row (lane & ~3) + i, so that store i of the
// warp writes 8 complete 128-byte lines (4 lanes each) instead of touching 32 different lines: a quarter of the LSU wavefronts.
// Two butterfly rounds over the 4 lanes of a group (xor 2, xor 1); each shuffle swaps one register between the two partners.
__device__ __forceinline__ void transpose_line_pieces(uint32_t (&a)[32], int lane) {
    const bool b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const uint32_t send = b1 ? a[r] : a[16 + r];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 2);
        a[r] = b1 ? recv : a[r];
        a[16 + r] = b1 ? a[16 + r] : recv;
    }
#pragma unroll
    for (int h = 0; h < 32; h += 16) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint32_t send = b0 ? a[h + r] : a[h + 8 + r];
            const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
            a[h + r] = b0 ? recv : a[h + r];
            a[h + 8 + r] = b0 ? a[h + 8 + r] : recv;
        }
    }
}

// P = w (1 + scale), Q = b (1 + scale) + shift for LayerNorm 1 and 2 of every layer, per event: one block per event.  pq row of an
// event: [layer][P_msa | Q_msa | P_mlp | Q_mlp][256]; mod row: [layer][shift_msa | scale_msa | gate_msa | shift_mlp | scale_mlp | gate_mlp][256];
// nrm: per layer (stride nrm_stride) norm1.w | norm1.b | norm2.w | norm2.b (models/diffusion_transformer.py:8-9,38-52).
struct ModPqParams { const float* mod; int ld_mod; const float* nrm; int nrm_stride; float* pq; int ld_pq; int layers; int e0; };
__global__ void __launch_bounds__(256) modpq_kernel(ModPqParams p) {
    const int ev = p.e0 + blockIdx.x;
    const float* m = p.mod + (size_t)ev * p.ld_mod;
    float* o = p.pq + (size_t)ev * p.ld_pq;
    for (int i = threadIdx.x; i < p.layers * 2 * (kChainH / 4); i += blockDim.x) {       // four columns per thread and step (all rows are 16-byte aligned)
        const int l = i / (2 * (kChainH / 4)), w = (i / (kChainH / 4)) & 1, c = (i % (kChainH / 4)) * 4;      // w: 0 = msa (norm1), 1 = mlp (norm2)
        const float* ml = m + (size_t)l * 6 * kChainH + w * 3 * kChainH;
        const float* nl = p.nrm + (size_t)l * p.nrm_stride + w * 2 * kChainH;
        const float4 sh = *reinterpret_cast<const float4*>(ml + c), sc = *reinterpret_cast<const float4*>(ml + kChainH + c);
        const float4 nw = *reinterpret_cast<const float4*>(nl + c), nb = *reinterpret_cast<const float4*>(nl + kChainH + c);
        const float4 s1 = make_float4(1.f + sc.x, 1.f + sc.y, 1.f + sc.z, 1.f + sc.w);
        float* ol = o + (size_t)l * 4 * kChainH + w * 2 * kChainH + c;
        *reinterpret_cast<float4*>(ol) = make_float4(nw.x * s1.x, nw.y * s1.y, nw.z * s1.z, nw.w * s1.w);
        *reinterpret_cast<float4*>(ol + kChainH) = make_float4(fmaf(nb.x, s1.x, sh.x), fmaf(nb.y, s1.y, sh.y), fmaf(nb.z, s1.z, sh.z), fmaf(nb.w, s1.w, sh.w));
    }
}

#ifdef SRHEP_TIMELINE      // build with -DSRHEP_TIMELINE to record the CTA timeline (tools/chain_dbg.py); costs registers, off in production
#define CHAIN_STAMP(tile, k) do { if (p.dbg && blockIdx.x == 0 && (tile) < 8) p.dbg[(tile) * 32 + (k)] = clock64(); } while (0)
#else
#define CHAIN_STAMP(tile, k) do { } while (0)
#endif

template <bool kFp16, bool kFirst = false, bool kSplit = false>
__global__ void __launch_bounds__(kChainThreads, kSplit ? 1 : 2) layer_chain_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a_lo,
                                                                                    const __grid_constant__ ChainParams p) {
    static_assert(!kSplit || kFp16, "the split (fp32-grade) mode runs on fp16 planes");
    constexpr uint32_t kABytes = kSplit ? 2 * kChainABytes : kChainABytes;          // hi [| lo]
    constexpr uint32_t kSlot = kSplit ? 2 * kChainSlotBytes : kChainSlotBytes;      // hi [| lo]
#ifdef SRHEP_NO_DEFER_X_STORE
    constexpr bool kDefer = false;
#else
    constexpr bool kDefer = true;       // the new residual row is stored by the pass that re-reads it from TMEM, so that the first pass only LOADS from global memory
#endif
    extern __shared__ __align__(1024) uint8_t chain_smem[];      // no static smem in this kernel: the dynamic window starts 1024-aligned
    uint8_t* s_a = chain_smem;
    if ((smem_u32(s_a) & 1023u) != 0) __trap();
    uint8_t* s_w = s_a + kABytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + kChainSlots * kSlot);
    uint64_t* a_full = bars;             // TMA -> MMA      attention-output tile landed
    uint64_t* a_free = bars + 14;        // [4] MMA -> TMA  the tile's last MMAs over k-block kb of A retired: that k-block may be overwritten (the next
                                         //     tile's A is requested k-block by k-block while the last stage still runs, not after it)
    uint64_t* w_full = bars + 2;         // [slots] TMA -> MMA
    uint64_t* w_empty = bars + 5;        // [slots] MMA -> TMA
    uint64_t* acc_full = bars + 8;       // [2] MMA -> epilogue  one 128-column half of a stage's accumulator complete
    uint64_t* epi_done = bars + 10;      // [2] epilogue -> MMA  the 4 warps of a column half are done with their half of TMEM
    uint64_t* a_written = bars + 12;     // epilogue -> MMA  all 8 warps rewrote the A operand (stages that feed another GEMM)
    uint64_t* par_full = bars + 13;      // bulk copies of the tile's per-event adaLN rows landed in the (dead) A buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.M + 127) / 128;
    constexpr uint32_t kTmemCols = 256;
    // first-layer mode (kFirst): the stages are feat_0 (K = 192: three k-blocks; epilogue = stage 2's without a residual to read), q, k, v
    constexpr int kKb0 = kFirst ? 3 : 4;                 // k-blocks of stage 0
    const int slots_per_tile = p.n_stages * 8 - (kFirst ? 2 : 0);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        if (kSplit) prefetch_tmap(&tmap_a_lo);
        mbar_init(a_full, 1); for (int i = 0; i < 4; ++i) mbar_init(&a_free[i], 1);
        for (int i = 0; i < kChainSlots; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&epi_done[i], 4); }
        mbar_init(a_written, 8); mbar_init(par_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t slot_it = 0, tile_i = 0;
            for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++tile_i) {
                for (int j = 0; j < slots_per_tile; ++j, ++slot_it) {
                    if (j == (tile_i == 0 ? 0 : 2)) {                 // the A tile: first thing of the kernel, else after two weight slots of run-ahead
                        CHAIN_STAMP(tile_i, 0);
                        mbar_expect_tx(a_full, kKb0 * 16384 * (kSplit ? 2 : 1));
#pragma unroll
                        for (int kb = 0; kb < kKb0; ++kb) {
                            if (tile_i > 0) mbar_wait(&a_free[p.a_early ? kb : 3], (tile_i - 1) & 1);
                            tma_load_2d(s_a + kb * 16384, &tmap_a, a_full, kb * 64, t * 128);
                            if (kSplit) tma_load_2d(s_a + kChainABytes + kb * 16384, &tmap_a_lo, a_full, kb * 64, t * 128);
                        }
                        CHAIN_STAMP(tile_i, 1);
                        if (!kFirst) {
                        // the fp32 residual tile of THIS tile (128 KB contiguous in the blocked layout) is first needed by the
                        // stage-0 epilogue, several microseconds from now: pull it into L2 meanwhile
                        const float* xt = p.x + (size_t)t * 128 * kChainH;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(xt + i * 8192), "r"(32768u) : "memory");
                        }
                    }
                    if (j == p.a_pf && t + (int)gridDim.x < m_tiles) {        // the next tile's A operand (attention output, 64 KB): towards L2 now, into shared memory when this tile's last MMAs retire
#pragma unroll
                        for (int kb2 = 0; kb2 < kKb0; ++kb2)
                            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&tmap_a), "r"(kb2 * 64), "r"((t + (int)gridDim.x) * 128) : "memory");
                    }
                    int g, nh, kb;                                            // column half outer: the two halves of the accumulator are a double buffer
                    if (kFirst && j < 6) { g = 0; nh = j / 3; kb = j % 3; }
                    else { const int jj = kFirst ? j + 2 : j; g = jj >> 3; nh = (jj >> 2) & 1; kb = jj & 3; }
                    const uint32_t s = slot_it % kChainSlots, ph = (slot_it / kChainSlots) & 1;
                    mbar_wait(&w_empty[s], ph ^ 1);
                    mbar_expect_tx(&w_full[s], kSlot);
                    bulk_load(s_w + s * kSlot, p.w[g] + (size_t)(kb * 256 + nh * 128) * 128, kChainSlotBytes, &w_full[s]);
                    if (kSplit) bulk_load(s_w + s * kSlot + kChainSlotBytes, p.w_lo[g] + (size_t)(kb * 256 + nh * 128) * 128, kChainSlotBytes, &w_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = umma_idesc_16(128, 128, kFp16 ? 1 : 0);
        uint32_t slot_it = 0, stage_it = 0, tile_i = 0, aw_it = 0;
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++tile_i) {
            for (int g = 0; g < p.n_stages; ++g, ++stage_it) {
                for (int nh = 0; nh < 2; ++nh) {
                    // this half of the accumulator was drained by the previous stage's epilogue (its 4 warps)
                    if (stage_it > 0) mbar_wait(&epi_done[nh], (stage_it - 1) & 1);
                    if (nh == 0) {
                        if (g == 0) { if (lane == 0) CHAIN_STAMP(tile_i, 2); mbar_wait(a_full, tile_i & 1); if (lane == 0) CHAIN_STAMP(tile_i, 3); }
                        else if (kFirst ? g == 1 : g <= 3) { mbar_wait(a_written, aw_it & 1); ++aw_it; }      // stages 1-3 (first-layer mode: stage 1) read the A operand the previous epilogue wrote
                    }
                    tc_fence_after();
                    const int kbn = (kFirst && g == 0) ? kKb0 : 4;
                    for (int kb = 0; kb < kbn; ++kb, ++slot_it) {
                        const uint32_t s = slot_it % kChainSlots, ph = (slot_it / kChainSlots) & 1;
                        mbar_wait(&w_full[s], ph);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t a_addr = smem_u32(s_a + kb * 16384), b_addr = smem_u32(s_w + s * kSlot);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                umma_bf16(tmem_base + nh * 128, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                                          (uint32_t)((kb | k) != 0));
                                if (kSplit) {      // hi.lo and lo.hi on the same accumulator
                                    umma_bf16(tmem_base + nh * 128, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + kChainSlotBytes + k * 32), idesc, 1u);
                                    umma_bf16(tmem_base + nh * 128, umma_desc_sw128(a_addr + kChainABytes + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, 1u);
                                }
                            }
                            tc_commit(&w_empty[s]);
                            if (nh == 1 && g == p.n_stages - 1) tc_commit(&a_free[kb]);
                            if (kb == kbn - 1) tc_commit(&acc_full[nh]);
                        }
                        __syncwarp();
                    }
                }
                if (lane == 0) CHAIN_STAMP(tile_i, 4 + g);
            }
        }
    } else {
        const int q = warp & 3, hh = (warp - 2) >> 2;
        const int rt = q * 32 + lane;                                  // row inside the tile = TMEM lane
        const uint32_t t_col = tmem_base + ((uint32_t)(q * 32) << 16) + hh * 128;
        const uint32_t a_sh = smem_u32(s_a);                            // 32-bit shared addresses keep the epilogue inside its register budget
        const uint32_t st_own = a_sh + (uint32_t)(hh * 128 + rt) * 8, st_oth = a_sh + (uint32_t)((hh ^ 1) * 128 + rt) * 8;   // LayerNorm partial sums: scratch in the (then dead) A buffer
        constexpr int fp16 = kFp16 ? 1 : 0;                            // compile-time operand format: one conversion per pair, no selects
        uint32_t stage_it = 0, tile_i = 0;
        auto stage_done = [&](bool wrote_a) {
            if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 11 + 2 * (int)(stage_it % (uint32_t)p.n_stages));
            fence_async_smem();                                        // A stores -> visible to the tensor-core (async) proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&epi_done[hh]); if (wrote_a) mbar_arrive(a_written); }
            ++stage_it;
        };
        uint32_t par_it = 0;
        const uint32_t par_sh = a_sh + 4096;                           // staged adaLN rows: [event in tile][3 arrays][256 floats]
        for (int t = blockIdx.x; t < m_tiles; t += gridDim.x, ++tile_i) {
            const int row = t * 128 + rt;
            const bool valid = row < p.M;
            const int ev0 = p.row_event[t * 128];
            const int ne = p.row_event[min(t * 128 + 127, p.M - 1)] - ev0 + 1;
            const bool staged = ne <= kChainStageEv;                   // CTA-uniform
            const int evt = valid ? p.row_event[row] : ev0;
            SRHEP_CHECK(p.M <= p.ext.rows_cap && t * 128 + 127 < ((p.ext.rows_cap + 127) & ~127));      // TMA boxes and the blocked residual tile stay inside the workspace
            SRHEP_CHECK(evt >= 0 && evt < p.ext.n_events && ev0 >= 0 && ev0 + ne <= p.ext.n_events);     // per-event adaLN rows
            const uint32_t xoff = (uint32_t)xblk_index(row, hh * 128);   // 32-bit element offsets keep the epilogue under its register budget
#define xrow (p.x + xoff)                                         /* + 1024 floats per 8-column group (32 B pieces of this row) */
            SRHEP_CHECK(!valid || (size_t)xoff + 15 * 1024 + 8 <= (size_t)((p.ext.rows_cap + 127) & ~127) * kChainH);      // last 32-byte piece of this thread's half row
            const uint32_t eo = (uint32_t)evt * (uint32_t)p.ld_mod + hh * 128;
            const uint32_t eq = (uint32_t)evt * (uint32_t)p.ld_pq + hh * 128;
            const uint32_t psm = par_sh + (uint32_t)(evt - ev0) * 3072 + hh * 512;       // this row's event, this thread's column half, array 0
            constexpr float inv_n = 1.0f / (float)kChainH;
            // One thread copies the tile's per-event rows of three adaLN arrays into the A buffer once every MMA of the stage has
            // retired (the buffer is dead until this stage's last pass rewrites it); everybody then reads them as shared-memory
            // broadcasts instead of paying an L2 round trip per 32-column chunk.
            auto stage_rows = [&](const float* a0, const float* a1, const float* a2, int ld0) {
                if (!staged) return;
                if (warp == 2 && lane == 0) {
                    const int na = 1 + (a1 != nullptr) + (a2 != nullptr);
                    mbar_expect_tx(par_full, (uint32_t)(ne * na) * 1024u);
                    for (int e = 0; e < ne; ++e) {
                        const size_t go = (size_t)(ev0 + e) * p.ld_pq;
                        uint8_t* dst = s_a + 4096 + e * 3072;
                        bulk_load(dst, a0 + (size_t)(ev0 + e) * ld0, 1024, par_full);
                        if (a1) bulk_load(dst + 1024, a1 + go, 1024, par_full);
                        if (a2) bulk_load(dst + 2048, a2 + go, 1024, par_full);
                    }
                }
                mbar_wait(par_full, par_it & 1);
                ++par_it;
            };

            // ---------------------------------------------------------------- stage 0: out-projection, residual, LN2 + modulate + LN
            if constexpr (!kFirst) {
                float xr[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) xr[j] = 0.f;
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) ldg256_stream(xrow + (j >> 3) * 1024, &xr[j]);   // in flight while the MMAs run
                }
                mbar_wait(&acc_full[hh], stage_it & 1);
                mbar_wait(&acc_full[hh ^ 1], stage_it & 1);               // every MMA of the stage retired: the A buffer is free for the scratch and the staged rows
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 10);
                tc_fence_after();
                stage_rows(p.gate_msa, p.p_mlp, p.q_mlp, p.ld_mod);
                uint64_t s1p = 0ull, s2p = 0ull;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_col + c * 32, r);
                    tmem_ld_wait();
                    chain_resid_chunk<false, kSplit>(r, xr, &p.cst[0][hh * 128 + c * 32], staged, psm + c * 128, p.gate_msa + eo + c * 32, s1p, s2p, p.wscale[0]);
                    if (valid) {
                        if (!kDefer) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);
                        }
                        if (c < 3) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) ldg256_stream(xrow + ((c + 1) * 4 + (j >> 3)) * 1024, &xr[j]);
                        }
                    }
                    tmem_st32(t_col + c * 32, r);
                }
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 22);
                tmem_st_wait();
                const float s1 = sum_f32x2(s1p), s2 = sum_f32x2(s2p);
                sts_f2(st_own, s1, s2);
                named_bar_sync(1, 256);
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 23);
                const float2 o1 = lds_f2(st_oth);
                float mean = (s1 + o1.x) * inv_n;
                float rstd = rsqrtf(fmaxf((s2 + o1.y) * inv_n - mean * mean, 0.f) + kLnEps);
                const uint64_t rs2 = pack_f32x2(rstd, rstd), nm2 = pack_f32x2(-mean * rstd, -mean * rstd);
                uint64_t t1p = 0ull, t2p = 0ull;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {                            // affine + adaLN modulate, parked back in TMEM
                    uint32_t r[32];
                    tmem_ld32(t_col + c * 32, r);
                    tmem_ld_wait();
                    if (kDefer && valid) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);      // x1
                    }
                    chain_ln_mod_chunk(r, rs2, nm2, staged, psm + 1024 + c * 128, psm + 2048 + c * 128, p.p_mlp + eq + c * 32, p.q_mlp + eq + c * 32, t1p, t2p);
                    tmem_st32(t_col + c * 32, r);
                }
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 24);
                tmem_st_wait();
                const float t1 = sum_f32x2(t1p), t2 = sum_f32x2(t2p);
                sts_f2(st_own + 2048, t1, t2);
                named_bar_sync(1, 256);
                const float2 o2 = lds_f2(st_oth + 2048);
                mean = (t1 + o2.x) * inv_n;
                rstd = rsqrtf(fmaxf((t2 + o2.y) * inv_n - mean * mean, 0.f) + kLnEps);
                named_bar_sync(1, 256);                                  // every thread has read its partner's sums and the staged rows: the buffer may be overwritten
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 25);
                const uint64_t rs2c = pack_f32x2(rstd, rstd), nm2c = pack_f32x2(-mean * rstd, -mean * rstd);
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {                            // the Dense's own non-affine LayerNorm (models/dense.py:62) -> A
                    uint32_t r[32];
                    tmem_ld32(t_col + c * 32, r);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint64_t xh = ffma2(pack_f32x2(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), rs2c, nm2c);
                        v[j] = f32x2_lo(xh); v[j + 1] = f32x2_hi(xh);
                    }
                    if (kSplit) chain_store_a_split(a_sh, rt, hh * 128 + c * 32, v); else chain_store_a(a_sh, rt, hh * 128 + c * 32, v, fp16);
                }
                stage_done(true);
            }
            // ---------------------------------------------------------------- stage 1: MLP hidden
            if constexpr (!kFirst) {
                mbar_wait(&acc_full[hh], stage_it & 1);
                mbar_wait(&acc_full[hh ^ 1], stage_it & 1);               // A is rewritten in place: every MMA of the stage must have retired
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 12);
                tc_fence_after();
                auto hidden_chunk = [&](const uint32_t (&r)[32], int c) {
                    const float* b = &p.cst[1][hh * 128 + c * 32];
                    float v[32];
                    const uint64_t slope2 = pack_f32x2(kLeaky, kLeaky);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {                       // LeakyReLU = max(x, 0.01 x), the product on a packed instruction
                        const float w0 = kSplit ? fmaf(__uint_as_float(r[j]), p.wscale[1], b[j]) : __uint_as_float(r[j]) + b[j];
                        const float w1 = kSplit ? fmaf(__uint_as_float(r[j + 1]), p.wscale[1], b[j + 1]) : __uint_as_float(r[j + 1]) + b[j + 1];
                        const uint64_t lk = fmul2(pack_f32x2(w0, w1), slope2);
                        v[j] = fmaxf(w0, f32x2_lo(lk)); v[j + 1] = fmaxf(w1, f32x2_hi(lk));
                    }
                    if (kSplit) chain_store_a_split(a_sh, rt, hh * 128 + c * 32, v); else chain_store_a(a_sh, rt, hh * 128 + c * 32, v, fp16);
                };
#ifdef SRHEP_CHAIN_TMEM_PIPE      // A/B: the tensor-memory read of chunk c + 1 in flight under the arithmetic of chunk c (two register buffers)
                if constexpr (!kSplit) {
                    uint32_t ra[32], rb[32];
                    tmem_ld32(t_col, ra);
                    tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < 3) { if (c & 1) tmem_ld32(t_col + (c + 1) * 32, ra); else tmem_ld32(t_col + (c + 1) * 32, rb); }
                        if (c & 1) hidden_chunk(rb, c); else hidden_chunk(ra, c);
                        if (c < 3) tmem_ld_wait();
                    }
                } else
#endif
                {
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_col + c * 32, r);
                    tmem_ld_wait();
                    hidden_chunk(r, c);
                }
                }
                stage_done(true);
            }
            // ---------------------------------------------------------------- stage 2: MLP output, residual, next layer's LN1 + modulate
            {
                const bool next = p.n_stages > 3;
                float xr[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) xr[j] = 0.f;
                if (!kFirst && valid) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) ldg256_stream(xrow + (j >> 3) * 1024, &xr[j]);   // x1, written by this thread in stage 0
                }
                mbar_wait(&acc_full[hh], stage_it & 1);
                mbar_wait(&acc_full[hh ^ 1], stage_it & 1);
                if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 14);
                tc_fence_after();
                // last layer: the A buffer is handed back to the producer (a_free) as soon as this stage's MMAs retire, so nothing may be staged in it
                const bool st2 = staged && next;
                if (st2) stage_rows(kFirst ? p.row_bias : p.gate_mlp, p.p_nxt, p.q_nxt, kFirst ? p.ld_row_bias : p.ld_mod);
                uint64_t s1p = 0ull, s2p = 0ull;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t r[32];
                    tmem_ld32(t_col + c * 32, r);
                    tmem_ld_wait();
                    if (kFirst) chain_first_chunk<kSplit>(r, &p.cst[2][hh * 128 + c * 32], st2, psm + c * 128, p.row_bias + (size_t)evt * p.ld_row_bias + hh * 128 + c * 32, s1p, s2p, p.wscale[0]);
                    else chain_resid_chunk<true, kSplit>(r, xr, &p.cst[2][hh * 128 + c * 32], st2, psm + c * 128, p.gate_mlp + eo + c * 32, s1p, s2p, p.wscale[2]);
                    if (valid) {
                        if (!kDefer || !next) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);
                        }
                        if (!kFirst && c < 3) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) ldg256_stream(xrow + ((c + 1) * 4 + (j >> 3)) * 1024, &xr[j]);
                        }
                    }
                    if (next) tmem_st32(t_col + c * 32, r);
                }
                if (next) {
                    tmem_st_wait();
                    const float s1 = sum_f32x2(s1p), s2 = sum_f32x2(s2p);
                    sts_f2(st_own, s1, s2);
                    named_bar_sync(1, 256);
                    const float2 o1 = lds_f2(st_oth);
                    const float mean = (s1 + o1.x) * inv_n;
                    const float rstd = rsqrtf(fmaxf((s2 + o1.y) * inv_n - mean * mean, 0.f) + kLnEps);
                    const uint64_t rs2 = pack_f32x2(rstd, rstd), nm2 = pack_f32x2(-mean * rstd, -mean * rstd);
                    uint64_t t1p = 0ull, t2p = 0ull;
                    // Tensor-memory reads are the resource the two CTAs of an SM share (57 B per clock and SM, tools/micro/ldtm_bw.cu), so the
                    // LayerNorm output goes STRAIGHT to the A operand instead of being parked in TMEM and copied in a third pass.  The scratch
                    // and the staged adaLN rows of up to four events live in k-block 0 of the A buffer (columns 0-63, bytes 0 .. 16 KB) and are
                    // still being read by other threads: the threads that own those columns keep their 64 packed values in registers until the
                    // barrier; every other column is written at once.
                    const bool direct = !kSplit && ne <= 4 && p.ln_direct;      // CTA-uniform
                    if (direct) {
#pragma unroll 1
                        for (int c = 2; c < 4; ++c) {                    // columns that nobody else is reading: straight to A
                            uint32_t r[32];
                            tmem_ld32(t_col + c * 32, r);
                            tmem_ld_wait();
                            if (kDefer && valid) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);      // x2
                            }
                            chain_ln_mod_chunk(r, rs2, nm2, staged, psm + 1024 + c * 128, psm + 2048 + c * 128, p.p_nxt + eq + c * 32, p.q_nxt + eq + c * 32, t1p, t2p);
                            uint32_t pk[16];
#pragma unroll
                            for (int j = 0; j < 32; j += 2) pk[j >> 1] = pack16(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), fp16);
                            chain_store_a_packed(a_sh, rt, hh * 128 + c * 32, pk);
                        }
                        uint32_t held[32];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {                    // columns [hh * 128, hh * 128 + 64): k-block 0 for the hh = 0 threads, processed last so that little else is live
                            uint32_t r[32];
                            tmem_ld32(t_col + c * 32, r);
                            tmem_ld_wait();
                            if (kDefer && valid) {
#pragma unroll
                                for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);      // x2
                            }
                            chain_ln_mod_chunk(r, rs2, nm2, staged, psm + 1024 + c * 128, psm + 2048 + c * 128, p.p_nxt + eq + c * 32, p.q_nxt + eq + c * 32, t1p, t2p);
#pragma unroll
                            for (int j = 0; j < 32; j += 2) held[c * 16 + (j >> 1)] = pack16(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), fp16);
                            if (hh != 0) chain_store_a_packed(a_sh, rt, hh * 128 + c * 32, &held[c * 16]);
                        }
                        named_bar_sync(1, 256);                          // all reads of the scratch and the staged rows are done: k-block 0 may be rewritten
                        if (hh == 0) { chain_store_a_packed(a_sh, rt, 0, &held[0]); chain_store_a_packed(a_sh, rt, 32, &held[16]); }
                    } else {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {                        // next layer's LN1 affine + modulate, parked in TMEM (the staged rows are still being read)
                        uint32_t r[32];
                        tmem_ld32(t_col + c * 32, r);
                        tmem_ld_wait();
                        if (kDefer && valid) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) stg256_stream(xrow + (c * 4 + (j >> 3)) * 1024, &r[j]);      // x2
                        }
                        chain_ln_mod_chunk(r, rs2, nm2, staged, psm + 1024 + c * 128, psm + 2048 + c * 128, p.p_nxt + eq + c * 32, p.q_nxt + eq + c * 32, t1p, t2p);
                        tmem_st32(t_col + c * 32, r);
                    }
                    tmem_st_wait();
                    named_bar_sync(1, 256);                              // all reads of the scratch and the staged rows are done: A may be rewritten
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r[32];
                        tmem_ld32(t_col + c * 32, r);
                        tmem_ld_wait();
                        float v[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                        if (kSplit) chain_store_a_split(a_sh, rt, hh * 128 + c * 32, v); else chain_store_a(a_sh, rt, hh * 128 + c * 32, v, fp16);
                    }
                    }
                }
                stage_done(next);
            }
            // ---------------------------------------------------------------- stages 3-5: q, k, v of the next layer
            if (p.n_stages > 3) {
#pragma unroll 1
                for (int g = 0; g < 3; ++g) {
                    mbar_wait(&acc_full[hh], stage_it & 1);
                    if (warp == 2 && lane == 0) CHAIN_STAMP(tile_i, 16 + 2 * g);
                    tc_fence_after();
                    // 64 columns (one 128-byte line per row) at a time: bias, pack, regroup across the 4 lanes of a group, store whole lines
                    const int row4 = t * 128 + (rt & ~3);                                   // first row of this lane's group of four
                    const size_t doff = (size_t)row4 * (3 * kChainH) + g * kChainH + hh * 128 + (lane & 3) * 16;
                    uint16_t* dst = reinterpret_cast<uint16_t*>(p.qkv) + doff;
#pragma unroll 1
                    for (int blk = 0; blk < 2; ++blk) {
                        uint32_t a[32], al[kSplit ? 32 : 1];
                        uint32_t rr[2][32];
#ifdef SRHEP_CHAIN_TMEM_PIPE      // A/B: both halves of the 64-column block requested at once (one wait instead of two dependent read / wait pairs)
                        constexpr bool kBoth = !kSplit;
                        if (kBoth) { tmem_ld32(t_col + blk * 64, rr[0]); tmem_ld32(t_col + blk * 64 + 32, rr[1]); tmem_ld_wait(); }
#else
                        constexpr bool kBoth = false;
#endif
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const int c = blk * 2 + half;
                            uint32_t (&r)[32] = rr[half];
                            if (!kBoth) { tmem_ld32(t_col + c * 32, r); tmem_ld_wait(); }
                            const float* b = &p.cst[3 + g][hh * 128 + c * 32];
                            const float ws = p.wscale[kFirst ? 1 + g : 3 + g];
                            if (g == 0) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                if (kSplit) split16(fmaf(__uint_as_float(r[j]), ws, b[j]), fmaf(__uint_as_float(r[j + 1]), ws, b[j + 1]), a[half * 16 + (j >> 1)], al[kSplit ? half * 16 + (j >> 1) : 0]);
                                else a[half * 16 + (j >> 1)] = pack16(__uint_as_float(r[j]) + b[j], __uint_as_float(r[j + 1]) + b[j + 1], fp16);
                            }
                            } else {
                            // K and V carry no bias here: a key bias shifts every score of a query row by the same q . b_k, which the softmax cancels, and
                            // the value bias passes through the attention unchanged (the weights of a row sum to 1), so it sits in the out-projection's bias
                            // (b_o + W_o b_v, folded when the weights are packed: bf16_forward.inl).  One add and one constant load per element less.
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                if (kSplit) split16(__uint_as_float(r[j]) * ws, __uint_as_float(r[j + 1]) * ws, a[half * 16 + (j >> 1)], al[kSplit ? half * 16 + (j >> 1) : 0]);
                                else a[half * 16 + (j >> 1)] = pack16(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), fp16);
                            }
                            }
                        }
#ifdef SRHEP_QKV_ROW_STORES       // A/B: every lane stores the 128-byte line of ITS row as four 32-byte pieces (no transposes, four times the LSU wavefronts)
                        if (valid) {
                            uint16_t* drow = reinterpret_cast<uint16_t*>(p.qkv) + (size_t)row * (3 * kChainH) + g * kChainH + hh * 128 + blk * 64;
#pragma unroll
                            for (int i = 0; i < 4; ++i) stg256(drow + i * 16, &a[8 * i]);
                        }
#else
                        transpose_line_pieces(a, lane);
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (row4 + i < p.M) { SRHEP_CHECK(row4 + i < p.ext.rows_cap); stg256(dst + (size_t)i * (3 * kChainH) + blk * 64, &a[8 * i]); }
#endif
                        if constexpr (kSplit) {
                            transpose_line_pieces(al, lane);
                            uint16_t* dlo = reinterpret_cast<uint16_t*>(p.qkv_lo) + doff;
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (row4 + i < p.M) stg256(dlo + (size_t)i * (3 * kChainH) + blk * 64, &al[8 * i]);
                        }
                    }
                    stage_done(false);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

#undef xrow

}  // namespace srhep

// Host side of the srhep C ABI (include/srhep.h): weight packing, event binding (varlen
// maps instead of masks), the per-pass evaluation schedule, CUDA-graph capture of one
// network evaluation, and the fixed-grid / dopri5 ODE drivers.
//
// Reference behaviour followed (paths relative to the reference repo root):
//   models/flow_model.py:167-264   FlowModel.forward        -> Engine::enqueue_eval
//   models/flow_model.py:302-329   FlowModel.generate_samples -> srhep_sample*
//   torchdiffeq.odeint (external)  fixed grid + dopri5      -> srhep_sample / srhep_sample_dopri5
#include "../../include/srhep.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX 3: ranges cost nothing unless a tool (ncu --nvtx, nsys) is attached

#include "kernels_f32.cuh"
#include "kernels_bf16.cuh"
#include "kernels_chain.cuh"
#include "kernels_head.cuh"
#include "kernels_head_fused.cuh"
#include "kernels_attn.cuh"
#include "kernels_attn3.cuh"
#include "kernels_embed.cuh"
#include "kernels_ode.cuh"

using namespace srhep;

namespace {

thread_local std::string g_create_error;

struct Lin { size_t w = 0, b = 0; int out = 0, in = 0; };   // offsets into the fp32 blob

struct Layout {
    Lin t0, t2, eta1, eta3, lay1, lay3, prx1, prx3, nsy1, nsy3, feat0;
    size_t layer_table = 0;
    struct Layer { Lin q, k, v, o, m1, m2, ada; size_t n1w, n1b, n2w, n2b; };
    std::vector<Layer> layers;
    size_t fn_w = 0, fn_b = 0, nv_w = 0, nv_b = 0;
    Lin vada, h1, h2, h3, h4;
    size_t freqs = 0;
    size_t total = 0;
};

// Same order as SrDims.param_order() in superresolutionhep_b200/config.py (which follows
// the reference state_dict).
Layout make_layout(const SrhepDims& d) {
    Layout L;
    size_t off = 0;
    auto lin = [&](Lin& l, int out, int in) { l.out = out; l.in = in; l.w = off; off += (size_t)out * in; l.b = off; off += out; };
    auto vec = [&](size_t& o, int n) { o = off; off += n; };
    lin(L.t0, d.t_emb, d.freq_dim);
    lin(L.t2, d.t_emb, d.t_emb);
    lin(L.eta1, d.etaphi_hid, d.etaphi_in + d.t_emb);
    lin(L.eta3, d.etaphi_out, d.etaphi_hid);
    vec(L.layer_table, 3 * d.layer_emb_dim);
    lin(L.lay1, d.layer_hid, d.layer_emb_dim + d.t_emb);
    lin(L.lay3, d.layer_out, d.layer_hid);
    lin(L.prx1, d.proxy_hid, 1 + d.t_emb);
    lin(L.prx3, d.proxy_out, d.proxy_hid);
    lin(L.nsy1, d.noisy_hid, 1 + d.t_emb);
    lin(L.nsy3, d.noisy_out, d.noisy_hid);
    lin(L.feat0, d.h_dim, d.cond + d.noisy_out + d.ctx);
    L.layers.resize(d.layers);
    for (auto& y : L.layers) {
        lin(y.q, d.h_dim, d.h_dim); lin(y.k, d.h_dim, d.h_dim); lin(y.v, d.h_dim, d.h_dim); lin(y.o, d.h_dim, d.h_dim);
        lin(y.m1, d.mlp_hid, d.h_dim); lin(y.m2, d.h_dim, d.mlp_hid);
        vec(y.n1w, d.h_dim); vec(y.n1b, d.h_dim); vec(y.n2w, d.h_dim); vec(y.n2b, d.h_dim);
        lin(y.ada, 6 * d.h_dim, d.ctx);
    }
    vec(L.fn_w, d.h_dim); vec(L.fn_b, d.h_dim);
    vec(L.nv_w, d.v_in); vec(L.nv_b, d.v_in);
    lin(L.vada, 2 * d.v_in, d.ctx);
    lin(L.h1, d.head_h1, d.v_in + d.ctx);
    lin(L.h2, d.head_h2, d.head_h1);
    lin(L.h3, d.head_h3, d.head_h2);
    lin(L.h4, 1, d.head_h3);
    vec(L.freqs, d.freq_dim / 2);
    L.total = off;
    return L;
}

struct Pass {
    int e0 = 0, e1 = 0;          // events
    int r0 = 0, r1 = 0;          // global rows
    int c0 = 0, c1 = 0;          // embedding chunks (global index)
    int w0 = 0, w1 = 0;          // attention work items (rows stored pass-local)
    cudaGraphExec_t exec = nullptr;
    int graph_nodes = 0;
};

}  // namespace

struct SrhepHandle {
    int device = 0;
    SrhepDims d{};
    int precision = SRHEP_PREC_FP32;    // operand format of the kernels (split mode: SRHEP_PREC_FP16 planes)
    int precision_req = SRHEP_PREC_FP32; // what srhep_create was asked for
    bool split = false;                 // 'highest' on tensor cores: every 16-bit operand is an fp16 (hi, lo) pair, products run as hi.hi + hi.lo + lo.hi (kernels_chain.cuh, kernels_attn3.cuh)
    Layout L;
    std::string err;
    uint64_t launches = 0;

    // weights
    float* w = nullptr;                 // the fp32 blob as uploaded
    float* wqkv = nullptr;              // [layers][3h, h]
    float* bqkv = nullptr;              // [layers][3h]
    float* wmod = nullptr;              // [mod_width, ctx]   all adaLN Linears stacked
    float* bmod = nullptr;              // [mod_width]
    float* mod_tbias = nullptr;         // [mod_width]   bmod + the time-embedding columns' contribution (sampling passes share t)
    float* r1 = nullptr;                // 4 x kMaxHid: row sums of the context part of the embed nets' first Linear
    Bf16Weights bw;                     // bf16 re-packed GEMM operands (SRHEP_PREC_BF16)
    int mod_width = 0;

    // diagnostic switches (environment, read once per API call: A/B comparisons in the tests and tools)
    struct Switches { bool no_chain = false, no_chain_first = false, attn_simt = false, attn_v1 = false, attn_v2 = false, no_lnfuse = false, head_fp32 = false, no_headchain = false, no_embed_tc = false, head_prep_scalar = false, head_prep_v4 = false, no_head_fused = false, chain_dbg = false, attn_dbg = false; int ctas_per_sm = 2; bool chain_a_early = true, chain_ln_direct = true; int chain_a_pf = -1; int only = 0; } sw;      // only: energy diagnostics (results wrong on purpose): 1 = launch the attention kernels only, 2 = the layer-chain kernels only, 3 = everything but those two
    // options
    int64_t pass_tokens = 0;
    int use_graph = 1;
    int debug = 0;

    // bound events
    bool bound = false;
    SrhepCond cond{};
    int B = 0; int64_t T = 0;
    std::vector<int32_t> cu_host;
    std::vector<Pass> passes;
    int max_pass_rows = 0;
    int* cu_dev = nullptr; int* row_event = nullptr;
    int *chunk_event = nullptr, *chunk_row = nullptr, *chunk_len = nullptr, *ev_chunk_start = nullptr;
    AttnWork* attn_work = nullptr;
    size_t cap_events = 0, cap_rows = 0, cap_chunks = 0, cap_work = 0, cap_event_bufs = 0;

    // per-event buffers [B, .]
    float *temb = nullptr, *ev_a = nullptr, *ev_stats = nullptr, *layer_out = nullptr, *ctx = nullptr, *silu_ctx = nullptr;
    float *mod = nullptr, *f0bias = nullptr, *partial = nullptr, *t_fill = nullptr;
    float* modpq = nullptr;             // [B, layers * 4 * h_dim]   P | Q vectors of every LayerNorm-modulate of the layer chain (kernels_chain.cuh: modpq_kernel)
    // per-pass workspace [max_pass_rows, .]
    float *tok_feat = nullptr, *xres = nullptr, *qkv = nullptr, *h1buf = nullptr;
    void *act_a = nullptr, *act_b = nullptr;      // GEMM A operands (fp32 or bf16): ln_out / attn_out / mlp hidden / head in
    void *qkv_lp = nullptr;                        // bf16 q|k|v (SRHEP_PREC_BF16)
    void *qkv_lo = nullptr, *act_b_lo = nullptr;   // split mode: low planes of q|k|v and of the attention output
    size_t cap_ws_rows = 0;
    // ODE state
    float *y_a = nullptr, *y_b = nullptr, *y_tmp = nullptr;
    float* kbuf[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* ybuf2 = nullptr;
    size_t cap_state = 0, cap_k = 0;
    double* red_dev = nullptr; double* red_host = nullptr;
    // stages
    StageParams* sp_dev = nullptr; StageParams* sp_host = nullptr; size_t cap_stages = 0, stage_stride = 0;
    int* stage_idx_dev = nullptr; size_t cap_stage_idx = 0;
    cudaEvent_t sp_done = nullptr;
    cudaStream_t cap_stream = nullptr;

    // device-resident dopri5 (dopri5.inl): control block, per-pass stage descriptors and cursors, time grid, the cached graph
    Dopri5Ctl* dp_ctl = nullptr; Dopri5Ctl* dp_ctl_host = nullptr;
    StageParams* dp_sp = nullptr; int* dp_idx = nullptr; size_t dp_cap_pass = 0;
    float* dp_tg = nullptr; size_t dp_cap_tg = 0;
    cudaGraphExec_t dp_exec = nullptr; cudaStream_t dp_body_stream = nullptr;
    int dp_init_nodes = 0, dp_body_nodes = 0;

    // debug taps (single pass)
    float* tap_layers = nullptr; float* tap_feat0 = nullptr; float* tap_final = nullptr; size_t cap_tap = 0;
};

namespace {

int fail(SrhepHandle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(h, call)                                                                          \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess)                                                               \
            return fail(h, e_ == cudaErrorMemoryAllocation ? SRHEP_E_NOMEM : SRHEP_E_CUDA, "%s:%d %s: %s", \
                        __FILE__, __LINE__, #call, cudaGetErrorString(e_));                   \
    } while (0)

template <typename T>
int ensure(SrhepHandle* h, T*& p, size_t& cap, size_t need_elems, size_t elems_per_unit = 1) {
    (void)elems_per_unit;
    if (need_elems <= cap && p) return 0;
    if (p) CK(h, cudaFree(p));
    p = nullptr;
    CK(h, cudaMalloc(&p, std::max<size_t>(need_elems, 1) * sizeof(T)));
    cap = need_elems;
    return 0;
}

template <typename T>
int dev_alloc(SrhepHandle* h, T*& p, size_t n) {
    if (p) { CK(h, cudaFree(p)); p = nullptr; }
    CK(h, cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

bool is_lp(const SrhepHandle* h) { return h->precision != SRHEP_PREC_FP32; }
// the tcgen05 kernels the split (fp32-grade) mode is built from exist for the shipped architecture only
bool split_supported(const SrhepDims& d) {
    return d.h_dim == 256 && d.mlp_hid == 256 && d.heads > 0 && d.h_dim / d.heads == 64 && (d.cond + d.noisy_out + 63) / 64 * 64 == 192 &&
           d.head_h1 == 128 && (d.v_in + d.ctx) % 64 == 0 && d.layers <= 64;
}
void read_switches(SrhepHandle* h) {
    auto on = [](const char* n) { const char* v = getenv(n); return v && *v && *v != '0'; };
    h->sw.no_chain = on("SRHEP_NO_CHAIN"); h->sw.no_chain_first = on("SRHEP_NO_CHAIN_FIRST"); h->sw.attn_simt = on("SRHEP_ATTN_SIMT"); h->sw.attn_v1 = on("SRHEP_ATTN_V1"); h->sw.attn_v2 = on("SRHEP_ATTN_V2");
    h->sw.no_lnfuse = on("SRHEP_NO_LNFUSE"); h->sw.head_fp32 = on("SRHEP_HEAD_FP32"); h->sw.no_headchain = on("SRHEP_NO_HEADCHAIN"); h->sw.no_embed_tc = on("SRHEP_NO_EMBED_TC"); h->sw.head_prep_scalar = on("SRHEP_HEAD_PREP_SCALAR"); h->sw.head_prep_v4 = on("SRHEP_HEAD_PREP_V4"); h->sw.no_head_fused = on("SRHEP_NO_HEAD_FUSED");
    h->sw.chain_dbg = on("SRHEP_CHAIN_DBG"); h->sw.attn_dbg = on("SRHEP_ATTN_DBG");
    { const char* v = getenv("SRHEP_CHAIN_A_EARLY"); h->sw.chain_a_early = !(v && *v == '0'); v = getenv("SRHEP_CHAIN_LN_DIRECT"); h->sw.chain_ln_direct = !(v && *v == '0'); }
    { const char* v = getenv("SRHEP_CHAIN_A_PF"); h->sw.chain_a_pf = v ? atoi(v) : -1; }      // -1 (default): no L2 prefetch of the next A tile; else the weight-slot index of a tile at which it is issued
    { const char* v = getenv("SRHEP_ONLY"); h->sw.only = v ? atoi(v) : 0; }
    { const char* v = getenv("SRHEP_CTAS_PER_SM"); h->sw.ctas_per_sm = (v && *v == '1') ? 1 : 2; }      // persistent grids of the chain / attention kernels: CTAs per SM
}
size_t act_elem_size(const SrhepHandle* h) { return is_lp(h) ? 2 : 4; }

int validate_dims(const SrhepDims& d) {
    auto bad = [&](const char* m) { return fail(nullptr, SRHEP_E_INVALID, "unsupported dims: %s", m); };
    if (d.h_dim <= 0 || d.h_dim % 32 || d.h_dim > 512) return bad("h_dim must be a multiple of 32, <= 512");
    if (d.heads <= 0 || d.h_dim % d.heads) return bad("h_dim % heads");
    const int hd = d.h_dim / d.heads;
    if (hd != 16 && hd != 32 && hd != 64) return bad("head dim must be 16, 32 or 64");
    if (d.t_emb <= 0 || d.t_emb > kMaxTemb) return bad("t_emb <= 128");
    if (d.freq_dim <= 0 || d.freq_dim % 2 || d.freq_dim > kMaxFreq) return bad("freq_dim even, <= 512");
    if (d.etaphi_in != 3) return bad("etaphi_in must be 3");
    if (d.etaphi_hid > kMaxHid || d.layer_hid > kMaxHid || d.proxy_hid > kMaxHid || d.noisy_hid > kMaxHid) return bad("embed hidden <= 64");
    if (d.layer_emb_dim <= 0 || d.layer_emb_dim > 16) return bad("layer_emb_dim <= 16");
    if (d.cond != d.etaphi_out + d.layer_out + d.proxy_out + 1 || d.ctx != d.t_emb + d.cond || d.v_in != d.h_dim + d.cond) return bad("derived dims inconsistent");
    if (d.cond + d.noisy_out > 192) return bad("cond + noisy_out <= 192");
    if (d.cond % 32 || d.ctx % 32 || (d.cond + d.noisy_out) % 4) return bad("cond, ctx multiples of 32");
    if (d.layer_out > kMaxTemb) return bad("layer_out <= 128");
    if (d.mlp_hid % 32 || d.mlp_hid > 512) return bad("mlp_hid multiple of 32");
    if (d.head_h1 % 32 || d.head_h1 > 256 || d.head_h2 % 32 || d.head_h2 > 128 || d.head_h3 % 32 || d.head_h3 > 64) return bad("head widths");
    if ((d.v_in + d.ctx) > 1024) return bad("v_in + ctx <= 1024");
    if (d.layers <= 0 || d.layers > 64) return bad("layers");
    return 0;
}

// ----------------------------------------------------------------------------------------
// launch helpers
// ----------------------------------------------------------------------------------------
struct Engine;
void bf16_forward(Engine& E, const Pass& p, const int* rev, const StageRef& st);      // bf16_forward.inl
int bf16_pack_weights(SrhepHandle* h, const float* weights_host);
void bf16_free_weights(SrhepHandle* h);
int bf16_on_bind(SrhepHandle* h);
int64_t default_pass_tokens(int precision);

// NVTX range over an API call / an evaluation / a kernel category (ncu --nvtx --nvtx-include "srhep_sample/" ...)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Profiler {
    std::vector<cudaEvent_t> ev;       // ev[0] = start; ev[i+1] recorded after launch i
    std::vector<int> cat;
};

struct Engine {
    SrhepHandle* h;
    cudaStream_t s;
    int rc = 0;
    Profiler* prof = nullptr;
    int cat = 0;
    bool head_done = false;          // the tcgen05 head chain already produced v and the ODE update for this evaluation
    bool x_blocked = false;          // residual stream of the current evaluation is in the blocked layout (tcgen05 chain path)

    const float* W(size_t off) const { return h->w + off; }

    void check(const char* what) {
        if (rc) return;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = fail(h, SRHEP_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
        ++h->launches;
        if (prof && !rc) {
            cudaEvent_t evt;
            if (cudaEventCreate(&evt) != cudaSuccess || cudaEventRecord(evt, s) != cudaSuccess) { rc = fail(h, SRHEP_E_CUDA, "profile event"); return; }
            prof->ev.push_back(evt); prof->cat.push_back(cat);
        }
    }

    // C = epilogue(A . W^T); A is fp32 here (per-event GEMMs and the whole 'highest' path)
    template <typename OutT>
    void gemm_f32(const float* A, int lda, const float* Wt, int ldw, OutT* C, int ldc, int M, int N, int K, const GemmEpilogue& ep) {
        if (rc || M <= 0) return;
        const bool big = M >= 256 && N >= 1024 && K % 16 == 0 && lda % 4 == 0 && ldw % 4 == 0 && !ep.resid &&
                         ((uintptr_t)A & 15) == 0 && ((uintptr_t)Wt & 15) == 0;
        if (big) {
            dim3 grid((M + 127) / 128, (N + 127) / 128);
            gemm_f32_big_kernel<OutT><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, C, ldc, M, N, K, ep);
        } else {
            dim3 grid((M + 63) / 64, (N + 63) / 64);
            gemm_f32_kernel<OutT><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, C, ldc, M, N, K, ep);
        }
        check("gemm_f32");
    }

    EmbedNetDev net(const Lin& l1, const Lin& l3, int d, int slot) const {
        EmbedNetDev n;
        n.w1 = W(l1.w); n.b1 = W(l1.b); n.r1 = h->r1 + slot * kMaxHid; n.w2 = W(l3.w); n.b2 = W(l3.b);
        n.d = d; n.hid = l1.out; n.out = l3.out;
        return n;
    }

    // reference-order softmax(QK^T / sqrt(hd)) V on CUDA cores (fp32 math; T = storage type)
    template <typename T>
    void attention_simt(const Pass& p, const T* qkv, T* out) {
        if (rc || p.w1 == p.w0) return;
        const SrhepDims& d = h->d;
        const int hd = d.h_dim / d.heads;
        dim3 grid(p.w1 - p.w0, d.heads);
        const float inv = 1.0f / sqrtf((float)hd);
        const AttnWork* wk = h->attn_work + p.w0;
        const int ld = 3 * d.h_dim;
        if (hd == 64) attn_f32_kernel<64, T><<<grid, 128, 0, s>>>(qkv, ld, qkv + d.h_dim, qkv + 2 * d.h_dim, ld, out, d.h_dim, wk, inv);
        else if (hd == 32) attn_f32_kernel<32, T><<<grid, 128, 0, s>>>(qkv, ld, qkv + d.h_dim, qkv + 2 * d.h_dim, ld, out, d.h_dim, wk, inv);
        else attn_f32_kernel<16, T><<<grid, 128, 0, s>>>(qkv, ld, qkv + d.h_dim, qkv + 2 * d.h_dim, ld, out, d.h_dim, wk, inv);
        check("attn_simt");
    }

    template <typename OutT>
    void ln_mod(const float* x, int M, int Wd, const float* lw, const float* lb, const float* shift, const float* scale,
                const int* row_event, int second, OutT* out) {
        if (rc || M <= 0) return;
        LnModParams q;
        q.x = x; q.ldx = Wd; q.M = M; q.W = Wd; q.ln_w = lw; q.ln_b = lb; q.shift = shift; q.scale = scale;
        q.ld_mod = h->mod_width; q.row_event = row_event; q.second_ln = second;
        ln_mod_kernel<OutT><<<(M + 7) / 8, 256, 0, s>>>(q, out, Wd);
        check("ln_mod");
    }

    template <typename OutT>
    void head_prep(const HeadPrepParams& q, OutT* hin, int ldh, OutT* hin_lo = nullptr) {
        if (rc || q.M <= 0) return;
        const SrhepDims& d = h->d;
        const int grid = (q.M + 7) / 8;
        const int ph = d.h_dim / 32, pc = d.cond / 32, px = d.ctx / 32;
        if (ph == 8 && pc == 3 && px == 5 && q.x_blocked && sizeof(OutT) == 2 && q.ldt % 4 == 0 && is_lp(h) && !h->sw.head_prep_scalar) {
            if (std::is_same<OutT, __half>::value && !hin_lo && !q.final_tap && !h->sw.head_prep_v4) {      // one reduction round per LayerNorm, two rows per round (kernels_head_fused.cuh)
                head_prep_v5_kernel<<<(q.M + 63) / 64, 256, 0, s>>>(q, (__half*)hin, ldh);
                check("head_prep_v5");
                return;
            }
            head_prep_v4_kernel<OutT><<<(q.M + 63) / 64, 256, 0, s>>>(q, hin, ldh, hin_lo);
            check("head_prep_v4");
            return;
        }
        if (hin_lo) { rc = fail(h, SRHEP_E_INVALID, "head_prep: the split mode needs the float4 kernel's shapes"); return; }
        if (ph == 8 && pc == 3 && px == 5) head_prep_kernel<OutT, 8, 3, 5><<<grid, 256, 0, s>>>(q, hin, ldh);
        else if (ph == 2 && pc == 3 && px == 5) head_prep_kernel<OutT, 2, 3, 5><<<grid, 256, 0, s>>>(q, hin, ldh);
        else { rc = fail(h, SRHEP_E_INVALID, "head_prep: unsupported (h,cond,ctx)/32 = (%d,%d,%d)", ph, pc, px); return; }
        check("head_prep");
    }

    // One network evaluation for one pass (FlowModel.forward, models/flow_model.py:167-264).
    void enqueue_eval(const Pass& p, const StageRef& st) {
        NvtxRange nvtx_eval("srhep_eval");
        const SrhepDims& d = h->d;
        const Layout& L = h->L;
        const int nE = p.e1 - p.e0, M = p.r1 - p.r0;
        if (nE <= 0) return;
        const bool lp = is_lp(h);
        const int* rev = h->row_event + p.r0;
        const int ncol = d.cond + d.noisy_out;

        const bool rest = h->sw.only == 0 || h->sw.only == 3;      // energy diagnostics: SRHEP_ONLY=1 / 2 launch the attention / chain kernels alone
        cat = SRHEP_CAT_EMBED;
        if (rest) {   // 1. per-event preparation
            EventPrepParams q;
            q.freqs = W(L.freqs); q.half = d.freq_dim / 2;
            q.wt0 = W(L.t0.w); q.bt0 = W(L.t0.b); q.wt2 = W(L.t2.w); q.bt2 = W(L.t2.b); q.t_emb = d.t_emb;
            q.etaphi = net(L.eta1, L.eta3, d.etaphi_in, 0);
            q.layer = net(L.lay1, L.lay3, d.layer_emb_dim, 1);
            q.proxy = net(L.prx1, L.prx3, 1, 2);
            q.noisy = net(L.nsy1, L.nsy3, 1, 3);
            q.layer_table = W(L.layer_table); q.layer_emb_dim = d.layer_emb_dim;
            q.temb = h->temb; q.ev_a = h->ev_a; q.ev_stats = h->ev_stats; q.layer_out = h->layer_out;
            q.stage = st; q.e0 = p.e0;
            // a sampling pass evaluates every event at the same time t: the timestep embedding and everything derived from it
            // alone is computed for ONE event and copied to the others (per-event times: srhep_velocity with t_event)
            const bool shared_t = st.t_event == nullptr;             // also for a single event: results must not depend on how a batch is cut into passes
            if (!rc) { event_prep_kernel<<<shared_t ? 1 : nE, shared_t ? 512 : 128, 0, s>>>(q);   /* the single block is pure latency: 16 warps shorten its matvec loops */ check("event_prep"); }
            if (shared_t && nE > 1 && !rc) {
                BroadcastPrepParams b;
                b.buf[0] = h->temb; b.len[0] = d.t_emb; b.buf[1] = h->ev_a; b.len[1] = 3 * kMaxHid;
                b.buf[2] = h->ev_stats; b.len[2] = 2; b.buf[3] = h->layer_out; b.len[3] = 3 * d.layer_out;
                b.e0 = p.e0;
                broadcast_prep_kernel<<<nE - 1, 128, 0, s>>>(b); check("broadcast_prep");
            }
        }
        const bool embed_tc = lp && h->bw.embed_tc && !h->sw.no_embed_tc && !h->split;      // split mode: the fp32 CUDA-core embedding, planes written by its store
        if (!rest) { if (M > 0 && lp) bf16_forward(*this, p, rev, st); return; }
        if (M > 0 && embed_tc) {   // 2. per-cell embeddings on the tensor core (kernels_embed.cuh)
            EmbedTcParams q = *static_cast<const EmbedTcParams*>(h->bw.embed_tpl);
            q.M = M; q.row0 = p.r0; q.lp_fp16 = h->precision == SRHEP_PREC_FP16;
            q.eta = h->cond.eta; q.cosphi = h->cond.cosphi; q.sinphi = h->cond.sinphi; q.e_proxy = h->cond.e_proxy; q.layer = h->cond.layer;
            q.stage = st; q.row_event = rev;
            q.ev_a = h->ev_a; q.ev_stats = h->ev_stats; q.layer_out = h->layer_out;
            q.w_img = h->bw.img + h->bw.embed_w;
            q.tok_feat = h->tok_feat; q.ld = ncol; q.tok_lp = h->bw.tok_lp; q.ld_lp = h->bw.feat0_kpad;
            if (!rc) { embed_tc_kernel<<<std::min((M + 127) / 128, 3 * 148), kEmbThreads, kEmbSmemBytes, s>>>(q); check("embed_tc"); }
        } else if (M > 0) {   // 2. per-cell embeddings + chunk sums
            EmbedTokParams q;
            q.etaphi = net(L.eta1, L.eta3, d.etaphi_in, 0);
            q.proxy = net(L.prx1, L.prx3, 1, 2);
            q.noisy = net(L.nsy1, L.nsy3, 1, 3);
            q.layer_out_dim = d.layer_out; q.t_emb = d.t_emb; q.cond = d.cond; q.ncol = ncol;
            q.eta = h->cond.eta; q.cosphi = h->cond.cosphi; q.sinphi = h->cond.sinphi; q.e_proxy = h->cond.e_proxy; q.layer = h->cond.layer;
            q.stage = st; q.row0 = p.r0; q.chunk0 = p.c0; q.chunk1 = p.c1;
            q.ev_a = h->ev_a; q.ev_stats = h->ev_stats; q.layer_out = h->layer_out;
            q.chunk_event = h->chunk_event; q.chunk_row = h->chunk_row; q.chunk_len = h->chunk_len;
            q.tok_feat = h->tok_feat; q.ld = ncol; q.partial = h->partial;
            q.tok_lp = lp ? (void*)h->bw.tok_lp : nullptr; q.ld_lp = h->bw.feat0_kpad; q.lp_fp16 = h->precision == SRHEP_PREC_FP16;
            q.tok_lp_lo = h->split ? (void*)h->bw.tok_lp_lo : nullptr;
            if (!rc) { embed_tokens_kernel<<<std::min(p.c1 - p.c0, 148 * 8), 192, 0, s>>>(q); check("embed_tokens"); }
        }
        if (embed_tc) {   // 3. context from the cond columns of tok_feat
            ContextRowsParams q;
            q.temb = h->temb; q.tok_feat = h->tok_feat; q.ld = ncol; q.row0 = p.r0; q.cu_seqlens = h->cu_dev;
            q.ctx = h->ctx; q.silu_ctx = h->silu_ctx; q.t_emb = d.t_emb; q.cond = d.cond; q.e0 = p.e0;
            if (!rc) { context_rows_kernel<<<nE, 384, 0, s>>>(q); check("context_rows"); }
        } else
        {   // 3. context
            ContextParams q;
            q.temb = h->temb; q.partial = h->partial; q.ev_chunk_start = h->ev_chunk_start; q.cu_seqlens = h->cu_dev;
            q.ctx = h->ctx; q.silu_ctx = h->silu_ctx; q.t_emb = d.t_emb; q.cond = d.cond; q.e0 = p.e0;
            if (!rc) { context_kernel<<<nE, 256, 0, s>>>(q); check("context"); }
        }
        cat = SRHEP_CAT_ADALN;
        {   // 4. every adaLN Linear of the evaluation in one GEMM; 5. context part of feat_0
            const float* sc = h->silu_ctx + (size_t)p.e0 * d.ctx;
            float* mo = h->mod + (size_t)p.e0 * h->mod_width;
            if (st.t_event == nullptr && d.t_emb % 16 == 0 && (d.ctx - d.t_emb) % 16 == 0) {
                // shared evaluation time: context = [time_emb | cond mean], so the time columns contribute the same vector to
                // every event: bias' = b + W[:, :t_emb] silu(time_emb) once, then only the cond columns per event
                GemmEpilogue e0; e0.bias = h->bmod;
                gemm_f32<float>(sc, d.ctx, h->wmod, d.ctx, h->mod_tbias, h->mod_width, 1, h->mod_width, d.t_emb, e0);
                GemmEpilogue ep; ep.bias = h->mod_tbias;
                gemm_f32<float>(sc + d.t_emb, d.ctx, h->wmod + d.t_emb, d.ctx, mo, h->mod_width, nE, h->mod_width, d.ctx - d.t_emb, ep);
            } else {
                GemmEpilogue ep; ep.bias = h->bmod;
                gemm_f32<float>(sc, d.ctx, h->wmod, d.ctx, mo, h->mod_width, nE, h->mod_width, d.ctx, ep);
            }
            if (h->modpq && h->bw.bias && !rc) {      // LayerNorm affine x adaLN modulation of every layer as one multiply-add per column for the layer chain
                ModPqParams q;
                q.mod = h->mod; q.ld_mod = h->mod_width; q.nrm = h->bw.bias + 3 * (size_t)d.h_dim; q.nrm_stride = (int)h->bw.bias_layer_stride;
                q.pq = h->modpq; q.ld_pq = d.layers * 4 * kChainH; q.layers = d.layers; q.e0 = p.e0;
                modpq_kernel<<<nE, 256, 0, s>>>(q); check("modpq");
            }
            GemmEpilogue e2; e2.bias = W(L.feat0.b);
            gemm_f32<float>(h->ctx + (size_t)p.e0 * d.ctx, d.ctx, W(L.feat0.w) + ncol, L.feat0.in, h->f0bias + (size_t)p.e0 * d.h_dim, d.h_dim,
                            nE, d.h_dim, d.ctx, e2);
        }
        if (M <= 0) return;
        const int H = d.h_dim;
        const float* mod = h->mod;
        float* x = h->xres;
        if (!lp) {
            float* a = (float*)h->act_a; float* b = (float*)h->act_b;
            cat = SRHEP_CAT_FEAT0;
            {   // 6. feat_0 token part + per-event bias, LeakyReLU
                GemmEpilogue ep; ep.row_bias = h->f0bias; ep.ld_row_bias = H; ep.row_event = rev; ep.act = 1;
                gemm_f32<float>(h->tok_feat, ncol, W(L.feat0.w), L.feat0.in, x, H, M, H, ncol, ep);
            }
            tap(h->tap_feat0, x, M);
            for (int l = 0; l < d.layers; ++l) {
                const Layout::Layer& y = L.layers[l];
                const float* ml = mod + (size_t)l * 6 * H;        // shift_msa | scale_msa | gate_msa | shift_mlp | scale_mlp | gate_mlp
                cat = SRHEP_CAT_LN;
                ln_mod<float>(x, M, H, W(y.n1w), W(y.n1b), ml, ml + H, rev, 0, a);
                cat = SRHEP_CAT_QKV;
                { GemmEpilogue ep; ep.bias = h->bqkv + (size_t)l * 3 * H;
                  gemm_f32<float>(a, H, h->wqkv + (size_t)l * 3 * H * H, H, h->qkv, 3 * H, M, 3 * H, H, ep); }
                cat = SRHEP_CAT_ATTN;
                attention_simt<float>(p, h->qkv, b);
                cat = SRHEP_CAT_OUT;
                { GemmEpilogue ep; ep.bias = W(y.o.b); ep.gate = ml + 2 * H; ep.ld_gate = h->mod_width; ep.row_event = rev; ep.resid = x; ep.ld_resid = H;
                  gemm_f32<float>(b, H, W(y.o.w), H, x, H, M, H, H, ep); }
                cat = SRHEP_CAT_LN;
                ln_mod<float>(x, M, H, W(y.n2w), W(y.n2b), ml + 3 * H, ml + 4 * H, rev, 1, a);
                cat = SRHEP_CAT_MLP1;
                { GemmEpilogue ep; ep.bias = W(y.m1.b); ep.act = 1;
                  gemm_f32<float>(a, H, W(y.m1.w), H, b, d.mlp_hid, M, d.mlp_hid, H, ep); }
                cat = SRHEP_CAT_MLP2;
                { GemmEpilogue ep; ep.bias = W(y.m2.b); ep.act = 1; ep.gate = ml + 5 * H; ep.ld_gate = h->mod_width; ep.row_event = rev; ep.resid = x; ep.ld_resid = H;
                  gemm_f32<float>(b, d.mlp_hid, W(y.m2.w), d.mlp_hid, x, H, M, H, d.mlp_hid, ep); }
                if (h->debug && h->tap_layers) tap(h->tap_layers + (size_t)l * h->cap_tap * H, x, M);
            }
            cat = SRHEP_CAT_HEAD;
            const int hw = d.v_in + d.ctx;
            head_prep<float>(head_params(p, x), a, hw);
            { GemmEpilogue ep; ep.bias = W(L.h1.b); ep.act = 1;
              gemm_f32<float>(a, hw, W(L.h1.w), hw, h->h1buf, d.head_h1, M, d.head_h1, hw, ep); }
        } else {
            bf16_forward(*this, p, rev, st);
        }
        cat = SRHEP_CAT_HEAD;
        if (!head_done) {   // head tail + ODE update
            HeadTailParams q;
            q.h1 = h->h1buf; q.ldh = d.head_h1; q.M = M;
            q.w2 = W(L.h2.w); q.b2 = W(L.h2.b); q.w3 = W(L.h3.w); q.b3 = W(L.h3.b); q.w4 = W(L.h4.w); q.b4 = W(L.h4.b);
            q.H1 = d.head_h1; q.H2 = d.head_h2; q.H3 = d.head_h3; q.final_ln = d.head_final_ln;
            q.stage = st;
            const size_t smem = ((size_t)d.head_h1 * d.head_h2 + (size_t)d.head_h2 * d.head_h3 + 2 * d.head_h3 + d.head_h2 + kHeadWarps * d.head_h1) * sizeof(float);
            const int grid = std::min((M + kHeadWarps - 1) / kHeadWarps, 148 * 4);
            if (!rc) { head_tail_kernel<<<grid, kHeadWarps * 32, smem, s>>>(q); check("head_tail"); }
        }
    }

    HeadPrepParams head_params(const Pass& p, const float* x) {
        const SrhepDims& d = h->d; const Layout& L = h->L;
        HeadPrepParams q;
        q.x = x; q.ldx = d.h_dim; q.tok_feat = h->tok_feat; q.ldt = d.cond + d.noisy_out;
        q.fn_w = W(L.fn_w); q.fn_b = W(L.fn_b); q.nv_w = W(L.nv_w); q.nv_b = W(L.nv_b);
        if (is_lp(h) && h->bw.bias) {          // 16-byte aligned copies for the float4 kernel
            const float* bf = h->bw.bias + h->bw.bias_fn;
            q.fn_w = bf; q.fn_b = bf + d.h_dim; q.nv_w = bf + 2 * d.h_dim; q.nv_b = bf + 2 * d.h_dim + d.v_in;
        }
        const float* mv = h->mod + (size_t)d.layers * 6 * d.h_dim;
        q.shift = mv; q.scale = mv + d.v_in; q.ld_mod = h->mod_width;
        q.ctx = h->ctx; q.ctx_dim = d.ctx; q.row_event = h->row_event + p.r0;
        q.M = p.r1 - p.r0; q.h = d.h_dim; q.cond = d.cond;
        q.final_tap = h->debug ? h->tap_final : nullptr;
        q.x_blocked = x_blocked ? 1 : 0;
        q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
        return q;
    }

    void tap(float* dst, const float* src, int M) {
        if (rc || !h->debug || !dst) return;
        if (x_blocked) {
            unblock_x_kernel<<<std::min((M * 256 + 255) / 256, 148 * 8), 256, 0, s>>>(src, dst, M);
            check("unblock_x");
            return;
        }
        cudaError_t e = cudaMemcpyAsync(dst, src, (size_t)M * h->d.h_dim * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) rc = fail(h, SRHEP_E_CUDA, "tap copy: %s", cudaGetErrorString(e));
    }

    void bump(int* idx) {
        if (rc) return;
        bump_stage_kernel<<<1, 32, 0, s>>>(idx);
        check("bump_stage");
    }

    void combine(const float* base, std::initializer_list<std::pair<const float*, float>> terms, float* out, size_t n) {
        if (rc || n == 0) return;
        CombineParams q; q.base = base; q.out = out; q.n = n; q.nk = 0;
        for (auto& t : terms) { if (t.second == 0.f) continue; q.k[q.nk] = t.first; q.c[q.nk] = t.second; ++q.nk; }
        for (int i = q.nk; i < 7; ++i) { q.k[i] = nullptr; q.c[i] = 0.f; }
        const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
        combine_kernel<<<grid, 256, 0, s>>>(q);
        check("combine");
    }
};

#include "bf16_forward.inl"

// ----------------------------------------------------------------------------------------
int alloc_for_binding(SrhepHandle* h) {
    const SrhepDims& d = h->d;
    const size_t B = std::max(h->B, 1);
    int rc;
    if ((rc = ensure(h, h->cu_dev, h->cap_events, B + 1))) return rc;
    if (B <= h->cap_event_bufs) return 0;                  // per-event buffers only ever grow
    if ((rc = dev_alloc(h, h->temb, B * d.t_emb))) return rc;
    if ((rc = dev_alloc(h, h->ev_a, B * 3 * kMaxHid))) return rc;
    if ((rc = dev_alloc(h, h->ev_stats, B * 2))) return rc;
    if ((rc = dev_alloc(h, h->layer_out, B * 3 * d.layer_out))) return rc;
    if ((rc = dev_alloc(h, h->ctx, B * d.ctx))) return rc;
    if ((rc = dev_alloc(h, h->silu_ctx, B * d.ctx))) return rc;
    if ((rc = dev_alloc(h, h->mod, B * h->mod_width))) return rc;
    if (h->bw.bias && h->d.h_dim == kChainH && (rc = dev_alloc(h, h->modpq, B * (size_t)h->d.layers * 4 * kChainH))) return rc;
    if ((rc = dev_alloc(h, h->f0bias, B * d.h_dim))) return rc;
    if ((rc = dev_alloc(h, h->t_fill, B))) return rc;
    if ((rc = dev_alloc(h, h->ev_chunk_start, B + 1))) return rc;
    h->cap_event_bufs = B;
    return 0;
}

int alloc_workspace(SrhepHandle* h) {
    const SrhepDims& d = h->d;
    size_t R = std::max(h->max_pass_rows, 1);
    if (R <= h->cap_ws_rows) return 0;
    R = (R + 4095) / 4096 * 4096;
    int rc;
    const size_t es = act_elem_size(h);
    const size_t wide = std::max<size_t>(d.v_in + d.ctx, std::max(d.h_dim, d.mlp_hid));
    auto re = [&](auto*& p, size_t bytes) -> int {
        if (p) { CK(h, cudaFree(p)); p = nullptr; }
        CK(h, cudaMalloc((void**)&p, bytes));
        return 0;
    };
    if ((rc = re(h->tok_feat, R * (d.cond + d.noisy_out) * sizeof(float)))) return rc;
    if ((rc = re(h->xres, R * d.h_dim * sizeof(float)))) return rc;
    if ((rc = re(h->h1buf, R * d.head_h1 * sizeof(float)))) return rc;
    if ((rc = re(h->act_a, R * wide * 4))) return rc;
    if ((rc = re(h->act_b, R * std::max(d.h_dim, d.mlp_hid) * es))) return rc;
    if (is_lp(h)) {
        if ((rc = re(h->qkv_lp, R * 3 * d.h_dim * 2))) return rc;
        CK(h, cudaMemset(h->qkv_lp, 0, R * 3 * d.h_dim * 2));    // rows past a pass's end are read (masked) by the attention tiles: keep them finite
        if (h->split) {
            if ((rc = re(h->qkv_lo, R * 3 * d.h_dim * 2))) return rc;
            CK(h, cudaMemset(h->qkv_lo, 0, R * 3 * d.h_dim * 2));
            if ((rc = re(h->act_b_lo, R * d.h_dim * 2))) return rc;
        }
    }
    else { if ((rc = re(h->qkv, R * 3 * d.h_dim * sizeof(float)))) return rc; }
    h->cap_ws_rows = R;
    if (is_lp(h) && (rc = bf16_on_bind(h))) return rc;     // tensor maps follow the workspace
    return 0;
}

void drop_graphs(SrhepHandle* h) {
    for (auto& p : h->passes) if (p.exec) { cudaGraphExecDestroy(p.exec); p.exec = nullptr; }
    if (h->dp_exec) { cudaGraphExecDestroy(h->dp_exec); h->dp_exec = nullptr; }
}

int ensure_state(SrhepHandle* h, bool need_k) {
    const size_t T = std::max<int64_t>(h->T, 1);
    if (T > h->cap_state) {
        for (float** p : {&h->y_a, &h->y_b, &h->y_tmp}) { if (*p) CK(h, cudaFree(*p)); *p = nullptr; CK(h, cudaMalloc(p, T * sizeof(float))); }
        h->cap_state = T;
    }
    if (need_k && T > h->cap_k) {
        for (int i = 0; i < 7; ++i) { if (h->kbuf[i]) CK(h, cudaFree(h->kbuf[i])); h->kbuf[i] = nullptr; CK(h, cudaMalloc(&h->kbuf[i], T * sizeof(float))); }
        if (h->ybuf2) CK(h, cudaFree(h->ybuf2)); h->ybuf2 = nullptr; CK(h, cudaMalloc(&h->ybuf2, T * sizeof(float)));
        h->cap_k = T;
    }
    return 0;
}

// Stage descriptors live at sp_dev[pass * stage_stride + k]; captured graphs bake that base
// pointer, so the stride only ever grows (and growing it drops the graphs).
int ensure_stages(SrhepHandle* h, size_t nst) {
    const size_t np = std::max<size_t>(h->passes.size(), 1);
    size_t stride = h->stage_stride;
    if (nst > stride) stride = std::max<size_t>(nst, std::max<size_t>(2 * stride, 64));
    if (stride != h->stage_stride || np * stride > h->cap_stages) {
        if (h->sp_dev) CK(h, cudaFree(h->sp_dev)); h->sp_dev = nullptr;
        if (h->sp_host) CK(h, cudaFreeHost(h->sp_host)); h->sp_host = nullptr;
        CK(h, cudaMalloc(&h->sp_dev, np * stride * sizeof(StageParams)));
        CK(h, cudaMallocHost(&h->sp_host, np * stride * sizeof(StageParams)));
        h->cap_stages = np * stride; h->stage_stride = stride;
        drop_graphs(h);
    }
    if (np > h->cap_stage_idx) {
        if (h->stage_idx_dev) CK(h, cudaFree(h->stage_idx_dev)); h->stage_idx_dev = nullptr;
        CK(h, cudaMalloc(&h->stage_idx_dev, np * sizeof(int)));
        h->cap_stage_idx = np;
        drop_graphs(h);      // captured graphs reference the old counter array
    }
    return 0;
}

// Captures (once per binding and pass) the evaluation of pass `pi` reading sp[stage_idx].
int get_graph(SrhepHandle* h, int pi, const StageParams* sp, int* idx) {
    Pass& p = h->passes[pi];
    if (p.exec) return 0;
    CK(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
    Engine E{h, h->cap_stream};
    const uint64_t before = h->launches;
    StageRef st; st.sp = sp; st.idx = idx;
    E.enqueue_eval(p, st);
    E.bump(idx);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
    p.graph_nodes = (int)(h->launches - before);
    h->launches = before;
    if (E.rc) { if (g) cudaGraphDestroy(g); return E.rc; }
    if (e != cudaSuccess) return fail(h, SRHEP_E_CUDA, "graph capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&p.exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(h, SRHEP_E_CUDA, "graph instantiate: %s", cudaGetErrorString(e));
    return 0;
}

// All passes, one evaluation each, direct launches: v = f(t, x) on the whole bound batch.
int eval_all(SrhepHandle* h, cudaStream_t s, const float* x, float t, const float* t_event, float* v,
             const float* base, float coef, float* out, Profiler* prof = nullptr) {
    Engine E{h, s};
    E.prof = prof;
    for (const Pass& p : h->passes) {
        StageRef st;
        st.fixed.t = t; st.fixed.coef = coef;
        st.fixed.x_in = x + p.r0;
        st.fixed.base = base ? base + p.r0 : nullptr;
        st.fixed.out = out ? out + p.r0 : nullptr;
        st.fixed.vout = v ? v + p.r0 : nullptr;
        st.t_event = t_event;
        E.enqueue_eval(p, st);
        if (E.rc) return E.rc;
    }
    return 0;
}

#include "dopri5.inl"

}  // namespace

// ========================================================================================
// C ABI
// ========================================================================================
extern "C" {

#ifdef SRHEP_BOUNDS
const char* srhep_version(void) { return "srhep 0.1 sm_100a bounds"; }      // the index-asserting debug build (common.cuh: SRHEP_CHECK)
#else
const char* srhep_version(void) { return "srhep 0.1 sm_100a"; }
#endif

size_t srhep_weight_count(const SrhepDims* dims) {
    if (!dims || dims->layers <= 0 || dims->layers > 1024) return 0;
    return make_layout(*dims).total;
}

const char* srhep_last_error(const SrhepHandle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

uint64_t srhep_launch_count(const SrhepHandle* h) { return h ? h->launches : 0; }

int srhep_create(int device, const SrhepDims* dims, const float* weights_host, size_t n_floats, int precision, SrhepHandle** out) {
    if (!dims || !weights_host || !out) return fail(nullptr, SRHEP_E_INVALID, "null argument");
    *out = nullptr;
    if (precision != SRHEP_PREC_FP32 && precision != SRHEP_PREC_BF16 && precision != SRHEP_PREC_FP16)
        return fail(nullptr, SRHEP_E_INVALID, "precision must be SRHEP_PREC_FP32, SRHEP_PREC_BF16 or SRHEP_PREC_FP16");
    int rc = validate_dims(*dims);
    if (rc) return rc;
    Layout L = make_layout(*dims);
    if (n_floats != L.total) return fail(nullptr, SRHEP_E_INVALID, "weight blob has %zu floats, dims need %zu", n_floats, L.total);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return fail(nullptr, SRHEP_E_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(nullptr, SRHEP_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail(nullptr, SRHEP_E_CUDA, "%s", cudaGetErrorString(ce));
    if (prop.major != 10) return fail(nullptr, SRHEP_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail(nullptr, SRHEP_E_CUDA, "%s", cudaGetErrorString(ce));

    SrhepHandle* h = new (std::nothrow) SrhepHandle();
    if (!h) return fail(nullptr, SRHEP_E_NOMEM, "host allocation failed");
    h->device = device; h->d = *dims; h->precision = h->precision_req = precision; h->L = L;
    {   // fp32-grade arithmetic on the tensor cores unless the CUDA-core reference-order path is asked for (diagnostics, A/B tests)
        const char* v = getenv("SRHEP_FP32_SIMT");
        const bool simt = v && *v && *v != '0';
        if (precision == SRHEP_PREC_FP32 && !simt && split_supported(*dims)) { h->split = true; h->precision = SRHEP_PREC_FP16; }
    }
    const SrhepDims& d = h->d;
    h->mod_width = 6 * d.h_dim * d.layers + 2 * d.v_in;
    h->pass_tokens = 0;

    auto cleanup = [&](int code) { g_create_error = h->err; srhep_destroy(h); return code; };
#define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(h, SRHEP_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); return cleanup(e_ == cudaErrorMemoryAllocation ? SRHEP_E_NOMEM : SRHEP_E_CUDA); } } while (0)
    CKC(cudaMalloc(&h->w, L.total * sizeof(float)));
    CKC(cudaMemcpy(h->w, weights_host, L.total * sizeof(float), cudaMemcpyHostToDevice));
    const int H = d.h_dim;
    {   // q|k|v stacked per layer
        std::vector<float> wq((size_t)d.layers * 3 * H * H), bq((size_t)d.layers * 3 * H);
        for (int l = 0; l < d.layers; ++l) {
            const Lin* ls[3] = {&L.layers[l].q, &L.layers[l].k, &L.layers[l].v};
            for (int j = 0; j < 3; ++j) {
                memcpy(&wq[((size_t)l * 3 + j) * H * H], weights_host + ls[j]->w, (size_t)H * H * sizeof(float));
                memcpy(&bq[((size_t)l * 3 + j) * H], weights_host + ls[j]->b, (size_t)H * sizeof(float));
            }
        }
        CKC(cudaMalloc(&h->wqkv, wq.size() * sizeof(float)));
        CKC(cudaMalloc(&h->bqkv, bq.size() * sizeof(float)));
        CKC(cudaMemcpy(h->wqkv, wq.data(), wq.size() * sizeof(float), cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(h->bqkv, bq.data(), bq.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    {   // all adaLN Linears stacked: layer l rows [l*6H, (l+1)*6H), head rows after
        std::vector<float> wm((size_t)h->mod_width * d.ctx), bm(h->mod_width);
        for (int l = 0; l < d.layers; ++l) {
            memcpy(&wm[(size_t)l * 6 * H * d.ctx], weights_host + L.layers[l].ada.w, (size_t)6 * H * d.ctx * sizeof(float));
            memcpy(&bm[(size_t)l * 6 * H], weights_host + L.layers[l].ada.b, (size_t)6 * H * sizeof(float));
        }
        memcpy(&wm[(size_t)d.layers * 6 * H * d.ctx], weights_host + L.vada.w, (size_t)2 * d.v_in * d.ctx * sizeof(float));
        memcpy(&bm[(size_t)d.layers * 6 * H], weights_host + L.vada.b, (size_t)2 * d.v_in * sizeof(float));
        CKC(cudaMalloc(&h->wmod, wm.size() * sizeof(float)));
        CKC(cudaMalloc(&h->bmod, bm.size() * sizeof(float)));
        CKC(cudaMalloc(&h->mod_tbias, bm.size() * sizeof(float)));
        CKC(cudaMemcpy(h->wmod, wm.data(), wm.size() * sizeof(float), cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(h->bmod, bm.data(), bm.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    {   // r1[j] = sum_k w1[j, d + k] over the time-embedding columns (LayerNorm mean term)
        std::vector<float> r(4 * kMaxHid, 0.f);
        const Lin* l1[4] = {&L.eta1, &L.lay1, &L.prx1, &L.nsy1};
        const int dd[4] = {d.etaphi_in, d.layer_emb_dim, 1, 1};
        for (int n = 0; n < 4; ++n)
            for (int j = 0; j < l1[n]->out; ++j) {
                double sacc = 0;
                for (int k = dd[n]; k < l1[n]->in; ++k) sacc += weights_host[l1[n]->w + (size_t)j * l1[n]->in + k];
                r[n * kMaxHid + j] = (float)sacc;
            }
        CKC(cudaMalloc(&h->r1, r.size() * sizeof(float)));
        CKC(cudaMemcpy(h->r1, r.data(), r.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    if (is_lp(h)) {
        rc = bf16_pack_weights(h, weights_host);
        if (rc) return cleanup(rc);
    }
    CKC(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    CKC(cudaEventCreateWithFlags(&h->sp_done, cudaEventDisableTiming));
    CKC(cudaMalloc(&h->red_dev, 8 * sizeof(double)));
    CKC(cudaMallocHost(&h->red_host, 8 * sizeof(double)));
    CKC(cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
#undef CKC
    *out = h;
    return SRHEP_OK;
}

int srhep_destroy(SrhepHandle* h) {
    if (!h) return SRHEP_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    drop_graphs(h);
    void* ptrs[] = {h->w, h->wqkv, h->bqkv, h->wmod, h->bmod, h->mod_tbias, h->r1, h->cu_dev, h->row_event, h->chunk_event, h->chunk_row, h->chunk_len,
                    h->ev_chunk_start, h->attn_work, h->temb, h->ev_a, h->ev_stats, h->layer_out, h->ctx, h->silu_ctx, h->mod, h->modpq, h->f0bias,
                    h->partial, h->t_fill, h->tok_feat, h->xres, h->qkv, h->h1buf, h->act_a, h->act_b, h->qkv_lp, h->qkv_lo, h->act_b_lo, h->y_a, h->y_b, h->y_tmp,
                    h->ybuf2, h->red_dev, h->sp_dev, h->stage_idx_dev, h->tap_layers, h->tap_feat0, h->tap_final};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (float* p : h->kbuf) if (p) cudaFree(p);
    bf16_free_weights(h);
    for (void* p : {(void*)h->dp_ctl, (void*)h->dp_sp, (void*)h->dp_idx, (void*)h->dp_tg}) if (p) cudaFree(p);
    if (h->dp_ctl_host) cudaFreeHost(h->dp_ctl_host);
    if (h->dp_body_stream) cudaStreamDestroy(h->dp_body_stream);
    if (h->sp_host) cudaFreeHost(h->sp_host);
    if (h->red_host) cudaFreeHost(h->red_host);
    if (h->sp_done) cudaEventDestroy(h->sp_done);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    delete h;
    return SRHEP_OK;
}

int srhep_set_pass_tokens(SrhepHandle* h, int64_t max_tokens) {
    if (!h) return SRHEP_E_INVALID;
    if (max_tokens < 0) return fail(h, SRHEP_E_INVALID, "max_tokens < 0");
    h->pass_tokens = max_tokens;
    return SRHEP_OK;
}

int srhep_set_use_graph(SrhepHandle* h, int enable) {
    if (!h) return SRHEP_E_INVALID;
    h->use_graph = enable ? 1 : 0;
    return SRHEP_OK;
}

int srhep_set_debug(SrhepHandle* h, int enable) {
    if (!h) return SRHEP_E_INVALID;
    h->debug = enable ? 1 : 0;
    drop_graphs(h);
    return SRHEP_OK;
}

int srhep_bind_events(SrhepHandle* h, const SrhepCond* c, const int32_t* cu, int32_t B, void* stream) {
    NvtxRange nvtx_api("srhep_bind_events");
    if (!h) return SRHEP_E_INVALID;
    if (!c || !cu || B < 0) return fail(h, SRHEP_E_INVALID, "null argument / negative event count");
    if (cu[0] != 0) return fail(h, SRHEP_E_INVALID, "cu_seqlens[0] must be 0");
    for (int i = 0; i < B; ++i) if (cu[i + 1] < cu[i]) return fail(h, SRHEP_E_INVALID, "cu_seqlens must be non-decreasing (event %d)", i);
    const int64_t T = cu[B];
    if (T > 0 && (!c->eta || !c->cosphi || !c->sinphi || !c->e_proxy || !c->layer)) return fail(h, SRHEP_E_INVALID, "null conditioning array");
    CK(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    CK(h, cudaStreamSynchronize(s));          // previous work may still read the maps rebuilt below
    drop_graphs(h);
    read_switches(h);
    h->bound = false; h->cond = *c; h->B = B; h->T = T;
    h->cu_host.assign(cu, cu + B + 1);

    // passes: whole events, at most pass_tokens rows each (an event larger than that gets its own pass)
    const int64_t cap = h->pass_tokens > 0 ? h->pass_tokens : default_pass_tokens(h->precision);
    h->passes.clear();
    std::vector<int> ch_event, ch_row, ch_len, ev_ch(B + 1, 0);
    std::vector<AttnWork> work;
    h->max_pass_rows = 0;
    {
        Pass p; p.e0 = 0; p.r0 = 0; p.c0 = 0; p.w0 = 0;
        for (int e = 0; e < B; ++e) {
            const int n = cu[e + 1] - cu[e];
            if (e > p.e0 && (int64_t)(cu[e + 1] - p.r0) > cap) {
                p.e1 = e; p.r1 = cu[e]; p.c1 = (int)ch_event.size(); p.w1 = (int)work.size();
                h->passes.push_back(p);
                p = Pass(); p.e0 = e; p.r0 = cu[e]; p.c0 = (int)ch_event.size(); p.w0 = (int)work.size();
            }
            ev_ch[e] = (int)ch_event.size();
            for (int o = 0; o < n; o += kChunk) { ch_event.push_back(e); ch_row.push_back(cu[e] + o); ch_len.push_back(std::min(kChunk, n - o)); }
            for (int o = 0; o < n; o += 128) work.push_back(AttnWork{cu[e] - p.r0 + o, std::min(128, n - o), cu[e] - p.r0, n});
        }
        ev_ch[B] = (int)ch_event.size();
        p.e1 = B; p.r1 = (int)T; p.c1 = (int)ch_event.size(); p.w1 = (int)work.size();
        if (B > 0) h->passes.push_back(p);
    }
    for (const Pass& p : h->passes) h->max_pass_rows = std::max(h->max_pass_rows, p.r1 - p.r0);
    // Attention work items of a pass are consumed round-robin by the G CTA columns of the attention grid (item w goes to column
    // w mod G).  In event order the columns' totals differ by 2 % (single_e) to 15 % (a few hundred multipart events): longest
    // first, every other round of G reversed (boustrophedon), leaves 0.2 - 2 %.  The sort is stable, so the query tiles of one
    // event stay neighbours and still share their K/V tiles in L2.  Results do not depend on the order (items are independent).
    for (const Pass& p : h->passes) {
        const int n = p.w1 - p.w0;
        const int G = std::max(1, std::min(n, h->sw.ctas_per_sm * 148 / h->d.heads));
        auto first = work.begin() + p.w0, last = work.begin() + p.w1;
        std::stable_sort(first, last, [](const AttnWork& a, const AttnWork& b) { return a.k_len > b.k_len; });
        for (int r0 = G; r0 < n; r0 += 2 * G) std::reverse(first + r0, first + std::min(n, r0 + G));
    }

    int rc;
    if ((rc = alloc_for_binding(h))) return rc;
    if ((rc = ensure(h, h->row_event, h->cap_rows, (size_t)T))) return rc;
    const size_t nch = ch_event.size();
    if (nch > h->cap_chunks || !h->chunk_event) {
        for (int** p : {&h->chunk_event, &h->chunk_row, &h->chunk_len}) { if (*p) CK(h, cudaFree(*p)); *p = nullptr; CK(h, cudaMalloc(p, std::max<size_t>(nch, 1) * sizeof(int))); }
        if (h->partial) CK(h, cudaFree(h->partial)); h->partial = nullptr;
        CK(h, cudaMalloc(&h->partial, std::max<size_t>(nch, 1) * h->d.cond * sizeof(float)));
        h->cap_chunks = nch;
    }
    if ((rc = ensure(h, h->attn_work, h->cap_work, work.size()))) return rc;
    if ((rc = alloc_workspace(h))) return rc;

    CK(h, cudaMemcpyAsync(h->cu_dev, cu, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (nch) {
        CK(h, cudaMemcpyAsync(h->chunk_event, ch_event.data(), nch * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(h, cudaMemcpyAsync(h->chunk_row, ch_row.data(), nch * sizeof(int), cudaMemcpyHostToDevice, s));
        CK(h, cudaMemcpyAsync(h->chunk_len, ch_len.data(), nch * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    CK(h, cudaMemcpyAsync(h->ev_chunk_start, ev_ch.data(), (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (!work.empty()) CK(h, cudaMemcpyAsync(h->attn_work, work.data(), work.size() * sizeof(AttnWork), cudaMemcpyHostToDevice, s));
    if (B > 0 && T > 0) {
        row_event_kernel<<<B, 128, 0, s>>>(h->cu_dev, h->row_event, B);
        CK(h, cudaGetLastError());
        ++h->launches;
    }
    CK(h, cudaStreamSynchronize(s));          // host vectors above go out of scope
    if (h->debug) {
        const size_t R = std::max(h->max_pass_rows, 1);
        if (R > h->cap_tap) {
            for (float** p : {&h->tap_layers, &h->tap_feat0, &h->tap_final}) { if (*p) CK(h, cudaFree(*p)); *p = nullptr; }
            CK(h, cudaMalloc(&h->tap_layers, R * h->d.h_dim * h->d.layers * sizeof(float)));
            CK(h, cudaMalloc(&h->tap_feat0, R * h->d.h_dim * sizeof(float)));
            CK(h, cudaMalloc(&h->tap_final, R * h->d.h_dim * sizeof(float)));
            h->cap_tap = R;
        }
    }
    h->bound = true;
    return SRHEP_OK;
}

int srhep_velocity(SrhepHandle* h, const float* x, const float* t, float* v, void* stream) {
    NvtxRange nvtx_api("srhep_velocity");
    if (!h) return SRHEP_E_INVALID;
    read_switches(h);
    if (!h->bound) return fail(h, SRHEP_E_STATE, "srhep_bind_events must be called first");
    if (h->B == 0) return SRHEP_OK;
    if (!t || (h->T > 0 && (!x || !v))) return fail(h, SRHEP_E_INVALID, "null argument");
    CK(h, cudaSetDevice(h->device));
    return eval_all(h, (cudaStream_t)stream, x, 0.f, t, v, nullptr, 0.f, nullptr);
}

int srhep_sample(SrhepHandle* h, const float* x0, const float* tg, int32_t n_steps, int32_t method, int32_t ret_seq,
                 float* x_seq, int32_t* nfe_out, void* stream) {
    NvtxRange nvtx_api("srhep_sample");
    if (!h) return SRHEP_E_INVALID;
    read_switches(h);
    if (!h->bound) return fail(h, SRHEP_E_STATE, "srhep_bind_events must be called first");
    if (n_steps < 1 || !tg) return fail(h, SRHEP_E_INVALID, "n_steps >= 1 and a time grid are required");
    if (method != SRHEP_EULER && method != SRHEP_MIDPOINT && method != SRHEP_RK4)
        return fail(h, SRHEP_E_INVALID, "srhep_sample handles euler / midpoint / rk4; use srhep_sample_dopri5 for dopri5");
    const size_t T = (size_t)h->T;
    if (nfe_out) *nfe_out = 0;
    if (T == 0) return SRHEP_OK;
    if (!x0 || !x_seq) return fail(h, SRHEP_E_INVALID, "null state pointer");
    CK(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_state(h, method == SRHEP_RK4))) return rc;
    const int nint = n_steps - 1;
    // solution[0] = y0
    if (ret_seq) { if (x_seq != x0) CK(h, cudaMemcpyAsync(x_seq, x0, T * sizeof(float), cudaMemcpyDeviceToDevice, s)); }
    else if (nint == 0) { if (x_seq != x0) CK(h, cudaMemcpyAsync(x_seq, x0, T * sizeof(float), cudaMemcpyDeviceToDevice, s)); return SRHEP_OK; }
    if (nint == 0) return SRHEP_OK;

    // state buffer of grid point j
    auto Y = [&](int j) -> float* {
        if (ret_seq) return x_seq + (size_t)j * T;
        if (j == nint) return x_seq;
        return (j & 1) ? h->y_a : h->y_b;
    };
    const float* y0ptr = ret_seq ? x_seq : x0;
    auto Yin = [&](int j) -> const float* { return j == 0 ? y0ptr : Y(j); };

    if (method == SRHEP_RK4) {
        // torchdiffeq rk4 = 3/8 rule; direct launches with stage combinations
        Engine E{h, s};
        for (int j = 0; j < nint; ++j) {
            const float t0 = tg[j], t1 = tg[j + 1], dt = t1 - t0, third = 1.0f / 3.0f;
            const float* y = Yin(j);
            float *k1 = h->kbuf[0], *k2 = h->kbuf[1], *k3 = h->kbuf[2], *k4 = h->kbuf[3], *yt = h->y_tmp;
            if ((rc = eval_all(h, s, y, t0, nullptr, k1, y, dt * third, yt))) return rc;
            if ((rc = eval_all(h, s, yt, t0 + dt * third, nullptr, k2, nullptr, 0.f, nullptr))) return rc;
            E.combine(y, {{k2, dt}, {k1, -dt * third}}, yt, T);
            if ((rc = eval_all(h, s, yt, t0 + dt * (2.0f / 3.0f), nullptr, k3, nullptr, 0.f, nullptr))) return rc;
            E.combine(y, {{k1, dt}, {k2, -dt}, {k3, dt}}, yt, T);
            if ((rc = eval_all(h, s, yt, t1, nullptr, k4, nullptr, 0.f, nullptr))) return rc;
            E.combine(y, {{k1, dt * 0.125f}, {k2, 3 * dt * 0.125f}, {k3, 3 * dt * 0.125f}, {k4, dt * 0.125f}}, Y(j + 1), T);
            if (E.rc) return E.rc;
        }
        if (nfe_out) *nfe_out = 4 * nint;
        return SRHEP_OK;
    }

    const int spi = method == SRHEP_EULER ? 1 : 2;         // evaluations per interval
    const int nst = spi * nint;
    const size_t np = h->passes.size();
    if ((rc = ensure_stages(h, (size_t)nst))) return rc;
    const size_t stride = h->stage_stride;
    CK(h, cudaEventSynchronize(h->sp_done));                 // previous upload finished reading sp_host
    for (size_t pi = 0; pi < np; ++pi) {
        const Pass& p = h->passes[pi];
        StageParams* sp = h->sp_host + pi * stride;
        for (int j = 0; j < nint; ++j) {
            const float t0 = tg[j], t1 = tg[j + 1], dt = t1 - t0;
            const float* y = Yin(j) + p.r0;
            float* yn = Y(j + 1) + p.r0;
            if (method == SRHEP_EULER) {
                sp[j] = StageParams{t0, dt, y, y, yn, nullptr};
            } else {
                const float half = 0.5f * dt;
                float* yt = h->y_tmp + p.r0;
                sp[2 * j] = StageParams{t0, half, y, y, yt, nullptr};
                sp[2 * j + 1] = StageParams{t0 + half, dt, yt, y, yn, nullptr};
            }
        }
    }
    CK(h, cudaMemcpyAsync(h->sp_dev, h->sp_host, np * stride * sizeof(StageParams), cudaMemcpyHostToDevice, s));
    CK(h, cudaEventRecord(h->sp_done, s));
    CK(h, cudaMemsetAsync(h->stage_idx_dev, 0, np * sizeof(int), s));
    // pass-outer, stage-inner: events are independent under a fixed grid, and one pass's
    // activations stay cache-resident across its whole trajectory.
    for (size_t pi = 0; pi < np; ++pi) {
        Pass& p = h->passes[pi];
        if (p.e1 == p.e0 || p.r1 == p.r0) continue;
        const StageParams* sp = h->sp_dev + pi * stride;
        int* idx = h->stage_idx_dev + pi;
        if (h->use_graph) {
            if ((rc = get_graph(h, (int)pi, sp, idx))) return rc;
            for (int k = 0; k < nst; ++k) CK(h, cudaGraphLaunch(p.exec, s));
            h->launches += (uint64_t)nst * p.graph_nodes;
        } else {
            Engine E{h, s};
            StageRef st; st.sp = sp; st.idx = idx;
            for (int k = 0; k < nst; ++k) { E.enqueue_eval(p, st); E.bump(idx); if (E.rc) return E.rc; }
        }
    }
    if (nfe_out) *nfe_out = nst;
    return SRHEP_OK;
}


// torchdiffeq's adaptive dopri5 (RKAdaptiveStepsizeODESolver with the Dormand-Prince tableau,
// restated in oracle/odeint.py and SURVEY Appendix C).  Device-resident by default (dopri5.inl,
// kernels_ode.cuh); the host-driven loop below is the use_graph = 0 variant.  Norms run over real
// cells only (the reference's include the padded slots, SURVEY 7 "Solver semantics").
int srhep_sample_dopri5(SrhepHandle* h, const float* x0, const float* tg, int32_t n_steps, float atol, float rtol,
                        int32_t ret_seq, float* x_seq, int32_t* stats_out, void* stream) {
    NvtxRange nvtx_api("srhep_sample_dopri5");
    if (!h) return SRHEP_E_INVALID;
    read_switches(h);
    if (!h->bound) return fail(h, SRHEP_E_STATE, "srhep_bind_events must be called first");
    if (n_steps < 1 || !tg) return fail(h, SRHEP_E_INVALID, "n_steps >= 1 and a time grid are required");
    if (!(atol > 0.f) || !(rtol >= 0.f)) return fail(h, SRHEP_E_INVALID, "atol > 0 and rtol >= 0 required");
    if (stats_out) stats_out[0] = stats_out[1] = stats_out[2] = 0;
    const size_t T = (size_t)h->T;
    if (T == 0) return SRHEP_OK;
    if (!x0 || !x_seq) return fail(h, SRHEP_E_INVALID, "null state pointer");
    for (int j = 1; j < n_steps; ++j) if (!(tg[j] > tg[j - 1])) return fail(h, SRHEP_E_INVALID, "time grid must be strictly increasing");
    CK(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    auto out_ptr = [&](int j) -> float* { return ret_seq ? x_seq + (size_t)j * T : x_seq; };
    if (ret_seq || n_steps == 1) { if (out_ptr(0) != x0) CK(h, cudaMemcpyAsync(out_ptr(0), x0, T * sizeof(float), cudaMemcpyDeviceToDevice, s)); }
    if (n_steps == 1) return SRHEP_OK;
    // Production: the whole adaptive loop runs on the device (one graph launch, no host round trip per step).  With graphs
    // switched off (srhep_set_use_graph(h, 0): diagnostics, A/B tests) the same arithmetic is driven from the host below,
    // which reads one scalar back per attempted step like the reference's torchdiffeq loop.
    if (h->use_graph) return dopri5_device(h, x0, tg, n_steps, atol, rtol, ret_seq, x_seq, stats_out, s);
    if ((rc = ensure_state(h, true))) return rc;
    Engine E{h, s};

    static const double alpha[6] = {1 / 5., 3 / 10., 4 / 5., 8 / 9., 1., 1.};
    static const double beta[6][6] = {
        {1 / 5., 0, 0, 0, 0, 0},
        {3 / 40., 9 / 40., 0, 0, 0, 0},
        {44 / 45., -56 / 15., 32 / 9., 0, 0, 0},
        {19372 / 6561., -25360 / 2187., 64448 / 6561., -212 / 729., 0, 0},
        {9017 / 3168., -355 / 33., 46732 / 5247., 49 / 176., -5103 / 18656., 0},
        {35 / 384., 0, 500 / 1113., 125 / 192., -2187 / 6784., 11 / 84.}};
    static const double c_err[7] = {35 / 384. - 1951 / 21600., 0, 500 / 1113. - 22642 / 50085., 125 / 192. - 451 / 720.,
                                    -2187 / 6784. + 12231 / 42400., 11 / 84. - 649 / 6300., -1. / 60.};
    static const double c_mid[7] = {6025192743. / 30085553152. / 2, 0, 51252292925. / 65400821598. / 2, -2691868925. / 45128329728. / 2,
                                    187940372067. / 1594534317056. / 2, -1776094331. / 19743644256. / 2, 11237099. / 235043384. / 2};
    int nfe = 0, n_acc = 0, n_rej = 0;
    // rms( (a - a2) / (atol + rtol * max(|s1|, |s2|)) )
    auto rms = [&](const float* a, const float* a2, const float* s1, const float* s2, double& out) -> int {
        CK(h, cudaMemsetAsync(h->red_dev, 0, sizeof(double), s));
        const int grid = (int)std::min<size_t>((T + 255) / 256, 148 * 8);
        scaled_sumsq_kernel<<<grid, 256, 0, s>>>(a, a2, s1, s2, atol, rtol, T, h->red_dev);
        CK(h, cudaGetLastError()); ++h->launches;
        CK(h, cudaMemcpyAsync(h->red_host, h->red_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
        CK(h, cudaStreamSynchronize(s));
        out = std::sqrt(h->red_host[0] / (double)T);
        return 0;
    };
    float* k[7]; for (int i = 0; i < 7; ++i) k[i] = h->kbuf[i];
    float* ycur = h->y_a; float* ynew = h->y_b; float* ymid = h->ybuf2; float* ytmp = h->y_tmp;
    CK(h, cudaMemcpyAsync(ycur, x0, T * sizeof(float), cudaMemcpyDeviceToDevice, s));
    double t0 = tg[0];
    if ((rc = eval_all(h, s, ycur, (float)t0, nullptr, k[0], nullptr, 0.f, nullptr))) return rc; ++nfe;
    // initial step size (order 4)
    double dt;
    {
        double d0, d1, d2;
        if ((rc = rms(ycur, nullptr, ycur, nullptr, d0))) return rc;
        if ((rc = rms(k[0], nullptr, ycur, nullptr, d1))) return rc;
        float h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6f : (float)(0.01 * d0 / d1);
        h0 = std::fabs(h0);
        E.combine(ycur, {{k[0], h0}}, ytmp, T); if (E.rc) return E.rc;
        if ((rc = eval_all(h, s, ytmp, (float)(t0 + (double)h0), nullptr, k[1], nullptr, 0.f, nullptr))) return rc; ++nfe;
        if ((rc = rms(k[1], k[0], ycur, nullptr, d2))) return rc;
        d2 = std::fabs(d2 / h0);
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = std::max(1e-6, (double)h0 * 1e-3);
        else h1 = std::pow(0.01 / std::max(d1, d2), 1.0 / 5.0);
        dt = std::min(100.0 * (double)h0, std::fabs(h1));
    }
    double t1 = t0;        // right end of the last accepted step; (t_lo, t1) is the interpolation interval
    double t_lo = t0;
    int next_out = 1;
    const int max_attempts = 1000000;
    for (int attempt = 0; next_out < n_steps; ++attempt) {
        if (attempt >= max_attempts) return fail(h, SRHEP_E_STATE, "dopri5: max_num_steps exceeded");
        const double ta = t1, tb = t1 + dt;
        const float dty = (float)dt, t0y = (float)ta, t1y = (float)tb;
        for (int st = 0; st < 6; ++st) {
            const float ti = alpha[st] == 1.0 ? t1y : t0y + (float)alpha[st] * dty;
            float* yi = st == 5 ? ynew : ytmp;
            CombineParams q; q.base = ycur; q.out = yi; q.n = T; q.nk = 0;
            for (int i = 0; i <= st; ++i) if (beta[st][i] != 0.0) { q.k[q.nk] = k[i]; q.c[q.nk] = (float)beta[st][i] * dty; ++q.nk; }
            for (int i = q.nk; i < 7; ++i) { q.k[i] = nullptr; q.c[i] = 0.f; }
            combine_kernel<<<(int)std::min<size_t>((T + 255) / 256, 148 * 8), 256, 0, s>>>(q);
            CK(h, cudaGetLastError()); ++h->launches;
            if ((rc = eval_all(h, s, yi, ti, nullptr, k[st + 1], nullptr, 0.f, nullptr))) return rc; ++nfe;
        }
        // error estimate and ratio
        {
            CombineParams q; q.base = nullptr; q.out = ytmp; q.n = T; q.nk = 0;
            for (int i = 0; i < 7; ++i) if (c_err[i] != 0.0) { q.k[q.nk] = k[i]; q.c[q.nk] = (float)c_err[i] * dty; ++q.nk; }
            for (int i = q.nk; i < 7; ++i) { q.k[i] = nullptr; q.c[i] = 0.f; }
            combine_kernel<<<(int)std::min<size_t>((T + 255) / 256, 148 * 8), 256, 0, s>>>(q);
            CK(h, cudaGetLastError()); ++h->launches;
        }
        double ratio;
        if ((rc = rms(ytmp, nullptr, ycur, ynew, ratio))) return rc;
        if (!(ratio == ratio)) return fail(h, SRHEP_E_STATE, "dopri5: non-finite error norm (non-finite input or an event with zero cells?)");
        if (ratio <= 1.0) {
            ++n_acc;
            // dense output on [ta, tb]: quartic through y0, y1, y_mid, f0, f1
            bool need_mid = next_out < n_steps && (double)tg[next_out] <= tb;
            if (need_mid) {
                CombineParams q; q.base = ycur; q.out = ymid; q.n = T; q.nk = 0;
                for (int i = 0; i < 7; ++i) if (c_mid[i] != 0.0) { q.k[q.nk] = k[i]; q.c[q.nk] = (float)c_mid[i] * dty; ++q.nk; }
                for (int i = q.nk; i < 7; ++i) { q.k[i] = nullptr; q.c[i] = 0.f; }
                combine_kernel<<<(int)std::min<size_t>((T + 255) / 256, 148 * 8), 256, 0, s>>>(q);
                CK(h, cudaGetLastError()); ++h->launches;
            }
            t_lo = ta; t1 = tb;
            while (next_out < n_steps && (double)tg[next_out] <= t1) {
                const float x = (float)(((double)tg[next_out] - t_lo) / (t1 - t_lo));
                const float x2 = x * x, x3 = x2 * x, x4 = x3 * x;
                // total = y0 + x dt f0 + x^2 c + x^3 b + x^4 a with
                //   a = 2dt(f1-f0) - 8(y1+y0) + 16 ym ; b = dt(5f0-3f1) + 18y0 + 14y1 - 32ym ; c = dt(f1-4f0) - 11y0 - 5y1 + 16ym
                const float cy0 = 1.f - 11.f * x2 + 18.f * x3 - 8.f * x4;
                const float cy1 = -5.f * x2 + 14.f * x3 - 8.f * x4;
                const float cym = 16.f * x2 - 32.f * x3 + 16.f * x4;
                const float cf0 = dty * (x - 4.f * x2 + 5.f * x3 - 2.f * x4);
                const float cf1 = dty * (x2 - 3.f * x3 + 2.f * x4);
                if (ret_seq || next_out == n_steps - 1) {
                    float* o = out_ptr(next_out);
                    if (x == 1.0f) CK(h, cudaMemcpyAsync(o, ynew, T * sizeof(float), cudaMemcpyDeviceToDevice, s));
                    else { E.combine(nullptr, {{ycur, cy0}, {ynew, cy1}, {ymid, cym}, {k[0], cf0}, {k[6], cf1}}, o, T); if (E.rc) return E.rc; }
                }
                ++next_out;
            }
            std::swap(ycur, ynew);
            std::swap(k[0], k[6]);
        } else {
            ++n_rej;
        }
        // step-size controller: safety 0.9, ifactor 10, dfactor 0.2, order 5
        if (ratio == 0.0) dt *= 10.0;
        else {
            const double dfac = ratio < 1.0 ? 1.0 : 0.2;
            dt *= std::min(10.0, std::max(0.9 / std::pow(ratio, 0.2), dfac));
        }
    }
    if (stats_out) { stats_out[0] = nfe; stats_out[1] = n_acc; stats_out[2] = n_rej; }
    return SRHEP_OK;
}


int srhep_profile(SrhepHandle* h, const float* x, float t, float* v, float* ms_by_cat, int32_t* launches_by_cat, void* stream) {
    if (!h) return SRHEP_E_INVALID;
    read_switches(h);
    if (!h->bound) return fail(h, SRHEP_E_STATE, "srhep_bind_events must be called first");
    if (!ms_by_cat || !launches_by_cat || (h->T > 0 && (!x || !v))) return fail(h, SRHEP_E_INVALID, "null argument");
    for (int i = 0; i < SRHEP_NCAT; ++i) { ms_by_cat[i] = 0.f; launches_by_cat[i] = 0; }
    if (h->B == 0) return SRHEP_OK;
    CK(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    Profiler P;
    cudaEvent_t e0;
    CK(h, cudaEventCreate(&e0));
    CK(h, cudaEventRecord(e0, s));
    P.ev.push_back(e0);
    int rc = eval_all(h, s, x, t, nullptr, v, nullptr, 0.f, nullptr, &P);
    cudaError_t ce = cudaStreamSynchronize(s);
    if (!rc && ce != cudaSuccess) rc = fail(h, SRHEP_E_CUDA, "profile sync: %s", cudaGetErrorString(ce));
    if (!rc)
        for (size_t i = 0; i < P.cat.size(); ++i) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, P.ev[i], P.ev[i + 1]) == cudaSuccess && P.cat[i] >= 0 && P.cat[i] < SRHEP_NCAT) {
                ms_by_cat[P.cat[i]] += ms; ++launches_by_cat[P.cat[i]];
            }
        }
    for (cudaEvent_t e : P.ev) cudaEventDestroy(e);
    return rc;
}

int srhep_get_tap(SrhepHandle* h, const char* name, float* out, size_t n_floats, void* stream) {
    if (!h) return SRHEP_E_INVALID;
    if (!h->debug || !h->bound) return fail(h, SRHEP_E_STATE, "srhep_set_debug(h, 1) before bind, then srhep_velocity, then srhep_get_tap");
    if (h->passes.size() != 1) return fail(h, SRHEP_E_STATE, "taps are kept for single-pass bindings only (%zu passes)", h->passes.size());
    if (!name || !out) return fail(h, SRHEP_E_INVALID, "null argument");
    const SrhepDims& d = h->d;
    const size_t B = h->B, T = h->T;
    const float* src = nullptr; size_t n = 0;
    const std::string nm(name);
    if (nm == "time_emb") { src = h->temb; n = B * d.t_emb; }
    else if (nm == "context") { src = h->ctx; n = B * d.ctx; }
    else if (nm == "tok_feat") { src = h->tok_feat; n = T * (d.cond + d.noisy_out); }
    else if (nm == "feat_0") { src = h->tap_feat0; n = T * d.h_dim; }
    else if (nm == "transformer_out") { src = h->tap_final; n = T * d.h_dim; }
    else if (nm == "mod") { src = h->mod; n = B * h->mod_width; }
    else if (nm.rfind("layer_", 0) == 0) {
        const int l = atoi(name + 6);
        if (l < 0 || l >= d.layers) return fail(h, SRHEP_E_INVALID, "no such layer tap %s", name);
        src = h->tap_layers + (size_t)l * h->cap_tap * d.h_dim; n = T * d.h_dim;
    } else return fail(h, SRHEP_E_INVALID, "unknown tap %s", name);
    if (n_floats != n) return fail(h, SRHEP_E_INVALID, "tap %s has %zu floats, caller passed %zu", name, n, n_floats);
    if (n) CK(h, cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SRHEP_OK;
}

}  // extern "C"

#include "pflow.inl"

#include "post.inl"

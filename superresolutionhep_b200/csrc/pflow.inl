// Host side of the pflow C ABI (include/pflow.h), included at the end of srhep.cu (one
// translation unit, so the generic fp32 kernels of kernels_f32.cuh are shared).
//
// Reference behaviour followed (paths relative to the reference repo root):
//   pflow/models/model_pf.py:56-74              SAPF.forward              -> pflow_forward
//   pflow/models/encoder.py:38-58               Encoder.forward           -> steps 1-4
//   pflow/models/cardinality_predictor.py:17-22 CardinalityPredictor      -> step 5
//   pflow/models/kinematics_predictor.py:99-135 KinematicsPredictor       -> steps 6-8
//   pflow/models/kinematics_predictor.py:24-57  AttnKinematicNet          -> step 9
#include "kernels_pflow.cuh"

namespace {

thread_local std::string g_pf_create_error;

struct PfLayer { Lin q, k, v, o, d1, d3, ada; size_t n1w, n1b, n2w, n2b; };
struct PfLayout {
    size_t table = 0; Lin ci0, ci2;
    std::vector<PfLayer> enc, kin;
    size_t enc_fn_w = 0, enc_fn_b = 0, kin_fn_w = 0, kin_fn_b = 0;
    Lin card[PFLOW_MAX_CARD_HIDDEN + 1];
    size_t part_table = 0; Lin part_proj;
    Lin kq, kk;
    size_t total = 0;
};

// Same order as SAPF.state_dict() of the reference (see tests/golden/pflow_pf_hr.pt).
PfLayout pf_make_layout(const PflowDims& d) {
    PfLayout L;
    size_t off = 0;
    const int H = d.h_dim;
    auto lin = [&](Lin& l, int out, int in) { l.out = out; l.in = in; l.w = off; off += (size_t)out * in; l.b = off; off += out; };
    auto vec = [&](size_t& o, size_t n) { o = off; off += n; };
    auto layer = [&](PfLayer& y) {
        lin(y.q, H, H); lin(y.k, H, H); lin(y.v, H, H); lin(y.o, H, H);
        lin(y.d1, H, H); lin(y.d3, H, H);
        vec(y.n1w, H); vec(y.n1b, H); vec(y.n2w, H); vec(y.n2b, H);
        lin(y.ada, 6 * H, H);
    };
    vec(L.table, (size_t)3 * d.layer_emb_dim);
    lin(L.ci0, H, 4 + d.layer_emb_dim); lin(L.ci2, H, H);
    L.enc.resize(d.enc_layers);
    for (auto& y : L.enc) layer(y);
    vec(L.enc_fn_w, H); vec(L.enc_fn_b, H);
    int win = H;
    for (int i = 0; i <= d.card_n_hidden; ++i) {
        const int wout = i < d.card_n_hidden ? d.card_hidden[i] : d.card_out;
        lin(L.card[i], wout, win); win = wout;
    }
    vec(L.part_table, (size_t)d.max_particles * d.part_emb_dim);
    lin(L.part_proj, H, d.part_emb_dim);
    L.kin.resize(d.kin_layers);
    for (auto& y : L.kin) layer(y);
    vec(L.kin_fn_w, H); vec(L.kin_fn_b, H);
    lin(L.kq, H, H); lin(L.kk, H, H);
    L.total = off;
    return L;
}

}  // namespace

struct PflowHandle {
    int device = 0;
    PflowDims d{};
    PflowVarTransform tr[3]{};
    PfLayout L;
    std::string err;
    uint64_t launches = 0;
    float* w = nullptr;                      // fp32 blob as uploaded
    float *wqkv_e = nullptr, *bqkv_e = nullptr;      // [enc_layers][192, 64] q|k|v stacked
    float *wkv_k = nullptr, *bkv_k = nullptr;        // [kin_layers][128, 64] k|v stacked
    float *wmod_e = nullptr, *bmod_e = nullptr, *wmod_k = nullptr, *bmod_k = nullptr;   // all adaLN Linears of a stack, stacked
    float* pe = nullptr;                     // [P, 64] initial particle embeddings
    // workspace
    size_t cap_rows = 0, cap_events = 0, cap_work = 0;
    float *x = nullptr, *a = nullptr, *b = nullptr, *qkv = nullptr, *enc = nullptr, *kproj = nullptr;
    int *cu_dev = nullptr, *row_event = nullptr, *prow_event = nullptr;
    AttnWork* work = nullptr;
    float *ctx = nullptr, *silu = nullptr, *mod_e = nullptr, *g = nullptr, *silu_g = nullptr, *mod_k = nullptr;
    float *px = nullptr, *pa = nullptr, *pb = nullptr, *pq = nullptr;
    uint8_t* part_mask = nullptr;
};

namespace {

int pf_fail(PflowHandle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (h) h->err = buf; else g_pf_create_error = buf;
    return code;
}

#define PCK(h, call)                                                                          \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return pf_fail(h, e_ == cudaErrorMemoryAllocation ? SRHEP_E_NOMEM : SRHEP_E_CUDA, "%s:%d %s: %s", \
                           __FILE__, __LINE__, #call, cudaGetErrorString(e_));                 \
    } while (0)

template <typename T>
int pf_realloc(PflowHandle* h, T*& p, size_t n) {
    if (p) { PCK(h, cudaFree(p)); p = nullptr; }
    PCK(h, cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)));
    return 0;
}

struct PfEngine {
    PflowHandle* h; cudaStream_t s; int rc = 0;
    const float* W(size_t off) const { return h->w + off; }
    void check(const char* what) {
        if (rc) return;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) rc = pf_fail(h, SRHEP_E_CUDA, "launch %s: %s", what, cudaGetErrorString(e));
        ++h->launches;
    }
    void gemm(const float* A, int lda, const float* Wt, int ldw, float* C, int ldc, int M, int N, int K, const GemmEpilogue& ep) {
        if (rc || M <= 0) return;
        dim3 grid((M + 63) / 64, (N + 63) / 64);
        gemm_f32_kernel<float><<<grid, 256, 0, s>>>(A, lda, Wt, ldw, C, ldc, M, N, K, ep);
        check("pf gemm");
    }
    void ln(const float* x, int M, const float* lw, const float* lb, const float* shift, const float* scale, int ld_mod,
            const int* row_event, int second, float* out) {
        if (rc || M <= 0) return;
        LnModParams q;
        q.x = x; q.ldx = kPfH; q.M = M; q.W = kPfH; q.ln_w = lw; q.ln_b = lb; q.shift = shift; q.scale = scale; q.ld_mod = ld_mod;
        q.row_event = row_event; q.second_ln = second;
        ln_mod_kernel<float><<<(M + 7) / 8, 256, 0, s>>>(q, out, kPfH);
        check("pf ln_mod");
    }
    // everything of a DiT layer after the attention: q += gate_msa * out(attn); q += gate_mlp * dense(modulate(LN2(q)))
    // (models/diffusion_transformer.py:47-52; Dense = LN -> Linear -> LeakyReLU -> Linear, no final activation)
    void post_attention(const PfLayer& y, const float* ml, int ld_mod, const int* row_event, float* x, float* attn, float* tmp, int M) {
        const int H = kPfH;
        { GemmEpilogue ep; ep.bias = W(y.o.b); ep.gate = ml + 2 * H; ep.ld_gate = ld_mod; ep.row_event = row_event; ep.resid = x; ep.ld_resid = H;
          gemm(attn, H, W(y.o.w), H, x, H, M, H, H, ep); }
        ln(x, M, W(y.n2w), W(y.n2b), ml + 3 * H, ml + 4 * H, ld_mod, row_event, 1, tmp);
        { GemmEpilogue ep; ep.bias = W(y.d1.b); ep.act = 1;
          gemm(tmp, H, W(y.d1.w), H, attn, H, M, H, H, ep); }
        { GemmEpilogue ep; ep.bias = W(y.d3.b); ep.gate = ml + 5 * H; ep.ld_gate = ld_mod; ep.row_event = row_event; ep.resid = x; ep.ld_resid = H;
          gemm(attn, H, W(y.d3.w), H, x, H, M, H, H, ep); }
    }
};

int pf_validate(const PflowDims& d) {
    auto bad = [&](const char* m) { return pf_fail(nullptr, SRHEP_E_INVALID, "unsupported pflow dims: %s", m); };
    if (d.h_dim != kPfH) return bad("h_dim must be 64");
    if (d.heads != 4) return bad("heads must be 4 (head dim 16)");
    if (d.enc_layers < 1 || d.enc_layers > 32 || d.kin_layers < 1 || d.kin_layers > 32) return bad("1..32 layers per stack");
    if (d.layer_emb_dim < 1 || d.layer_emb_dim > 8) return bad("layer_emb_dim in 1..8");
    if (d.max_particles < 1 || d.max_particles > kPfMaxP) return bad("max_particles in 1..8");
    if (d.part_emb_dim < 1 || d.part_emb_dim > 64) return bad("part_emb_dim in 1..64");
    if (d.card_n_hidden < 0 || d.card_n_hidden > PFLOW_MAX_CARD_HIDDEN) return bad("at most 4 hidden layers in the cardinality head");
    for (int i = 0; i < d.card_n_hidden; ++i) if (d.card_hidden[i] < 1 || d.card_hidden[i] > 128) return bad("cardinality hidden width in 1..128");
    if (d.card_out < 1 || d.card_out > 128) return bad("card_out in 1..128");
    return 0;
}

}  // namespace

extern "C" {

size_t pflow_weight_count(const PflowDims* d) {
    if (!d || pf_validate(*d)) return 0;
    return pf_make_layout(*d).total;
}

const char* pflow_last_error(const PflowHandle* h) { return h ? h->err.c_str() : g_pf_create_error.c_str(); }
uint64_t pflow_launch_count(const PflowHandle* h) { return h ? h->launches : 0; }

int pflow_destroy(PflowHandle* h) {
    if (!h) return SRHEP_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {h->w, h->wqkv_e, h->bqkv_e, h->wkv_k, h->bkv_k, h->wmod_e, h->bmod_e, h->wmod_k, h->bmod_k, h->pe, h->x, h->a, h->b, h->qkv,
                    h->enc, h->kproj, h->cu_dev, h->row_event, h->prow_event, h->work, h->ctx, h->silu, h->mod_e, h->g, h->silu_g, h->mod_k,
                    h->px, h->pa, h->pb, h->pq, h->part_mask};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete h;
    return SRHEP_OK;
}

int pflow_create(int device, const PflowDims* dims, const float* wh, size_t n_floats, const PflowVarTransform* tr, PflowHandle** out) {
    if (!dims || !wh || !tr || !out) return pf_fail(nullptr, SRHEP_E_INVALID, "null argument");
    *out = nullptr;
    int rc = pf_validate(*dims);
    if (rc) return rc;
    PfLayout L = pf_make_layout(*dims);
    if (n_floats != L.total) return pf_fail(nullptr, SRHEP_E_INVALID, "weight blob has %zu floats, dims need %zu", n_floats, L.total);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) return pf_fail(nullptr, SRHEP_E_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return pf_fail(nullptr, SRHEP_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    cudaDeviceProp prop;
    if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return pf_fail(nullptr, SRHEP_E_CUDA, "%s", cudaGetErrorString(ce));
    if (prop.major != 10) return pf_fail(nullptr, SRHEP_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return pf_fail(nullptr, SRHEP_E_CUDA, "%s", cudaGetErrorString(ce));
    PflowHandle* h = new (std::nothrow) PflowHandle();
    if (!h) return pf_fail(nullptr, SRHEP_E_NOMEM, "host allocation failed");
    h->device = device; h->d = *dims; h->L = L;
    for (int i = 0; i < 3; ++i) h->tr[i] = tr[i];
    const PflowDims& d = h->d;
    const int H = d.h_dim;
    auto up = [&](float*& dst, const std::vector<float>& v) -> int {
        PCK(h, cudaMalloc(&dst, std::max<size_t>(v.size(), 1) * sizeof(float)));
        PCK(h, cudaMemcpy(dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
        return 0;
    };
    auto cleanup = [&](int code) { g_pf_create_error = h->err; pflow_destroy(h); return code; };
    std::vector<float> blob(wh, wh + L.total);
    if ((rc = up(h->w, blob))) return cleanup(rc);
    {   // stacked projections and adaLN Linears
        std::vector<float> wq((size_t)d.enc_layers * 3 * H * H), bq((size_t)d.enc_layers * 3 * H);
        std::vector<float> wm((size_t)d.enc_layers * 6 * H * H), bm((size_t)d.enc_layers * 6 * H);
        for (int l = 0; l < d.enc_layers; ++l) {
            const Lin* ls[3] = {&L.enc[l].q, &L.enc[l].k, &L.enc[l].v};
            for (int j = 0; j < 3; ++j) {
                memcpy(&wq[((size_t)l * 3 + j) * H * H], wh + ls[j]->w, (size_t)H * H * sizeof(float));
                memcpy(&bq[((size_t)l * 3 + j) * H], wh + ls[j]->b, H * sizeof(float));
            }
            memcpy(&wm[(size_t)l * 6 * H * H], wh + L.enc[l].ada.w, (size_t)6 * H * H * sizeof(float));
            memcpy(&bm[(size_t)l * 6 * H], wh + L.enc[l].ada.b, (size_t)6 * H * sizeof(float));
        }
        if ((rc = up(h->wqkv_e, wq)) || (rc = up(h->bqkv_e, bq)) || (rc = up(h->wmod_e, wm)) || (rc = up(h->bmod_e, bm))) return cleanup(rc);
        std::vector<float> wk((size_t)d.kin_layers * 2 * H * H), bk((size_t)d.kin_layers * 2 * H);
        std::vector<float> wm2((size_t)d.kin_layers * 6 * H * H), bm2((size_t)d.kin_layers * 6 * H);
        for (int l = 0; l < d.kin_layers; ++l) {
            const Lin* ls[2] = {&L.kin[l].k, &L.kin[l].v};
            for (int j = 0; j < 2; ++j) {
                memcpy(&wk[((size_t)l * 2 + j) * H * H], wh + ls[j]->w, (size_t)H * H * sizeof(float));
                memcpy(&bk[((size_t)l * 2 + j) * H], wh + ls[j]->b, H * sizeof(float));
            }
            memcpy(&wm2[(size_t)l * 6 * H * H], wh + L.kin[l].ada.w, (size_t)6 * H * H * sizeof(float));
            memcpy(&bm2[(size_t)l * 6 * H], wh + L.kin[l].ada.b, (size_t)6 * H * sizeof(float));
        }
        if ((rc = up(h->wkv_k, wk)) || (rc = up(h->bkv_k, bk)) || (rc = up(h->wmod_k, wm2)) || (rc = up(h->bmod_k, bm2))) return cleanup(rc);
    }
    {   // particle_proj(Embedding(arange(P))): the same P rows for every event (kinematics_predictor.py:84-91)
        std::vector<float> pe((size_t)d.max_particles * H);
        for (int p = 0; p < d.max_particles; ++p)
            for (int c = 0; c < H; ++c) {
                float a = wh[L.part_proj.b + c];
                for (int k = 0; k < d.part_emb_dim; ++k) a = fmaf(wh[L.part_proj.w + (size_t)c * d.part_emb_dim + k], wh[L.part_table + (size_t)p * d.part_emb_dim + k], a);
                pe[(size_t)p * H + c] = a;
            }
        if ((rc = up(h->pe, pe))) return cleanup(rc);
    }
    *out = h;
    return SRHEP_OK;
}

int pflow_forward(PflowHandle* h, const PflowCells* c, const int32_t* cu, int32_t B, const uint8_t* part_mask_in, float* logits, int32_t* n_pred,
                  float* kin_pred, float* inc, void* stream) {
    if (!h) return SRHEP_E_INVALID;
    if (!c || !cu || B < 0 || !logits || !kin_pred) return pf_fail(h, SRHEP_E_INVALID, "null argument / negative event count");
    if (cu[0] != 0) return pf_fail(h, SRHEP_E_INVALID, "cu_seqlens[0] must be 0");
    for (int i = 0; i < B; ++i) if (cu[i + 1] < cu[i]) return pf_fail(h, SRHEP_E_INVALID, "cu_seqlens must be non-decreasing (event %d)", i);
    if (B == 0) return SRHEP_OK;
    const int T = cu[B];
    if (T > 0 && (!c->e || !c->eta || !c->cosphi || !c->sinphi || !c->phi || !c->e_raw || !c->eta_raw || !c->layer || !inc))
        return pf_fail(h, SRHEP_E_INVALID, "null cell array");
    PCK(h, cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const PflowDims& d = h->d; const PfLayout& L = h->L;
    const int H = d.h_dim, P = d.max_particles, BP = B * P;
    int rc;
    // ---- workspace
    if ((size_t)T > h->cap_rows || !h->x) {
        PCK(h, cudaStreamSynchronize(s));
        const size_t R = ((size_t)std::max(T, 1) + 1023) / 1024 * 1024;
        if ((rc = pf_realloc(h, h->x, R * H)) || (rc = pf_realloc(h, h->a, R * H)) || (rc = pf_realloc(h, h->b, R * H)) ||
            (rc = pf_realloc(h, h->qkv, R * 3 * H)) || (rc = pf_realloc(h, h->enc, R * H)) || (rc = pf_realloc(h, h->kproj, R * H)) ||
            (rc = pf_realloc(h, h->row_event, R))) return rc;
        h->cap_rows = R;
    }
    if ((size_t)B > h->cap_events || !h->ctx) {
        PCK(h, cudaStreamSynchronize(s));
        const size_t E = ((size_t)B + 255) / 256 * 256;
        if ((rc = pf_realloc(h, h->cu_dev, E + 1)) || (rc = pf_realloc(h, h->ctx, E * H)) || (rc = pf_realloc(h, h->silu, E * H)) ||
            (rc = pf_realloc(h, h->mod_e, E * 6 * H * d.enc_layers)) || (rc = pf_realloc(h, h->g, E * H)) || (rc = pf_realloc(h, h->silu_g, E * H)) ||
            (rc = pf_realloc(h, h->mod_k, E * 6 * H * d.kin_layers)) || (rc = pf_realloc(h, h->px, E * P * H)) || (rc = pf_realloc(h, h->pa, E * P * H)) ||
            (rc = pf_realloc(h, h->pb, E * P * H)) || (rc = pf_realloc(h, h->pq, E * P * H)) || (rc = pf_realloc(h, h->prow_event, E * P)) ||
            (rc = pf_realloc(h, h->part_mask, E * P))) return rc;
        h->cap_events = E;
    }
    std::vector<AttnWork> work;
    for (int e = 0; e < B; ++e) {
        const int n = cu[e + 1] - cu[e];
        for (int o = 0; o < n; o += 128) work.push_back(AttnWork{cu[e] + o, std::min(128, n - o), cu[e], n});
    }
    if (work.size() > h->cap_work || !h->work) {
        PCK(h, cudaStreamSynchronize(s));
        if ((rc = pf_realloc(h, h->work, work.size() + 64))) return rc;
        h->cap_work = work.size() + 64;
    }
    PCK(h, cudaMemcpyAsync(h->cu_dev, cu, (size_t)(B + 1) * sizeof(int), cudaMemcpyHostToDevice, s));
    if (!work.empty()) PCK(h, cudaMemcpyAsync(h->work, work.data(), work.size() * sizeof(AttnWork), cudaMemcpyHostToDevice, s));
    std::vector<int> pre(BP);
    for (int i = 0; i < BP; ++i) pre[i] = i / P;
    PCK(h, cudaMemcpyAsync(h->prow_event, pre.data(), (size_t)BP * sizeof(int), cudaMemcpyHostToDevice, s));
    PCK(h, cudaStreamSynchronize(s));                      // host vectors above go out of scope; also orders workspace reuse
    PfEngine E{h, s};
    if (T > 0) { row_event_kernel<<<B, 128, 0, s>>>(h->cu_dev, h->row_event, B); E.check("row_event"); }

    // ---- 1. cell initialisation, 2. context = masked mean, 3. all adaLN Linears of the encoder in one GEMM
    if (T > 0) {
        PfCellInitParams q;
        q.in = *c; q.M = T; q.emb_dim = d.layer_emb_dim; q.table = E.W(L.table);
        q.w0 = E.W(L.ci0.w); q.b0 = E.W(L.ci0.b); q.w2 = E.W(L.ci2.w); q.b2 = E.W(L.ci2.b); q.out = h->x;
        pf_cell_init_kernel<<<std::min((T + 31) / 32, 148 * 4), 256, 0, s>>>(q); E.check("pf_cell_init");
    }
    pf_event_mean_kernel<<<B, 256, 0, s>>>(h->x, h->cu_dev, h->ctx, h->silu); E.check("pf_event_mean");
    const int ldm_e = 6 * H * d.enc_layers, ldm_k = 6 * H * d.kin_layers;
    { GemmEpilogue ep; ep.bias = h->bmod_e; E.gemm(h->silu, H, h->wmod_e, H, h->mod_e, ldm_e, B, ldm_e, H, ep); }
    // ---- 4. encoder DiT layers (self-attention over the event's cells)
    for (int l = 0; l < d.enc_layers && T > 0; ++l) {
        const PfLayer& y = L.enc[l];
        const float* ml = h->mod_e + (size_t)l * 6 * H;
        E.ln(h->x, T, E.W(y.n1w), E.W(y.n1b), ml, ml + H, ldm_e, h->row_event, 0, h->a);
        { GemmEpilogue ep; ep.bias = h->bqkv_e + (size_t)l * 3 * H;
          E.gemm(h->a, H, h->wqkv_e + (size_t)l * 3 * H * H, H, h->qkv, 3 * H, T, 3 * H, H, ep); }
        if (!E.rc && !work.empty()) {
            dim3 grid((unsigned)work.size(), d.heads);
            attn_f32_kernel<16, float><<<grid, 128, 0, s>>>(h->qkv, 3 * H, h->qkv + H, h->qkv + 2 * H, 3 * H, h->b, H, h->work, 0.25f);
            E.check("pf attn");
        }
        E.post_attention(y, ml, ldm_e, h->row_event, h->x, h->b, h->a, T);
    }
    E.ln(h->x, T, E.W(L.enc_fn_w), E.W(L.enc_fn_b), nullptr, nullptr, 0, h->row_event, 0, h->enc);
    // ---- 5. cardinality head on the masked mean of the encoded cells; part_mask
    pf_event_mean_kernel<<<B, 256, 0, s>>>(h->enc, h->cu_dev, h->g, h->silu_g); E.check("pf_event_mean");
    {
        PfCardParams q;
        q.g = h->g; q.n_hidden = d.card_n_hidden; q.width[0] = H;
        for (int i = 0; i <= d.card_n_hidden; ++i) { q.width[i + 1] = L.card[i].out; q.w[i] = E.W(L.card[i].w); q.b[i] = E.W(L.card[i].b); }
        q.logits = logits; q.n_pred = n_pred; q.part_mask_in = part_mask_in; q.part_mask = h->part_mask; q.P = P;
        pf_cardinality_kernel<<<B, 128, 0, s>>>(q); E.check("pf_cardinality");
    }
    // ---- 6. decoder: particle queries, adaLN from the same masked mean
    pf_bcast_particles_kernel<<<std::min((BP * H + 255) / 256, 148 * 4), 256, 0, s>>>(h->pe, h->px, BP, P); E.check("pf_bcast_particles");
    { GemmEpilogue ep; ep.bias = h->bmod_k; E.gemm(h->silu_g, H, h->wmod_k, H, h->mod_k, ldm_k, B, ldm_k, H, ep); }
    // ---- 7. decoder DiT layers: cross-attention particles -> modulated LN1(cells)
    for (int l = 0; l < d.kin_layers; ++l) {
        const PfLayer& y = L.kin[l];
        const float* ml = h->mod_k + (size_t)l * 6 * H;
        E.ln(h->enc, T, E.W(y.n1w), E.W(y.n1b), ml, ml + H, ldm_k, h->row_event, 0, h->a);
        { GemmEpilogue ep; ep.bias = h->bkv_k + (size_t)l * 2 * H;
          E.gemm(h->a, H, h->wkv_k + (size_t)l * 2 * H * H, H, h->qkv, 2 * H, T, 2 * H, H, ep); }
        { GemmEpilogue ep; ep.bias = E.W(y.q.b); E.gemm(h->px, H, E.W(y.q.w), H, h->pq, H, BP, H, H, ep); }
        if (!E.rc) {
            PfCrossParams q; q.q = h->pq; q.kv = h->qkv; q.cu = h->cu_dev; q.part_mask = h->part_mask; q.P = P; q.out = h->pa;
            pf_cross_attn_kernel<16><<<B, 512, 0, s>>>(q); E.check("pf_cross_attn");
        }
        E.post_attention(y, ml, ldm_k, h->prow_event, h->px, h->pa, h->pb, BP);
    }
    E.ln(h->px, BP, E.W(L.kin_fn_w), E.W(L.kin_fn_b), nullptr, nullptr, 0, h->prow_event, 0, h->pb);
    // ---- 8./9. AttnKinematicNet
    { GemmEpilogue ep; ep.bias = E.W(L.kq.b); E.gemm(h->pb, H, E.W(L.kq.w), H, h->pq, H, BP, H, H, ep); }
    { GemmEpilogue ep; ep.bias = E.W(L.kk.b); E.gemm(h->enc, H, E.W(L.kk.w), H, h->kproj, H, T, H, H, ep); }
    if (!E.rc) {
        PfKinParams q;
        q.qp = h->pq; q.kp = h->kproj; q.cu = h->cu_dev; q.part_mask = h->part_mask; q.P = P; q.T = T;
        q.e_raw = c->e_raw; q.eta_raw = c->eta_raw; q.phi = c->phi;
        for (int i = 0; i < 3; ++i) q.tr[i] = h->tr[i];
        q.inc = inc; q.kin = kin_pred;
        pf_kin_kernel<<<B, 256, 0, s>>>(q); E.check("pf_kin");
    }
    return E.rc;
}

}  // extern "C"

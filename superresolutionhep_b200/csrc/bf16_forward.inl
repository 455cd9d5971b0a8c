// Host side of the bf16 tcgen05 path (included by srhep.cu inside its anonymous namespace):
// weight re-packing into the swizzled shared-memory image, TMA tensor maps over the pass
// workspace, and the layer schedule of FlowModel.forward with bf16 GEMM operands.

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// bf16 row-major [rows, cols] with row pitch ld (elements); box = 64 columns x 128 rows, 128B swizzle
int make_a_tmap(SrhepHandle* h, CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows = kGemmBM, int force_fp16 = 0) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return fail(h, SRHEP_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(m, ((h->precision == SRHEP_PREC_FP16 || force_fp16) ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16), 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, SRHEP_E_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                                       (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return 0;
}

uint16_t f2h(float f) { const __half hv = __float2half_rn(f); uint16_t u; memcpy(&u, &hv, 2); return u; }

uint16_t f2bf(float f) {
    uint32_t u; memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);     // NaN
    u += 0x7fffu + ((u >> 16) & 1);                                               // round to nearest even
    return (uint16_t)(u >> 16);
}

// W fp32 [N, K] (row pitch ldw) -> image [N / BN][kpad / 64][BN rows x 128 B], 128B-swizzled, K zero-padded
// plane = 1: the LOW fp16 plane of the split mode, fp16(w - float(fp16(w)))
void pack_weight(std::vector<uint8_t>& img, size_t off, const float* W, int ldw, int N, int K, int kpad, int BN, bool fp16, int plane = 0, float mul = 1.f) {
    const int nkb = kpad / 64;
    for (int n = 0; n < N; ++n) {
        const int nt = n / BN, nr = n % BN;
        for (int kb = 0; kb < nkb; ++kb) {
            uint8_t* row = img.data() + off + (((size_t)nt * nkb + kb) * BN + nr) * 128;
            for (int c = 0; c < 8; ++c) {
                uint16_t* dst = reinterpret_cast<uint16_t*>(row + ((c ^ (nr & 7)) * 16));
                for (int j = 0; j < 8; ++j) {
                    const int k = kb * 64 + c * 8 + j;
                    const float wv = k < K ? W[(size_t)n * ldw + k] * mul : 0.f;      // mul is a power of two: exact
                    if (plane) { const __half hv = __float2half_rn(wv); dst[j] = f2h(wv - __half2float(hv)); }
                    else dst[j] = k < K ? (fp16 ? f2h(wv) : f2bf(wv)) : (uint16_t)0;
                }
            }
        }
    }
}

int bf16_pack_weights(SrhepHandle* h, const float* wh) {
    const SrhepDims& d = h->d; const Layout& L = h->L; Bf16Weights& bw = h->bw;
    const int H = d.h_dim;
    if (H != 256 || d.mlp_hid != 256) return fail(h, SRHEP_E_INVALID, "bf16 path: h_dim = mlp_hid = 256 expected (got %d, %d)", H, d.mlp_hid);
    if (d.head_h1 != 128 || (d.v_in + d.ctx) % 64) return fail(h, SRHEP_E_INVALID, "bf16 path: head_h1 = 128 and (v_in + ctx) %% 64 == 0 expected");
    if (H / d.heads != 64) return fail(h, SRHEP_E_INVALID, "bf16 path: head dim 64 expected");
    if (d.layers > 64) return fail(h, SRHEP_E_INVALID, "bf16 path: at most 64 layers");
    const int ncol = d.cond + d.noisy_out;
    bw.feat0_kpad = (ncol + 63) / 64 * 64;
    size_t off = 0;
    auto take = [&](size_t n_rows, size_t kpad) { size_t o = off; off += n_rows * kpad * 2; return o; };
    bw.feat0 = take(H, bw.feat0_kpad);
    for (int l = 0; l < d.layers; ++l) { bw.qkv[l] = take(3 * H, H); bw.out[l] = take(H, H); bw.mlp1[l] = take(H, H); bw.mlp2[l] = take(H, H); }
    const int hk = d.v_in + d.ctx;
    bw.head1 = take(d.head_h1, hk);
    bw.head_chain = hk == kHeadK1 && d.head_h1 == kHeadH1 && d.head_h2 == kHeadH2 && d.head_h3 == kHeadH3 && !h->split;      // split mode: the head runs in fp32 on the CUDA cores
    if (bw.head_chain) { bw.head2 = take(kHeadH2, kHeadH1); bw.head3 = take(kHeadH3, kHeadH2); }
    bw.embed_tc = d.etaphi_in == 3 && d.etaphi_hid == 64 && d.proxy_hid == 64 && d.noisy_hid == 64 && d.etaphi_out == 32 && d.proxy_out == 31 &&
                  d.noisy_out == 64 && d.layer_out == 32 && d.t_emb == 64 && d.cond == 96 && !h->split;
    if (bw.embed_tc) bw.embed_w = take(128, 64);
    std::vector<uint8_t> img(off);
    const bool fp16 = h->precision == SRHEP_PREC_FP16;
    // split mode: each matrix is stored as W * 2^s with max |W| 2^s in [256, 512), so that the LOW plane (~2^-12 of the value) is a normal fp16
    // number for every weight that matters (as stored, the weights are ~0.06 and their low planes would be subnormals with an absolute
    // quantum of 6e-8: 5e-7 of the weight); 2^-s goes to the kernel, which applies it to the accumulator (exact)
    auto pow2_scale = [&](const float* W, int ldw, int N, int K) -> float {
        if (!h->split) return 1.f;
        float mx = 0.f;
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) mx = std::max(mx, fabsf(W[(size_t)n * ldw + k]));
        if (!(mx > 0.f) || !std::isfinite(mx)) return 1.f;
        int e; frexpf(mx, &e);                       // mx = f * 2^e, f in [0.5, 1)
        return ldexpf(1.f, 9 - e);                   // mx * 2^(9-e) in [256, 512)
    };
    float m_feat0 = pow2_scale(wh + L.feat0.w, L.feat0.in, H, ncol);
    bw.ws_feat0 = 1.f / m_feat0;
    std::vector<float> m_qkv(3 * d.layers, 1.f), m_out(d.layers, 1.f), m_m1(d.layers, 1.f), m_m2(d.layers, 1.f);
    pack_weight(img, bw.feat0, wh + L.feat0.w, L.feat0.in, H, ncol, bw.feat0_kpad, 256, fp16, 0, m_feat0);
    for (int l = 0; l < d.layers; ++l) {
        const Layout::Layer& y = L.layers[l];
        const Lin* qkv[3] = {&y.q, &y.k, &y.v};
        for (int j = 0; j < 3; ++j) { m_qkv[3 * l + j] = pow2_scale(wh + qkv[j]->w, H, H, H); bw.ws_qkv[l][j] = 1.f / m_qkv[3 * l + j]; }
        m_out[l] = pow2_scale(wh + y.o.w, H, H, H); m_m1[l] = pow2_scale(wh + y.m1.w, H, H, H); m_m2[l] = pow2_scale(wh + y.m2.w, H, H, H);
        bw.ws_out[l] = 1.f / m_out[l]; bw.ws_mlp1[l] = 1.f / m_m1[l]; bw.ws_mlp2[l] = 1.f / m_m2[l];
        for (int j = 0; j < 3; ++j) pack_weight(img, bw.qkv[l] + (size_t)j * H * H * 2, wh + qkv[j]->w, H, H, H, H, 256, fp16, 0, m_qkv[3 * l + j]);
        pack_weight(img, bw.out[l], wh + y.o.w, H, H, H, H, 256, fp16, 0, m_out[l]);
        pack_weight(img, bw.mlp1[l], wh + y.m1.w, H, H, H, H, 256, fp16, 0, m_m1[l]);
        pack_weight(img, bw.mlp2[l], wh + y.m2.w, H, H, H, H, 256, fp16, 0, m_m2[l]);
    }
    // The velocity head ends in a 32-term dot product with cancellation: its operand rounding dominates the error of v (bf16 operands: rel-L2 2e-2
    // on v for 3e-3 on the transformer output).  Every head operand is a LayerNorm output or a weight, far inside fp16 range, so the fused
    // head runs on fp16 operands (8x finer mantissa, same tcgen05 rate) whatever the operand format of the transformer is.
    const float m_head1 = pow2_scale(wh + L.h1.w, hk, d.head_h1, hk);
    bw.ws_head1 = 1.f / m_head1;
    pack_weight(img, bw.head1, wh + L.h1.w, hk, d.head_h1, hk, hk, 128, fp16 || bw.head_chain, 0, m_head1);
    if (bw.head_chain) {
        pack_weight(img, bw.head2, wh + L.h2.w, kHeadH1, kHeadH2, kHeadH1, kHeadH1, kHeadH2, true);
        pack_weight(img, bw.head3, wh + L.h3.w, kHeadH2, kHeadH3, kHeadH2, kHeadH2, kHeadH3, true);
        memcpy(bw.head_b1, wh + L.h1.b, sizeof bw.head_b1); memcpy(bw.head_b2, wh + L.h2.b, sizeof bw.head_b2);
        memcpy(bw.head_b3, wh + L.h3.b, sizeof bw.head_b3); memcpy(bw.head_w4, wh + L.h4.w, sizeof bw.head_w4);
        bw.head_b4 = wh[L.h4.b];
    }
    if (bw.embed_tc) {
        // block-diagonal second Linear of the three per-cell nets: GEMM column g <- its own net's 64 hidden units (column 63 = padding)
        std::vector<float> w2((size_t)128 * 64, 0.f);
        for (int o = 0; o < 32; ++o) memcpy(&w2[(size_t)o * 64], wh + L.eta3.w + (size_t)o * 64, 64 * sizeof(float));
        for (int o = 0; o < 31; ++o) memcpy(&w2[(size_t)(32 + o) * 64], wh + L.prx3.w + (size_t)o * 64, 64 * sizeof(float));
        for (int o = 0; o < 64; ++o) memcpy(&w2[(size_t)(64 + o) * 64], wh + L.nsy3.w + (size_t)o * 64, 64 * sizeof(float));
        pack_weight(img, bw.embed_w, w2.data(), 64, 128, 64, 64, 128, true);
        EmbedTcParams* q = new (std::nothrow) EmbedTcParams();
        if (!q) return fail(h, SRHEP_E_NOMEM, "host allocation failed");
        memset(q, 0, sizeof *q);
        const Lin* l1[3] = {&L.eta1, &L.prx1, &L.nsy1};
        const int dd[3] = {3, 1, 1};
        for (int n = 0; n < 3; ++n)
            for (int j = 0; j < 64; ++j) {
                const float* row = wh + l1[n]->w + (size_t)j * l1[n]->in;
                double sacc = 0;
                for (int k = dd[n]; k < l1[n]->in; ++k) sacc += row[k];
                q->r1[n * 64 + j] = (float)sacc; q->b1[n * 64 + j] = wh[l1[n]->b + j]; q->w0[n * 64 + j] = row[0];
                if (n == 0) { q->w1[j] = row[1]; q->w2[j] = row[2]; }
            }
        for (int o = 0; o < 32; ++o) q->b2[o] = wh[L.eta3.b + o];
        for (int o = 0; o < 31; ++o) q->b2[32 + o] = wh[L.prx3.b + o];
        for (int o = 0; o < 64; ++o) q->b2[64 + o] = wh[L.nsy3.b + o];
        q->te = (float)d.t_emb;
        bw.embed_tpl = q;
        CK(h, cudaFuncSetAttribute(embed_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEmbSmemBytes));
    }
    CK(h, cudaMalloc(&bw.img, img.size()));
    CK(h, cudaMemcpy(bw.img, img.data(), img.size(), cudaMemcpyHostToDevice));
    if (h->split) {      // low planes of the chain's weights at the same offsets
        std::vector<uint8_t> lo(off, 0);
        pack_weight(lo, bw.feat0, wh + L.feat0.w, L.feat0.in, H, ncol, bw.feat0_kpad, 256, true, 1, m_feat0);
        for (int l = 0; l < d.layers; ++l) {
            const Layout::Layer& y = L.layers[l];
            const Lin* qkv[3] = {&y.q, &y.k, &y.v};
            for (int j = 0; j < 3; ++j) pack_weight(lo, bw.qkv[l] + (size_t)j * H * H * 2, wh + qkv[j]->w, H, H, H, H, 256, true, 1, m_qkv[3 * l + j]);
            pack_weight(lo, bw.out[l], wh + y.o.w, H, H, H, H, 256, true, 1, m_out[l]);
            pack_weight(lo, bw.mlp1[l], wh + y.m1.w, H, H, H, H, 256, true, 1, m_m1[l]);
            pack_weight(lo, bw.mlp2[l], wh + y.m2.w, H, H, H, H, 256, true, 1, m_m2[l]);
        }
        pack_weight(lo, bw.head1, wh + L.h1.w, hk, d.head_h1, hk, hk, 128, true, 1, m_head1);
        CK(h, cudaMalloc(&bw.img_lo, lo.size()));
        CK(h, cudaMemcpy(bw.img_lo, lo.data(), lo.size(), cudaMemcpyHostToDevice));
    }
    bw.bytes = img.size();
    bw.bias_layer_stride = 7 * (size_t)H;      // out.b | mlp1.b | mlp2.b | norm1.w | norm1.b | norm2.w | norm2.b
    bw.bias_head1 = bw.bias_layer_stride * d.layers;
    bw.bias_fn = bw.bias_head1 + d.head_h1;
    std::vector<float> b(bw.bias_fn + 2 * (size_t)H + 2 * (size_t)d.v_in);
    for (int l = 0; l < d.layers; ++l) {
        const Layout::Layer& y = L.layers[l];
        // out-projection bias + W_o b_v: the tcgen05 paths project V WITHOUT its bias (every attention row's weights sum to 1, so the value
        // bias reaches the out-projection unchanged) and K without its bias (a key bias adds the same q . b_k to every score of a row: the
        // softmax cancels it).  Exact in real arithmetic; models/attention.py:112-123, 238-265.
        for (int i = 0; i < H; ++i) {
            double acc = wh[y.o.b + i];
            for (int j = 0; j < H; ++j) acc += (double)wh[y.o.w + (size_t)i * H + j] * (double)wh[y.v.b + j];
            b[l * bw.bias_layer_stride + i] = (float)acc;
        }
        memcpy(&b[l * bw.bias_layer_stride + H], wh + y.m1.b, H * sizeof(float));
        memcpy(&b[l * bw.bias_layer_stride + 2 * H], wh + y.m2.b, H * sizeof(float));
        memcpy(&b[l * bw.bias_layer_stride + 3 * H], wh + y.n1w, H * sizeof(float));
        memcpy(&b[l * bw.bias_layer_stride + 4 * H], wh + y.n1b, H * sizeof(float));
        memcpy(&b[l * bw.bias_layer_stride + 5 * H], wh + y.n2w, H * sizeof(float));
        memcpy(&b[l * bw.bias_layer_stride + 6 * H], wh + y.n2b, H * sizeof(float));
    }
    memcpy(&b[bw.bias_head1], wh + L.h1.b, d.head_h1 * sizeof(float));
    memcpy(&b[bw.bias_fn], wh + L.fn_w, H * sizeof(float)); memcpy(&b[bw.bias_fn + H], wh + L.fn_b, H * sizeof(float));
    memcpy(&b[bw.bias_fn + 2 * H], wh + L.nv_w, d.v_in * sizeof(float)); memcpy(&b[bw.bias_fn + 2 * H + d.v_in], wh + L.nv_b, d.v_in * sizeof(float));
    CK(h, cudaMalloc(&bw.bias, b.size() * sizeof(float)));
    CK(h, cudaMemcpy(bw.bias, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice));
    bw.bias_h = (float*)malloc(b.size() * sizeof(float));
    bw.bqkv_h = (float*)malloc((size_t)d.layers * 3 * H * sizeof(float));
    if (!bw.bias_h || !bw.bqkv_h) return fail(h, SRHEP_E_NOMEM, "host allocation failed");
    memcpy(bw.bias_h, b.data(), b.size() * sizeof(float));
    for (int l = 0; l < d.layers; ++l) {
        const Lin* qkv[3] = {&L.layers[l].q, &L.layers[l].k, &L.layers[l].v};
        for (int j = 0; j < 3; ++j) {
            if (j == 0) memcpy(bw.bqkv_h + ((size_t)l * 3 + j) * H, wh + qkv[j]->b, H * sizeof(float));
            else memset(bw.bqkv_h + ((size_t)l * 3 + j) * H, 0, H * sizeof(float));          // K, V: see the out-projection bias above
        }
    }
    CK(h, cudaMalloc(&bw.bqkv_dev, (size_t)d.layers * 3 * H * sizeof(float)));
    CK(h, cudaMemcpy(bw.bqkv_dev, bw.bqkv_h, (size_t)d.layers * 3 * H * sizeof(float), cudaMemcpyHostToDevice));
    CK(h, cudaFuncSetAttribute(gemm_bf16_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_bf16_smem_bytes<256>(4)));
    CK(h, cudaFuncSetAttribute(gemm_bf16_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_bf16_smem_bytes<256>(4)));
    CK(h, cudaFuncSetAttribute(attn_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmemBytes));
    CK(h, cudaFuncSetAttribute(attn2_bf16_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAtt2SmemBytes));
    CK(h, cudaFuncSetAttribute(attn2_bf16_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(attn2_bf16_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAtt2SmemBytes));
    CK(h, cudaFuncSetAttribute(attn2_bf16_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(attn3_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Att3Cfg<false>::kSmemBytes));
    CK(h, cudaFuncSetAttribute(attn3_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(attn3_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Att3Cfg<false>::kSmemBytes));
    CK(h, cudaFuncSetAttribute(attn3_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(head_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmemBytes));
    CK(h, cudaFuncSetAttribute(head_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHFSmemBytes));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytes));
    CK(h, cudaFuncSetAttribute(layer_chain_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CK(h, cudaFuncSetAttribute(gemm_bf16_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_bf16_smem_bytes<128>(hk / 64)));
    if (h->split) {
        CK(h, cudaFuncSetAttribute(layer_chain_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytesSplit));
        CK(h, cudaFuncSetAttribute(layer_chain_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemBytesSplit));
        CK(h, cudaFuncSetAttribute(attn3_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Att3Cfg<true>::kSmemBytes));
        CK(h, cudaFuncSetAttribute(head_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHeadSmemBytesSplit));
    }
    return 0;
}

void bf16_free_weights(SrhepHandle* h) {
    if (h->bw.img) cudaFree(h->bw.img);
    if (h->bw.img_lo) cudaFree(h->bw.img_lo);
    if (h->bw.tok_lp_lo) cudaFree(h->bw.tok_lp_lo);
    if (h->bw.bias) cudaFree(h->bw.bias);
    if (h->bw.bqkv_dev) cudaFree(h->bw.bqkv_dev);
    if (h->bw.tok_lp) cudaFree(h->bw.tok_lp);
    free(h->bw.bias_h); free(h->bw.bqkv_h);
    delete static_cast<EmbedTcParams*>(h->bw.embed_tpl);
    h->bw = Bf16Weights();
}

// (re)build the A-operand tensor maps after the pass workspace was (re)allocated
int bf16_on_bind(SrhepHandle* h) {
    const SrhepDims& d = h->d; Bf16Weights& bw = h->bw;
    const uint64_t R = h->cap_ws_rows;
    if (bw.tok_lp) { CK(h, cudaFree(bw.tok_lp)); bw.tok_lp = nullptr; }
    CK(h, cudaMalloc(&bw.tok_lp, R * bw.feat0_kpad * 2));
    CK(h, cudaMemset(bw.tok_lp, 0, R * bw.feat0_kpad * 2));      // the K padding columns stay zero; the embedding kernel writes the rest
    int rc;
    if ((rc = make_a_tmap(h, &bw.tm_ln, h->act_a, R, d.h_dim, d.h_dim))) return rc;
    if ((rc = make_a_tmap(h, &bw.tm_hin, h->act_a, R, d.v_in + d.ctx, d.v_in + d.ctx, kGemmBM, bw.head_chain ? 1 : 0))) return rc;      // the fused head always runs on fp16 operands
    if ((rc = make_a_tmap(h, &bw.tm_b, h->act_b, R, d.h_dim, d.h_dim))) return rc;
    if ((rc = make_a_tmap(h, &bw.tm_tok, bw.tok_lp, R, bw.feat0_kpad, bw.feat0_kpad))) return rc;
    if ((rc = make_a_tmap(h, &bw.tm_qkv, h->qkv_lp, R, 3 * d.h_dim, 3 * d.h_dim))) return rc;
    if ((rc = make_a_tmap(h, &bw.tm_kv64, h->qkv_lp, R, 3 * d.h_dim, 3 * d.h_dim, kAtt2KvTile))) return rc;
    if (h->split) {
        if (bw.tok_lp_lo) { CK(h, cudaFree(bw.tok_lp_lo)); bw.tok_lp_lo = nullptr; }
        CK(h, cudaMalloc(&bw.tok_lp_lo, R * bw.feat0_kpad * 2));
        CK(h, cudaMemset(bw.tok_lp_lo, 0, R * bw.feat0_kpad * 2));
        if ((rc = make_a_tmap(h, &bw.tm_b_lo, h->act_b_lo, R, d.h_dim, d.h_dim))) return rc;
        if ((rc = make_a_tmap(h, &bw.tm_tok_lo, bw.tok_lp_lo, R, bw.feat0_kpad, bw.feat0_kpad))) return rc;
        if ((rc = make_a_tmap(h, &bw.tm_qkv_lo, h->qkv_lo, R, 3 * d.h_dim, 3 * d.h_dim))) return rc;
        if ((rc = make_a_tmap(h, &bw.tm_kv64_lo, h->qkv_lo, R, 3 * d.h_dim, 3 * d.h_dim, kAtt2KvTile))) return rc;
        // the head operand's low plane lives in the upper half of act_a (R x wide x 4 bytes; the high plane takes R x (v_in + ctx) x 2)
        if ((rc = make_a_tmap(h, &bw.tm_hin_lo, (uint8_t*)h->act_a + R * (size_t)(d.v_in + d.ctx) * 2, R, d.v_in + d.ctx, d.v_in + d.ctx))) return rc;
    }
    return 0;
}

int64_t default_pass_tokens(int precision) { return precision != SRHEP_PREC_FP32 ? (1 << 21) : (1 << 19); }

template <int BN, bool kLN = false>
void launch_gemm_bf16(Engine& E, const CUtensorMap& tm, int M, int K, int N, const uint8_t* w_img, void* C, int ldc, int out_bf16,
                      const GemmEpilogue& ep) {
    if (E.rc || M <= 0) return;
    GemmBf16Params p;
    p.M = M; p.num_kb = K / 64; p.w_img = w_img; p.C = C; p.ldc = ldc; p.out_bf16 = out_bf16; p.ep = ep;
    p.fp16 = E.h->precision == SRHEP_PREC_FP16;
    p.c_blocked = (kLN && !ep.resid && E.x_blocked && out_bf16 == 0) ? 1 : 0;
    const int m_tiles = (M + kGemmBM - 1) / kGemmBM, n_tiles = N / BN;
    dim3 grid(std::max(1, std::min(m_tiles, 148 / n_tiles)), n_tiles);
    static_assert(!kLN || BN == 256, "the fused LayerNorm epilogue maps 256 epilogue threads to 256 columns");
    gemm_bf16_kernel<BN, kLN><<<grid, kGemmThreads, gemm_bf16_smem_bytes<BN>(p.num_kb), E.s>>>(tm, p);
    E.check("gemm_bf16");
}

const float* wh_b1(SrhepHandle* h) { return h->bw.bias_h + h->bw.bias_head1; }      // host copy of the first head bias

void launch_attn_bf16(Engine& E, const Pass& p, __nv_bfloat16* out) {
    if (E.rc || p.w1 == p.w0) return;
    if (E.h->sw.only == 2 || E.h->sw.only == 3) return;
    SrhepHandle* h = E.h;
    const SrhepDims& d = h->d;
    AttnBf16Params q;
    static_assert(sizeof(AttnItem) == sizeof(AttnWork), "work item layout");
    q.items = reinterpret_cast<const AttnItem*>(h->attn_work + p.w0); q.n_items = p.w1 - p.w0;
    q.out = out; q.out_lo = h->split ? (__nv_bfloat16*)h->act_b_lo : nullptr; q.ldo = d.h_dim; q.h_dim = d.h_dim;
    q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
    q.scale_log2 = 1.4426950408889634f / sqrtf((float)(d.h_dim / d.heads));
    q.fp16 = h->precision == SRHEP_PREC_FP16;
    q.dbg = nullptr;
    static long long* adbg = nullptr;
    static int adbg_calls = 0;
    const bool dbg = h->sw.attn_dbg && (++adbg_calls == 8);
    if (dbg) { if (!adbg) cudaMalloc(&adbg, 256 * sizeof(long long)); cudaMemsetAsync(adbg, 0, 256 * sizeof(long long), E.s); q.dbg = adbg; }
    dim3 grid(std::max(1, std::min(q.n_items, h->sw.ctas_per_sm * 148 / d.heads)), d.heads);
    if (h->split) {                                 // fp32-grade: (hi, lo) planes, one CTA per SM
        dim3 g1(std::max(1, std::min(q.n_items, 148 / d.heads)), d.heads);
        attn3_kernel<true, true><<<g1, kAtt3Threads, Att3Cfg<true>::kSmemBytes, E.s>>>(h->bw.tm_qkv, h->bw.tm_kv64, h->bw.tm_qkv_lo, h->bw.tm_kv64_lo, q);
    } else
    if (!h->sw.attn_v1 && !h->sw.attn_v2) {        // third generation: P in tensor memory, producer running ahead across items
        if (q.fp16) attn3_kernel<true, false><<<grid, kAtt3Threads, Att3Cfg<false>::kSmemBytes, E.s>>>(h->bw.tm_qkv, h->bw.tm_kv64, h->bw.tm_qkv, h->bw.tm_kv64, q);
        else attn3_kernel<false, false><<<grid, kAtt3Threads, Att3Cfg<false>::kSmemBytes, E.s>>>(h->bw.tm_qkv, h->bw.tm_kv64, h->bw.tm_qkv, h->bw.tm_kv64, q);
    } else if (!h->sw.attn_v1) {
        if (q.fp16) attn2_bf16_kernel<true><<<grid, kAtt2Threads, kAtt2SmemBytes, E.s>>>(h->bw.tm_qkv, h->bw.tm_kv64, q);
        else attn2_bf16_kernel<false><<<grid, kAtt2Threads, kAtt2SmemBytes, E.s>>>(h->bw.tm_qkv, h->bw.tm_kv64, q);
    }
    else
    attn_bf16_kernel<<<grid, kAttnThreads, kAttnSmemBytes, E.s>>>(h->bw.tm_qkv, q);
    E.check("attn_bf16");
    if (dbg) {
        long long hb[256];
        cudaStreamSynchronize(E.s);
        cudaMemcpy(hb, adbg, sizeof hb, cudaMemcpyDeviceToHost);
        const long long t0 = hb[0];
        for (int i = 0; i < 4; ++i) {
            fprintf(stderr, "[attn dbg] item %d:", i);
            for (int k = 0; k < 64; ++k) if (hb[i * 64 + k]) fprintf(stderr, " %d=%lld", k, hb[i * 64 + k] - t0);
            fprintf(stderr, "\n");
        }
    }
}

// One launch for the row-local part of DiT layer l: out-projection ... q|k|v of layer l + 1 (kernels_chain.cuh)
void launch_chain(Engine& E, int M, int l, const int* rev) {
    if (E.rc || M <= 0) return;
    if (E.h->sw.only == 1 || E.h->sw.only == 3) return;
    SrhepHandle* h = E.h;
    const SrhepDims& d = h->d; Bf16Weights& bw = h->bw;
    const int H = d.h_dim;
    const bool last = l + 1 == d.layers;
    const float* ml = h->mod + (size_t)l * 6 * H;
    ChainParams q{};
    q.M = M; q.n_stages = last ? 3 : 6; q.fp16 = h->precision == SRHEP_PREC_FP16; q.a_early = h->sw.chain_a_early; q.ln_direct = h->sw.chain_ln_direct;
    q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
    q.a_pf = h->sw.chain_a_pf;                                       // SRHEP_CHAIN_A_PF: measured at slots 8 / 24 / 32 / 40, never a gain: off by default
#ifdef SRHEP_BOUNDS
    if (getenv("SRHEP_BOUNDS_SELFTEST")) q.ext.rows_cap = 1;      // tests/test_gpu_bounds.py: proves that a violated extent is caught (the kernel traps)
#endif
    q.row_event = rev; q.x = h->xres;
    q.w[0] = bw.img + bw.out[l]; q.w[1] = bw.img + bw.mlp1[l]; q.w[2] = bw.img + bw.mlp2[l];
    const float* blh = bw.bias_h + l * bw.bias_layer_stride;
    memcpy(q.cst[0], blh, 3 * H * sizeof(float));                      // out.b | mlp1.b | mlp2.b
    const float* pq = h->modpq + (size_t)l * 4 * H;                    // [P_msa | Q_msa | P_mlp | Q_mlp] of layer l (modpq_kernel)
    q.gate_msa = ml + 2 * H; q.p_mlp = pq + 2 * H; q.q_mlp = pq + 3 * H; q.gate_mlp = ml + 5 * H;
    q.ld_mod = h->mod_width; q.ld_pq = d.layers * 4 * H;
    if (!last) {
        for (int j = 0; j < 3; ++j) q.w[3 + j] = bw.img + bw.qkv[l + 1] + (size_t)j * H * H * 2;
        q.qkv_lo = h->qkv_lo;
        memcpy(q.cst[3], bw.bqkv_h + (size_t)(l + 1) * 3 * H, 3 * H * sizeof(float));
        q.p_nxt = pq + 4 * H; q.q_nxt = pq + 5 * H;                        // layer l + 1: norm1 x (scale_msa, shift_msa)
        q.qkv = h->qkv_lp;
    }
    const int m_tiles = (M + 127) / 128;
    const int grid = std::max(1, std::min(m_tiles, (h->split ? 1 : h->sw.ctas_per_sm) * 148));
    if (h->split) for (int g = 0; g < 6; ++g) q.w_lo[g] = q.w[g] ? bw.img_lo + (q.w[g] - bw.img) : nullptr;
    q.wscale[0] = bw.ws_out[l]; q.wscale[1] = bw.ws_mlp1[l]; q.wscale[2] = bw.ws_mlp2[l];
    for (int j = 0; j < 3; ++j) q.wscale[3 + j] = last ? 1.f : bw.ws_qkv[l + 1][j];
    static long long* dbg_dev = nullptr;
    const bool dbg = h->sw.chain_dbg && l == 1;
    if (dbg) { if (!dbg_dev) cudaMalloc(&dbg_dev, 256 * sizeof(long long)); cudaMemsetAsync(dbg_dev, 0, 256 * sizeof(long long), E.s); q.dbg = dbg_dev; }
    if (h->split) layer_chain_kernel<true, false, true><<<grid, kChainThreads, kChainSmemBytesSplit, E.s>>>(bw.tm_b, bw.tm_b_lo, q);
    else if (q.fp16) layer_chain_kernel<true><<<grid, kChainThreads, kChainSmemBytes, E.s>>>(bw.tm_b, bw.tm_b, q);
    else layer_chain_kernel<false><<<grid, kChainThreads, kChainSmemBytes, E.s>>>(bw.tm_b, bw.tm_b, q);
    E.check("layer_chain");
    if (dbg) {
        long long hbuf[256];
        cudaStreamSynchronize(E.s);
        cudaMemcpy(hbuf, dbg_dev, sizeof hbuf, cudaMemcpyDeviceToHost);
        static const char* names[26] = {"P a_free", "P A issued", "M tile start", "M a_full", "M g0", "M g1", "M g2", "M g3", "M g4", "M g5",
                                        "E0 acc", "E0 done", "E1 acc", "E1 done", "E2 acc", "E2 done", "E3 acc", "E3 done", "E4 acc", "E4 done", "E5 acc", "E5 done", "E0 passA", "E0 bar1", "E0 passB", "E0 bar3"};
        const long long t0 = hbuf[1];
        for (int t = 0; t < 4; ++t) {
            fprintf(stderr, "[chain dbg] tile %d (cycles since first A issue):", t);
            for (int k = 0; k < 26; ++k) fprintf(stderr, " %s=%lld", names[k], hbuf[t * 32 + k] ? hbuf[t * 32 + k] - t0 : -1);
            fprintf(stderr, "\n");
        }
    }
}

// feat_0 -> LN1 . modulate of layer 0 -> q|k|v of layer 0 in ONE launch: the chain kernel in its first-layer mode (stages
// feat_0, q, k, v).  Replaces the feat_0 GEMM and the layer-0 q|k|v GEMM: the LayerNorm output never travels through HBM.
void launch_chain_first(Engine& E, int M, const int* rev) {
    if (E.rc || M <= 0) return;
    SrhepHandle* h = E.h;
    const SrhepDims& d = h->d; Bf16Weights& bw = h->bw;
    const int H = d.h_dim;
    ChainParams q{};
    q.M = M; q.n_stages = 4; q.fp16 = h->precision == SRHEP_PREC_FP16; q.a_early = h->sw.chain_a_early; q.ln_direct = h->sw.chain_ln_direct;
    q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
    q.row_event = rev; q.x = h->xres;
    q.a_pf = h->sw.chain_a_pf;
    q.w[0] = bw.img + bw.feat0;
    for (int j = 0; j < 3; ++j) q.w[1 + j] = bw.img + bw.qkv[0] + (size_t)j * H * H * 2;
    // cst[2] (bias of the residual-producing stage) stays zero: feat_0's bias is inside the per-event rows
    memcpy(q.cst[3], bw.bqkv_h, 3 * H * sizeof(float));
    q.row_bias = h->f0bias; q.ld_row_bias = H;
    q.p_nxt = h->modpq; q.q_nxt = h->modpq + H;                        // layer 0: norm1 x (scale_msa, shift_msa)
    q.gate_msa = q.gate_mlp = h->mod; q.p_mlp = q.q_mlp = h->modpq;    // unused in this mode
    q.ld_mod = h->mod_width; q.ld_pq = d.layers * 4 * H;
    q.qkv = h->qkv_lp; q.qkv_lo = h->qkv_lo;
    const int m_tiles = (M + 127) / 128;
    const int grid = std::max(1, std::min(m_tiles, (h->split ? 1 : h->sw.ctas_per_sm) * 148));
    if (h->split) for (int g = 0; g < 6; ++g) q.w_lo[g] = q.w[g] ? bw.img_lo + (q.w[g] - bw.img) : nullptr;
    q.wscale[0] = bw.ws_feat0; for (int j = 0; j < 3; ++j) q.wscale[1 + j] = bw.ws_qkv[0][j];
    if (h->split) layer_chain_kernel<true, true, true><<<grid, kChainThreads, kChainSmemBytesSplit, E.s>>>(bw.tm_tok, bw.tm_tok_lo, q);
    else if (q.fp16) layer_chain_kernel<true, true><<<grid, kChainThreads, kChainSmemBytes, E.s>>>(bw.tm_tok, bw.tm_tok, q);
    else layer_chain_kernel<false, true><<<grid, kChainThreads, kChainSmemBytes, E.s>>>(bw.tm_tok, bw.tm_tok, q);
    E.check("layer_chain_first");
}

void bf16_forward(Engine& E, const Pass& p, const int* rev, const StageRef& st) {
    SrhepHandle* h = E.h;
    const SrhepDims& d = h->d; const Layout& L = h->L; Bf16Weights& bw = h->bw;
    const int M = p.r1 - p.r0, H = d.h_dim;
    const int fp16 = h->precision == SRHEP_PREC_FP16;
    float* x = h->xres;
    const float* mod = h->mod;
    const bool chain = (H == kChainH && d.mlp_hid == kChainH) && (!h->sw.no_chain || h->split);      // split mode exists for the chain path only
    E.x_blocked = chain;
    __nv_bfloat16* a = (__nv_bfloat16*)h->act_a; __nv_bfloat16* b = (__nv_bfloat16*)h->act_b;
    __nv_bfloat16* qkv = (__nv_bfloat16*)h->qkv_lp;
    E.cat = SRHEP_CAT_FEAT0;      // the 16-bit A operand of feat_0 was written by the embedding kernel
    // LayerNorm + adaLN modulate of the freshly produced residual row is fused into the epilogue of the GEMM
    // that produces it: ln1 of layer l rides on feat_0 (l = 0) / the previous layer's MLP2, ln2 on the out-projection.
    auto with_ln = [&](GemmEpilogue& ep, int layer, bool second) {
        const float* ml = mod + (size_t)layer * 6 * H;
        const float* bl = bw.bias + layer * bw.bias_layer_stride;      // 16-byte aligned copies of the LayerNorm affine
        ep.ln_out = a; ep.ld_ln = H; ep.ld_lnmod = h->mod_width; ep.ln_second = second ? 1 : 0;
        if (!second) { ep.ln_w = bl + 3 * H; ep.ln_b = bl + 4 * H; ep.ln_shift = ml; ep.ln_scale = ml + H; }
        else { ep.ln_w = bl + 5 * H; ep.ln_b = bl + 6 * H; ep.ln_shift = ml + 3 * H; ep.ln_scale = ml + 4 * H; }
    };
    const bool first_fused = chain && bw.feat0_kpad == 192 && (!h->sw.no_chain_first || h->split);
    const bool rest = h->sw.only == 0 || h->sw.only == 3;
    if (!rest) { }
    else if (first_fused) launch_chain_first(E, M, rev);
    else
    { GemmEpilogue ep; ep.row_bias = h->f0bias; ep.ld_row_bias = H; ep.row_event = rev; ep.act = 1;
      with_ln(ep, 0, false);
      launch_gemm_bf16<256, true>(E, bw.tm_tok, M, bw.feat0_kpad, H, bw.img + bw.feat0, x, H, 0, ep); }
    E.tap(h->tap_feat0, x, M);
    if (chain) {
        // layer 0's q|k|v come from the feat_0 GEMM's fused LN1; every later projection rides in the previous layer's chain kernel
        E.cat = SRHEP_CAT_QKV;
        if (!first_fused)
        { GemmEpilogue ep; ep.bias = bw.bqkv_dev;
          launch_gemm_bf16<256>(E, bw.tm_ln, M, H, 3 * H, bw.img + bw.qkv[0], qkv, 3 * H, 1, ep); }
        for (int l = 0; l < d.layers; ++l) {
            E.cat = SRHEP_CAT_ATTN;
            if (!fp16 && h->sw.attn_simt) E.attention_simt<__nv_bfloat16>(p, qkv, b);
            else launch_attn_bf16(E, p, b);
            E.cat = SRHEP_CAT_CHAIN;
            launch_chain(E, M, l, rev);
            if (h->debug && h->tap_layers) E.tap(h->tap_layers + (size_t)l * h->cap_tap * H, x, M);
        }
    } else
    for (int l = 0; l < d.layers; ++l) {
        const float* ml = mod + (size_t)l * 6 * H;
        const float* bl = bw.bias + l * bw.bias_layer_stride;
        E.cat = SRHEP_CAT_QKV;
        { GemmEpilogue ep; ep.bias = bw.bqkv_dev + (size_t)l * 3 * H;
          launch_gemm_bf16<256>(E, bw.tm_ln, M, H, 3 * H, bw.img + bw.qkv[l], qkv, 3 * H, 1, ep); }
        E.cat = SRHEP_CAT_ATTN;
        if (!fp16 && h->sw.attn_simt) E.attention_simt<__nv_bfloat16>(p, qkv, b);
        else launch_attn_bf16(E, p, b);
        E.cat = SRHEP_CAT_OUT;
        { GemmEpilogue ep; ep.bias = bl; ep.gate = ml + 2 * H; ep.ld_gate = h->mod_width; ep.row_event = rev; ep.resid = x; ep.ld_resid = H;
          if (h->sw.no_lnfuse) {
              launch_gemm_bf16<256>(E, bw.tm_b, M, H, H, bw.img + bw.out[l], x, H, 0, ep);
              const Layout::Layer& y = L.layers[l];
              E.cat = SRHEP_CAT_LN;
              E.ln_mod<__nv_bfloat16>(x, M, H, E.W(y.n2w), E.W(y.n2b), ml + 3 * H, ml + 4 * H, rev, 1, a);
          } else {
          with_ln(ep, l, true);
          launch_gemm_bf16<256, true>(E, bw.tm_b, M, H, H, bw.img + bw.out[l], x, H, 0, ep); } }
        E.cat = SRHEP_CAT_MLP1;
        { GemmEpilogue ep; ep.bias = bl + H; ep.act = 1;
          launch_gemm_bf16<256>(E, bw.tm_ln, M, H, H, bw.img + bw.mlp1[l], b, H, 1, ep); }
        E.cat = SRHEP_CAT_MLP2;
        { GemmEpilogue ep; ep.bias = bl + 2 * H; ep.act = 1; ep.gate = ml + 5 * H; ep.ld_gate = h->mod_width; ep.row_event = rev; ep.resid = x; ep.ld_resid = H;
          if (l + 1 < d.layers) { with_ln(ep, l + 1, false); launch_gemm_bf16<256, true>(E, bw.tm_b, M, H, H, bw.img + bw.mlp2[l], x, H, 0, ep); }
          else launch_gemm_bf16<256>(E, bw.tm_b, M, H, H, bw.img + bw.mlp2[l], x, H, 0, ep); }
        if (h->debug && h->tap_layers) E.tap(h->tap_layers + (size_t)l * h->cap_tap * H, x, M);
    }
    E.cat = SRHEP_CAT_HEAD;
    if (!rest) { E.head_done = true; return; }
    const int hw = d.v_in + d.ctx;
    const bool split_head = h->split && hw == kHeadK1 && d.head_h1 == kHeadH1 && !h->sw.head_fp32;
    if (split_head) {      // first head GEMM on the tensor cores (hi / lo planes), LeakyReLU(h1 + b1) as fp32 rows; the tail follows on the CUDA cores
        __half* hlo = (__half*)((uint8_t*)h->act_a + h->cap_ws_rows * (size_t)hw * 2);
        E.head_prep<__half>(E.head_params(p, x), (__half*)a, hw, hlo);
        if (!E.rc) {
            HeadChainParams q{};
            q.M = M; q.fp16 = 1; q.final_ln = d.head_final_ln; q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
            q.w1 = bw.img + bw.head1; q.w1_lo = bw.img_lo + bw.head1; q.ws1 = bw.ws_head1; q.h1out = h->h1buf;
            memcpy(q.b1, wh_b1(h), sizeof q.b1);
            q.stage = st;
            const int m_tiles = (M + 127) / 128;
            head_chain_kernel<true><<<std::max(1, std::min(m_tiles, 148)), kHeadThreads, kHeadSmemBytesSplit, E.s>>>(bw.tm_hin, bw.tm_hin_lo, q);
            E.check("head_chain_split");
        }
        E.head_done = false;
    } else
    if (h->sw.head_fp32 || h->split) {
        E.head_prep<float>(E.head_params(p, x), (float*)h->act_a, hw);
        GemmEpilogue ep; ep.bias = E.W(L.h1.b); ep.act = 1;
        E.gemm_f32<float>((float*)h->act_a, hw, E.W(L.h1.w), hw, h->h1buf, d.head_h1, M, d.head_h1, hw, ep);
    } else {
    const bool fused_head = bw.head_chain && !h->sw.no_headchain;
    // the whole head in ONE kernel (kernels_head_fused.cuh): the preparation warps write the operand rows straight into shared memory
    const bool one_kernel = fused_head && !h->sw.no_head_fused && !h->sw.head_prep_scalar && !h->sw.head_prep_v4 && !(h->debug && h->tap_final) && E.x_blocked &&
                            d.h_dim == 256 && d.cond == 96 && d.ctx == 160 && (d.cond + d.noisy_out) % 4 == 0;
    if (one_kernel) {
        if (!E.rc) {
            HeadFusedParams f{};
            f.q = E.head_params(p, x);
            HeadChainParams& q = f.c;
            q.M = M; q.fp16 = 1; q.final_ln = d.head_final_ln; q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
            q.w1 = bw.img + bw.head1; q.w2 = bw.img + bw.head2; q.w3 = bw.img + bw.head3;
            memcpy(q.b1, bw.head_b1, sizeof q.b1); memcpy(q.b2, bw.head_b2, sizeof q.b2); memcpy(q.b3, bw.head_b3, sizeof q.b3);
            memcpy(q.w4, bw.head_w4, sizeof q.w4); q.b4 = bw.head_b4;
            q.stage = st;
            { const char* v = getenv("SRHEP_HF_DIAG"); f.diag = v ? atoi(v) : 0; }
            const int m_tiles = (M + 127) / 128;
            head_fused_kernel<<<std::max(1, std::min(m_tiles, 148)), kHFThreads, kHFSmemBytes, E.s>>>(f);
            E.check("head_fused");
            E.head_done = true;
        }
    } else {
    if (fp16 || fused_head) E.head_prep<__half>(E.head_params(p, x), (__half*)a, hw);
    else E.head_prep<__nv_bfloat16>(E.head_params(p, x), a, hw);
    E.head_done = false;
    if (fused_head) {
        if (!E.rc) {
            HeadChainParams q;
            q.M = M; q.fp16 = 1; q.final_ln = d.head_final_ln; q.ext.rows_cap = (int)h->cap_ws_rows; q.ext.n_events = h->B;
            q.w1 = bw.img + bw.head1; q.w2 = bw.img + bw.head2; q.w3 = bw.img + bw.head3;
            memcpy(q.b1, bw.head_b1, sizeof q.b1); memcpy(q.b2, bw.head_b2, sizeof q.b2); memcpy(q.b3, bw.head_b3, sizeof q.b3);
            memcpy(q.w4, bw.head_w4, sizeof q.w4); q.b4 = bw.head_b4;
            q.stage = st;
            const int m_tiles = (M + 127) / 128;
            head_chain_kernel<false><<<std::max(1, std::min(m_tiles, 148)), kHeadThreads, kHeadSmemBytes, E.s>>>(bw.tm_hin, bw.tm_hin, q);
            E.check("head_chain");
            E.head_done = true;
        }
    } else
    { GemmEpilogue ep; ep.bias = bw.bias + bw.bias_head1; ep.act = 1;
      launch_gemm_bf16<128>(E, bw.tm_hin, M, hw, d.head_h1, bw.img + bw.head1, h->h1buf, d.head_h1, 0, ep); }
    }
    }
}

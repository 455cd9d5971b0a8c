// Host side of the device-resident dopri5 (kernels_ode.cuh): builds, once per binding, ONE graph = initial step-size
// selection + a conditional WHILE node whose body is one attempted step, and launches it per srhep_sample_dopri5 call.
// Included by srhep.cu inside its anonymous namespace.

void drop_dopri_graph(SrhepHandle* h) {
    if (h->dp_exec) { cudaGraphExecDestroy(h->dp_exec); h->dp_exec = nullptr; }
}

int build_dopri_graph(SrhepHandle* h) {
    const size_t np = h->passes.size();
    const int gridT = (int)std::min<size_t>(((size_t)h->T + 255) / 256, 148 * 8);
    // static part of the stage descriptors (the controller kernel only rewrites .t)
    std::vector<StageParams> sp(np * kDopriStageSlots);
    for (size_t pi = 0; pi < np; ++pi) {
        const int r0 = h->passes[pi].r0;
        for (int st = 0; st < 6; ++st)
            sp[pi * kDopriStageSlots + st] = StageParams{0.f, 0.f, (st == 5 ? h->y_b : h->y_tmp) + r0, nullptr, nullptr, h->kbuf[st + 1] + r0};
        sp[pi * kDopriStageSlots + 6] = StageParams{0.f, 0.f, h->y_a + r0, nullptr, nullptr, h->kbuf[0] + r0};
        sp[pi * kDopriStageSlots + 7] = StageParams{0.f, 0.f, h->y_tmp + r0, nullptr, nullptr, h->kbuf[1] + r0};
    }
    if (np > h->dp_cap_pass) {
        if (h->dp_sp) CK(h, cudaFree(h->dp_sp)); h->dp_sp = nullptr;
        if (h->dp_idx) CK(h, cudaFree(h->dp_idx)); h->dp_idx = nullptr;
        CK(h, cudaMalloc(&h->dp_sp, np * kDopriStageSlots * sizeof(StageParams)));
        CK(h, cudaMalloc(&h->dp_idx, np * sizeof(int)));
        h->dp_cap_pass = np;
    }
    CK(h, cudaMemcpy(h->dp_sp, sp.data(), sp.size() * sizeof(StageParams), cudaMemcpyHostToDevice));

    DopriBufs b; b.ycur = h->y_a; b.ynew = h->y_b; b.ytmp = h->y_tmp;
    for (int i = 0; i < 7; ++i) b.k[i] = h->kbuf[i];
    DopriStages ds; ds.sp = h->dp_sp; ds.idx = h->dp_idx; ds.n_pass = (int)np;
    Dopri5Ctl* ctl = h->dp_ctl;

    cudaStream_t cs = h->cap_stream, bs = h->dp_body_stream;
    CK(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t graph = nullptr;
    cudaStreamCaptureStatus cstat;
    const cudaGraphNode_t* deps = nullptr;
    size_t ndeps = 0;
    cudaGraphConditionalHandle handle;
    cudaError_t e = cudaStreamGetCaptureInfo(cs, &cstat, nullptr, &graph, &deps, &ndeps);
    if (e == cudaSuccess) e = cudaGraphConditionalHandleCreate(&handle, graph, 0, cudaGraphCondAssignDefault);
    auto abort_capture = [&](const char* what, cudaError_t err) {
        cudaGraph_t g = nullptr; cudaStreamEndCapture(cs, &g); if (g) cudaGraphDestroy(g);
        return fail(h, SRHEP_E_CUDA, "dopri5 graph: %s: %s", what, cudaGetErrorString(err));
    };
    if (e != cudaSuccess) return abort_capture("conditional handle", e);

    const uint64_t before = h->launches;
    auto evals = [&](Engine& E, bool bump) {
        for (size_t pi = 0; pi < np; ++pi) {
            const Pass& p = h->passes[pi];
            if (p.e1 == p.e0 || p.r1 == p.r0) continue;
            StageRef r; r.sp = h->dp_sp + pi * kDopriStageSlots; r.idx = h->dp_idx + pi;
            E.enqueue_eval(p, r);
            if (bump) E.bump(h->dp_idx + pi);
        }
    };
    Engine E{h, cs};
    dopri_controller_kernel<<<1, 32, 0, cs>>>(ctl, -1, ds, h->dp_tg, handle); E.check("dopri_begin");
    evals(E, false);                                                                   // f0 = f(t0, y0)
    dopri_init_norm_kernel<<<gridT, 256, 0, cs>>>(ctl, 0, b); E.check("dopri_init_norm");
    dopri_controller_kernel<<<1, 32, 0, cs>>>(ctl, 0, ds, h->dp_tg, handle); E.check("dopri_controller");
    dopri_trial_kernel<<<gridT, 256, 0, cs>>>(ctl, b); E.check("dopri_trial");
    evals(E, false);                                                                   // f(t0 + h0, y0 + h0 f0)
    dopri_init_norm_kernel<<<gridT, 256, 0, cs>>>(ctl, 1, b); E.check("dopri_init_norm");
    dopri_controller_kernel<<<1, 32, 0, cs>>>(ctl, 1, ds, h->dp_tg, handle); E.check("dopri_controller");
    if (E.rc) { abort_capture("init launches", cudaSuccess); return E.rc; }
    h->dp_init_nodes = (int)(h->launches - before);

    if ((e = cudaStreamGetCaptureInfo(cs, &cstat, nullptr, &graph, &deps, &ndeps)) != cudaSuccess) return abort_capture("capture info", e);
    cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
    cp.conditional.handle = handle;
    cp.conditional.type = cudaGraphCondTypeWhile;
    cp.conditional.size = 1;
    cudaGraphNode_t cnode;
    if ((e = cudaGraphAddNode(&cnode, graph, deps, ndeps, &cp)) != cudaSuccess) return abort_capture("conditional node", e);
    cudaGraph_t body = cp.conditional.phGraph_out[0];
    if ((e = cudaStreamUpdateCaptureDependencies(cs, &cnode, 1, cudaStreamSetCaptureDependencies)) != cudaSuccess) return abort_capture("capture dependencies", e);

    if ((e = cudaStreamBeginCaptureToGraph(bs, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return abort_capture("body capture", e);
    const uint64_t before_body = h->launches;
    Engine EB{h, bs};
    for (int st = 0; st < 6; ++st) {
        dopri_stage_kernel<<<gridT, 256, 0, bs>>>(ctl, st, b); EB.check("dopri_stage");
        evals(EB, true);
    }
    dopri_error_kernel<<<gridT, 256, 0, bs>>>(ctl, b); EB.check("dopri_error");
    dopri_controller_kernel<<<1, 32, 0, bs>>>(ctl, 2, ds, h->dp_tg, handle); EB.check("dopri_controller");
    dopri_commit_kernel<<<gridT, 256, 0, bs>>>(ctl, b, h->dp_tg); EB.check("dopri_commit");
    h->dp_body_nodes = (int)(h->launches - before_body);
    h->launches = before;
    e = cudaStreamEndCapture(bs, nullptr);
    if (e != cudaSuccess || EB.rc) { abort_capture("body end", e); return EB.rc ? EB.rc : SRHEP_E_CUDA; }
    if ((e = cudaStreamEndCapture(cs, &graph)) != cudaSuccess) return fail(h, SRHEP_E_CUDA, "dopri5 graph: end capture: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&h->dp_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { h->dp_exec = nullptr; return fail(h, SRHEP_E_CUDA, "dopri5 graph: instantiate: %s", cudaGetErrorString(e)); }
    return 0;
}

// One graph launch, one read-back of the control block at the end (statistics + status).
int dopri5_device(SrhepHandle* h, const float* x0, const float* tg, int n_steps, float atol, float rtol, int ret_seq, float* x_seq,
                  int32_t* stats_out, cudaStream_t s) {
    const size_t T = (size_t)h->T;
    int rc;
    const float* ya = h->y_a;
    if ((rc = ensure_state(h, true))) return rc;
    if (h->y_a != ya) drop_dopri_graph(h);
    if (!h->dp_ctl) {
        CK(h, cudaMalloc(&h->dp_ctl, sizeof(Dopri5Ctl)));
        CK(h, cudaMallocHost(&h->dp_ctl_host, sizeof(Dopri5Ctl)));
        CK(h, cudaStreamCreateWithFlags(&h->dp_body_stream, cudaStreamNonBlocking));
    }
    if ((size_t)n_steps > h->dp_cap_tg) {
        drop_dopri_graph(h);
        if (h->dp_tg) CK(h, cudaFree(h->dp_tg)); h->dp_tg = nullptr;
        const size_t cap = std::max<size_t>(1024, (size_t)n_steps);
        CK(h, cudaMalloc(&h->dp_tg, cap * sizeof(float)));
        h->dp_cap_tg = cap;
    }
    if (!h->dp_exec && (rc = build_dopri_graph(h))) return rc;
    Dopri5Ctl* c = h->dp_ctl_host;
    memset(c, 0, sizeof *c);
    c->t = tg[0]; c->atol = atol; c->rtol = rtol; c->n_steps = n_steps; c->ret_seq = ret_seq; c->T = (long long)T; c->x_seq = x_seq;
    c->max_attempts = 1000000;
    CK(h, cudaMemcpyAsync(h->dp_ctl, c, sizeof *c, cudaMemcpyHostToDevice, s));
    CK(h, cudaMemcpyAsync(h->dp_tg, tg, (size_t)n_steps * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(h, cudaMemcpyAsync(h->y_a, x0, T * sizeof(float), cudaMemcpyDeviceToDevice, s));
    CK(h, cudaGraphLaunch(h->dp_exec, s));
    CK(h, cudaMemcpyAsync(c, h->dp_ctl, sizeof *c, cudaMemcpyDeviceToHost, s));
    CK(h, cudaStreamSynchronize(s));            // the only synchronisation of the call: statistics and status of the finished integration
    h->launches += (uint64_t)h->dp_init_nodes + (uint64_t)c->attempts * h->dp_body_nodes;
    if (stats_out) { stats_out[0] = c->nfe; stats_out[1] = c->accepted; stats_out[2] = c->rejected; }
    if (c->status == 2) return fail(h, SRHEP_E_STATE, "dopri5: non-finite error norm (non-finite input or an event with zero cells?)");
    if (c->status == 3) return fail(h, SRHEP_E_STATE, "dopri5: max_num_steps exceeded");
    if (c->status != 1) return fail(h, SRHEP_E_STATE, "dopri5: loop ended in state %d", c->status);
    return SRHEP_OK;
}

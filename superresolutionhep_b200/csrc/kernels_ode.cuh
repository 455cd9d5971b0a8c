// Device-resident adaptive dopri5 (torchdiffeq's RKAdaptiveStepsizeODESolver with the Dormand-Prince tableau; call site
// models/flow_model.py:315-324; restated in oracle/odeint.py).
//
// The whole integration is ONE graph launch: t, dt, the accept flag, the output-grid cursor and the statistics live in a
// control block in device memory; a single-thread controller kernel makes every decision the host-driven loop makes
// (initial step size, error ratio, accept / reject, next dt, which output grid points the accepted step covers) and sets
// the handle of a conditional WHILE node whose body is one attempted step: 6 x (stage combination + network evaluation),
// the fused error estimate + scaled rms reduction, the controller, and the commit kernel (dense output for every grid
// point inside the step, y <- y1, f0 <- f1).  Nothing is read back until the loop has ended.
#pragma once
#include "common.cuh"

namespace srhep {

constexpr int kDopriStageSlots = 8;      // per pass: stage descriptors 0-5 = the attempt, 6 = f(t0, y0), 7 = the initial-step trial

struct Dopri5Ctl {
    // set by the host before the launch
    double t;                 // left end of the next attempt
    float  atol, rtol;
    int    n_steps, ret_seq;
    long long T;              // cells
    float* x_seq;             // (n_steps, T) if ret_seq else (T)
    int    max_attempts;
    // maintained on the device
    double dt;
    double sum_a, sum_b;      // reduction targets (scaled sums of squares)
    float  dty;               // (float)dt of the attempt in flight
    float  h0;                // trial step of the initial step-size selection
    double d1;
    // decision of the last controller run, read by the commit kernel
    int    accept, out_a, out_b;
    double t_lo, t_hi;        // the accepted step
    float  dty_commit;
    // bookkeeping
    int    next_out, nfe, accepted, rejected, attempts, status;     // status: 0 running, 1 done, 2 non-finite error norm, 3 max_attempts
};

__constant__ double kDopriAlpha[6] = {1 / 5., 3 / 10., 4 / 5., 8 / 9., 1., 1.};
__constant__ float kDopriBeta[6][6] = {
    {(float)(1 / 5.), 0, 0, 0, 0, 0},
    {(float)(3 / 40.), (float)(9 / 40.), 0, 0, 0, 0},
    {(float)(44 / 45.), (float)(-56 / 15.), (float)(32 / 9.), 0, 0, 0},
    {(float)(19372 / 6561.), (float)(-25360 / 2187.), (float)(64448 / 6561.), (float)(-212 / 729.), 0, 0},
    {(float)(9017 / 3168.), (float)(-355 / 33.), (float)(46732 / 5247.), (float)(49 / 176.), (float)(-5103 / 18656.), 0},
    {(float)(35 / 384.), 0, (float)(500 / 1113.), (float)(125 / 192.), (float)(-2187 / 6784.), (float)(11 / 84.)}};
__constant__ float kDopriErr[7] = {(float)(35 / 384. - 1951 / 21600.), 0, (float)(500 / 1113. - 22642 / 50085.), (float)(125 / 192. - 451 / 720.),
                                   (float)(-2187 / 6784. + 12231 / 42400.), (float)(11 / 84. - 649 / 6300.), (float)(-1. / 60.)};
__constant__ float kDopriMid[7] = {(float)(6025192743. / 30085553152. / 2), 0, (float)(51252292925. / 65400821598. / 2), (float)(-2691868925. / 45128329728. / 2),
                                   (float)(187940372067. / 1594534317056. / 2), (float)(-1776094331. / 19743644256. / 2), (float)(11237099. / 235043384. / 2)};

struct DopriBufs { float* ycur; float* ynew; float* ytmp; float* k[7]; };

__device__ __forceinline__ double block_sum_double(double acc, double* sred) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sred[warp] = acc;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    __syncthreads();
    return t;
}

// stage input of attempt stage `st`: out = ycur + sum_{i <= st} (beta[st][i] * dty) * k[i]      (same float operations as combine_kernel)
__global__ void __launch_bounds__(256) dopri_stage_kernel(const Dopri5Ctl* __restrict__ ctl, int st, DopriBufs b) {
    const float dty = ctl->dty;
    const size_t n = (size_t)ctl->T;
    float c[6]; int nk = 0; const float* kp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) if (i <= st && kDopriBeta[st][i] != 0.f) { c[nk] = __fmul_rn(kDopriBeta[st][i], dty); kp[nk] = b.k[i]; ++nk; }
    float* out = st == 5 ? b.ynew : b.ytmp;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) if (j < nk) acc = fmaf(c[j], kp[j][i], acc);
        out[i] = b.ycur[i] + acc;
    }
}

// ytmp = ycur + h0 * k0 (initial step-size selection trial)
__global__ void __launch_bounds__(256) dopri_trial_kernel(const Dopri5Ctl* __restrict__ ctl, DopriBufs b) {
    const float h0 = ctl->h0;
    const size_t n = (size_t)ctl->T;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        b.ytmp[i] = b.ycur[i] + fmaf(h0, b.k[0][i], 0.f);
}

// mode 0: sum_a += (ycur / sc)^2, sum_b += (k0 / sc)^2   mode 1: sum_a += ((k1 - k0) / sc)^2      sc = atol + rtol |ycur|
__global__ void __launch_bounds__(256) dopri_init_norm_kernel(Dopri5Ctl* ctl, int mode, DopriBufs b) {
    __shared__ double sred[8];
    const float atol = ctl->atol, rtol = ctl->rtol;
    const size_t n = (size_t)ctl->T;
    double a = 0.0, c = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float y = b.ycur[i];
        const float sc = atol + rtol * fabsf(y);
        if (mode == 0) {
            const float q0 = y / sc, q1 = b.k[0][i] / sc;
            a += (double)q0 * (double)q0; c += (double)q1 * (double)q1;
        } else {
            const float q = (b.k[1][i] - b.k[0][i]) / sc;
            a += (double)q * (double)q;
        }
    }
    a = block_sum_double(a, sred);
    if (mode == 0) c = block_sum_double(c, sred);
    if (threadIdx.x == 0) { atomicAdd(&ctl->sum_a, a); if (mode == 0) atomicAdd(&ctl->sum_b, c); }
}

// error estimate of the attempt and its scaled sum of squares:  err = sum_i (c_err[i] * dty) k[i];  q = err / (atol + rtol max(|y0|, |y1|))
__global__ void __launch_bounds__(256) dopri_error_kernel(Dopri5Ctl* ctl, DopriBufs b) {
    __shared__ double sred[8];
    const float atol = ctl->atol, rtol = ctl->rtol, dty = ctl->dty;
    const size_t n = (size_t)ctl->T;
    float c[7]; const float* kp[7]; int nk = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) if (kDopriErr[i] != 0.f) { c[nk] = __fmul_rn(kDopriErr[i], dty); kp[nk] = b.k[i]; ++nk; }
    double acc = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float e = 0.f;
#pragma unroll
        for (int j = 0; j < 7; ++j) if (j < nk) e = fmaf(c[j], kp[j][i], e);
        const float sc = fmaxf(fabsf(b.ycur[i]), fabsf(b.ynew[i]));
        const float q = e / (atol + rtol * sc);
        acc += (double)q * (double)q;
    }
    acc = block_sum_double(acc, sred);
    if (threadIdx.x == 0) atomicAdd(&ctl->sum_a, acc);
}

// per-pass stage descriptors: [pass][kDopriStageSlots]; the controller rewrites the evaluation times and rewinds the stage cursors
struct DopriStages { StageParams* sp; int* idx; int n_pass; };

__device__ __forceinline__ void dopri_write_attempt(Dopri5Ctl* c, const DopriStages& s) {
    const float dty = (float)c->dt, t0y = (float)c->t, t1y = (float)(c->t + c->dt);
    c->dty = dty;
    for (int p = 0; p < s.n_pass; ++p) {
        for (int st = 0; st < 6; ++st)
            s.sp[p * kDopriStageSlots + st].t = kDopriAlpha[st] == 1.0 ? t1y : __fadd_rn(t0y, __fmul_rn((float)kDopriAlpha[st], dty));
        s.idx[p] = 0;
    }
}

// phase -1: first node of the graph
// phase 0: after f0 and the first two norms -> h0 and the trial evaluation's time
// phase 1: after the trial evaluation and its norm -> first dt, first attempt's stage times, loop condition
// phase 2: after an attempt's error norm -> accept / reject, next dt, outputs covered, next attempt, loop condition
__global__ void dopri_controller_kernel(Dopri5Ctl* c, int phase, DopriStages s, const float* __restrict__ tg, cudaGraphConditionalHandle handle) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double T = (double)c->T;
    bool go = true;
    if (phase < 0) {                    // start of the graph: the first evaluation is f(t0, y0) (stage slot 6)
        for (int p = 0; p < s.n_pass; ++p) { s.sp[p * kDopriStageSlots + 6].t = (float)c->t; s.idx[p] = 6; }
        return;
    }
    if (phase == 0) {
        const double d0 = sqrt(c->sum_a / T), d1 = sqrt(c->sum_b / T);
        float h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6f : (float)(0.01 * d0 / d1);
        h0 = fabsf(h0);
        c->h0 = h0; c->d1 = d1; c->sum_a = 0.0; c->sum_b = 0.0;
        c->nfe = 1;
        for (int p = 0; p < s.n_pass; ++p) { s.sp[p * kDopriStageSlots + 7].t = (float)(c->t + (double)h0); s.idx[p] = 7; }
        return;
    }
    if (phase == 1) {
        const double h0 = (double)c->h0, d1 = c->d1;
        const double d2 = fabs(sqrt(c->sum_a / T) / h0);
        double h1;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        c->dt = fmin(100.0 * h0, fabs(h1));
        c->sum_a = 0.0;
        c->nfe = 2;
        c->next_out = 1; c->accepted = 0; c->rejected = 0; c->attempts = 0; c->status = 0; c->accept = 0;
        dopri_write_attempt(c, s);
        go = c->n_steps > 1;
    } else {
        const double ratio = sqrt(c->sum_a / T);
        c->sum_a = 0.0;
        c->nfe += 6; c->attempts += 1;
        c->accept = 0; c->out_a = c->out_b = c->next_out;
        if (!(ratio == ratio) || isinf(ratio)) { c->status = 2; go = false; }
        else {
            if (ratio <= 1.0) {
                c->accept = 1; c->accepted += 1;
                c->t_lo = c->t; c->t_hi = c->t + c->dt; c->dty_commit = c->dty;
                int j = c->next_out;
                while (j < c->n_steps && (double)tg[j] <= c->t_hi) ++j;
                c->out_b = j; c->next_out = j;
                c->t = c->t_hi;
            } else c->rejected += 1;
            if (ratio == 0.0) c->dt *= 10.0;
            else {
                const double dfac = ratio < 1.0 ? 1.0 : 0.2;
                c->dt *= fmin(10.0, fmax(0.9 / pow(ratio, 0.2), dfac));
            }
            if (c->next_out >= c->n_steps) { c->status = 1; go = false; }
            else if (c->attempts >= c->max_attempts) { c->status = 3; go = false; }
        }
        if (go) dopri_write_attempt(c, s);
    }
    cudaGraphSetConditional(handle, go ? 1u : 0u);
}

// accepted step: 4th-order dense output at every grid point inside (t_lo, t_hi] (quartic through y0, y1, y_mid, f0, f1), then y <- y1, f0 <- f1
__global__ void __launch_bounds__(256) dopri_commit_kernel(const Dopri5Ctl* __restrict__ ctl, DopriBufs b, const float* __restrict__ tg) {
    if (!ctl->accept) return;
    __shared__ float coef[6];
    const size_t n = (size_t)ctl->T;
    const int oa = ctl->out_a, ob = ctl->out_b, last = ctl->n_steps - 1;
    const float dty = ctl->dty_commit;
    float cm[7]; const float* kp[7]; int nk = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) if (kDopriMid[i] != 0.f) { cm[nk] = __fmul_rn(kDopriMid[i], dty); kp[nk] = b.k[i]; ++nk; }
    for (int j = oa; j < ob; ++j) {
        const bool wanted = ctl->ret_seq || j == last;
        if (threadIdx.x == 0) {
            const float x = (float)(((double)tg[j] - ctl->t_lo) / (ctl->t_hi - ctl->t_lo));
            const float x2 = __fmul_rn(x, x), x3 = __fmul_rn(x2, x), x4 = __fmul_rn(x3, x);
            // y0 + x dt f0 + x^2 c + x^3 b + x^4 a,  a = 2dt(f1-f0) - 8(y1+y0) + 16 ym, b = dt(5f0-3f1) + 18y0 + 14y1 - 32ym, c = dt(f1-4f0) - 11y0 - 5y1 + 16ym
            coef[0] = 1.f - 11.f * x2 + 18.f * x3 - 8.f * x4;
            coef[1] = -5.f * x2 + 14.f * x3 - 8.f * x4;
            coef[2] = 16.f * x2 - 32.f * x3 + 16.f * x4;
            coef[3] = dty * (x - 4.f * x2 + 5.f * x3 - 2.f * x4);
            coef[4] = dty * (x2 - 3.f * x3 + 2.f * x4);
            coef[5] = x;
        }
        __syncthreads();
        if (wanted) {
            float* o = ctl->x_seq + (ctl->ret_seq ? (size_t)j * n : 0);
            const bool at_end = coef[5] == 1.0f;
            for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
                const float y1 = b.ynew[i];
                if (at_end) { o[i] = y1; continue; }
                const float y0 = b.ycur[i];
                float m = 0.f;
#pragma unroll
                for (int q = 0; q < 7; ++q) if (q < nk) m = fmaf(cm[q], kp[q][i], m);
                const float ym = y0 + m;
                float acc = fmaf(coef[0], y0, 0.f);
                acc = fmaf(coef[1], y1, acc); acc = fmaf(coef[2], ym, acc); acc = fmaf(coef[3], b.k[0][i], acc); acc = fmaf(coef[4], b.k[6][i], acc);
                o[i] = acc;
            }
        }
        __syncthreads();
    }
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        b.ycur[i] = b.ynew[i];
        b.k[0][i] = b.k[6][i];
    }
}

}  // namespace srhep

// Host side + kernels of include/srhep_post.h (included at the end of srhep.cu).
//   inference.py:146-152,163-287 + utility/target_transformation.py:17-33  -> srpost_ensemble_unscale
//   inference.py:291-310 -> pflow/dataset_pf.py:81-92,136-147              -> srpost_select_cells
#include "../../include/srhep_post.h"

namespace srhep {

__device__ __forceinline__ float target_inverse(const SrpostTargetTransform& t, float y, float proxy_raw) {
    if (t.standard) y = fmaf(y, t.std, t.mean);
    float ratio = 1.0f / (1.0f + expf(-y));
    ratio = (ratio - t.alpha) / (1.0f - 2.0f * t.alpha);
    return ratio * proxy_raw * t.f;
}

__global__ void __launch_bounds__(256) ensemble_unscale_kernel(const float* __restrict__ x, int n_ens, int n_store, size_t T, const float* __restrict__ proxy,
                                                              SrpostTargetTransform tt, float unit, float* nn_avg, float* e_avg, float* e_raw) {
    const size_t n = (size_t)n_store * T;
    const float inv = 1.0f / (float)n_ens;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float pr = proxy[i % T];
        float s = 0.f, se = 0.f;
        for (int e = 0; e < n_ens; ++e) {
            const float v = x[(size_t)e * n + i];
            s += v;
            se += target_inverse(tt, v, pr) * unit;                    // unscale each member, then average (inference.py:241-276)
        }
        if (nn_avg) nn_avg[i] = s * inv;
        if (e_avg) e_avg[i] = target_inverse(tt, s * inv, pr) * unit;  // average the network outputs, then unscale (inference.py:199-201)
        if (e_raw) e_raw[i] = se * inv;
    }
}

// cells above threshold per event
__global__ void __launch_bounds__(256) select_count_kernel(const float* __restrict__ e, const int* __restrict__ cu, float thr, int* counts) {
    __shared__ int red[8];
    const int ev = blockIdx.x;
    int c = 0;
    for (int r = cu[ev] + threadIdx.x; r < cu[ev + 1]; r += 256) c += e[r] > thr ? 1 : 0;
    c = (int)warp_sum((float)c);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) { int s = 0; for (int w = 0; w < 8; ++w) s += red[w]; counts[ev] = s; }
}

// exclusive scan of counts[0..B) into cu_out[0..B]; one block
__global__ void __launch_bounds__(1024) select_scan_kernel(const int* counts, int B, int* cu_out) {
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    const int per = (B + 1023) / 1024;
    const int b0 = min(tid * per, B), b1 = min(b0 + per, B);
    int s = 0;
    for (int i = b0; i < b1; ++i) s += counts[i];
    part[tid] = s;
    __syncthreads();
    if (tid == 0) { int acc = 0; for (int i = 0; i < 1024; ++i) { const int v = part[i]; part[i] = acc; acc += v; } cu_out[B] = acc; }
    __syncthreads();
    int acc = part[tid];
    for (int i = b0; i < b1; ++i) { cu_out[i] = acc; acc += counts[i]; }
}

__global__ void __launch_bounds__(256) select_scatter_kernel(const float* __restrict__ e, const float* __restrict__ eta_raw, const float* __restrict__ phi,
                                                            const int* __restrict__ layer, const int* __restrict__ cu, const int* __restrict__ cu_out, float thr,
                                                            PflowVarTransform tr_e, PflowVarTransform tr_eta, SrpostPflowOut o) {
    __shared__ int woff[8];
    __shared__ int base;
    const int ev = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = cu_out[ev];
    __syncthreads();
    for (int r0 = cu[ev]; r0 < cu[ev + 1]; r0 += 256) {
        const int r = r0 + threadIdx.x;
        const bool keep = r < cu[ev + 1] && e[r] > thr;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) woff[warp] = __popc(m);
        __syncthreads();
        int off = base;
        for (int w = 0; w < warp; ++w) off += woff[w];
        if (keep) {
            const int d = off + __popc(m & ((1u << lane) - 1u));
            const float er = e[r], et = eta_raw[r], ph = phi[r];
            o.e_raw[d] = er; o.eta_raw[d] = et; o.phi[d] = ph; o.layer[d] = layer[r];
            o.e[d] = pf_var_forward(tr_e, er); o.eta[d] = pf_var_forward(tr_eta, et);
            o.cosphi[d] = cosf(ph); o.sinphi[d] = sinf(ph);
        }
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += woff[w]; base += t; }
        __syncthreads();
    }
}

}  // namespace srhep

namespace {
thread_local std::string g_post_error;
int post_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_post_error = buf;
    return code;
}
}  // namespace

extern "C" {

const char* srpost_last_error(void) { return g_post_error.c_str(); }

int srpost_ensemble_unscale(const float* x, int32_t n_ens, int32_t n_store, int64_t T, const float* proxy, const SrpostTargetTransform* tt, float unit,
                            float* nn_avg, float* e_avg, float* e_raw, void* stream) {
    if (n_ens < 1 || n_store < 0 || T < 0 || !tt) return post_fail(SRHEP_E_INVALID, "n_ens >= 1, n_store >= 0, n_cells >= 0 and a transform are required");
    if ((size_t)n_store * (size_t)T == 0) return SRHEP_OK;
    if (!x || !proxy) return post_fail(SRHEP_E_INVALID, "null argument");
    if (!(tt->alpha >= 0.f && tt->alpha < 0.5f)) return post_fail(SRHEP_E_INVALID, "alpha must be in [0, 0.5)");
    const size_t n = (size_t)n_store * (size_t)T;
    const int grid = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    ensemble_unscale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n_ens, n_store, (size_t)T, proxy, *tt, unit, nn_avg, e_avg, e_raw);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return post_fail(SRHEP_E_CUDA, "launch ensemble_unscale: %s", cudaGetErrorString(e));
    return SRHEP_OK;
}

int srpost_select_cells(const float* e_pred, const float* eta_raw, const float* phi, const int32_t* layer, const int32_t* cu_in, int32_t B, float thr,
                        const PflowVarTransform* tr_e, const PflowVarTransform* tr_eta, const SrpostPflowOut* o, int32_t* cu_out, void* stream) {
    if (B < 0 || !cu_in || !cu_out || !tr_e || !tr_eta || !o) return post_fail(SRHEP_E_INVALID, "null argument / negative event count");
    cudaStream_t s = (cudaStream_t)stream;
    if (B == 0) { cudaError_t e = cudaMemsetAsync(cu_out, 0, sizeof(int32_t), s); return e == cudaSuccess ? SRHEP_OK : post_fail(SRHEP_E_CUDA, "%s", cudaGetErrorString(e)); }
    if (!e_pred || !eta_raw || !phi || !layer || !o->e || !o->eta || !o->cosphi || !o->sinphi || !o->phi || !o->e_raw || !o->eta_raw || !o->layer)
        return post_fail(SRHEP_E_INVALID, "null cell array");
    int* counts = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&counts, (size_t)B * sizeof(int), s);
    if (e != cudaSuccess) return post_fail(e == cudaErrorMemoryAllocation ? SRHEP_E_NOMEM : SRHEP_E_CUDA, "cudaMallocAsync: %s", cudaGetErrorString(e));
    select_count_kernel<<<B, 256, 0, s>>>(e_pred, cu_in, thr, counts);
    select_scan_kernel<<<1, 1024, 0, s>>>(counts, B, cu_out);
    select_scatter_kernel<<<B, 256, 0, s>>>(e_pred, eta_raw, phi, layer, cu_in, cu_out, thr, *tr_e, *tr_eta, *o);
    e = cudaGetLastError();
    cudaFreeAsync(counts, s);
    if (e != cudaSuccess) return post_fail(SRHEP_E_CUDA, "launch select_cells: %s", cudaGetErrorString(e));
    return SRHEP_OK;
}

}  // extern "C"

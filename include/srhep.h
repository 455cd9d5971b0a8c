/*
 * srhep.h -- C ABI of the B200 (sm_100a) super-resolution sampling hot path.
 *
 * The reference (etiennedreyer/SuperResolutionHEP) is pure Python/PyTorch and has no FFI
 * of its own; this header is the boundary a maintainer binds with ctypes/cffi from the
 * reference's Python call sites.  Every entry point cites the
 * reference interface it replaces (paths relative to the reference repo root).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross the boundary;
 *   - all functions return 0 on success, a negative SRHEP_E_* code otherwise, and never
 *     throw; srhep_last_error() gives the message of the last failure on that handle
 *     (or of the last failed srhep_create / pflow_create when handle == NULL);
 *   - "dev" pointers are device memory of the handle's device, "host" pointers are host
 *     memory; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - events are PACKED: event i owns rows cu_seqlens[i] .. cu_seqlens[i+1]-1 of every
 *     per-cell array (the reference's padded (B, Nmax, 1) tensors with q_mask compacted,
 *     dataset.py:341-349); T = cu_seqlens[B];
 *   - a handle is not thread-safe; calls on one handle are serialised by the caller.
 */
#ifndef SRHEP_H_
#define SRHEP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRHEP_OK            0
#define SRHEP_E_INVALID    -1   /* bad argument / unsupported dimension */
#define SRHEP_E_CUDA       -2   /* a CUDA runtime / driver call failed  */
#define SRHEP_E_NOMEM      -3
#define SRHEP_E_STATE      -4   /* call order violated                  */

/* arithmetic mode of the dense contractions (inference.py:330 `-p/--precision`):
 * FP32 = 'highest': fp32-grade on the tensor cores (every operand of a contraction is a pair of fp16 planes x = hi + lo,
 * products run as hi.hi + hi.lo + lo.hi with fp32 accumulation; rel. error 2e-6 on the transformer output); the
 * environment variable SRHEP_FP32_SIMT=1, read by srhep_create, selects reference-order fp32 FFMA on the CUDA cores instead.
 * BF16 = bf16 tcgen05 MMA operands with fp32 accumulation.  In every mode LayerNorm statistics, softmax, residual
 * stream and ODE state stay fp32. */
#define SRHEP_PREC_FP32 0
#define SRHEP_PREC_BF16 1
#define SRHEP_PREC_FP16 2   /* as BF16 but fp16 tcgen05 operands: same speed, 8x finer mantissa; activations behind a
                             * LayerNorm and the weights are well inside fp16 range */

/* fixed-grid solvers of torchdiffeq.odeint (call site models/flow_model.py:315-324) */
#define SRHEP_EULER    0
#define SRHEP_MIDPOINT 1
#define SRHEP_RK4      2
#define SRHEP_DOPRI5   3

/* Dimensions derived from the YAML `flow_model` block exactly as FlowModel.__init__ does
 * (models/flow_model.py:29-110).  All int32, same order as SrDimsC in config.py. */
typedef struct SrhepDims {
    int32_t h_dim, heads, layers, t_emb, freq_dim;
    int32_t etaphi_in, etaphi_hid, etaphi_out;
    int32_t layer_emb_dim, layer_hid, layer_out;
    int32_t proxy_hid, proxy_out;
    int32_t noisy_hid, noisy_out;
    int32_t mlp_hid;
    int32_t head_h1, head_h2, head_h3, head_final_ln;
    int32_t cond, ctx, v_in;
} SrhepDims;

/* Packed per-cell conditioning inputs (the keys FlowModel.forward reads,
 * models/flow_model.py:187-189): fp32 (T) each, `layer` int32 (T) in {0,1,2}. */
typedef struct SrhepCond {
    const float*   eta;
    const float*   cosphi;
    const float*   sinphi;
    const float*   e_proxy;
    const int32_t* layer;
} SrhepCond;

typedef struct SrhepHandle SrhepHandle;

/* Number of fp32 values srhep_create expects in `weights_host` for `dims`: every tensor of
 * FlowModel.state_dict() flattened in SrDims.param_order() order, followed by
 * the freq_dim/2 sinusoid frequencies of TimestepEmbedder (models/utils.py:152-154). */
size_t srhep_weight_count(const SrhepDims* dims);

/* Replaces SupResLightning(...).load_state_dict(ckpt['state_dict']).eval().cuda()
 * (inference.py:74-83): uploads and re-packs the weights on `device`. */
int srhep_create(int device, const SrhepDims* dims, const float* weights_host, size_t n_floats,
                 int precision, SrhepHandle** out);
int srhep_destroy(SrhepHandle* h);
const char* srhep_last_error(const SrhepHandle* h);

/* Upper bound on the real cells processed per pass (events are independent, so a batch is
 * cut into passes that keep the activation workspace L2-resident).  0 = library default. */
int srhep_set_pass_tokens(SrhepHandle* h, int64_t max_tokens);
/* 1 = capture each pass's evaluation in a CUDA graph and replay it per step (default 1). */
int srhep_set_use_graph(SrhepHandle* h, int enable);

/* Binds a batch of packed events: replaces moving the collate_graphs dict to the device
 * (inference.py:141-143) + q_mask handling.  `cond` arrays are dev pointers that must stay
 * valid until the next bind; `cu_seqlens_host` has B+1 entries, non-decreasing, [0] == 0. */
int srhep_bind_events(SrhepHandle* h, const SrhepCond* cond_dev, const int32_t* cu_seqlens_host,
                      int32_t n_events, void* stream);

/* FlowModel.forward(batch, noisy_input, time_step) (models/flow_model.py:167-264) on the
 * bound events: x_dev (T) noisy input, t_dev (B) per-event time, v_dev (T) velocity. */
int srhep_velocity(SrhepHandle* h, const float* x_dev, const float* t_dev, float* v_dev, void* stream);

/* FlowModel.generate_samples(batch, n_steps, method, ret_seq) (models/flow_model.py:302-329)
 * for the fixed-grid methods.  x0_dev (T) is the initial noise (the reference draws
 * randn_like(e_proxy) at :319 -- the caller draws it so that RNG parity stays in PyTorch);
 * t_grid_host = torch.linspace(0, 1, n_steps) as fp32 (n_steps >= 2 values);
 * x_seq_dev: (n_steps, T) if ret_seq (row 0 = x0) else (T) final state.
 * nfe_out (host, may be NULL) receives the number of network evaluations. */
int srhep_sample(SrhepHandle* h, const float* x0_dev, const float* t_grid_host, int32_t n_steps,
                 int32_t method, int32_t ret_seq, float* x_seq_dev, int32_t* nfe_out, void* stream);

/* Adaptive dopri5 with the step controller of torchdiffeq (atol/rtol as at
 * models/flow_model.py:321-323; the reference's default method, :303); rms norms are taken over
 * real cells only.  Device-resident: t, dt, accept/reject and the output-grid cursor live in
 * device memory and the adaptive loop is a conditional WHILE node of one CUDA graph, so the
 * call launches once and synchronises once, at its end, to return the statistics (with
 * srhep_set_use_graph(h, 0) the loop is driven from the host, one scalar read back per attempted
 * step like torchdiffeq).  stats_out (host, 3 ints, may be NULL): nfe, accepted, rejected. */
int srhep_sample_dopri5(SrhepHandle* h, const float* x0_dev, const float* t_grid_host, int32_t n_steps,
                        float atol, float rtol, int32_t ret_seq, float* x_seq_dev,
                        int32_t* stats_out, void* stream);

/* Parity hook: copies a named intermediate of the LAST srhep_velocity call (single pass
 * only) into out_dev.  Names: "time_emb" (B,t_emb) "context" (B,ctx) "tok_feat" (T,cond+noisy_out)
 * "feat_0" "layer_<i>" "transformer_out" (T,h_dim).  Requires srhep_set_debug(h, 1). */
int srhep_set_debug(SrhepHandle* h, int enable);
int srhep_get_tap(SrhepHandle* h, const char* name, float* out_dev, size_t n_floats, void* stream);

/* Measurement hook (bench.py's roofline leg): one srhep_velocity-equivalent evaluation at a
 * single time t over all passes, launched directly (no graph) with a CUDA event after every
 * kernel on `stream`; returns the device time and launch count per kernel category.
 * Synchronises the stream.  ms_by_cat / launches_by_cat: host arrays of SRHEP_NCAT. */
#define SRHEP_CAT_EMBED 0   /* timestep embedding, cell embedding nets, masked mean    */
#define SRHEP_CAT_ADALN 1   /* per-event GEMMs: all adaLN Linears, feat_0 context part */
#define SRHEP_CAT_FEAT0 2   /* feat_0_mlp token GEMM                                    */
#define SRHEP_CAT_LN    3   /* LayerNorm + adaLN modulate                               */
#define SRHEP_CAT_QKV   4   /* q|k|v projection GEMM                                    */
#define SRHEP_CAT_ATTN  5   /* varlen attention                                         */
#define SRHEP_CAT_OUT   6   /* attention output projection + gate + residual            */
#define SRHEP_CAT_MLP1  7   /* layer MLP first Linear + LeakyReLU                       */
#define SRHEP_CAT_MLP2  8   /* layer MLP second Linear + LeakyReLU + gate + residual    */
#define SRHEP_CAT_HEAD  9   /* velocity head + ODE update                               */
#define SRHEP_CAT_CHAIN 10  /* fused layer chain: out-proj .. MLP .. next layer's LN1 + q|k|v  */
#define SRHEP_NCAT     11
int srhep_profile(SrhepHandle* h, const float* x_dev, float t, float* v_dev, float* ms_by_cat,
                  int32_t* launches_by_cat, void* stream);

/* Kernels launched by this handle since creation (bench.py's gpu_launches claim). */
uint64_t srhep_launch_count(const SrhepHandle* h);

/* Library self-description: "srhep <version> sm_100a" */
const char* srhep_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SRHEP_H_ */

/*
 * pflow.h -- C ABI of the B200 (sm_100a) particle-flow forward (BASELINE.json configs[4]):
 * the reference's SAPF network run on the cells of the SR output.
 *
 * Replaces PflowLightning(config_mv, config_t, inference=True).net(batch) = SAPF.forward
 * (inference_pf.py:76,135; pflow/models/model_pf.py:56-74; pflow/models/encoder.py:38-58;
 * pflow/models/cardinality_predictor.py:17-22; pflow/models/kinematics_predictor.py:24-57,99-135).
 * Same conventions as srhep.h: plain pointers and sizes, 0 / negative SRHEP_E_* status codes,
 * packed events (cell rows cu_seqlens[i] .. cu_seqlens[i+1]-1 belong to event i), every call
 * takes the cudaStream_t to run on, a handle is bound to one device and is not thread-safe.
 */
#ifndef PFLOW_H_
#define PFLOW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFLOW_MAX_CARD_HIDDEN 4

/* Architecture, read from the YAML `pf_model` block (pflow/configs/model_and_var.yml:9-70). */
typedef struct PflowDims {
    int32_t h_dim, heads;                       /* 64, 4 */
    int32_t enc_layers, kin_layers;             /* DiT layers of the cell encoder / the particle decoder */
    int32_t layer_emb_dim;                      /* encoder.layer_emb_dim */
    int32_t max_particles, part_emb_dim;        /* max_particles, kinematics_predictor.init_particles.embedding_dim */
    int32_t card_n_hidden;                      /* len(cardinality_predictor.hidden_layers) <= 4 */
    int32_t card_hidden[PFLOW_MAX_CARD_HIDDEN];
    int32_t card_out;                           /* max_particles + 1 */
} PflowDims;

/* utility/transformation.py:VarTransformation.forward = scale(trans(x)) */
#define PFLOW_TRANS_NONE 0
#define PFLOW_TRANS_POW 1            /* pow(x, m)                    */
#define PFLOW_TRANS_POW_SIGNED 2     /* sign(x) * |x|^m              */
#define PFLOW_SCALE_NONE 0
#define PFLOW_SCALE_MINMAX 1         /* (x-min)/(max-min)*(hi-lo)+lo */
#define PFLOW_SCALE_STANDARD 2       /* (x-mean)/std                 */
typedef struct PflowVarTransform {
    int32_t trans; float m;
    int32_t scale; float mean, std, min, max, lo, hi;
} PflowVarTransform;

/* Packed per-cell inputs: the keys SAPF.forward reads from the collate_fn dict
 * (pflow/dataset_pf.py:246-259) with cell_mask compacted.  fp32 (T) each, layer int32 (T). */
typedef struct PflowCells {
    const float* e; const float* eta; const float* cosphi; const float* sinphi;   /* scaled inputs of the encoder       */
    const float* phi; const float* e_raw; const float* eta_raw;                   /* raw inputs of AttnKinematicNet     */
    const int32_t* layer;
} PflowCells;

typedef struct PflowHandle PflowHandle;

/* Number of fp32 values pflow_create expects: every tensor of SAPF.state_dict() flattened in the
 * reference's own order (encoder.*, cardinality_predictor.*, kinematics_predictor.*). */
size_t pflow_weight_count(const PflowDims* dims);

/* Replaces PflowLightning(...).load_state_dict(ckpt['state_dict']).eval().cuda() and
 * kin_net.set_trans_dicts (pflow/lightning_pf.py:52-58): transforms[0..2] = pt, eta, e. */
int pflow_create(int device, const PflowDims* dims, const float* weights_host, size_t n_floats,
                 const PflowVarTransform* transforms, PflowHandle** out);
int pflow_destroy(PflowHandle* h);
const char* pflow_last_error(const PflowHandle* h);

/* SAPF.forward(batch) on packed events.
 *   part_mask_dev : NULL = inference mode (n_pred = argmax(logits), part_mask = arange(P) < n_pred,
 *                   model_pf.py:65-67); else uint8 (B, P), the `batch['part_mask']` of training mode.
 *   logits_dev    : (B, card_out) fp32          n_pred_dev : (B) int32 argmax (may be NULL)
 *   kin_pred_dev  : (B, P, 4) fp32 = (pt, eta, phi, e) transformed
 *   inc_dev       : (P, T) fp32 incidence weights, particle-major packed cells
 *                   (= inc_weights (B, P, Nmax) of the reference with padded cells dropped). */
int pflow_forward(PflowHandle* h, const PflowCells* cells_dev, const int32_t* cu_seqlens_host, int32_t n_events,
                  const uint8_t* part_mask_dev, float* logits_dev, int32_t* n_pred_dev, float* kin_pred_dev,
                  float* inc_dev, void* stream);

/* Kernels launched by this handle since creation. */
uint64_t pflow_launch_count(const PflowHandle* h);

#ifdef __cplusplus
}
#endif
#endif /* PFLOW_H_ */

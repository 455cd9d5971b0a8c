/*
 * srhep_post.h -- C ABI of the two "next" rows of the hot path: what the
 * reference does in per-event Python loops between the sampler and the particle-flow model.
 *
 *   srpost_ensemble_unscale : ensemble mean + TargetTransformation.inverse of the sampled trajectories
 *                             (inference.py:146-152 and the per-event loop :163-287,
 *                             utility/target_transformation.py:17-33) for all events and stored grid
 *                             points in one pass over packed cells.
 *   srpost_select_cells     : the SR -> pflow hand-off that the reference does through ROOT files
 *                             (inference.py:291-310 -> pflow/dataset_pf.py:81-92,136-147): keep the cells
 *                             with E_pred above the threshold, compact them per event (order preserved),
 *                             and derive the scaled pflow inputs (VarTransformation.forward, cos/sin phi).
 *
 * Same conventions as srhep.h: plain pointers and sizes, 0 / negative SRHEP_E_* codes, packed events, an
 * explicit cudaStream_t; the functions are stateless (current device).
 */
#ifndef SRHEP_POST_H_
#define SRHEP_POST_H_

#include <stddef.h>
#include <stdint.h>

#include "pflow.h"

#ifdef __cplusplus
extern "C" {
#endif

/* configs/single_e/model_and_var.yml:132-137 `target_transform` (utility/target_transformation.py):
 * inverse(y) = inv_trans(inv_scale(y)); inv_scale: y * std + mean (scale_mode "standard") or none;
 * inv_trans ("logit_ratio"): ((sigmoid(z) - alpha) / (1 - 2 alpha)) * proxy_raw * f. */
typedef struct SrpostTargetTransform {
    int32_t standard;              /* 1: scale_mode == "standard", 0: no scaling */
    float mean, std, alpha, f;
} SrpostTargetTransform;

/* samples_dev : (n_ens, n_store, T) fp32 -- member e, stored grid point s, packed cell
 * proxy_raw_dev : (T) e_proxy_raw [GeV];  unit = 1e3 (GeV -> MeV, inference.py:199)
 * nn_avg_dev    : (n_store, T) mean over members of the raw network output        (`raw_nn_pred*`)
 * e_avg_raw_dev : (n_store, T) inverse(mean over members)  * unit                  (`e_pred_avg_raw*`)
 * e_raw_dev     : (n_store, T) mean over members of inverse(member) * unit         (`e_pred_raw*`)
 * any output may be NULL. */
int srpost_ensemble_unscale(const float* samples_dev, int32_t n_ens, int32_t n_store, int64_t n_cells, const float* proxy_raw_dev,
                            const SrpostTargetTransform* tt, float unit, float* nn_avg_dev, float* e_avg_raw_dev, float* e_raw_dev,
                            void* stream);

/* In:  packed SR cells of B events (cu_in_dev: B + 1 offsets on the device): predicted energy [MeV], eta_raw, phi, layer.
 * Out: cells with e_pred > threshold, compacted per event in the original order, as the pflow model wants them
 *      (pflow/dataset_pf.py:136-147): PflowCells-compatible arrays of capacity T, and cu_out_dev (B + 1).
 *      tr_e / tr_eta: `var_transform` entries `e` and `eta` of the pflow config. */
typedef struct SrpostPflowOut {
    float* e; float* eta; float* cosphi; float* sinphi; float* phi; float* e_raw; float* eta_raw; int32_t* layer;
} SrpostPflowOut;
int srpost_select_cells(const float* e_pred_dev, const float* eta_raw_dev, const float* phi_dev, const int32_t* layer_dev,
                        const int32_t* cu_in_dev, int32_t n_events, float threshold, const PflowVarTransform* tr_e,
                        const PflowVarTransform* tr_eta, const SrpostPflowOut* out_dev, int32_t* cu_out_dev, void* stream);

const char* srpost_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SRHEP_POST_H_ */

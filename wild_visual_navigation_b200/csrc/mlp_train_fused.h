// wvn-b200: internal interface of the fused online train step (mlp_train_fused.cu).
#pragma once

#include <cuda_runtime.h>

#include "mlp_train.h"

namespace wvn {

// Device-resident scalars of one step.  The six leading doubles are plain sums (incl. the row count), so a
// data-parallel step all-reduces them in ONE call between the forward and the backward kernels.
struct FusedScalars {
  double sum_lr;    // sum of loss_reco over valid rows
  double sum_lr2;   // sum of loss_reco^2 over valid rows
  double sum_raw;   // sum of (trav - y)^2 over all live rows
  double n_valid;   // number of valid (labelled) rows
  double n_rows;    // number of live rows (the loss' N; global after the all-reduce)
  double reserved;
  float mean, std;  // ConfidenceGenerator state after the update
  float loss_total, loss_trav, loss_reco, loss_trav_conf;
};

struct FusedTrainer;

// scalars_ext (sizeof(FusedScalars) bytes) / grads_ext (n_params + 1 floats): caller-owned device buffers, or NULL to
// let the trainer allocate them with the rest of its workspace (everything is allocated here, nothing per step).
int fused_trainer_create(const MlpShape& s, int max_rows, const LossCfg& loss, const AdamCfg& adam, void* scalars_ext,
                         float* grads_ext, FusedTrainer** out);
void fused_trainer_destroy(FusedTrainer* t);
int fused_comm_unique_id(void* id128);
int fused_trainer_init_comm(FusedTrainer* t, const void* id128, int rank, int world);
// phase_mask: 1 = forward + statistics (+ their all-reduce), 2 = backward + weight gradients (+ gradient all-reduce),
// 4 = loss metrics + Adam; 7 = the whole step.
int fused_train_step(FusedTrainer* t, float* params, float* exp_avg, float* exp_avg_sq, long long* step_counter,
                     const float* x, int groups, int rows_per_group, const int* n_rows, const float* y,
                     const unsigned char* y_valid, float* cg_mean, float* cg_std, float* conf_out, float* metrics,
                     int phase_mask, cudaStream_t stream);

}  // namespace wvn

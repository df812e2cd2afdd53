/* ofb_train.h -- C ABI of the Q-learning update path in libofb.so (sm_100a).
 *
 * Replaces, for the bi-head pointer model, the Keras calls on the far side of the forward
 * (paths under /root/reference/ofighters):
 *   Trainer.replay's targets      agents/qlearnIA_V2.py:237-280   -> ofb_trainer_td_targets
 *   model.fit(x, y, epochs=1, batch_size=B)          :281-286    -> ofb_trainer_fit
 *        = one optimiser step on  mse(output1) + mse(output2)  with BatchNormalization in
 *          training mode (batch statistics, moving averages updated) and Adam(lr)   :188,308
 *   model weights (save / load_model)                :70,298      -> ofb_trainer_get_weights
 *
 * Conventions are those of ofb.h: int return codes, ofb_last_error(), caller-owned device buffers,
 * asynchronous on `stream`, no CPU fallback.  The handle owns the fp32 master weights, Adam's
 * moments and the activation workspace for max_batch samples (sized at create time; nothing is
 * allocated by fit).
 *
 * Weights travel as ONE flat fp32 array in Keras layer order with Keras layouts -- the order of
 * ofighters_b200.policy.WEIGHT_SPEC: conv1/kernel [3,3,2,8], conv1/bias, norm1/gamma, beta, mean,
 * var, conv2 ... conv4, dense1/kernel [5008,100], bias, dense2, output1, updense1, upconv1/kernel,
 * bias, upnorm1/gamma, beta, mean, var, ... upconv4/kernel [3,3,8,1], bias  (571 730 values).
 *
 * Arithmetic: fp32 throughout (Keras' default dtype), sums over a batch accumulated in fp64.
 * Keras defaults that matter: BatchNormalization momentum 0.99, epsilon 1e-3, batch variance is the
 * biased one; Adam beta1 0.9, beta2 0.999, epsilon 1e-7, update
 *   lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  p -= lr_t * m / (sqrt(v) + epsilon).
 * Keras / TensorFlow are un-pinned and absent: PARITY UNPINNED, the checker is the torch-autograd
 * restatement oracle/policy_train_torch.py.
 */
#ifndef OFB_TRAIN_H
#define OFB_TRAIN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_TRAIN_N_PARAMS 571730

typedef struct ofb_trainer ofb_trainer;

typedef struct ofb_train_config {
    float lr;              /* Adam(lr=0.0001)                     qlearnIA_V2.py:304,308 */
    float beta1, beta2, adam_eps;
    float bn_momentum, bn_eps;
    int32_t bn_unbiased_moving_var;   /* 0 = Keras (biased), 1 = TF fused kernels (n / (n - 1)) */
    int32_t max_batch;     /* samples per fit the workspace is sized for (reference: batch_size = 8) */
    int32_t reserved[8];
} ofb_train_config;

void ofb_train_default_config(ofb_train_config *cfg);

int ofb_trainer_create(const float *weights_flat_host, int64_t n_params, const ofb_train_config *cfg, int device,
                       ofb_trainer **out);
int ofb_trainer_destroy(ofb_trainer *t);

/* model.fit on one batch (one Adam step).
 *   maps_bits_dev   uint32 [B, 2, 5000]    the samples' images (bit y*400+x; ch0 ship_map, ch1 laser_map)
 *   vec_dev         float  [B, 8]
 *   target_act_dev  float  [B, 2]          y for output1
 *   target_ptr_dev  float  [B, 400, 400]   y for output2
 *   loss_dev        float  [3]             (total, mse(output1), mse(output2)) of the batch BEFORE the update,
 *                                          i.e. history['loss'][0]; may be NULL */
int ofb_trainer_fit(ofb_trainer *t, const uint32_t *maps_bits_dev, const float *vec_dev, const float *target_act_dev,
                    const float *target_ptr_dev, int batch, float *loss_dev, void *stream);

/* Training-mode forward only (the predictions fit() differentiates): act [B,2], ptr [B,400,400], either may be NULL. */
int ofb_trainer_forward(ofb_trainer *t, const uint32_t *maps_bits_dev, const float *vec_dev, int batch, float *act_dev,
                        float *ptr_dev, void *stream);

/* Trainer.replay's targets (agents/qlearnIA_V2.py:262-280) from predictions on obs and next_obs:
 *   target[b]      = act_obs[b];      target[b][iaction[b]]            = r[b] + gamma * max(act_next[b]) * !done[b]
 *   ptr_target[b]  = ptr_obs[b];      ptr_target[b][px[b]][py[b]]      = r[b] + gamma * max(ptr_next[b]) * !done[b]
 * (the reference indexes the [row, col] map with the (x, y) pointer tuple -- reproduced).  act_* float [B,2], ptr_* float
 * [B,400,400], iaction int32 [B], pointer int32 [B,2] = (x, y), reward float [B], done uint8 [B].  target_act_dev may alias
 * act_obs_dev and target_ptr_dev may alias ptr_obs_dev. */
int ofb_trainer_td_targets(const float *act_obs_dev, const float *ptr_obs_dev, const float *act_next_dev,
                           const float *ptr_next_dev, const int32_t *iaction_dev, const int32_t *pointer_dev,
                           const float *reward_dev, const uint8_t *done_dev, float gamma, int batch,
                           float *target_act_dev, float *target_ptr_dev, void *stream);

/* Flat copies to HOST (synchronise `stream` first): current weights / gradients of the last fit. */
int ofb_trainer_get_weights(ofb_trainer *t, float *weights_flat_host, void *stream);
int ofb_trainer_get_grads(ofb_trainer *t, float *grads_flat_host, void *stream);
int64_t ofb_trainer_steps(const ofb_trainer *t);   /* optimiser steps taken (Adam's t) */

#ifdef __cplusplus
}
#endif
#endif /* OFB_TRAIN_H */

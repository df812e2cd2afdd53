/* ofb_policy.h -- C ABI of the bi-head pointer policy forward in libofb.so (sm_100a).
 *
 * Replaces, for batches of arenas, the Keras calls of the reference's Trainer
 * (paths under /root/reference/ofighters):
 *   model definition            agents/qlearnIA_V2.py:123-190   -> ofb_policy_create (weights in Keras layouts)
 *   model.predict([img, vec])   agents/qlearnIA_V2.py:208-215   -> ofb_policy_forward (act, ptr)
 *   argmax decode               agents/qlearnIA_V2.py:218-220   -> ofb_policy_forward (iaction, xy)
 *   random_play / eps-greedy    agents/qlearnIA_V2.py:199-204,317-321 + QlearnIA.play :447-456
 *                                                                -> ofb_policy_write_actions
 *
 * Conventions are those of ofb.h: int return codes, ofb_last_error(), caller-owned device buffers,
 * asynchronous on `stream`, no CPU fallback.  The handle owns the folded weights and a workspace
 * for the intermediate activations (sized at create time; nothing is allocated by forward).
 *
 * Arithmetic: BatchNormalization (inference, eps 1e-3) is folded into the conv weights; conv2-4,
 * the 5000-wide flat slice of dense1 and upconv3-4 run as bf16 x bf16 -> fp32 tensor-core
 * contractions (tcgen05) with bf16 activations between layers; conv1 (binary input), the 8-value
 * vector slice of dense1, dense2, output1, updense1, upconv1 and upconv2 stay in fp32.  Bilinear x2 upsampling uses TF2 half-pixel
 * centres with edge clamp by default (ofb_policy_create_opts offers the TF1.x legacy kernel).
 */
#ifndef OFB_POLICY_H
#define OFB_POLICY_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ofb_policy ofb_policy;

/* One conv layer in Keras layout: kernel HWIO [3][3][cin][cout], bias [cout], and the following
 * BatchNormalization's gamma/beta/moving_mean/moving_variance [cout] (all NULL = no BN). HOST fp32. */
typedef struct ofb_conv_weights {
    const float *kernel, *bias, *gamma, *beta, *mean, *var;
} ofb_conv_weights;

/* Dense layer in Keras layout: kernel [in][out], bias [out].  HOST fp32. */
typedef struct ofb_dense_weights {
    const float *kernel, *bias;
} ofb_dense_weights;

typedef struct ofb_policy_weights {
    ofb_conv_weights conv[4];      /* conv1 2->8, conv2..4 8->8        qlearnIA_V2.py:129-147 */
    ofb_dense_weights dense1;      /* 5008 -> 100 ([vector(8), flat(5000)])          :154-155 */
    ofb_dense_weights dense2;      /* 100 -> 50                                      :158     */
    ofb_dense_weights output1;     /* 50 -> 2                                        :160     */
    ofb_dense_weights updense1;    /* 100 -> 625                                     :163     */
    ofb_conv_weights upconv[4];    /* 1->2, 2->4, 4->8, 8->1 (last one without BN)   :166-186 */
} ofb_policy_weights;

enum { OFB_ENGINE_TENSOR = 0,      /* tcgen05 kernels (product path) */
       OFB_ENGINE_CUDA_CORE = 1 }; /* same arithmetic on CUDA cores: the numerical twin used to
                                      validate the tensor-core kernels layer by layer */

/* max_ships = largest n_arenas * ships_per_arena a forward call may carry per chunk (the library
 * loops over chunks of that size); workspace = 0.18 MB per ship, plus 1.5 MB per ship allocated on the first use of
 * an alternative kernel or of the validation taps.  0 = default (1024). */
int ofb_policy_create(const ofb_policy_weights *w_host, int device, int max_ships, ofb_policy **out);
/* The same with options (flags, OR-ed):
 *   OFB_POLICY_BILINEAR_TF1   UpSampling2D(interpolation='bilinear') (agents/qlearnIA_V2.py:166,171,177,183) with the TF1.x /
 *                             standalone-Keras legacy kernel (out[2i] = in[i], out[2i+1] = (in[i] + in[i+1]) / 2) instead of TF2's
 *                             half-pixel centres: the reference pins no Keras / TensorFlow version, so both are offered
 *   OFB_POLICY_UNFUSED_TAIL   upconv3 and upconv4 as two kernels through HBM (measurement aid; default = one fused kernel)
 *   OFB_POLICY_DENSE_TRUNK    conv1 + conv2 on the dense tcgen05 kernel instead of the sparse one
 *   OFB_POLICY_CC_SPARSE_TRUNK  the sparse trunk with conv2 on CUDA cores (round 1's kernel) instead of tensor cores
 *   OFB_POLICY_UNFUSED_TRUNK  conv1+conv2, conv3, conv4 as three kernels through HBM (default: one sparse kernel for the whole trunk)
 *   OFB_POLICY_TAIL_PAIR      the fused tail as CTA pairs (tcgen05 cta_group::2, M = 256 MMAs over two SMs that share the weight operand);
 *                             identical results, measured no faster than one CTA per SM, kept as a selectable experiment */
enum { OFB_POLICY_BILINEAR_TF1 = 1, OFB_POLICY_UNFUSED_TAIL = 2, OFB_POLICY_DENSE_TRUNK = 4, OFB_POLICY_CC_SPARSE_TRUNK = 8,
       OFB_POLICY_UNFUSED_TRUNK = 16, OFB_POLICY_TAIL_PAIR = 32 };
int ofb_policy_create_opts(const ofb_policy_weights *w_host, int device, int max_ships, int flags, ofb_policy **out);
int ofb_policy_destroy(ofb_policy *p);
/* load new weights into an existing handle (Trainer.fit refreshed them; load_model, agents/qlearnIA_V2.py:70): folds like
 * ofb_policy_create and overwrites the resident copy after `stream` has drained; the workspace is kept. */
int ofb_policy_set_weights(ofb_policy *p, const ofb_policy_weights *w_host, void *stream);
int ofb_policy_set_engine(ofb_policy *p, int engine);
/* validation taps: when enabled the fused tail kernel also writes upconv3's output for ofb_policy_debug_tap(6) */
int ofb_policy_set_taps(ofb_policy *p, int enable);

/* model.predict + decode for n_arenas arenas with ships_per_arena policy-driven ships each.
 *   maps_bits_dev  uint32 [A, 2, 5000]   bit y*400+x, ch0 ship_map, ch1 laser_map (ofb_raster OFB_MAP_BITS)
 *   vec_dev        float  [A*P, 8]       observation heads of the policy ships, arena-major
 * outputs (each may be NULL):
 *   act_dev        float  [A*P, 2]       output1
 *   ptr_dev        float  [A*P, 400,400] output2 (dense pointer map; only for predict()-style callers)
 *   iaction_dev    int32  [A*P]          argmax(act), ties -> lowest index
 *   xy_dev         int32  [A*P, 2]       (x, y) = (k % 400, k / 400), k = first flat argmax of ptr[row, col]
 */
int ofb_policy_forward(ofb_policy *p, const uint32_t *maps_bits_dev, const float *vec_dev, int64_t n_arenas,
                       int ships_per_arena, float *act_dev, float *ptr_dev, int32_t *iaction_dev, int32_t *xy_dev,
                       void *stream);

/* Exploration schedule of the shared trainer (lib/epsilon.py:36-86 behind Trainer.epsilon, agents/qlearnIA_V2.py:53,195-201)
 * as a closed-form function of the number t of decay_epsilon() calls made so far, so that the device evaluates it itself:
 *   OFB_EPS_CONST   eps(t) = start
 *   OFB_EPS_COSINE  eps(t) = start * (cos(2 pi (t mod period) / period) + 1) / 2          (Epsilon_cos, amplitude = start)
 *   OFB_EPS_DECAY   eps(t) = start * decay^min(t, n),  n = first step count with start * decay^n <= floor  (Epsilon_decay:
 *                   the reference stops multiplying once epsilon is at or below its soft minimum)                        */
enum { OFB_EPS_CONST = 0, OFB_EPS_COSINE = 1, OFB_EPS_DECAY = 2 };
typedef struct ofb_eps_schedule {
    int32_t kind;
    int32_t reserved;
    double start, period, decay, floor;  /* doubles: 0.9999^t must not drift from the host's value over 10^5 steps */
} ofb_eps_schedule;

/* QlearnIA.play's action vector (agents/qlearnIA_V2.py:447-456): row (shoot, thrust, x, y) =
 * (iaction == 0, iaction == 1, x, y) written to actions[arena, ship_index[p], :] (int16 [A,S,4]).
 * With epsilon > 0 a ship acts randomly (iaction U{0,1}, x,y U{0..399}; :199-204,317-321) when its
 * Philox draw keyed by (seed; arena0 + arena, ship, step) is <= epsilon. */
int ofb_policy_write_actions(const int32_t *iaction_dev, const int32_t *xy_dev, int64_t n_arenas, int ships_per_arena,
                             const int32_t *ship_index_dev, int n_ships_total, float epsilon, uint64_t seed,
                             int64_t arena0, uint32_t step, int16_t *actions_dev, void *stream);

/* The same with the trainer's schedule evaluated on the device at schedule step t, and with the action that is PLAYED
 * written back over iaction_dev / xy_dev (in / out), which is what QlearnIA.play remembers as previous_action /
 * previous_pointer (:399-401).  force_random != 0 = the collecting phase (total_steps < collecting_steps, :394-396):
 * every row is random_play() and iaction_dev / xy_dev need not hold a prediction.  eps_out_dev (optional, float[1])
 * receives the epsilon that was used. */
int ofb_policy_play_actions(int32_t *iaction_dev, int32_t *xy_dev, int64_t n_arenas, int ships_per_arena,
                            const int32_t *ship_index_dev, int n_ships_total, const ofb_eps_schedule *sched, double t,
                            int force_random, uint64_t seed, int64_t arena0, uint32_t step, int16_t *actions_dev,
                            float *eps_out_dev, void *stream);
/* eps(t) evaluated on the host by the same code (logging: QlearnIA.epsilons, agents/qlearnIA_V2.py:364). */
float ofb_eps_value(const ofb_eps_schedule *sched, double t);

/* Dense image [B,400,400,2] (NHWC; fmt OFB_MAP_BF16 / OFB_MAP_U8 / 3 = float32) -> bit maps, for
 * predict([img, vec]) callers that hold Keras-style images.  A pixel is set iff it is non-zero. */
int ofb_policy_pack_image(const void *img_dev, int fmt, int64_t n, uint32_t *maps_bits_dev, void *stream);

/* Per-layer device times (measurement aid): enable = 1 brackets every kernel of the following forward
 * calls with CUDA events; enable = 0 stops, synchronises and writes the accumulated milliseconds per
 * layer to ms_out[9] = (trunk12, conv3, conv4, dense1, heads, up3, up4, argmax, tail = fused up3 + up4 + argmax). */
int ofb_policy_profile(ofb_policy *p, int enable, float *ms_out);

/* Debug / validation tap: copies the first n_items entries of an intermediate activation of the
 * LAST forward call's first chunk into dst_dev (device, n_items * stride elements).
 * which (stride in elements): 0 pool1 [200,200,8] (320000, cuda-core engine only), 1 pool2 [100,100,8]
 * (80000), 2 pool3 [50,50,8] (20000), 3 flat [5000] (5120), 4 dense1 flat-part pre-activation [100]
 * (100, fp32), 5 up2 [100,100,8] (80000), 6 up3 [200,200,8] (320000).  bf16 except (4). */
int ofb_policy_debug_tap(ofb_policy *p, int which, int64_t n_items, void *dst_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* OFB_POLICY_H */

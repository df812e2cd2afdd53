/* ofb.h -- C ABI of libofb.so: the B200 (sm_100a) drop-in for Ofighters' data-parallel hot path.
 *
 * The reference (Chthi/Ofighters) is pure Python and has no FFI of its own; this header declares
 * the entry points a Python binding of that path needs, each citing the reference interface it
 * replaces (paths under /root/reference/ofighters).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (OFB_E_*); the message of the last error
 *     on the calling thread is ofb_last_error().
 *   - every pointer marked "dev" is DEVICE memory owned by the caller; the handle owns only its
 *     arena state allocated at create time.  Nothing is allocated on the step path.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All calls are
 *     asynchronous on that stream unless stated otherwise.
 *   - a handle is bound to one device and is not thread-safe.  There is no CPU fallback: without
 *     a CUDA device ofb_create fails with OFB_E_CUDA.
 *
 * Tensor formats
 *   actions   int16 [N, S, 4] = (shoot, thrust, pointing_x, pointing_y)     lib/action.py:12-56
 *   obs_vec   float [N, S, 8] = (reward, can_shoot=1, pointing_x, pointing_y, W, H, x, y)
 *                                                                            lib/observation.py:113-123
 *   spawn     int32 [N, S, 2] = (x, y), 0..W / 0..H inclusive                lib/battleground.py:114
 *   maps      OFB_MAP_BITS : uint32 [N, 2, W*H/32], bit (y*W + x) LSB-first, ch0 ships, ch1 lasers
 *             OFB_MAP_BF16 : bf16   [N, H, W, 2]  (NHWC, the Keras input of qlearnIA_V2.py:208)
 *             OFB_MAP_U8   : uint8  [N, H, W, 2]
 */
#ifndef OFB_H
#define OFB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OFB_ABI_VERSION 1

enum {
    OFB_OK = 0,
    OFB_E_ARG = -1,      /* bad argument */
    OFB_E_CUDA = -2,     /* CUDA runtime error (message in ofb_last_error) */
    OFB_E_NOMEM = -3,
    OFB_E_STATE = -4     /* e.g. laser-slot overflow detected */
};

enum { OFB_MAP_BITS = 0, OFB_MAP_BF16 = 1, OFB_MAP_U8 = 2 };

/* scripted bots of agents/agent.py:99-155 (+ the stress distribution of SURVEY 8(d) config 4) */
enum { OFB_BOT_IDLE = 0, OFB_BOT_RANDOM = 1, OFB_BOT_TURRET = 2, OFB_BOT_RUNNER = 3,
       OFB_BOT_THRUST = 4, OFB_BOT_SHOOT = 5, OFB_BOT_STRESS = 6,
       OFB_BOT_EXTERNAL = 255 /* leave the ship's action row untouched (policy / host bot) */ };

/* Arena constants; defaults = the reference's module constants (SURVEY section 5 "config"). */
typedef struct ofb_config {
    int32_t n_ships;        /* ships per arena, 1..32          lib/ofighters.py:53 (7)       */
    int32_t laser_cap;      /* laser slots per arena, 0 = auto (max(128, 16*n_ships))         */
    int32_t width, height;  /* lib/observation.py:10-11 (400, 400); width*height % 32 == 0    */
    int32_t max_time;       /* lib/ofighters.py:59 (200) -- informational, host enforces it   */
    int32_t reward_kill, reward_death, reward_aim, reward_trajectory;  /* qlearnIA_V2.py:39-44 */
    int32_t reserved[7];
} ofb_config;

typedef struct ofb_arenas ofb_arenas;   /* opaque: N arenas' struct-of-arrays state on one GPU */

/* Flat views for export / import (all dev pointers, any may be NULL to skip). */
typedef struct ofb_state_view {
    int32_t *time, *n_lasers, *kills, *deaths, *shots, *overflow, *episode, *near_ties;  /* [N]   */
    int32_t *ship_x, *ship_y, *ship_px, *ship_py, *ship_hull, *ship_reward,
            *ship_score, *ship_steps;                                                     /* [N,S] */
    uint8_t *ship_alive;                                                                  /* [N,S] */
    double  *laser_x, *laser_y, *laser_dx, *laser_dy;                                     /* [N,L] */
    uint8_t *laser_owner, *laser_destroyed;                                               /* [N,L] */
} ofb_state_view;

int         ofb_abi_version(void);
const char *ofb_last_error(void);
void        ofb_default_config(ofb_config *cfg);

/* Battleground.__init__ for N arenas (lib/battleground.py:13-106): ships placed at spawn[N,S,2]
 * (dev, int32), pointing = own position, hull 1, flying, no lasers, time 0. */
int ofb_create(const ofb_config *cfg, int64_t n_arenas, int device, const int32_t *spawn_dev,
               void *stream, ofb_arenas **out);
int ofb_destroy(ofb_arenas *h);
int ofb_laser_cap(const ofb_arenas *h);
int64_t ofb_state_stride(const ofb_arenas *h);   /* bytes of HBM per arena (for sizing) */

/* Battleground.restart + Ship.reset + Agent.reset (lib/battleground.py:108-117, lib/ship.py:92-106,
 * agents/agent.py:59-64) for the arenas with mask[a] != 0 (mask NULL = all).  Before zeroing, the
 * episode's [sum score, kills, deaths, shots, ships, arenas] are added into stats_dev[6] (int64,
 * may be NULL) -- the payload of the per-episode all-reduce. */
int ofb_reset(ofb_arenas *h, const uint8_t *mask_dev, const int32_t *spawn_dev, int64_t *stats_dev,
              void *stream);

/* Battleground.generate_frame(actions) (lib/battleground.py:153-160) preceded by the score fold of
 * Agent.step (agents/agent.py:66-74) and the destroyed-laser pruning of the Tk controller
 * (lib/ofighters.py:704-707).  If obs_out_dev != NULL the per-ship observation head of the NEXT
 * request_actions() is written there (float [N,S,8]). */
int ofb_step(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, void *stream);

/* Same step driven from HOST buffers (a host-side bot loop calling Battleground.frame,
 * lib/battleground.py:163-166): actions_host int16 [N,S,4] is copied to the device, the step runs,
 * and the next observation heads are copied back to obs_host float [N,S,8] (may be NULL).  Both
 * host buffers should be pinned; the call is asynchronous on `stream` -- synchronise before reading. */
int ofb_step_host(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *stream);

/* Pipelined form of ofb_step_host for a loop that feeds one action batch per frame (a recorded action tape, or a
 * host bot that does not need frame k's observations before sending frame k+1): the H2D copy of the next frame
 * and the D2H copy of the previous frame's observation heads run on streams owned by the handle and overlap the
 * kernels on `stream`.  obs_host is valid, and actions_host reusable, after ofb_host_wait(). */
int ofb_step_host_async(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *stream);
int ofb_host_wait(ofb_arenas *h);

/* Running statistics of the CURRENT episode, without resetting anything: [sum over ships of score + pending reward, kills,
 * deaths, shots, ships, arenas] is ADDED into stats_dev[6] (int64; zero it first for totals) -- what the reference's score
 * consumers read mid-episode (agent.score, battleground.last_x_time_rewards: lib/laser.py:58, lib/ship.py:229). */
int ofb_stats(const ofb_arenas *h, int64_t *stats_dev, void *stream);

/* Observation.analyse_ship head for every ship (lib/observation.py:101-123). */
int ofb_obs_vec(const ofb_arenas *h, float *out_dev, void *stream);

/* Observation.analyse_battleground (lib/observation.py:79-95): ship_map / laser_map of every arena. */
int ofb_raster(const ofb_arenas *h, void *out_dev, int format, void *stream);

/* Scripted bots (agents/agent.py:99-155): counter-based Philox4x32-10 keyed by (seed; global arena
 * id = arena0 + a, ship, step).  kinds_dev (uint8 [S], may be NULL) gives one bot kind per ship
 * index like the reference's ships={"behavior": n} map (lib/battleground.py:21-28,79-81);
 * when NULL every ship uses bot_kind. */
int ofb_bot_actions(const ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed,
                    int64_t arena0, uint32_t step, int16_t *actions_dev, void *stream);
/* request_actions + generate_frame in one launch (Battleground.frame, lib/battleground.py:163-166, with the
 * scripted bots of agents/agent.py): every ship whose kind is not OFB_BOT_EXTERNAL draws its action inside the
 * step kernel (same Philox counters as ofb_bot_actions, so results are identical to the two-call form);
 * external ships read their row from actions_dev (may be NULL when no ship is external). */
int ofb_step_bots(ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed, int64_t arena0, uint32_t step,
                  const int16_t *actions_dev, float *obs_out_dev, void *stream);

/* Battleground.frame (lib/battleground.py:163-166) as ONE launch: generate_frame followed by the new Observation's
 * ship_map / laser_map (lib/observation.py:79-95) written to maps_bits_dev in OFB_MAP_BITS format.  A persistent,
 * warp-specialised kernel steps the arenas and rasterises them from shared memory, so the state is read once and the
 * step's latency hides behind the raster's HBM writes.  Results are bit-identical to ofb_step / ofb_step_bots followed
 * by ofb_raster(OFB_MAP_BITS); a configuration that does not fit the kernel's shared-memory ring runs as those two
 * launches (ofb_debug_last_frame_fused tells which). */
int ofb_frame(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, void *maps_bits_dev, void *stream);
int ofb_frame_bots(ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed, int64_t arena0, uint32_t step,
                   const int16_t *actions_dev, float *obs_out_dev, void *maps_bits_dev, void *stream);
/* ofb_step_host_async with the fused frame kernel (maps stay in HBM at maps_bits_dev). */
int ofb_frame_host_async(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *maps_bits_dev, void *stream);
/* The 5 entries of an observation head that are not constants (lib/observation.py:119-123: can_shoot is always 1, dim is the
 * map's size) as int16 [n_rows, 5] = [reward, pointing.x, pointing.y, pos.x, pos.y] -- all exact small integers: 10 instead of
 * 32 bytes per ship on the device -> host path. */
int ofb_obs_pack_i16(const float *obs_dev, int16_t *out_dev, int64_t n_rows, void *stream);
/* ofb_frame_host_async with the observation heads copied back in that compact form (obs16_host: pinned int16 [N, S, 5]). */
int ofb_frame_host_async_i16(ofb_arenas *h, const int16_t *actions_host, int16_t *obs16_host, void *maps_bits_dev, void *stream);

/* Debug aid for profiling the fused frame kernel: when buf_dev != NULL every following ofb_frame* launch writes
 * per-warp-role cycle counters to it (int64 [grid][32][8]); NULL switches it off. */
int ofb_debug_frame_prof(long long *buf_dev);
/* 1 if the calling thread's last ofb_frame* call ran as the single fused launch, 0 if as the two-launch form, -1 before. */
int ofb_debug_last_frame_fused(void);

/* randint(0, W) x randint(0, H) spawn draws of lib/battleground.py:79-81,114. */
int ofb_random_spawn(int64_t n_arenas, int n_ships, int width, int height, uint64_t seed, int64_t arena0,
                     uint32_t episode, int32_t *spawn_dev, void *stream);

int ofb_state_export(const ofb_arenas *h, const ofb_state_view *view, void *stream);
int ofb_state_import(ofb_arenas *h, const ofb_state_view *view, void *stream);

/* ---- policy forward (agents/qlearnIA_V2.py:123-235), see ofb_policy.h ---- */

#ifdef __cplusplus
}
#endif
#endif /* OFB_H */

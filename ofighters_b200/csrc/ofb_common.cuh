// Shared definitions for libofb's arena kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ofb.h"

#define OFB_R_SHIP 8        // lib/ship.py:43
#define OFB_R_LASER 2       // lib/laser.py:23
#define OFB_SHIP_SPEED 8    // lib/ship.py:24
#define OFB_LASER_SPEED 10  // lib/ship.py:86 x lib/laser.py:13

// Per-arena block in HBM.  Everything one arena owns is contiguous, and the part a step needs -- header, ships and
// the LIVE prefix of the laser list -- is one run of bytes, so that a tile can pull it into shared memory with one or
// two bulk asynchronous copies (cp.async.bulk) of exactly the bytes that are live:
//
//   +0          int32 hdr[8]   = time, n_lasers, kills, deaths, shots, overflow, episode, near_ties
//   +32         uint4 ship[SP] (SP = S rounded up to 8), one 16-byte record per ship:
//                   .x = x | y << 16 | alive << 31      (0 <= x, y <= 16384)
//                   .y = (pointing_x & 0xffff) | pointing_y << 16
//                   .z = score
//                   .w = (pending reward & 0xffff) | hull << 16   (hull saturates at -32768; only its sign matters)
//               agent.steps is not stored: it equals hdr.time (both ++ every frame, both zeroed by a reset)
//   +off_laser  laser groups of 8 slots, 288 B each; lasers are a dense, append-ordered list [0, n_lasers):
//                   double x[8], y[8], dx[8], dy[8]   (dx, dy = per-frame increment, fixed at spawn, lib/laser.py:39-47)
//                   uint32 meta[8]                    bits 0-7 owner ship, bit 8 destroyed
//   stride = total rounded up to 128 B
enum { HDR_TIME = 0, HDR_NLASERS, HDR_KILLS, HDR_DEATHS, HDR_SHOTS, HDR_OVERFLOW, HDR_EPISODE, HDR_NEARTIES };
#define OFB_GROUP_BYTES 288
#define OFB_G_Y 64
#define OFB_G_DX 128
#define OFB_G_DY 192
#define OFB_G_META 256

struct ArenaLayout {
    int S, SP, L, W, H;
    int off_ship, off_laser;
    int stride;
    int r_kill, r_death, r_aim, r_traj;
};

struct ofb_arenas {
    ArenaLayout lay;
    ofb_config cfg;
    int64_t n_arenas;
    int device;
    char *state;      // n_arenas * stride bytes
    int16_t *stage_actions;   // [N,S,4] device staging for ofb_step_host
    float *stage_obs;         // [N,S,8]
    // pipelined host frames (ofb_step_host_async): double-buffered staging, one stream per copy direction
    int16_t *pipe_actions[2];
    float *pipe_obs[2];
    int16_t *pipe_obs16[2];        // compact observation heads (ofb_obs_pack_i16) of the pipelined host loop
    cudaStream_t s_h2d, s_d2h;
    cudaEvent_t ev_h2d[2], ev_step[2], ev_d2h[2];
    unsigned long long host_seq;
    int pipe_ready;
};

static inline ArenaLayout make_layout(const ofb_config &c) {
    ArenaLayout l;
    l.S = c.n_ships;
    l.SP = (c.n_ships + 7) & ~7;
    l.L = c.laser_cap;
    l.W = c.width;
    l.H = c.height;
    l.off_ship = 32;
    l.off_laser = l.off_ship + 16 * l.SP;
    l.stride = (l.off_laser + (l.L / 8) * OFB_GROUP_BYTES + 127) & ~127;
    l.r_kill = c.reward_kill;
    l.r_death = c.reward_death;
    l.r_aim = c.reward_aim;
    l.r_traj = c.reward_trajectory;
    return l;
}

// ---- packed ship record ----
struct ShipRec {
    int x, y, px, py, score, reward, hull;
    bool alive;
};
__host__ __device__ __forceinline__ ShipRec ship_unpack(const uint4 v) {
    ShipRec r;
    r.x = (int)(v.x & 0xffffu);
    r.y = (int)((v.x >> 16) & 0x7fffu);
    r.alive = (v.x >> 31) != 0u;
    r.px = (int)(short)(v.y & 0xffffu);
    r.py = (int)(short)(v.y >> 16);
    r.score = (int)v.z;
    r.reward = (int)(short)(v.w & 0xffffu);
    r.hull = (int)(short)(v.w >> 16);
    return r;
}
__host__ __device__ __forceinline__ uint4 ship_pack(const ShipRec &r) {
    const int hull = r.hull < -32768 ? -32768 : r.hull;
    uint4 v;
    v.x = (unsigned)r.x | ((unsigned)r.y << 16) | (r.alive ? 0x80000000u : 0u);
    v.y = ((unsigned)r.px & 0xffffu) | ((unsigned)r.py << 16);
    v.z = (unsigned)r.score;
    v.w = ((unsigned)r.reward & 0xffffu) | ((unsigned)hull << 16);
    return v;
}
// address of laser slot k's x inside an arena block (y, dx, dy at +OFB_G_*; meta at laser_meta_off)
__host__ __device__ __forceinline__ int laser_off(int off_laser, int k) { return off_laser + (k >> 3) * OFB_GROUP_BYTES + (k & 7) * 8; }
__host__ __device__ __forceinline__ int laser_meta_off(int off_laser, int k) {
    return off_laser + (k >> 3) * OFB_GROUP_BYTES + OFB_G_META + (k & 7) * 4;
}

// ---- Philox4x32-10 (Salmon et al., SC'11); counter = (arena, ship, step, stream), key = seed ----
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
    return (uint32_t)(((uint64_t)a * b) >> 32);
}

// Scripted bots of agents/agent.py:99-155 (+ the stress distribution of SURVEY 8(d) config 4).  bot_draw() is the random
// part for ship `ship` of global arena `arena` at frame `step` (it does not depend on the arena's state, so the fused step
// kernel draws it while its state is still in flight); bot_action() completes it into the action row
// (shoot | thrust << 16, px | py << 16), keeping the ship's current pointing unless the bot re-points.
struct BotDraw {
    int shoot, thrust, repoint, rx, ry;
};
// the counter-based random words of (arena, ship, step): independent of the bot kind and of the arena's state
__device__ __forceinline__ uint4 bot_random(uint64_t seed, long long arena, int ship, uint32_t step) {
    uint32_t c[4] = {(uint32_t)arena, (uint32_t)ship, step, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return make_uint4(c[0], c[1], c[2], c[3]);
}
__device__ __forceinline__ BotDraw bot_interpret(int kind, const uint4 c, int W, int H) {
    BotDraw d = {0, 0, 0, (int)mulhi32(c.y, (uint32_t)W + 1), (int)mulhi32(c.z, (uint32_t)H + 1)};
    switch (kind) {
    case OFB_BOT_RANDOM: {                                    // agents/agent.py:123-133
        const uint32_t k = mulhi32(c.x, 3u);
        d.shoot = k == 0; d.thrust = k == 1; d.repoint = k == 2;
    } break;
    case OFB_BOT_TURRET:                                      // agents/agent.py:136-144
        d.shoot = mulhi32(c.x, 10u) < 8;
        d.repoint = mulhi32(c.w, 10u) < 3;
        break;
    case OFB_BOT_RUNNER:                                      // agents/agent.py:147-155
        d.thrust = mulhi32(c.x, 10u) < 9;
        d.repoint = mulhi32(c.w, 10u) < 1;
        break;
    case OFB_BOT_THRUST: d.thrust = 1; break;                 // agents/agent.py:107-112
    case OFB_BOT_SHOOT: d.shoot = 1; break;                   // agents/agent.py:115-120
    case OFB_BOT_STRESS:                                      // agents/qlearnIA_V2.py:317-321, shoot forced
        d.shoot = 1; d.thrust = (int)(c.x & 1u); d.repoint = 1;
        d.rx = (int)mulhi32(c.y, (uint32_t)W); d.ry = (int)mulhi32(c.z, (uint32_t)H);
        break;
    default: break;                                           // idle: agents/agent.py:99-104
    }
    return d;
}
__device__ __forceinline__ BotDraw bot_draw(int kind, uint64_t seed, long long arena, int ship, uint32_t step, int W, int H) {
    return bot_interpret(kind, bot_random(seed, arena, ship, step), W, H);
}
__device__ __forceinline__ int2 bot_apply(const BotDraw &d, int px, int py) {
    if (d.repoint) { px = d.rx; py = d.ry; }
    return make_int2((d.shoot & 0xffff) | (d.thrust << 16), (px & 0xffff) | (py << 16));
}
__device__ __forceinline__ int2 bot_action(int kind, uint64_t seed, long long arena, int ship, uint32_t step, int px, int py,
                                           int W, int H) {
    return bot_apply(bot_draw(kind, seed, arena, ship, step, W, H), px, py);
}

void ofb_set_error(const char *fmt, ...);
#define OFB_CUDA_CHECK(expr)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            ofb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return OFB_E_CUDA;                                                            \
        }                                                                                 \
    } while (0)

// Opt-in dynamic shared memory is a per-device (per-context) attribute of a kernel: a call site remembers, per calling thread
// AND per device, the largest size it has configured, so that one thread driving handles on several GPUs configures each.
struct SmemAttrCache {
    int sz[32];
    template <class K> cudaError_t ensure(K kernel, int bytes) {
        int d = 0;
        cudaGetDevice(&d);
        int &c = sz[d & 31];
        if (bytes <= c) return cudaSuccess;
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) c = bytes;
        return e;
    }
};

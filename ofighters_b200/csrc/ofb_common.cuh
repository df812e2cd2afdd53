// Shared definitions for libofb's arena kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ofb.h"

#define OFB_R_SHIP 8        // lib/ship.py:43
#define OFB_R_LASER 2       // lib/laser.py:23
#define OFB_SHIP_SPEED 8    // lib/ship.py:24
#define OFB_LASER_SPEED 10  // lib/ship.py:86 x lib/laser.py:13

// Per-arena block in HBM.  Everything one arena owns is contiguous so that the warp (or
// sub-warp tile) that steps it touches a single run of 32-byte sectors:
//
//   +0    int32 hdr[8]   = time, n_lasers, kills, deaths, shots, overflow, episode, near_ties
//   +32   int32 ship[8][SP]  (struct-of-arrays over ships; SP = S rounded up to 8)
//                          x, y, px, py, reward, score, steps, flags(bit0 alive | hull<<8)
//   +off_lx     double laser_x[L]      lasers are a dense, append-ordered list [0, n_lasers)
//   +off_ly     double laser_y[L]
//   +off_ldx    double laser_dx[L]     per-frame increment, fixed at spawn (lib/laser.py:39-47)
//   +off_ldy    double laser_dy[L]
//   +off_lmeta  uint32 laser_meta[L]   bits 0-7 owner ship, bit 8 destroyed
//   stride = total rounded up to 128 B
enum { HDR_TIME = 0, HDR_NLASERS, HDR_KILLS, HDR_DEATHS, HDR_SHOTS, HDR_OVERFLOW, HDR_EPISODE, HDR_NEARTIES };
enum { SF_X = 0, SF_Y, SF_PX, SF_PY, SF_REWARD, SF_SCORE, SF_STEPS, SF_FLAGS };

struct ArenaLayout {
    int S, SP, L, W, H;
    int off_ship, off_lx, off_ly, off_ldx, off_ldy, off_lmeta;
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
    l.off_lx = l.off_ship + 8 * l.SP * 4;
    l.off_ly = l.off_lx + l.L * 8;
    l.off_ldx = l.off_ly + l.L * 8;
    l.off_ldy = l.off_ldx + l.L * 8;
    l.off_lmeta = l.off_ldy + l.L * 8;
    l.stride = (l.off_lmeta + l.L * 4 + 127) & ~127;
    l.r_kill = c.reward_kill;
    l.r_death = c.reward_death;
    l.r_aim = c.reward_aim;
    l.r_traj = c.reward_trajectory;
    return l;
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

// Scripted bots of agents/agent.py:99-155 (+ the stress distribution of SURVEY 8(d) config 4): the action row
// (shoot | thrust << 16, px | py << 16) of ship `ship` of global arena `arena` at frame `step`; (px, py) is the
// ship's current pointing, kept unless the bot re-points.  Shared by k_bot_actions and the fused step kernel.
__device__ __forceinline__ int2 bot_action(int kind, uint64_t seed, long long arena, int ship, uint32_t step, int px, int py,
                                           int W, int H) {
    uint32_t c[4] = {(uint32_t)arena, (uint32_t)ship, step, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    int shoot = 0, thrust = 0;
    const int rx = (int)mulhi32(c[1], (uint32_t)W + 1), ry = (int)mulhi32(c[2], (uint32_t)H + 1);
    switch (kind) {
    case OFB_BOT_RANDOM: {                                    // agents/agent.py:123-133
        const uint32_t k = mulhi32(c[0], 3u);
        shoot = k == 0; thrust = k == 1;
        if (k == 2) { px = rx; py = ry; }
    } break;
    case OFB_BOT_TURRET:                                      // agents/agent.py:136-144
        shoot = mulhi32(c[0], 10u) < 8;
        if (mulhi32(c[3], 10u) < 3) { px = rx; py = ry; }
        break;
    case OFB_BOT_RUNNER:                                      // agents/agent.py:147-155
        thrust = mulhi32(c[0], 10u) < 9;
        if (mulhi32(c[3], 10u) < 1) { px = rx; py = ry; }
        break;
    case OFB_BOT_THRUST: thrust = 1; break;                   // agents/agent.py:107-112
    case OFB_BOT_SHOOT: shoot = 1; break;                     // agents/agent.py:115-120
    case OFB_BOT_STRESS:                                      // agents/qlearnIA_V2.py:317-321, shoot forced
        shoot = 1; thrust = (int)(c[0] & 1u);
        px = (int)mulhi32(c[1], (uint32_t)W); py = (int)mulhi32(c[2], (uint32_t)H);
        break;
    default: break;                                           // idle: agents/agent.py:99-104
    }
    return make_int2((shoot & 0xffff) | (thrust << 16), (px & 0xffff) | (py << 16));
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

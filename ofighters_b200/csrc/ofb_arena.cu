// Handle management + the small arena kernels: create/reset (Battleground.__init__/restart),
// observation head, scripted bots, state export/import.  Compiled with -fmad=false.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include "ofb_common.cuh"

// ---------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";
void ofb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char *ofb_last_error(void) { return g_err; }
extern "C" int ofb_abi_version(void) { return OFB_ABI_VERSION; }

extern "C" void ofb_default_config(ofb_config *c) {
    memset(c, 0, sizeof(*c));
    c->n_ships = 7;            // lib/ofighters.py:53
    c->laser_cap = 0;
    c->width = 400;            // lib/observation.py:10-11
    c->height = 400;
    c->max_time = 200;         // lib/ofighters.py:59
    c->reward_kill = 0;        // agents/qlearnIA_V2.py:39-44
    c->reward_death = 0;
    c->reward_aim = 2;
    c->reward_trajectory = 1;
}

// ---------------------------------------------------------------- init / reset
// One thread per (arena, ship slot SP); thread 0 of the arena also writes the header.
__global__ void k_init(char *state, ArenaLayout lay, const int32_t *spawn, long long n_arenas) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long a = t / lay.SP;
    int i = (int)(t % lay.SP);
    if (a >= n_arenas) return;
    char *base = state + a * (long long)lay.stride;
    uint4 *ship = reinterpret_cast<uint4 *>(base + lay.off_ship);
    if (i == 0) {
        int4 *hdr = reinterpret_cast<int4 *>(base);
        hdr[0] = make_int4(0, 0, 0, 0);
        hdr[1] = make_int4(0, 0, 0, 0);
    }
    ShipRec r = {};
    if (i < lay.S) {
        r.x = spawn[(a * lay.S + i) * 2];
        r.y = spawn[(a * lay.S + i) * 2 + 1];
        r.px = r.x;                                          // pointing = own position (lib/ship.py:52)
        r.py = r.y;
        r.hull = 1;                                          // flying, hull 1 (lib/ship.py:45,55)
        r.alive = true;
    }
    ship[i] = ship_pack(r);
}

// Battleground.restart (lib/battleground.py:108-117) + Ship.reset (lib/ship.py:92-106) +
// Agent.reset (agents/agent.py:59-64).  One warp per arena; the episode's statistics are
// warp-reduced, then block-reduced in shared memory, then one atomicAdd per block and stat.
__global__ void __launch_bounds__(256)
k_reset(char *state, ArenaLayout lay, const uint8_t *mask, const int32_t *spawn,
        unsigned long long *stats, long long n_arenas) {
    __shared__ long long s_acc[6];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool ok = a < n_arenas && (!mask || mask[a]);
    long long acc_score = 0;
    int k = 0, d = 0, sh = 0;
    if (ok) {
        char *base = state + a * (long long)lay.stride;
        int *hdr = reinterpret_cast<int *>(base);
        uint4 *ship = reinterpret_cast<uint4 *>(base + lay.off_ship);
        if (lane < lay.S) {
            ShipRec r = ship_unpack(ship[lane]);
            acc_score = r.score;                             // scores.append(score)
            const int nx = spawn[(a * lay.S + lane) * 2], ny = spawn[(a * lay.S + lane) * 2 + 1];
            r.px = r.x;                                      // pointing = OLD position
            r.py = r.y;
            r.x = nx ? nx : r.x;                             // "x or self.body.x": 0 keeps old
            r.y = ny ? ny : r.y;
            r.score = 0;
            r.alive = true;                                  // flying again; hull NOT restored
            // pending reward survives the reset (agents/agent.py:59-64)
            ship[lane] = ship_pack(r);
        }
        k = hdr[HDR_KILLS]; d = hdr[HDR_DEATHS]; sh = hdr[HDR_SHOTS];
    }
    __syncwarp();
    if (ok && lane == 0) {
        int *hdr = reinterpret_cast<int *>(state + a * (long long)lay.stride);
        hdr[HDR_TIME] = 0; hdr[HDR_NLASERS] = 0; hdr[HDR_KILLS] = 0; hdr[HDR_DEATHS] = 0;
        hdr[HDR_SHOTS] = 0; hdr[HDR_EPISODE] += 1;
    }
    if (stats) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc_score += __shfl_xor_sync(0xffffffffu, acc_score, o);
        if (lane == 0 && ok) {
            atomicAdd((unsigned long long *)&s_acc[0], (unsigned long long)acc_score);
            atomicAdd((unsigned long long *)&s_acc[1], (unsigned long long)k);
            atomicAdd((unsigned long long *)&s_acc[2], (unsigned long long)d);
            atomicAdd((unsigned long long *)&s_acc[3], (unsigned long long)sh);
            atomicAdd((unsigned long long *)&s_acc[4], (unsigned long long)lay.S);
            atomicAdd((unsigned long long *)&s_acc[5], 1ull);
        }
        __syncthreads();
        if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0)
            atomicAdd(&stats[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
    }
}

// Read-only twin of k_reset's statistics: the RUNNING episode's [sum of score + pending reward, kills, deaths, shots, ships,
// arenas] added into stats[6] (what the reference's ScoreGraph / last_x_time_rewards consumers read mid-episode).
__global__ void __launch_bounds__(256)
k_stats(const char *state, ArenaLayout lay, unsigned long long *stats, long long n_arenas) {
    __shared__ long long s_acc[6];
    if (threadIdx.x < 6) s_acc[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool ok = a < n_arenas;
    long long acc_score = 0;
    int k = 0, d = 0, sh = 0;
    if (ok) {
        const char *base = state + a * (long long)lay.stride;
        const int *hdr = reinterpret_cast<const int *>(base);
        if (lane < lay.S) {
            const ShipRec r = ship_unpack(reinterpret_cast<const uint4 *>(base + lay.off_ship)[lane]);
            acc_score = (long long)r.score + r.reward;        // the pending reward folds into the score at the next Agent.step
        }
        k = hdr[HDR_KILLS]; d = hdr[HDR_DEATHS]; sh = hdr[HDR_SHOTS];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc_score += __shfl_xor_sync(0xffffffffu, acc_score, o);
    if (lane == 0 && ok) {
        atomicAdd((unsigned long long *)&s_acc[0], (unsigned long long)acc_score);
        atomicAdd((unsigned long long *)&s_acc[1], (unsigned long long)k);
        atomicAdd((unsigned long long *)&s_acc[2], (unsigned long long)d);
        atomicAdd((unsigned long long *)&s_acc[3], (unsigned long long)sh);
        atomicAdd((unsigned long long *)&s_acc[4], (unsigned long long)lay.S);
        atomicAdd((unsigned long long *)&s_acc[5], 1ull);
    }
    __syncthreads();
    if (threadIdx.x < 6 && s_acc[threadIdx.x] != 0) atomicAdd(&stats[threadIdx.x], (unsigned long long)s_acc[threadIdx.x]);
}

// ---------------------------------------------------------------- observation head
__global__ void k_obs_vec(const char *state, ArenaLayout lay, float4 *out, long long n_ships_total) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ships_total) return;
    long long a = t / lay.S;
    int i = (int)(t % lay.S);
    const ShipRec r = ship_unpack(reinterpret_cast<const uint4 *>(state + a * (long long)lay.stride + lay.off_ship)[i]);
    out[t * 2] = make_float4((float)r.reward, 1.0f, (float)r.px, (float)r.py);
    out[t * 2 + 1] = make_float4((float)lay.W, (float)lay.H, (float)r.x, (float)r.y);
}

// ---------------------------------------------------------------- scripted bots
__global__ void k_bot_actions(const char *state, ArenaLayout lay, int kind, const uint8_t *kinds, uint64_t seed,
                              long long arena0, uint32_t step, int2 *actions, long long n_ships_total) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ships_total) return;
    long long a = t / lay.S;
    int i = (int)(t % lay.S);
    if (kinds) kind = kinds[i];
    if (kind == OFB_BOT_EXTERNAL) return;                     // row is written by someone else (policy, host bot)
    const ShipRec r = ship_unpack(reinterpret_cast<const uint4 *>(state + a * (long long)lay.stride + lay.off_ship)[i]);
    actions[t] = bot_action(kind, seed, arena0 + a, i, step, r.px, r.py, lay.W, lay.H);
}

__global__ void k_random_spawn(int S, int W, int H, uint64_t seed, long long arena0, uint32_t episode, int2 *spawn,
                               long long n_ships_total) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_ships_total) return;
    uint32_t c[4] = {(uint32_t)(arena0 + t / S), (uint32_t)(t % S), episode, 1u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    spawn[t] = make_int2((int)mulhi32(c[0], (uint32_t)W + 1), (int)mulhi32(c[1], (uint32_t)H + 1));
}

// ---------------------------------------------------------------- export / import
template <bool IMPORT>
__global__ void k_xfer(char *state, ArenaLayout lay, ofb_state_view v, long long n_arenas) {
    const long long a = blockIdx.x;
    if (a >= n_arenas) return;
    char *base = state + a * (long long)lay.stride;
    int *hdr = reinterpret_cast<int *>(base);
    uint4 *ship = reinterpret_cast<uint4 *>(base + lay.off_ship);
    const int S = lay.S, L = lay.L;
    int32_t *hv[8] = {v.time, v.n_lasers, v.kills, v.deaths, v.shots, v.overflow, v.episode, v.near_ties};
    if (threadIdx.x < 8 && hv[threadIdx.x]) {
        if (IMPORT) hdr[threadIdx.x] = hv[threadIdx.x][a];
        else hv[threadIdx.x][a] = hdr[threadIdx.x];
    }
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const long long o = a * S + i;
        ShipRec r = ship_unpack(ship[i]);
        if (IMPORT) {
            if (v.ship_x) r.x = v.ship_x[o];
            if (v.ship_y) r.y = v.ship_y[o];
            if (v.ship_px) r.px = v.ship_px[o];
            if (v.ship_py) r.py = v.ship_py[o];
            if (v.ship_reward) r.reward = v.ship_reward[o];
            if (v.ship_score) r.score = v.ship_score[o];
            if (v.ship_hull) r.hull = v.ship_hull[o];
            if (v.ship_alive) r.alive = v.ship_alive[o] != 0;
            ship[i] = ship_pack(r);                          // ship_steps is derived (= time) and ignored on import
        } else {
            if (v.ship_x) v.ship_x[o] = r.x;
            if (v.ship_y) v.ship_y[o] = r.y;
            if (v.ship_px) v.ship_px[o] = r.px;
            if (v.ship_py) v.ship_py[o] = r.py;
            if (v.ship_reward) v.ship_reward[o] = r.reward;
            if (v.ship_score) v.ship_score[o] = r.score;
            if (v.ship_steps) v.ship_steps[o] = hdr[HDR_TIME];       // agent.steps == battleground.time (see ofb_common.cuh)
            if (v.ship_hull) v.ship_hull[o] = r.hull;
            if (v.ship_alive) v.ship_alive[o] = r.alive ? 1 : 0;
        }
    }
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        const long long o = a * L + k;
        double *lx = reinterpret_cast<double *>(base + laser_off(lay.off_laser, k));
        unsigned *lm = reinterpret_cast<unsigned *>(base + laser_meta_off(lay.off_laser, k));
        if (IMPORT) {
            if (v.laser_x) lx[0] = v.laser_x[o];
            if (v.laser_y) lx[OFB_G_Y / 8] = v.laser_y[o];
            if (v.laser_dx) lx[OFB_G_DX / 8] = v.laser_dx[o];
            if (v.laser_dy) lx[OFB_G_DY / 8] = v.laser_dy[o];
            if (v.laser_owner && v.laser_destroyed)
                *lm = (unsigned)v.laser_owner[o] | (v.laser_destroyed[o] ? 0x100u : 0u);
        } else {
            const bool live = k < hdr[HDR_NLASERS];
            const unsigned m = live ? *lm : 0u;
            if (v.laser_x) v.laser_x[o] = live ? lx[0] : 0.0;
            if (v.laser_y) v.laser_y[o] = live ? lx[OFB_G_Y / 8] : 0.0;
            if (v.laser_dx) v.laser_dx[o] = live ? lx[OFB_G_DX / 8] : 0.0;
            if (v.laser_dy) v.laser_dy[o] = live ? lx[OFB_G_DY / 8] : 0.0;
            if (v.laser_owner) v.laser_owner[o] = (uint8_t)(m & 0xffu);
            if (v.laser_destroyed) v.laser_destroyed[o] = (uint8_t)((m >> 8) & 1u);
        }
    }
}

// ---------------------------------------------------------------- C ABI
static inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }

extern "C" int ofb_create(const ofb_config *cfg, int64_t n_arenas, int device, const int32_t *spawn_dev,
                          void *stream, ofb_arenas **out) {
    if (!cfg || !out || !spawn_dev || n_arenas <= 0) { ofb_set_error("ofb_create: bad argument"); return OFB_E_ARG; }
    ofb_config c = *cfg;
    if (c.n_ships < 1 || c.n_ships > 32) { ofb_set_error("ofb_create: n_ships must be 1..32"); return OFB_E_ARG; }
    if (c.width < 32 || c.height < 1 || c.width > 16384 || c.height > 16384 || (c.width * c.height) % 128 != 0) {
        ofb_set_error("ofb_create: width*height must be a multiple of 128 and fit int16 coordinates");
        return OFB_E_ARG;
    }
    const int rw[4] = {c.reward_kill, c.reward_death, c.reward_aim, c.reward_trajectory};
    for (int i = 0; i < 4; i++)
        if (rw[i] < -500 || rw[i] > 500) { ofb_set_error("ofb_create: rewards must be within +-500 (pending reward is a 16-bit field)"); return OFB_E_ARG; }
    if (c.laser_cap <= 0) c.laser_cap = c.n_ships * 16 > 128 ? c.n_ships * 16 : 128;
    c.laser_cap = (c.laser_cap + 31) & ~31;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        ofb_set_error("ofb_create: no CUDA device (libofb has no CPU fallback)");
        return OFB_E_CUDA;
    }
    OFB_CUDA_CHECK(cudaSetDevice(device));
    ofb_arenas *h = new (std::nothrow) ofb_arenas();
    if (!h) return OFB_E_NOMEM;
    h->cfg = c;
    h->lay = make_layout(c);
    h->n_arenas = n_arenas;
    h->device = device;
    h->state = nullptr;
    cudaError_t e = cudaMalloc(&h->state, (size_t)n_arenas * h->lay.stride);
    if (e != cudaSuccess) {
        ofb_set_error("ofb_create: cudaMalloc(%zu) failed: %s", (size_t)n_arenas * h->lay.stride, cudaGetErrorString(e));
        delete h;
        return OFB_E_NOMEM;
    }
    h->stage_actions = nullptr;
    h->stage_obs = nullptr;
    h->pipe_ready = 0;
    h->host_seq = 0;
    const size_t n_ship = (size_t)n_arenas * c.n_ships;
    if (cudaMalloc(&h->stage_actions, n_ship * 4 * sizeof(int16_t)) != cudaSuccess ||
        cudaMalloc(&h->stage_obs, n_ship * 8 * sizeof(float)) != cudaSuccess) {
        ofb_set_error("ofb_create: cudaMalloc of the host-staging buffers failed");
        cudaFree(h->stage_actions);
        cudaFree(h->state);
        delete h;
        return OFB_E_NOMEM;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long nt = n_arenas * h->lay.SP;
    k_init<<<nblocks(nt, 256), 256, 0, st>>>(h->state, h->lay, spawn_dev, n_arenas);
    e = cudaGetLastError();
    if (e != cudaSuccess) {
        ofb_set_error("ofb_create: init kernel failed: %s", cudaGetErrorString(e));
        cudaFree(h->stage_actions);
        cudaFree(h->stage_obs);
        cudaFree(h->state);
        delete h;
        return OFB_E_CUDA;
    }
    *out = h;
    return OFB_OK;
}

// streams, events and double-buffered staging of the pipelined host frame; created on first use
int ofb_pipe_init(ofb_arenas *h) {
    if (h->pipe_ready) return OFB_OK;
    const size_t n_ship = (size_t)h->n_arenas * h->lay.S;
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    for (int i = 0; i < 2; i++) {
        OFB_CUDA_CHECK(cudaMalloc(&h->pipe_actions[i], n_ship * 4 * sizeof(int16_t)));
        OFB_CUDA_CHECK(cudaMalloc(&h->pipe_obs[i], n_ship * 8 * sizeof(float)));
        OFB_CUDA_CHECK(cudaMalloc(&h->pipe_obs16[i], n_ship * 5 * sizeof(int16_t)));
        OFB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_h2d[i], cudaEventDisableTiming));
        OFB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_step[i], cudaEventDisableTiming));
        OFB_CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_d2h[i], cudaEventDisableTiming));
    }
    OFB_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    OFB_CUDA_CHECK(cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    h->host_seq = 0;
    h->pipe_ready = 1;
    return OFB_OK;
}

extern "C" int ofb_destroy(ofb_arenas *h) {
    if (!h) return OFB_OK;
    cudaSetDevice(h->device);
    if (h->pipe_ready) {
        cudaStreamSynchronize(h->s_h2d);
        cudaStreamSynchronize(h->s_d2h);
        for (int i = 0; i < 2; i++) {
            cudaFree(h->pipe_actions[i]);
            cudaFree(h->pipe_obs[i]);
            cudaFree(h->pipe_obs16[i]);
            cudaEventDestroy(h->ev_h2d[i]);
            cudaEventDestroy(h->ev_step[i]);
            cudaEventDestroy(h->ev_d2h[i]);
        }
        cudaStreamDestroy(h->s_h2d);
        cudaStreamDestroy(h->s_d2h);
    }
    cudaFree(h->stage_actions);
    cudaFree(h->stage_obs);
    cudaFree(h->state);
    delete h;
    return OFB_OK;
}

extern "C" int ofb_laser_cap(const ofb_arenas *h) { return h ? h->lay.L : OFB_E_ARG; }
extern "C" int64_t ofb_state_stride(const ofb_arenas *h) { return h ? (int64_t)h->lay.stride : OFB_E_ARG; }

extern "C" int ofb_reset(ofb_arenas *h, const uint8_t *mask_dev, const int32_t *spawn_dev, int64_t *stats_dev,
                         void *stream) {
    if (!h || !spawn_dev) { ofb_set_error("ofb_reset: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    k_reset<<<nblocks(h->n_arenas * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        h->state, h->lay, mask_dev, spawn_dev, reinterpret_cast<unsigned long long *>(stats_dev), h->n_arenas);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_stats(const ofb_arenas *h, int64_t *stats_dev, void *stream) {
    if (!h || !stats_dev) { ofb_set_error("ofb_stats: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    if (h->n_arenas == 0) return OFB_OK;
    k_stats<<<nblocks(h->n_arenas * 32, 256), 256, 0, (cudaStream_t)stream>>>(h->state, h->lay,
                                                                             reinterpret_cast<unsigned long long *>(stats_dev), h->n_arenas);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_obs_vec(const ofb_arenas *h, float *out_dev, void *stream) {
    if (!h || !out_dev) { ofb_set_error("ofb_obs_vec: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    const long long nt = h->n_arenas * h->lay.S;
    k_obs_vec<<<nblocks(nt, 256), 256, 0, (cudaStream_t)stream>>>(h->state, h->lay, reinterpret_cast<float4 *>(out_dev), nt);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// The 5 entries of an observation head that are not constants (lib/observation.py:119-123: can_shoot = 1 and the map's
// dimensions never change), as int16: [reward, pointing.x, pointing.y, pos.x, pos.y] -- all small exact integers.
__global__ void k_obs_pack_i16(const float4 *__restrict__ obs, int16_t *__restrict__ out, long long n_rows) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const float4 a = obs[2 * r], b = obs[2 * r + 1];
    int16_t *o = out + 5 * r;
    o[0] = (int16_t)a.x; o[1] = (int16_t)a.z; o[2] = (int16_t)a.w; o[3] = (int16_t)b.z; o[4] = (int16_t)b.w;
}

extern "C" int ofb_obs_pack_i16(const float *obs_dev, int16_t *out_dev, int64_t n_rows, void *stream) {
    if (!obs_dev || !out_dev || n_rows < 0) { ofb_set_error("ofb_obs_pack_i16: bad argument"); return OFB_E_ARG; }
    if (n_rows == 0) return OFB_OK;
    k_obs_pack_i16<<<nblocks(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(obs_dev), out_dev, n_rows);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_bot_actions(const ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed,
                               int64_t arena0, uint32_t step, int16_t *actions_dev, void *stream) {
    if (!h || !actions_dev || bot_kind < 0 || (bot_kind > OFB_BOT_STRESS && bot_kind != OFB_BOT_EXTERNAL)) {
        ofb_set_error("ofb_bot_actions: bad argument");
        return OFB_E_ARG;
    }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    const long long nt = h->n_arenas * h->lay.S;
    k_bot_actions<<<nblocks(nt, 256), 256, 0, (cudaStream_t)stream>>>(h->state, h->lay, bot_kind, kinds_dev, seed, arena0, step,
                                                                       reinterpret_cast<int2 *>(actions_dev), nt);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_random_spawn(int64_t n_arenas, int n_ships, int width, int height, uint64_t seed, int64_t arena0,
                                uint32_t episode, int32_t *spawn_dev, void *stream) {
    if (!spawn_dev || n_arenas <= 0 || n_ships < 1) { ofb_set_error("ofb_random_spawn: bad argument"); return OFB_E_ARG; }
    const long long nt = n_arenas * n_ships;
    k_random_spawn<<<nblocks(nt, 256), 256, 0, (cudaStream_t)stream>>>(n_ships, width, height, seed, arena0, episode,
                                                                        reinterpret_cast<int2 *>(spawn_dev), nt);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_state_export(const ofb_arenas *h, const ofb_state_view *view, void *stream) {
    if (!h || !view) { ofb_set_error("ofb_state_export: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    k_xfer<false><<<(unsigned)h->n_arenas, 128, 0, (cudaStream_t)stream>>>(h->state, h->lay, *view, h->n_arenas);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_state_import(ofb_arenas *h, const ofb_state_view *view, void *stream) {
    if (!h || !view) { ofb_set_error("ofb_state_import: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    k_xfer<true><<<(unsigned)h->n_arenas, 128, 0, (cudaStream_t)stream>>>(h->state, h->lay, *view, h->n_arenas);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

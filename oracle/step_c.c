/* Batched C restatement of the Ofighters arena step -- TEST INFRASTRUCTURE, not product.
 *
 * Same rules as oracle/step_py.py, one loop nest per
 * reference function, for N independent arenas held in plain [N,...] arrays.
 * It is pinned by tests/test_oracle_c.py against the traces of the real
 * reference in tests/golden/.  Uses libm (sqrt, pow, atan2, atan, fmod) the
 * way CPython does, so on the same libm it is bit-identical to the reference.
 * Build: oracle/build.py (gcc -O2 -fno-builtin -ffp-contract=off -fopenmp).
 *
 * Reference lines restated (paths under /root/reference/ofighters):
 *   ofo_step        lib/battleground.py:146-166, lib/laser.py:36-62, lib/ship.py:127-230,303-339,
 *                   agents/agent.py:66-74, lib/form.py:24-31,42-44,75-83,148-188,298-307
 *   ofo_reset       lib/battleground.py:108-117, lib/ship.py:92-106, agents/agent.py:59-64
 *   ofo_obs_vec     lib/observation.py:101-123
 *   ofo_raster_bits lib/observation.py:79-95, lib/form.py:222-228 (+ skimage.draw.disk, see disk.py)
 *   ofo_bot_actions agents/agent.py:99-155, agents/qlearnIA_V2.py:317-321 (distributions only;
 *                   the RNG is Philox4x32-10, not CPython's MT19937 -- actions are inputs)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

typedef struct {
    int32_t n_arenas, n_ships, lcap, width, height;
    int32_t r_kill, r_death, r_aim, r_traj;
    int32_t *time, *n_lasers, *kills, *deaths, *shots, *overflow, *episode;           /* [N]   */
    int32_t *ship_x, *ship_y, *ship_px, *ship_py, *ship_hull, *ship_reward,
            *ship_score, *ship_steps;                                                   /* [N,S] */
    uint8_t *ship_alive;                                                                /* [N,S] */
    double  *laser_x, *laser_y;                                                         /* [N,L] */
    int32_t *laser_fx, *laser_fy, *laser_px, *laser_py;                                 /* [N,L] */
    uint8_t *laser_owner, *laser_destroyed;                                             /* [N,L] */
    /* episode statistics accumulated by ofo_reset: [sum score, kills, deaths, shots, ships, arenas] */
    int64_t *stats;
} ofo_state;

#define R_SHIP 8
#define R_LASER 2
#define SHIP_SPEED 8
#define LASER_SPEED 10

/* CPython float ** 2  ->  libm pow */
static inline double py_sq(double v) { return pow(v, 2.0); }

/* CPython float % float (Objects/floatobject.c float_rem) */
static inline double py_mod(double a, double b) {
    double m = fmod(a, b);
    if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}

static int on_trajectory(int x, int y, int px, int py, int ox, int oy) {
    const double two_pi = 2 * M_PI;
    double shooting = atan2((double)(py - y), (double)(px - x)) + M_PI;
    if (shooting == 0.0) return 0;
    double target = atan2((double)(oy - y), (double)(ox - x)) + M_PI;
    if (target == 0.0) return 0;
    long d2 = (long)(x - ox) * (x - ox) + (long)(y - oy) * (y - oy);
    double d = sqrt((double)d2);
    double ang = (d == 0.0) ? two_pi : atan((double)R_SHIP / d);
    double sup = py_mod(target + ang, two_pi);
    double inf = py_mod(target + -ang, two_pi);
    return inf <= shooting && shooting <= sup;
}

static void arena_step(ofo_state *st, int a, const int16_t *act) {
    const int S = st->n_ships, L = st->lcap, W = st->width, H = st->height;
    int32_t *sx = st->ship_x + (size_t)a * S, *sy = st->ship_y + (size_t)a * S;
    int32_t *spx = st->ship_px + (size_t)a * S, *spy = st->ship_py + (size_t)a * S;
    int32_t *hull = st->ship_hull + (size_t)a * S, *rew = st->ship_reward + (size_t)a * S;
    int32_t *score = st->ship_score + (size_t)a * S, *steps = st->ship_steps + (size_t)a * S;
    uint8_t *alive = st->ship_alive + (size_t)a * S;
    double *lx = st->laser_x + (size_t)a * L, *ly = st->laser_y + (size_t)a * L;
    int32_t *lfx = st->laser_fx + (size_t)a * L, *lfy = st->laser_fy + (size_t)a * L;
    int32_t *lpx = st->laser_px + (size_t)a * L, *lpy = st->laser_py + (size_t)a * L;
    uint8_t *lown = st->laser_owner + (size_t)a * L, *ldes = st->laser_destroyed + (size_t)a * L;
    int n = st->n_lasers[a];

    /* A7 prune (lib/ofighters.py:704-707), order preserved */
    int w = 0;
    for (int k = 0; k < n; k++) {
        if (ldes[k]) continue;
        if (w != k) {
            lx[w] = lx[k]; ly[w] = ly[k]; lfx[w] = lfx[k]; lfy[w] = lfy[k];
            lpx[w] = lpx[k]; lpy[w] = lpy[k]; lown[w] = lown[k]; ldes[w] = 0;
        }
        w++;
    }
    n = w;
    /* A1 score fold (agents/agent.py:66-74) */
    for (int i = 0; i < S; i++) { steps[i]++; score[i] += rew[i]; rew[i] = 0; }
    /* A2 */
    st->time[a]++;
    /* A3 lasers in list order (lib/laser.py:36-62) */
    for (int k = 0; k < n; k++) {
        long dX = lpx[k] - lfx[k], dY = lpy[k] - lfy[k];
        double dist = sqrt((double)(dX * dX + dY * dY));
        if (dist != 0.0) {
            lx[k] += (double)(dX * LASER_SPEED) / dist;
            ly[k] += (double)(dY * LASER_SPEED) / dist;
        }
        int explode = 0;
        for (int s = 0; s < S; s++) {
            if (alive[s] && sqrt(py_sq(lx[k] - sx[s]) + py_sq(ly[k] - sy[s])) <= (double)(R_LASER + R_SHIP)) {
                rew[lown[k]] += st->r_kill;
                st->kills[a]++;
                hull[s] -= 1;
                if (hull[s] <= 0) { rew[s] += st->r_death; st->deaths[a]++; alive[s] = 0; }
                explode = 1;
            }
        }
        if (explode || lx[k] < 0 || ly[k] < 0 || lx[k] >= W || ly[k] >= H) ldes[k] = 1;
    }
    /* A4 ships in index order (lib/ship.py:303-339) */
    for (int i = 0; i < S; i++) {
        if (!alive[i]) continue;
        const int16_t *ac = act + ((size_t)a * S + i) * 4;
        int shoot = ac[0] != 0, thrust = ac[1] != 0;
        spx[i] = ac[2]; spy[i] = ac[3];
        if (thrust) {
            long dX = spx[i] - sx[i], dY = spy[i] - sy[i];
            double dist = sqrt((double)(dX * dX + dY * dY));
            if (dist != 0.0) {
                double dx = (double)(dX * SHIP_SPEED) / dist, dy = (double)(dY * SHIP_SPEED) / dist;
                int nx = (int)((double)sx[i] + dx), ny = (int)((double)sy[i] + dy);
                nx = nx < 0 ? 0 : nx; ny = ny < 0 ? 0 : ny;
                sx[i] = nx > W - 1 ? W - 1 : nx;
                sy[i] = ny > H - 1 ? H - 1 : ny;
            }
        }
        if (shoot) {
            long dX = spx[i] - sx[i], dY = spy[i] - sy[i];
            long d2 = dX * dX + dY * dY;
            double dist = sqrt((double)d2);
            if (dist == 0.0) continue;
            int in_r = R_SHIP + R_LASER;
            int px0 = (int)((double)sx[i] + (double)(dX * in_r) / dist);
            int py0 = (int)((double)sy[i] + (double)(dY * in_r) / dist);
            int fx = px0, fy = py0;
            if (dist <= (double)(R_SHIP + R_LASER)) { fx = sx[i]; fy = sy[i]; }
            st->shots[a]++;
            if (n < L) {
                lx[n] = px0; ly[n] = py0; lfx[n] = fx; lfy[n] = fy; lpx[n] = spx[i]; lpy[n] = spy[i];
                lown[n] = (uint8_t)i; ldes[n] = 0; n++;
            } else st->overflow[a]++;
            int aimed = 0, traj = 0;
            for (int j = 0; j < S; j++) {
                if (j == i || !alive[j]) continue;
                long ex = sx[j] - spx[i], ey = sy[j] - spy[i];
                if (sqrt((double)(ex * ex + ey * ey)) <= (double)R_SHIP) aimed = 1;
                if (on_trajectory(sx[i], sy[i], spx[i], spy[i], sx[j], sy[j])) traj = 1;
            }
            if (aimed) rew[i] += st->r_aim;
            if (traj) rew[i] += st->r_traj;
        }
    }
    st->n_lasers[a] = n;
}

static void arena_obs(const ofo_state *st, int a, float *out) {
    const int S = st->n_ships;
    for (int i = 0; i < S; i++) {
        size_t o = (size_t)a * S + i;
        float *v = out + o * 8;
        v[0] = (float)st->ship_reward[o]; v[1] = 1.0f;
        v[2] = (float)st->ship_px[o]; v[3] = (float)st->ship_py[o];
        v[4] = (float)st->width; v[5] = (float)st->height;
        v[6] = (float)st->ship_x[o]; v[7] = (float)st->ship_y[o];
    }
}

void ofo_step(ofo_state *st, const int16_t *actions, float *obs_out) {
#pragma omp parallel for schedule(static)
    for (int a = 0; a < st->n_arenas; a++) {
        arena_step(st, a, actions);
        if (obs_out) arena_obs(st, a, obs_out);
    }
}

void ofo_obs_vec(const ofo_state *st, float *out) {
#pragma omp parallel for schedule(static)
    for (int a = 0; a < st->n_arenas; a++) arena_obs(st, a, out);
}

/* A8: mask NULL = all arenas; spawn [N,S,2] */
void ofo_reset(ofo_state *st, const uint8_t *mask, const int32_t *spawn) {
    const int S = st->n_ships;
    int64_t acc[6] = {0, 0, 0, 0, 0, 0};
    for (int a = 0; a < st->n_arenas; a++) {
        if (mask && !mask[a]) continue;
        acc[1] += st->kills[a]; acc[2] += st->deaths[a]; acc[3] += st->shots[a]; acc[5] += 1;
        st->time[a] = 0; st->n_lasers[a] = 0;
        st->kills[a] = st->deaths[a] = st->shots[a] = 0;
        st->episode[a]++;
        for (int i = 0; i < S; i++) {
            size_t o = (size_t)a * S + i;
            acc[0] += st->ship_score[o]; acc[4] += 1;
            st->ship_steps[o] = 0; st->ship_score[o] = 0;
            st->ship_px[o] = st->ship_x[o]; st->ship_py[o] = st->ship_y[o];
            int nx = spawn[o * 2], ny = spawn[o * 2 + 1];
            if (nx) st->ship_x[o] = nx;
            if (ny) st->ship_y[o] = ny;
            st->ship_alive[o] = 1;
        }
    }
    if (st->stats) for (int k = 0; k < 6; k++) st->stats[k] += acc[k];
}

/* A5: bit (y*W + x) of channel c, LSB-first in uint32 words; out [N,2,W*H/32] */
static void draw_disk(uint32_t *bits, double cy, double cx, double R, int H, int W) {
    long ulr = (long)ceil(cy - R), ulc = (long)ceil(cx - R);
    long lrr = (long)floor(cy + R), lrc = (long)floor(cx + R);
    if (ulr < 0) ulr = 0;
    if (ulc < 0) ulc = 0;
    if (lrr > H - 1) lrr = H - 1;
    if (lrc > W - 1) lrc = W - 1;
    double scr = cy - (double)ulr, scc = cx - (double)ulc;
    for (long i = 0; i <= lrr - ulr; i++) {
        double dr = ((double)i - scr) / R;
        for (long j = 0; j <= lrc - ulc; j++) {
            double dc = ((double)j - scc) / R;
            if (dr * dr + dc * dc < 1.0) {
                size_t b = (size_t)(ulr + i) * W + (size_t)(ulc + j);
                bits[b >> 5] |= 1u << (b & 31);
            }
        }
    }
}

void ofo_raster_bits(const ofo_state *st, uint32_t *out) {
    const int S = st->n_ships, L = st->lcap, W = st->width, H = st->height;
    const size_t words = (size_t)W * H / 32;
#pragma omp parallel for schedule(static)
    for (int a = 0; a < st->n_arenas; a++) {
        uint32_t *sm = out + (size_t)a * 2 * words, *lm = sm + words;
        memset(sm, 0, 2 * words * sizeof(uint32_t));
        for (int i = 0; i < S; i++) {
            size_t o = (size_t)a * S + i;
            if (st->ship_alive[o]) draw_disk(sm, st->ship_y[o], st->ship_x[o], R_SHIP, H, W);
        }
        for (int k = 0; k < st->n_lasers[a]; k++)
            draw_disk(lm, st->laser_y[(size_t)a * L + k], st->laser_x[(size_t)a * L + k], R_LASER, H, W);
    }
}

/* ---- Philox4x32-10 (Salmon et al. 2011), counter = (arena, ship, step, stream), key = seed ---- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* bot kinds: 0 idle, 1 random, 2 turret, 3 runner, 4 thrust, 5 shoot (agents/agent.py:99-155),
 *            6 stress = qlearnIA_V2.random_play with shoot forced */
void ofo_bot_actions(const ofo_state *st, int kind, uint64_t seed, int64_t arena0, uint32_t step,
                     int16_t *actions) {
    const int S = st->n_ships, W = st->width, H = st->height;
#pragma omp parallel for schedule(static)
    for (int a = 0; a < st->n_arenas; a++) {
        for (int i = 0; i < S; i++) {
            size_t o = (size_t)a * S + i;
            uint32_t c[4] = {(uint32_t)(arena0 + a), (uint32_t)i, step, 0u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            int shoot = 0, thrust = 0, px = st->ship_px[o], py = st->ship_py[o];
            int rx = (int)mulhi32(c[1], (uint32_t)W + 1), ry = (int)mulhi32(c[2], (uint32_t)H + 1);
            switch (kind) {
            case 1: { uint32_t k = mulhi32(c[0], 3u); shoot = k == 0; thrust = k == 1;
                      if (k == 2) { px = rx; py = ry; } } break;
            case 2: shoot = mulhi32(c[0], 10u) < 8; if (mulhi32(c[3], 10u) < 3) { px = rx; py = ry; } break;
            case 3: thrust = mulhi32(c[0], 10u) < 9; if (mulhi32(c[3], 10u) < 1) { px = rx; py = ry; } break;
            case 4: thrust = 1; break;
            case 5: shoot = 1; break;
            case 6: shoot = 1; thrust = (int)(c[0] & 1u);
                    px = (int)mulhi32(c[1], (uint32_t)W); py = (int)mulhi32(c[2], (uint32_t)H); break;
            default: break;
            }
            int16_t *ac = actions + o * 4;
            ac[0] = (int16_t)shoot; ac[1] = (int16_t)thrust; ac[2] = (int16_t)px; ac[3] = (int16_t)py;
        }
    }
}

/* spawn draws: U{0..W} x U{0..H} inclusive (lib/battleground.py:79-81,114), stream 1 */
void ofo_random_spawn(const ofo_state *st, uint64_t seed, int64_t arena0, uint32_t episode, int32_t *spawn) {
    const int S = st->n_ships, W = st->width, H = st->height;
    for (int a = 0; a < st->n_arenas; a++)
        for (int i = 0; i < S; i++) {
            uint32_t c[4] = {(uint32_t)(arena0 + a), (uint32_t)i, episode, 1u};
            philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
            size_t o = (size_t)a * S + i;
            spawn[o * 2] = (int32_t)mulhi32(c[0], (uint32_t)W + 1);
            spawn[o * 2 + 1] = (int32_t)mulhi32(c[1], (uint32_t)H + 1);
        }
}

int ofo_abi_version(void) { return 1; }

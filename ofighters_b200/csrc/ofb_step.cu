// K1 -- fused arena step for sm_100a (compiled with -fmad=false: every fp64 op is an IEEE rn op,
// which is what makes laser positions / hit tests / thrust truncation bit-identical to CPython).
//
// Restates Battleground.generate_frame + Agent.step's score fold + the Tk controller's laser
// pruning (reference: lib/battleground.py:153-160, lib/laser.py:36-62, lib/ship.py:127-230,303-339,
// agents/agent.py:66-74, lib/ofighters.py:704-707; rules A1-A4/A7 of SURVEY.md Appendix A).
//
// Mapping: one LPA-lane tile (LPA = 8/16/32 >= ships per arena) steps one arena, 32/LPA arenas per
// warp.  In the laser phase a lane owns a laser slot; in the ship phase lane i owns ship i.  The
// reference's sequential semantics are recovered with ballots:
//   * lasers are kept as a dense list in append order, so "lowest lane of the lowest chunk" ==
//     "first laser in list order": a ship alive at frame start dies to the first colliding laser,
//     and later lasers pass through it (laser.py:52-62);
//   * ship i's shoot-rewards read ships j<i after their thrust and j>i before it (ship.py:158-177
//     inside the loop of battleground.py:159-160).
#include <stdlib.h>
#include "ofb_common.cuh"
#include "ofb_raster_dev.cuh"

#define FULL 0xffffffffu

// largest double whose correctly rounded sqrt is <= 10.0 (collide: distance <= r_laser + r_ship)
#define D2_HIT_MAX 0x1.9000000000001p+6

// ---- enemy_on_trajectory (lib/ship.py:179-210) ------------------------------------------------
// Reference formula, evaluated with the device's fp64 libm.  Only reached for geometrically
// borderline pairs (see on_trajectory below); counted in hdr[HDR_NEARTIES].
__device__ __noinline__ bool on_trajectory_formula(int ux, int uy, int vx, int vy) {
    const double PI = 3.141592653589793, TWO_PI = 6.283185307179586;
    double shooting = atan2((double)uy, (double)ux) + PI;
    if (shooting == 0.0) return false;
    double target = atan2((double)vy, (double)vx) + PI;
    if (target == 0.0) return false;
    long long vv = (long long)vx * vx + (long long)vy * vy;
    double d = __dsqrt_rn((double)vv);
    double ang = (d == 0.0) ? TWO_PI : atan(__ddiv_rn(8.0, d));
    double s1 = target + ang, s2 = target - ang;
    double sup = fmod(s1, TWO_PI);              // s1 >= 0
    double inf = fmod(s2, TWO_PI);
    if (inf != 0.0) { if (inf < 0.0) inf += TWO_PI; } else inf = 0.0;   // CPython float_rem
    return inf <= shooting && shooting <= sup;
}

// Exact integer form of the same predicate away from its decision boundaries:
//   touch  <=>  angle(u, v) <= atan(8/|v|)   and the cone [t-a, t+a] does not straddle 0/2pi
// with u = pointing - me, v = enemy - me.  cos^2 of both sides are rationals of the integer
// coordinates, so the comparison is exact in int64; pairs within 1e-10 rad of a boundary (where
// the reference's outcome is decided by libm rounding) fall back to the formula.
__device__ __noinline__ bool on_trajectory(int ux, int uy, int vx, int vy, int &near_ties) {
    long long uu = (long long)ux * ux + (long long)uy * uy;
    long long vv = (long long)vx * vx + (long long)vy * vy;
    // enemy on my own pixel: target = pi, cone half-angle = 2*pi, sup = inf = pi exactly in fp64
    // (3*pi is representable) -> touch iff the shot angle is exactly pi, i.e. along +x.
    if (vv == 0) return uy == 0 && ux > 0;
    long long dot = (long long)ux * vx + (long long)uy * vy;
    if (dot <= 0) return false;                       // angle >= pi/2 > atan(8/d)
    long long diff = dot * dot * (vv + 64) - uu * vv * vv;
    double scale = (double)uu * (double)vv * (double)(vv + 64);
    bool near = fabs((double)diff) <= 1e-10 * scale;
    if (!near && diff < 0) return false;
    if (!near) {
        if (vx >= 0) return true;                     // cone axis >= pi/2 away from the 0/2pi seam
        long long lw = (long long)vx * vx * (vv + 64) - vv * vv;
        double sw = (double)vv * (double)(vv + 64);
        if (fabs((double)lw) > 1e-10 * sw) return lw <= 0;
    }
    near_ties++;
    return on_trajectory_formula(ux, uy, vx, vy);
}

struct BotSpec {
    int enabled, kind;
    const uint8_t *kinds;
    uint64_t seed;
    long long arena0;
    uint32_t step;
};

// ---- mbarrier / bulk-copy wrappers (TMA-class copies without a tensor map) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        if (++spins > (1u << 24)) __trap();              // a lost arrival must fail loudly, not hang the GPU
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#define STEP_THREADS 128
#define STEP_BAR_BYTES 128               // 4 warp mbarriers in front of the tiles

// Post-step entity list of one arena, kept in shared memory for the fused frame kernel's raster warps (k_frame):
//   +0   uint32 n            lasers in the list after the step (just-destroyed ones included, lib/observation.py:89-93)
//   +16  uint32 ship_xy[SP]  x | y << 16 | alive << 31
//   +16 + 4 * SP             double2 lxy[PC]   the first PC = min(L, 64) lasers; the raster reads the rare ones beyond
//                                              from the arena's block in HBM (written by the same CTA just before)
#define FR_POST_CAP 64
__host__ __device__ __forceinline__ int post_slot_bytes(int SP, int PC) { return 16 + 4 * SP + 16 * PC; }


// One LPA-lane tile per arena.  The arena's header, ships and the LIVE prefix of its laser list reach shared memory
// as bulk asynchronous copies counted on the warp's mbarrier:
//   phase 0        header + ships + the first C0 = LPA/4 laser groups (issued before anything is known about the arena)
//   phase 1        the remaining live groups of the first CH = LPA groups, issued as soon as n_lasers has landed and
//                  overlapped with the first two chunk iterations
//   phase p + 1    pass p >= 1 (more than 8 * LPA live lasers: rare), same buffer
// Everything the loop reads comes from the shared-memory snapshot; results go straight to HBM (only the fields that
// changed: x, y, meta of live lasers; dx, dy only for entries the compaction moved), so a tile never waits on a global
// load inside the loop and the in-place compaction cannot race with its own reads.
// phase 0 of a tile: header + ships + the first LPA/4 laser groups of `arena` -> tb, counted on mbar
template <int LPA>
__device__ __forceinline__ void step_tile_load0(char *__restrict__ state, const ArenaLayout &lay, const long long arena, const bool ok,
                                                unsigned char *tb, uint64_t *mbar) {
    if ((threadIdx.x & 31) % LPA == 0) {
        if (ok) {
            const unsigned bytes0 = (unsigned)(lay.off_laser + OFB_GROUP_BYTES * min(LPA / 4, lay.L >> 3));
            mbar_expect_tx(mbar, bytes0);
            bulk_g2s(tb, state + arena * (long long)lay.stride, bytes0, mbar);
        } else mbar_arrive(mbar);
    }
}

template <int LPA, int PIT, bool POST>
__device__ __forceinline__ void step_tile(char *__restrict__ state, const ArenaLayout &lay, const int2 *__restrict__ actions,
                                          float4 *__restrict__ obs_out, const long long arena, const bool ok, const BotSpec &bots,
                                          unsigned char *tb, uint64_t *mbar, const uint32_t ph, uint32_t &ph_used,
                                          unsigned char *post, const int post_cap) {
    constexpr int CH = PIT * LPA / 8;                    // laser groups staged per pass = PIT chunk iterations
    constexpr int C0 = LPA / 4;                          // groups copied with the header = 2 chunk iterations
    static_assert(PIT > 2, "the second copy is awaited before chunk iteration 2");
    constexpr unsigned GM = (LPA == 32) ? 0xffffffffu : ((1u << LPA) - 1u);
    const int lane = threadIdx.x & 31;
    const int g = lane / LPA, gl = lane % LPA;
    const unsigned gshift = g * LPA;
    const unsigned tmask = GM << gshift;                 // lanes of this arena's tile
    const bool leader = gl == 0;
    const int S = lay.S, L = lay.L, off_l = lay.off_laser;
    char *base = state + (ok ? arena : 0) * (long long)lay.stride;
    const int G_cap = L >> 3;
    double2 *post_lxy = reinterpret_cast<double2 *>(post + 16 + 4 * lay.SP);

    if (!POST) step_tile_load0<LPA>(state, lay, arena, ok, tb, mbar);      // k_frame issues it one unit ahead

    // ---- while the state is in flight: the action row, or the scripted bot's random words (they need nothing from the
    //      arena; the kind byte is loaded first and interpreted after the Philox rounds)
    const bool is_ship = ok && gl < S;
    int2 act = make_int2(0, 0);
    BotDraw draw = {};
    int kind = OFB_BOT_EXTERNAL;
    if (is_ship) {
        if (bots.enabled) {
            kind = bots.kinds ? (int)__ldg(bots.kinds + gl) : bots.kind;
            const uint4 rnd = bot_random(bots.seed, bots.arena0 + arena, gl, bots.step);
            if (kind != OFB_BOT_EXTERNAL) draw = bot_interpret(kind, rnd, lay.W, lay.H);
        }
        if (kind == OFB_BOT_EXTERNAL) act = actions[arena * S + gl];
    }

    mbar_wait(mbar, ph & 1u);
    int n = ok ? reinterpret_cast<const int *>(tb)[HDR_NLASERS] : 0;     // the other header words are re-read at the end
    int dk = 0;                                          // kills (= deaths: hull is 1 and never restored) of this frame
    const int G_live = min((n + 7) >> 3, G_cap);
    if (leader) {
        const int hi = min(G_live, CH);
        if (ok && hi > C0) {
            const unsigned bytes1 = (unsigned)(OFB_GROUP_BYTES * (hi - C0));
            mbar_expect_tx(mbar, bytes1);
            bulk_g2s(tb + off_l + OFB_GROUP_BYTES * C0, base + off_l + OFB_GROUP_BYTES * C0, bytes1, mbar);
        } else mbar_arrive(mbar);
    }
    // the ship record stays in the tile; the loop only carries what it can change (alive, pending reward)
    const uint4 *srec = reinterpret_cast<const uint4 *>(tb + lay.off_ship);
    bool alive = false;
    int rew = 0;                                         // reward earned this frame (the old pending reward folds into the score)
    {
        ShipRec me = {};
        if (is_ship) me = ship_unpack(srec[gl]);
        alive = me.alive;
        if (is_ship && kind != OFB_BOT_EXTERNAL) act = bot_apply(draw, me.px, me.py);
    }
    int near_ties = 0;

    // centres of the ships that are alive at frame start, compacted, as floats for the box pre-filter; kept with their ship
    // indices in the (now consumed) ship region of the tile
    float2 *spos = reinterpret_cast<float2 *>(tb + off_l + CH * OFB_GROUP_BYTES);      // scratch behind the staged groups
    unsigned char *sid = reinterpret_cast<unsigned char *>(spos) + 8 * lay.SP;
    const unsigned alive_bits = (__ballot_sync(FULL, alive) >> gshift) & GM;
    const int n_alive = __popc(alive_bits);
    unsigned alive_c = n_alive >= 32 ? 0xffffffffu : ((1u << n_alive) - 1u);      // by compact index
    const int na_max = __reduce_max_sync(FULL, n_alive);
    __syncwarp();
    if (alive) {
        const int r = __popc(alive_bits & ((1u << gl) - 1u));
        const unsigned xy = srec[gl].x;
        spos[r] = make_float2((float)(xy & 0xffffu), (float)((xy >> 16) & 0x7fffu));
        sid[r] = (unsigned char)gl;
    }
    __syncwarp();

    // ---- A3 + A7: lasers in list order; entries destroyed last frame are dropped on load ----
    // A ship alive at frame start dies to the FIRST laser (list order) that collides with it, and only
    // that laser explodes on it (lib/laser.py:52-62).  Collisions are rare, so each lane first builds
    // the bit mask of ships its laser touches -- float box pre-filter, then the exact fp64 distance the
    // reference computes -- and the ballots that recover the list order run only when a mask is non-zero.
    int iters = (n + LPA - 1) / LPA;
    iters = __reduce_max_sync(FULL, iters);
    ph_used = 2u + (uint32_t)max(0, (iters + PIT - 1) / PIT - 1);      // mbarrier phases this call completes (warp-uniform)
    int w = 0;                                           // compaction write cursor
    for (int it = 0; it < iters; it++) {
        const int pit = it % PIT;                        // chunk iteration within the pass
        if (it == 2) mbar_wait(mbar, (ph + 1u) & 1u);
        else if (pit == 0 && it > 0) {                   // pass p: restage groups [p * CH, (p + 1) * CH)
            const int p = it / PIT;
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (leader) {
                const int lo = p * CH, hi = min(G_live, lo + CH);
                if (ok && hi > lo) {
                    const unsigned bytes = (unsigned)(OFB_GROUP_BYTES * (hi - lo));
                    mbar_expect_tx(mbar, bytes);
                    bulk_g2s(tb + off_l, base + off_l + OFB_GROUP_BYTES * lo, bytes, mbar);
                } else mbar_arrive(mbar);
            }
            mbar_wait(mbar, (ph + (uint32_t)p + 1u) & 1u);
        }
        const int k = it * LPA + gl;
        const unsigned char *lp = tb + off_l + ((pit * LPA + gl) >> 3) * OFB_GROUP_BYTES + (gl & 7) * 8;
        const unsigned meta = *reinterpret_cast<const unsigned *>(lp + OFB_G_META - (gl & 7) * 4);
        double x = *reinterpret_cast<const double *>(lp), y = *reinterpret_cast<const double *>(lp + OFB_G_Y);
        const double dx = *reinterpret_cast<const double *>(lp + OFB_G_DX), dy = *reinterpret_cast<const double *>(lp + OFB_G_DY);
        const bool live = ok && k < n && !(meta & 0x100u);
        if (live) {
            x = __dadd_rn(x, dx);                        // lib/laser.py:46-47
            y = __dadd_rn(y, dy);
        }
        const float fx = (float)x, fy = (float)y;
        unsigned hm = 0;
        for (int j = 0; j < na_max; j++) {               // entries beyond this tile's n_alive are masked by alive_c
            const float2 c = spos[j];
            const float m = fmaxf(fabsf(fx - c.x), fabsf(fy - c.y));
            hm |= (m <= 10.5f ? 1u : 0u) << j;           // superset of the radius-10 disk
        }
        hm = live ? (hm & alive_c) : 0u;
        bool hit_any = false;
        if (__ballot_sync(FULL, hm != 0u) & tmask) {     // some laser of this tile's chunk may touch a ship (rare)
            unsigned cand = __reduce_or_sync(tmask, hm);
            while (cand) {
                const int j = __ffs(cand) - 1;
                cand &= cand - 1;
                const int s = sid[j];
                const unsigned sxy = srec[s].x;
                const double ddx = __dsub_rn(x, (double)(int)(sxy & 0xffffu));
                const double ddy = __dsub_rn(y, (double)(int)((sxy >> 16) & 0x7fffu));
                const double d2 = __dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy));
                const bool hit = ((hm >> j) & 1u) && d2 <= D2_HIT_MAX;
                const unsigned b = (__ballot_sync(tmask, hit) >> gshift) & GM;
                if (!b) continue;
                const int killer = __ffs(b) - 1;
                if (lay.r_kill != 0) {
                    const int own = __shfl_sync(tmask, (int)(meta & 0xffu), gshift + killer);
                    if (gl == own) rew += lay.r_kill;    // lib/laser.py:57
                }
                if (gl == killer) hit_any = true;
                alive_c &= ~(1u << j);
                dk += 1;                                 // events (1, t) and (10, t)
                if (gl == s) {                           // lib/ship.py:127-131,225-230: hull -= 1 <= 0 -> destroyed
                    alive = false;
                    rew += lay.r_death;
                }
            }
        }
        const bool destroyed = live && (hit_any || x < 0.0 || y < 0.0 || x >= (double)lay.W || y >= (double)lay.H);
        const unsigned lv = (__ballot_sync(FULL, live) >> gshift) & GM;
        const int pos = w + __popc(lv & ((1u << gl) - 1u));
        if (live) {
            char *gp = base + laser_off(off_l, pos);
            *reinterpret_cast<double *>(gp) = x;
            *reinterpret_cast<double *>(gp + OFB_G_Y) = y;
            *reinterpret_cast<unsigned *>(base + laser_meta_off(off_l, pos)) = (meta & 0xffu) | (destroyed ? 0x100u : 0u);
            if (POST && pos < post_cap) post_lxy[pos] = make_double2(x, y);
            if (pos != k) {
                *reinterpret_cast<double *>(gp + OFB_G_DX) = dx;
                *reinterpret_cast<double *>(gp + OFB_G_DY) = dy;
            }
        }
        w += __popc(lv);
    }

    // ---- the ship record again (A1: score fold, dead ships included -- agents/agent.py:66-74, lib/ship.py:260-262) ----
    ShipRec me = {};
    if (is_ship) me = ship_unpack(srec[gl]);
    int sx = me.x, sy = me.y, spx = me.px, spy = me.py;
    const int score = me.score + me.reward;
    const int hull = me.hull - ((me.alive && !alive) ? 1 : 0);

    // ---- A4: ships in index order (lib/ship.py:303-339) ----
    const int old_x = sx, old_y = sy;
    bool shooter = false;
    double nlx = 0.0, nly = 0.0, ndx = 0.0, ndy = 0.0;
    if (is_ship && alive) {
        const int a_shoot = (short)(act.x & 0xffff), a_thrust = (short)(act.x >> 16);
        spx = (short)(act.y & 0xffff);
        spy = (short)(act.y >> 16);
        if (a_thrust) {                                  // lib/ship.py:213-222
            const int dX = spx - sx, dY = spy - sy;
            const int d2 = dX * dX + dY * dY;
            if (d2 != 0) {
                const double dist = __dsqrt_rn((double)d2);
                const double mx = __ddiv_rn((double)(dX * OFB_SHIP_SPEED), dist);
                const double my = __ddiv_rn((double)(dY * OFB_SHIP_SPEED), dist);
                int nx = (int)__dadd_rn((double)sx, mx);
                int ny = (int)__dadd_rn((double)sy, my);
                sx = min(lay.W - 1, max(0, nx));
                sy = min(lay.H - 1, max(0, ny));
            }
        }
        if (a_shoot) {                                   // lib/ship.py:134-156, lib/form.py:159-188
            const int dX = spx - sx, dY = spy - sy;
            const int d2 = dX * dX + dY * dY;
            if (d2 != 0) {
                const double dist = __dsqrt_rn((double)d2);
                const int in_r = OFB_R_SHIP + OFB_R_LASER;
                const int px0 = (int)__dadd_rn((double)sx, __ddiv_rn((double)(dX * in_r), dist));
                const int py0 = (int)__dadd_rn((double)sy, __ddiv_rn((double)(dY * in_r), dist));
                int fx = px0, fy = py0;
                if (d2 <= in_r * in_r) { fx = sx; fy = sy; }       // lib/ship.py:147-148
                const int fX = spx - fx, fY = spy - fy;            // lib/laser.py:39-45
                const int f2 = fX * fX + fY * fY;
                if (f2 != 0) {
                    const double fd = __dsqrt_rn((double)f2);
                    ndx = __ddiv_rn((double)(fX * OFB_LASER_SPEED), fd);
                    ndy = __ddiv_rn((double)(fY * OFB_LASER_SPEED), fd);
                }
                nlx = (double)px0;
                nly = (double)py0;
                shooter = true;
            }
        }
    }
    // shoot rewards: enemies j<i are seen after their move, j>i before it
    int n_shots = 0, n_over = 0;
    {
        const unsigned alive_now = (__ballot_sync(FULL, alive) >> gshift) & GM;
        const unsigned any_shooter = __ballot_sync(FULL, shooter) & tmask;
        bool aimed = false, traj = false;
        if (any_shooter) {
            for (int j = 0; j < S; j++) {
                const int jnx = __shfl_sync(tmask, sx, gshift + j), jny = __shfl_sync(tmask, sy, gshift + j);
                const int jox = __shfl_sync(tmask, old_x, gshift + j), joy = __shfl_sync(tmask, old_y, gshift + j);
                if (shooter && j != gl && ((alive_now >> j) & 1u)) {
                    const int ox = j < gl ? jnx : jox, oy = j < gl ? jny : joy;
                    const int ax = ox - spx, ay = oy - spy;
                    if (ax * ax + ay * ay <= OFB_R_SHIP * OFB_R_SHIP) aimed = true;     // lib/ship.py:165-169
                    if (!traj && on_trajectory(spx - sx, spy - sy, ox - sx, oy - sy, near_ties)) traj = true;
                }
            }
        }
        if (aimed) rew += lay.r_aim;
        if (traj) rew += lay.r_traj;
        // append the new lasers in ship order (lib/ship.py:151)
        const unsigned sh = any_shooter >> gshift;
        const int slot = w + __popc(sh & ((1u << gl) - 1u));
        if (shooter && slot < L) {
            char *gp = base + laser_off(off_l, slot);
            *reinterpret_cast<double *>(gp) = nlx;
            *reinterpret_cast<double *>(gp + OFB_G_Y) = nly;
            *reinterpret_cast<double *>(gp + OFB_G_DX) = ndx;
            *reinterpret_cast<double *>(gp + OFB_G_DY) = ndy;
            *reinterpret_cast<unsigned *>(base + laser_meta_off(off_l, slot)) = (unsigned)gl;
            if (POST && slot < post_cap) post_lxy[slot] = make_double2(nlx, nly);
        }
        const int want = w + __popc(sh);
        n_shots = __popc(sh);
        n_over = max(0, want - L);
        n = min(want, L);
    }
    int nt = __reduce_add_sync(tmask, near_ties);

    // ---- store ----
    if (is_ship) {
        ShipRec o;
        o.x = sx; o.y = sy; o.px = spx; o.py = spy; o.score = score; o.reward = rew; o.hull = hull; o.alive = alive;
        const uint4 packed = ship_pack(o);
        reinterpret_cast<uint4 *>(base + lay.off_ship)[gl] = packed;
        if (POST) reinterpret_cast<unsigned *>(post + 16)[gl] = packed.x;
        if (obs_out) {                                   // lib/observation.py:113-123
            float4 *ob = obs_out + (arena * S + gl) * 2;
            ob[0] = make_float4((float)rew, 1.0f, (float)spx, (float)spy);
            ob[1] = make_float4((float)lay.W, (float)lay.H, (float)sx, (float)sy);
        }
    }
    if (ok && leader) {                                  // A2: time += 1
        const int4 h0 = reinterpret_cast<const int4 *>(tb)[0], h1 = reinterpret_cast<const int4 *>(tb)[1];
        int4 *h4 = reinterpret_cast<int4 *>(base);
        h4[0] = make_int4(h0.x + 1, n, h0.z + dk, h0.w + dk);
        h4[1] = make_int4(h1.x + n_shots, h1.y + n_over, h1.z, h1.w + nt);
        if (POST) *reinterpret_cast<unsigned *>(post) = (unsigned)n;
    }
}

template <int LPA, int PIT, int MINB>
__global__ void __launch_bounds__(STEP_THREADS, MINB)
k_step(char *__restrict__ state, const ArenaLayout lay, const int2 *__restrict__ actions,
       float4 *__restrict__ obs_out, long long n_arenas, const BotSpec bots, const int tile_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int APW = 32 / LPA;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / LPA;
    const long long warp_global = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long arena = warp_global * APW + g;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem) + warp;
    unsigned char *tb = smem + STEP_BAR_BYTES + (warp * APW + g) * tile_bytes;
    if (lane == 0) {
        mbar_init(mbar, APW);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t ph_used;
    step_tile<LPA, PIT, false>(state, lay, actions, obs_out, arena, arena < n_arenas, bots, tb, mbar, 0u, ph_used, nullptr, 0);
}


// ---- K1 + K2 fused: Battleground.frame (lib/battleground.py:163-166) = generate_frame + the new Observation ----------------
// Persistent, warp-specialised, one CTA per SM over a contiguous range of arenas:
//   * SW stepper warps run step_tile over the range (unit u = 32/LPA arenas -> warp u % SW) and leave every arena's
//     post-step entity list in a ring of K shared-memory slots (full / empty mbarriers per slot);
//   * NG raster groups of 8 warps consume the ring (group q takes arenas i = q mod NG, in order): zero one of the
//     group's NBUF bitmap buffers, compose both disk maps from the slot (one thread per ship row / per laser pixel),
//     hand the 2 * W*H/8 bytes to one bulk store, and move on while that store drains.
// The step's latency chain (two dependent bulk loads + the fp64 loop) therefore hides behind the raster's HBM writes,
// the arena state is read once per frame, and a frame is one launch.
#define FR_RASTER_WARPS 8                // per raster group
#define FR_BAR_BYTES 1024                // tile mbarriers [SW <= 32] at +0, full[K] at +256, empty[K] behind them (K <= 48)
#define FR_MAX_K 48

template <int LPA, int PIT, int SW, int NG, int NBUF>
__global__ void __launch_bounds__((SW + NG * FR_RASTER_WARPS) * 32, 1)
k_frame(char *__restrict__ state, const ArenaLayout lay, const int2 *__restrict__ actions, float4 *__restrict__ obs_out,
        long long n_arenas, const BotSpec bots, const int tile_bytes, const int slot_bytes, const int post_cap, const int K,
        uint32_t *__restrict__ maps_out, long long *__restrict__ prof, const int dbg) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int APW = 32 / LPA;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *tile_bar = reinterpret_cast<uint64_t *>(smem);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 256), *empty = full + K;
    unsigned char *tiles = smem + FR_BAR_BYTES;
    unsigned char *slots = tiles + SW * APW * tile_bytes;
    const int words = (lay.W * lay.H) >> 5;
    const unsigned map_bytes = (unsigned)(2 * words * 4);
    unsigned char *bitmaps = smem + ((FR_BAR_BYTES + SW * APW * tile_bytes + K * slot_bytes + 127) & ~127);
    const long long a0 = n_arenas * blockIdx.x / gridDim.x, a1 = n_arenas * (blockIdx.x + 1) / gridDim.x;
    const int cnt = (int)(a1 - a0);

    if ((int)threadIdx.x < SW) mbar_init(&tile_bar[threadIdx.x], APW);
    else if ((int)threadIdx.x - SW < 2 * K) mbar_init(&full[threadIdx.x - SW], 1);      // full[K] and empty[K] are contiguous
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    if (warp < SW) {
        // ---------------- stepper warps
        const int g = lane / LPA, gl = lane % LPA;
        uint32_t ph = 0;
        long long t_wait = 0, t_begin = clock64(), t_unit_max = 0;
        int n_units = 0;
        unsigned char *tb = tiles + (warp * APW + g) * tile_bytes;
        for (int u = warp; u * APW < cnt; u += SW) {
            const int i = u * APW + g;
            const bool ok = i < cnt;
            const int slot = i % K, r = i / K;
            step_tile_load0<LPA>(state, lay, a0 + i, ok, tb, &tile_bar[warp]);
            const long long tw = clock64();
            if (ok && r > 0) mbar_wait(&empty[slot], (uint32_t)((r - 1) & 1));      // the raster has left arena i - K
            __syncwarp();
            const long long tu = clock64();
            t_wait += tu - tw;
            uint32_t used;
            step_tile<LPA, PIT, true>(state, lay, actions, obs_out, a0 + i, ok, bots, tb, &tile_bar[warp], ph, used,
                                      slots + slot * slot_bytes, post_cap);
            ph += used;
            t_unit_max = max(t_unit_max, clock64() - tu);
            n_units++;
            __syncwarp();
            if (ok && gl == 0) mbar_arrive(&full[slot]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");            // tile reads before the next unit's bulk copy
        }
        if (prof && lane == 0) {                         // debug: cycles this stepper warp spent in total / waiting for a slot
            prof[((long long)blockIdx.x * 32 + warp) * 8 + 0] = clock64() - t_begin;
            prof[((long long)blockIdx.x * 32 + warp) * 8 + 1] = t_wait;
            prof[((long long)blockIdx.x * 32 + warp) * 8 + 2] = t_unit_max;
            prof[((long long)blockIdx.x * 32 + warp) * 8 + 3] = n_units;
        }
        return;
    }

    // ---------------- raster groups
    constexpr int RT = FR_RASTER_WARPS * 32;
    const int q = (warp - SW) / FR_RASTER_WARPS;
    const int rt = threadIdx.x - (SW + q * FR_RASTER_WARPS) * 32;
    const int W = lay.W, H = lay.H, S = lay.S;
    const int rows = 2 * OFB_R_SHIP - 1;
    unsigned char *gbuf = bitmaps + (size_t)q * NBUF * map_bytes;
    for (int x = rt; x < NBUF * (int)(map_bytes / 16); x += RT) reinterpret_cast<uint4 *>(gbuf)[x] = make_uint4(0, 0, 0, 0);
    // this thread's first ship-row and laser-row items are the same for every arena
    const int s_i = rt / rows, s_dr = rt % rows - (OFB_R_SHIP - 1);
    const int l_t = RT - 1 - rt, l_k = l_t / 5, l_i = l_t % 5;       // laser rows from the top thread down
    // Re-zeroing: a buffer comes back after NBUF arenas of this group.  Each thread remembers the (at most two + two) words
    // its items of that arena touched and clears just those; an arena with more items than threads asks for a full wipe.
    int zs[NBUF], zl[NBUF];
    bool zfull[NBUF];
#pragma unroll
    for (int j = 0; j < NBUF; j++) { zs[j] = -1; zl[j] = -1; zfull[j] = false; }
    int use = 0;                                         // arenas this group has rasterised
    long long pt[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
#define FR_PROF(j) do { if (prof) { const long long now = clock64(); pt[j] += now - tp; tp = now; } } while (0)
    for (int i = q; i < cnt; i += NG, use++) {
        long long tp = prof ? clock64() : 0;
        const int slot = i % K, r = i / K;
        uint32_t *bits = reinterpret_cast<uint32_t *>(gbuf + (size_t)(use % NBUF) * map_bytes);
        if (rt == 0 && use >= NBUF)                      // the store that last used this buffer has read it out
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(NBUF - 1) : "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(RT) : "memory");
        FR_PROF(0);
        if (zfull[0]) {
            uint4 *b4 = reinterpret_cast<uint4 *>(bits);
            for (int x = rt; x < (int)(map_bytes / 16); x += RT) b4[x] = make_uint4(0, 0, 0, 0);
        } else {
            if (zs[0] >= 0) { bits[zs[0]] = 0u; bits[zs[0] + 1] = 0u; }            // (word + 1 may belong to the laser map: also zero)
            if (zl[0] >= 0) { bits[zl[0]] = 0u; if (zl[0] + 1 < 2 * words) bits[zl[0] + 1] = 0u; }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(RT) : "memory");
        FR_PROF(1);
        mbar_wait(&full[slot], (uint32_t)(r & 1));
        FR_PROF(2);
        const unsigned char *post = slots + slot * slot_bytes;
        const int n = (int)*reinterpret_cast<const unsigned *>(post);
        const unsigned *sxy = reinterpret_cast<const unsigned *>(post + 16);
        const double2 *lxy = reinterpret_cast<const double2 *>(post + 16 + 4 * lay.SP);
        int ws = -1, wl = -1;
        if (l_t < n * 5 && !(dbg & 1)) {
            double2 c;
            if (l_k < post_cap) c = lxy[l_k];
            else {                                       // beyond the slot: the stepper's own store to the arena block
                const double *lp = reinterpret_cast<const double *>(state + (a0 + i) * (long long)lay.stride + laser_off(lay.off_laser, l_k));
                c = make_double2(lp[0], lp[OFB_G_Y / 8]);
            }
            wl = raster_laser_row(bits + words, W, H, c.x, c.y, l_i);
            if (wl >= 0) wl += words;
        }
        if (rt < S * rows && !(dbg & 2)) ws = raster_ship_row(bits, W, H, sxy[s_i], s_dr);
        const bool many = n * 5 > RT || S * rows > RT;   // (group-uniform) more items than threads: generic loops, full wipe later
        if (many) {
            for (int t = l_t + RT; t < n * 5; t += RT) {
                const int k = t / 5;
                double2 c;
                if (k < post_cap) c = lxy[k];
                else {
                    const double *lp = reinterpret_cast<const double *>(state + (a0 + i) * (long long)lay.stride + laser_off(lay.off_laser, k));
                    c = make_double2(lp[0], lp[OFB_G_Y / 8]);
                }
                raster_laser_row(bits + words, W, H, c.x, c.y, t % 5);
            }
            for (int t = rt + RT; t < S * rows; t += RT) raster_ship_row(bits, W, H, sxy[t / rows], t % rows - (OFB_R_SHIP - 1));
        }
#pragma unroll
        for (int j = 0; j + 1 < NBUF; j++) { zs[j] = zs[j + 1]; zl[j] = zl[j + 1]; zfull[j] = zfull[j + 1]; }
        zs[NBUF - 1] = ws; zl[NBUF - 1] = wl; zfull[NBUF - 1] = many;
        FR_PROF(3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(RT) : "memory");
        FR_PROF(4);
        if (rt == 0) {
            char *dst = reinterpret_cast<char *>(maps_out) + (a0 + i) * (long long)map_bytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(bits)), "r"(map_bytes)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            mbar_arrive(&empty[slot]);
        }
    }
    if (rt == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (prof && rt == 0) {                               // debug: cycles of this group's leader per phase of the loop
        long long *o = prof + ((long long)blockIdx.x * 32 + SW + q) * 8;
        o[0] = clock64() - t_begin;
        for (int j = 0; j < 5; j++) o[1 + j] = pt[j];
        o[6] = use;
    }
#undef FR_PROF
}

// lanes per arena: the smallest tile that holds the ships, widened while the batch is too small to fill the GPU
// (more lanes = more lasers per pass and more warps in flight)
static inline int lpa_for(int S, long long n_arenas) {
    int lpa = S <= 8 ? 8 : (S <= 16 ? 16 : 32);
    while (lpa < 32 && n_arenas * lpa / 32 < 148 * 12) lpa *= 2;
    return lpa;
}

// shared-memory image of one arena tile: the arena block's prefix (header, ships, PIT * lpa / 8 laser groups) followed
// by the scratch of the compacted ship centres (8 + 1 bytes per ship), sized so that consecutive tiles start 64 B apart
// modulo the 128-byte bank period (two tiles' 64-byte group rows then never share a bank)
static inline int step_tile_bytes(const ArenaLayout &lay, int lpa, int pit) {
    const int raw = lay.off_laser + OFB_GROUP_BYTES * (pit * lpa / 8) + 9 * lay.SP;
    return ((raw + 63) & ~127) + 64;                     // smallest size >= raw that is 64 modulo 128
}

template <int LPA, int PIT, int MINB>
static int launch_step_t(ofb_arenas *h, const int2 *act, float4 *obs, const BotSpec &bots, cudaStream_t st) {
    const int apw = 32 / LPA;
    const long long warps = (h->n_arenas + apw - 1) / apw;
    const long long blocks = (warps * 32 + STEP_THREADS - 1) / STEP_THREADS;
    if (blocks == 0) return OFB_OK;
    const int tile = step_tile_bytes(h->lay, LPA, PIT);
    const int smem = STEP_BAR_BYTES + (STEP_THREADS / 32) * apw * tile;
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_step<LPA, PIT, MINB>, smem));
    k_step<LPA, PIT, MINB><<<(unsigned)blocks, STEP_THREADS, smem, st>>>(h->state, h->lay, act, obs, h->n_arenas, bots, tile);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

static int launch_step(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, const BotSpec &bots, void *stream) {
    OFB_CUDA_CHECK(cudaSetDevice(h->device));           // a thread may drive handles on several GPUs
    cudaStream_t st = (cudaStream_t)stream;
    const int lpa = lpa_for(h->lay.S, h->n_arenas);
    const int2 *act = reinterpret_cast<const int2 *>(actions_dev);
    float4 *obs = reinterpret_cast<float4 *>(obs_out_dev);
    // 6 chunk iterations (48 lasers at 8 lanes per arena) staged per pass, 7 CTAs = 28 warps per SM: measured best of
    // (8 iterations, 5 CTAs), (6, 7), (5, 8) on B200 -- profiles/r01_step_tuning.md
    if (lpa == 8) return launch_step_t<8, 6, 7>(h, act, obs, bots, st);
    if (lpa == 16) return launch_step_t<16, 6, 7>(h, act, obs, bots, st);
    return launch_step_t<32, 6, 7>(h, act, obs, bots, st);
}

extern "C" int ofb_step(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, void *stream) {
    if (!h || !actions_dev) { ofb_set_error("ofb_step: null argument"); return OFB_E_ARG; }
    BotSpec none = {};
    return launch_step(h, actions_dev, obs_out_dev, none, stream);
}

extern "C" int ofb_step_bots(ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed, int64_t arena0, uint32_t step,
                             const int16_t *actions_dev, float *obs_out_dev, void *stream) {
    if (!h || bot_kind < 0 || (bot_kind > OFB_BOT_STRESS && bot_kind != OFB_BOT_EXTERNAL)) {
        ofb_set_error("ofb_step_bots: bad argument");
        return OFB_E_ARG;
    }
    if (!actions_dev && (kinds_dev || bot_kind == OFB_BOT_EXTERNAL)) {
        ofb_set_error("ofb_step_bots: ships of kind 'external' need an actions buffer");
        return OFB_E_ARG;
    }
    BotSpec b = {1, bot_kind, kinds_dev, seed, (long long)arena0, step};
    return launch_step(h, actions_dev, obs_out_dev, b, stream);
}

// ---- fused frame: launch ------------------------------------------------------------------------------------------
int ofb_raster_bits_launch(const ofb_arenas *h, void *out_dev, cudaStream_t st);     // ofb_raster.cu

// Tuning knobs of the fused frame kernel (experiments and the geometry tests): only looked at when OFB_FRAME_TUNE is set, so
// that a production launch costs one getenv.
static thread_local bool g_frame_tune = false;
static int env_int(const char *name, int dflt) {
    if (!g_frame_tune) return dflt;
    const char *v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// debug: 1 if the calling thread's last ofb_frame* call ran as the single fused launch, 0 if as k_step + k_raster
static thread_local int g_last_frame_fused = -1;
extern "C" int ofb_debug_last_frame_fused(void) { return g_last_frame_fused; }
// debug: when set (ofb_debug_frame_prof), k_frame writes per-warp-role cycle counters there: long long [grid][32][8]
static long long *g_frame_prof = nullptr;
extern "C" int ofb_debug_frame_prof(long long *buf_dev) { g_frame_prof = buf_dev; return OFB_OK; }

template <int LPA, int PIT, int SW, int NG, int NBUF>
static int launch_frame_t(ofb_arenas *h, const int2 *act, float4 *obs, const BotSpec &bots, uint32_t *maps, cudaStream_t st,
                          int n_sm, int smem_max, bool *fits) {
    constexpr int APW = 32 / LPA;
    const int tile = step_tile_bytes(h->lay, LPA, PIT);
    const int post_cap = min(h->lay.L, env_int("OFB_FRAME_POSTCAP", FR_POST_CAP));
    const int slot = post_slot_bytes(h->lay.SP, post_cap);
    const int map_bytes = (h->lay.W * h->lay.H / 32) * 8;
    const int fixed = FR_BAR_BYTES + SW * APW * tile + 128 + NG * NBUF * map_bytes;
    int K = (smem_max - fixed) / slot;
    K = min(K, min(FR_MAX_K, env_int("OFB_FRAME_K", 4 * SW * APW)));
    // Every slot must have ONE producer (stepper warp, tile) and ONE consumer (raster group) for the phase parities of
    // its full / empty barriers to be sound: arenas i and i + K share a slot, so K is a multiple of both the arenas in
    // flight per stepper round (SW * APW) and the number of raster groups.
    int unit = SW * APW;
    while (unit % NG) unit += SW * APW;
    K -= K % unit;
    *fits = K >= unit && (map_bytes % 16) == 0;
    if (!*fits) return OFB_OK;
    const int smem = fixed + K * slot;
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_frame<LPA, PIT, SW, NG, NBUF>, smem));
    const unsigned grid = (unsigned)(h->n_arenas < (int64_t)n_sm ? h->n_arenas : (int64_t)n_sm);
    k_frame<LPA, PIT, SW, NG, NBUF><<<grid, (SW + NG * FR_RASTER_WARPS) * 32, smem, st>>>(h->state, h->lay, act, obs, h->n_arenas,
                                                                                         bots, tile, slot, post_cap, K, maps, g_frame_prof,
                                                                                         g_frame_prof ? env_int("OFB_FRAME_DBG", 0) : 0);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// One launch when the arena's post-step list fits the shared-memory ring (default arenas); otherwise (e.g. 32 ships x 2048
// laser slots) the same result as two launches, K1 then K2.
static int launch_frame(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, const BotSpec &bots, void *maps_dev,
                        void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int2 *act = reinterpret_cast<const int2 *>(actions_dev);
    float4 *obs = reinterpret_cast<float4 *>(obs_out_dev);
    uint32_t *maps = reinterpret_cast<uint32_t *>(maps_dev);
    if (h->n_arenas == 0) return OFB_OK;
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    {
        const char *tune = getenv("OFB_FRAME_TUNE");
        g_frame_tune = tune && *tune && *tune != '0';
    }
    static thread_local int n_sm = 0, smem_max = 0, dev_cached = -1;
    if (dev_cached != h->device) {
        OFB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, h->device));
        OFB_CUDA_CHECK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device));
        dev_cached = h->device;
    }
    const int S = h->lay.S;
    // Geometry (profiles/r01_step_tuning.md): up to a few hundred arenas per SM the kernel is bound by one step latency plus
    // the per-SM store rate and two raster groups with 4 stepper warps are best (4 096 arenas: 36.8 us); long ranges are
    // HBM-bound and want stepper throughput instead: one arena per warp, 12 stepper warps, one raster group with three
    // buffers (131 072 arenas: 917 vs 975 us).
    const bool long_range = h->n_arenas >= 256ll * n_sm;
    const int lpa = env_int("OFB_FRAME_LPA", (S <= 16 && !long_range) ? 16 : 32);
    const int sw = env_int("OFB_FRAME_SW", long_range ? 12 : 4), ng = env_int("OFB_FRAME_NG", long_range ? 1 : 2);
    const int nbuf = env_int("OFB_FRAME_NBUF", long_range ? 3 : 2);
    bool fits = false;
    int rc = OFB_OK;
    if (!env_int("OFB_FRAME_SPLIT", 0) && lpa >= S) {
#define FR_CASE(L_, P_, W_, G_, B_) \
        if (lpa == L_ && sw == W_ && ng == G_ && nbuf == B_) \
            rc = launch_frame_t<L_, P_, W_, G_, B_>(h, act, obs, bots, maps, st, n_sm, smem_max, &fits);
        FR_CASE(8, 6, 8, 2, 2) FR_CASE(8, 6, 8, 1, 3) FR_CASE(8, 6, 8, 2, 1)
        FR_CASE(16, 4, 4, 2, 2) FR_CASE(16, 4, 8, 2, 2) FR_CASE(16, 4, 8, 1, 3) FR_CASE(16, 4, 4, 3, 1)
        FR_CASE(32, 3, 4, 2, 2) FR_CASE(32, 3, 8, 2, 2) FR_CASE(32, 3, 12, 1, 3)
#undef FR_CASE
        if (rc != OFB_OK) return rc;
    }
    g_last_frame_fused = fits ? 1 : 0;
    if (fits) return OFB_OK;
    rc = launch_step(h, actions_dev, obs_out_dev, bots, stream);
    if (rc != OFB_OK) return rc;
    return ofb_raster_bits_launch(h, maps_dev, st);
}

extern "C" int ofb_frame(ofb_arenas *h, const int16_t *actions_dev, float *obs_out_dev, void *maps_bits_dev, void *stream) {
    if (!h || !actions_dev || !maps_bits_dev) { ofb_set_error("ofb_frame: null argument"); return OFB_E_ARG; }
    BotSpec none = {};
    return launch_frame(h, actions_dev, obs_out_dev, none, maps_bits_dev, stream);
}

extern "C" int ofb_frame_bots(ofb_arenas *h, int bot_kind, const uint8_t *kinds_dev, uint64_t seed, int64_t arena0, uint32_t step,
                              const int16_t *actions_dev, float *obs_out_dev, void *maps_bits_dev, void *stream) {
    if (!h || !maps_bits_dev || bot_kind < 0 || (bot_kind > OFB_BOT_STRESS && bot_kind != OFB_BOT_EXTERNAL)) {
        ofb_set_error("ofb_frame_bots: bad argument");
        return OFB_E_ARG;
    }
    if (!actions_dev && (kinds_dev || bot_kind == OFB_BOT_EXTERNAL)) {
        ofb_set_error("ofb_frame_bots: ships of kind 'external' need an actions buffer");
        return OFB_E_ARG;
    }
    BotSpec b = {1, bot_kind, kinds_dev, seed, (long long)arena0, step};
    return launch_frame(h, actions_dev, obs_out_dev, b, maps_bits_dev, stream);
}

// Host-buffer form of ofb_step: the drop-in call for a host-side bot loop (Battleground.frame with
// Python bots).  actions_host / obs_host should be pinned; everything is asynchronous on `stream`.
extern "C" int ofb_step_host(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *stream) {
    if (!h || !actions_host) { ofb_set_error("ofb_step_host: null argument"); return OFB_E_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_ship = (size_t)h->n_arenas * h->lay.S;
    OFB_CUDA_CHECK(cudaMemcpyAsync(h->stage_actions, actions_host, n_ship * 4 * sizeof(int16_t),
                                   cudaMemcpyHostToDevice, st));
    int rc = ofb_step(h, h->stage_actions, obs_host ? h->stage_obs : nullptr, st);
    if (rc != OFB_OK) return rc;
    if (obs_host)
        OFB_CUDA_CHECK(cudaMemcpyAsync(obs_host, h->stage_obs, n_ship * 8 * sizeof(float), cudaMemcpyDeviceToHost, st));
    return OFB_OK;
}

// Pipelined form of ofb_step_host for a host loop that replays / produces one action batch per frame: the
// H2D copy of frame k+1 and the D2H copy of frame k's observation heads run on their own streams, so both PCIe
// directions overlap the kernels of `stream` (step k, raster k).  Staging is double-buffered:
//   s_h2d : [wait step k-2] actions_host -> pipe_actions[k&1]
//   stream: [wait h2d k] [wait d2h k-2] step k (reads pipe_actions[k&1], writes pipe_obs[k&1])
//   s_d2h : [wait step k] pipe_obs[k&1] -> obs_host
// obs_host is valid after ofb_host_wait(); actions_host may be reused after the same call (or after the next
// ofb_step_host_async returns two frames later).
int ofb_pipe_init(ofb_arenas *h);
extern "C" int ofb_obs_pack_i16(const float *obs_dev, int16_t *out_dev, int64_t n_rows, void *stream);
static int step_host_async(ofb_arenas *h, const int16_t *actions_host, float *obs_host, int16_t *obs16_host, void *maps_bits_dev,
                           void *stream) {
    if (!h || !actions_host) { ofb_set_error("ofb_step_host_async: null argument"); return OFB_E_ARG; }
    int rc = ofb_pipe_init(h);
    if (rc != OFB_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n_ship = (size_t)h->n_arenas * h->lay.S;
    const int i = (int)(h->host_seq & 1ull);
    const bool want_obs = obs_host || obs16_host;
    if (h->host_seq >= 2) OFB_CUDA_CHECK(cudaStreamWaitEvent(h->s_h2d, h->ev_step[i], 0));    // step k-2 has read pipe_actions[i]
    OFB_CUDA_CHECK(cudaMemcpyAsync(h->pipe_actions[i], actions_host, n_ship * 4 * sizeof(int16_t), cudaMemcpyHostToDevice, h->s_h2d));
    OFB_CUDA_CHECK(cudaEventRecord(h->ev_h2d[i], h->s_h2d));
    OFB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_h2d[i], 0));
    if (want_obs && h->host_seq >= 2) OFB_CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_d2h[i], 0));   // pipe_obs[i] has left
    rc = maps_bits_dev ? ofb_frame(h, h->pipe_actions[i], want_obs ? h->pipe_obs[i] : nullptr, maps_bits_dev, st)
                       : ofb_step(h, h->pipe_actions[i], want_obs ? h->pipe_obs[i] : nullptr, st);
    if (rc != OFB_OK) return rc;
    if (obs16_host && (rc = ofb_obs_pack_i16(h->pipe_obs[i], h->pipe_obs16[i], (int64_t)n_ship, st)) != OFB_OK) return rc;
    OFB_CUDA_CHECK(cudaEventRecord(h->ev_step[i], st));
    // ev_d2h[i] also stands for "actions_host of frame k has been consumed" (ofb_host_wait): it must follow the step -- and
    // with it the H2D copy the step waited for -- even when no observation heads are copied back
    OFB_CUDA_CHECK(cudaStreamWaitEvent(h->s_d2h, h->ev_step[i], 0));
    if (obs_host)
        OFB_CUDA_CHECK(cudaMemcpyAsync(obs_host, h->pipe_obs[i], n_ship * 8 * sizeof(float), cudaMemcpyDeviceToHost, h->s_d2h));
    if (obs16_host)
        OFB_CUDA_CHECK(cudaMemcpyAsync(obs16_host, h->pipe_obs16[i], n_ship * 5 * sizeof(int16_t), cudaMemcpyDeviceToHost, h->s_d2h));
    OFB_CUDA_CHECK(cudaEventRecord(h->ev_d2h[i], h->s_d2h));
    h->host_seq++;
    return OFB_OK;
}

extern "C" int ofb_step_host_async(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *stream) {
    return step_host_async(h, actions_host, obs_host, nullptr, nullptr, stream);
}
// Same pipeline with the fused frame kernel: the frame's observation maps (OFB_MAP_BITS) are written to maps_bits_dev.
extern "C" int ofb_frame_host_async(ofb_arenas *h, const int16_t *actions_host, float *obs_host, void *maps_bits_dev, void *stream) {
    if (!maps_bits_dev) { ofb_set_error("ofb_frame_host_async: null maps buffer"); return OFB_E_ARG; }
    return step_host_async(h, actions_host, obs_host, nullptr, maps_bits_dev, stream);
}
// ... and the observation heads copied back in the compact form of ofb_obs_pack_i16 (10 instead of 32 bytes per ship).
extern "C" int ofb_frame_host_async_i16(ofb_arenas *h, const int16_t *actions_host, int16_t *obs16_host, void *maps_bits_dev, void *stream) {
    if (!maps_bits_dev || !obs16_host) { ofb_set_error("ofb_frame_host_async_i16: null buffer"); return OFB_E_ARG; }
    return step_host_async(h, actions_host, nullptr, obs16_host, maps_bits_dev, stream);
}

// Block until every copy queued by ofb_step_host_async has completed (obs_host readable, actions_host reusable).
extern "C" int ofb_host_wait(ofb_arenas *h) {
    if (!h) { ofb_set_error("ofb_host_wait: null argument"); return OFB_E_ARG; }
    if (!h->pipe_ready || h->host_seq == 0) return OFB_OK;
    OFB_CUDA_CHECK(cudaEventSynchronize(h->ev_d2h[(h->host_seq - 1) & 1ull]));
    return OFB_OK;
}

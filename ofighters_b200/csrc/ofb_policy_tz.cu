// Pointer head on tensor cores, "block-Toeplitz" form (sm_100a, tcgen05 + TMEM + bulk async copies).
//
// A 3x3 convolution with 8 channels in and a handful out is a poor GEMM if one output pixel is one row of M: every
// 16-byte pixel read as A operand feeds only Cout multiply-accumulates, and the tensor pipe spends its time reading
// shared memory (measured: ~75 cycles per M=128 x K=16 tile whatever N is).  Here one row of M is a BLOCK of B = 8
// consecutive pixels of an image row, K runs over the B + 2 input pixels the block's outputs depend on (x 8 channels),
// and N = B x Cout' holds all outputs of the block; the weights become a banded (Toeplitz) matrix with 3 of every
// B + 2 pixel-columns non-zero.  The zero columns are wasted math on a pipe that has math to spare; what it buys is
// (B + 2) / (3 B) = 0.42x the A-operand bytes per output and an epilogue thread that owns 8 x 4 outputs at once.
//
// Activations therefore live de-interleaved by x mod 8 ("plane layout"): plane q of an H x H x 8 image is the array
//     plane[q][y * P + xb + 1]  (16 B = 8 bf16 channels each),   P = H/8 + 1,   pixel x = 8 xb + q
// so that for M row m = y * P + xb the K-chunk of pixel offset p (x = 8 xb + p - 1) is
//     p = 0      plane 7 at m          p = 1..8   plane p-1 at m + 1          p = 9   plane 0 at m + 2
// i.e. every chunk is a K-major / no-swizzle UMMA operand (8 consecutive m = one 128-byte core matrix) and a vertical
// tap is the same descriptor shifted by P slots.  Slot 0 of a row of plane 7 is the left halo (x = -1), slot 0 of row
// y + 1 of plane 0 the right halo (x = H) of row y; the consumer fills them after staging.
//
// k_tz_up4: upsampling4 + upconv4 (qlearnIA_V2.py:184-186) + the pointer argmax (:218-220).  The bilinear x2 is folded
// into 4 output phases on the 200 x 200 grid (N = 8 px x 4 phases = 32), so one strip of 19 rows is 4 tiles x 15 MMAs;
// the 1-pixel border ring, where the convolution's zero padding breaks the folded form, is corrected after the tile
// loop exactly as in the pixel-linear kernels (ofb_policy_dev.cuh: up_ring_tap).
#include "ofb_common.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_tc_ptx.cuh"

#define TZ4_H 200                     // input grid of upconv4
#define TZ4_P 26                      // slots per plane row: 25 blocks + 1 halo
#define TZ4_R 14                      // rows per strip: 14 * 26 = 364 M rows = 3 tiles of 128
#define TZ4_T 3
#define TZ4_PS 440                    // slots per plane in shared memory: (R + 2) * P + 2 + tile overrun (20), rounded to 8
#define TZ4_NST 3                     // stages of the strip buffer / TMEM accumulators (two in flight while one is computed: ~110 KB per SM)
#define TZ4_N 32
#define TZ4_NT 160                    // 4 draining warps (one per TMEM lane quarter) + 1 producer / MMA-issuer warp
#define TZ4_STRIPS ((TZ4_H + TZ4_R - 1) / TZ4_R)

struct Tz4Args {
    const __nv_bfloat16 *in;          // plane layout [item][8][200][26][8]
    const __nv_bfloat16 *wt;          // B operand [3 dy][5 k-steps][2 chunks][32 n][8 cin]
    const float *ring_w;              // un-phased fp32 weights [9][8][1] (ring pixels)
    float bias;
    float *ptr_out;                   // optional dense map [item][400][400]
    float *amax_val;                  // [item][TZ4_STRIPS]
    int *amax_idx;
    long long *dbg;                   // optional: clock64 stamps of CTA (5, 0) at the phase boundaries
    int legacy;                       // TF1.x legacy bilinear in the ring corrections
};

#define TZ4_NG 3                      // draining groups of 4 warps; group g owns stage g (TZ4_NG == TZ4_NST)
#define TZ4_NTP (32 * (4 * TZ4_NG + 1 + TZ4_T))                   // persistent kernel: 2 groups of 4 draining warps (one per stage) + TMA producer warp + MMA / halo warp
#define TZ4_ABYTES (8 * TZ4_PS * 16)
#define TZ4_WBYTES (3 * 5 * 2 * TZ4_N * 16)
#define TZ4_RING (4 * TZ4_R + 400)    // floats: 2R left, 2R right, 400 top / bottom
#define TZ4_SCR (2 * (TZ4_R + 2) * 3 + 202 * 3 + 2)   // ring-correction scratch: column dots, row dots

struct Tz4Smem {
    static constexpr unsigned off_w = 0;
    static constexpr unsigned off_a = TZ4_WBYTES;                          // TZ4_NST stages
    static constexpr unsigned off_ring = off_a + TZ4_NST * TZ4_ABYTES;     // one per draining group
    static constexpr unsigned off_corr = off_ring + TZ4_NG * TZ4_RING * 4; // one per draining group: ring corrections
    static constexpr unsigned off_scr = off_corr + TZ4_NG * TZ4_RING * 4;  // one per draining group
    static constexpr unsigned off_aux = off_scr + TZ4_NG * TZ4_SCR * 4;    // 72 ring weights + argmax scratch (16 floats per group)
    static constexpr unsigned off_bar = (off_aux + (72 + 16 * TZ4_NG) * 4 + 15) & ~15u;
    // barriers: wbar, full_a[NST], a_empty[NST], halo[NST], halo2[NST], tmem_empty[NST], acc_full[NST][T]
    static constexpr unsigned n_bar = 1 + 5 * TZ4_NST + TZ4_NST * TZ4_T;
    static constexpr unsigned total = off_bar + n_bar * 8 + 16;
};

// plane-layout strip in shared memory as an image accessor: rows y0-1 .. y0+R, cols -1 .. 200 (halos filled)
struct PlaneStrip4 {
    const uint4 *pl;
    int y0;
    __device__ __forceinline__ uint4 operator()(int y, int x) const { return pl[(x & 7) * TZ4_PS + (y - y0 + 1) * TZ4_P + (x >> 3) + 1]; }
};

__device__ __forceinline__ void named_sync_128(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

// The 32 accumulator columns of one M row (y_local, block xb) of upconv4 -> biased values with the hi-res border pixels
// masked out (-inf); when `ring` is given the masked values are parked there for the ring pass (column slots own the
// corners).  Column n = xo * 4 + a * 2 + b is output pixel (2 i + a, 16 xb + 2 xo + b).
__device__ __forceinline__ void up4_row_values(const uint32_t *r, float bias, int xb, int yl, int i, float *ring, float *v) {
    const float NINF = -INFINITY;
#pragma unroll
    for (int n = 0; n < 32; n++) v[n] = __uint_as_float(r[n]) + bias;
    if (xb == 0) {
        if (ring) { ring[2 * yl] = v[0]; ring[2 * yl + 1] = v[2]; }
        v[0] = NINF; v[2] = NINF;
    }
    if (xb == 24) {
        if (ring) { ring[2 * TZ4_R + 2 * yl] = v[29]; ring[2 * TZ4_R + 2 * yl + 1] = v[31]; }
        v[29] = NINF; v[31] = NINF;
    }
    if (i == 0 || i == TZ4_H - 1) {
        const int ph_a = i == 0 ? 0 : 1;
#pragma unroll
        for (int rem = 0; rem < 16; rem++) {
            const int X = 16 * xb + rem;
            if (X == 0 || X == 2 * TZ4_H - 1) continue;
            if (ph_a == 0) { if (ring) ring[4 * TZ4_R + X] = v[(rem >> 1) * 4 + (rem & 1)]; v[(rem >> 1) * 4 + (rem & 1)] = NINF; }
            else { if (ring) ring[4 * TZ4_R + X] = v[(rem >> 1) * 4 + 2 + (rem & 1)]; v[(rem >> 1) * 4 + 2 + (rem & 1)] = NINF; }
        }
    }
}

// Persistent, warp-specialised: one CTA per SM walks the (item, strip) work list with two stages of everything --
//   warp 4 (producer)  : bulk copies of strip k + 1 while strip k is computed            full_a[s] / a_empty[s]
//   warp 5 (MMA)       : halo slots, then 4 tiles x 15 MMAs into TMEM stage s             halo[s], acc_full[s][t] / tmem_empty[s]
//   warps 0-11 (drain) : group g = warps 4g..4g+3 owns stage g (every third strip): TMEM -> registers -> masked max,
//                        border ring, argmax partial.  The index of the maximum is recovered by re-reading the one tile
//                        that holds it, so the per-tile work is branch-free.
static_assert(TZ4_NG == TZ4_NST, "group g drains stage g");
__global__ void __launch_bounds__(TZ4_NTP, 1)
k_tz_up4(const Tz4Args a, const int n_work) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4 *sw = reinterpret_cast<uint4 *>(smem + Tz4Smem::off_w);
    float *aux = reinterpret_cast<float *>(smem + Tz4Smem::off_aux);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Tz4Smem::off_bar);
    uint64_t *wbar = bars, *full_a = bars + 1, *a_empty = full_a + TZ4_NST, *halo = a_empty + TZ4_NST, *halo2 = halo + TZ4_NST,
             *tmem_empty = halo2 + TZ4_NST, *acc_full = tmem_empty + TZ4_NST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Tz4Smem::off_bar + Tz4Smem::n_bar * 8);
#define TZ_STAMP(k, who) do { if (a.dbg && tid == (who) && blockIdx.x == 5) a.dbg[k] = clock64(); } while (0)
    // per-work stamps of CTA 5: dbg[16 + 8 * k + j] for the CTA's k-th work item (k < 14)
#define TZ_STAMPK(k, j) do { if (a.dbg && blockIdx.x == 5 && (k) < 14) a.dbg[16 + 8 * (k) + (j)] = clock64(); } while (0)
    TZ_STAMP(0, 0);

    if (tid == 0) {
        mbar_init(wbar, 1);
        for (int s = 0; s < TZ4_NST; s++) {
            mbar_init(&full_a[s], 1);
            mbar_init(&a_empty[s], 4 + TZ4_T);             // 4 draining warps (ring corrections read) + the MMA warps (tcgen05.commit)
            mbar_init(&halo[s], 1);
            mbar_init(&halo2[s], 1);                       // released once, awaited by the other MMA warps
            mbar_init(&tmem_empty[s], 4);
            for (int t = 0; t < TZ4_T; t++) mbar_init(&acc_full[s * TZ4_T + t], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * TZ4_NG) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            mbar_expect_tx(wbar, TZ4_WBYTES + 72 * 4);
            bulk_g2s(sw, a.wt, TZ4_WBYTES, wbar);
            bulk_g2s(aux, a.ring_w, 72 * 4, wbar);
            int k = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, k++) {
                const int s = k % TZ4_NST, item = w / TZ4_STRIPS, y0 = (w % TZ4_STRIPS) * TZ4_R;
                if (k >= TZ4_NST) mbar_wait(&a_empty[s], (uint32_t)((k / TZ4_NST - 1) & 1));
                TZ_STAMPK(k, 0);                                           // loads issued
                uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz4Smem::off_a + s * TZ4_ABYTES);
                const int ylo = max(y0 - 1, 0), yhi = min(y0 + TZ4_R, TZ4_H - 1);
                const uint32_t run = (uint32_t)(yhi - ylo + 1) * TZ4_P * 16;
                const bool top = y0 == 0, bot = y0 + TZ4_R > TZ4_H - 1;        // a halo row outside the image is needed
                mbar_expect_tx(&full_a[s], 8u * (run + (top ? TZ4_P * 16 : 0) + (bot ? TZ4_P * 16 : 0)));
                const uint8_t *src = reinterpret_cast<const uint8_t *>(a.in) + (size_t)item * (8 * TZ4_H * TZ4_P * 16);
                for (int q = 0; q < 8; q++) {
                    const uint8_t *pq = src + (size_t)q * TZ4_H * TZ4_P * 16;
                    uint4 *dq = sa + q * TZ4_PS;
                    bulk_g2s(dq + (ylo - (y0 - 1)) * TZ4_P, pq + (size_t)ylo * TZ4_P * 16, run, &full_a[s]);
                    if (top) bulk_g2s(dq, pq, TZ4_P * 16, &full_a[s]);                                         // row -1 := row 0
                    if (bot) bulk_g2s(dq + (TZ4_H - (y0 - 1)) * TZ4_P, pq + (size_t)(TZ4_H - 1) * TZ4_P * 16, TZ4_P * 16, &full_a[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp > 4 * TZ4_NG) {
        // ------------------------------------------------------------ halo slots + MMA issue.  A single thread issues an
        // MMA every ~75 cycles (descriptor moves to uniform registers), far slower than the pipe retires these small
        // MMAs, so one warp per tile issues: warp 9 also fixes the halo slots first.
        mbar_wait(wbar, 0);
        constexpr uint32_t IDESC = instr_desc(TZ4_N);
        const uint32_t sw16 = smem_u32(sw) >> 4;
        const int half = warp - 4 * TZ4_NG - 1;            // tile issued by this warp
        int k = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, k++) {
            const int s = k % TZ4_NST;
            const uint32_t par = (uint32_t)((k / TZ4_NST) & 1);
            uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz4Smem::off_a + s * TZ4_ABYTES);
            if (half == 0) {
                mbar_wait(&full_a[s], par);
                if (lane == 0) TZ_STAMPK(k, 1);                            // loads landed
                // x = -1 := x = 0 (plane 7 slot 0 of the row), x = 200 := x = 199 (plane 0 slot 0 of the next row)
                for (int i = lane; i < 2 * (TZ4_R + 2); i += 32) {
                    const int r = i >> 1;
                    if (i & 1) sa[0 * TZ4_PS + (r + 1) * TZ4_P] = sa[7 * TZ4_PS + r * TZ4_P + 25];
                    else sa[7 * TZ4_PS + r * TZ4_P] = sa[0 * TZ4_PS + r * TZ4_P + 1];
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) { mbar_arrive(&halo[s]); mbar_arrive(&halo2[s]); }   // ring pass of the draining warps / second issuer
            } else {
                mbar_wait(&halo2[s], par);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            // the whole warp stays converged so that the descriptors are warp-uniform values; one elected lane issues
            if (k >= TZ4_NST) { mbar_wait(&tmem_empty[s], par ^ 1u); tc_fence_after(); }
            if (half == 0 && lane == 0) TZ_STAMPK(k, 2);                   // MMA issue starts
            {
                const bool leader = elect_one();
                const uint32_t sa16 = smem_u32(sa) >> 4;
                const int t = half;                                        // warp 4 NG + 1 + t issues tile t
                const uint32_t d = tmem_base + (uint32_t)((s * TZ4_T + t) * TZ4_N);
#pragma unroll
                for (int u = 0; u < 3; u++)
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {
                        // first / second K-chunk of the step: (plane, slot offset); see the header comment
                        const int q0 = ks == 0 || ks == 4 ? 0 : 2 * ks - 1, o0 = ks == 4 ? 2 : 1;
                        const uint32_t lbo = (ks == 0 || ks == 4) ? 7u * TZ4_PS - 1u : (uint32_t)TZ4_PS;
                        const uint64_t ad = smem_desc(sa16 + (uint32_t)(q0 * TZ4_PS + u * TZ4_P + 128 * t + o0), lbo, 8);
                        const uint64_t bd = smem_desc(sw16 + (uint32_t)((u * 5 + ks) * 2 * TZ4_N), TZ4_N, 8);
                        if (leader) tc_mma(d, ad, bd, IDESC, (u | ks) ? 1u : 0u);
                    }
                if (leader) {
                    tc_commit(&acc_full[s * TZ4_T + t]);
                    tc_commit(&a_empty[s]);                                // the strip buffer is free once these MMAs have read it
                }
                if (half == 0 && lane == 0) TZ_STAMPK(k, 3);               // MMA issue done
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------ drain: thread = M row (y_local, block) of the tile
        mbar_wait(wbar, 0);                               // ring weights
        const int grp = warp >> 2, gt = tid & 127, gw = warp & 3;     // group g drains every other work item
        float *ring = reinterpret_cast<float *>(smem + Tz4Smem::off_ring) + grp * TZ4_RING;
        float *sv = aux + 72 + 16 * grp;                  // per group: 4 warp maxima, then 4 (value, index) pairs
        int *si = reinterpret_cast<int *>(sv + 8);
        int k = grp;
        for (int w = blockIdx.x + grp * gridDim.x; w < n_work; w += TZ4_NG * gridDim.x, k += TZ4_NG) {
            const int s = k % TZ4_NST, item = w / TZ4_STRIPS, strip = w % TZ4_STRIPS, y0 = strip * TZ4_R;
            const uint32_t par = (uint32_t)((k / TZ4_NST) & 1);
            const int rows_valid = min(TZ4_R, TZ4_H - y0);
            const uint4 *sa = reinterpret_cast<const uint4 *>(smem + Tz4Smem::off_a + s * TZ4_ABYTES);
            const uint32_t tm = tmem_base + ((uint32_t)(gw * 32) << 16) + (uint32_t)(s * TZ4_T * TZ4_N);
            float *corr = reinterpret_cast<float *>(smem + Tz4Smem::off_corr) + grp * TZ4_RING;
            constexpr int Wo = 2 * TZ4_H;
            const bool first = strip == 0, last = strip == TZ4_STRIPS - 1;
            const int nslots = 4 * TZ4_R + ((first || last) ? Wo : 0);             // only the first / last strip own a border row
            // ---- border ring, part 1 (while the MMAs run): the contribution of the taps that fall outside the image, per
            //      ring pixel (4 lanes per pixel: lanes 0..2 evaluate one out-of-range tap each, lane 0 also the two extra
            //      taps of a corner).  It needs the staged strip only, so the strip buffer is released right after it.
            mbar_wait(&halo[s], par);                     // strip + halo slots are in place (acquire)
            {
                // Every out-of-range tap of a border-COLUMN pixel sits in column -1 (or 400) of the extended upsampled map,
                // which is a vertical blend of low-res column 0 (199) alone: cdots[side][row][dy] = w[dy][edge dx] . L[row][edge],
                // then 6 terms per ring pixel.  Likewise a border-ROW pixel (first / last strip) only sees the vertical blend of
                // low-res rows -1/0 (199/200): rdots[x][dx] = w[edge dy][dx] . blend[x], 6 terms per pixel.  Corner pixels live
                // in the column slots and add the two row taps that are not in their column.
                float *cdots = reinterpret_cast<float *>(smem + Tz4Smem::off_scr) + grp * TZ4_SCR, *rdots = cdots + 2 * (TZ4_R + 2) * 3;
                const int dye = first ? 0 : 2, Yedge = first ? 0 : Wo - 1;
                int rlo = 0, rhi = 0; float rwl = 0.f, rwh = 0.f;
                bil_tap_ext(Yedge + dye - 1, rlo, rhi, rwl, rwh, a.legacy);
                if (gt < 2 * (TZ4_R + 2) * 3) {
                    const int side = gt / ((TZ4_R + 2) * 3), rem = gt % ((TZ4_R + 2) * 3), r = rem / 3, dy = rem % 3;
                    float x[8];
                    unpack_bf8(sa[(side ? 7 : 0) * TZ4_PS + r * TZ4_P + (side ? 25 : 1)], x);
                    const float *wp = aux + (dy * 3 + (side ? 2 : 0)) * 8;
                    float d = 0.f;
#pragma unroll
                    for (int ci = 0; ci < 8; ci++) d += x[ci] * wp[ci];
                    cdots[gt] = d;
                }
                if (first || last) {
                    const PlaneStrip4 Ls{sa, y0};
                    for (int q = gt; q < 202 * 3; q += 128) {
                        const int x = q / 3 - 1, dx = q % 3;
                        float lo[8], hi[8];
                        unpack_bf8(Ls(rlo, x), lo);
                        unpack_bf8(Ls(rhi, x), hi);
                        const float *wp = aux + (dye * 3 + dx) * 8;
                        float d = 0.f;
#pragma unroll
                        for (int ci = 0; ci < 8; ci++) d += (rwl * lo[ci] + rwh * hi[ci]) * wp[ci];
                        rdots[q] = d;
                    }
                }
                named_sync_128(1 + 2 * TZ4_NG + grp);
                for (int slot = gt; slot < nslots; slot += 128) {
                    float c = 0.f;
                    if (slot < 4 * TZ4_R) {
                        const int side = slot / (2 * TZ4_R), Yl = slot % (2 * TZ4_R), Y = 2 * y0 + Yl;
                        const float *ds = cdots + side * (TZ4_R + 2) * 3;
#pragma unroll
                        for (int dy = 0; dy < 3; dy++) {
                            int lo, hi; float wlo, whi;
                            bil_tap_ext(Y + dy - 1, lo, hi, wlo, whi, a.legacy);
                            c += wlo * ds[(lo - y0 + 1) * 3 + dy] + whi * ds[(hi - y0 + 1) * 3 + dy];
                        }
                        if ((first || last) && Y == Yedge) {        // corner: the two taps of the border row outside its column
                            const int X = side ? Wo - 1 : 0;
#pragma unroll
                            for (int dx = 0; dx < 3; dx++) {
                                if (dx == (side ? 2 : 0)) continue;
                                int lo, hi; float wlo, whi;
                                bil_tap_ext(X + dx - 1, lo, hi, wlo, whi, a.legacy);
                                c += wlo * rdots[(lo + 1) * 3 + dx] + whi * rdots[(hi + 1) * 3 + dx];
                            }
                        }
                    } else {
                        const int X = slot - 4 * TZ4_R;
#pragma unroll
                        for (int dx = 0; dx < 3; dx++) {
                            int lo, hi; float wlo, whi;
                            bil_tap_ext(X + dx - 1, lo, hi, wlo, whi, a.legacy);
                            c += wlo * rdots[(lo + 1) * 3 + dx] + whi * rdots[(hi + 1) * 3 + dx];
                        }
                    }
                    corr[slot] = c;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[s]);      // this warp is done with stage s of the strip buffer
            if (gt == 0) TZ_STAMPK(k, 4);                 // ring corrections done
            float best_v = -INFINITY;
            int bt = 0;
            for (int t = 0; t < TZ4_T; t++) {
                mbar_wait(&acc_full[s * TZ4_T + t], par);
                tc_fence_after();
                uint32_t r[32];
                tc_ld32(tm + (uint32_t)(t * TZ4_N), r);
                tc_wait_ld();
                const int m = 128 * t + gt, yl = m / TZ4_P, xb = m - yl * TZ4_P;
                if (yl >= rows_valid || xb >= 25) continue;
                const int i = y0 + yl;
                float v[32];
                up4_row_values(r, a.bias, xb, yl, i, ring, v);
                if (a.ptr_out) {                          // optional dense map (predict()'s second output); ring pixels follow later
#pragma unroll
                    for (int n = 0; n < 32; n++)
                        if (v[n] != -INFINITY)
                            a.ptr_out[(size_t)item * 4 * TZ4_H * TZ4_H + (2 * i + ((n >> 1) & 1)) * 2 * TZ4_H + 16 * xb + 2 * (n >> 2) + (n & 1)] = v[n];
                }
#pragma unroll
                for (int h = 16; h > 0; h >>= 1)          // tree maximum
#pragma unroll
                    for (int n = 0; n < h; n++) v[n] = fmaxf(v[n], v[n + h]);
                if (v[0] > best_v) { best_v = v[0]; bt = t; }          // strict: the earliest tile (lowest indices) keeps a tie
            }
            if (gt == 0) TZ_STAMPK(k, 5);                 // tiles drained
            // ---- maximum of the strip's interior, then its first index in C order: only the thread(s) that hold the
            //      maximum look at their 32 values again (TMEM stage s is still intact)
            float wv = best_v;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wv = fmaxf(wv, __shfl_xor_sync(0xffffffffu, wv, o));
            if (lane == 0) sv[gw] = wv;
            named_sync_128(1 + grp);                      // warp maxima, ring values and ring corrections of all four warps are visible
            const float V = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
            const bool cand = best_v == V && V > -INFINITY;
            int best_i = 0x7fffffff;
            unsigned bal = __ballot_sync(0xffffffffu, cand);
            while (bal) {
                const int tt = __shfl_sync(0xffffffffu, bt, __ffs(bal) - 1);
                uint32_t r[32];
                tc_ld32(tm + (uint32_t)(tt * TZ4_N), r);
                tc_wait_ld();
                const bool mine = cand && bt == tt;
                if (mine) {
                    const int m = 128 * tt + gt, yl = m / TZ4_P, xb = m - yl * TZ4_P, i = y0 + yl;
                    float v[32];
                    up4_row_values(r, a.bias, xb, yl, i, nullptr, v);
#pragma unroll
                    for (int n = 31; n >= 0; n--) {       // descending C order within the thread: the last match is the lowest index
                        const int ph_a = (n >> 4) & 1, rem = n & 15;             // order: a, then xo, then b
                        if (v[(rem >> 1) * 4 + ph_a * 2 + (rem & 1)] == V) best_i = (2 * i + ph_a) * 2 * TZ4_H + 16 * xb + rem;
                    }
                }
                bal &= ~__ballot_sync(0xffffffffu, mine);
            }
            best_v = cand ? V : -INFINITY;
            tc_fence_before();                            // this warp's TMEM reads of stage s are complete
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[s]);

            // ---- border ring, part 2: folded value - out-of-range taps = the true (zero-padded) convolution
            for (int slot = gt; slot < nslots; slot += 128) {
                int Y, X;
                bool valid = true;
                if (slot < 2 * TZ4_R) { Y = 2 * y0 + slot; X = 0; valid = slot < 2 * rows_valid; }
                else if (slot < 4 * TZ4_R) { Y = 2 * y0 + slot - 2 * TZ4_R; X = Wo - 1; valid = slot - 2 * TZ4_R < 2 * rows_valid; }
                else { X = slot - 4 * TZ4_R; valid = X != 0 && X != Wo - 1; Y = first ? 0 : Wo - 1; }
                if (!valid) continue;
                const float val = ring[slot] - corr[slot];
                const int idx = Y * Wo + X;
                if (a.ptr_out) a.ptr_out[(size_t)item * Wo * Wo + idx] = val;
                if (amax_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
            }
            amax_warp(best_v, best_i);
            if (lane == 0) { sv[4 + gw] = best_v; si[4 + gw] = best_i; }
            named_sync_128(1 + TZ4_NG + grp);
            if (gt == 0) {
                for (int q = 1; q < 4; q++)
                    if (amax_better(sv[4 + q], si[4 + q], best_v, best_i)) { best_v = sv[4 + q]; best_i = si[4 + q]; }
                a.amax_val[(size_t)item * TZ4_STRIPS + strip] = best_v;
                a.amax_idx[(size_t)item * TZ4_STRIPS + strip] = best_i;
                TZ_STAMPK(k, 6);                          // work item finished
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    TZ_STAMP(12, 0);
}

// ================================================================================================
// k_tz_up3: upsampling3 + upconv3 + BN + ReLU (qlearnIA_V2.py:178-181), 100 x 100 x 4 -> 200 x 200 x 8.
// Same scheme with blocks of B = 4 pixels: the input (upconv2's output, written by k_heads) is de-interleaved by
// x mod 4 into 4 planes of P = 26 slots per row; K = 6 pixel chunks (3 K-steps), N = 4 px x 4 phases x 8 channels = 128.
// A strip is 9 rows = 234 M rows = 2 tiles x 9 MMAs.  The epilogue thread owns 4 low-res pixels = 8 x 2 output pixels
// and writes them straight in k_tz_up4's plane layout (its 8 x-values are exactly the 8 planes of one block).
// ================================================================================================
#define TZ3_H 100
#define TZ3_P 26
#define TZ3_R 9
#define TZ3_T 2
#define TZ3_PS 312                    // (R + 2) * P + 2 + tile overrun (256 - 234), rounded to 8
#define TZ3_N 128
#define TZ3_NST 2                     // stages = draining groups (2 x 2 tiles x 128 columns = all of TMEM)
#define TZ3_NTP (32 * (4 * TZ3_NST + 1 + TZ3_T + TZ3_NST))     // draining groups, producer, MMA warps, one ring warp per stage
#define TZ3_STRIPS ((TZ3_H + TZ3_R - 1) / TZ3_R)
#define TZ3_WBYTES (3 * 3 * 2 * TZ3_N * 16)
#define TZ3_ABYTES (4 * TZ3_PS * 16)
#define TZ3_RING (4 * TZ3_R + 200)    // ring pixels of a strip: 2R left, 2R right, 200 top / bottom; 8 floats each
#define TZ3_SCR ((2 * (TZ3_R + 2) * 3 + 102 * 3) * 8)

struct Tz3Args {
    const __nv_bfloat16 *in;          // plane layout [item][4][100][26][8] (channels 4..7 zero)
    const __nv_bfloat16 *wt;          // B operand [3 dy][3 k-steps][2 chunks][128 n][8 cin]
    const float *bias;                // [32] = bias[co] per phase
    const float *ring_w;              // un-phased fp32 weights [9][4][8]
    __nv_bfloat16 *out;               // plane layout [item][8][200][26][8]
    long long *dbg;                   // optional per-work clock stamps of CTA 5
    int legacy;
};

struct Tz3Smem {
    static constexpr unsigned off_w = 0;
    static constexpr unsigned off_a = TZ3_WBYTES;
    static constexpr unsigned off_ring = off_a + TZ3_NST * TZ3_ABYTES;             // per group
    static constexpr unsigned off_corr = off_ring + TZ3_NST * TZ3_RING * 32;        // per group
    static constexpr unsigned off_scr = off_corr + TZ3_NST * TZ3_RING * 32;         // per group: column / row dots
    static constexpr unsigned off_aux = off_scr + TZ3_NST * TZ3_SCR * 4;            // 288 ring weights
    static constexpr unsigned off_bar = off_aux + 288 * 4;
    // barriers: wbar, full_a[NST], a_empty[NST], halo[NST], halo2[NST], tmem_empty[NST], acc_full[NST][T]
    // ... ring_full[NST], ring_free[NST]
    static constexpr unsigned n_bar = 1 + 7 * TZ3_NST + TZ3_NST * TZ3_T;
    static constexpr unsigned total = off_bar + n_bar * 8 + 16;
};

struct PlaneStrip3 {
    const uint4 *pl;
    int y0;
    __device__ __forceinline__ uint4 operator()(int y, int x) const { return pl[(x & 3) * TZ3_PS + (y - y0 + 1) * TZ3_P + (x >> 2) + 1]; }
};

// Persistent and warp-specialised like k_tz_up4: warp 8 = bulk-copy producer, warps 9-10 = one MMA warp per tile (warp 9
// also fills the halo slots), warps 0-7 = two draining groups (group g owns stage g) that add bias, apply ReLU and store
// bf16 pixels in k_tz_up4's plane layout; the border ring is corrected with the separable form (column / row dots).
__global__ void __launch_bounds__(TZ3_NTP, 1)
k_tz_up3(const Tz3Args a, const int n_work) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4 *sw = reinterpret_cast<uint4 *>(smem + Tz3Smem::off_w);
    float *aux = reinterpret_cast<float *>(smem + Tz3Smem::off_aux);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Tz3Smem::off_bar);
    uint64_t *wbar = bars, *full_a = bars + 1, *a_empty = full_a + TZ3_NST, *halo = a_empty + TZ3_NST, *halo2 = halo + TZ3_NST,
             *tmem_empty = halo2 + TZ3_NST, *acc_full = tmem_empty + TZ3_NST, *ring_full = acc_full + TZ3_NST * TZ3_T,
             *ring_free = ring_full + TZ3_NST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Tz3Smem::off_bar + Tz3Smem::n_bar * 8);

    if (tid == 0) {
        mbar_init(wbar, 1);
        for (int s = 0; s < TZ3_NST; s++) {
            mbar_init(&full_a[s], 1);
            mbar_init(&a_empty[s], 1 + TZ3_T);             // the ring warp (its corrections read the strip) + the MMA warps
            mbar_init(&ring_full[s], 4);                   // the draining warps have parked their ring pixels
            mbar_init(&ring_free[s], 1);                   // the ring warp has consumed them
            mbar_init(&halo[s], 1);
            mbar_init(&halo2[s], 1);
            mbar_init(&tmem_empty[s], 4);
            for (int t = 0; t < TZ3_T; t++) mbar_init(&acc_full[s * TZ3_T + t], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4 * TZ3_NST) {
        // ------------------------------------------------------------ producer
        if (lane == 0) {
            mbar_expect_tx(wbar, TZ3_WBYTES + 288 * 4);
            bulk_g2s(sw, a.wt, TZ3_WBYTES, wbar);
            bulk_g2s(aux, a.ring_w, 288 * 4, wbar);
            int k = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, k++) {
                const int s = k % TZ3_NST, item = w / TZ3_STRIPS, y0 = (w % TZ3_STRIPS) * TZ3_R;
                if (k >= TZ3_NST) mbar_wait(&a_empty[s], (uint32_t)((k / TZ3_NST - 1) & 1));
                uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz3Smem::off_a + s * TZ3_ABYTES);
                const int ylo = max(y0 - 1, 0), yhi = min(y0 + TZ3_R, TZ3_H - 1);
                const uint32_t run = (uint32_t)(yhi - ylo + 1) * TZ3_P * 16;
                const bool top = y0 == 0, bot = y0 + TZ3_R > TZ3_H - 1;
                mbar_expect_tx(&full_a[s], 4u * (run + (top ? TZ3_P * 16 : 0) + (bot ? TZ3_P * 16 : 0)));
                const uint8_t *src = reinterpret_cast<const uint8_t *>(a.in) + (size_t)item * (4 * TZ3_H * TZ3_P * 16);
                for (int q = 0; q < 4; q++) {
                    const uint8_t *pq = src + (size_t)q * TZ3_H * TZ3_P * 16;
                    uint4 *dq = sa + q * TZ3_PS;
                    bulk_g2s(dq + (ylo - (y0 - 1)) * TZ3_P, pq + (size_t)ylo * TZ3_P * 16, run, &full_a[s]);
                    if (top) bulk_g2s(dq, pq, TZ3_P * 16, &full_a[s]);
                    if (bot) bulk_g2s(dq + (TZ3_H - (y0 - 1)) * TZ3_P, pq + (size_t)(TZ3_H - 1) * TZ3_P * 16, TZ3_P * 16, &full_a[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp > 4 * TZ3_NST && warp <= 4 * TZ3_NST + TZ3_T) {
        // ------------------------------------------------------------ halo slots + MMA issue (one converged warp per tile)
        mbar_wait(wbar, 0);
        constexpr uint32_t IDESC = instr_desc(TZ3_N);
        const uint32_t sw16 = smem_u32(sw) >> 4;
        const int t = warp - 4 * TZ3_NST - 1;
        const bool leader = elect_one();
        int k = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, k++) {
            const int s = k % TZ3_NST;
            const uint32_t par = (uint32_t)((k / TZ3_NST) & 1);
            uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz3Smem::off_a + s * TZ3_ABYTES);
            if (t == 0) {
                mbar_wait(&full_a[s], par);
                for (int i = lane; i < 2 * (TZ3_R + 2); i += 32) {
                    const int r = i >> 1;
                    if (i & 1) sa[0 * TZ3_PS + (r + 1) * TZ3_P] = sa[3 * TZ3_PS + r * TZ3_P + 25];     // x = 100 := x = 99
                    else sa[3 * TZ3_PS + r * TZ3_P] = sa[0 * TZ3_PS + r * TZ3_P + 1];                 // x = -1 := x = 0
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) { mbar_arrive(&halo[s]); mbar_arrive(&halo2[s]); }
            } else {
                mbar_wait(&halo2[s], par);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            }
            if (k >= TZ3_NST) { mbar_wait(&tmem_empty[s], par ^ 1u); tc_fence_after(); }
            const uint32_t sa16 = smem_u32(sa) >> 4;
            const uint32_t d = tmem_base + (uint32_t)((s * TZ3_T + t) * TZ3_N);
#pragma unroll
            for (int u = 0; u < 3; u++)
#pragma unroll
                for (int ks = 0; ks < 3; ks++) {
                    const int q0 = ks == 1 ? 1 : 0, o0 = ks == 2 ? 2 : 1;
                    const uint32_t lbo = ks == 1 ? (uint32_t)TZ3_PS : 3u * TZ3_PS - 1u;
                    const uint64_t ad = smem_desc(sa16 + (uint32_t)(q0 * TZ3_PS + u * TZ3_P + 128 * t + o0), lbo, 8);
                    const uint64_t bd = smem_desc(sw16 + (uint32_t)((u * 3 + ks) * 2 * TZ3_N), TZ3_N, 8);
                    if (leader) tc_mma(d, ad, bd, IDESC, (u | ks) ? 1u : 0u);
                }
            if (leader) {
                tc_commit(&acc_full[s * TZ3_T + t]);
                tc_commit(&a_empty[s]);
                if (t == 0) TZ_STAMPK(k, 6);
            }
            __syncwarp();
        }
    } else if (warp > 4 * TZ3_NST + TZ3_T) {
        // ------------------------------------------------------------ ring warps: warp (4 NST + T + 1 + s) owns the border ring
        // of stage s.  Part 1 (while the MMAs run): the out-of-range taps of every ring pixel in the separable form (see
        // k_tz_up4); part 2 (once the draining group has parked the folded values): folded - taps, ReLU, store again.
        mbar_wait(wbar, 0);                               // ring weights
        const int s = warp - (4 * TZ3_NST + TZ3_T + 1);
        const uint4 *sa = reinterpret_cast<const uint4 *>(smem + Tz3Smem::off_a + s * TZ3_ABYTES);
        const float *ring = reinterpret_cast<const float *>(smem + Tz3Smem::off_ring) + s * TZ3_RING * 8;
        float *corr = reinterpret_cast<float *>(smem + Tz3Smem::off_corr) + s * TZ3_RING * 8;
        float *cdots = reinterpret_cast<float *>(smem + Tz3Smem::off_scr) + s * TZ3_SCR, *rdots = cdots + 2 * (TZ3_R + 2) * 3 * 8;
        constexpr int Wo = 2 * TZ3_H;
        int k = s;
        for (int w = blockIdx.x + s * gridDim.x; w < n_work; w += TZ3_NST * gridDim.x, k += TZ3_NST) {
            const int item = w / TZ3_STRIPS, strip = w % TZ3_STRIPS, y0 = strip * TZ3_R;
            const uint32_t par = (uint32_t)((k / TZ3_NST) & 1);
            const int rows_valid = min(TZ3_R, TZ3_H - y0);
            const bool first = strip == 0, last = strip == TZ3_STRIPS - 1;
            const int nslots = 4 * TZ3_R + ((first || last) ? Wo : 0);
            __nv_bfloat16 *dst = a.out + (size_t)item * POL_UP3_ITEM;
            mbar_wait(&halo[s], par);
            const int dye = first ? 0 : 2, Yedge = first ? 0 : Wo - 1;
            int rlo = 0, rhi = 0; float rwl = 0.f, rwh = 0.f;
            bil_tap_ext(Yedge + dye - 1, rlo, rhi, rwl, rwh, a.legacy);
            for (int q = lane; q < 2 * (TZ3_R + 2) * 3; q += 32) {
                const int side = q / ((TZ3_R + 2) * 3), rem = q % ((TZ3_R + 2) * 3), r = rem / 3, dy = rem % 3;
                float x[8], d[8];
                unpack_bf8(sa[(side ? 3 : 0) * TZ3_PS + r * TZ3_P + (side ? 25 : 1)], x);
                const float *wp = aux + (dy * 3 + (side ? 2 : 0)) * 32;
#pragma unroll
                for (int co = 0; co < 8; co++) d[co] = 0.f;
#pragma unroll
                for (int ci = 0; ci < 4; ci++)
#pragma unroll
                    for (int co = 0; co < 8; co++) d[co] += x[ci] * wp[ci * 8 + co];
#pragma unroll
                for (int co = 0; co < 8; co++) cdots[q * 8 + co] = d[co];
            }
            if (first || last) {
                const PlaneStrip3 Ls{sa, y0};
                for (int q = lane; q < 102 * 3; q += 32) {
                    const int x = q / 3 - 1, dx = q % 3;
                    float lo[8], hi[8], d[8];
                    unpack_bf8(Ls(rlo, x), lo);
                    unpack_bf8(Ls(rhi, x), hi);
                    const float *wp = aux + (dye * 3 + dx) * 32;
#pragma unroll
                    for (int co = 0; co < 8; co++) d[co] = 0.f;
#pragma unroll
                    for (int ci = 0; ci < 4; ci++) {
                        const float u = rwl * lo[ci] + rwh * hi[ci];
#pragma unroll
                        for (int co = 0; co < 8; co++) d[co] += u * wp[ci * 8 + co];
                    }
#pragma unroll
                    for (int co = 0; co < 8; co++) rdots[q * 8 + co] = d[co];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_empty[s]);      // the strip buffer is no longer read by this warp
            for (int slot = lane; slot < nslots; slot += 32) {
                float c[8];
#pragma unroll
                for (int co = 0; co < 8; co++) c[co] = 0.f;
                if (slot < 4 * TZ3_R) {
                    const int side = slot / (2 * TZ3_R), Yl = slot % (2 * TZ3_R), Y = 2 * y0 + Yl;
                    const float *ds = cdots + side * (TZ3_R + 2) * 3 * 8;
#pragma unroll
                    for (int dy = 0; dy < 3; dy++) {
                        int lo, hi; float wlo, whi;
                        bil_tap_ext(Y + dy - 1, lo, hi, wlo, whi, a.legacy);
                        const float *dl = ds + ((lo - y0 + 1) * 3 + dy) * 8, *dh = ds + ((hi - y0 + 1) * 3 + dy) * 8;
#pragma unroll
                        for (int co = 0; co < 8; co++) c[co] += wlo * dl[co] + whi * dh[co];
                    }
                    if ((first || last) && Y == Yedge) {        // corner: the two taps of the border row outside its column
                        const int X = side ? Wo - 1 : 0;
                        for (int dx = 0; dx < 3; dx++) {
                            if (dx == (side ? 2 : 0)) continue;
                            int lo, hi; float wlo, whi;
                            bil_tap_ext(X + dx - 1, lo, hi, wlo, whi, a.legacy);
                            const float *dl = rdots + ((lo + 1) * 3 + dx) * 8, *dh = rdots + ((hi + 1) * 3 + dx) * 8;
#pragma unroll
                            for (int co = 0; co < 8; co++) c[co] += wlo * dl[co] + whi * dh[co];
                        }
                    }
                } else {
                    const int X = slot - 4 * TZ3_R;
#pragma unroll
                    for (int dx = 0; dx < 3; dx++) {
                        int lo, hi; float wlo, whi;
                        bil_tap_ext(X + dx - 1, lo, hi, wlo, whi, a.legacy);
                        const float *dl = rdots + ((lo + 1) * 3 + dx) * 8, *dh = rdots + ((hi + 1) * 3 + dx) * 8;
#pragma unroll
                        for (int co = 0; co < 8; co++) c[co] += wlo * dl[co] + whi * dh[co];
                    }
                }
#pragma unroll
                for (int co = 0; co < 8; co++) corr[slot * 8 + co] = c[co];
            }
            __syncwarp();
            mbar_wait(&ring_full[s], par);                // the draining group has parked (and stored, uncorrected) the ring pixels
            for (int slot = lane; slot < nslots; slot += 32) {
                int Y, X;
                bool valid = true;
                if (slot < 2 * TZ3_R) { Y = 2 * y0 + slot; X = 0; valid = slot < 2 * rows_valid; }
                else if (slot < 4 * TZ3_R) { Y = 2 * y0 + slot - 2 * TZ3_R; X = Wo - 1; valid = slot - 2 * TZ3_R < 2 * rows_valid; }
                else { X = slot - 4 * TZ3_R; valid = X != 0 && X != Wo - 1; Y = first ? 0 : Wo - 1; }
                if (!valid) continue;
                float o[8];
#pragma unroll
                for (int co = 0; co < 8; co++) o[co] = ring[slot * 8 + co] - corr[slot * 8 + co];
                *reinterpret_cast<uint4 *>(dst + pol_plane200_off(Y, X)) = pack_relu_bf8(o);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring_free[s]);
        }
    } else {
        // ------------------------------------------------------------ draining groups: bias + ReLU + bf16, straight to HBM in
        // k_tz_up4's plane layout.  Every pixel is stored; the pixels of the border ring are also parked (pre-activation) for
        // the stage's ring warp, which stores them again, corrected, after this group's stores (mbarrier release / acquire).
        const int grp = warp >> 2, gt = tid & 127, gw = warp & 3, s = grp;
        float *ring = reinterpret_cast<float *>(smem + Tz3Smem::off_ring) + grp * TZ3_RING * 8;
        float biasr[32];
#pragma unroll
        for (int i = 0; i < 32; i++) biasr[i] = a.bias[i];
        constexpr int Wo = 2 * TZ3_H;
        int k = grp;
        for (int w = blockIdx.x + grp * gridDim.x; w < n_work; w += TZ3_NST * gridDim.x, k += TZ3_NST) {
            const int item = w / TZ3_STRIPS, strip = w % TZ3_STRIPS, y0 = strip * TZ3_R;
            const uint32_t par = (uint32_t)((k / TZ3_NST) & 1);
            const int rows_valid = min(TZ3_R, TZ3_H - y0);
            __nv_bfloat16 *dst = a.out + (size_t)item * POL_UP3_ITEM;
            if (gt == 0) TZ_STAMPK(k, 0);
            if (k >= TZ3_NST) mbar_wait(&ring_free[s], par ^ 1u);        // the ring scratch of the previous item has been consumed
            for (int t = 0; t < TZ3_T; t++) {
                mbar_wait(&acc_full[s * TZ3_T + t], par);
                tc_fence_after();
                if (gt == 0) TZ_STAMPK(k, 2 + t);
                const int m = 128 * t + gt, yl = m / TZ3_P, xb = m - yl * TZ3_P;
                const bool valid = yl < rows_valid && xb < 25;
                const int i = y0 + yl;
                const bool edge = xb == 0 || xb == 24 || i == 0 || i == TZ3_H - 1;     // owns pixels of the border ring
                // pixel (Y = 2 i + a, X = 8 xb + 2 xo + b) lives in plane 2 xo + b at slot Y * 26 + xb + 1: one base per thread,
                // compile-time offsets per (xo, a, b)
                __nv_bfloat16 *pix = dst + ((2 * i) * 26 + xb + 1) * 8;
                const uint32_t tm = tmem_base + ((uint32_t)(gw * 32) << 16) + (uint32_t)((s * TZ3_T + t) * TZ3_N);
                uint32_t ra[32], rb[32];                  // two TMEM loads in flight: pixel xo is processed while xo + 1 arrives
                tc_ld32(tm, ra);
#pragma unroll
                for (int xo = 0; xo < 4; xo++) {          // 32 columns = the 4 phases x 8 channels of low-res pixel 4 xb + xo
                    uint32_t *r = (xo & 1) ? rb : ra;
                    tc_wait_ld();
                    if (xo < 3) tc_ld32(tm + (uint32_t)((xo + 1) * 32), (xo & 1) ? ra : rb);
                    if (!valid) continue;
#pragma unroll
                    for (int ph = 0; ph < 4; ph++) {
                        float o[8];
#pragma unroll
                        for (int co = 0; co < 8; co++) o[co] = __uint_as_float(r[ph * 8 + co]) + biasr[ph * 8 + co];
                        *reinterpret_cast<uint4 *>(pix + ((2 * xo + (ph & 1)) * 200 + (ph >> 1)) * 26 * 8) = pack_relu_bf8(o);
                        if (edge) {
                            // the row's halo slot is never read from HBM, but leaving it unwritten would make the row's first
                            // 32-byte sector a partial write (read-modify-write in DRAM)
                            if (xb == 0) *reinterpret_cast<uint4 *>(pix - 8 + ((2 * xo + (ph & 1)) * 200 + (ph >> 1)) * 26 * 8) = make_uint4(0, 0, 0, 0);
                            const int Y = 2 * i + (ph >> 1), X = 8 * xb + 2 * xo + (ph & 1);
                            if (Y == 0 || Y == Wo - 1 || X == 0 || X == Wo - 1) {
                                float *rbp = ring + 8 * (X == 0 ? Y - 2 * y0 : (X == Wo - 1 ? 2 * TZ3_R + Y - 2 * y0 : 4 * TZ3_R + X));
#pragma unroll
                                for (int co = 0; co < 8; co++) rbp[co] = o[co];
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&tmem_empty[s]); mbar_arrive(&ring_full[s]); }
            if (gt == 0) TZ_STAMPK(k, 4);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

// ---------------------------------------------------------------- host side
static long long *g_tz_dbg3 = nullptr;                   // per-work stamps of k_tz_up3 (OFB_TZ_DEBUG=up3)
extern "C" int ofb_policy_tz_debug3(long long *dev_buf) { g_tz_dbg3 = dev_buf; return OFB_OK; }
static long long *g_tz_dbg = nullptr;                    // device buffer of 16 stamps, see ofb_policy_tz_debug
extern "C" int ofb_policy_tz_debug(long long *dev_buf) { g_tz_dbg = dev_buf; return OFB_OK; }

int pol_tz_up4(const ofb_policy *p, const __nv_bfloat16 *in, float *ptr_out, float *amax_val, int *amax_idx, int n_items,
               cudaStream_t st) {
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_tz_up4, (int)Tz4Smem::total));
    int n_sm = 148;
    OFB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
    if (n_items == 0) return OFB_OK;
    Tz4Args a = {};
    a.in = in; a.wt = p->w.u4_tz; a.ring_w = p->w.u4_w; a.bias = p->u4_bias;
    a.ptr_out = ptr_out; a.amax_val = amax_val; a.amax_idx = amax_idx; a.legacy = p->w.bil_legacy;
    a.dbg = g_tz_dbg;
    const int n_work = n_items * TZ4_STRIPS;
    k_tz_up4<<<n_work < n_sm ? n_work : n_sm, TZ4_NTP, Tz4Smem::total, st>>>(a, n_work);     // one persistent CTA per SM
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

int pol_tz_up4_parts() { return TZ4_STRIPS; }

int pol_tz_up3(const ofb_policy *p, const __nv_bfloat16 *in, __nv_bfloat16 *out, int n_items, cudaStream_t st) {
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_tz_up3, (int)Tz3Smem::total));
    int n_sm = 148;
    OFB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
    if (n_items == 0) return OFB_OK;
    Tz3Args a = {};
    a.in = in; a.wt = p->w.u3_tz; a.bias = p->w.u3_pb; a.ring_w = p->w.u3_w; a.out = out; a.legacy = p->w.bil_legacy;
    a.dbg = g_tz_dbg3;
    const int n_work = n_items * TZ3_STRIPS;
    k_tz_up3<<<n_work < n_sm ? n_work : n_sm, TZ3_NTP, Tz3Smem::total, st>>>(a, n_work);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// ================================================================================================
// k_tz_trunk12: conv1 + BN + ReLU + pool (from the bit maps) -> conv2 + BN + ReLU + pool (qlearnIA_V2.py:129-139),
// 400 x 400 x 2 bits -> 100 x 100 x 8, persistent and warp-specialised like k_tz_up4.
//
// conv1's pooled output is a constant vector ("background") wherever the 4 x 4 map patch under a pooled pixel holds no
// set bit -- ~96 % of an arena -- so the strip buffers are kept FILLED with the background between work items: the
// producer warps only scan the strip's bit rows, evaluate the dirty pixels exactly (9-bit stencil pattern LUT, as in the
// pixel-linear kernel) and put the background back after the MMAs have read the strip.  conv2 is the block-Toeplitz GEMM
// (N = 8 px x 8 channels = 64, 15 MMAs per tile); the draining warps pool along x in registers (their 8 pixels are 4
// pooling pairs) and along y through a shared-memory stage.
// ================================================================================================
#define TZ2_H 200                     // conv2 grid
#define TZ2_P 26
#define TZ2_R 14                      // conv2 rows per strip (even: 7 pooled rows); 14 * 26 = 364 M rows = 3 tiles
#define TZ2_T 3
#define TZ2_PS 440
#define TZ2_N 64
#define TZ2_NST 2                     // stages = draining groups
#define TZ2_STRIPS ((TZ2_H + TZ2_R - 1) / TZ2_R)
#define TZ2_NPW 8                     // producer warps: two groups of 4, group g owns stage g (every other work item)
#define TZ2_NPT 128                   // threads of one producer group
#define TZ2_NTP (32 * (4 * TZ2_NST + TZ2_NPW + TZ2_T))      // 8 draining + 8 producer + 3 MMA warps
#define TZ2_ABYTES (8 * TZ2_PS * 16)
#define TZ2_WBYTES (3 * 5 * 2 * TZ2_N * 16)
#define TZ2_BITW 456                  // words of one staged bit map: (2 R + 6) rows x 50 B + alignment slack
#define TZ2_WL ((TZ2_R + 2) * TZ2_H)  // dirty-pixel list capacity = every pixel of the strip
#define TZ2_STG 384                   // M rows of the pooling stage

struct Tz2Args {
    const uint32_t *maps;             // [item][2][5000]
    const __nv_bfloat16 *wt;          // conv2 block-Toeplitz B operand [3][5][2][64][8]
    const float *bias;                // conv2 bias [8]
    const float *c1_lut, *c1_b;       // conv1 pattern LUT [2][512][8], bias [8]
    __nv_bfloat16 *out;               // pool2 NHWC [item][100][100][8]
    long long *dbg;
};

struct Tz2Smem {
    static constexpr unsigned off_w = 0;
    static constexpr unsigned off_a = TZ2_WBYTES;
    static constexpr unsigned off_bits = off_a + TZ2_NST * TZ2_ABYTES;                 // [stage][2 maps][BITW]
    static constexpr unsigned off_wl = off_bits + TZ2_NST * 2 * 2 * TZ2_BITW * 4;       // [stage][WL] ushort   (bits: [stage][2 buffers][2 maps][BITW])
    static constexpr unsigned off_stg = (off_wl + TZ2_NST * TZ2_WL * 2 + 15) & ~15u;    // [group][STG][4] uint4
    static constexpr unsigned off_misc = off_stg + TZ2_NST * TZ2_STG * 64;              // counters, biases
    static constexpr unsigned off_bar = off_misc + 128;
    // barriers: wbar, bits_full[NST][2 buffers] (the second half of the array), full_a[NST], a_empty[NST], tmem_empty[NST], acc_full[NST][T]
    static constexpr unsigned n_bar = 1 + 5 * TZ2_NST + TZ2_NST * TZ2_T;
    static constexpr unsigned total = off_bar + n_bar * 8 + 16;
};

__device__ __forceinline__ void named_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(TZ2_NTP, 1)
k_tz_trunk12(const Tz2Args a, const int n_work) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint4 *sw = reinterpret_cast<uint4 *>(smem + Tz2Smem::off_w);
    int *misc = reinterpret_cast<int *>(smem + Tz2Smem::off_misc);          // [0..1] dirty counts per stage
    float *sbias = reinterpret_cast<float *>(misc + 8);                      // conv1 bias [8], conv2 bias [8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Tz2Smem::off_bar);
    uint64_t *wbar = bars, *bits_full = bars + 1, *full_a = bits_full + TZ2_NST, *a_empty = full_a + TZ2_NST,
             *tmem_empty = a_empty + TZ2_NST, *acc_full = tmem_empty + TZ2_NST, *bits_full2 = acc_full + TZ2_NST * TZ2_T;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Tz2Smem::off_bar + Tz2Smem::n_bar * 8);

    if (tid == 0) {
        mbar_init(wbar, 1);
        for (int s = 0; s < TZ2_NST; s++) {
            mbar_init(&bits_full[s], 1);
            mbar_init(&bits_full2[s], 1);
            mbar_init(&full_a[s], 4);                      // the 4 warps of the stage's producer group
            mbar_init(&a_empty[s], TZ2_T);                 // the MMA warps (tcgen05.commit)
            mbar_init(&tmem_empty[s], 4);                  // the 4 warps of the draining group
            for (int t = 0; t < TZ2_T; t++) mbar_init(&acc_full[s * TZ2_T + t], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 16) sbias[tid] = tid < 8 ? a.c1_b[tid] : a.bias[tid - 8];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 4 * TZ2_NST && warp < 4 * TZ2_NST + TZ2_NPW) {
        // ------------------------------------------------------------ producer groups: group g = 4 warps owns stage g
        const int pg = (warp - 4 * TZ2_NST) >> 2, pt = (tid - 128 * TZ2_NST) & 127, s = pg;
        float bg[8];
#pragma unroll
        for (int co = 0; co < 8; co++) bg[co] = fmaxf(sbias[co], 0.f);
        const uint4 bgq = pack_bf8(bg), zq = make_uint4(0, 0, 0, 0);
        uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz2Smem::off_a + s * TZ2_ABYTES);
        unsigned short *wl = reinterpret_cast<unsigned short *>(smem + Tz2Smem::off_wl) + s * TZ2_WL;
        // the stage starts clean: background everywhere, zeros in the x halo slots (slot 0 of planes 0 and 7)
        for (int i = pt; i < 8 * TZ2_PS; i += TZ2_NPT) {
            const int q = i / TZ2_PS, t = (i % TZ2_PS) % TZ2_P;
            sa[i] = ((q == 0 || q == 7) && t == 0) ? zq : bgq;
        }
        auto issue_bits = [&](int j, int w) {               // one thread: stage the bit rows of the group's j-th work item w
            const int item = w / TZ2_STRIPS, y0 = (w % TZ2_STRIPS) * TZ2_R;
            const int r0 = max(2 * y0 - 3, 0), r1 = min(2 * (y0 + TZ2_R) + 2, POL_W - 1);
            const int b0 = (r0 * 50) & ~15, b1 = min(((r1 + 1) * 50 + 15 + 16) & ~15, POL_WORDS * 4);
            uint32_t *sb = reinterpret_cast<uint32_t *>(smem + Tz2Smem::off_bits) + (s * 2 + (j & 1)) * 2 * TZ2_BITW;
            uint64_t *bar = (j & 1) ? &bits_full2[s] : &bits_full[s];
            const uint8_t *src = reinterpret_cast<const uint8_t *>(a.maps) + (size_t)item * 2 * POL_WORDS * 4;
            mbar_expect_tx(bar, 2u * (uint32_t)(b1 - b0));
            bulk_g2s(sb, src + b0, (uint32_t)(b1 - b0), bar);
            bulk_g2s(sb + TZ2_BITW, src + POL_WORDS * 4 + b0, (uint32_t)(b1 - b0), bar);
        };
        const int wstep = TZ2_NST * gridDim.x;
        if (pt == 0) {
            if (pg == 0) {
                mbar_expect_tx(wbar, TZ2_WBYTES);
                bulk_g2s(sw, a.wt, TZ2_WBYTES, wbar);
            }
            if (blockIdx.x + pg * gridDim.x < n_work) issue_bits(0, blockIdx.x + pg * gridDim.x);
        }
        named_sync_n(1 + pg, TZ2_NPT);
        int j = 0, prev_y0 = -1000;                         // j = index among the group's work items
        for (int w = blockIdx.x + pg * gridDim.x; w < n_work; w += wstep, j++) {
            const int y0 = (w % TZ2_STRIPS) * TZ2_R;
            const uint32_t par = (uint32_t)(j & 1);         // stage s is used once per group item: its barriers flip every item
            const uint32_t *sb = reinterpret_cast<const uint32_t *>(smem + Tz2Smem::off_bits) + (s * 2 + (j & 1)) * 2 * TZ2_BITW;
            if (j >= 1) {
                // ---- the MMAs of the previous item in this stage are done: put the background back
                mbar_wait(&a_empty[s], par ^ 1u);
                const int n_old = misc[s];
                for (int i = pt; i < n_old; i += TZ2_NPT) {
                    const int e = wl[i], r = e / TZ2_H, px = e % TZ2_H;
                    sa[(px & 7) * TZ2_PS + r * TZ2_P + (px >> 3) + 1] = bgq;
                }
                const int zr = prev_y0 == 0 ? 0 : (prev_y0 + TZ2_R >= TZ2_H ? TZ2_H - (prev_y0 - 1) : -1);
                if (zr >= 0)
                    for (int i = pt; i < 8 * 25; i += TZ2_NPT) sa[(i / 25) * TZ2_PS + zr * TZ2_P + (i % 25) + 1] = bgq;
            }
            prev_y0 = y0;
            named_sync_n(1 + pg, TZ2_NPT);                  // every thread of the group has left the previous item
            if (pt == 0) {
                misc[s] = 0;
                if (w + wstep < n_work) issue_bits(j + 1, w + wstep);                        // prefetch the group's next bit rows
            }
            mbar_wait((j & 1) ? &bits_full2[s] : &bits_full[s], (uint32_t)((j >> 1) & 1));
            named_sync_n(1 + pg, TZ2_NPT);
            // ---- scan: pooled pixels whose 4 x 4 map patch holds a set bit
            const int r0 = max(2 * y0 - 3, 0), bits_w0 = ((r0 * 50) & ~15) >> 2;
            const uint32_t *smap = sb - bits_w0, *lmap = sb + TZ2_BITW - bits_w0;
            constexpr int groups = TZ2_H / 8;
            for (int g = pt; g < (TZ2_R + 2) * groups; g += TZ2_NPT) {
                const int r = g / groups, gx = g % groups, py = y0 - 1 + r;
                if (py < 0 || py >= TZ2_H) continue;
                // 18 map columns 16 gx - 1 .. 16 gx + 16 of the 4 map rows 2 py - 1 .. 2 py + 2
                uint32_t rs[4], any = 0, colmask = 0x3FFFFu;
                if (gx == 0) colmask &= ~1u;
                if (gx == groups - 1) colmask &= ~(1u << 17);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int rr = 2 * py - 1 + i;
                    rs[i] = 0;
                    if (rr >= 0 && rr < POL_W) {
                        int b = rr * POL_W + 16 * gx - 1;
                        const int sh = b < 0 ? 1 : 0;
                        b = max(b, 0);
                        const int wd = b >> 5;
                        rs[i] = ((__funnelshift_r(smap[wd], smap[wd + 1], b & 31) | __funnelshift_r(lmap[wd], lmap[wd + 1], b & 31)) << sh) & colmask;
                    }
                    any |= rs[i];
                }
                if (!any) continue;
                const uint32_t orr = rs[0] | rs[1] | rs[2] | rs[3];
#pragma unroll
                for (int kk = 0; kk < 8; kk++)
                    if ((orr >> (2 * kk)) & 0xFu) wl[atomicAdd(&misc[s], 1)] = (unsigned short)(r * TZ2_H + gx * 8 + kk);
            }
            named_sync_n(1 + pg, TZ2_NPT);
            // ---- exact conv1 + pool for the dirty pixels; zero rows outside the image (conv2's padding)
            {
                const int n_new = misc[s];
                for (int i = pt; i < n_new; i += TZ2_NPT) {
                    const int e = wl[i], r = e / TZ2_H, px = e % TZ2_H, py = y0 - 1 + r;
                    float v[8];
                    conv1_pool_pixel(conv1_patch(smap, py, px), conv1_patch(lmap, py, px), a.c1_lut, sbias, v);
                    sa[(px & 7) * TZ2_PS + r * TZ2_P + (px >> 3) + 1] = pack_bf8(v);
                }
                const int zr = y0 == 0 ? 0 : (y0 + TZ2_R >= TZ2_H ? TZ2_H - (y0 - 1) : -1);
                if (zr >= 0)
                    for (int i = pt; i < 8 * 25; i += TZ2_NPT) sa[(i / 25) * TZ2_PS + zr * TZ2_P + (i % 25) + 1] = zq;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[s]);
        }
    } else if (warp >= 4 * TZ2_NST + TZ2_NPW) {
        // ------------------------------------------------------------ MMA warps: one tile each, converged, elected lane issues
        mbar_wait(wbar, 0);
        constexpr uint32_t IDESC = instr_desc(TZ2_N);
        const uint32_t sw16 = smem_u32(sw) >> 4;
        const int t = warp - (4 * TZ2_NST + TZ2_NPW);
        const bool leader = elect_one();
        int k = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, k++) {
            const int s = k & 1;
            const uint32_t par = (uint32_t)((k >> 1) & 1);
            mbar_wait(&full_a[s], par);
            if (k >= 2) mbar_wait(&tmem_empty[s], par ^ 1u);
            tc_fence_after();
            const uint32_t sa16 = smem_u32(smem + Tz2Smem::off_a + s * TZ2_ABYTES) >> 4;
            const uint32_t d = tmem_base + (uint32_t)((s * TZ2_T + t) * TZ2_N);
#pragma unroll
            for (int u = 0; u < 3; u++)
#pragma unroll
                for (int ks = 0; ks < 5; ks++) {
                    const int q0 = ks == 0 || ks == 4 ? 0 : 2 * ks - 1, o0 = ks == 4 ? 2 : 1;
                    const uint32_t lbo = (ks == 0 || ks == 4) ? 7u * TZ2_PS - 1u : (uint32_t)TZ2_PS;
                    const uint64_t ad = smem_desc(sa16 + (uint32_t)(q0 * TZ2_PS + u * TZ2_P + 128 * t + o0), lbo, 8);
                    const uint64_t bd = smem_desc(sw16 + (uint32_t)((u * 5 + ks) * 2 * TZ2_N), TZ2_N, 8);
                    if (leader) tc_mma(d, ad, bd, IDESC, (u | ks) ? 1u : 0u);
                }
            if (leader) {
                tc_commit(&acc_full[s * TZ2_T + t]);
                tc_commit(&a_empty[s]);
            }
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------ draining groups: bias + ReLU + 2 x 2 max-pool
        const int grp = warp >> 2, gt = tid & 127, gw = warp & 3, s = grp;
        uint4 *stg = reinterpret_cast<uint4 *>(smem + Tz2Smem::off_stg) + grp * TZ2_STG * 4;
        float b2[8];
#pragma unroll
        for (int co = 0; co < 8; co++) b2[co] = sbias[8 + co];
        int k = grp;
        for (int w = blockIdx.x + grp * gridDim.x; w < n_work; w += TZ2_NST * gridDim.x, k += TZ2_NST) {
            const int item = w / TZ2_STRIPS, y0 = (w % TZ2_STRIPS) * TZ2_R;
            const uint32_t par = (uint32_t)((k >> 1) & 1);
            const int rows_valid = min(TZ2_R, TZ2_H - y0);
            for (int t = 0; t < TZ2_T; t++) {
                mbar_wait(&acc_full[s * TZ2_T + t], par);
                tc_fence_after();
                const int m = 128 * t + gt, yl = m / TZ2_P, xb = m - yl * TZ2_P;
                const bool valid = yl < rows_valid && xb < 25;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {            // 32 columns = pixels 4 hh .. 4 hh + 3 of the block, 8 channels each
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((uint32_t)(gw * 32) << 16) + (uint32_t)((s * TZ2_T + t) * TZ2_N + hh * 32), r);
                    tc_wait_ld();
                    if (!valid) continue;
#pragma unroll
                    for (int j = 0; j < 2; j++) {           // pooling pair (x even, x odd): relu(max(.) + bias), then bf16
                        float o[8];
#pragma unroll
                        for (int co = 0; co < 8; co++)
                            o[co] = fmaxf(fmaxf(__uint_as_float(r[(2 * j) * 8 + co]), __uint_as_float(r[(2 * j + 1) * 8 + co])) + b2[co], 0.f);
                        stg[m * 4 + hh * 2 + j] = pack_bf8(o);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[s]);
            named_sync_n(3 + grp, 128);
            // ---- pooling along y and the NHWC store of the strip's rows_valid / 2 output rows
            __nv_bfloat16 *dst = a.out + (size_t)item * (100 * 100 * 8) + (size_t)(y0 / 2) * 100 * 8;
            for (int idx = gt; idx < (rows_valid / 2) * 100; idx += 128) {
                const int pr = idx / 100, X = idx % 100, xb = X >> 2, j = X & 3;
                const uint4 q0 = stg[((2 * pr) * TZ2_P + xb) * 4 + j], q1 = stg[((2 * pr + 1) * TZ2_P + xb) * 4 + j];
                uint4 o;
                const uint32_t *p0 = &q0.x, *p1 = &q1.x;
                uint32_t *po = &o.x;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    __nv_bfloat162 mm = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(p0 + c), *reinterpret_cast<const __nv_bfloat162 *>(p1 + c));
                    po[c] = *reinterpret_cast<uint32_t *>(&mm);
                }
                *reinterpret_cast<uint4 *>(dst + (size_t)idx * 8) = o;
            }
            named_sync_n(3 + grp, 128);                    // the stage may be overwritten by the group's next item
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
}

int pol_tz_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st) {
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_tz_trunk12, (int)Tz2Smem::total));
    int n_sm = 148;
    OFB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
    if (n_items == 0) return OFB_OK;
    Tz2Args a = {};
    a.maps = maps; a.wt = p->w.c2_tz; a.bias = p->w.cb[0]; a.c1_lut = p->w.c1_lut; a.c1_b = p->w.c1_b; a.out = out;
    a.dbg = g_tz_dbg;
    const int n_work = n_items * TZ2_STRIPS;
    k_tz_trunk12<<<n_work < n_sm ? n_work : n_sm, TZ2_NTP, Tz2Smem::total, st>>>(a, n_work);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// ================================================================================================
// k_tc_dense1: the flat slice of dense1 (qlearnIA_V2.py:154-155) as a plain tcgen05 GEMM,
//     hflat[arena][n] = sum_k flat[arena][k] * W[n][k],   M = 128 arenas per CTA, N = 128 (100 used), K = 5120.
// Both operands are K-major in HBM ([row][5120] bf16).  Eight loader warps move K = 64 slabs into shared memory with
// coalesced 16-byte loads, storing them as UMMA core matrices [k/8][row][16 B] (LBO = one k-chunk plane, padded by 16 B
// so that the 8 chunks of a row fall into different banks); one converged warp issues 4 MMAs per slab.
// ================================================================================================
#define D1_K POL_FLAT_PITCH           // 5120
#define D1_SLAB 64                    // K per stage
#define D1_NSLAB (D1_K / D1_SLAB)     // 80
#define D1_PLANE (128 * 16 + 16)      // bytes of one k-chunk plane (128 rows x 16 B + pad)
#define D1_OPBYTES (8 * D1_PLANE)     // one operand slab
#define D1_NST 4
#define D1_NT (32 * 9)                // 8 loader warps + 1 MMA warp

__global__ void __launch_bounds__(D1_NT, 1)
k_tc_dense1(const __nv_bfloat16 *__restrict__ flat, const __nv_bfloat16 *__restrict__ wt, float *__restrict__ hflat, int n_items) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + D1_NST * 2 * D1_OPBYTES);     // full[NST], empty[NST], done
    uint64_t *full = bars, *empty = bars + D1_NST, *done = bars + 2 * D1_NST;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * D1_NST + 1);
    const int a0 = blockIdx.x * 128;
    if (tid == 0) {
        for (int s = 0; s < D1_NST; s++) { mbar_init(&full[s], 8); mbar_init(&empty[s], 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ---- loaders: thread -> (row, k-chunk) pieces of 16 B; 8 consecutive threads read 128 contiguous bytes of one row
        const int kc = tid & 7, r0 = tid >> 3;             // rows r0, r0 + 32, r0 + 64, r0 + 96
        for (int sl = 0; sl < D1_NSLAB; sl++) {
            const int s = sl % D1_NST;
            if (sl >= D1_NST) mbar_wait(&empty[s], (uint32_t)((sl / D1_NST - 1) & 1));
            uint8_t *sA = smem + s * 2 * D1_OPBYTES, *sB = sA + D1_OPBYTES;
            uint4 va[4], vb[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int row = r0 + 32 * i;
                va[i] = (a0 + row < n_items) ? *reinterpret_cast<const uint4 *>(flat + (size_t)(a0 + row) * D1_K + sl * D1_SLAB + kc * 8)
                                             : make_uint4(0, 0, 0, 0);
                vb[i] = *reinterpret_cast<const uint4 *>(wt + (size_t)row * D1_K + sl * D1_SLAB + kc * 8);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int row = r0 + 32 * i;
                *reinterpret_cast<uint4 *>(sA + kc * D1_PLANE + row * 16) = va[i];
                *reinterpret_cast<uint4 *>(sB + kc * D1_PLANE + row * 16) = vb[i];
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
        // ---- epilogue: warps 0-3 read the accumulator (lane quarter = warp), thread = arena
        if (warp < 4) {
            mbar_wait(done, 0);
            tc_fence_after();
            const int arena = a0 + tid;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), r);
                tc_wait_ld();
                if (arena < n_items)
#pragma unroll
                    for (int n = 0; n < 32; n++)
                        if (c * 32 + n < 100) hflat[(size_t)arena * 100 + c * 32 + n] = __uint_as_float(r[n]);
            }
        }
    } else {
        // ---- MMA warp (converged, elected lane issues)
        constexpr uint32_t IDESC = instr_desc(128);
        const bool leader = elect_one();
        for (int sl = 0; sl < D1_NSLAB; sl++) {
            const int s = sl % D1_NST;
            mbar_wait(&full[s], (uint32_t)((sl / D1_NST) & 1));
            tc_fence_after();
            const uint32_t a16 = smem_u32(smem + s * 2 * D1_OPBYTES) >> 4, b16 = a16 + D1_OPBYTES / 16;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t ad = smem_desc(a16 + (uint32_t)(2 * j * (D1_PLANE / 16)), D1_PLANE / 16, 8);
                const uint64_t bd = smem_desc(b16 + (uint32_t)(2 * j * (D1_PLANE / 16)), D1_PLANE / 16, 8);
                if (leader) tc_mma(tmem_base, ad, bd, IDESC, (sl | j) ? 1u : 0u);
            }
            if (leader) tc_commit(&empty[s]);
        }
        if (leader) tc_commit(done);
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

int pol_tc_dense1(const ofb_policy *p, const __nv_bfloat16 *flat, float *hflat, int n_items, cudaStream_t st) {
    const int smem = D1_NST * 2 * D1_OPBYTES + (2 * D1_NST + 1) * 8 + 16;
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_tc_dense1, smem));
    if (n_items == 0) return OFB_OK;
    k_tc_dense1<<<(n_items + 127) / 128, D1_NT, smem, st>>>(flat, p->w.d1_wt, hflat, n_items);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

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
#define TZ4_R 19                      // rows per strip: 19 * 26 = 494 M rows = 4 tiles of 128
#define TZ4_T 4
#define TZ4_PS 576                    // slots per plane in shared memory: (R + 2) * P + 2 + tile overrun, rounded to 8
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
};

struct Tz4Smem {
    static constexpr unsigned off_w = 0;                                   // 3 * 5 * 2 * 32 * 16 B
    static constexpr unsigned off_a = 15360;
    static constexpr unsigned off_ring = off_a + 8 * TZ4_PS * 16;          // floats: 2R left, 2R right, 400 top/bottom
    static constexpr unsigned off_aux = off_ring + (4 * TZ4_R + 400) * 4;  // 72 ring weights + argmax scratch
    static constexpr unsigned off_bar = (off_aux + (72 + 16) * 4 + 15) & ~15u;
    static constexpr unsigned total = off_bar + (TZ4_T + 1) * 8 + 16;
};

// plane-layout strip in shared memory as an image accessor: rows y0-1 .. y0+R, cols -1 .. 200 (halos filled)
struct PlaneStrip4 {
    const uint4 *pl;
    int y0;
    __device__ __forceinline__ uint4 operator()(int y, int x) const { return pl[(x & 7) * TZ4_PS + (y - y0 + 1) * TZ4_P + (x >> 3) + 1]; }
};

__global__ void __launch_bounds__(TZ4_NT)
k_tz_up4(const Tz4Args a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.y, y0 = blockIdx.x * TZ4_R;
    const int rows_valid = min(TZ4_R, TZ4_H - y0);
    uint4 *sw = reinterpret_cast<uint4 *>(smem + Tz4Smem::off_w);
    uint4 *sa = reinterpret_cast<uint4 *>(smem + Tz4Smem::off_a);
    float *ring = reinterpret_cast<float *>(smem + Tz4Smem::off_ring);
    float *aux = reinterpret_cast<float *>(smem + Tz4Smem::off_aux);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + Tz4Smem::off_bar);      // full[T], load
    uint64_t *lbar = &bars[TZ4_T];
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Tz4Smem::off_bar + (TZ4_T + 1) * 8);
#define TZ_STAMP(k, who) do { if (a.dbg && tid == (who) && blockIdx.x == 5 && blockIdx.y == 0) a.dbg[k] = clock64(); } while (0)
    TZ_STAMP(0, 0);

    // ---- producer: barriers + bulk copies of the strip's rows, one run of rows per plane (+ a replicated row at the
    //      top / bottom of the image)
    if (tid == 128) {
        for (int t = 0; t <= TZ4_T; t++) mbar_init(&bars[t], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int ylo = max(y0 - 1, 0), yhi = min(y0 + TZ4_R, TZ4_H - 1);
        const uint32_t run = (uint32_t)(yhi - ylo + 1) * TZ4_P * 16;
        const bool top = y0 == 0, bot = y0 + TZ4_R > TZ4_H - 1;                // a halo row outside the image is needed
        mbar_expect_tx(lbar, 8u * (run + (top ? TZ4_P * 16 : 0) + (bot ? TZ4_P * 16 : 0)) + 3 * 5 * 2 * TZ4_N * 16 + 72 * 4);
        bulk_g2s(sw, a.wt, 3 * 5 * 2 * TZ4_N * 16, lbar);                      // weights ride the same barrier
        bulk_g2s(aux, a.ring_w, 72 * 4, lbar);
        const uint8_t *src = reinterpret_cast<const uint8_t *>(a.in) + (size_t)item * (8 * TZ4_H * TZ4_P * 16);
        for (int q = 0; q < 8; q++) {
            const uint8_t *pq = src + (size_t)q * TZ4_H * TZ4_P * 16;
            uint4 *dq = sa + q * TZ4_PS;
            bulk_g2s(dq + (ylo - (y0 - 1)) * TZ4_P, pq + (size_t)ylo * TZ4_P * 16, run, lbar);
            if (top) bulk_g2s(dq, pq, TZ4_P * 16, lbar);                                                  // row -1 := row 0
            if (bot) bulk_g2s(dq + (TZ4_H - (y0 - 1)) * TZ4_P, pq + (size_t)(TZ4_H - 1) * TZ4_P * 16, TZ4_P * 16, lbar);   // row 200 := row 199
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    __syncthreads();                                     // barrier init visible
    TZ_STAMP(1, 0);
    mbar_wait(lbar, 0);                                  // rows have landed (acquire for every thread)
    TZ_STAMP(2, 0);
    // ---- halos: x = -1 := x = 0 (plane 7 slot 0 of the row), x = 200 := x = 199 (plane 0 slot 0 of the next row)
    for (int i = tid; i < 2 * (TZ4_R + 2); i += TZ4_NT) {
        const int r = i >> 1;
        if (i & 1) sa[0 * TZ4_PS + (r + 1) * TZ4_P] = sa[7 * TZ4_PS + r * TZ4_P + 25];
        else sa[7 * TZ4_PS + r * TZ4_P] = sa[0 * TZ4_PS + r * TZ4_P + 1];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    TZ_STAMP(3, 0);

    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    if (warp == 4) {
        // ---- MMA issuer: 4 tiles x (3 vertical taps x 5 K-steps), every tile has its own TMEM accumulator
        if (lane == 0) {
            constexpr uint32_t IDESC = instr_desc(TZ4_N);
            const uint32_t sa16 = smem_u32(sa) >> 4, sw16 = smem_u32(sw) >> 4;
            for (int t = 0; t < TZ4_T; t++) {
                const uint32_t d = tmem_base + (uint32_t)(t * TZ4_N);
#pragma unroll
                for (int u = 0; u < 3; u++)
#pragma unroll
                    for (int ks = 0; ks < 5; ks++) {
                        // first / second K-chunk of the step: (plane, slot offset); see the header comment
                        const int q0 = ks == 0 || ks == 4 ? 0 : 2 * ks - 1, o0 = ks == 4 ? 2 : 1;
                        const uint32_t lbo = (ks == 0 || ks == 4) ? 7u * TZ4_PS - 1u : (uint32_t)TZ4_PS;
                        const uint64_t ad = smem_desc(sa16 + (uint32_t)(q0 * TZ4_PS + u * TZ4_P + 128 * t + o0), lbo, 8);
                        const uint64_t bd = smem_desc(sw16 + (uint32_t)((u * 5 + ks) * 2 * TZ4_N), TZ4_N, 8);
                        tc_mma(d, ad, bd, IDESC, (u | ks) ? 1u : 0u);
                    }
                tc_commit(&bars[t]);
            }
            TZ_STAMP(4, 128);
        }
        __syncwarp();
    } else {
        // ---- drain: thread = M row (y_local, block) of the tile; 32 columns = 8 pixels x 4 phases
        for (int t = 0; t < TZ4_T; t++) {
            mbar_wait(&bars[t], 0);
            tc_fence_after();
            TZ_STAMP(5 + t, 0);
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t * TZ4_N), r);
            tc_wait_ld();
            const int m = 128 * t + tid, yl = m / TZ4_P, xb = m - yl * TZ4_P;
            if (yl >= rows_valid || xb >= 25) continue;
            const int i = y0 + yl;
            float v[32];
#pragma unroll
            for (int n = 0; n < 32; n++) v[n] = __uint_as_float(r[n]) + a.bias;
            // ring pixels (hi-res border) leave through the ring buffer and are corrected after the tile loop; here they
            // are masked out with -inf so that every thread takes the same branch-free path.  Column slots own the corners.
            const float NINF = -INFINITY;
            if (xb == 0) { ring[2 * yl] = v[0]; ring[2 * yl + 1] = v[2]; v[0] = NINF; v[2] = NINF; }
            if (xb == 24) { ring[2 * TZ4_R + 2 * yl] = v[29]; ring[2 * TZ4_R + 2 * yl + 1] = v[31]; v[29] = NINF; v[31] = NINF; }
            if (i == 0 || i == TZ4_H - 1) {
                const int ph_a = i == 0 ? 0 : 1;
#pragma unroll
                for (int rem = 0; rem < 16; rem++) {
                    const int X = 16 * xb + rem;
                    if (X == 0 || X == 2 * TZ4_H - 1) continue;
                    if (ph_a == 0) { ring[4 * TZ4_R + X] = v[(rem >> 1) * 4 + (rem & 1)]; v[(rem >> 1) * 4 + (rem & 1)] = NINF; }
                    else { ring[4 * TZ4_R + X] = v[(rem >> 1) * 4 + 2 + (rem & 1)]; v[(rem >> 1) * 4 + 2 + (rem & 1)] = NINF; }
                }
            }
            if (a.ptr_out) {                              // optional dense map (predict()'s second output); ring pixels follow later
#pragma unroll
                for (int n = 0; n < 32; n++)
                    if (v[n] != NINF)
                        a.ptr_out[(size_t)item * 4 * TZ4_H * TZ4_H + (2 * i + ((n >> 1) & 1)) * 2 * TZ4_H + 16 * xb + 2 * (n >> 2) + (n & 1)] = v[n];
            }
            float mx = v[0];
#pragma unroll
            for (int n = 1; n < 32; n++) mx = fmaxf(mx, v[n]);
            if (mx >= best_v) {                           // rare after the first tiles: locate the first maximum in C order
                int idx = 0x7fffffff;
#pragma unroll
                for (int n = 31; n >= 0; n--) {           // descending C order within the thread: the last match is the lowest index
                    const int ph_a = (n >> 4) & 1, rem = n & 15;                 // order: a, then xo, then b
                    const int nn = (rem >> 1) * 4 + ph_a * 2 + (rem & 1);
                    if (v[nn] == mx) idx = (2 * i + ph_a) * 2 * TZ4_H + 16 * xb + rem;
                }
                if (amax_better(mx, idx, best_v, best_i)) { best_v = mx; best_i = idx; }
            }
        }
    }
    TZ_STAMP(9, 0);
    tc_fence_before();
    __syncthreads();                                     // ring values are visible to every thread
    TZ_STAMP(10, 0);

    // ---- border ring: take the out-of-range taps back out of the folded result (4 lanes per ring pixel: lanes 0..2
    //      evaluate one out-of-range tap each, lane 0 also the two extra taps of a corner), then consume the pixel
    {
        constexpr int Wo = 2 * TZ4_H;
        const bool first = blockIdx.x == 0, last = blockIdx.x == gridDim.x - 1;
        const int nslots = 4 * TZ4_R + ((first || last) ? Wo : 0);             // only the first / last strip own a border row
        const PlaneStrip4 Ls{sa, y0};
        for (int base = 0; base < nslots * 4; base += TZ4_NT) {
            const int itm = base + tid, slot = itm >> 2, j = itm & 3;
            int Y = 0, X = 0;
            bool valid = slot < nslots;
            if (slot < 2 * TZ4_R) { Y = 2 * y0 + slot; X = 0; valid = valid && slot < 2 * rows_valid; }
            else if (slot < 4 * TZ4_R) { Y = 2 * y0 + slot - 2 * TZ4_R; X = Wo - 1; valid = valid && slot - 2 * TZ4_R < 2 * rows_valid; }
            else {
                X = slot - 4 * TZ4_R;
                valid = valid && (first || last) && X != 0 && X != Wo - 1;
                Y = first ? 0 : Wo - 1;
            }
            // a strip that is both first and last does not exist (11 strips); the top row belongs to the first, the bottom to the last
            float acc = 0.f;
            if (valid && j < 3) {
                const bool xedge = (X == 0 || X == Wo - 1), yedge = (Y == 0 || Y == Wo - 1);
                const int dxe = X == 0 ? 0 : 2, dye = Y == 0 ? 0 : 2;
                if (xedge) up_ring_tap<8, 1>(Ls, Y, X, j, dxe, aux, &acc);
                else up_ring_tap<8, 1>(Ls, Y, X, dye, j, aux, &acc);
                if (xedge && yedge && j == 0)                       // corner: the two remaining taps of the edge row
                    for (int dx = 0; dx < 3; dx++)
                        if (dx != dxe) up_ring_tap<8, 1>(Ls, Y, X, dye, dx, aux, &acc);
            }
            const float t1 = __shfl_down_sync(0xffffffffu, acc, 1), t2 = __shfl_down_sync(0xffffffffu, acc, 2);
            const float corr = (acc + t1) + t2;
            if (!valid || j != 0) continue;
            const float val = ring[slot] - corr;
            const int idx = Y * Wo + X;
            if (a.ptr_out) a.ptr_out[(size_t)item * Wo * Wo + idx] = val;
            if (amax_better(val, idx, best_v, best_i)) { best_v = val; best_i = idx; }
        }
    }
    TZ_STAMP(11, 0);
    {
        float *sv = aux + 72;
        int *si = reinterpret_cast<int *>(aux + 80);
        amax_warp(best_v, best_i);
        if (lane == 0) { sv[warp] = best_v; si[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < TZ4_NT / 32; k++)
                if (amax_better(sv[k], si[k], best_v, best_i)) { best_v = sv[k]; best_i = si[k]; }
            a.amax_val[(size_t)item * gridDim.x + blockIdx.x] = best_v;
            a.amax_idx[(size_t)item * gridDim.x + blockIdx.x] = best_i;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
    TZ_STAMP(12, 0);
}

// ---------------------------------------------------------------- host side
static long long *g_tz_dbg = nullptr;                    // device buffer of 16 stamps, see ofb_policy_tz_debug
extern "C" int ofb_policy_tz_debug(long long *dev_buf) { g_tz_dbg = dev_buf; return OFB_OK; }

int pol_tz_up4(const ofb_policy *p, const __nv_bfloat16 *in, float *ptr_out, float *amax_val, int *amax_idx, int n_items,
               cudaStream_t st) {
    static thread_local bool configured = false;
    if (!configured) {
        OFB_CUDA_CHECK(cudaFuncSetAttribute(k_tz_up4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Tz4Smem::total));
        configured = true;
    }
    if (n_items == 0) return OFB_OK;
    Tz4Args a = {};
    a.in = in; a.wt = p->w.u4_tz; a.ring_w = p->w.u4_w; a.bias = p->u4_bias;
    a.ptr_out = ptr_out; a.amax_val = amax_val; a.amax_idx = amax_idx;
    a.dbg = g_tz_dbg;
    k_tz_up4<<<dim3(TZ4_STRIPS, n_items), TZ4_NT, Tz4Smem::total, st>>>(a);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

int pol_tz_up4_parts() { return TZ4_STRIPS; }

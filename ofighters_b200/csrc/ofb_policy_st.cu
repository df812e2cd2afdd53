// Sparse trunk12 on tensor cores (sm_100a): conv1 + BN + ReLU + pool -> conv2 + BN + ReLU + pool (agents/qlearnIA_V2.py:129-139),
// 400 x 400 x 2 bits -> 100 x 100 x 8 bf16, evaluated only where it can differ from the empty-arena answer.
//
// The maps are > 96 % zeros, so pool1 (200 x 200 x 8) equals one constant vector bg1 except at the "dirty" pixels whose
// 4 x 4-bit receptive field holds a set bit, and a pool2 cell equals the precomputed empty-arena value of its border class
// unless one of the 4 x 4 pool1 pixels it reads is dirty.  Per arena (default scene: 150 - 1 000 dirty cells of 10 000):
//   1. both bit maps -> shared memory (one bulk async copy); meanwhile every cell of the output gets its empty-arena value;
//   2. bit arithmetic only: M = ship | laser re-aligned to 13 words per row, D1 = dirty pool1 pixels (200 x 200 bits),
//      D2 = dirty cells (100 x 100 bits), row prefix sums of both -> compact lists without atomics, in raster order;
//   3. conv1 for the dirty pool1 pixels only (9-bit stencil LUT exactly like the other engines) -> V1[k];
//   4. dirty cells 128 at a time: ONE M row of a tcgen05 MMA per cell -- K = the cell's 4 x 4 pool1 patch x 8 channels
//      (each patch entry is bg1, a V1 entry found by a popcount prefix, or zero outside the grid = conv2's padding),
//      N = 4 conv2 pixels x 8 channels, B = conv2's weights scattered over the patch (8 K-steps); the draining thread owns
//      a cell: + bias, max over its 4 pixels, ReLU, bf16, one 16-byte store.
// So the only per-pixel CUDA-core work left is the LUT of step 3 (~1 200 pixels per arena) and the 16 patch lookups per
// dirty cell; conv2's 2 304 MACs per dirty cell run on the tensor pipe.  Scenes with more dirty pixels than the lists hold
// (32-ship stress arenas) are processed in bands of cell rows sized from the prefix sums.
// Arithmetic: bf16 weights and activations, fp32 accumulation, like the twin engines (the summation order differs).
#include "ofb_common.cuh"
#include "ofb_policy.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_tc_ptx.cuh"

#ifndef ST_NT
#define ST_NT 384
#endif
#define ST_NG (ST_NT / 128)                 // gather slices: thread = (cell of the tile, every ST_NG-th patch position)
#define ST_CAP1 1280                      // dirty pool1 pixels per band
#define ST_CAP2 2048                      // dirty cells per band
#define ST_RW 13                          // 32-bit words of one 400-bit map row

struct StSmem {
    static constexpr int off_maps = 0;                                 // ship map | laser map (40 000 B); later the MMA's A tile:
    static constexpr int a_bytes = 16 * 128 * 16;                      //   [16 patch positions][128 cells][8 ch bf16]
    static constexpr int off_v1 = 40000;                               // uint4 [CAP1]; before that Q [201][13] words (10 452 B)
    static constexpr int off_l1 = off_v1 + ST_CAP1 * 16;               // u32 [CAP1]: py << 16 | px
    static constexpr int off_l2 = off_l1 + ST_CAP1 * 4;                // u16 [CAP2]: Y * 100 + X
    static constexpr int off_d1 = off_l2 + ST_CAP2 * 2;                // u32 [200][7]
    static constexpr int off_d2 = off_d1 + 200 * 7 * 4;                // u32 [100][4]
    static constexpr int off_p1 = off_d2 + 100 * 4 * 4;                // int [200][7] dirty pool1 pixels before word (py, w), raster order
    static constexpr int off_p2 = off_p1 + 200 * 7 * 4;                // int [100][4] dirty cells before word (Y, w)
    static constexpr int off_rb1 = off_p2 + 100 * 4 * 4;               // int [201] prefix of dirty pool1 pixels per row
    static constexpr int off_rb2 = off_rb1 + 202 * 4;                  // int [101] prefix of dirty cells per row
    static constexpr int off_b = (off_rb2 + 102 * 4 + 15) & ~15;       // conv2 as B operand [8 ks][2 chunks][32 n][8 k] bf16
    static constexpr int b_bytes = 8 * 2 * 32 * 16;                    //   x 3: conv2, conv3, conv4
    static constexpr int off_misc = off_b + 3 * b_bytes;               // c1 bias [8] f32, conv2 bias [8] f32, bg1 (uint4), band ints [8], scan [16],
    static constexpr int off_bar = off_misc + 32 + 32 + 16 + 32 + 64 + 64;   // conv3 / conv4 bias [16] f32; then 2 mbarriers, tmem slot
    static constexpr int bytes = off_bar + 16 + 16;
};
// levels 3 / 4 (fused trunk): D3 [50][2], P3 [50][2], L3 u16 [2500], L4 u16 [625] in the map region behind A tile 0; D4 / P4 [25] in RB1;
// A tile 1 over V1 .. the head of D1 (D2 survives: the restore reads it)
static_assert(StSmem::a_bytes <= 40000 && ST_CAP1 * 16 >= 201 * ST_RW * 4 && 40000 - StSmem::a_bytes >= 400 + 400 + 5000 + 1250 &&
              StSmem::off_v1 + StSmem::a_bytes <= StSmem::off_d2, "k_st_trunk12: aliasing");
static_assert(StSmem::bytes <= 113 * 1024, "k_st_trunk12: two CTAs per SM");

// 32 bits of map row r starting at column 32 c (rows are 400 bits = 12.5 words: odd rows start mid-word)
__device__ __forceinline__ uint32_t st_row_chunk(const uint32_t *__restrict__ m, int r, int c) {
    const int b = r * POL_W + 32 * c, w = b >> 5;
    const uint32_t v = __funnelshift_r(m[w], m[min(w + 1, POL_WORDS - 1)], b & 31);
    return c == ST_RW - 1 ? (v & 0xFFFFu) : v;                        // the 13th chunk holds the row's last 16 columns
}
// bit p of the result = bit 2p of x
__device__ __forceinline__ uint32_t st_even_bits(uint64_t x) {
    x &= 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)x;
}
// Output bit p' (p = 32 w + p') = OR of the input bits 2p-1 .. 2p+2, from the input words 2w-1 .. 2w+2 (wa, wb, wc, wd; words
// outside the row are passed as 0).
__device__ __forceinline__ uint32_t st_window4(uint32_t wa, uint32_t wb, uint32_t wc, uint32_t wd) {
    const uint64_t v = (uint64_t)wb | ((uint64_t)wc << 32);
    const uint64_t prev = wa >> 31, next = wd & 1u;
    const uint64_t g = ((v << 1) | prev) | v;                          // g[b] = in[b - 1] | in[b]
    const uint64_t g64 = (v >> 63) | next;                             // g[64]
    return st_even_bits(g | (g >> 2) | (g64 << 62));                   // bit 2p': g[2p'] | g[2p' + 2]
}

// The whole trunk in this kernel (fz.fuse): pool2 and pool3 live in per-CTA dense images in global memory (L2-resident scratch,
// initialised with the empty-arena values; only the dirty cells are written and put back afterwards), levels 3 and 4 are the
// same cell-patch MMAs on the dirty cells of D3 = window(D2) / D4 = window(D3), and only `flat` (25 x 25 x 8) reaches HBM.
struct StFuse {
    int fuse;
    __nv_bfloat16 *flat;               // [item][POL_FLAT_PITCH]
    uint4 *scratch;                    // [gridDim.x][ST_SCRATCH_CELLS]
    __nv_bfloat16 *tap2, *tap3;        // optional dense copies of pool2 / pool3 (validation taps)
};
__device__ __forceinline__ int st_cls(int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); }

// One level of the fused trunk: the `n` cells of `list` (cell = Y * NDST + X on the NDST x NDST output grid) from the dense
// NSRC x NSRC image `src` (global; written by this CTA before the last barrier -> read at L2), 128 cells per MMA tile.
// The A tile is double-buffered (atile0 / atile1): the gather of tile t + 1 -- cp.async straight from L2 into shared memory,
// every patch position of a thread in flight together -- runs under the MMAs and the drain of tile t.
template <int NSRC, int NDST>
__device__ __forceinline__ void st_gather(const uint4 *src, const uint16_t *list, int tb, int n, uint4 *atile) {
    // item = (patch row, cell, column j of the patch): 4 neighbouring lanes fetch the 64 contiguous bytes of one patch row, so a
    // warp's 32 requests fall into a few 128-byte lines (the L1 -> L2 request rate is what this gather costs)
    const int nb = min(128, n - tb);
#pragma unroll
    for (int k = 0; k < (2048 + ST_NT - 1) / ST_NT; k++) {
        const int idx = threadIdx.x + ST_NT * k, j = idx & 3, c = (idx >> 2) & 127, row = idx >> 9;
        if (idx >= 2048) break;
        if (c < nb) {
            const int cell = list[tb + c], Y = cell / NDST, X = cell - Y * NDST, qy = 2 * Y - 1 + row, qx = 2 * X - 1 + j;
            uint4 *dst = atile + (row * 4 + j) * 128 + c;
            if (qy >= 0 && qy < NSRC && qx >= 0 && qx < NSRC)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src + qy * NSRC + qx) : "memory");
            else
                *dst = make_uint4(0u, 0u, 0u, 0u);                                      // outside the grid: the convolution's zero padding
        }
    }
}
template <int NSRC, int NDST>
__device__ __forceinline__ void st_level(const uint4 *src, const uint16_t *list, int n, uint4 *atile0, uint4 *atile1, uint32_t b16,
                                         const float *bias, uint4 *dst, uint32_t tmem_base, uint64_t *mma_bar, uint32_t &n_mma) {
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t IDESC = instr_desc(32);
    if (n > 0) st_gather<NSRC, NDST>(src, list, 0, n, atile0);
    int buf = 0;
    for (int tb = 0; tb < n; tb += 128, buf ^= 1) {
        const int nb = min(128, n - tb);
        uint4 *atile = buf ? atile1 : atile0;
        asm volatile("cp.async.wait_all;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();                                   // tile tb is in shared memory; the previous tile's drain has left TMEM
        if (warp == 4) {
            tc_fence_after();
            const bool leader = elect_one();
            const uint32_t a16 = smem_u32(atile) >> 4;
#pragma unroll
            for (int ks = 0; ks < 8; ks++) {
                const uint64_t ad = smem_desc(a16 + (uint32_t)(2 * ks * 128), 128, 8);
                const uint64_t bd = smem_desc(b16 + (uint32_t)(ks * 2 * 32), 32, 8);
                if (leader) tc_mma(tmem_base, ad, bd, IDESC, ks ? 1u : 0u);
            }
            if (leader) tc_commit(mma_bar);
            __syncwarp();
        }
        // the other buffer was last read by the MMAs of the tile before this one, whose completion every thread has waited for
        if (tb + 128 < n) st_gather<NSRC, NDST>(src, list, tb + 128, n, buf ? atile0 : atile1);
        mbar_wait(mma_bar, n_mma & 1u);
        n_mma++;
        tc_fence_after();
        if (warp < 4) {
            uint32_t r[32];
            tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
            tc_wait_ld();
            if (tid < nb) {
                float o[8];
#pragma unroll
                for (int co = 0; co < 8; co++)
                    o[co] = fmaxf(fmaxf(__uint_as_float(r[co]), __uint_as_float(r[8 + co])),
                                  fmaxf(__uint_as_float(r[16 + co]), __uint_as_float(r[24 + co]))) + bias[co];
                dst[list[tb + tid]] = pack_relu_bf8(o);
            }
        }
    }
    tc_fence_before();
    __syncthreads();                                       // the level's image is complete; TMEM and both A tiles are free
}

__global__ void __launch_bounds__(ST_NT, 2)
k_st_trunk12(const uint32_t *__restrict__ maps, const PolicyDev w, __nv_bfloat16 *__restrict__ out, const int n_items, long long *stamps,
             const StFuse fz) {
#define ST_STAMP(j) do { if (stamps && blockIdx.x == 0 && tid == 0 && it < 8) stamps[it * 16 + (j)] = clock64(); } while (0)
    extern __shared__ __align__(128) uint8_t st_smem[];
    const uint32_t *bs = reinterpret_cast<const uint32_t *>(st_smem + StSmem::off_maps), *bl = bs + POL_WORDS;
    uint4 *atile = reinterpret_cast<uint4 *>(st_smem + StSmem::off_maps);
    uint4 *v1 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_v1);
    uint32_t *qrow = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_v1);       // Q [201][13], dead before V1 is written
    uint32_t *l1 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_l1);
    int *p1 = reinterpret_cast<int *>(st_smem + StSmem::off_p1), *p2 = reinterpret_cast<int *>(st_smem + StSmem::off_p2);
    uint16_t *l2 = reinterpret_cast<uint16_t *>(st_smem + StSmem::off_l2);
    uint32_t *d1 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d1), *d2 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d2);
    int *rb1 = reinterpret_cast<int *>(st_smem + StSmem::off_rb1), *rb2 = reinterpret_cast<int *>(st_smem + StSmem::off_rb2);
    float *c1b = reinterpret_cast<float *>(st_smem + StSmem::off_misc), *b2 = c1b + 8;
    uint4 *bg1v = reinterpret_cast<uint4 *>(st_smem + StSmem::off_misc + 64);
    int *band = reinterpret_cast<int *>(st_smem + StSmem::off_misc + 80);         // Y1, p_lo, p_hi of the current band
    int *scan = reinterpret_cast<int *>(st_smem + StSmem::off_misc + 112);        // warp totals of the row scan
    uint64_t *mbar = reinterpret_cast<uint64_t *>(st_smem + StSmem::off_bar);      // [0] maps landed, [1] MMAs done
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_bar + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < 8) { c1b[tid] = w.c1_b[tid]; b2[tid] = w.cb[0][tid]; }
    if (tid == 0) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = w.sp_bg1[k];
        *bg1v = pack_bf8(v);
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 8 * 2 * 32; i += ST_NT) {
        reinterpret_cast<uint4 *>(st_smem + StSmem::off_b)[i] = reinterpret_cast<const uint4 *>(w.c2_st)[i];
        if (fz.fuse) {
            reinterpret_cast<uint4 *>(st_smem + StSmem::off_b + StSmem::b_bytes)[i] = reinterpret_cast<const uint4 *>(w.c3_st)[i];
            reinterpret_cast<uint4 *>(st_smem + StSmem::off_b + 2 * StSmem::b_bytes)[i] = reinterpret_cast<const uint4 *>(w.c4_st)[i];
        }
    }
    float *b34 = reinterpret_cast<float *>(st_smem + StSmem::off_misc + 176);      // conv3 bias [8], conv4 bias [8]
    if (tid < 8) { b34[tid] = w.cb[1][tid]; b34[8 + tid] = w.cb[2][tid]; }
    uint4 *scr2 = fz.scratch + (size_t)blockIdx.x * ST_SCRATCH_CELLS, *scr3 = scr2 + 100 * 100;
    const uint4 *bg3 = reinterpret_cast<const uint4 *>(w.sp_bg3), *bg4 = reinterpret_cast<const uint4 *>(w.sp_bg4);
    if (fz.fuse) {                                          // this CTA's pool2 / pool3 images start out as the empty arena's
        const uint4 *bg2i = reinterpret_cast<const uint4 *>(w.sp_bg2);
        for (int i = tid; i < 100 * 100; i += ST_NT) scr2[i] = __ldg(bg2i + st_cls(i / 100, 100) * 3 + st_cls(i % 100, 100));
        for (int i = tid; i < 50 * 50; i += ST_NT) scr3[i] = __ldg(bg3 + st_cls(i / 50, 50) * 3 + st_cls(i % 50, 50));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(32u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the B operand was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint4 *bg2 = reinterpret_cast<const uint4 *>(w.sp_bg2);      // [9 classes] bf16 x 8
    constexpr uint32_t IDESC = instr_desc(32);
    uint32_t n_loads = 0, n_mma = 0;

    int it = -1;
    for (int a = blockIdx.x; a < n_items; a += gridDim.x) {
        it++;
        ST_STAMP(0);
        const uint32_t *src = maps + (size_t)a * 2 * POL_WORDS;
        __nv_bfloat16 *dsta = out + (size_t)a * 100 * 100 * 8;
        // ---- 1. maps -> shared memory; the output's empty-arena values meanwhile
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar[0], 2 * POL_WORDS * 4);
            bulk_g2s(st_smem + StSmem::off_maps, src, 2 * POL_WORDS * 4, &mbar[0]);
        }
        if (fz.fuse) {                                     // only `flat` leaves the kernel: its empty-arena values first
            uint4 *fl = reinterpret_cast<uint4 *>(fz.flat + (size_t)a * POL_FLAT_PITCH);
            for (int i = tid; i < 625; i += ST_NT) fl[i] = __ldg(bg4 + st_cls(i / 25, 25) * 3 + st_cls(i % 25, 25));
        } else if ((tid & 127) < 100) {                    // a thread keeps its column and walks every other row
            const int X = tid & 127, cx = X == 0 ? 0 : (X == 99 ? 2 : 1);
            const uint4 top = __ldg(bg2 + cx), mid = __ldg(bg2 + 3 + cx), bot = __ldg(bg2 + 6 + cx);
            uint4 *col = reinterpret_cast<uint4 *>(dsta) + X;
            for (int Y = tid >> 7; Y < 100; Y += ST_NT / 128) col[Y * 100] = Y == 0 ? top : (Y == 99 ? bot : mid);
        }
        ST_STAMP(1);
        mbar_wait(&mbar[0], n_loads & 1u);
        n_loads++;
        ST_STAMP(2);
        // ---- 2a. Q[j] = M[2j-1] | M[2j] (M = ship | laser; j = 0 .. 200), 13 aligned words per row: the 4 map rows a pool1 row
        //          depends on are Q[py] | Q[py + 1]
        for (int i = tid; i < 201 * ST_RW; i += ST_NT) {
            const int j = i / ST_RW, c = i - j * ST_RW;
            uint32_t v = 0u;
            if (j > 0) v = st_row_chunk(bs, 2 * j - 1, c) | st_row_chunk(bl, 2 * j - 1, c);
            if (j < 200) v |= st_row_chunk(bs, 2 * j, c) | st_row_chunk(bl, 2 * j, c);
            qrow[i] = v;
        }
        __syncthreads();
        ST_STAMP(3);
        // ---- 2b. D1[py] bit px: the 4 x 4 bits (rows 2py-1 .. 2py+2, columns 2px-1 .. 2px+2) hold a set bit
        for (int i = tid; i < 200 * 7; i += ST_NT) {
            const int py = i / 7, wd = i - py * 7, c0 = 2 * wd;
            const uint32_t *q0 = qrow + py * ST_RW, *q1 = q0 + ST_RW;
            const uint32_t wa = c0 > 0 ? (q0[c0 - 1] | q1[c0 - 1]) : 0u, wb = q0[c0] | q1[c0];
            const uint32_t wc = c0 + 1 < ST_RW ? (q0[c0 + 1] | q1[c0 + 1]) : 0u, we = c0 + 2 < ST_RW ? (q0[c0 + 2] | q1[c0 + 2]) : 0u;
            uint32_t dd = st_window4(wa, wb, wc, we);
            if (wd == 6) dd &= 0xFFu;
            d1[i] = dd;
        }
        __syncthreads();
        // ---- 2c. D2[Y] bit X: one of the 4 x 4 pool1 pixels (rows 2Y-1 .. 2Y+2, columns 2X-1 .. 2X+2) is dirty
        for (int i = tid; i < 100 * 4; i += ST_NT) {
            const int Y = i >> 2, wd = i & 3, c0 = 2 * wd;
            uint32_t wa = 0u, wb = 0u, wc = 0u, we = 0u;
            for (int r = max(2 * Y - 1, 0); r <= min(2 * Y + 2, 199); r++) {
                const uint32_t *dr = d1 + r * 7;
                if (c0 > 0) wa |= dr[c0 - 1];
                wb |= dr[c0];
                if (c0 + 1 < 7) wc |= dr[c0 + 1];
                if (c0 + 2 < 7) we |= dr[c0 + 2];
            }
            uint32_t dd = st_window4(wa, wb, wc, we);
            if (wd == 3) dd &= 0xFu;
            d2[i] = dd;
        }
        __syncthreads();
        // ---- 2d. prefix sums in raster order, both levels in one block scan: thread t = pool1 row t (low 16 bits) and cell row t
        //          (high 16 bits; the totals stay below 2^16); then the prefix in front of every word
        {
            int c1 = 0, c2 = 0;
            if (tid < 200)
#pragma unroll
                for (int k = 0; k < 7; k++) c1 += __popc(d1[tid * 7 + k]);
            if (tid < 100)
#pragma unroll
                for (int k = 0; k < 4; k++) c2 += __popc(d2[tid * 4 + k]);
            const int mine = c1 | (c2 << 16);
            int inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            if (lane == 31) scan[warp] = inc;
            __syncthreads();
            int before = 0;
            for (int q = 0; q < warp; q++) before += scan[q];
            const int excl = before + inc - mine;
            if (tid < 200) {
                int run = excl & 0xFFFF;
                rb1[tid] = run;
#pragma unroll
                for (int k = 0; k < 7; k++) { p1[tid * 7 + k] = run; run += __popc(d1[tid * 7 + k]); }
                if (tid == 199) rb1[200] = run;
            }
            if (tid < 100) {
                int run = excl >> 16;
                rb2[tid] = run;
#pragma unroll
                for (int k = 0; k < 4; k++) { p2[tid * 4 + k] = run; run += __popc(d2[tid * 4 + k]); }
                if (tid == 99) rb2[100] = run;
            }
        }
        __syncthreads();
        ST_STAMP(4);
        // ---- 3 / 4. bands of cell rows whose dirty pixels / cells fit the lists (one band for a default arena)
        int Y0 = 0, nband = 0;
        while (Y0 < 100) {
            if (tid == 0) {
                int Y1 = 99;
                const int plo = max(2 * Y0 - 1, 0);
                if (rb1[200] - rb1[plo] > ST_CAP1 || rb2[100] - rb2[Y0] > ST_CAP2) {       // rare: more dirty pixels than the lists hold
                    Y1 = Y0;
                    while (Y1 + 1 < 100 && rb1[min(2 * (Y1 + 1) + 2, 199) + 1] - rb1[plo] <= ST_CAP1 && rb2[Y1 + 2] - rb2[Y0] <= ST_CAP2) Y1++;
                }
                band[0] = Y1; band[1] = plo; band[2] = min(2 * Y1 + 2, 199);
            }
            __syncthreads();
            const int Y1 = band[0], plo = band[1], phi = band[2];
            const int base1 = rb1[plo], n1 = rb1[phi + 1] - base1, base2 = rb2[Y0], n2 = rb2[Y1 + 1] - base2;
            if (n2 > 0) {
                if (nband > 0) {                           // a later band: the A tile has overwritten the maps -- fetch them again
                    __syncthreads();
                    if (tid == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_expect_tx(&mbar[0], 2 * POL_WORDS * 4);
                        bulk_g2s(st_smem + StSmem::off_maps, src, 2 * POL_WORDS * 4, &mbar[0]);
                    }
                    mbar_wait(&mbar[0], n_loads & 1u);
                    n_loads++;
                }
                // the band's dirty cells in raster order: entry index = prefix in front of the word + set bits before it in the word
                for (int i = tid; i < (Y1 - Y0 + 1) * 4; i += ST_NT) {
                    const int Y = Y0 + (i >> 2), wd = i & 3;
                    uint32_t bits = d2[Y * 4 + wd];
                    int k = p2[Y * 4 + wd] - base2;
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        l2[k++] = (uint16_t)(Y * 100 + 32 * wd + b);
                    }
                }
                for (int i = tid; i < (phi - plo + 1) * 7; i += ST_NT) {
                    const int py = plo + i / 7, wd = i % 7;
                    uint32_t bits = d1[py * 7 + wd];
                    int k = p1[py * 7 + wd] - base1;
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        l1[k++] = ((uint32_t)py << 16) | (uint32_t)(32 * wd + b);
                    }
                }
                __syncthreads();
                ST_STAMP(5);
                // ---- 3. conv1 + ReLU + pool for the dirty pool1 pixels (9-bit stencil LUT, like the twin engines), one list
                //         entry per thread and round (V1 is indexed in raster order, like the list)
                for (int e = tid; e < n1; e += ST_NT) {
                    const int py = (int)(l1[e] >> 16), px = (int)(l1[e] & 0xFFFFu);
                    const uint32_t ps = conv1_patch(bs, py, px), pl = conv1_patch(bl, py, px);
                    float v[8];
                    conv1_pool_pixel(ps, pl, w.c1_lut, c1b, v);
                    v1[e] = pack_bf8(v);
                }
                __syncthreads();                           // the maps are dead from here on: the A tile takes their place
                ST_STAMP(6);
                if (stamps && blockIdx.x == 0 && tid == 0 && it < 8) { stamps[it * 16 + 8] = n1; stamps[it * 16 + 9] = n2; }
                // ---- 4. conv2 + pool on the tensor pipe, 128 dirty cells per MMA tile
                for (int tb = 0; tb < n2; tb += 128) {
                    const int nb = min(128, n2 - tb), c = tid & 127;
                    if (c < nb) {
                        const int cell = l2[tb + c], Y = cell / 100, X = cell - Y * 100;
#pragma unroll
                        for (int k = 0; k < (16 + ST_NG - 1) / ST_NG; k++) {
                            const int pos = (tid >> 7) + ST_NG * k, qy = 2 * Y - 1 + (pos >> 2), qx = 2 * X - 1 + (pos & 3);
                            if (pos >= 16) break;
                            uint4 val = make_uint4(0u, 0u, 0u, 0u);                     // outside the grid: conv2's zero padding
                            if (qy >= 0 && qy < 200 && qx >= 0 && qx < 200) {
                                const uint32_t *dr = d1 + qy * 7;
                                const int wd = qx >> 5, bit = qx & 31;
                                if ((dr[wd] >> bit) & 1u) val = v1[p1[qy * 7 + wd] - base1 + __popc(dr[wd] & ((1u << bit) - 1u))];
                                else val = *bg1v;
                            }
                            atile[pos * 128 + c] = val;
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMAs
                    tc_fence_before();
                    __syncthreads();
                    if (warp == 4) {
                        tc_fence_after();
                        const bool leader = elect_one();
                        const uint32_t a16 = smem_u32(atile) >> 4, b16 = smem_u32(st_smem + StSmem::off_b) >> 4;
#pragma unroll
                        for (int ks = 0; ks < 8; ks++) {
                            const uint64_t ad = smem_desc(a16 + (uint32_t)(2 * ks * 128), 128, 8);
                            const uint64_t bd = smem_desc(b16 + (uint32_t)(ks * 2 * 32), 32, 8);
                            if (leader) tc_mma(tmem_base, ad, bd, IDESC, ks ? 1u : 0u);
                        }
                        if (leader) tc_commit(&mbar[1]);
                        __syncwarp();
                    }
                    mbar_wait(&mbar[1], n_mma & 1u);
                    n_mma++;
                    tc_fence_after();
                    if (warp < 4) {                        // thread = cell: + bias, max over its 4 conv2 pixels, ReLU, bf16
                        uint32_t r[32];
                        tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
                        tc_wait_ld();
                        if (tid < nb) {
                            float o[8];
#pragma unroll
                            for (int co = 0; co < 8; co++)
                                o[co] = fmaxf(fmaxf(__uint_as_float(r[co]), __uint_as_float(r[8 + co])),
                                              fmaxf(__uint_as_float(r[16 + co]), __uint_as_float(r[24 + co]))) + b2[co];
                            if (fz.fuse) scr2[l2[tb + tid]] = pack_relu_bf8(o);
                            else *reinterpret_cast<uint4 *>(dsta + (size_t)l2[tb + tid] * 8) = pack_relu_bf8(o);
                        }
                    }
                    tc_fence_before();
                    __syncthreads();                       // TMEM and the A tile are free again
                }
            }
            Y0 = Y1 + 1;
            nband++;
            __syncthreads();                               // everyone has read the band's bounds before thread 0 writes the next ones
        }
        __syncthreads();                                   // nobody still reads this arena's lists / bitmaps
        if (fz.fuse) {
            // ---- 5. level 3: D3 = dirty pool3 cells (50 x 50), from D2; lists by popcount prefix (one warp); conv3 on the dirty cells
            // (shared memory in levels 3 / 4: A tile 0 = the first 32 KB of the map region, bitmaps and lists in the 7 232 B after
            //  it, A tile 1 over V1 / L1 / L2 / the head of D1 -- all dead by now; D2 stays for the restore)
            uint32_t *d3 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_maps + StSmem::a_bytes);
            int *p3 = reinterpret_cast<int *>(d3 + 100);
            uint16_t *l3 = reinterpret_cast<uint16_t *>(p3 + 100), *l4 = l3 + 2500;
            uint32_t *d4 = reinterpret_cast<uint32_t *>(rb1);
            int *p4 = rb1 + 32;
            uint4 *atile1 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_v1);
            int *cnt = band;                               // n3, n4
            if (tid < 100) {
                const int Y = tid >> 1, wd = tid & 1, c0 = 2 * wd;
                uint32_t wa = 0u, wb = 0u, wc = 0u, we = 0u;
                for (int r = max(2 * Y - 1, 0); r <= min(2 * Y + 2, 99); r++) {
                    const uint32_t *dr = d2 + r * 4;
                    if (c0 > 0) wa |= dr[c0 - 1];
                    wb |= dr[c0];
                    wc |= dr[c0 + 1];
                    if (c0 + 2 < 4) we |= dr[c0 + 2];
                }
                uint32_t dd = st_window4(wa, wb, wc, we);
                if (wd == 1) dd &= 0x3FFFFu;               // 50 cells per row
                d3[tid] = dd;
            }
            __syncthreads();
            if (tid < 25) {                                // D4 = dirty pool4 cells (25 x 25), from D3
                uint32_t wb = 0u, wc = 0u;
                for (int r = max(2 * tid - 1, 0); r <= min(2 * tid + 2, 49); r++) { wb |= d3[r * 2]; wc |= d3[r * 2 + 1]; }
                d4[tid] = st_window4(0u, wb, wc, 0u) & 0x1FFFFFFu;
            }
            __syncthreads();
            if (warp == 0) {                               // row-major prefixes of D3 (100 words) and D4 (25 words)
                int run = 0;
                for (int base = 0; base < 100; base += 32) {
                    const int i = base + lane, c = i < 100 ? __popc(d3[i]) : 0;
                    int inc = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                    if (i < 100) p3[i] = run + inc - c;
                    run += __shfl_sync(0xffffffffu, inc, 31);
                }
                const int c4 = lane < 25 ? __popc(d4[lane]) : 0;
                int inc = c4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                if (lane < 25) p4[lane] = inc - c4;
                if (lane == 31) { cnt[0] = run; cnt[1] = inc; }
            }
            __syncthreads();
            if (tid < 100) {
                uint32_t bits = d3[tid];
                int k = p3[tid];
                while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; l3[k++] = (uint16_t)((tid >> 1) * 50 + 32 * (tid & 1) + b); }
            } else if (tid >= 128 && tid < 153) {
                const int Y = tid - 128;
                uint32_t bits = d4[Y];
                int k = p4[Y];
                while (bits) { const int b = __ffs(bits) - 1; bits &= bits - 1; l4[k++] = (uint16_t)(Y * 25 + b); }
            }
            __syncthreads();
            const int n3 = cnt[0], n4 = cnt[1];
            const uint32_t b16 = smem_u32(st_smem + StSmem::off_b) >> 4;
            st_level<100, 50>(scr2, l3, n3, atile, atile1, b16 + StSmem::b_bytes / 16, b34, scr3, tmem_base, &mbar[1], n_mma);
            // ---- 6. level 4: conv4 on the dirty pool4 cells, straight into `flat` (NHWC flatten = 16 bytes per cell, raster order)
            st_level<50, 25>(scr3, l4, n4, atile, atile1, b16 + 2 * StSmem::b_bytes / 16, b34 + 8,
                             reinterpret_cast<uint4 *>(fz.flat + (size_t)a * POL_FLAT_PITCH), tmem_base, &mbar[1], n_mma);
            // ---- 7. validation taps, then the dirty cells of both images go back to their empty-arena values
            if (fz.tap2) {
                uint4 *t2 = reinterpret_cast<uint4 *>(fz.tap2) + (size_t)a * 10000, *t3 = reinterpret_cast<uint4 *>(fz.tap3) + (size_t)a * 2500;
                for (int i = tid; i < 10000; i += ST_NT) t2[i] = __ldcg(scr2 + i);
                for (int i = tid; i < 2500; i += ST_NT) t3[i] = __ldcg(scr3 + i);
                __syncthreads();
            }
            const uint4 *bg2i = reinterpret_cast<const uint4 *>(w.sp_bg2);
            for (int i = tid; i < 100 * 4; i += ST_NT) {
                uint32_t bits = d2[i];
                const int Y = i >> 2;
                while (bits) {
                    const int X = 32 * (i & 3) + __ffs(bits) - 1;
                    bits &= bits - 1;
                    scr2[Y * 100 + X] = __ldg(bg2i + st_cls(Y, 100) * 3 + st_cls(X, 100));
                }
            }
            for (int e = tid; e < n3; e += ST_NT) {
                const int cell = l3[e];
                scr3[cell] = __ldg(bg3 + st_cls(cell / 50, 50) * 3 + st_cls(cell % 50, 50));
            }
            __syncthreads();
        }
        ST_STAMP(7);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(32u));
}

static long long *g_st_stamps = nullptr;
extern "C" int ofb_policy_st_stamps(long long *dev_buf) { g_st_stamps = dev_buf; return OFB_OK; }

static int st_launch(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, const StFuse &fz, cudaStream_t st) {
    if (n_items <= 0) return OFB_OK;
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_st_trunk12, (int)StSmem::bytes));
    int n_sm = 148;
    OFB_CUDA_CHECK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, p->device));
    int grid = n_items < 2 * n_sm ? n_items : 2 * n_sm;
    if (grid > ST_MAX_CTAS) grid = ST_MAX_CTAS;
    k_st_trunk12<<<grid, ST_NT, StSmem::bytes, st>>>(maps, p->w, out, n_items, g_st_stamps, fz);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

int pol_st_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st) {
    StFuse fz = {};
    return st_launch(p, maps, out, n_items, fz, st);
}

int pol_st_trunk(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *flat, __nv_bfloat16 *tap2, __nv_bfloat16 *tap3, int n_items,
                 cudaStream_t st) {
    StFuse fz = {};
    fz.fuse = 1; fz.flat = flat; fz.scratch = p->ws.st_scratch; fz.tap2 = tap2; fz.tap3 = tap2 ? tap3 : nullptr;
    return st_launch(p, maps, nullptr, n_items, fz, st);
}

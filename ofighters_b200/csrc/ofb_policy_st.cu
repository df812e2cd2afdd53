// Sparse trunk on tensor cores (sm_100a): conv1 + BN + ReLU + pool -> ... -> conv4 + BN + ReLU + pool (agents/qlearnIA_V2.py:129-147),
// 400 x 400 x 2 bits -> 25 x 25 x 8 bf16 (`flat`), evaluated only where it can differ from the empty-arena answer.
//
// The maps are > 96 % zeros, so pool1 (200 x 200 x 8) equals one constant vector bg1 except at the "dirty" pixels whose
// 4 x 4-bit receptive field holds a set bit, and a pooled cell of any deeper level equals the precomputed empty-arena value of its
// border class unless one of the 4 x 4 cells it reads is dirty.  Per arena (default scene: 150 - 1 000 dirty pool2 cells of 10 000):
//   1. both bit maps -> shared memory (one bulk async copy); meanwhile `flat` gets its empty-arena values;
//   2. bit arithmetic only: Q = pairwise OR of map rows re-aligned to 13 words per row, D1 = dirty pool1 pixels (200 x 200 bits),
//      D2 / D3 / D4 = dirty cells of the levels above (a 4-wide window OR + even-bit compress of the level below), row-major
//      popcount prefixes of every level -> compact lists without atomics, in raster order;
//   3. conv1 for the dirty pool1 pixels only (9-bit stencil LUT exactly like the other engines) -> V1[k];
//   4. per level, dirty cells 128 at a time: ONE M row of a tcgen05 MMA per cell -- K = the cell's 4 x 4 input patch x 8 channels,
//      N = 4 conv pixels x 8 channels, B = the conv's weights scattered over the patch (8 K-steps); the draining thread owns a cell:
//      + bias, max over its 4 pixels, ReLU, bf16.  The computed cells of a level stay in shared memory as a compact list in raster
//      order (V1 -> V2 -> V3); a patch row of the next level is 4 adjacent entries, found with ONE bitmap word pair and ONE prefix
//      (entry = the empty-arena value of its border class, a V entry, or zero outside the grid = the convolution's padding).
// Only `flat` reaches HBM.  Arenas whose dirty cells exceed the lists (32-ship stress scenes, all-ones maps) take the fallback:
// level 2 in bands of cell rows sized from the prefix sums, pool2 / pool3 in per-CTA dense images in global memory (L2-resident
// scratch, initialised on first use with the empty arena's values; only the dirty cells are written and put back afterwards).
// Arithmetic: bf16 weights and activations, fp32 accumulation, like the twin engines (the summation order differs).
#include "ofb_common.cuh"
#include "ofb_policy.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_tc_ptx.cuh"

#ifndef ST_NT
#define ST_NT 384
#endif
#define ST_CAP1 1280                      // dirty pool1 pixels per band
#define ST_CAP2 2048                      // dirty pool2 cells per band
#define ST_VCAP 1024                      // compact path: dirty pool2 / pool3 cells of an arena kept in shared memory
#define ST_RW 13                          // 32-bit words of one 400-bit map row

struct StSmem {
    static constexpr int off_maps = 0;                                 // ship map | laser map (40 000 B); later the MMA's A tile:
    static constexpr int a_bytes = 16 * 128 * 16;                      //   [16 patch positions][128 cells][8 ch bf16]
    // V1 .. P1 are contiguous: all four are dead after level 2, when the region holds conv3's and conv4's B operands and V3
    static constexpr int off_v1 = 40000;                               // uint4 [CAP1]; before that Q [201][13] words (10 452 B)
    static constexpr int off_l1 = off_v1 + ST_CAP1 * 16;               // u32 [CAP1]: py << 16 | px
    static constexpr int off_d1 = off_l1 + ST_CAP1 * 4;                // u32 [200][7]
    static constexpr int off_p1 = off_d1 + 200 * 7 * 4;                // int [200][7] dirty pool1 pixels before word (py, w), raster order
    static constexpr int off_l2 = off_p1 + 200 * 7 * 4;                // u16 [CAP2]: Y * 100 + X; compact path: L3 = the second half
    static constexpr int off_d2 = off_l2 + ST_CAP2 * 2;                // u32 [100][4]
    static constexpr int off_p2 = off_d2 + 100 * 4 * 4;                // int [100][4] dirty cells before word (Y, w)
    static constexpr int off_d3 = off_p2 + 100 * 4 * 4;                // u32 [50][2]
    static constexpr int off_p3 = off_d3 + 100 * 4;                    // int [50][2]
    static constexpr int off_d4 = off_p3 + 100 * 4;                    // u32 [25] (+ 7)
    static constexpr int off_p4 = off_d4 + 32 * 4;                     // int [25] (+ 7)
    static constexpr int off_l4 = off_p4 + 32 * 4;                     // u16 [625] (+ 15): Y * 25 + X
    static constexpr int off_rb1 = off_l4 + 640 * 2;                   // int [201] prefix of dirty pool1 pixels per row
    static constexpr int off_rb2 = off_rb1 + 202 * 4;                  // int [101] prefix of dirty cells per row
    static constexpr int off_b = (off_rb2 + 102 * 4 + 15) & ~15;       // conv2 as B operand [8 ks][2 chunks][32 n][8 k] bf16
    static constexpr int b_bytes = 8 * 2 * 32 * 16;
    static constexpr int off_v2 = off_b + b_bytes;                     // uint4 [VCAP]: pool2's dirty cells, raster order
    static constexpr int off_misc = off_v2 + ST_VCAP * 16;             // c1 bias [8], conv2 bias [8], conv3 / conv4 bias [16] f32, bg1 (uint4),
                                                                       // band ints [8], scan [16], bg1 / bg2 / bg3 classes: 3 x uint4 [9 + a zero entry]
    static constexpr int off_bar = off_misc + 32 + 32 + 64 + 16 + 32 + 64 + 160 + 160 + 160;    // 6 mbarriers, tmem slot
    static constexpr int bytes = off_bar + 48 + 8;
    // levels 3 / 4: conv3's and conv4's B operands and V3 over the dead V1 .. P1
    static constexpr int off_b34 = off_v1;
    static constexpr int off_v3 = off_v1 + 2 * b_bytes;
};
// fallback path, levels 3 / 4: L3 u16 [2500], L4 u16 [625] in the map region behind the A tile
static_assert(StSmem::a_bytes <= 40000 && ST_CAP1 * 16 >= 201 * ST_RW * 4 && 40000 - StSmem::a_bytes >= 5000 + 1250 &&
              StSmem::off_v3 + ST_VCAP * 16 <= StSmem::off_l2 && 2 * ST_VCAP <= ST_CAP2, "k_st_trunk12: aliasing");
static_assert(StSmem::bytes <= 113 * 1024, "k_st_trunk12: two CTAs per SM");

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

struct StFuse {
    int fuse;                          // 1 = the whole trunk (levels 2 .. 4 -> flat); 0 = conv1 + conv2 only (dense pool2 -> `out`)
    __nv_bfloat16 *flat;               // [item][POL_FLAT_PITCH]
    uint4 *scratch;                    // fallback path: [gridDim.x][ST_SCRATCH_CELLS]
    __nv_bfloat16 *tap2, *tap3;        // optional dense copies of pool2 / pool3 (validation taps)
};
__device__ __forceinline__ int st_cls(int i, int n) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); }

// The A tile of 128 cells of `list` (cell = Y * NDST + X on the NDST x NDST output grid; input grid NSRC = 2 NDST): item = (cell,
// patch row) = the 4 adjacent input entries (2Y - 1 + r, 2X - 1 .. 2X + 2).  Compact source: bitmap D [NSRC][RW words] of the
// computed ("dirty") entries, P = entries before each word in raster order, V = their values in raster order, every other entry
// = the empty-arena value of its border class (bg[9]; one value when CONST_BG), zero outside the grid.
template <int NSRC, int NDST, int RW, bool CONST_BG>
struct StGatherCompact {
    const uint32_t *D;
    const int *P;
    int base;
    const uint4 *V, *bg;                                   // bg[9] classes (bg[0] only when CONST_BG), bg[9] = zeros
    const uint16_t *list;
    uint4 *atile;
    // A thread of warps 4 .. (the draining warps 0 - 3 sit the gather out) takes TWO patch rows of one cell: the cell is decoded
    // once, and the code is straight-line (no data-dependent branches, every load's address is valid) so that the two rows overlap
    // in the pipeline.  An entry is fetched through a shared-memory ADDRESS chosen among {V entry, class value, zeros}.
    __device__ __forceinline__ void operator()(int tb, int nb) const {
        static_assert(ST_NT >= 384, "two rows per thread need 256 gathering threads");
        const int t = (int)threadIdx.x - 128, c = t & 127, r0 = (t >> 7) * 2;
        if (t < 0 || t >= 256 || (c & ~31) >= nb) return;                // (warp-uniform)
        const uint32_t v16 = smem_u32(V), bg16 = smem_u32(bg), zero16 = bg16 + 9 * 16;
        const bool on = c < nb;
        const int cell = list[tb + (on ? c : 0)], Y = cell / NDST, X = cell - Y * NDST, x0 = 2 * X - 1;
        const int start = max(x0, 0), w = start >> 5, sh = start & 31;
        const uint32_t below = (1u << sh) - 1u;
        uint32_t addr[2][4];
        uint4 val[2][4];
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const int qy = 2 * Y - 1 + r0 + k;
            const bool row_in = qy >= 0 && qy < NSRC;
            const int qc = min(max(qy, 0), NSRC - 1);
            const uint32_t lo = D[qc * RW + w], hi = D[qc * RW + w + 1];        // (hi matters only inside the row: see the masks of D)
            const int pw = P[qc * RW + w];
            uint32_t m = __funnelshift_r(lo, hi, sh) & 0xFu;
            if (x0 < 0) m = (m << 1) & 0xFu;
            if (!row_in) m = 0u;
            const uint32_t vi = v16 + (uint32_t)(pw - base + __popc(lo & below)) * 16u;
            const uint32_t cls16 = CONST_BG ? bg16 : bg16 + (uint32_t)(st_cls(qc, NSRC) * 3) * 16u;
            // interior columns: a dirty entry = the next V entry of the row, anything else = the row's middle class (or zeros)
            const uint32_t plain = !row_in ? zero16 : (CONST_BG ? cls16 : cls16 + 16u);
            uint32_t nextv = vi;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const bool dirty = (m >> j) & 1u;
                addr[k][j] = dirty ? nextv : plain;
                nextv += dirty ? 16u : 0u;
            }
            if (X == 0 || X == NDST - 1) {                 // first / last cell of a row: the grid's border column and the padding
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int qx = x0 + j;
                    if ((m >> j) & 1u) continue;
                    if (!row_in || qx < 0 || qx >= NSRC) addr[k][j] = zero16;
                    else if (!CONST_BG && (qx == 0 || qx == NSRC - 1)) addr[k][j] = cls16 + (qx == 0 ? 0u : 32u);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 2; k++)
#pragma unroll
            for (int j = 0; j < 4; j++)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(val[k][j].x), "=r"(val[k][j].y), "=r"(val[k][j].z), "=r"(val[k][j].w) : "r"(addr[k][j]) : "memory");
        if (on) {
            uint4 *dst = atile + (r0 * 4) * 128 + c;
#pragma unroll
            for (int k = 0; k < 2; k++)
#pragma unroll
                for (int j = 0; j < 4; j++) dst[(k * 4 + j) * 128] = val[k][j];
        }
    }
};
// Fallback source: the dense NSRC x NSRC image `src` in global memory (written by this CTA before the last barrier -> read at L2)
template <int NSRC, int NDST>
struct StGatherDense {
    const uint4 *src;
    const uint16_t *list;
    uint4 *atile;
    __device__ __forceinline__ void operator()(int tb, int nb) const {
        for (int idx = (int)threadIdx.x - 128; idx < 2048; idx += ST_NT - 128) {                 // (the gathering warps 4 ..)
            const int j = idx & 3, c = (idx >> 2) & 127, row = idx >> 9;
            if (c >= nb) continue;
            const int cell = list[tb + c], Y = cell / NDST, X = cell - Y * NDST, qy = 2 * Y - 1 + row, qx = 2 * X - 1 + j;
            uint4 val = make_uint4(0u, 0u, 0u, 0u);                                      // outside the grid: the convolution's zero padding
            if (qy >= 0 && qy < NSRC && qx >= 0 && qx < NSRC) val = __ldcg(src + qy * NSRC + qx);
            atile[(row * 4 + j) * 128 + c] = val;
        }
    }
};

// One level: the `n` cells of `list`, 128 per MMA tile, as a two-role pipeline.  Warps 4 .. gather tile t into the A tile (after the
// MMAs of tile t - 1 have finished reading it), meet at a named barrier, and warp 4 issues the tile's 8 MMAs (N = 32) into TMEM
// stage t & 1; warps 0 - 3 drain tile t - 1 meanwhile: + bias, max over the cell's 4 conv pixels, ReLU, bf16 into dst[list[.]]
// (by_list) or dst[position in the list].  `gt` counts the tiles of the whole kernel (stage and barrier parities).  The level ends
// with a block barrier: every drained cell is visible, TMEM and the A tile are free.  `pf_*`: after the gather of the level's LAST
// tile one thread may start a bulk copy (the next levels' B operands into memory this level's gathers were the last readers of).
template <class G>
__device__ __forceinline__ void st_tiles(const G &gather, const uint16_t *list, int n, uint4 *atile, uint32_t b16, const float *bias,
                                         uint4 *dst, bool by_list, uint32_t tmem_base, uint64_t *mma_done, uint64_t *tfree, uint32_t &gt,
                                         uint64_t *wait_bar, uint32_t wait_parity, uint64_t *pf_bar, void *pf_dst, const void *pf_src0,
                                         const void *pf_src1, long long *stamp) {
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t IDESC = instr_desc(32);
    if (n <= 0) return;
    for (int tb = 0; tb < n; tb += 128, gt++) {
        const int nb = min(128, n - tb);
        const uint32_t stage = gt & 1u, use = gt >> 1;
        if (warp >= 4) {
            if (tb > 0) mbar_wait(&mma_done[stage ^ 1u], ((gt - 1u) >> 1) & 1u);       // the previous tile's MMAs have read the A tile
            gather(tb, nb);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the MMAs
            asm volatile("bar.sync 1, %0;" ::"n"(ST_NT - 128) : "memory");
            if (warp == 4) {
                if (wait_bar && tb == 0) mbar_wait(wait_bar, wait_parity);     // this level's B operand has landed
                if (use >= 1u) { mbar_wait(&tfree[stage], (use - 1u) & 1u); tc_fence_after(); }   // the stage's last tile has been drained
                const bool leader = elect_one();
                const uint32_t a16 = smem_u32(atile) >> 4, d = tmem_base + stage * 32u;
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {
                    const uint64_t ad = smem_desc(a16 + (uint32_t)(2 * ks * 128), 128, 8);
                    const uint64_t bd = smem_desc(b16 + (uint32_t)(ks * 2 * 32), 32, 8);
                    if (leader) tc_mma(d, ad, bd, IDESC, ks ? 1u : 0u);
                }
                if (leader) tc_commit(&mma_done[stage]);
                __syncwarp();
                if (pf_bar && tb + 128 >= n && elect_one()) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_expect_tx(pf_bar, 2 * StSmem::b_bytes);
                    bulk_g2s(pf_dst, pf_src0, StSmem::b_bytes, pf_bar);
                    bulk_g2s(reinterpret_cast<uint8_t *>(pf_dst) + StSmem::b_bytes, pf_src1, StSmem::b_bytes, pf_bar);
                }
                __syncwarp();
            }
        } else {
            mbar_wait(&mma_done[stage], use & 1u);
            tc_fence_after();
            if (stamp && tb == 0) stamp[2] = clock64();
            uint32_t r[32];
            tc_ld32(tmem_base + stage * 32u + ((uint32_t)(warp * 32) << 16), r);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&tfree[stage]);
            if (tid < nb) {
                float o[8];
#pragma unroll
                for (int co = 0; co < 8; co++)
                    o[co] = fmaxf(fmaxf(__uint_as_float(r[co]), __uint_as_float(r[8 + co])),
                                  fmaxf(__uint_as_float(r[16 + co]), __uint_as_float(r[24 + co]))) + bias[co];
                dst[by_list ? (int)list[tb + tid] : tb + tid] = pack_relu_bf8(o);
            }
            if (stamp && tb == 0) stamp[3] = clock64();
        }
    }
    tc_fence_before();
    __syncthreads();
}

// raster-order list of the set bits of bitmap word `wd` (NW words per row, NCOL cells per row), starting at entry k
template <int NW, int NCOL>
__device__ __forceinline__ void st_list_word(uint32_t bits, int wd, int k, uint16_t *list) {
    const int Y = wd / NW, x0 = 32 * (wd - Y * NW);
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        list[k++] = (uint16_t)(Y * NCOL + x0 + b);
    }
}

__global__ void __launch_bounds__(ST_NT, 2)
k_st_trunk12(const uint32_t *__restrict__ maps, const PolicyDev w, __nv_bfloat16 *__restrict__ out, const int n_items, long long *stamps,
             const StFuse fz) {
#define ST_STAMP(j) do { if (stamps && blockIdx.x == 0 && tid == 0 && it < 8) stamps[it * 16 + (j)] = clock64(); } while (0)
    extern __shared__ __align__(128) uint8_t st_smem[];
    const uint32_t *bs = reinterpret_cast<const uint32_t *>(st_smem + StSmem::off_maps), *bl = bs + POL_WORDS;
    uint4 *atile = reinterpret_cast<uint4 *>(st_smem + StSmem::off_maps);
    uint4 *v1 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_v1);
    uint4 *v2 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_v2), *v3 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_v3);
    uint32_t *qrow = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_v1);       // Q [201][13], dead before V1 is written
    uint32_t *l1 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_l1);
    int *p1 = reinterpret_cast<int *>(st_smem + StSmem::off_p1), *p2 = reinterpret_cast<int *>(st_smem + StSmem::off_p2);
    int *p3 = reinterpret_cast<int *>(st_smem + StSmem::off_p3), *p4 = reinterpret_cast<int *>(st_smem + StSmem::off_p4);
    uint16_t *l2 = reinterpret_cast<uint16_t *>(st_smem + StSmem::off_l2);
    uint32_t *d1 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d1), *d2 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d2);
    uint32_t *d3 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d3), *d4 = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_d4);
    int *rb1 = reinterpret_cast<int *>(st_smem + StSmem::off_rb1), *rb2 = reinterpret_cast<int *>(st_smem + StSmem::off_rb2);
    float *c1b = reinterpret_cast<float *>(st_smem + StSmem::off_misc), *b2 = c1b + 8, *b34 = c1b + 16;   // conv3 bias [8], conv4 bias [8]
    uint4 *bg1v = reinterpret_cast<uint4 *>(st_smem + StSmem::off_misc + 128);
    int *band = reinterpret_cast<int *>(st_smem + StSmem::off_misc + 144);        // Y1, p_lo, p_hi of the current band; [4] n3, [5] n4
    int *scan = reinterpret_cast<int *>(st_smem + StSmem::off_misc + 176);        // warp totals of the row scan
    uint4 *bgc1 = reinterpret_cast<uint4 *>(st_smem + StSmem::off_misc + 240), *bgc2 = bgc1 + 10, *bgc3 = bgc2 + 10;   // [9] = zeros
    // [0] maps landed, [1] [2] MMAs of TMEM stage 0 / 1 done, [3] [4] stage 0 / 1 drained, [5] conv3 / conv4 operands landed
    uint64_t *mbar = reinterpret_cast<uint64_t *>(st_smem + StSmem::off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(st_smem + StSmem::off_bar + 48);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid < 8) { c1b[tid] = w.c1_b[tid]; b2[tid] = w.cb[0][tid]; b34[tid] = w.cb[1][tid]; b34[8 + tid] = w.cb[2][tid]; }
    if (tid >= 32 && tid < 41) bgc2[tid - 32] = reinterpret_cast<const uint4 *>(w.sp_bg2)[tid - 32];
    if (tid >= 64 && tid < 73 && fz.fuse) bgc3[tid - 64] = reinterpret_cast<const uint4 *>(w.sp_bg3)[tid - 64];
    if (tid >= 96 && tid < 99) bgc1[(tid - 96) * 10 + 9] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; k++) v[k] = w.sp_bg1[k];
        *bg1v = pack_bf8(v);
        bgc1[0] = *bg1v;
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_init(&mbar[2], 1);
        mbar_init(&mbar[3], 4);
        mbar_init(&mbar[4], 4);
        mbar_init(&mbar[5], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 8 * 2 * 32; i += ST_NT)
        reinterpret_cast<uint4 *>(st_smem + StSmem::off_b)[i] = reinterpret_cast<const uint4 *>(w.c2_st)[i];
    uint4 *scr2 = fz.scratch + (size_t)blockIdx.x * ST_SCRATCH_CELLS, *scr3 = scr2 + 100 * 100;
    const uint4 *bg2g = reinterpret_cast<const uint4 *>(w.sp_bg2), *bg3g = reinterpret_cast<const uint4 *>(w.sp_bg3);
    const uint4 *bg4g = reinterpret_cast<const uint4 *>(w.sp_bg4);
    bool scratch_ready = false;                             // the fallback's dense images are initialised on first use
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the B operand was written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t b16 = smem_u32(st_smem + StSmem::off_b) >> 4, b34_16 = smem_u32(st_smem + StSmem::off_b34) >> 4;
    uint32_t n_loads = 0, gt = 0, n_b34 = 0;

    int it = -1;
    for (int a = blockIdx.x; a < n_items; a += gridDim.x) {
        it++;
        ST_STAMP(0);
        long long *tstamp = (stamps && blockIdx.x == 0 && tid == 0 && it < 8) ? stamps + 128 + it * 16 : nullptr;
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
            for (int i = tid; i < 625; i += ST_NT) fl[i] = __ldg(bg4g + st_cls(i / 25, 25) * 3 + st_cls(i % 25, 25));
        } else if ((tid & 127) < 100) {                    // a thread keeps its column and walks every other row
            const int X = tid & 127, cx = X == 0 ? 0 : (X == 99 ? 2 : 1);
            const uint4 top = __ldg(bg2g + cx), mid = __ldg(bg2g + 3 + cx), bot = __ldg(bg2g + 6 + cx);
            uint4 *col = reinterpret_cast<uint4 *>(dsta) + X;
            for (int Y = tid >> 7; Y < 100; Y += ST_NT / 128) col[Y * 100] = Y == 0 ? top : (Y == 99 ? bot : mid);
        }
        ST_STAMP(1);
        mbar_wait(&mbar[0], n_loads & 1u);
        n_loads++;
        ST_STAMP(2);
        // ---- 2a. Q[j] = M[2j-1] | M[2j] (M = ship | laser; j = 0 .. 200), 13 aligned words per row: the 4 map rows a pool1 row
        //          depends on are Q[py] | Q[py + 1]
        // (a map row is 12.5 words: row 2j starts at word 25 j, row 2j - 1 sixteen bits into word 25 j - 13)
        for (int i = tid; i < 201 * ST_RW; i += ST_NT) {
            const int j = i / ST_RW, c = i - j * ST_RW, we = 25 * j + c, wo = we - 13;
            uint32_t v = 0u;
            if (j < 200) {
                v = bs[we] | bl[we];
                if (c == ST_RW - 1) v &= 0xFFFFu;              // the 13th chunk holds the row's last 16 columns
            }
            if (j > 0) {
                const uint32_t lo = bs[wo] | bl[wo], hi = c == ST_RW - 1 ? 0u : (bs[wo + 1] | bl[wo + 1]);
                v |= __funnelshift_r(lo, hi, 16);
            }
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
        // ---- 2d. prefix sums in raster order, levels 1 and 2 in one block scan: thread t = pool1 row t (low 16 bits) and cell row t
        //          (high 16 bits; the totals stay below 2^16); then the prefix in front of every word.  Beside it (threads that
        //          have no row): D3 = dirty pool3 cells (50 x 50) from D2, then D4 (25 x 25) from D3 and both their prefixes.
        {
            int c1 = 0, c2 = 0;
            if (tid < 200)
#pragma unroll
                for (int k = 0; k < 7; k++) c1 += __popc(d1[tid * 7 + k]);
            if (tid < 100)
#pragma unroll
                for (int k = 0; k < 4; k++) c2 += __popc(d2[tid * 4 + k]);
            if (fz.fuse && tid >= 256 && tid < 356) {
                const int t = tid - 256, Y = t >> 1, wd = t & 1, c0 = 2 * wd;
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
                d3[t] = dd;
            }
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
            if (fz.fuse && warp == 8) {                    // D4, its prefix
                uint32_t dd = 0u;
                if (lane < 25) {
                    uint32_t wb = 0u, wc = 0u;
                    for (int r = max(2 * lane - 1, 0); r <= min(2 * lane + 2, 49); r++) { wb |= d3[r * 2]; wc |= d3[r * 2 + 1]; }
                    dd = st_window4(0u, wb, wc, 0u) & 0x1FFFFFFu;
                    d4[lane] = dd;
                }
                const int c4 = __popc(dd);
                int i4 = c4;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, i4, o); if (lane >= o) i4 += t; }
                if (lane < 25) p4[lane] = i4 - c4;
                if (lane == 31) band[5] = i4;
            } else if (fz.fuse && warp == 9) {             // row-major prefix of D3 (100 words)
                int run = 0;
                for (int base = 0; base < 100; base += 32) {
                    const int i = base + lane, c = i < 100 ? __popc(d3[i]) : 0;
                    int i3 = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, i3, o); if (lane >= o) i3 += t; }
                    if (i < 100) p3[i] = run + i3 - c;
                    run += __shfl_sync(0xffffffffu, i3, 31);
                }
                if (lane == 0) band[4] = run;
            }
        }
        __syncthreads();
        ST_STAMP(4);
        const int n3 = fz.fuse ? band[4] : 0, n4 = fz.fuse ? band[5] : 0;
        // the compact path: every level's dirty cells fit the shared-memory lists (one band)
        const bool compact = fz.fuse && rb1[200] <= ST_CAP1 && rb2[100] <= ST_VCAP && n3 <= ST_VCAP;
        uint16_t *l3 = compact ? l2 + ST_VCAP : reinterpret_cast<uint16_t *>(st_smem + StSmem::off_maps + StSmem::a_bytes);
        uint16_t *l4 = compact ? reinterpret_cast<uint16_t *>(st_smem + StSmem::off_l4) : l3 + 2500;
        const bool need_b34 = n3 > 0;                       // (D4 is a window of D3: n3 == 0 implies n4 == 0)
        bool b34_issued = false;
        if (fz.fuse && !compact && !scratch_ready) {       // this CTA's pool2 / pool3 images start out as the empty arena's
            for (int i = tid; i < 100 * 100; i += ST_NT) scr2[i] = __ldg(bg2g + st_cls(i / 100, 100) * 3 + st_cls(i % 100, 100));
            for (int i = tid; i < 50 * 50; i += ST_NT) scr3[i] = __ldg(bg3g + st_cls(i / 50, 50) * 3 + st_cls(i % 50, 50));
            scratch_ready = true;
        }
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
                    const int wd = Y0 * 4 + i;
                    st_list_word<4, 100>(d2[wd], wd, p2[wd] - base2, l2);
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
                if (compact) {                             // the lists of levels 3 and 4, by the threads with the fewest items above
                    const int t2 = ST_NT - 1 - tid;
                    if (t2 < 100) st_list_word<2, 50>(d3[t2], t2, p3[t2], l3);
                    else if (t2 < 125) st_list_word<1, 25>(d4[t2 - 100], t2 - 100, p4[t2 - 100], l4);
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
                if (stamps && blockIdx.x == 0 && tid == 0 && it < 8) {
                    stamps[it * 16 + 8] = n1; stamps[it * 16 + 9] = n2; stamps[it * 16 + 10] = n3; stamps[it * 16 + 11] = n4;
                }
                // ---- 4. conv2 + pool on the tensor pipe, 128 dirty cells per MMA tile
                const StGatherCompact<200, 100, 7, true> g2 = {d1, p1, base1, v1, bgc1, l2, atile};
                uint4 *dst2 = compact ? v2 : (fz.fuse ? scr2 : reinterpret_cast<uint4 *>(dsta));
                const bool pf = compact && need_b34;
                st_tiles(g2, l2, n2, atile, b16, b2, dst2, !compact, tmem_base, &mbar[1], &mbar[3], gt, nullptr, 0u, pf ? &mbar[5] : nullptr,
                         st_smem + StSmem::off_b34, w.c3_st, w.c4_st, tstamp);
                b34_issued = b34_issued || pf;
            }
            Y0 = Y1 + 1;
            nband++;
            __syncthreads();                               // everyone has read the band's bounds before thread 0 writes the next ones
        }
        ST_STAMP(7);
        if (fz.fuse) {
            if (!compact) {
                // fallback: the lists of levels 3 / 4 behind the A tile (the maps are dead)
                if (tid < 100) st_list_word<2, 50>(d3[tid], tid, p3[tid], l3);
                else if (tid >= 128 && tid < 153) st_list_word<1, 25>(d4[tid - 128], tid - 128, p4[tid - 128], l4);
            }
            if (need_b34 && !b34_issued && tid == 0) {     // conv3's and conv4's B operands over V1 (dead: the barrier above)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(&mbar[5], 2 * StSmem::b_bytes);
                bulk_g2s(st_smem + StSmem::off_b34, w.c3_st, StSmem::b_bytes, &mbar[5]);
                bulk_g2s(st_smem + StSmem::off_b34 + StSmem::b_bytes, w.c4_st, StSmem::b_bytes, &mbar[5]);
            }
            __syncthreads();
            uint4 *flat_a = reinterpret_cast<uint4 *>(fz.flat + (size_t)a * POL_FLAT_PITCH);
            // ---- 5. level 3: conv3 on the dirty pool3 cells;  6. level 4: conv4 on the dirty pool4 cells, straight into `flat`
            //         (NHWC flatten = 16 bytes per cell, raster order)
            if (compact) {
                const StGatherCompact<100, 50, 4, false> g3 = {d2, p2, 0, v2, bgc2, l3, atile};
                st_tiles(g3, l3, n3, atile, b34_16, b34, v3, false, tmem_base, &mbar[1], &mbar[3], gt, &mbar[5], n_b34 & 1u, nullptr, nullptr,
                         nullptr, nullptr, tstamp ? tstamp + 4 : nullptr);
                ST_STAMP(12);
                const StGatherCompact<50, 25, 2, false> g4 = {d3, p3, 0, v3, bgc3, l4, atile};
                st_tiles(g4, l4, n4, atile, b34_16 + StSmem::b_bytes / 16, b34 + 8, flat_a, true, tmem_base, &mbar[1], &mbar[3], gt, nullptr, 0u,
                         nullptr, nullptr, nullptr, nullptr, tstamp ? tstamp + 8 : nullptr);
            } else {
                const StGatherDense<100, 50> g3 = {scr2, l3, atile};
                st_tiles(g3, l3, n3, atile, b34_16, b34, scr3, true, tmem_base, &mbar[1], &mbar[3], gt, &mbar[5], n_b34 & 1u, nullptr, nullptr,
                         nullptr, nullptr, nullptr);
                ST_STAMP(12);
                const StGatherDense<50, 25> g4 = {scr3, l4, atile};
                st_tiles(g4, l4, n4, atile, b34_16 + StSmem::b_bytes / 16, b34 + 8, flat_a, true, tmem_base, &mbar[1], &mbar[3], gt, nullptr, 0u,
                         nullptr, nullptr, nullptr, nullptr, nullptr);
            }
            if (need_b34) n_b34++;
            ST_STAMP(13);
            // ---- 7. validation taps (dense pool2 / pool3)
            if (fz.tap2) {
                uint4 *t2 = reinterpret_cast<uint4 *>(fz.tap2) + (size_t)a * 10000, *t3 = reinterpret_cast<uint4 *>(fz.tap3) + (size_t)a * 2500;
                if (compact) {
                    const int n2 = rb2[100];
                    for (int i = tid; i < 10000; i += ST_NT) t2[i] = bgc2[st_cls(i / 100, 100) * 3 + st_cls(i % 100, 100)];
                    for (int i = tid; i < 2500; i += ST_NT) t3[i] = bgc3[st_cls(i / 50, 50) * 3 + st_cls(i % 50, 50)];
                    __syncthreads();
                    for (int e = tid; e < n2; e += ST_NT) t2[l2[e]] = v2[e];
                    for (int e = tid; e < n3; e += ST_NT) t3[l3[e]] = v3[e];
                } else {
                    for (int i = tid; i < 10000; i += ST_NT) t2[i] = __ldcg(scr2 + i);
                    for (int i = tid; i < 2500; i += ST_NT) t3[i] = __ldcg(scr3 + i);
                }
                __syncthreads();
            }
            if (!compact) {                                // fallback: the dirty cells of both images go back to their empty-arena values
                for (int i = tid; i < 100 * 4; i += ST_NT) {
                    uint32_t bits = d2[i];
                    const int Y = i >> 2;
                    while (bits) {
                        const int X = 32 * (i & 3) + __ffs(bits) - 1;
                        bits &= bits - 1;
                        scr2[Y * 100 + X] = __ldg(bg2g + st_cls(Y, 100) * 3 + st_cls(X, 100));
                    }
                }
                for (int e = tid; e < n3; e += ST_NT) {
                    const int cell = l3[e];
                    scr3[cell] = __ldg(bg3g + st_cls(cell / 50, 50) * 3 + st_cls(cell % 50, 50));
                }
            }
            __syncthreads();                               // nobody still reads this arena's lists / bitmaps / V lists
        }
        ST_STAMP(14);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u));
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

// Tensor-core engine of the policy forward: 3x3 convolutions as tcgen05 "shifted GEMMs" (sm_100a).
//
// Activations are NHWC bf16 with 8 channels = one 16-byte row of a UMMA core matrix per pixel.  A CTA
// stages a strip of the input image in shared memory as a LINEAR pixel array with pitch P (one or two
// halo pixels per row) and never builds an im2col matrix: for the tile of 128 consecutive output
// positions q = 128 t .. 128 t + 127, tap (dy, dx) of the 3x3 stencil is the same array shifted by
// dy*P + dx pixels, so the A operand of every tcgen05.mma is just a shared-memory descriptor
// (K-major, no swizzle: 8 pixels x 16 B = one core matrix, SBO = 128 B between 8-pixel groups,
// LBO = distance between the two taps that make up one K = 16 step).  Nine taps + one zero tap =
// five MMAs (M = 128, N = 16 or 32, K = 16) per tile, accumulated in TMEM; the 128 threads then read
// their accumulator row with tcgen05.ld and run the fused epilogue:
//   conv2/3/4 : + bias, ReLU, bf16 -> smem stage -> 2x2 max-pool -> HBM
//   upconv3   : bilinear x2 folded into 4 output phases (N = 4 x 8): + bias, ReLU -> 4 pixels of HBM
//   upconv4   : 4 phases of the single output channel: + bias -> running argmax (+ optional dense map)
// conv1 never touches HBM: its pooled output is generated straight into conv2's shared-memory strip
// from the bit maps (background constant + exact evaluation near set bits).
// The input rows reach shared memory as bulk asynchronous copies (cp.async.bulk + mbarrier).  A dedicated
// warp issues the MMAs into a ring of 8 TMEM accumulators (full / empty mbarriers per accumulator) while
// eight warps drain completed tiles.  Measured on B200: these kernels are bound by instruction issue and by
// the tensor pipe's shared-memory operand reads (~75 cycles per M=128 x K=16 A tile whatever N is), not by
// its math; folding the dx taps into N (2 MMAs per tile instead of 5) was tried and lost to its heavier
// epilogue (shuffles + lane-quarter exchange) -- see DESIGN.md section 5.
#include "ofb_common.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_tc_ptx.cuh"

enum { M_CONV_GMEM = 0, M_CONV_BITS = 1, M_UP3 = 2, M_UP4 = 3 };
#define TC_R 10                       // image rows per strip (all layer heights are multiples of 10)
#define TC_NT 288                     // threads per CTA: 8 draining warps (two per TMEM lane quarter) + 1 MMA-issuer warp
#define TC_RING 8                     // TMEM accumulators (tiles in flight) per CTA
#define TC_BITS_WORDS 352             // words of one bit map staged per strip: 27 rows x 50 B + alignment slack

struct TcArgs {
    const void *in;                   // bf16 [item][H][H][8], or uint32 bit maps [item][2][5000]
    const __nv_bfloat16 *wt;          // [10][N][8] B-operand image
    const float *bias;                // [N]
    const float *aux_w, *aux_b;       // conv1 fp32 weights (BITS) / un-phased fp32 weights (ring of UP3, UP4)
    __nv_bfloat16 *out;               // pooled / upsampled activations
    float *ptr_out;                   // UP4: optional dense map [item][400][400]
    float *amax_val;                  // UP4: [item][gridDim.x]
    int *amax_idx;
    long long in_item_stride, out_item_stride;   // in elements
    int H;                            // input height = width (conv resolution)
    long long *dbg;                   // optional: clock64 stamps of CTA (1, 0) at the phase boundaries
};

// shared-memory plan of one CTA (host and device agree through this)
struct TcPlan {
    int P, tiles, sin_pixels;
    unsigned off_sin, off_stage, off_aux, off_bar, total;
};
__host__ __device__ inline TcPlan tc_plan(int mode, int N, int H) {
    TcPlan p;
    p.P = H + ((mode == M_UP3 || mode == M_UP4) ? 2 : 1);
    p.tiles = (TC_R * p.P + 127) / 128;
    p.sin_pixels = 128 * p.tiles + 2 * p.P + 8;
    p.off_sin = (unsigned)((POL_TAPS * N * 16 + 127) & ~127);
    p.off_stage = p.off_sin + (unsigned)p.sin_pixels * 16;
    const unsigned stage = (mode == M_CONV_GMEM || mode == M_CONV_BITS) ? (unsigned)(TC_R * p.P) * 16      // (BITS: also holds the bit rows first)
                                                                         : (unsigned)(4 * TC_R + 4 * H) * (mode == M_UP3 ? 32 : 4);   // ring buffer
    p.off_aux = p.off_stage + stage;
    // aux: ring weights (<= 288 floats) + argmax scratch, or conv1 bias + the strip's bit rows of both maps
    p.off_bar = p.off_aux + (mode == M_CONV_BITS ? 64 : 1280);
    p.total = p.off_bar + (2 * TC_RING + 1) * 8 + 16;   // full[RING] + load + empty[RING] mbarriers, TMEM slot
    return p;
}

// low-res strip in shared memory as an image accessor (rows y0-1 .. y0+R, cols -1 .. W, replicated halo)
struct StripImage {
    const uint4 *sin;
    int P, y0;
    __device__ __forceinline__ uint4 operator()(int y, int x) const { return sin[(y - y0 + 1) * P + x + 1]; }
};
// slot of a ring pixel in the strip's ring buffer: left column, right column, then the top / bottom row by X
__device__ __forceinline__ int ring_slot(int y_local, int X, int Wo) {
    return X == 0 ? y_local : (X == Wo - 1 ? 2 * TC_R + y_local : 4 * TC_R + X);
}

template <int MODE, int N>
__global__ void __launch_bounds__(TC_NT)
k_tc_conv(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr bool REPL = (MODE == M_UP3 || MODE == M_UP4);
    constexpr bool POOL = (MODE == M_CONV_GMEM || MODE == M_CONV_BITS);
    const int W = a.H;
    const TcPlan pl = tc_plan(MODE, N, W);
    const int P = pl.P, T = pl.tiles;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int item = blockIdx.y, y0 = blockIdx.x * TC_R;

    uint4 *sw = reinterpret_cast<uint4 *>(smem);
    uint4 *sin = reinterpret_cast<uint4 *>(smem + pl.off_sin);
    uint4 *stage = reinterpret_cast<uint4 *>(smem + pl.off_stage);
    float *ring = reinterpret_cast<float *>(smem + pl.off_stage);     // UP modes reuse the stage region
    float *aux = reinterpret_cast<float *>(smem + pl.off_aux);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + pl.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + pl.off_bar + (2 * TC_RING + 1) * 8);
    uint64_t *ebars = bars + TC_RING + 1;                 // accumulator b drained by its 4 warps -> may be overwritten

#define TC_STAMP(k) do { if (a.dbg && tid == 0 && blockIdx.x == 1 && blockIdx.y == 0) a.dbg[k] = clock64(); } while (0)
    TC_STAMP(0);
    // ---- one-time setup: barriers, bulk copies of the input rows, weights
    uint64_t *lbar = &bars[TC_RING];                     // completion of the bulk (TMA-class) row copies
    const int rows_in = TC_R + 2;
    const uint8_t *in_item = reinterpret_cast<const uint8_t *>(a.in) +
                             (MODE == M_CONV_BITS ? (size_t)item * 2 * POL_WORDS * 4 : (size_t)item * a.in_item_stride * 2);
    uint32_t *sbits = reinterpret_cast<uint32_t *>(smem + pl.off_stage);  // BITS: 2 x TC_BITS_WORDS words of the two maps (dead before the epilogue)
    uint2 *wl = reinterpret_cast<uint2 *>(smem + pl.off_stage + 2 * TC_BITS_WORDS * 4);   // BITS: work list (patch, sin index)
    int *wl_count = reinterpret_cast<int *>(aux + 8);
    int bits_w0 = 0;                                     // first map word held in sbits
    if (MODE == M_CONV_BITS) {
        const int r0 = max(2 * y0 - 3, 0), r1 = min(2 * (y0 + TC_R) + 2, POL_W - 1);
        bits_w0 = ((r0 * 50) & ~15) >> 2;                // rows are 50 B; bulk copies move 16 B units
        if (tid == 32) {
            const int b0 = bits_w0 * 4, b1 = min(((r1 + 1) * 50 + 15 + 16) & ~15, POL_WORDS * 4);
            for (int t = 0; t <= 2 * TC_RING; t++) mbar_init(&bars[t], t > TC_RING ? 4 : 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(lbar, 2u * (uint32_t)(b1 - b0));
            bulk_g2s(sbits, in_item + b0, (uint32_t)(b1 - b0), lbar);
            bulk_g2s(sbits + TC_BITS_WORDS, in_item + POL_WORDS * 4 + b0, (uint32_t)(b1 - b0), lbar);
        }
    } else if (tid == 32) {
        for (int t = 0; t <= 2 * TC_RING; t++) mbar_init(&bars[t], t > TC_RING ? 4 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int nrows = 0;
        for (int ry = 0; ry < rows_in; ry++) {
            const int y = y0 - 1 + ry;
            nrows += (REPL || (y >= 0 && y < W)) ? 1 : 0;
        }
        mbar_expect_tx(lbar, (uint32_t)(nrows * W * 16));
        for (int ry = 0; ry < rows_in; ry++) {
            int y = y0 - 1 + ry;
            if (!REPL && (y < 0 || y >= W)) continue;
            y = min(max(y, 0), W - 1);
            bulk_g2s(sin + ry * P + 1, in_item + (size_t)y * W * 16, (uint32_t)(W * 16), lbar);
        }
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wt);
        for (int i = tid; i < POL_TAPS * N; i += TC_NT) sw[i] = src[i];
    }
    if (MODE == M_CONV_BITS) {
        if (tid < 8) aux[tid] = a.aux_b[tid];            // conv1 bias; its pattern LUT (a.aux_w) stays in global/L1
        if (tid == 8) *wl_count = 0;
    } else if (MODE == M_UP3) {
        for (int i = tid; i < 9 * 4 * 8; i += TC_NT) aux[i] = a.aux_w[i];
    } else if (MODE == M_UP4) {
        for (int i = tid; i < 72; i += TC_NT) aux[i] = a.aux_w[i];
    }
    float biasr[POOL ? 8 : (MODE == M_UP3 ? 32 : 4)];
#pragma unroll
    for (int i = 0; i < (int)(sizeof(biasr) / sizeof(float)); i++) biasr[i] = a.bias[i];

    // ---- halo pixels, out-of-image rows and the tail (generic stores; the rows themselves arrive by bulk copy)
    if (MODE != M_CONV_BITS) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in_item);
        for (int i = tid; i < rows_in * 2; i += TC_NT) {
            const int ry = i >> 1, right = i & 1;
            const int y = min(max(y0 - 1 + ry, 0), W - 1);
            if (REPL) sin[ry * P + (right ? P - 1 : 0)] = src[(size_t)y * W + (right ? W - 1 : 0)];
            else if (!right) sin[ry * P] = make_uint4(0, 0, 0, 0);
        }
        if (!REPL)
            for (int ry = 0; ry < rows_in; ry += rows_in - 1) {           // only the first / last staged row can be outside
                const int y = y0 - 1 + ry;
                if (y >= 0 && y < W) continue;
                for (int c = tid; c < W; c += TC_NT) sin[ry * P + 1 + c] = make_uint4(0, 0, 0, 0);
            }
    }
    for (int i = rows_in * P + tid; i < pl.sin_pixels; i += TC_NT) sin[i] = make_uint4(0, 0, 0, 0);

    if (MODE == M_CONV_BITS) {
        __syncthreads();                                 // barrier init + aux visible
        TC_STAMP(1);
        mbar_wait(lbar, 0);                              // bit rows have landed
        TC_STAMP(2);
        const uint32_t *smap = sbits - bits_w0, *lmap = sbits + TC_BITS_WORDS - bits_w0;
        float bg[8];
#pragma unroll
        for (int co = 0; co < 8; co++) bg[co] = fmaxf(aux[co], 0.f);
        const uint4 bgq = pack_bf8(bg);
        const int groups = W / 8;                        // 8 pooled pixels per work item
        for (int g = tid; g < rows_in * groups; g += TC_NT) {
            const int ry = g / groups, gx = g % groups, py = y0 - 1 + ry;
            uint4 *dst = sin + ry * P + 1 + gx * 8;
            const int rot = tid & 7;                     // rotate the store order: 8 lanes hit 8 different bank groups
            if (py < 0 || py >= W) {
#pragma unroll
                for (int k = 0; k < 8; k++) dst[(k + rot) & 7] = make_uint4(0, 0, 0, 0);
                continue;
            }
            // 18 map columns 16 gx - 1 .. 16 gx + 16 of the 4 map rows 2 py - 1 .. 2 py + 2
            uint32_t rs[4], rl[4], any = 0;
            uint32_t colmask = 0x3FFFFu;
            if (gx == 0) colmask &= ~1u;
            if (gx == groups - 1) colmask &= ~(1u << 17);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int r = 2 * py - 1 + i;
                rs[i] = rl[i] = 0;
                if (r >= 0 && r < POL_W) {
                    int b = r * POL_W + 16 * gx - 1;
                    const int sh = b < 0 ? 1 : 0;
                    b = max(b, 0);
                    const int wd = b >> 5;
                    rs[i] = (__funnelshift_r(smap[wd], smap[wd + 1], b & 31) << sh) & colmask;
                    rl[i] = (__funnelshift_r(lmap[wd], lmap[wd + 1], b & 31) << sh) & colmask;
                }
                any |= rs[i] | rl[i];
            }
            if (!any) {
#pragma unroll
                for (int k = 0; k < 8; k++) dst[(k + rot) & 7] = bgq;
                continue;
            }
            // pixels that see a set bit go to a work list, so that the (rare, heavier) exact evaluations are
            // spread over all threads afterwards instead of serialising in the few threads that met them
            for (int kk = 0; kk < 8; kk++) {
                const int k = (kk + rot) & 7;
                uint32_t ps = 0, pq = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ps |= ((rs[i] >> (2 * k)) & 0xFu) << (4 * i);
                    pq |= ((rl[i] >> (2 * k)) & 0xFu) << (4 * i);
                }
                if ((ps | pq) == 0) { dst[k] = bgq; continue; }
                const int slot = atomicAdd(wl_count, 1);
                wl[slot] = make_uint2(ps | (pq << 16), (uint32_t)(ry * P + 1 + gx * 8 + k));
            }
        }
        __syncthreads();
        {
            const int n_work = *wl_count;
            for (int i = tid; i < n_work; i += TC_NT) {
                const uint2 e = wl[i];
                float v[8];
                conv1_pool_pixel(e.x & 0xFFFFu, e.x >> 16, a.aux_w, aux, v);
                sin[e.y] = pack_bf8(v);
            }
        }
        for (int ry = tid; ry < rows_in; ry += TC_NT) sin[ry * P] = make_uint4(0, 0, 0, 0);    // shared halo column
    }
    TC_STAMP(3);
    // ---- TMEM: a ring of TC_RING accumulators of N columns each (allocated late so that a CTA waiting for
    //      columns has already staged its strip)
    uint32_t TMEM_COLS = 32;
    while (TMEM_COLS < (uint32_t)(min(T, TC_RING) * N)) TMEM_COLS <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    TC_STAMP(4);
    if (MODE != M_CONV_BITS) mbar_wait(lbar, 0);         // bulk-copied rows have landed (acquire for every thread)
    const uint32_t tmem_base = *tmem_slot;
    TC_STAMP(5);

    // ---- descriptors
    constexpr uint32_t IDESC = instr_desc(N);
    const uint32_t sin16 = smem_u32(sin) >> 4, sw16 = smem_u32(sw) >> 4;
    // called by the whole (converged) issuer warp: the descriptors are warp-uniform values, one elected lane issues.
    // (Issuing from inside an `if (lane == 0)` region costs ~80 cycles per MMA in register -> uniform-register moves;
    // this form issues back to back and leaves the tensor pipe's operand reads, ~39 cycles per MMA, as the limit.)
    auto issue_tile = [&](int t, bool leader) {
        const uint32_t d = tmem_base + (uint32_t)((t % TC_RING) * N);
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const int t0 = 2 * j, t1 = 2 * j + 1;
            const int off0 = (t0 / 3) * P + (t0 % 3);
            const int off1 = t1 < 9 ? (t1 / 3) * P + (t1 % 3) : off0 + 1;     // tap 9: zero weights
            const uint64_t ad = smem_desc(sin16 + (uint32_t)(128 * t + off0), (uint32_t)(off1 - off0), 8);
            const uint64_t bd = smem_desc(sw16 + (uint32_t)(t0 * N), (uint32_t)N, 8);
            if (leader) tc_mma(d, ad, bd, IDESC, j > 0 ? 1u : 0u);
        }
        if (leader) tc_commit(&bars[t % TC_RING]);
    };

    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    // warp 8 = MMA issuer: queues tile t as soon as accumulator t % TC_RING has been drained (empty barrier);
    // warps 0-7 = drain: warps 0-3 take even tiles, warps 4-7 odd tiles, one TMEM lane quarter each
    if (warp == 8) {
        const bool leader = elect_one();
        for (int t = 0; t < T; t++) {
            if (t >= TC_RING) {
                mbar_wait(&ebars[t % TC_RING], (uint32_t)(((t / TC_RING) - 1) & 1));
                tc_fence_after();
            }
            issue_tile(t, leader);
        }
        __syncwarp();
    }
    TC_STAMP(6);
    for (int t = 0; t < T && warp < 8; t++) {
        if ((t & 1) != (warp >> 2)) continue;
        mbar_wait(&bars[t % TC_RING], (uint32_t)((t / TC_RING) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((t % TC_RING) * N);
        const int q = 128 * t + (tid & 127);
        if (POOL) {
            uint32_t r[8];
            tc_ld8(taddr, r);
            tc_wait_ld();
            float v[8];
#pragma unroll
            for (int co = 0; co < 8; co++) v[co] = fmaxf(__uint_as_float(r[co]) + biasr[co], 0.f);
            if (q < TC_R * P) stage[q] = pack_bf8(v);
        } else if (MODE == M_UP3) {
            uint32_t r[32];
            tc_ld8(taddr, r); tc_ld8(taddr + 8, r + 8); tc_ld8(taddr + 16, r + 16); tc_ld8(taddr + 24, r + 24);
            tc_wait_ld();
            const int rr = q / P, c = q % P;
            if (rr < TC_R && c < W) {
                const int i = y0 + rr, j = c;
                __nv_bfloat16 *dst = a.out + (size_t)item * a.out_item_stride;
#pragma unroll
                for (int ph = 0; ph < 4; ph++) {
                    const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
                    float o[8];
#pragma unroll
                    for (int co = 0; co < 8; co++) o[co] = __uint_as_float(r[ph * 8 + co]) + biasr[ph * 8 + co];
                    if (Y == 0 || Y == 2 * W - 1 || X == 0 || X == 2 * W - 1) {      // ring: corrected after the tile loop
                        float *rb = ring + 8 * ring_slot(Y - 2 * y0, X, 2 * W);
#pragma unroll
                        for (int co = 0; co < 8; co++) rb[co] = o[co];
                        continue;
                    }
#pragma unroll
                    for (int co = 0; co < 8; co++) o[co] = fmaxf(o[co], 0.f);
                    *reinterpret_cast<uint4 *>(dst + pol_plane200_off(Y, X)) = pack_bf8(o);      // plane layout for k_tz_up4
                }
            }
        } else {    // M_UP4
            uint32_t r[4];
            tc_ld4(taddr, r);
            tc_wait_ld();
            const int rr = q / P, c = q % P;
            if (rr < TC_R && c < W) {
                const int i = y0 + rr, j = c;
#pragma unroll
                for (int ph = 0; ph < 4; ph++) {
                    const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
                    float o = __uint_as_float(r[ph]) + biasr[ph];
                    if (Y == 0 || Y == 2 * W - 1 || X == 0 || X == 2 * W - 1) { ring[ring_slot(Y - 2 * y0, X, 2 * W)] = o; continue; }
                    const int idx = Y * 2 * W + X;
                    if (a.ptr_out) a.ptr_out[(size_t)item * 4 * W * W + idx] = o;
                    if (amax_better(o, idx, best_v, best_i)) { best_v = o; best_i = idx; }
                }
            }
        }
        tc_fence_before();                               // this warp's TMEM reads of the tile are complete
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&ebars[t % TC_RING]);
    }
    TC_STAMP(7);
    tc_fence_before();
    __syncthreads();                                     // stage rows / ring values are visible to every thread
    TC_STAMP(8);

    if (REPL) {
        // border ring: take the out-of-range taps back out of the folded result, then consume the pixel.
        // Four lanes per ring pixel: lanes 0..2 evaluate one out-of-range tap each (lane 0 also the two extra
        // taps of a corner), summed in lane order by shuffles.
        constexpr int CIN = (MODE == M_UP3) ? 4 : 8, COUT = (MODE == M_UP3) ? 8 : 1;
        const int Wo = 2 * W, nslots = 4 * TC_R + Wo;
        const bool first = blockIdx.x == 0, last = blockIdx.x == gridDim.x - 1;
        const StripImage Ls{sin, P, y0};
        for (int base = 0; base < nslots * 4; base += TC_NT) {
            const int itm = base + tid, slot = itm >> 2, j = itm & 3;
            int Y = 0, X = 0;
            bool valid = slot < nslots;
            if (slot < 2 * TC_R) { Y = 2 * y0 + slot; X = 0; }
            else if (slot < 4 * TC_R) { Y = 2 * y0 + slot - 2 * TC_R; X = Wo - 1; }
            else {
                X = slot - 4 * TC_R;
                valid = valid && (first || last) && X != 0 && X != Wo - 1;
                Y = first ? 0 : Wo - 1;
            }
            float acc[COUT];
#pragma unroll
            for (int co = 0; co < COUT; co++) acc[co] = 0.f;
            if (valid && j < 3) {
                const bool xedge = (X == 0 || X == Wo - 1), yedge = (Y == 0 || Y == Wo - 1);
                const int dxe = X == 0 ? 0 : 2, dye = Y == 0 ? 0 : 2;
                if (xedge) up_ring_tap<CIN, COUT>(Ls, Y, X, j, dxe, aux, acc);
                else up_ring_tap<CIN, COUT>(Ls, Y, X, dye, j, aux, acc);
                if (xedge && yedge && j == 0)                       // corner: the two remaining taps of the edge row
                    for (int dx = 0; dx < 3; dx++)
                        if (dx != dxe) up_ring_tap<CIN, COUT>(Ls, Y, X, dye, dx, aux, acc);
            }
            float o[COUT];
#pragma unroll
            for (int co = 0; co < COUT; co++) {
                const float t1 = __shfl_down_sync(0xffffffffu, acc[co], 1), t2 = __shfl_down_sync(0xffffffffu, acc[co], 2);
                o[co] = (acc[co] + t1) + t2;
            }
            if (!valid || j != 0) continue;
            if (MODE == M_UP3) {
#pragma unroll
                for (int co = 0; co < COUT; co++) o[co] = fmaxf(ring[COUT * slot + co] - o[co], 0.f);
                float o8[8];
#pragma unroll
                for (int co = 0; co < 8; co++) o8[co] = o[co % COUT];
                *reinterpret_cast<uint4 *>(a.out + (size_t)item * a.out_item_stride + pol_plane200_off(Y, X)) = pack_bf8(o8);
            } else {
                const float v = ring[slot] - o[0];
                const int idx = Y * Wo + X;
                if (a.ptr_out) a.ptr_out[(size_t)item * Wo * Wo + idx] = v;
                if (amax_better(v, idx, best_v, best_i)) { best_v = v; best_i = idx; }
            }
        }
    }
    if (POOL) {
        const int wo = W / 2;
        __nv_bfloat16 *dst = a.out + (size_t)item * a.out_item_stride;
        for (int pp = tid; pp < (TC_R / 2) * wo; pp += TC_NT) {
            const int pr = pp / wo, pc = pp % wo;
            const uint4 q0 = stage[(2 * pr) * P + 2 * pc], q1 = stage[(2 * pr) * P + 2 * pc + 1];
            const uint4 q2 = stage[(2 * pr + 1) * P + 2 * pc], q3 = stage[(2 * pr + 1) * P + 2 * pc + 1];
            uint4 o;
            const uint32_t *p0 = &q0.x, *p1 = &q1.x, *p2 = &q2.x, *p3 = &q3.x;
            uint32_t *po = &o.x;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                __nv_bfloat162 m = __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162 *>(p0 + k), *reinterpret_cast<const __nv_bfloat162 *>(p1 + k)),
                                           __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(p2 + k), *reinterpret_cast<const __nv_bfloat162 *>(p3 + k)));
                po[k] = *reinterpret_cast<uint32_t *>(&m);
            }
            *reinterpret_cast<uint4 *>(dst + ((size_t)(y0 / 2 + pr) * wo + pc) * 8) = o;
        }
    }
    if (MODE == M_UP4) {
        float *sv = aux + 128;
        int *si = reinterpret_cast<int *>(aux + 144);
        amax_warp(best_v, best_i);
        if ((tid & 31) == 0) { sv[warp] = best_v; si[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < TC_NT / 32; k++)
                if (amax_better(sv[k], si[k], best_v, best_i)) { best_v = sv[k]; best_i = si[k]; }
            a.amax_val[(size_t)item * gridDim.x + blockIdx.x] = best_v;
            a.amax_idx[(size_t)item * gridDim.x + blockIdx.x] = best_i;
        }
    }
    // ---- teardown
    TC_STAMP(9);
    tc_fence_before();
    __syncthreads();
    TC_STAMP(10);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// ---------------------------------------------------------------- host launchers
static long long *g_tc_dbg = nullptr;                    // device buffer of 4 x 16 stamps (mode-major), see ofb_policy_tc_debug
extern "C" int ofb_policy_tc_debug(long long *dev_buf) { g_tc_dbg = dev_buf; return OFB_OK; }

template <int MODE, int N>
static int launch(const TcArgs &a, int n_items, cudaStream_t st) {
    const TcPlan pl = tc_plan(MODE, N, a.H);
    static thread_local unsigned configured = 0;
    if (pl.total > configured) {
        OFB_CUDA_CHECK(cudaFuncSetAttribute(k_tc_conv<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total));
        configured = pl.total;
    }
    if (n_items == 0) return OFB_OK;
    TcArgs b = a;
    b.dbg = g_tc_dbg ? g_tc_dbg + 16 * MODE : nullptr;
    k_tc_conv<MODE, N><<<dim3(a.H / TC_R, n_items), TC_NT, pl.total, st>>>(b);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

int pol_tc_conv_pool(const ofb_policy *p, int layer, const __nv_bfloat16 *in, __nv_bfloat16 *out, int hin, int n_items,
                     long long out_item_stride, cudaStream_t st) {
    TcArgs a = {};
    a.in = in; a.wt = p->w.cw[layer]; a.bias = p->w.cb[layer]; a.out = out;
    a.in_item_stride = (long long)hin * hin * 8; a.out_item_stride = out_item_stride; a.H = hin;
    return launch<M_CONV_GMEM, 16>(a, n_items, st);
}

int pol_tc_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st) {
    TcArgs a = {};
    a.in = maps; a.wt = p->w.cw[0]; a.bias = p->w.cb[0]; a.aux_w = p->w.c1_lut; a.aux_b = p->w.c1_b; a.out = out;
    a.out_item_stride = 100 * 100 * 8; a.H = 200;
    return launch<M_CONV_BITS, 16>(a, n_items, st);
}

int pol_tc_up3(const ofb_policy *p, const __nv_bfloat16 *in, __nv_bfloat16 *out, int n_items, cudaStream_t st) {
    TcArgs a = {};
    a.in = in; a.wt = p->w.u3_pw; a.bias = p->w.u3_pb; a.aux_w = p->w.u3_w; a.aux_b = p->w.u3_b; a.out = out;
    a.in_item_stride = 100 * 100 * 8; a.out_item_stride = POL_UP3_ITEM; a.H = 100;
    return launch<M_UP3, 32>(a, n_items, st);
}

int pol_tc_up4(const ofb_policy *p, const __nv_bfloat16 *in, float *ptr_out, float *amax_val, int *amax_idx, int n_items,
               cudaStream_t st) {
    TcArgs a = {};
    a.in = in; a.wt = p->w.u4_pw; a.bias = p->w.u4_pb; a.aux_w = p->w.u4_w; a.aux_b = p->w.u4_b;
    a.ptr_out = ptr_out; a.amax_val = amax_val; a.amax_idx = amax_idx;
    a.in_item_stride = 200 * 200 * 8; a.H = 200;
    return launch<M_UP4, 16>(a, n_items, st);
}

int pol_tc_up4_parts() { return 200 / TC_R; }

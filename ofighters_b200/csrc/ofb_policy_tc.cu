// conv3 / conv4 of the policy trunk (qlearnIA_V2.py:141-151) on tensor cores: 3x3 'same' convolution 8 -> 8 + BN + ReLU +
// 2x2 max-pool as a tcgen05 "shifted GEMM" (sm_100a).  The large layers (conv1+conv2, upconv3, upconv4) use the
// block-Toeplitz kernels of ofb_policy_tz.cu; these two are 4 % of the forward's MACs and keep the simpler pixel-linear form.
//
// Activations are NHWC bf16 with 8 channels = one 16-byte row of a UMMA core matrix per pixel.  A CTA stages a strip of
// 10 image rows (+ halo) in shared memory as a LINEAR pixel array with pitch P = W + 1 (one shared zero halo pixel per
// row) and never builds an im2col matrix: for the tile of 128 consecutive output positions q = 128 t .. 128 t + 127, tap
// (dy, dx) of the stencil is the same array shifted by dy * P + dx pixels, so the A operand of every tcgen05.mma is just a
// shared-memory descriptor (K-major, no swizzle: 8 pixels x 16 B = one core matrix, SBO = 128 B between 8-pixel groups,
// LBO = distance between the two taps that make up one K = 16 step).  Nine taps + one zero tap = five MMAs (M = 128,
// N = 16, K = 16) per tile, accumulated in TMEM.  Rows arrive as bulk asynchronous copies (cp.async.bulk + mbarrier); a
// converged warp issues the MMAs (one elected lane; descriptors stay in uniform registers) into a ring of 8 TMEM
// accumulators while eight warps drain completed tiles: + bias, ReLU, bf16 -> shared-memory stage -> 2x2 max-pool -> HBM.
#include "ofb_common.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_tc_ptx.cuh"

#define TC_R 10                       // image rows per strip (100 and 50 are multiples of 10)
#define TC_NT 288                     // 8 draining warps (two per TMEM lane quarter) + 1 MMA-issuer warp
#define TC_RING 8                     // TMEM accumulators (tiles in flight) per CTA
#define TC_N 16                       // MMA N: 8 output channels padded to 16

struct TcArgs {
    const __nv_bfloat16 *in;          // [item][H][H][8]
    const __nv_bfloat16 *wt;          // [10 taps][16][8] B-operand image
    const float *bias;                // [16]
    __nv_bfloat16 *out;               // pooled activations [item][H/2][H/2][8] (item stride given)
    long long out_item_stride;        // in elements
    int H;                            // input height = width
};

// shared-memory plan of one CTA (host and device agree through this)
struct TcPlan {
    int P, tiles, sin_pixels;
    unsigned off_sin, off_stage, off_bar, total;
};
__host__ __device__ inline TcPlan tc_plan(int H) {
    TcPlan p;
    p.P = H + 1;
    p.tiles = (TC_R * p.P + 127) / 128;
    p.sin_pixels = 128 * p.tiles + 2 * p.P + 8;
    p.off_sin = (unsigned)((POL_TAPS * TC_N * 16 + 127) & ~127);
    p.off_stage = p.off_sin + (unsigned)p.sin_pixels * 16;
    p.off_bar = p.off_stage + (unsigned)(TC_R * p.P) * 16;
    p.total = p.off_bar + (2 * TC_RING + 1) * 8 + 16;   // full[RING] + load + empty[RING] mbarriers, TMEM slot
    return p;
}

__global__ void __launch_bounds__(TC_NT)
k_tc_conv_pool(const TcArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int W = a.H;
    const TcPlan pl = tc_plan(W);
    const int P = pl.P, T = pl.tiles;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int item = blockIdx.y, y0 = blockIdx.x * TC_R;

    uint4 *sw = reinterpret_cast<uint4 *>(smem);
    uint4 *sin = reinterpret_cast<uint4 *>(smem + pl.off_sin);
    uint4 *stage = reinterpret_cast<uint4 *>(smem + pl.off_stage);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + pl.off_bar);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + pl.off_bar + (2 * TC_RING + 1) * 8);
    uint64_t *lbar = &bars[TC_RING];                     // completion of the bulk row copies
    uint64_t *ebars = bars + TC_RING + 1;                // accumulator b drained by its 4 warps -> may be overwritten

    // ---- barriers, bulk copies of the input rows, weights
    const int rows_in = TC_R + 2;
    const uint8_t *in_item = reinterpret_cast<const uint8_t *>(a.in) + (size_t)item * W * W * 16;
    if (tid == 32) {
        for (int t = 0; t <= 2 * TC_RING; t++) mbar_init(&bars[t], t > TC_RING ? 4 : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int nrows = 0;
        for (int ry = 0; ry < rows_in; ry++) nrows += (y0 - 1 + ry >= 0 && y0 - 1 + ry < W) ? 1 : 0;
        mbar_expect_tx(lbar, (uint32_t)(nrows * W * 16));
        for (int ry = 0; ry < rows_in; ry++) {
            const int y = y0 - 1 + ry;
            if (y < 0 || y >= W) continue;
            bulk_g2s(sin + ry * P + 1, in_item + (size_t)y * W * 16, (uint32_t)(W * 16), lbar);
        }
    }
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.wt);
        for (int i = tid; i < POL_TAPS * TC_N; i += TC_NT) sw[i] = src[i];
    }
    float biasr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) biasr[i] = a.bias[i];
    // ---- zero halo column, out-of-image rows and the tail (generic stores; the rows themselves arrive by bulk copy)
    for (int ry = tid; ry < rows_in; ry += TC_NT) sin[ry * P] = make_uint4(0, 0, 0, 0);
    for (int ry = 0; ry < rows_in; ry += rows_in - 1) {    // only the first / last staged row can be outside
        const int y = y0 - 1 + ry;
        if (y >= 0 && y < W) continue;
        for (int c = tid; c < W; c += TC_NT) sin[ry * P + 1 + c] = make_uint4(0, 0, 0, 0);
    }
    for (int i = rows_in * P + tid; i < pl.sin_pixels; i += TC_NT) sin[i] = make_uint4(0, 0, 0, 0);

    // ---- TMEM: a ring of TC_RING accumulators of 16 columns each
    uint32_t TMEM_COLS = 32;
    while (TMEM_COLS < (uint32_t)(min(T, TC_RING) * TC_N)) TMEM_COLS <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(lbar, 0);                                  // bulk-copied rows have landed (acquire for every thread)
    const uint32_t tmem_base = *tmem_slot;

    constexpr uint32_t IDESC = instr_desc(TC_N);
    const uint32_t sin16 = smem_u32(sin) >> 4, sw16 = smem_u32(sw) >> 4;
    // warp 8 = MMA issuer: queues tile t as soon as accumulator t % TC_RING has been drained (empty barrier).  The whole
    // warp stays converged so that the descriptors are warp-uniform values; one elected lane issues.
    if (warp == 8) {
        const bool leader = elect_one();
        for (int t = 0; t < T; t++) {
            if (t >= TC_RING) {
                mbar_wait(&ebars[t % TC_RING], (uint32_t)(((t / TC_RING) - 1) & 1));
                tc_fence_after();
            }
            const uint32_t d = tmem_base + (uint32_t)((t % TC_RING) * TC_N);
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const int t0 = 2 * j, t1 = 2 * j + 1;
                const int off0 = (t0 / 3) * P + (t0 % 3);
                const int off1 = t1 < 9 ? (t1 / 3) * P + (t1 % 3) : off0 + 1;     // tap 9: zero weights
                const uint64_t ad = smem_desc(sin16 + (uint32_t)(128 * t + off0), (uint32_t)(off1 - off0), 8);
                const uint64_t bd = smem_desc(sw16 + (uint32_t)(t0 * TC_N), (uint32_t)TC_N, 8);
                if (leader) tc_mma(d, ad, bd, IDESC, j > 0 ? 1u : 0u);
            }
            if (leader) tc_commit(&bars[t % TC_RING]);
        }
        __syncwarp();
    }
    // warps 0-7 = drain: warps 0-3 take even tiles, warps 4-7 odd tiles, one TMEM lane quarter each
    for (int t = 0; t < T && warp < 8; t++) {
        if ((t & 1) != (warp >> 2)) continue;
        mbar_wait(&bars[t % TC_RING], (uint32_t)((t / TC_RING) & 1));
        tc_fence_after();
        uint32_t r[8];
        tc_ld8(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((t % TC_RING) * TC_N), r);
        tc_wait_ld();
        float v[8];
#pragma unroll
        for (int co = 0; co < 8; co++) v[co] = __uint_as_float(r[co]) + biasr[co];
        const int q = 128 * t + (tid & 127);
        if (q < TC_R * P) stage[q] = pack_relu_bf8(v);
        tc_fence_before();                               // this warp's TMEM reads of the tile are complete
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&ebars[t % TC_RING]);
    }
    tc_fence_before();
    __syncthreads();                                     // stage rows are visible to every thread

    // ---- 2x2 max-pool of the staged strip -> HBM
    {
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
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

int pol_tc_conv_pool(const ofb_policy *p, int layer, const __nv_bfloat16 *in, __nv_bfloat16 *out, int hin, int n_items,
                     long long out_item_stride, cudaStream_t st) {
    const TcPlan pl = tc_plan(hin);
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_tc_conv_pool, (int)pl.total));
    if (n_items == 0) return OFB_OK;
    TcArgs a = {};
    a.in = in; a.wt = p->w.cw[layer]; a.bias = p->w.cb[layer]; a.out = out;
    a.out_item_stride = out_item_stride; a.H = hin;
    k_tc_conv_pool<<<dim3(hin / TC_R, n_items), TC_NT, pl.total, st>>>(a);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// Internal definitions of the policy forward (see include/ofb_policy.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ofb.h"
#include "../../include/ofb_policy.h"

#define POL_W 400
#define POL_WORDS 5000           // W*H/32 words per bit map
#define POL_TAPS 10              // 9 taps + 1 zero tap (K = 10 * 8 = 80 = 5 UMMA K-steps)
#define POL_FLAT 5000            // 25*25*8
#define POL_FLAT_PITCH 5120      // padded K of dense1's flat part (zeros beyond 5000)

// Device-side weights, BN already folded (see fold_weights in ofb_policy.cu).
struct PolicyDev {
    // trunk
    float *c1_w;              // [9][2][8]  (tap, cin, cout)  fp32
    float *c1_b;              // [8]
    float *c1_lut;            // [2][512][8] sums of c1_w over the set taps of a 9-bit stencil pattern
    __nv_bfloat16 *cw[3];     // conv2..4: [10][16][8] (tap, n = cout padded to 16, cin)  bf16
    float *cb[3];             // [16]
    // dense1
    float *d1_wv;             // [8][100]     vector slice, fp32 (values up to 400 are not bf16-exact)
    __nv_bfloat16 *d1_wf;     // [5000][100]  flat slice, (k, n)  -- CUDA-core engine
    __nv_bfloat16 *d1_wt;     // [128][5120]  flat slice transposed + zero padded (n, k) -- tensor engine
    float *d1_b;              // [100]
    // heads
    float *d2_w, *d2_b;       // [100][50], [50]
    float *o1_w, *o1_b;       // [50][2], [2]
    float *ud_w, *ud_b;       // [100][625], [625]
    float *u1_w, *u1_b;       // upconv1 [9][1][2], [2]
    float *u2_w, *u2_b;       // upconv2 [9][2][4], [4]
    float *u1_pw, *u2_pw;     // the same, bilinear x2 folded into 4 output phases: [9][1][8], [9][2][16] (n = phase * cout + co)
    float *u2_tf;             // upconv2 (folded) as a tcgen05 kind::tf32 B operand: [5 k-steps][2 chunks][32 n][4 lanes]
    float *u1_cw, *u2_cw;     // ... per border class of the low-res pixel (first / middle / last row x column): [9 cls][9][1][8], [9 cls][9][2][16]
    // upconv3 / upconv4: bilinear x2 folded into 4 output phases on the low-res grid
    float *u3_w, *u3_b;       // [9][4][8], [8]   un-phased fp32 (ring pixels)
    __nv_bfloat16 *u3_pw;     // [10][32][8]      (tap, n = phase*8 + cout, cin padded to 8)
    float *u3_pb;             // [32]
    float *u4_w, *u4_b;       // [9][8][1], [1]
    __nv_bfloat16 *u4_pw;     // [10][16][8]      (tap, n = phase (4 used), cin)
    float *u4_pb;             // [16]
    float *sp_bg1;            // [8]     pool1 of an empty arena = bf16(relu(conv1 bias)), as fp32          (sparse trunk)
    __nv_bfloat16 *sp_bg2;    // [9][8]  pool2 of an empty arena per border class (Y in {0, mid, 99}) x (X in {0, mid, 99})
    __nv_bfloat16 *c2_tz;     // block-Toeplitz B operand of conv2: [3 u][5 k-steps][2 chunks][64 n = xo*8 + cout][8 cin]
    __nv_bfloat16 *u3_tz;     // block-Toeplitz B operand of upconv3: [3 u][3 k-steps][2 chunks][128 n = xo*32 + phase*8 + cout][8 cin]
    __nv_bfloat16 *u4_tz;     // block-Toeplitz B operand of upconv4: [3 u][5 k-steps][2 chunks][32 n = xo*4 + phase][8 cin]
    int bil_legacy;           // 1 = TF1.x legacy bilinear in the scalar border code (the folded operands carry it themselves)
    __nv_bfloat16 *c2_st, *c3_st, *c4_st;   // conv2..4 as B operands of the sparse tensor trunk: [8 k-steps][2 chunks][32 n = px*8 + cout][8 cin]
    __nv_bfloat16 *sp_bg3, *sp_bg4;         // [9][8] pool3 / pool4 of an empty arena per border class
    uint8_t *tail_blob;       // operands + constants of the fused tail kernel (ofb_policy_tail.cuh: TL_WBYTES)
    uint8_t *tail_blob2;      // the same for CTA pairs: [2 ranks][TL_WBYTES_H], each rank's half of every operand's columns
};

struct PolicyWork {
    __nv_bfloat16 *pool1;     // [Ca][200*200*8]
    __nv_bfloat16 *pool2;     // [Ca][100*100*8]
    __nv_bfloat16 *pool3;     // [Ca][50*50*8]
    __nv_bfloat16 *flat;      // [Ca][5120]
    float *hflat;             // [Ca][100]   dense1 flat-part pre-activation
    __nv_bfloat16 *up2;       // [Cs][POL_UP2_ITEM]: tensor engine = plane layout [4][100][26][8], CUDA-core engine = NHWC [100][100][8]
    __nv_bfloat16 *up3;       // [Cs][POL_UP3_ITEM]: tensor engine = plane layout [8][200][26][8], CUDA-core engine = NHWC [200][200][8]
    uint4 *st_scratch;        // sparse tensor trunk: per CTA a dense pool2 (100 x 100) + pool3 (50 x 50) image, [ST_MAX_CTAS][12500]
    float *amax_val;          // [Cs][AMAX_PARTS]
    int *amax_idx;            // [Cs][AMAX_PARTS]
};
#define AMAX_PARTS 160
#define ST_MAX_CTAS 320                // upper bound of the sparse tensor trunk's grid (2 CTAs per SM)
#define ST_SCRATCH_CELLS (100 * 100 + 50 * 50)
#define POL_UP2_ITEM (4 * 100 * 26 * 8)   // elements of one upconv2 output in plane layout (>= 100*100*8)
#define POL_UP3_ITEM (8 * 200 * 26 * 8)   // elements of one upconv3 output in plane layout (>= 200*200*8)

struct ofb_policy {
    int device;
    int max_ships;
    int engine;
    float u4_bias;            // upconv4's (single) bias, host copy
    PolicyDev w;
    PolicyWork ws;
    int dense_trunk;          // 1 = the dense tcgen05 trunk12 (k_tz_trunk12); 2 = the CUDA-core sparse one (k_sp_trunk12); 0 = sparse + tcgen05 (k_st_trunk12)
    int unfused_trunk;        // 1 = trunk12, conv3, conv4 as three kernels through HBM; 0 = the whole trunk in k_st_trunk
    int unfused_tail;         // 1 = upconv3 / upconv4 as two kernels through HBM (k_tz_up3, k_tz_up4); 0 = the fused tail (k_tz_tail)
    int tail_pair;            // 1 = the fused tail runs as CTA pairs (tcgen05 cta_group::2, M = 256); 0 = one CTA per SM on its own
    int taps;                 // 1 = the fused tail also writes upconv3's output (validation taps)
    int bilinear_legacy;      // 0 = TF2 half-pixel bilinear x2 (default), 1 = TF1.x legacy (asymmetric) UpSampling2D
    int profiling;            // when set, forward brackets every kernel with CUDA events
    void *prof;               // std::vector<ProfEvent>*
    void *arena_blob;         // single allocation holding all weights
    size_t arena_bytes;
    void *work_blob;          // single allocation holding the workspace
    void *wide_blob;          // second allocation (on first use): pool1 / pool2 / pool3 / up3 of the alternative kernels and the taps
};

void ofb_set_error(const char *fmt, ...);

// tensor-core (tcgen05) kernels, ofb_policy_tc.cu
int pol_tc_conv_pool(const ofb_policy *p, int layer, const __nv_bfloat16 *in, __nv_bfloat16 *out, int hin, int n_items,
                     long long out_item_stride, cudaStream_t st);
// block-Toeplitz kernels, ofb_policy_tz.cu
int pol_tz_up4(const ofb_policy *p, const __nv_bfloat16 *in, float *ptr_out, float *amax_val, int *amax_idx, int n_items,
               cudaStream_t st);
int pol_tz_up4_parts();
int pol_tc_dense1(const ofb_policy *p, const __nv_bfloat16 *flat, float *hflat, int n_items, cudaStream_t st);
int pol_tz_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st);
// sparse trunk12 on CUDA cores, ofb_policy_sp.cu
int pol_sp_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st);
// sparse trunk12 with conv2 on tensor cores (one MMA row per dirty cell), ofb_policy_st.cu
int pol_st_trunk12(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *out, int n_items, cudaStream_t st);
// ... the whole trunk (conv1 .. conv4 + pools) in the same kernel: bit maps -> flat [item][POL_FLAT_PITCH]; tap2 / tap3 (optional) receive pool2 / pool3
int pol_st_trunk(const ofb_policy *p, const uint32_t *maps, __nv_bfloat16 *flat, __nv_bfloat16 *tap2, __nv_bfloat16 *tap3, int n_items,
                 cudaStream_t st);
int pol_tz_up3(const ofb_policy *p, const __nv_bfloat16 *in, __nv_bfloat16 *out, int n_items, cudaStream_t st);
// fused upconv3 -> upconv4 -> argmax, ofb_policy_tail.cu (input: upconv2's output in the pairs layout)
int pol_tz_tail(const ofb_policy *p, const __nv_bfloat16 *up2_pairs, float *ptr_out, int32_t *xy, __nv_bfloat16 *up3_dbg, int n_items,
                cudaStream_t st);
// element offset of pixel (y, x) of a 100 x 100 x 8 image in plane layout (4 planes of x mod 4)
__host__ __device__ __forceinline__ int pol_plane100_off(int y, int x) { return (((x & 3) * 100 + y) * 26 + (x >> 2) + 1) * 8; }
// element offset of pixel (Y, X) of a 200 x 200 x 8 image in plane layout (ofb_policy_tz.cu)
__host__ __device__ __forceinline__ int pol_plane200_off(int Y, int X) { return (((X & 7) * 200 + Y) * 26 + (X >> 3) + 1) * 8; }

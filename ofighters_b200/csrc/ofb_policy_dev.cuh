// Device helpers shared by the CUDA-core and the tensor-core policy kernels, so that the pieces
// that are NOT contractions (conv1 on binary input, bilinear taps, border ring of the phase-folded
// up-convolutions, argmax ordering) are one piece of code in both engines.
#pragma once
#include "ofb_policy.cuh"

__device__ __forceinline__ float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&t);
}
// max(x, 0) and round-to-nearest bf16 of two floats in one instruction (a -> low half, b -> high half)
__device__ __forceinline__ uint32_t pack_relu_bf2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint4 pack_relu_bf8(const float *v) {
    return make_uint4(pack_relu_bf2(v[0], v[1]), pack_relu_bf2(v[2], v[3]), pack_relu_bf2(v[4], v[5]), pack_relu_bf2(v[6], v[7]));
}
__device__ __forceinline__ uint4 pack_bf8(const float *v) {
    return make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}
__device__ __forceinline__ void unpack_bf8(const uint4 &q, float *v) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// 4 consecutive bits of a bit map starting at bit index b (b may be -1: bit -1 reads as 0).
__device__ __forceinline__ uint32_t bits4(const uint32_t *__restrict__ m, int b) {
    const int sh = b < 0 ? 1 : 0;
    b = max(b, 0);
    const int w = b >> 5;
    const uint32_t lo = m[w], hi = m[min(w + 1, POL_WORDS - 1)];
    return ((__funnelshift_r(lo, hi, b & 31) << sh) & 0xFu);
}

// Bits of the 4x4 input patch (rows 2py-1..2py+2, cols 2px-1..2px+2) that one pooled conv1 output
// pixel depends on; out-of-map bits are 0.  Returns rows packed 4 bits each (row i in bits 4i..4i+3).
__device__ __forceinline__ uint32_t conv1_patch(const uint32_t *__restrict__ m, int py, int px) {
    uint32_t colmask = 0xFu;
    if (px == 0) colmask &= 0xEu;
    if (px == POL_W / 2 - 1) colmask &= 0x7u;
    uint32_t p = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int r = 2 * py - 1 + i;
        if (r >= 0 && r < POL_W) p |= (bits4(m, r * POL_W + 2 * px - 1) & colmask) << (4 * i);
    }
    return p;
}

// conv1 (3x3 'same', 2 -> 8, BN folded) + ReLU + 2x2 max-pool for one pooled pixel whose 4x4 patch is
// (ps, pl).  The input is binary, so the 3x3 x 1-channel stencil is a 9-bit pattern: lut[ch][pattern][co]
// holds sum_{set taps} w[tap][ch][co] (fp32, summed in tap order on the host) and a conv output is
// bias + lut[ship][pattern_s] + lut[laser][pattern_l].
__device__ __forceinline__ uint32_t conv1_pattern(uint32_t p, int i, int j) {
    return ((p >> (4 * i + j)) & 7u) | (((p >> (4 * (i + 1) + j)) & 7u) << 3) | (((p >> (4 * (i + 2) + j)) & 7u) << 6);
}
__device__ __forceinline__ void conv1_pool_pixel(uint32_t ps, uint32_t pl, const float *__restrict__ lut,
                                                 const float *__restrict__ b, float *out) {
#pragma unroll
    for (int co = 0; co < 8; co++) out[co] = 0.0f;             // ReLU output >= 0
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            // (entry 0 = no tap set = 0.0f: most dirty pixels see only one of the two maps, so half of the lookups are skipped)
            const uint32_t qs = conv1_pattern(ps, i, j), ql = conv1_pattern(pl, i, j);
            const float4 *ls = reinterpret_cast<const float4 *>(lut + (size_t)qs * 8);
            const float4 *ll = reinterpret_cast<const float4 *>(lut + (size_t)(512 + ql) * 8);
            const float4 zero = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const float4 s0 = qs ? __ldg(ls) : zero, s1 = qs ? __ldg(ls + 1) : zero, l0 = ql ? __ldg(ll) : zero, l1 = ql ? __ldg(ll + 1) : zero;
            const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
            const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
            for (int co = 0; co < 8; co++) out[co] = fmaxf(out[co], (b[co] + sv[co]) + lv[co]);
        }
}

// TF2 bilinear x2 (half-pixel centres, edge clamp): upsampled coordinate Y of a length-n axis reads
// low-res lo/hi with weights wlo/whi.
__device__ __forceinline__ void bil_tap(int Y, int n, int &lo, int &hi, float &wlo, float &whi) {
    const int i = Y >> 1;
    if (Y & 1) { lo = i; hi = min(i + 1, n - 1); wlo = 0.75f; whi = 0.25f; }
    else { lo = max(i - 1, 0); hi = i; wlo = 0.25f; whi = 0.75f; }
}

// Border ring of [bilinear x2 -> conv3x3 'same'] in the phase-folded form.  The folded weights assume
// the upsampled map continues past its border (replicate-extended low-res input), whereas the
// convolution zero-pads it: for the output pixels on the 1-pixel ring, the contribution of the taps
// that fall outside [0, 2n) must be taken back out:
//     out_true = out_folded - sum_{(dy,dx) out of range} w[dy][dx] . U~[Y+dy-1][X+dx-1]
// with U~ the same half-pixel interpolation evaluated on the replicate-extended input.
// `L(y, x)` returns the 8 bf16 channels of low-res pixel (y, x) for y, x in [-1, n] (replicated).
// legacy != 0: TF1.x UpSampling2D (U[2i] = L[i], U[2i+1] = (L[i] + L[i+1]) / 2) instead of TF2's half-pixel centres
__device__ __forceinline__ void bil_tap_ext(int Y, int &lo, int &hi, float &wlo, float &whi, int legacy = 0) {
    const int i = Y >> 1;                                   // arithmetic shift: -1 -> -1
    if (legacy) {
        lo = i; hi = i + (Y & 1); wlo = (Y & 1) ? 0.5f : 1.0f; whi = (Y & 1) ? 0.5f : 0.0f;
        return;
    }
    if (Y & 1) { lo = i; hi = i + 1; wlo = 0.75f; whi = 0.25f; }
    else { lo = i - 1; hi = i; wlo = 0.25f; whi = 0.75f; }
}

// contribution of ONE out-of-range tap (dy, dx) of output pixel (Y, X): acc[co] += w[dy][dx][:, co] . U~[Y+dy-1][X+dx-1]
template <int CIN, int COUT, class Acc>
__device__ __forceinline__ void up_ring_tap(const Acc &L, int Y, int X, int dy, int dx, const float *__restrict__ w, float *acc,
                                            int legacy = 0) {
    int ylo, yhi, xlo, xhi; float wyl, wyh, wxl, wxh;
    bil_tap_ext(Y + dy - 1, ylo, yhi, wyl, wyh, legacy);
    bil_tap_ext(X + dx - 1, xlo, xhi, wxl, wxh, legacy);
    float a[8], c[8], d[8], e[8];
    unpack_bf8(L(ylo, xlo), a);
    unpack_bf8(L(ylo, xhi), c);
    unpack_bf8(L(yhi, xlo), d);
    unpack_bf8(L(yhi, xhi), e);
    const int tap = dy * 3 + dx;
#pragma unroll
    for (int ci = 0; ci < CIN; ci++) {
        const float u = wyl * (wxl * a[ci] + wxh * c[ci]) + wyh * (wxl * d[ci] + wxh * e[ci]);
#pragma unroll
        for (int co = 0; co < COUT; co++) acc[co] += u * w[(tap * CIN + ci) * COUT + co];
    }
}

template <int CIN, int COUT, class Acc>
__device__ __forceinline__ void up_ring_correct(const Acc &L, int n, int Y, int X, const float *__restrict__ w, float *v, int legacy = 0) {
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; co++) acc[co] = 0.f;
    for (int dy = 0; dy < 3; dy++) {
        const int YY = Y + dy - 1;
        for (int dx = 0; dx < 3; dx++) {
            const int XX = X + dx - 1;
            if (YY >= 0 && YY < 2 * n && XX >= 0 && XX < 2 * n) continue;
            up_ring_tap<CIN, COUT>(L, Y, X, dy, dx, w, acc, legacy);
        }
    }
#pragma unroll
    for (int co = 0; co < COUT; co++) v[co] -= acc[co];
}

// accessor over a dense [n][n][8] bf16 image in global memory, replicate-clamped
struct GlobalImage {
    const __nv_bfloat16 *p;
    int n;
    __device__ __forceinline__ uint4 operator()(int y, int x) const {
        y = min(max(y, 0), n - 1);
        x = min(max(x, 0), n - 1);
        return *reinterpret_cast<const uint4 *>(p + ((size_t)y * n + x) * 8);
    }
};

// argmax ordering of np.argmax on the flat C-order map: larger value wins, ties -> lower index.
__device__ __forceinline__ bool amax_better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

__device__ __forceinline__ void amax_warp(float &v, int &i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        if (amax_better(ov, oi, v, i)) { v = ov; i = oi; }
    }
}

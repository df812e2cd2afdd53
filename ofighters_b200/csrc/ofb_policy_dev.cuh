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

// conv1 (3x3 'same', 2 -> 8, BN folded) + ReLU + 2x2 max-pool for one pooled pixel whose patch is
// (ps, pl).  w = [9][2][8] fp32, b = [8].  Accumulation order: bias, then taps 0..8, ship channel
// before laser channel.
__device__ __forceinline__ void conv1_pool_pixel(uint32_t ps, uint32_t pl, const float *__restrict__ w,
                                                 const float *__restrict__ b, float *out) {
#pragma unroll
    for (int co = 0; co < 8; co++) out[co] = 0.0f;             // ReLU output >= 0
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            float acc[8];
#pragma unroll
            for (int co = 0; co < 8; co++) acc[co] = b[co];
#pragma unroll
            for (int dy = 0; dy < 3; dy++)
#pragma unroll
                for (int dx = 0; dx < 3; dx++) {
                    const int bit = 4 * (i + dy) + (j + dx), tap = dy * 3 + dx;
                    if ((ps >> bit) & 1u) {
#pragma unroll
                        for (int co = 0; co < 8; co++) acc[co] += w[(tap * 2 + 0) * 8 + co];
                    }
                    if ((pl >> bit) & 1u) {
#pragma unroll
                        for (int co = 0; co < 8; co++) acc[co] += w[(tap * 2 + 1) * 8 + co];
                    }
                }
#pragma unroll
            for (int co = 0; co < 8; co++) out[co] = fmaxf(out[co], acc[co]);
        }
}

// TF2 bilinear x2 (half-pixel centres, edge clamp): upsampled coordinate Y of a length-n axis reads
// low-res lo/hi with weights wlo/whi.
__device__ __forceinline__ void bil_tap(int Y, int n, int &lo, int &hi, float &wlo, float &whi) {
    const int i = Y >> 1;
    if (Y & 1) { lo = i; hi = min(i + 1, n - 1); wlo = 0.75f; whi = 0.25f; }
    else { lo = max(i - 1, 0); hi = i; wlo = 0.25f; whi = 0.75f; }
}

// One output pixel (Y, X) of [bilinear x2 -> conv3x3 'same' (zero pad)] straight from the definition,
// input L = [n][n][8] bf16 (CIN real channels), weights w = [9][CIN][COUT] fp32, bias [COUT].
// Used for the 1-pixel border ring, where the zero padding of the upsampled map breaks the
// phase-folded form.
template <int CIN, int COUT>
__device__ __forceinline__ void up_ring_pixel(const __nv_bfloat16 *__restrict__ L, int n, int Y, int X,
                                              const float *__restrict__ w, const float *__restrict__ b, float *out) {
#pragma unroll
    for (int co = 0; co < COUT; co++) out[co] = b[co];
    for (int dy = 0; dy < 3; dy++) {
        const int YY = Y + dy - 1;
        if (YY < 0 || YY >= 2 * n) continue;
        int ylo, yhi; float wyl, wyh;
        bil_tap(YY, n, ylo, yhi, wyl, wyh);
        for (int dx = 0; dx < 3; dx++) {
            const int XX = X + dx - 1;
            if (XX < 0 || XX >= 2 * n) continue;
            int xlo, xhi; float wxl, wxh;
            bil_tap(XX, n, xlo, xhi, wxl, wxh);
            float a[8], c[8], d[8], e[8];
            unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)ylo * n + xlo) * 8), a);
            unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)ylo * n + xhi) * 8), c);
            unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)yhi * n + xlo) * 8), d);
            unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)yhi * n + xhi) * 8), e);
            const int tap = dy * 3 + dx;
#pragma unroll
            for (int ci = 0; ci < CIN; ci++) {
                const float u = wyl * (wxl * a[ci] + wxh * c[ci]) + wyh * (wxl * d[ci] + wxh * e[ci]);
#pragma unroll
                for (int co = 0; co < COUT; co++) out[co] += u * w[(tap * CIN + ci) * COUT + co];
            }
        }
    }
}

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

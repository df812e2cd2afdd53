// Disk rasterisation shared by K2 (ofb_raster.cu) and the fused frame kernel (ofb_step.cu).  Both TUs are compiled
// with -fmad=false; the explicit _rn intrinsics make the fp64 sequence independent of that flag anyway.
//
// skimage.draw.disk as called by Circle.binary_draw (lib/form.py:222-228):
//     ul = ceil(c - R), lr = floor(c + R) clipped to the map, sc = c - ul,
//     pixel (i, j) of the box set iff ((i - sc_r)/R)^2 + ((j - sc_c)/R)^2 < 1   (fp64, strict).
// R is 8 or 2, so the divisions are exact scalings and the test is dr*dr + dc*dc < R*R with
// dr = fl(i - fl(c - ul)).  Ship centres are integers -> the same test in int32.
#pragma once
#include "ofb_common.cuh"

__device__ __forceinline__ void set_bit_range(uint32_t *bits, unsigned b0, unsigned b1) {
    const unsigned w0 = b0 >> 5, w1 = b1 >> 5;
    const uint32_t lo = ~0u << (b0 & 31u), hi = (2u << (b1 & 31u)) - 1u;
    if (w0 == w1) atomicOr(&bits[w0], lo & hi);
    else {
        atomicOr(&bits[w0], lo);
        for (unsigned w = w0 + 1; w < w1; w++) atomicOr(&bits[w], ~0u);
        atomicOr(&bits[w1], hi);
    }
}

// half-width of the radius-8 disk's row |dr| = 0..7 (largest dc with dr^2 + dc^2 < 64): 7,7,7,7,6,6,5,3
__device__ __forceinline__ int ship_half_width(int dr) { return (int)((0x35667777u >> (4 * abs(dr))) & 15u); }

// one row (dr in [-(R-1), R-1]) of the radius-8 disk of a ship; sxy = x | y << 16 | alive << 31.  Returns the index of the
// first word written (the row spans at most that word and the next), -1 if nothing was drawn.
__device__ __forceinline__ int raster_ship_row(uint32_t *bits, int W, int H, unsigned sxy, int dr) {
    if (!(sxy >> 31)) return -1;
    const int cx = (int)(sxy & 0xffffu), y = (int)((sxy >> 16) & 0x7fffu) + dr;
    if (y < 0 || y >= H) return -1;
    const int hw = ship_half_width(dr);
    const int c0 = max(0, cx - hw), c1 = min(W - 1, cx + hw);
    if (c0 > c1) return -1;
    set_bit_range(bits, (unsigned)(y * W + c0), (unsigned)(y * W + c1));
    return (y * W + c0) >> 5;
}

// the radius-2 disk of a laser centred on (cx, cy): <= 5x5 candidate pixels in fp64
__device__ __forceinline__ void raster_laser(uint32_t *lbits, int W, int H, double cx, double cy) {
    const double R = (double)OFB_R_LASER;
    long long ulr = (long long)ceil(__dsub_rn(cy, R)), ulc = (long long)ceil(__dsub_rn(cx, R));
    long long lrr = (long long)floor(__dadd_rn(cy, R)), lrc = (long long)floor(__dadd_rn(cx, R));
    ulr = ulr < 0 ? 0 : ulr;
    ulc = ulc < 0 ? 0 : ulc;
    lrr = lrr > H - 1 ? H - 1 : lrr;
    lrc = lrc > W - 1 ? W - 1 : lrc;
    const double scr = __dsub_rn(cy, (double)ulr), scc = __dsub_rn(cx, (double)ulc);
    for (long long i = 0; i <= lrr - ulr; i++) {
        const double dr = __dsub_rn((double)i, scr);
        const double dr2 = __dmul_rn(dr, dr);
        for (long long j = 0; j <= lrc - ulc; j++) {
            const double dc = __dsub_rn((double)j, scc);
            if (__dadd_rn(dr2, __dmul_rn(dc, dc)) < R * R) {
                const unsigned b = (unsigned)((ulr + i) * W + (ulc + j));
                atomicOr(&lbits[b >> 5], 1u << (b & 31u));
            }
        }
    }
}

// The same disk, one row i of the <= 5-row box per call (the fused frame kernel spreads a laser over 5 threads).  A centre
// further than R + 1 from the map touches no pixel (the clipped box is empty), which also keeps everything below inside
// int32; otherwise the arithmetic is the sequence above, so the bits are identical.  The 5 pixel tests are independent
// (fixed trip count, predicated) so that their fp64 latencies overlap.  Returns the index of the first word written (the
// row spans at most that word and the next), -1 if nothing was drawn.
__device__ __forceinline__ int raster_laser_row(uint32_t *lbits, int W, int H, double cx, double cy, int i) {
    const double R = (double)OFB_R_LASER;
    if (!(cx > -(R + 1.0) && cx < (double)W + R + 1.0 && cy > -(R + 1.0) && cy < (double)H + R + 1.0)) return -1;
    const int ulr = max(0, __double2int_ru(__dsub_rn(cy, R))), ulc = max(0, __double2int_ru(__dsub_rn(cx, R)));
    const int lrr = min(H - 1, __double2int_rd(__dadd_rn(cy, R))), lrc = min(W - 1, __double2int_rd(__dadd_rn(cx, R)));
    if (i > lrr - ulr) return -1;
    const double dr = __dsub_rn((double)i, __dsub_rn(cy, (double)ulr));
    const double dr2 = __dmul_rn(dr, dr), scc = __dsub_rn(cx, (double)ulc);
    const int nj = lrc - ulc;
    uint32_t m = 0;                                      // the row's <= 5 pixels start at bit b0 and span at most two words
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const double dc = __dsub_rn((double)j, scc);
        if (j <= nj && __dadd_rn(dr2, __dmul_rn(dc, dc)) < R * R) m |= 1u << j;
    }
    if (!m) return -1;
    const unsigned b0 = (unsigned)((ulr + i) * W + ulc), sh = b0 & 31u;
    atomicOr(&lbits[b0 >> 5], m << sh);
    if (sh > 27u && (m >> (32u - sh))) atomicOr(&lbits[(b0 >> 5) + 1], m >> (32u - sh));
    return (int)(b0 >> 5);
}

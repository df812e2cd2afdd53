// Policy forward of the bi-head pointer model (agents/qlearnIA_V2.py:123-235): handle, weight
// folding, orchestration, and the CUDA-core engine.  The tensor-core (tcgen05) kernels of the
// product path live in ofb_policy_tc.cu; both engines share ofb_policy_dev.cuh.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "ofb_common.cuh"
#include "ofb_policy_dev.cuh"
#include "ofb_policy_tail.cuh"
#include "ofb_tc_ptx.cuh"

struct ProfEvent { cudaEvent_t a, b; int layer; };
enum { L_TRUNK12 = 0, L_CONV3, L_CONV4, L_DENSE1, L_HEADS, L_UP3, L_UP4, L_ARGMAX, L_TAIL, L_COUNT };

// ------------------------------------------------------------------------------------------------
// host-side folding (BN into conv, bilinear x2 into 4 output phases) and upload
// ------------------------------------------------------------------------------------------------
static void fold_conv(const ofb_conv_weights &c, int cin, int cout, std::vector<float> &w, std::vector<float> &b) {
    w.assign((size_t)9 * cin * cout, 0.f);
    b.assign(cout, 0.f);
    for (int co = 0; co < cout; co++) {
        float s = 1.f, sh = c.bias ? c.bias[co] : 0.f;
        if (c.gamma) {                                        // BatchNormalization, eps = 1e-3 (Keras default)
            s = c.gamma[co] / std::sqrt(c.var[co] + 1e-3f);
            sh = (sh - c.mean[co]) * s + c.beta[co];
        }
        b[co] = sh;
        for (int t = 0; t < 9; t++)
            for (int ci = 0; ci < cin; ci++) w[((size_t)t * cin + ci) * cout + co] = c.kernel[((size_t)t * cin + ci) * cout + co] * s;
    }
}

// [9][cin][cout] fp32 -> UMMA B-operand image [10 taps][npad][8 cin] bf16 (K-major rows of 16 B)
static std::vector<__nv_bfloat16> pack_taps(const std::vector<float> &w, int cin, int cout, int npad) {
    std::vector<__nv_bfloat16> o((size_t)POL_TAPS * npad * 8, __float2bfloat16(0.f));
    for (int t = 0; t < 9; t++)
        for (int ci = 0; ci < cin; ci++)
            for (int co = 0; co < cout; co++) o[((size_t)t * npad + co) * 8 + ci] = __float2bfloat16(w[((size_t)t * cin + ci) * cout + co]);
    return o;
}

// c[a][d][u] = coefficient of L[i+u-1] in U[2i + a + d - 1] for a bilinear x2 upsampling U of L: a = output phase, d = conv
// tap, on a replicate-extended L.  TF2 / Keras >= 2.2.? half-pixel centres: U[2i] = .25 L[i-1] + .75 L[i], U[2i+1] = .75 L[i] +
// .25 L[i+1];  TF1.x legacy (asymmetric, align_corners=False): U[2i] = L[i], U[2i+1] = .5 (L[i] + L[i+1]).
struct PhaseTab { float c[2][3][3]; };
static PhaseTab phase_tab(int legacy) {
    static const PhaseTab tf2 = {{{{0.75f, 0.25f, 0.f}, {0.25f, 0.75f, 0.f}, {0.f, 0.75f, 0.25f}},
                                  {{0.25f, 0.75f, 0.f}, {0.f, 0.75f, 0.25f}, {0.f, 0.25f, 0.75f}}}};
    static const PhaseTab tf1 = {{{{0.5f, 0.5f, 0.f}, {0.f, 1.f, 0.f}, {0.f, 0.5f, 0.5f}},
                                  {{0.f, 1.f, 0.f}, {0.f, 0.5f, 0.5f}, {0.f, 0.f, 1.f}}}};
    return legacy ? tf1 : tf2;
}
// The same for the FIRST row / column of the image (phase 0): tap 0 reads the convolution's zero padding, and the edge
// clamp of the upsampling folds L[-1] onto L[0].
static PhaseTab low_border(PhaseTab t) {
    for (int d = 0; d < 3; d++) {
        if (d == 0) { t.c[0][0][0] = t.c[0][0][1] = t.c[0][0][2] = 0.f; continue; }
        t.c[0][d][1] += t.c[0][d][0];
        t.c[0][d][0] = 0.f;
    }
    return t;
}
// ... and for the LAST row / column (phase 1): tap 2 reads the padding, L[n] folds onto L[n-1].
static PhaseTab high_border(PhaseTab t) {
    for (int d = 0; d < 3; d++) {
        if (d == 2) { t.c[1][2][0] = t.c[1][2][1] = t.c[1][2][2] = 0.f; continue; }
        t.c[1][d][1] += t.c[1][d][2];
        t.c[1][d][2] = 0.f;
    }
    return t;
}

// [9][cin][cout] -> bilinear-x2 phase-folded fp32 [9 (u, v)][cin][4 * cout], column = (a*2+b)*cout + co; ty / tx = the
// coefficient tables of the vertical / horizontal direction
static std::vector<float> fold_phase(const std::vector<float> &w, int cin, int cout, const PhaseTab &ty, const PhaseTab &tx) {
    std::vector<float> o((size_t)9 * cin * 4 * cout, 0.f);
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++)
            for (int u = 0; u < 3; u++)
                for (int v = 0; v < 3; v++)
                    for (int ci = 0; ci < cin; ci++)
                        for (int co = 0; co < cout; co++) {
                            double s = 0.0;
                            for (int dy = 0; dy < 3; dy++)
                                for (int dx = 0; dx < 3; dx++)
                                    s += (double)w[((size_t)(dy * 3 + dx) * cin + ci) * cout + co] * ty.c[a][dy][u] * tx.c[b][dx][v];
                            o[((size_t)(u * 3 + v) * cin + ci) * 4 * cout + (a * 2 + b) * cout + co] = (float)s;
                        }
    return o;
}

// phase-folded fp32 [9 (u, v)][cin][coutp] -> block-Toeplitz B operand [3 u][nks][2 chunks][B * coutp n][8 cin] bf16 for
// blocks of B pixels: n = xo * coutp + c, chunk of pixel offset p (input x = B xb + p - 1) carries tap v = p - xo.
// K-step pairs: (1, 0), (2, 3), ..., (B - 2, B - 1), (B + 1, B)  -- see ofb_policy_tz.cu
static std::vector<__nv_bfloat16> pack_toeplitz(const std::vector<float> &pf, int cin, int coutp, int B) {
    const int nks = (B + 2) / 2, N = B * coutp;
    std::vector<__nv_bfloat16> o((size_t)3 * nks * 2 * N * 8, __float2bfloat16(0.f));
    for (int u = 0; u < 3; u++)
        for (int ks = 0; ks < nks; ks++)
            for (int c = 0; c < 2; c++) {
                int pix;
                if (ks == 0) pix = c == 0 ? 1 : 0;
                else if (ks == nks - 1) pix = c == 0 ? B + 1 : B;
                else pix = 2 * ks + c;
                for (int xo = 0; xo < B; xo++) {
                    const int v = pix - xo;
                    if (v < 0 || v > 2) continue;
                    for (int cc = 0; cc < coutp; cc++)
                        for (int ci = 0; ci < cin; ci++)
                            o[((((size_t)u * nks + ks) * 2 + c) * N + xo * coutp + cc) * 8 + ci] =
                                __float2bfloat16(pf[((size_t)(u * 3 + v) * cin + ci) * coutp + cc]);
                }
            }
    return o;
}

// Operands of the fused tail kernel (layouts: ofb_policy_tail.cuh / .cu).  w3 = upconv3 [9][4][8] and w4 = upconv4 [9][8][1],
// BN folded; b3 = upconv3's bias.
static std::vector<uint8_t> pack_tail(const std::vector<float> &w3, const std::vector<float> &b3, const std::vector<float> &w4,
                                      int legacy) {
    std::vector<uint8_t> blob(TL_WBYTES, 0);
    auto put = [&](size_t byte_off, float v) { *reinterpret_cast<__nv_bfloat16 *>(blob.data() + byte_off) = __float2bfloat16(v); };
    const PhaseTab T = phase_tab(legacy), TLo = low_border(T), THi = high_border(T);
    // ---- upconv3: K lanes of chunk cc < 3 = pixel offsets p = 2 cc + (l >> 2) (input x = 4 xb + p - 1), channel l & 3; chunk 3 =
    //      spare lanes: l < 4 -> L[i][0] (M rows xb = 0), l >= 4 -> L[i][99] (xb = 24), weights = minus the out-of-range taps
    auto pack3 = [&](size_t base, const PhaseTab &ty, int u, int ks, int n_cols, bool only_a, int a_sel) {
        const std::vector<float> pf = fold_phase(w3, 4, 8, ty, T);
        for (int c = 0; c < 2; c++) {
            const int cc = 2 * ks + c;
            for (int n = 0; n < n_cols; n++) {
                int xo, a, b, co;
                if (only_a) { xo = n >> 4; a = a_sel; b = (n >> 3) & 1; co = n & 7; }
                else { xo = n >> 5; a = (n >> 4) & 1; b = (n >> 3) & 1; co = n & 7; }
                for (int l = 0; l < 8; l++) {
                    float val = 0.f;
                    const int ci = l & 3;
                    if (cc < 3) {
                        const int v = 2 * cc + (l >> 2) - xo;
                        if (v >= 0 && v <= 2) val = pf[((size_t)(u * 3 + v) * 4 + ci) * 32 + (a * 2 + b) * 8 + co];
                    } else if ((l < 4 && xo == 0 && b == 0) || (l >= 4 && xo == 3 && b == 1)) {
                        const int dx = l < 4 ? 0 : 2;
                        double sacc = 0.0;
                        for (int dy = 0; dy < 3; dy++) sacc += (double)w3[((size_t)(dy * 3 + dx) * 4 + ci) * 8 + co] * ty.c[a][dy][u];
                        val = (float)-sacc;
                    }
                    put(base + ((size_t)c * n_cols + n) * 16 + l * 2, val);
                }
            }
        }
    };
    for (int u = 0; u < 3; u++)
        for (int ks = 0; ks < 2; ks++) pack3(TL_OFF_B3 + (size_t)(u * 2 + ks) * 2 * 128 * 16, T, u, ks, 128, false, 0);
    for (int set = 0; set < 2; set++)                       // top: phase a = 0 of row 0, taps u = 1, 2; bottom: a = 1 of row 99, u = 0, 1
        for (int ui = 0; ui < 2; ui++)
            for (int ks = 0; ks < 2; ks++)
                pack3(TL_OFF_B3V + (size_t)(((set * 2 + ui) * 2) + ks) * 2 * 64 * 16, set == 0 ? TLo : THi, set == 0 ? ui + 1 : ui, ks, 64,
                      true, set);
    // ---- upconv4 (2-row form): output row r of the pair uses tap u = dy - r; pixel offset p of (ks, chunk): input X = 8 xb + p - 1
    const std::vector<float> pf4 = fold_phase(w4, 8, 1, T, T), pfL = fold_phase(w4, 8, 1, T, TLo), pfR = fold_phase(w4, 8, 1, T, THi);
    const std::vector<float> pfT = fold_phase(w4, 8, 1, TLo, T), pfB = fold_phase(w4, 8, 1, THi, T);
    auto pix_of = [](int ks, int c) { return ks == 0 ? (c == 0 ? 1 : 0) : (ks == 4 ? (c == 0 ? 9 : 8) : 2 * ks + c); };
    const size_t b4off[4] = {0, 480, 1280, 2080};
    for (int dy = 0; dy < 4; dy++) {
        const int nn = (dy == 0 || dy == 3) ? 48 : 80, n0 = dy == 3 ? 32 : 0;
        for (int ks = 0; ks < 5; ks++)
            for (int c = 0; c < 2; c++) {
                const int p = pix_of(ks, c);
                for (int j = 0; j < nn; j++) {
                    const int n = n0 + j;
                    for (int ci = 0; ci < 8; ci++) {
                        float val = 0.f;
                        int r = -1, cph = 0, xo = 0, d = 0;
                        const std::vector<float> *src = &pf4;
                        if (n < 4) { r = 0; cph = n >> 1; if (n & 1) { xo = 7; d = 1; src = &pfR; } else { xo = 0; d = 0; src = &pfL; } }
                        else if (n >= 8 && n < 72) { r = (n - 8) >> 5; const int rem = (n - 8) & 31; xo = rem >> 2; cph = (rem >> 1) & 1; d = rem & 1; }
                        else if (n >= 72 && n < 76) { r = 1; cph = (n - 72) >> 1; if (n & 1) { xo = 7; d = 1; src = &pfR; } else { xo = 0; d = 0; src = &pfL; } }
                        if (r >= 0) {
                            const int u = dy - r, v = p - xo;
                            if (u >= 0 && u <= 2 && v >= 0 && v <= 2) val = (*src)[((size_t)(u * 3 + v) * 8 + ci) * 4 + cph * 2 + d];
                        }
                        put(TL_OFF_B4 + (b4off[dy] + (size_t)(ks * 2 + c) * nn + j) * 16 + ci * 2, val);
                    }
                }
            }
    }
    for (int set = 0; set < 2; set++)                       // V = 0: row r = 0, phase c = 0, taps u = dy = 1, 2; V = 399: r = 1, c = 1, u = dy - 1 = 0, 1
        for (int dyi = 0; dyi < 2; dyi++)
            for (int ks = 0; ks < 5; ks++)
                for (int c = 0; c < 2; c++) {
                    const int p = pix_of(ks, c), u = set == 0 ? dyi + 1 : dyi;
                    const std::vector<float> &src = set == 0 ? pfT : pfB;
                    for (int n = 0; n < 16; n++)
                        for (int ci = 0; ci < 8; ci++) {
                            const int xo = n >> 1, d = n & 1, v = p - xo;
                            const float val = (v >= 0 && v <= 2) ? src[((size_t)(u * 3 + v) * 8 + ci) * 4 + set * 2 + d] : 0.f;
                            put(TL_OFF_B4V + ((size_t)((((set * 2 + dyi) * 5) + ks) * 2 + c) * 16 + n) * 16 + ci * 2, val);
                        }
                }
    // ---- floats: upconv3's bias, then the 4 corner pixels' weights [corner][uu][vv][ci] (bf16-rounded like the operands):
    //      the corner reads upconv3 pixels (Y0 + uu, X0 + vv), (Y0, X0) = (0 | 198, 0 | 198)
    // upconv3's bias rides the tensor pipe: one extra MMA per tile whose A operand is a column of ones (K lane 0).  (Adding it in
    // the drain instead was measured: 11.0 vs 10.3 ms per 65 536 ships -- the drain's instruction count is what the ring waits for.)
    for (int n = 0; n < 128; n++) put(TL_OFF_BIAS3 + (size_t)n * 16, b3[n & 7]);
    float *aux = reinterpret_cast<float *>(blob.data() + TL_OFF_AUX);
    for (int c = 0; c < 8; c++) aux[c] = b3[c];
    for (int cid = 0; cid < 4; cid++) {
        const bool bottom = cid >= 2, right = cid & 1;
        const std::vector<float> pfc = fold_phase(w4, 8, 1, bottom ? THi : TLo, right ? THi : TLo);
        for (int uu = 0; uu < 2; uu++)
            for (int vv = 0; vv < 2; vv++)
                for (int ci = 0; ci < 8; ci++) {
                    const int u = bottom ? uu : uu + 1, v = right ? vv : vv + 1;
                    const float val = pfc[((size_t)(u * 3 + v) * 8 + ci) * 4 + (bottom ? 2 : 0) + (right ? 1 : 0)];
                    aux[8 + ((cid * 2 + uu) * 2 + vv) * 8 + ci] = __bfloat162float(__float2bfloat16(val));
                }
    }
    return blob;
}

// The blob of pack_tail for CTA pairs: every B operand [blocks][2 chunks][N][16 B] sliced along N, rank r keeps the columns
// [r N/2, (r + 1) N/2); the floats unchanged.  -> [2 ranks][TL_WBYTES_H]
static std::vector<uint8_t> split_tail_blob(const std::vector<uint8_t> &blob) {
    std::vector<uint8_t> out((size_t)2 * TL_WBYTES_H, 0);
    struct Op { size_t off; int blocks, n; };
    const Op ops[] = {{TL_OFF_B3, 6, 128}, {TL_OFF_B3V, 8, 64}, {TL_OFF_B4, 5, 48}, {TL_OFF_B4 + 480 * 16, 5, 80}, {TL_OFF_B4 + 1280 * 16, 5, 80},
                      {TL_OFF_B4 + 2080 * 16, 5, 48}, {TL_OFF_B4V, 20, 16}, {TL_OFF_BIAS3, 1, 128}};
    for (int r = 0; r < 2; r++) {
        uint8_t *dst = out.data() + (size_t)r * TL_WBYTES_H;
        for (const Op &o : ops)
            for (int b = 0; b < o.blocks * 2; b++)          // (block, chunk)
                memcpy(dst + o.off / 2 + (size_t)b * (o.n / 2) * 16, blob.data() + o.off + ((size_t)b * o.n + (size_t)r * (o.n / 2)) * 16,
                       (size_t)(o.n / 2) * 16);
        memcpy(dst + TL_OFF_AUX / 2, blob.data() + TL_OFF_AUX, TL_AUX_FLOATS * 4);
    }
    return out;
}

struct Uploader {
    std::vector<char> host;
    std::vector<std::pair<void **, size_t>> fix;
    template <class T> void add(T **dst, const T *src, size_t n) {
        size_t off = (host.size() + 255) & ~(size_t)255;
        host.resize(off + n * sizeof(T));
        memcpy(host.data() + off, src, n * sizeof(T));
        fix.push_back({reinterpret_cast<void **>(dst), off});
    }
    template <class T> void add(T **dst, const std::vector<T> &v) { add(dst, v.data(), v.size()); }
};

// conv (8 -> 8, 3x3 'same') + bias + ReLU + 2x2 max-pool of a field that only depends on the border class of a pixel --
// lower[(cy * 3 + cx) * 8 + c] for class (first / middle / last row, first / middle / last column) of an n x n grid -- gives a
// field of the same kind on the n/2 x n/2 grid.  bf16 weights and activations, fp32 sums, like the kernels.
static std::vector<__nv_bfloat16> pool_of_class_field(const std::vector<float> &lower, int n, const std::vector<float> &w,
                                                      const std::vector<float> &b) {
    std::vector<__nv_bfloat16> up((size_t)9 * 8);
    auto cls = [&](int i) { return i == 0 ? 0 : (i == n - 1 ? 2 : 1); };
    const int no = n / 2;
    for (int cy = 0; cy < 3; cy++)
        for (int cx = 0; cx < 3; cx++)
            for (int co = 0; co < 8; co++) {
                float best = 0.f;                          // ReLU
                for (int i = 0; i < 2; i++)
                    for (int j = 0; j < 2; j++) {
                        const int y = 2 * (cy == 0 ? 0 : (cy == 2 ? no - 1 : no / 2)) + i, x = 2 * (cx == 0 ? 0 : (cx == 2 ? no - 1 : no / 2)) + j;
                        float sacc = b[co];
                        for (int dy = 0; dy < 3; dy++)
                            for (int dx = 0; dx < 3; dx++) {
                                const int yy = y + dy - 1, xx = x + dx - 1;
                                if (yy < 0 || yy >= n || xx < 0 || xx >= n) continue;
                                for (int ci = 0; ci < 8; ci++)
                                    sacc += lower[(size_t)(cls(yy) * 3 + cls(xx)) * 8 + ci] *
                                            __bfloat162float(__float2bfloat16(w[((size_t)(dy * 3 + dx) * 8 + ci) * 8 + co]));
                            }
                        best = std::max(best, sacc);
                    }
                up[(size_t)(cy * 3 + cx) * 8 + co] = __float2bfloat16(best);
            }
    return up;
}

// conv (8 -> 8) as the B operand of the sparse tensor trunk (ofb_policy_st.cu): one M row per pooled cell, K = its 4 x 4 input
// patch (position r * 4 + c) x 8 channels, N = the cell's 2 x 2 conv pixels (i * 2 + j) x 8 channels: pixel (i, j) reads patch
// entry (r, c) with tap (r - i, c - j).  [16 pos][32 n][8 cin] = [8 k-steps][2 chunks][32][8]
static std::vector<__nv_bfloat16> pack_cell_patch(const std::vector<float> &w) {
    std::vector<__nv_bfloat16> st((size_t)16 * 32 * 8, __float2bfloat16(0.f));
    for (int pos = 0; pos < 16; pos++)
        for (int n = 0; n < 32; n++)
            for (int ci = 0; ci < 8; ci++) {
                const int dy = (pos >> 2) - ((n >> 3) >> 1), dx = (pos & 3) - ((n >> 3) & 1), co = n & 7;
                if (dy < 0 || dy > 2 || dx < 0 || dx > 2) continue;
                st[((size_t)pos * 32 + n) * 8 + ci] = __float2bfloat16(w[((size_t)(dy * 3 + dx) * 8 + ci) * 8 + co]);
            }
    return st;
}

// Folds the Keras weights (BN into the convs, bilinear x2 into output phases, operand packs of the tensor-core kernels) into
// the upload blob.  The blob's layout depends only on the architecture, so ofb_policy_set_weights can rebuild it in place.
static void build_weight_blob(const ofb_policy_weights *wh, Uploader &up, PolicyDev &d, float &u4_bias, int legacy) {
    std::vector<float> w, b;
    up.host.reserve((size_t)4 << 20);                       // (the blob is 2.9 MB: no reallocation while it grows)
    const PhaseTab PT = phase_tab(legacy);
    // trunk
    fold_conv(wh->conv[0], 2, 8, w, b);
    up.add(&d.c1_w, w); up.add(&d.c1_b, b);
    {
        std::vector<float> lut((size_t)2 * 512 * 8, 0.f);
        for (int ch = 0; ch < 2; ch++)
            for (int pat = 0; pat < 512; pat++)
                for (int co = 0; co < 8; co++) {
                    float s = 0.f;
                    for (int t = 0; t < 9; t++)
                        if ((pat >> t) & 1) s += w[((size_t)t * 2 + ch) * 8 + co];
                    lut[((size_t)ch * 512 + pat) * 8 + co] = s;
                }
        up.add(&d.c1_lut, lut);
    }
    std::vector<float> bg1(8);                            // pool1 of an empty arena: conv1 of zeros is its bias whatever the padding
    for (int co = 0; co < 8; co++) bg1[co] = __bfloat162float(__float2bfloat16(std::max(b[co], 0.f)));
    up.add(&d.sp_bg1, bg1);
    std::vector<float> bg_lower;
    for (int l = 0; l < 3; l++) {
        fold_conv(wh->conv[l + 1], 8, 8, w, b);
        if (l == 0) {
            // pool2 of an empty arena per border class: a cell covers conv2 pixels (2Y + i, 2X + j); the taps that fall outside
            // the 200 x 200 grid read the zero padding.  bf16 weights and activations, fp32 sums, like the kernels.
            std::vector<__nv_bfloat16> bg2((size_t)9 * 8);
            for (int cy = 0; cy < 3; cy++)
                for (int cx = 0; cx < 3; cx++)
                    for (int co = 0; co < 8; co++) {
                        float best = 0.f;                      // ReLU
                        for (int i = 0; i < 2; i++)
                            for (int j = 0; j < 2; j++) {
                                const int y = (cy == 0 ? 0 : (cy == 2 ? 198 : 100)) + i, x = (cx == 0 ? 0 : (cx == 2 ? 198 : 100)) + j;
                                float s = b[co];
                                for (int dy = 0; dy < 3; dy++)
                                    for (int dx = 0; dx < 3; dx++) {
                                        if (y + dy - 1 < 0 || y + dy - 1 > 199 || x + dx - 1 < 0 || x + dx - 1 > 199) continue;
                                        for (int ci = 0; ci < 8; ci++)
                                            s += bg1[ci] * __bfloat162float(__float2bfloat16(w[((size_t)(dy * 3 + dx) * 8 + ci) * 8 + co]));
                                    }
                                best = std::max(best, s);
                            }
                        bg2[(size_t)(cy * 3 + cx) * 8 + co] = __float2bfloat16(best);
                    }
            up.add(&d.sp_bg2, bg2);
        }
        b.resize(16, 0.f);
        up.add(&d.cw[l], pack_taps(w, 8, 8, 16));
        if (l == 0) up.add(&d.c2_tz, pack_toeplitz(w, 8, 8, 8));
        {
            // sparse tensor trunk: the three 8 -> 8 convolutions as cell-patch operands, and the empty-arena value of a pooled
            // cell per border class at every level (each level's classes follow from the level below; level 2 = sp_bg2 above)
            if (l == 0) {
                bg_lower.assign((size_t)9 * 8, 0.f);
                for (int k = 0; k < 9; k++)
                    for (int c = 0; c < 8; c++) bg_lower[(size_t)k * 8 + c] = bg1[c];
            }
            const std::vector<__nv_bfloat16> bgu = pool_of_class_field(bg_lower, 200 >> l, w, b);
            for (size_t k = 0; k < bgu.size(); k++) bg_lower[k] = __bfloat162float(bgu[k]);
            if (l == 1) up.add(&d.sp_bg3, bgu);
            if (l == 2) up.add(&d.sp_bg4, bgu);
            up.add(l == 0 ? &d.c2_st : (l == 1 ? &d.c3_st : &d.c4_st), pack_cell_patch(w));
        }
        up.add(&d.cb[l], b);
    }
    // dense1: rows 0..7 = vector slice (fp32), rows 8..5007 = flat slice (bf16)
    {
        const float *k = wh->dense1.kernel;
        up.add(&d.d1_wv, k, 8 * 100);
        // (1.1 M values on the path of every weight refresh -- Trainer.replay -> ofb_policy_set_weights: the round-to-nearest-even
        //  conversion inline on the bit pattern, the transpose in blocks of 64 rows that stay in L1)
        auto bf16_bits = [](float f) -> uint16_t {
            uint32_t u;
            memcpy(&u, &f, 4);
            if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);     // NaN stays NaN
            return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
        };
        std::vector<__nv_bfloat16> wf((size_t)POL_FLAT * 100), wt((size_t)128 * POL_FLAT_PITCH);
        uint16_t *wf16 = reinterpret_cast<uint16_t *>(wf.data()), *wt16 = reinterpret_cast<uint16_t *>(wt.data());
        memset(wt16, 0, wt.size() * 2);
        const float *kf = k + 8 * 100;
        for (size_t e = 0; e < (size_t)POL_FLAT * 100; e++) wf16[e] = bf16_bits(kf[e]);
        for (int i0 = 0; i0 < POL_FLAT; i0 += 64) {
            const int i1 = std::min(i0 + 64, (int)POL_FLAT);
            for (int j = 0; j < 100; j++)
                for (int i = i0; i < i1; i++) wt16[(size_t)j * POL_FLAT_PITCH + i] = wf16[(size_t)i * 100 + j];
        }
        up.add(&d.d1_wf, wf); up.add(&d.d1_wt, wt);
        up.add(&d.d1_b, wh->dense1.bias, 100);
    }
    up.add(&d.d2_w, wh->dense2.kernel, 100 * 50); up.add(&d.d2_b, wh->dense2.bias, 50);
    up.add(&d.o1_w, wh->output1.kernel, 50 * 2); up.add(&d.o1_b, wh->output1.bias, 2);
    up.add(&d.ud_w, wh->updense1.kernel, 100 * 625); up.add(&d.ud_b, wh->updense1.bias, 625);
    // the folded weights per border class of the low-res pixel: on the image's first / last row or column the convolution's zero
    // padding and the upsampling's edge clamp change the fold (low_border / high_border), not the code that applies it
    auto class_folds = [&](const std::vector<float> &wc, int cin, int cout) {
        std::vector<float> all;
        const PhaseTab tabs[3] = {low_border(PT), PT, high_border(PT)};
        for (int cy = 0; cy < 3; cy++)
            for (int cx = 0; cx < 3; cx++) {
                const std::vector<float> f = fold_phase(wc, cin, cout, tabs[cy], tabs[cx]);
                all.insert(all.end(), f.begin(), f.end());
            }
        return all;
    };
    fold_conv(wh->upconv[0], 1, 2, w, b); up.add(&d.u1_w, w); up.add(&d.u1_b, b); up.add(&d.u1_pw, fold_phase(w, 1, 2, PT, PT));
    up.add(&d.u1_cw, class_folds(w, 1, 2));
    fold_conv(wh->upconv[1], 2, 4, w, b); up.add(&d.u2_w, w); up.add(&d.u2_b, b); up.add(&d.u2_pw, fold_phase(w, 2, 4, PT, PT));
    up.add(&d.u2_cw, class_folds(w, 2, 4));
    {
        // upconv2 as a tcgen05 kind::tf32 B operand (k_heads): [5 k-steps][2 chunks][32 n = xo * 16 + phase * 4 + co][4 lanes = pixel
        // of the chunk * 2 + cin]; chunk (u, c) of the 3 x 3 chunk neighbourhood holds the input pixels 2k + 2c + p, the output pixel
        // 2k + xo reads input pixel 2k + xo + v - 1  ->  v = 2c + p - xo + 1
        const std::vector<float> pf2 = fold_phase(w, 2, 4, PT, PT);
        std::vector<float> tf((size_t)5 * 2 * 32 * 4, 0.f);
        for (int idx = 0; idx < 9; idx++) {
            const int u = idx / 3, c = idx % 3 - 1, j = idx / 2, ch = idx % 2;
            for (int n = 0; n < 32; n++)
                for (int l = 0; l < 4; l++) {
                    const int xo = n >> 4, ph = (n >> 2) & 3, co = n & 3, pp = l >> 1, ci = l & 1, v = 2 * c + pp - xo + 1;
                    if (v < 0 || v > 2) continue;
                    float wv = pf2[((size_t)(u * 3 + v) * 2 + ci) * 16 + ph * 4 + co];
                    uint32_t bits;                        // round to tf32 (10 mantissa bits), nearest even: the pipe would truncate
                    memcpy(&bits, &wv, 4);
                    bits = (bits + 0xFFFu + ((bits >> 13) & 1u)) & ~0x1FFFu;
                    memcpy(&wv, &bits, 4);
                    tf[(((size_t)j * 2 + ch) * 32 + n) * 4 + l] = wv;
                }
        }
        up.add(&d.u2_tf, tf);
    }
    fold_conv(wh->upconv[2], 4, 8, w, b);
    const std::vector<float> w3 = w, b3 = b;
    up.add(&d.u3_w, w); up.add(&d.u3_b, b);
    { const std::vector<float> pf = fold_phase(w, 4, 8, PT, PT);
      up.add(&d.u3_pw, pack_taps(pf, 4, 32, 32));
      up.add(&d.u3_tz, pack_toeplitz(pf, 4, 32, 4)); }
    { std::vector<float> pb(32); for (int n = 0; n < 32; n++) pb[n] = b[n % 8]; up.add(&d.u3_pb, pb); }
    fold_conv(wh->upconv[3], 8, 1, w, b);
    up.add(&d.u4_w, w); up.add(&d.u4_b, b);
    { const std::vector<float> pf = fold_phase(w, 8, 1, PT, PT);
      up.add(&d.u4_pw, pack_taps(pf, 8, 4, 16));
      up.add(&d.u4_tz, pack_toeplitz(pf, 8, 4, 8)); }
    {
        const std::vector<uint8_t> tb = pack_tail(w3, b3, w, legacy);
        up.add(&d.tail_blob, tb);
        up.add(&d.tail_blob2, split_tail_blob(tb));
    }
    u4_bias = b[0];
    { std::vector<float> pb(16, 0.f); for (int n = 0; n < 4; n++) pb[n] = b[0]; up.add(&d.u4_pb, pb); }

}

extern "C" int ofb_policy_create(const ofb_policy_weights *wh, int device, int max_ships, ofb_policy **out) {
    return ofb_policy_create_opts(wh, device, max_ships, 0, out);
}

extern "C" int ofb_policy_create_opts(const ofb_policy_weights *wh, int device, int max_ships, int flags, ofb_policy **out) {
    if (!wh || !out || (flags & ~(OFB_POLICY_BILINEAR_TF1 | OFB_POLICY_UNFUSED_TAIL | OFB_POLICY_DENSE_TRUNK | OFB_POLICY_CC_SPARSE_TRUNK | OFB_POLICY_UNFUSED_TRUNK | OFB_POLICY_TAIL_PAIR))) {
        ofb_set_error("ofb_policy_create: bad argument");
        return OFB_E_ARG;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        ofb_set_error("ofb_policy_create: no CUDA device (libofb has no CPU fallback)");
        return OFB_E_CUDA;
    }
    OFB_CUDA_CHECK(cudaSetDevice(device));
    if (max_ships <= 0) max_ships = 1024;
    ofb_policy *p = new (std::nothrow) ofb_policy();
    if (!p) return OFB_E_NOMEM;
    memset(p, 0, sizeof(*p));
    p->device = device;
    p->max_ships = max_ships;
    p->engine = OFB_ENGINE_TENSOR;
    // trunk12: the sparse CUDA-core kernel (ofb_policy_sp.cu) is the default -- 26 % faster than the dense tcgen05 one over a
    // 200-frame episode of the default arena, 12 % slower at the laser peak around frame 30 (profiles/r01_step_tuning.md);
    // OFB_POLICY_DENSE_TRUNK=1 selects the dense kernel when the handle is created
    { const char *e = getenv("OFB_POLICY_DENSE_TRUNK"); p->dense_trunk = ((e && *e && *e != '0') || (flags & OFB_POLICY_DENSE_TRUNK)) ? 1 : 0; }
    { const char *e = getenv("OFB_POLICY_CC_SPARSE_TRUNK"); if (!p->dense_trunk && ((e && *e && *e != '0') || (flags & OFB_POLICY_CC_SPARSE_TRUNK))) p->dense_trunk = 2; }
    // the tail: fused upconv3 -> upconv4 -> argmax by default; the two-kernel form is kept for A / B measurements
    { const char *e = getenv("OFB_POLICY_UNFUSED_TRUNK"); p->unfused_trunk = ((e && *e && *e != '0') || (flags & OFB_POLICY_UNFUSED_TRUNK)) ? 1 : 0; }
    { const char *e = getenv("OFB_POLICY_UNFUSED_TAIL"); p->unfused_tail = ((e && *e && *e != '0') || (flags & OFB_POLICY_UNFUSED_TAIL)) ? 1 : 0; }
    { const char *e = getenv("OFB_POLICY_TAIL_PAIR"); p->tail_pair = ((e && *e && *e != '0') || (flags & OFB_POLICY_TAIL_PAIR)) ? 1 : 0; }
    p->bilinear_legacy = (flags & OFB_POLICY_BILINEAR_TF1) ? 1 : 0;
    p->prof = new std::vector<ProfEvent>();

    Uploader up;
    PolicyDev &d = p->w;
    build_weight_blob(wh, up, d, p->u4_bias, p->bilinear_legacy);
    p->arena_bytes = up.host.size();

    cudaError_t e = cudaMalloc(&p->arena_blob, up.host.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->arena_blob, up.host.data(), up.host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        ofb_set_error("ofb_policy_create: weight upload failed: %s", cudaGetErrorString(e));
        cudaFree(p->arena_blob); delete p; return OFB_E_CUDA;
    }
    for (auto &f : up.fix) *f.first = static_cast<char *>(p->arena_blob) + f.second;
    d.bil_legacy = p->bilinear_legacy;

    // workspace
    const size_t C = (size_t)max_ships;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    // (pool1 / pool2 / pool3 / up3 -- 1.5 MB per ship, needed only by the alternative kernels and the validation taps -- are
    //  allocated on first use: ensure_wide_workspace)
    const size_t o_fl = take(C * POL_FLAT_PITCH * 2), o_hf = take(C * 100 * 4);
    const size_t o_u2 = take(C * POL_UP2_ITEM * 2);
    const size_t o_av = take(C * AMAX_PARTS * 4), o_ai = take(C * AMAX_PARTS * 4);
    const size_t o_sc = take((size_t)ST_MAX_CTAS * ST_SCRATCH_CELLS * 16);
    e = cudaMalloc(&p->work_blob, off);
    if (e != cudaSuccess) {
        ofb_set_error("ofb_policy_create: cudaMalloc(workspace %zu bytes for %d ships) failed: %s", off, max_ships, cudaGetErrorString(e));
        cudaFree(p->arena_blob); delete p; return OFB_E_NOMEM;
    }
    char *wb = static_cast<char *>(p->work_blob);
    cudaMemset(wb + o_fl, 0, C * POL_FLAT_PITCH * 2);            // K padding of dense1 must read as zero
    cudaMemset(wb + o_u2, 0, C * POL_UP2_ITEM * 2);              // channels 4..7 of up2 (plane layout) / the never-written entries of the pairs layout stay zero
    p->ws.pool1 = p->ws.pool2 = p->ws.pool3 = p->ws.up3 = nullptr;
    p->wide_blob = nullptr;
    p->ws.flat = reinterpret_cast<__nv_bfloat16 *>(wb + o_fl);
    p->ws.hflat = reinterpret_cast<float *>(wb + o_hf);
    p->ws.up2 = reinterpret_cast<__nv_bfloat16 *>(wb + o_u2);
    p->ws.amax_val = reinterpret_cast<float *>(wb + o_av);
    p->ws.amax_idx = reinterpret_cast<int *>(wb + o_ai);
    p->ws.st_scratch = reinterpret_cast<uint4 *>(wb + o_sc);
    OFB_CUDA_CHECK(cudaDeviceSynchronize());
    *out = p;
    return OFB_OK;
}

extern "C" int ofb_policy_destroy(ofb_policy *p) {
    if (!p) return OFB_OK;
    cudaSetDevice(p->device);
    cudaFree(p->arena_blob);
    cudaFree(p->work_blob);
    cudaFree(p->wide_blob);
    if (p->prof) {
        for (auto &e : *static_cast<std::vector<ProfEvent> *>(p->prof)) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
        delete static_cast<std::vector<ProfEvent> *>(p->prof);
    }
    delete p;
    return OFB_OK;
}

// New weights into an existing handle (after a Trainer.fit step): same folding as ofb_policy_create, copied over the
// resident blob once `stream` has drained.  The workspace is untouched.
extern "C" int ofb_policy_set_weights(ofb_policy *p, const ofb_policy_weights *wh, void *stream) {
    if (!p || !wh) { ofb_set_error("ofb_policy_set_weights: null argument"); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaSetDevice(p->device));
    Uploader up;
    PolicyDev scratch;
    float u4_bias = 0.f;
    build_weight_blob(wh, up, scratch, u4_bias, p->bilinear_legacy);
    if (up.host.size() != p->arena_bytes) { ofb_set_error("ofb_policy_set_weights: blob size mismatch"); return OFB_E_STATE; }
    OFB_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
    OFB_CUDA_CHECK(cudaMemcpy(p->arena_blob, up.host.data(), up.host.size(), cudaMemcpyHostToDevice));
    p->u4_bias = u4_bias;
    return OFB_OK;
}

// Per-layer device times: enable = 1 starts collecting (CUDA events around every kernel of forward);
// enable = 0 stops, synchronises and returns the accumulated milliseconds per layer in ms_out[9]
// (trunk12, conv3, conv4, dense1, heads, up3, up4, argmax, tail = the fused up3 + up4 + argmax kernel).
extern "C" int ofb_policy_profile(ofb_policy *p, int enable, float *ms_out) {
    if (!p) { ofb_set_error("ofb_policy_profile: null handle"); return OFB_E_ARG; }
    auto *v = static_cast<std::vector<ProfEvent> *>(p->prof);
    if (enable) { p->profiling = 1; return OFB_OK; }
    p->profiling = 0;
    OFB_CUDA_CHECK(cudaDeviceSynchronize());
    float acc[L_COUNT] = {0};
    for (auto &e : *v) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e.a, e.b);
        acc[e.layer] += ms;
        cudaEventDestroy(e.a); cudaEventDestroy(e.b);
    }
    v->clear();
    if (ms_out) for (int i = 0; i < L_COUNT; i++) ms_out[i] = acc[i];
    return OFB_OK;
}

extern "C" int ofb_policy_set_taps(ofb_policy *p, int enable) {
    if (!p) { ofb_set_error("ofb_policy_set_taps: null handle"); return OFB_E_ARG; }
    p->taps = enable ? 1 : 0;
    return OFB_OK;
}

extern "C" int ofb_policy_set_engine(ofb_policy *p, int engine) {
    if (!p || (engine != OFB_ENGINE_TENSOR && engine != OFB_ENGINE_CUDA_CORE)) {
        ofb_set_error("ofb_policy_set_engine: bad argument");
        return OFB_E_ARG;
    }
    if (engine != p->engine) {
        // the engines keep upconv2's output in different layouts; the entries of the pairs layout that k_heads never writes
        // (spare K lanes, plane tails) must read as zero for the fused tail
        OFB_CUDA_CHECK(cudaSetDevice(p->device));
        OFB_CUDA_CHECK(cudaDeviceSynchronize());
        OFB_CUDA_CHECK(cudaMemset(p->ws.up2, 0, (size_t)p->max_ships * POL_UP2_ITEM * 2));
    }
    p->engine = engine;
    return OFB_OK;
}

// ------------------------------------------------------------------------------------------------
// CUDA-core engine
// ------------------------------------------------------------------------------------------------
// conv1 + BN + ReLU + pool straight from the bit maps: one thread per pooled pixel.
__global__ void __launch_bounds__(256)
k_trunk1_cc(const uint32_t *__restrict__ maps, PolicyDev w, __nv_bfloat16 *__restrict__ out) {
    __shared__ float sb[8];
    if (threadIdx.x < 8) sb[threadIdx.x] = w.c1_b[threadIdx.x];
    __syncthreads();
    const int a = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 200 * 200) return;
    const int py = p / 200, px = p % 200;
    const uint32_t *sm = maps + (size_t)a * 2 * POL_WORDS, *lm = sm + POL_WORDS;
    const uint32_t ps = conv1_patch(sm, py, px), pl = conv1_patch(lm, py, px);
    float v[8];
    if ((ps | pl) == 0) {
#pragma unroll
        for (int co = 0; co < 8; co++) v[co] = fmaxf(sb[co], 0.f);
    } else conv1_pool_pixel(ps, pl, w.c1_lut, sb, v);
    *reinterpret_cast<uint4 *>(out + ((size_t)a * 40000 + p) * 8) = pack_bf8(v);
}

// generic 8 -> 8 conv3x3 'same' + bias + ReLU + 2x2 max-pool, NHWC bf16; one thread per pooled pixel
__global__ void __launch_bounds__(128)
k_conv_pool_cc(const __nv_bfloat16 *__restrict__ in, const __nv_bfloat16 *__restrict__ wt, const float *__restrict__ bias,
               __nv_bfloat16 *__restrict__ out, int hin, long long out_item_stride) {
    __shared__ float sw[9 * 8 * 8];                             // [tap][cout][cin]
    __shared__ float sb[8];
    for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) {
        const int t = i / 64, co = (i / 8) % 8, ci = i % 8;
        sw[i] = bf2f(wt[((size_t)t * 16 + co) * 8 + ci]);
    }
    if (threadIdx.x < 8) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const int ho = hin / 2;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= ho * ho) return;
    const int py = p / ho, px = p % ho;
    const __nv_bfloat16 *src = in + (size_t)blockIdx.y * hin * hin * 8;
    float acc[4][8];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int co = 0; co < 8; co++) acc[q][co] = sb[co];
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const int y = 2 * py - 1 + r;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int x = 2 * px - 1 + c;
            if (y < 0 || y >= hin || x < 0 || x >= hin) continue;
            float v[8];
            unpack_bf8(*reinterpret_cast<const uint4 *>(src + ((size_t)y * hin + x) * 8), v);
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int dy = r - i, dx = c - j;
                    if (dy < 0 || dy > 2 || dx < 0 || dx > 2) continue;
                    const float *wp = sw + (dy * 3 + dx) * 64;
#pragma unroll
                    for (int co = 0; co < 8; co++) {
                        float s = acc[i * 2 + j][co];
#pragma unroll
                        for (int ci = 0; ci < 8; ci++) s += v[ci] * wp[co * 8 + ci];
                        acc[i * 2 + j][co] = s;
                    }
                }
        }
    }
    float o[8];
#pragma unroll
    for (int co = 0; co < 8; co++)
        o[co] = fmaxf(fmaxf(fmaxf(acc[0][co], acc[1][co]), fmaxf(acc[2][co], acc[3][co])), 0.f);
    *reinterpret_cast<uint4 *>(out + (size_t)blockIdx.y * out_item_stride + (size_t)p * 8) = pack_bf8(o);
}

// dense1, flat slice: hflat[a][j] = sum_k flat[a][k] * Wf[k][j].  4 arenas per block; 256 threads =
// 128 output lanes (100 used) x 2 halves of every K tile, halves summed through shared memory.
#define D1_ARENAS 4
__global__ void __launch_bounds__(256)
k_dense1_cc(const __nv_bfloat16 *__restrict__ flat, const __nv_bfloat16 *__restrict__ wf, float *__restrict__ hflat, int n_items) {
    constexpr int KT = 1000;
    __shared__ __align__(16) float sf[D1_ARENAS][KT];
    __shared__ float red[D1_ARENAS][128];
    const int a0 = blockIdx.x * D1_ARENAS, j = threadIdx.x & 127, kh = threadIdx.x >> 7;
    float acc[D1_ARENAS];
#pragma unroll
    for (int a = 0; a < D1_ARENAS; a++) acc[a] = 0.f;
    for (int k0 = 0; k0 < POL_FLAT; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < D1_ARENAS * KT; i += blockDim.x) {
            const int a = i / KT, k = i % KT;
            sf[a][k] = (a0 + a < n_items) ? bf2f(flat[(size_t)(a0 + a) * POL_FLAT_PITCH + k0 + k]) : 0.f;
        }
        __syncthreads();
        if (j < 100) {
            const int kb = kh * (KT / 2);
#pragma unroll 2
            for (int k = kb; k < kb + KT / 2; k += 4) {
                float wv[4];
#pragma unroll
                for (int q = 0; q < 4; q++) wv[q] = bf2f(wf[(size_t)(k0 + k + q) * 100 + j]);
#pragma unroll
                for (int a = 0; a < D1_ARENAS; a++) {
                    const float4 f = *reinterpret_cast<const float4 *>(&sf[a][k]);
                    acc[a] += f.x * wv[0];
                    acc[a] += f.y * wv[1];
                    acc[a] += f.z * wv[2];
                    acc[a] += f.w * wv[3];
                }
            }
        }
    }
    if (kh == 1) {
#pragma unroll
        for (int a = 0; a < D1_ARENAS; a++) red[a][j] = acc[a];
    }
    __syncthreads();
    if (kh == 0 && j < 100)
        for (int a = 0; a < D1_ARENAS; a++)
            if (a0 + a < n_items) hflat[(size_t)(a0 + a) * 100 + j] = acc[a] + red[a][j];
}

// 16-byte shared-memory load the compiler may not hoist out of a loop
__device__ __forceinline__ float4 lds128_volatile(const float *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// Border-ring correction of a phase-folded [bilinear x2 -> conv3x3 'same'] for an fp32 [n][n][CIN] source in shared memory
// (see up_ring_correct in ofb_policy_dev.cuh): v[co] -= sum over the taps outside [0, 2n) of w[dy][dx][:, co] . U~.
template <int CIN, int COUT>
__device__ __forceinline__ void ring_correct_f32(const float *__restrict__ src, int n, int Y, int X, const float *__restrict__ w, float *v,
                                                 int legacy) {
    for (int dy = 0; dy < 3; dy++) {
        const int YY = Y + dy - 1;
        for (int dx = 0; dx < 3; dx++) {
            const int XX = X + dx - 1;
            if (YY >= 0 && YY < 2 * n && XX >= 0 && XX < 2 * n) continue;
            int ylo, yhi, xlo, xhi; float wyl, wyh, wxl, wxh;
            bil_tap_ext(YY, ylo, yhi, wyl, wyh, legacy);
            bil_tap_ext(XX, xlo, xhi, wxl, wxh, legacy);
            ylo = min(max(ylo, 0), n - 1); yhi = min(max(yhi, 0), n - 1);
            xlo = min(max(xlo, 0), n - 1); xhi = min(max(xhi, 0), n - 1);
#pragma unroll
            for (int ci = 0; ci < CIN; ci++) {
                const float u = wyl * (wxl * src[(ylo * n + xlo) * CIN + ci] + wxh * src[(ylo * n + xhi) * CIN + ci]) +
                                wyh * (wxl * src[(yhi * n + xlo) * CIN + ci] + wxh * src[(yhi * n + xhi) * CIN + ci]);
#pragma unroll
                for (int co = 0; co < COUT; co++) v[co] -= u * w[((dy * 3 + dx) * CIN + ci) * COUT + co];
            }
        }
    }
}

// Everything between dense1's flat part and upconv3, per ship, in fp32 on CUDA cores:
// dense1 (vector slice + bias + ReLU), dense2, output1 (+ argmax), updense1, upsampling1 + upconv1, upsampling2 + upconv2.
// The two small up-convolutions run in the phase-folded form on their low-res grids (25 x 25 and 50 x 50), so no upsampled
// map is materialised and the block needs ~26 KB of shared memory (8 blocks per SM instead of 2).
// upconv2's output pixel (y, x) (4 channels = 8 bytes) in the "pairs" layout the fused tail stages with bulk copies
// (ofb_policy_tail.cu): plane A entry (y+1)*26 + s = pixels (4s-1, 4s), plane B = pixels (4s+1, 4s+2), plane S (spare K
// lanes) = L[y][0] in the low half of entry s = 0 and L[y][99] in the high half of entry s = 24; rows -1 / 100 and columns
// -1 / 100 replicate the edge.
__device__ __forceinline__ void up2_pairs_store_row(uint8_t *item, int row, int x, uint2 v) {
    const int s = (x + 1) >> 2, rr = (x + 1) & 3;
    uint8_t *e = item + ((size_t)(rr >> 1) * TL_UP2_PLANE + (size_t)row * TL_P + s) * 16 + (rr & 1) * 8;
    *reinterpret_cast<uint2 *>(e) = v;
    if (x == 0) {
        *reinterpret_cast<uint2 *>(item + ((size_t)row * TL_P) * 16) = v;                                   // x = -1 := x = 0
        *reinterpret_cast<uint2 *>(item + ((size_t)2 * TL_UP2_PLANE + (size_t)row * TL_P) * 16) = v;        // spare, left
    }
    if (x == 99) {
        *reinterpret_cast<uint2 *>(item + ((size_t)row * TL_P + 25) * 16 + 8) = v;                          // x = 100 := x = 99
        *reinterpret_cast<uint2 *>(item + ((size_t)2 * TL_UP2_PLANE + (size_t)row * TL_P + 24) * 16 + 8) = v;   // spare, right
    }
}
__device__ __forceinline__ void up2_pairs_store(uint8_t *item, int y, int x, uint2 v) {
    up2_pairs_store_row(item, y + 1, x, v);
    if (y == 0) up2_pairs_store_row(item, 0, x, v);
    if (y == 99) up2_pairs_store_row(item, 101, x, v);
}

__device__ __forceinline__ float to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}
#define HEADS_G 4                                        // ships per CTA: the dense layers' weights are read once per group
// upconv1's output a1 (50 x 50 x 2 fp32).  CUDA-core upconv2: plain [y][x][c].  Tensor-core upconv2 (tcgen05 kind::tf32, the product
// path): a1 IS the A operand -- 16-byte chunks of 2 pixels x 2 channels, 27 chunks per row (one halo chunk on either side) and one
// halo row above / below, replicated edges: chunk (y + 1) * 27 + (x >> 1) + 1.
#define HEADS_A1_CHUNKS (52 * 27 + 32)
#define HEADS_A1_FLOATS (HEADS_A1_CHUNKS * 4)
#define HEADS_U2_TILES 11                                // 52 * 27 = 1404 M rows (chunks incl. halos) in tiles of 128
#define HEADS_SMEM_FLOATS (HEADS_G * (128 + 64 + 640) + HEADS_A1_FLOATS + 80 + 304 + 32 + 80 + 5 * 2 * 32 * 4 + 16)
__global__ void __launch_bounds__(256, 4)
k_heads(const float *__restrict__ hflat, const float *__restrict__ vec, PolicyDev w, int ships_per_arena, int n_ships, float *__restrict__ act_out,
        int *__restrict__ iaction_out, __nv_bfloat16 *__restrict__ up2_out, int plane_layout) {
    extern __shared__ __align__(16) float sm[];
    float *hh = sm, *dd2 = hh + HEADS_G * 128, *u_all = dd2 + HEADS_G * 64, *a1 = u_all + HEADS_G * 640, *w1p = a1 + HEADS_A1_FLOATS,
          *w2p = w1p + 80, *w1r = w2p + 304, *w2r = w1r + 32, *u2b = w2r + 80;       // u2b: upconv2 as a tf32 B operand (16-byte aligned)
    uint64_t *mbar = reinterpret_cast<uint64_t *>(u2b + 5 * 2 * 32 * 4);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mbar + 1);
    const bool tc2 = plane_layout == 2;                  // the fused-tail path: upconv2 on the tensor pipe
    // (the tensor pipe reads fp32 operands as tf32 by dropping mantissa bits: round to nearest here instead)
#define A1V(v) (tc2 ? to_tf32(v) : (v))
#define A1I(Y, X) (tc2 ? ((((Y) + 1) * 27 + ((X) >> 1) + 1) * 4 + ((X) & 1) * 2) : (((Y) * 50 + (X)) * 2))
    const int s0 = blockIdx.x * HEADS_G, tid = threadIdx.x, nt = blockDim.x;
    const int ng = min(HEADS_G, n_ships - s0);
    // folded weights: upconv1 [9][1][8] (+ un-phased [9][1][2] for the ring), upconv2 [9][2][16] (+ [9][2][4]), biases
    for (int i = tid; i < 72; i += nt) w1p[i] = w.u1_pw[i];
    for (int i = tid; i < 288; i += nt) w2p[i] = w.u2_pw[i];
    if (tid < 18) w1r[tid] = w.u1_w[tid];
    for (int i = tid; i < 72; i += nt) w2r[i] = w.u2_w[i];
    if (tid < 2) w1p[72 + tid] = w.u1_b[tid];
    if (tid < 4) w2p[288 + tid] = w.u2_b[tid];
    uint32_t tmem_base = 0;
    if (tc2) {
        for (int i = tid; i < 5 * 2 * 32 * 4; i += nt) u2b[i] = w.u2_tf[i];
        for (int i = 52 * 27 * 4 + tid; i < HEADS_A1_FLOATS; i += nt) a1[i] = 0.f;     // read (against zero weights) by the last rows' pad chunk
        if (tid == 0) { mbar_init(mbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_slot;
    }
    uint32_t n_mma = 0;
    // dense1: [vector(8), flat(5000)] -> 100, ReLU      (qlearnIA_V2.py:154-155)
    if (tid < 100) {
        float wv[8];
#pragma unroll
        for (int i = 0; i < 8; i++) wv[i] = w.d1_wv[i * 100 + tid];
        const float bias = w.d1_b[tid];
#pragma unroll
        for (int g = 0; g < HEADS_G; g++) {
            if (g >= ng) break;
            const int s = s0 + g;
            float acc = bias;
#pragma unroll
            for (int i = 0; i < 8; i++) acc += vec[(size_t)s * 8 + i] * wv[i];
            acc += hflat[(size_t)(s / ships_per_arena) * 100 + tid];
            hh[g * 128 + tid] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    // dense2 100 -> 50 ReLU (:158); updense1 100 -> 625 ReLU (:163): every weight is loaded once for the group's ships
    for (int j = tid; j < 50 + 625; j += nt) {
        float acc[HEADS_G];
        if (j < 50) {
#pragma unroll
            for (int g = 0; g < HEADS_G; g++) acc[g] = w.d2_b[j];
            for (int i = 0; i < 100; i++) {
                const float wt = w.d2_w[i * 50 + j];
#pragma unroll
                for (int g = 0; g < HEADS_G; g++) acc[g] += hh[g * 128 + i] * wt;
            }
#pragma unroll
            for (int g = 0; g < HEADS_G; g++) dd2[g * 64 + j] = fmaxf(acc[g], 0.f);
        } else {
            const int q = j - 50;
#pragma unroll
            for (int g = 0; g < HEADS_G; g++) acc[g] = w.ud_b[q];
#pragma unroll 10
            for (int i = 0; i < 100; i++) {
                const float wt = w.ud_w[i * 625 + q];
#pragma unroll
                for (int g = 0; g < HEADS_G; g++) acc[g] += hh[g * 128 + i] * wt;
            }
#pragma unroll
            for (int g = 0; g < HEADS_G; g++) u_all[g * 640 + q] = fmaxf(acc[g], 0.f);
        }
    }
    __syncthreads();
    // output1 50 -> 2 linear (:160) and its argmax (:218, ties -> lowest index)
    if (tid < ng) {
        const int s = s0 + tid;
        const float *d2 = dd2 + tid * 64;
        float a0 = w.o1_b[0], a1v = w.o1_b[1];
        for (int i = 0; i < 50; i++) { a0 += d2[i] * w.o1_w[i * 2]; a1v += d2[i] * w.o1_w[i * 2 + 1]; }
        if (act_out) { act_out[(size_t)s * 2] = a0; act_out[(size_t)s * 2 + 1] = a1v; }
        if (iaction_out) iaction_out[s] = a1v > a0 ? 1 : 0;
    }
  for (int g = 0; g < ng; g++) {                          // the pointer head's small up-convolutions, one ship at a time
    const int s = s0 + g;
    const float *u = u_all + g * 640;
    if (g > 0) __syncthreads();                           // the previous ship's a1 has been consumed
    // upsampling1 + upconv1 1 -> 2 + BN + ReLU (:166-169): one thread per pixel of the 25 x 25 grid, 4 phases x 2 channels
    for (int p = tid; p < 625; p += nt) {
        const int i = p / 25, j = p % 25;
        float acc[8];
#pragma unroll
        for (int n = 0; n < 8; n++) acc[n] = w1p[72 + (n & 1)];
#pragma unroll
        for (int uu = 0; uu < 3; uu++)
#pragma unroll
            for (int vv = 0; vv < 3; vv++) {
                const float x = u[min(max(i + uu - 1, 0), 24) * 25 + min(max(j + vv - 1, 0), 24)];
#pragma unroll
                for (int n = 0; n < 8; n++) acc[n] += x * w1p[(uu * 3 + vv) * 8 + n];
            }
#pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
            if (Y == 0 || Y == 49 || X == 0 || X == 49) continue;        // border ring: second loop
            a1[A1I(Y, X)] = A1V(fmaxf(acc[ph * 2], 0.f));
            a1[A1I(Y, X) + 1] = A1V(fmaxf(acc[ph * 2 + 1], 0.f));
        }
    }
#pragma unroll 1
    for (int rp = tid; rp < 4 * 49; rp += nt) {                          // the 196 pixels of the border ring, with the correction
        const int side = rp / 49, q = rp % 49;
        const int Y = side == 0 ? 0 : (side == 1 ? 49 : (side == 2 ? q + 1 : q)), X = side == 0 ? q : (side == 1 ? q + 1 : (side == 2 ? 0 : 49));
        const int i = Y >> 1, j = X >> 1, ph = (Y & 1) * 2 + (X & 1);
        float o[2] = {w1p[72], w1p[73]};
        const float *wc = w.u1_cw + ((i == 0 ? 0 : (i == 24 ? 2 : 1)) * 3 + (j == 0 ? 0 : (j == 24 ? 2 : 1))) * 72;    // border-class fold
        for (int t = 0; t < 9; t++) {
            const float x = u[min(max(i + t / 3 - 1, 0), 24) * 25 + min(max(j + t % 3 - 1, 0), 24)];
            o[0] += x * __ldg(wc + t * 8 + ph * 2);
            o[1] += x * __ldg(wc + t * 8 + ph * 2 + 1);
        }
        a1[A1I(Y, X)] = A1V(fmaxf(o[0], 0.f));
        a1[A1I(Y, X) + 1] = A1V(fmaxf(o[1], 0.f));
    }
    __syncthreads();
    // upsampling2 + upconv2 2 -> 4 + BN + ReLU (:172-175): one thread per pixel of the 50 x 50 grid, 4 phases x 4 channels
    // -> bf16 with channels padded to 8
    __nv_bfloat16 *dst = up2_out + (size_t)s * POL_UP2_ITEM;   // plane_layout 1: de-interleaved by x mod 4 for k_tz_up3; 2: pairs for k_tz_tail
    uint8_t *dstp = reinterpret_cast<uint8_t *>(up2_out) + (size_t)s * TL_UP2_ITEM_BYTES;
    if (tc2) {
        // ---- upconv2 on the tensor pipe (tcgen05 kind::tf32): one M row per 2-pixel chunk of a1, K = the 3 x 3 chunks around it
        //      (9 x 4 values + 4 lanes of zero weights = 5 K-steps), N = 2 pixels x 4 phases x 4 channels.  a1 is the A operand as it
        //      lies (its halo chunks / rows replicate the edges); the border ring of the output comes from the scalar loop below.
        for (int q = tid; q < 100; q += nt) {             // halo chunks of rows 0 .. 49: x = -1 := x = 0, x = 50 := x = 49
            const int r = 1 + (q >> 1), side = q & 1;
            const float2 e = *reinterpret_cast<const float2 *>(a1 + A1I(r - 1, side ? 49 : 0));
            *reinterpret_cast<float4 *>(a1 + (r * 27 + (side ? 26 : 0)) * 4) = make_float4(e.x, e.y, e.x, e.y);
        }
        __syncthreads();
        for (int q = tid; q < 2 * 27; q += nt) {          // halo rows: row -1 := row 0, row 50 := row 49
            const int side = q / 27, c = q % 27;
            *reinterpret_cast<float4 *>(a1 + ((side ? 51 : 0) * 27 + c) * 4) = *reinterpret_cast<const float4 *>(a1 + ((side ? 50 : 1) * 27 + c) * 4);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        tc_fence_before();
        __syncthreads();
        const uint32_t a16 = smem_u32(a1) >> 4, b16 = smem_u32(u2b) >> 4;
        constexpr uint32_t ID_TF32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int warp = tid >> 5, grp = warp >> 2;
        for (int t0 = 0; t0 < HEADS_U2_TILES; t0 += 4) {  // rounds of 4 tiles = 4 x 32 TMEM columns
            const int nt4 = min(4, HEADS_U2_TILES - t0);
            if (warp == 0) {
                tc_fence_after();
                const bool leader = elect_one();
                for (int tt = 0; tt < nt4; tt++) {
                    const int m0 = 128 * (t0 + tt);
#pragma unroll
                    for (int j = 0; j < 5; j++) {
                        // chunk pairs (u, c): (0,-1)(0,0) | (0,1)(1,-1) | (1,0)(1,1) | (2,-1)(2,0) | (2,1)(pad); offsets in chunks
                        const int o0 = j == 0 ? -28 : (j == 1 ? -26 : (j == 2 ? 0 : (j == 3 ? 26 : 28)));
                        const uint32_t lbo = j == 1 ? 25u : 1u;
                        const uint64_t ad = smem_desc(a16 + (uint32_t)(m0 + o0), lbo, 8);
                        const uint64_t bd = smem_desc(b16 + (uint32_t)(j * 2 * 32), 32, 8);
                        if (leader)
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_base + (uint32_t)(tt * 32)),
                                         "l"(ad), "l"(bd), "r"(ID_TF32), "r"(j ? 1u : 0u) : "memory");
                    }
                }
                if (leader) tc_commit(mbar);
                __syncwarp();
            }
            mbar_wait(mbar, n_mma & 1u);
            n_mma++;
            tc_fence_after();
            for (int tt = grp; tt < nt4; tt += 2) {       // warps 0-3 drain the even tiles of the round, warps 4-7 the odd ones
                uint32_t r[32];
                tc_ld32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tt * 32), r);
                tc_wait_ld();
                const int m = 128 * (t0 + tt) + (tid & 127), ip = m / 27, sc = m - ip * 27;
                const bool valid = ip >= 1 && ip <= 50 && sc >= 1 && sc <= 25;     // not a halo chunk
                const int i = ip - 1, k = sc - 1;
                const float *bb = w2p + 288;
                // the chunk's 2 x 4 output pixels x = 4k .. 4k+3 of rows y = 2i, 2i+1 in the pairs layout: x = 4k -> plane A entry k
                // (high half), 4k+1 / 4k+2 -> plane B entry k (one 16-byte store), 4k+3 -> plane A entry k+1 (low half).  The border
                // ring (rows 0 / 99, columns 0 / 99) and its replicas belong to the scalar loop below.
#pragma unroll
                for (int pa = 0; pa < 2; pa++) {
                    const int y = 2 * i + pa;
                    uint2 px[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t *rv = r + (q >> 1) * 16 + (pa * 2 + (q & 1)) * 4;
                        px[q] = make_uint2(pack_relu_bf2(__uint_as_float(rv[0]) + bb[0], __uint_as_float(rv[1]) + bb[1]),
                                           pack_relu_bf2(__uint_as_float(rv[2]) + bb[2], __uint_as_float(rv[3]) + bb[3]));
                    }
                    if (valid && y != 0 && y != 99) {
                        uint8_t *e = dstp + ((size_t)(y + 1) * TL_P + k) * 16;
                        if (k != 0) *reinterpret_cast<uint2 *>(e + 8) = px[0];
                        *reinterpret_cast<uint4 *>(e + (size_t)TL_UP2_PLANE * 16) = make_uint4(px[1].x, px[1].y, px[2].x, px[2].y);
                        if (k != 24) *reinterpret_cast<uint2 *>(e + 16) = px[3];
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
        }
    }
    for (int p = tid; p < (tc2 ? 0 : 2500); p += nt) {
        const int i = p / 50, j = p % 50;
        float acc[16];
#pragma unroll
        for (int n = 0; n < 16; n++) acc[n] = w2p[288 + (n & 3)];
#pragma unroll
        for (int uu = 0; uu < 3; uu++)
#pragma unroll
            for (int vv = 0; vv < 3; vv++) {
                const float2 x = *reinterpret_cast<const float2 *>(a1 + A1I(min(max(i + uu - 1, 0), 49), min(max(j + vv - 1, 0), 49)));
                // volatile: keeps the 288 loop-invariant weights in shared memory instead of (spilled) registers
                const float *wp = w2p + (uu * 3 + vv) * 32;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 wa = lds128_volatile(wp + 4 * q), wb = lds128_volatile(wp + 16 + 4 * q);
                    acc[q * 4 + 0] += x.x * wa.x + x.y * wb.x;
                    acc[q * 4 + 1] += x.x * wa.y + x.y * wb.y;
                    acc[q * 4 + 2] += x.x * wa.z + x.y * wb.z;
                    acc[q * 4 + 3] += x.x * wa.w + x.y * wb.w;
                }
            }
#pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            const int y = 2 * i + (ph >> 1), x = 2 * j + (ph & 1);
            if (y == 0 || y == 99 || x == 0 || x == 99) continue;        // border ring: second loop
            const uint2 px = make_uint2(pack_bf2(fmaxf(acc[ph * 4], 0.f), fmaxf(acc[ph * 4 + 1], 0.f)),
                                        pack_bf2(fmaxf(acc[ph * 4 + 2], 0.f), fmaxf(acc[ph * 4 + 3], 0.f)));
            if (plane_layout == 2) { up2_pairs_store(dstp, y, x, px); continue; }
            // full 16-byte pixels (channels 4..7 = 0): partial 32-byte sectors would turn into read-modify-writes in DRAM
            *reinterpret_cast<uint4 *>(dst + (plane_layout ? pol_plane100_off(y, x) : (y * 100 + x) * 8)) = make_uint4(px.x, px.y, 0u, 0u);
        }
    }
#pragma unroll 1
    for (int rp = tid; rp < 4 * 99; rp += nt) {                          // the 396 pixels of the border ring, with the correction
        const int side = rp / 99, q = rp % 99;
        const int y = side == 0 ? 0 : (side == 1 ? 99 : (side == 2 ? q + 1 : q)), x = side == 0 ? q : (side == 1 ? q + 1 : (side == 2 ? 0 : 99));
        const int i = y >> 1, j = x >> 1, ph = (y & 1) * 2 + (x & 1);
        float c[4] = {w2p[288], w2p[289], w2p[290], w2p[291]};
        const float *wc = w.u2_cw + ((i == 0 ? 0 : (i == 49 ? 2 : 1)) * 3 + (j == 0 ? 0 : (j == 49 ? 2 : 1))) * 288;   // border-class fold
        for (int t = 0; t < 9; t++) {
            const float2 v = *reinterpret_cast<const float2 *>(a1 + A1I(min(max(i + t / 3 - 1, 0), 49), min(max(j + t % 3 - 1, 0), 49)));
            const float4 wa = __ldg(reinterpret_cast<const float4 *>(wc + t * 32 + ph * 4)), wb = __ldg(reinterpret_cast<const float4 *>(wc + t * 32 + 16 + ph * 4));
            c[0] += v.x * wa.x + v.y * wb.x;
            c[1] += v.x * wa.y + v.y * wb.y;
            c[2] += v.x * wa.z + v.y * wb.z;
            c[3] += v.x * wa.w + v.y * wb.w;
        }
        if (plane_layout == 2) {
            up2_pairs_store(dstp, y, x, make_uint2(pack_bf2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f)), pack_bf2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f))));
            continue;
        }
        __nv_bfloat16 *q16 = dst + (plane_layout ? pol_plane100_off(y, x) : (y * 100 + x) * 8);
        *reinterpret_cast<uint4 *>(q16) = make_uint4(pack_bf2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f)), pack_bf2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f)), 0u, 0u);
        if (plane_layout == 1 && x == 0) *reinterpret_cast<uint4 *>(q16 - 8) = make_uint4(0u, 0u, 0u, 0u);     // the row's halo slot
    }
    if (plane_layout == 1)                                               // halo slots of planes 1..3 (x = 1..3 are interior pixels)
        for (int rp = tid; rp < 3 * 100; rp += nt)
            *reinterpret_cast<uint4 *>(dst + pol_plane100_off(rp % 100, 1 + rp / 100) - 8) = make_uint4(0u, 0u, 0u, 0u);
  }
    if (tc2) {
        tc_fence_before();
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
    }
#undef A1I
#undef A1V
}

// upconv3 (bilinear x2 folded into 4 phases) 4 -> 8 + BN + ReLU: one thread per low-res pixel
__global__ void __launch_bounds__(128)
k_up3_cc(const __nv_bfloat16 *__restrict__ in, PolicyDev w, __nv_bfloat16 *__restrict__ out) {
    __shared__ float sw[9 * 32 * 4];                            // [tap][n][cin<4]
    __shared__ float sb[32], rw[9 * 4 * 8];
    for (int i = threadIdx.x; i < 9 * 32 * 4; i += blockDim.x) sw[i] = bf2f(w.u3_pw[(size_t)(i / 4) * 8 + (i % 4)]);
    for (int i = threadIdx.x; i < 9 * 4 * 8; i += blockDim.x) rw[i] = w.u3_w[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = w.u3_pb[threadIdx.x];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 100 * 100) return;
    const int i = p / 100, j = p % 100;
    const __nv_bfloat16 *L = in + (size_t)blockIdx.y * POL_UP2_ITEM;
    __nv_bfloat16 *dst = out + (size_t)blockIdx.y * POL_UP3_ITEM;
    float acc[32];
#pragma unroll
    for (int n = 0; n < 32; n++) acc[n] = sb[n];
#pragma unroll
    for (int u = 0; u < 3; u++)
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const int yy = min(max(i + u - 1, 0), 99), xx = min(max(j + v - 1, 0), 99);
            float x[8];
            unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)yy * 100 + xx) * 8), x);
            const float *wp = sw + (u * 3 + v) * 128;
#pragma unroll
            for (int n = 0; n < 32; n++)
#pragma unroll
                for (int ci = 0; ci < 4; ci++) acc[n] += x[ci] * wp[n * 4 + ci];
        }
#pragma unroll
    for (int ph = 0; ph < 4; ph++) {
        const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
        float o[8];
#pragma unroll
        for (int co = 0; co < 8; co++) o[co] = acc[ph * 8 + co];
        if (Y == 0 || Y == 199 || X == 0 || X == 199) up_ring_correct<4, 8>(GlobalImage{L, 100}, 100, Y, X, rw, o, w.bil_legacy);
#pragma unroll
        for (int co = 0; co < 8; co++) o[co] = fmaxf(o[co], 0.f);
        *reinterpret_cast<uint4 *>(dst + ((size_t)Y * 200 + X) * 8) = pack_bf8(o);
    }
}

// upconv4 (phase-folded) 8 -> 1, linear, + per-block argmax partials; optional dense map
__global__ void __launch_bounds__(256)
k_up4_cc(const __nv_bfloat16 *__restrict__ in, PolicyDev w, float *__restrict__ ptr_out, float *__restrict__ amax_val,
         int *__restrict__ amax_idx) {
    __shared__ float sw[9 * 4 * 8], rw[72], sv[8];
    __shared__ int si[8];
    for (int i = threadIdx.x; i < 9 * 4 * 8; i += blockDim.x) sw[i] = bf2f(w.u4_pw[(size_t)((i / 32) * 16 + (i / 8) % 4) * 8 + (i % 8)]);
    for (int i = threadIdx.x; i < 72; i += blockDim.x) rw[i] = w.u4_w[i];
    __syncthreads();
    const float pb = w.u4_pb[0];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const __nv_bfloat16 *L = in + (size_t)blockIdx.y * POL_UP3_ITEM;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    if (p < 200 * 200) {
        const int i = p / 200, j = p % 200;
        float acc[4] = {pb, pb, pb, pb};
#pragma unroll
        for (int u = 0; u < 3; u++)
#pragma unroll
            for (int v = 0; v < 3; v++) {
                const int yy = min(max(i + u - 1, 0), 199), xx = min(max(j + v - 1, 0), 199);
                float x[8];
                unpack_bf8(*reinterpret_cast<const uint4 *>(L + ((size_t)yy * 200 + xx) * 8), x);
                const float *wp = sw + (u * 3 + v) * 32;
#pragma unroll
                for (int ph = 0; ph < 4; ph++)
#pragma unroll
                    for (int ci = 0; ci < 8; ci++) acc[ph] += x[ci] * wp[ph * 8 + ci];
            }
#pragma unroll
        for (int ph = 0; ph < 4; ph++) {
            const int Y = 2 * i + (ph >> 1), X = 2 * j + (ph & 1);
            float o = acc[ph];
            if (Y == 0 || Y == 399 || X == 0 || X == 399) up_ring_correct<8, 1>(GlobalImage{L, 200}, 200, Y, X, rw, &o, w.bil_legacy);
            const int idx = Y * 400 + X;
            if (ptr_out) ptr_out[(size_t)blockIdx.y * 160000 + idx] = o;
            if (amax_better(o, idx, bv, bi)) { bv = o; bi = idx; }
        }
    }
    amax_warp(bv, bi);
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        bv = threadIdx.x < 8 ? sv[threadIdx.x] : -INFINITY;
        bi = threadIdx.x < 8 ? si[threadIdx.x] : 0x7fffffff;
        amax_warp(bv, bi);
        if (threadIdx.x == 0) {
            amax_val[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = bv;
            amax_idx[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = bi;
        }
    }
}

// final argmax over the per-block partials: one warp per ship; (x, y) = (k % 400, k / 400)
__global__ void k_argmax_final(const float *__restrict__ val, const int *__restrict__ idx, int parts, int n_ships, int *__restrict__ xy) {
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_ships) return;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < parts; k += 32) {
        const float v = val[(size_t)s * parts + k];
        const int i = idx[(size_t)s * parts + k];
        if (amax_better(v, i, bv, bi)) { bv = v; bi = i; }
    }
    amax_warp(bv, bi);
    if (lane == 0) { xy[s * 2] = bi % POL_W; xy[s * 2 + 1] = bi / POL_W; }
}

// ------------------------------------------------------------------------------------------------
// orchestration
// ------------------------------------------------------------------------------------------------
struct ProfScope {
    ofb_policy *p; cudaStream_t st; cudaEvent_t a; int layer;
    ProfScope(ofb_policy *p_, int layer_, cudaStream_t st_) : p(p_), st(st_), a(nullptr), layer(layer_) {
        if (p->profiling) { cudaEventCreate(&a); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (p->profiling) {
            cudaEvent_t b; cudaEventCreate(&b); cudaEventRecord(b, st);
            static_cast<std::vector<ProfEvent> *>(p->prof)->push_back({a, b, layer});
        }
    }
};
// The intermediates only the alternative kernels (dense / unfused trunk, two-kernel tail, CUDA-core engine) and the validation
// taps need -- pool1, pool2, pool3, upconv3's output: 1.5 MB per ship -- live in a second allocation made on first use, so that
// the default path can be created for 131 072 ships (24 GB) instead of 220 GB.
static int ensure_wide_workspace(ofb_policy *p) {
    if (p->wide_blob) return OFB_OK;
    const size_t C = (size_t)p->max_ships;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t o_p1 = take(C * 200 * 200 * 8 * 2), o_p2 = take(C * 100 * 100 * 8 * 2), o_p3 = take(C * 50 * 50 * 8 * 2);
    const size_t o_u3 = take(C * POL_UP3_ITEM * 2);
    const cudaError_t e = cudaMalloc(&p->wide_blob, off);
    if (e != cudaSuccess) {
        p->wide_blob = nullptr;
        ofb_set_error("ofb_policy: cudaMalloc(%zu bytes of intermediates for %d ships: alternative kernels / taps) failed: %s", off, p->max_ships,
                      cudaGetErrorString(e));
        return OFB_E_NOMEM;
    }
    char *wb = static_cast<char *>(p->wide_blob);
    p->ws.pool1 = reinterpret_cast<__nv_bfloat16 *>(wb + o_p1);
    p->ws.pool2 = reinterpret_cast<__nv_bfloat16 *>(wb + o_p2);
    p->ws.pool3 = reinterpret_cast<__nv_bfloat16 *>(wb + o_p3);
    p->ws.up3 = reinterpret_cast<__nv_bfloat16 *>(wb + o_u3);
    return OFB_OK;
}

static int forward_chunk(ofb_policy *p, const uint32_t *maps, const float *vec, int A, int P, float *act, float *ptr, int32_t *iact,
                         int32_t *xy, cudaStream_t st) {
    const PolicyDev &w = p->w;
    PolicyWork &ws = p->ws;
    const int S = A * P;
    const bool tc = p->engine == OFB_ENGINE_TENSOR;
    int rc;
    if (!tc || p->dense_trunk || p->unfused_trunk || p->unfused_tail || p->taps)
        if ((rc = ensure_wide_workspace(p)) != OFB_OK) return rc;
    if (tc && !p->dense_trunk && !p->unfused_trunk) {
        // the whole trunk (conv1 .. conv4 + pools) in one sparse kernel: pool2 / pool3 never reach HBM (unless taps are on)
        ProfScope ps(p, L_TRUNK12, st);
        if ((rc = pol_st_trunk(p, maps, ws.flat, p->taps ? ws.pool2 : nullptr, p->taps ? ws.pool3 : nullptr, A, st)) != OFB_OK) return rc;
    } else if (tc) {
        { ProfScope ps(p, L_TRUNK12, st);
          rc = p->dense_trunk == 1 ? pol_tz_trunk12(p, maps, ws.pool2, A, st)
                                   : (p->dense_trunk == 2 ? pol_sp_trunk12(p, maps, ws.pool2, A, st) : pol_st_trunk12(p, maps, ws.pool2, A, st));
          if (rc != OFB_OK) return rc; }
        { ProfScope ps(p, L_CONV3, st); if ((rc = pol_tc_conv_pool(p, 1, ws.pool2, ws.pool3, 100, A, 50 * 50 * 8, st)) != OFB_OK) return rc; }
        { ProfScope ps(p, L_CONV4, st); if ((rc = pol_tc_conv_pool(p, 2, ws.pool3, ws.flat, 50, A, POL_FLAT_PITCH, st)) != OFB_OK) return rc; }
    } else {
        { ProfScope ps(p, L_TRUNK12, st);
          k_trunk1_cc<<<dim3((40000 + 255) / 256, A), 256, 0, st>>>(maps, w, ws.pool1);
          k_conv_pool_cc<<<dim3((10000 + 127) / 128, A), 128, 0, st>>>(ws.pool1, w.cw[0], w.cb[0], ws.pool2, 200, 100 * 100 * 8); }
        { ProfScope ps(p, L_CONV3, st);
          k_conv_pool_cc<<<dim3((2500 + 127) / 128, A), 128, 0, st>>>(ws.pool2, w.cw[1], w.cb[1], ws.pool3, 100, 50 * 50 * 8); }
        { ProfScope ps(p, L_CONV4, st);
          k_conv_pool_cc<<<dim3((625 + 127) / 128, A), 128, 0, st>>>(ws.pool3, w.cw[2], w.cb[2], ws.flat, 50, POL_FLAT_PITCH); }
    }
    { ProfScope ps(p, L_DENSE1, st);
      if (tc) { if ((rc = pol_tc_dense1(p, ws.flat, ws.hflat, A, st)) != OFB_OK) return rc; }
      else k_dense1_cc<<<(A + D1_ARENAS - 1) / D1_ARENAS, 256, 0, st>>>(ws.flat, w.d1_wf, ws.hflat, A); }
    { ProfScope ps(p, L_HEADS, st);
      k_heads<<<(S + HEADS_G - 1) / HEADS_G, 256, HEADS_SMEM_FLOATS * sizeof(float), st>>>(ws.hflat, vec, w, P, S, act, iact, ws.up2,
                                                                 tc ? (p->unfused_tail ? 1 : 2) : 0); }
    if (!xy && !ptr) { OFB_CUDA_CHECK(cudaGetLastError()); return OFB_OK; }
    int parts;
    if (tc && !p->unfused_tail) {
        // fused tail: upconv3 never reaches HBM, (x, y) comes straight out of the kernel
        ProfScope ps(p, L_TAIL, st);
        if ((rc = pol_tz_tail(p, ws.up2, ptr, xy, p->taps ? ws.up3 : nullptr, S, st)) != OFB_OK) return rc;
        OFB_CUDA_CHECK(cudaGetLastError());
        return OFB_OK;
    }
    if (tc) {
        { ProfScope ps(p, L_UP3, st); if ((rc = pol_tz_up3(p, ws.up2, ws.up3, S, st)) != OFB_OK) return rc; }
        { ProfScope ps(p, L_UP4, st); if ((rc = pol_tz_up4(p, ws.up3, ptr, ws.amax_val, ws.amax_idx, S, st)) != OFB_OK) return rc; }
        parts = pol_tz_up4_parts();
    } else {
        { ProfScope ps(p, L_UP3, st); k_up3_cc<<<dim3((10000 + 127) / 128, S), 128, 0, st>>>(ws.up2, w, ws.up3); }
        parts = (40000 + 255) / 256;
        { ProfScope ps(p, L_UP4, st); k_up4_cc<<<dim3(parts, S), 256, 0, st>>>(ws.up3, w, ptr, ws.amax_val, ws.amax_idx); }
    }
    if (xy) { ProfScope ps(p, L_ARGMAX, st); k_argmax_final<<<(S * 32 + 127) / 128, 128, 0, st>>>(ws.amax_val, ws.amax_idx, parts, S, xy); }
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_policy_forward(ofb_policy *p, const uint32_t *maps, const float *vec, int64_t n_arenas, int P, float *act,
                                  float *ptr, int32_t *iact, int32_t *xy, void *stream) {
    if (!p || !maps || !vec || n_arenas < 0 || P < 1 || P > p->max_ships) {
        ofb_set_error("ofb_policy_forward: bad argument");
        return OFB_E_ARG;
    }
    OFB_CUDA_CHECK(cudaSetDevice(p->device));
    static thread_local SmemAttrCache attr = {};
    OFB_CUDA_CHECK(attr.ensure(k_heads, (int)(HEADS_SMEM_FLOATS * sizeof(float))));
    const int64_t chunk = p->max_ships / P;                     // arenas per chunk
    for (int64_t a0 = 0; a0 < n_arenas; a0 += chunk) {
        const int A = (int)((n_arenas - a0) < chunk ? (n_arenas - a0) : chunk);
        const int64_t s0 = a0 * P;
        int rc = forward_chunk(p, maps + a0 * 2 * POL_WORDS, vec + s0 * 8, A, P, act ? act + s0 * 2 : nullptr,
                               ptr ? ptr + s0 * 160000 : nullptr, iact ? iact + s0 : nullptr, xy ? xy + s0 * 2 : nullptr,
                               (cudaStream_t)stream);
        if (rc != OFB_OK) return rc;
    }
    return OFB_OK;
}

// ------------------------------------------------------------------------------------------------
// action vector of QlearnIA.play (+ eps-greedy random_play), image packing, debug taps
// ------------------------------------------------------------------------------------------------
// eps(t) of include/ofb_policy.h (double precision: t reaches 10^5..10^6 and the cosine's argument must not lose it)
__host__ __device__ inline float eps_at(const ofb_eps_schedule &s, double t) {
    if (s.kind == OFB_EPS_COSINE) {
        const double ph = fmod(t, s.period) / s.period;
        return (float)(s.start * (cos(ph * 6.283185307179586) + 1.0) * 0.5);
    }
    if (s.kind == OFB_EPS_DECAY) {
        double n = 0.0;
        if (s.start > s.floor && s.decay > 0.0 && s.decay < 1.0) n = ceil(log(s.floor / s.start) / log(s.decay));
        return (float)(s.start * pow(s.decay, fmin(t, n)));
    }
    return (float)s.start;
}

// WRITEBACK: the played action replaces the prediction in iact / xy (what QlearnIA.play remembers, :399-401)
template <bool WRITEBACK>
__global__ void k_write_actions(int *__restrict__ iact, int *__restrict__ xy, long long n_rows, int P,
                                const int *__restrict__ ship_index, int S, ofb_eps_schedule sched, double t, int force_random,
                                uint64_t seed, long long arena0, uint32_t step, int2 *__restrict__ actions, float *__restrict__ eps_out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const float eps = force_random ? 1.0f : eps_at(sched, t);
    if (r == 0 && eps_out) *eps_out = eps;
    if (r >= n_rows) return;
    const long long a = r / P;
    const int ship = ship_index[r % P];
    int ia = 0, x = 0, y = 0;
    if (!force_random) { ia = iact[r]; x = xy[r * 2]; y = xy[r * 2 + 1]; }
    if (eps > 0.f) {
        uint32_t c[4] = {(uint32_t)(arena0 + a), (uint32_t)ship, step, 2u};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        if (force_random || (float)(c[0] >> 8) * (1.0f / 16777216.0f) <= eps) {     // np.random.rand() <= epsilon  (:201)
            ia = (int)(c[1] & 1u);                                   // random.randint(0, action_size - 1)
            x = (int)mulhi32(c[2], POL_W);                           // randint(0, DEFAULT_WIDTH - 1)
            y = (int)mulhi32(c[3], POL_W);
            if (WRITEBACK) { iact[r] = ia; xy[r * 2] = x; xy[r * 2 + 1] = y; }
        }
    }
    const int shoot = ia == 0, thrust = ia == 1;                     // act_vector[iaction] = 1   (:449)
    actions[a * S + ship] = make_int2(shoot | (thrust << 16), (x & 0xffff) | (y << 16));
}

extern "C" int ofb_policy_write_actions(const int32_t *iact, const int32_t *xy, int64_t n_arenas, int P, const int32_t *ship_index,
                                        int S, float eps, uint64_t seed, int64_t arena0, uint32_t step, int16_t *actions,
                                        void *stream) {
    if (!iact || !xy || !ship_index || !actions || P < 1 || S < P) { ofb_set_error("ofb_policy_write_actions: bad argument"); return OFB_E_ARG; }
    const long long n = n_arenas * P;
    if (n == 0) return OFB_OK;
    ofb_eps_schedule sc;
    memset(&sc, 0, sizeof(sc));
    sc.kind = OFB_EPS_CONST; sc.start = (double)eps;
    k_write_actions<false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        const_cast<int *>(iact), const_cast<int *>(xy), n, P, ship_index, S, sc, 0.0, 0, seed, arena0, step, reinterpret_cast<int2 *>(actions), nullptr);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

extern "C" int ofb_policy_play_actions(int32_t *iact, int32_t *xy, int64_t n_arenas, int P, const int32_t *ship_index, int S,
                                       const ofb_eps_schedule *sched, double t, int force_random, uint64_t seed, int64_t arena0,
                                       uint32_t step, int16_t *actions, float *eps_out, void *stream) {
    if (!iact || !xy || !ship_index || !actions || !sched || P < 1 || S < P || t < 0.0 || sched->kind < OFB_EPS_CONST ||
        sched->kind > OFB_EPS_DECAY || (sched->kind == OFB_EPS_COSINE && !(sched->period > 0.0))) {
        ofb_set_error("ofb_policy_play_actions: bad argument");
        return OFB_E_ARG;
    }
    const long long n = n_arenas * P;
    if (n == 0) return OFB_OK;
    k_write_actions<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        iact, xy, n, P, ship_index, S, *sched, t, force_random, seed, arena0, step, reinterpret_cast<int2 *>(actions), eps_out);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// host-side evaluation of the same closed form (tests / logging use it through ofb_eps_value)
extern "C" float ofb_eps_value(const ofb_eps_schedule *sched, double t) { return sched ? eps_at(*sched, t) : 0.f; }

template <class T>
__global__ void k_pack_image(const T *__restrict__ img, long long n_words, uint32_t *__restrict__ maps) {
    const long long wi = (long long)blockIdx.x * blockDim.x + threadIdx.x;       // word index over [B][5000]
    if (wi >= n_words) return;
    const long long b = wi / POL_WORDS;
    const int w = (int)(wi % POL_WORDS);
    const T *src = img + ((size_t)b * 160000 + (size_t)w * 32) * 2;
    uint32_t s = 0, l = 0;
    for (int k = 0; k < 32; k++) {
        if (src[k * 2] != T(0)) s |= 1u << k;
        if (src[k * 2 + 1] != T(0)) l |= 1u << k;
    }
    maps[(size_t)b * 2 * POL_WORDS + w] = s;
    maps[(size_t)b * 2 * POL_WORDS + POL_WORDS + w] = l;
}

extern "C" int ofb_policy_pack_image(const void *img, int fmt, int64_t n, uint32_t *maps, void *stream) {
    if (!img || !maps || n < 0) { ofb_set_error("ofb_policy_pack_image: bad argument"); return OFB_E_ARG; }
    const long long nw = n * POL_WORDS;
    if (nw == 0) return OFB_OK;
    const unsigned blocks = (unsigned)((nw + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (fmt == OFB_MAP_U8) k_pack_image<uint8_t><<<blocks, 256, 0, st>>>(static_cast<const uint8_t *>(img), nw, maps);
    else if (fmt == OFB_MAP_BF16) k_pack_image<uint16_t><<<blocks, 256, 0, st>>>(static_cast<const uint16_t *>(img), nw, maps);
    else if (fmt == 3) k_pack_image<float><<<blocks, 256, 0, st>>>(static_cast<const float *>(img), nw, maps);
    else { ofb_set_error("ofb_policy_pack_image: unknown format %d", fmt); return OFB_E_ARG; }
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

// plane layout [item][8][200][26][8] (ofb_policy_tz.cu) -> NHWC [item][200][200][8]
__global__ void k_plane200_to_nhwc(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n_pixels) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pixels) return;
    const long long item = t / 40000;
    const int p = (int)(t % 40000), Y = p / 200, X = p % 200;
    *reinterpret_cast<uint4 *>(dst + t * 8) = *reinterpret_cast<const uint4 *>(src + item * POL_UP3_ITEM + pol_plane200_off(Y, X));
}

__global__ void k_plane100_to_nhwc(const __nv_bfloat16 *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n_pixels) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pixels) return;
    const long long item = t / 10000;
    const int p = (int)(t % 10000), y = p / 100, x = p % 100;
    *reinterpret_cast<uint4 *>(dst + t * 8) = *reinterpret_cast<const uint4 *>(src + item * POL_UP2_ITEM + pol_plane100_off(y, x));
}

// pairs layout (see up2_pairs_store) -> NHWC [item][100][100][8] (channels 4..7 = 0)
__global__ void k_pairs_to_nhwc(const uint8_t *__restrict__ src, __nv_bfloat16 *__restrict__ dst, long long n_pixels) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pixels) return;
    const long long item = t / 10000;
    const int p = (int)(t % 10000), y = p / 100, x = p % 100;
    const int s = (x + 1) >> 2, rr = (x + 1) & 3;
    const uint2 v = *reinterpret_cast<const uint2 *>(src + item * TL_UP2_ITEM_BYTES +
                                                     ((size_t)(rr >> 1) * TL_UP2_PLANE + (size_t)(y + 1) * TL_P + s) * 16 + (rr & 1) * 8);
    *reinterpret_cast<uint4 *>(dst + t * 8) = make_uint4(v.x, v.y, 0u, 0u);
}

extern "C" int ofb_policy_debug_tap(ofb_policy *p, int which, int64_t n_items, void *dst_dev, void *stream) {
    if (!p || !dst_dev || n_items < 0 || n_items > p->max_ships) { ofb_set_error("ofb_policy_debug_tap: bad argument"); return OFB_E_ARG; }
    const bool fused = p->engine == OFB_ENGINE_TENSOR && !p->unfused_tail;
    if ((which == 0 || which == 1 || which == 2 || which == 6) && !p->wide_blob) {
        ofb_set_error("ofb_policy_debug_tap: tap %d was not produced -- the default path keeps it on chip; enable taps (ofb_policy_set_taps) or "
                      "select the alternative kernels before the forward", which);
        return OFB_E_STATE;
    }
    if (which == 5 && fused) {                                   // the fused tail takes upconv2's output in the pairs layout
        const long long np = n_items * 10000;
        if (np) k_pairs_to_nhwc<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint8_t *>(p->ws.up2), static_cast<__nv_bfloat16 *>(dst_dev), np);
        OFB_CUDA_CHECK(cudaGetLastError());
        return OFB_OK;
    }
    if (which == 6 && fused) {                                   // ... and writes upconv3's output (NHWC) only when taps are enabled
        if (!p->taps) { ofb_set_error("ofb_policy_debug_tap: enable taps (ofb_policy_set_taps) before the forward to see upconv3's output"); return OFB_E_STATE; }
        OFB_CUDA_CHECK(cudaMemcpyAsync(dst_dev, p->ws.up3, (size_t)n_items * 320000 * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return OFB_OK;
    }
    if (which == 5 && p->engine == OFB_ENGINE_TENSOR) {          // ... and upconv2's in the 4-plane layout
        const long long np = n_items * 10000;
        if (np) k_plane100_to_nhwc<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            p->ws.up2, static_cast<__nv_bfloat16 *>(dst_dev), np);
        OFB_CUDA_CHECK(cudaGetLastError());
        return OFB_OK;
    }
    if (which == 6 && p->engine == OFB_ENGINE_TENSOR) {          // the tensor engine keeps upconv3's output in plane layout
        const long long np = n_items * 40000;
        if (np) k_plane200_to_nhwc<<<(unsigned)((np + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            p->ws.up3, static_cast<__nv_bfloat16 *>(dst_dev), np);
        OFB_CUDA_CHECK(cudaGetLastError());
        return OFB_OK;
    }
    const void *src;
    size_t stride, pitch = 0;                                   // pitch: distance between items in the workspace when it is not `stride`
    switch (which) {
    case 0: src = p->ws.pool1; stride = 200 * 200 * 8 * 2; break;
    case 1: src = p->ws.pool2; stride = 100 * 100 * 8 * 2; break;
    case 2: src = p->ws.pool3; stride = 50 * 50 * 8 * 2; break;
    case 3: src = p->ws.flat; stride = POL_FLAT_PITCH * 2; break;
    case 4: src = p->ws.hflat; stride = 100 * 4; break;
    case 5: src = p->ws.up2; stride = 100 * 100 * 8 * 2; pitch = POL_UP2_ITEM * 2; break;
    case 6: src = p->ws.up3; stride = 200 * 200 * 8 * 2; pitch = POL_UP3_ITEM * 2; break;
    default: ofb_set_error("ofb_policy_debug_tap: unknown tap %d", which); return OFB_E_ARG;
    }
    if (n_items == 0) return OFB_OK;
    OFB_CUDA_CHECK(cudaMemcpy2DAsync(dst_dev, stride, src, pitch ? pitch : stride, stride, (size_t)n_items, cudaMemcpyDeviceToDevice,
                                     (cudaStream_t)stream));
    return OFB_OK;
}

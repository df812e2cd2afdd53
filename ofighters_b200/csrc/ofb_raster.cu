// K2 -- observation raster for sm_100a (compiled with -fmad=false).
//
// Restates Observation.analyse_battleground (lib/observation.py:79-95): ship_map gets a radius-8
// disk per playable ship, laser_map a radius-2 disk per laser in the list (just-destroyed and
// off-map ones included), both indexed [row = y, col = x]; the disk is skimage.draw.disk as
// called by Circle.binary_draw (lib/form.py:222-228):
//     ul = ceil(c - R), lr = floor(c + R) clipped to the map, sc = c - ul,
//     pixel (i, j) of the box set iff ((i - sc_r)/R)^2 + ((j - sc_c)/R)^2 < 1   (fp64, strict).
// R is 8 or 2, so the divisions are exact scalings and the test is dr*dr + dc*dc < R*R with
// dr = fl(i - fl(c - ul)).  Ship centres are integers -> the same test in int32.
//
// One CTA per arena.  Both bitmaps (W*H/8 bytes each) are composed in shared memory with
// atomicOr, then leave the SM either as one TMA bulk store (OFB_MAP_BITS, 40 000 B / arena) or
// expanded to the dense NHWC bf16 / u8 tensor Keras' predict() takes (qlearnIA_V2.py:208).
#include <cuda_bf16.h>
#include "ofb_common.cuh"
#include "ofb_raster_dev.cuh"

__global__ void __launch_bounds__(256)
k_raster(const char *__restrict__ state, const ArenaLayout lay, void *__restrict__ out, int format,
         long long n_arenas) {
    extern __shared__ __align__(128) uint32_t bits[];      // [2][words]
    const int W = lay.W, H = lay.H;
    const int words = (W * H) >> 5;
    const long long a = blockIdx.x;
    const char *base = state + a * (long long)lay.stride;
    const int *hdr = reinterpret_cast<const int *>(base);
    const uint4 *ship = reinterpret_cast<const uint4 *>(base + lay.off_ship);

    const int n = hdr[HDR_NLASERS];
    {
        uint4 *b4 = reinterpret_cast<uint4 *>(bits);
        for (int i = threadIdx.x; i < (2 * words) / 4; i += blockDim.x) b4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // ships: one thread per (ship, row)
    const int rows = 2 * OFB_R_SHIP - 1;
    for (int t = threadIdx.x; t < lay.S * rows; t += blockDim.x)
        raster_ship_row(bits, W, H, ship[t / rows].x, t % rows - (OFB_R_SHIP - 1));
    // lasers: one thread per laser
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double *lp = reinterpret_cast<const double *>(base + laser_off(lay.off_laser, k));
        raster_laser(bits + words, W, H, lp[0], lp[OFB_G_Y / 8]);
    }

    if (format == OFB_MAP_BITS) {
        // generic-proxy writes -> visible to the async proxy, then one TMA bulk store per arena
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)(2 * words * 4);
            char *dst = reinterpret_cast<char *>(out) + a * (long long)bytes;
            const unsigned src = (unsigned)__cvta_generic_to_shared(bits);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"(src), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    const uint32_t *sm = bits, *lm = bits + words;
    if (format == OFB_MAP_BF16) {
        // 4 pixels x 2 channels x bf16 = 16 B per store; bf16(1.0) = 0x3F80
        uint4 *o = reinterpret_cast<uint4 *>(out) + a * (long long)(W * H / 4);
        for (int q = threadIdx.x; q < W * H / 4; q += blockDim.x) {
            const unsigned p = (unsigned)q * 4u;
            const uint32_t s4 = (sm[p >> 5] >> (p & 31u)) & 0xfu, l4 = (lm[p >> 5] >> (p & 31u)) & 0xfu;
            uint4 v;
            v.x = ((s4 & 1u) ? 0x3F80u : 0u) | ((l4 & 1u) ? 0x3F800000u : 0u);
            v.y = ((s4 & 2u) ? 0x3F80u : 0u) | ((l4 & 2u) ? 0x3F800000u : 0u);
            v.z = ((s4 & 4u) ? 0x3F80u : 0u) | ((l4 & 4u) ? 0x3F800000u : 0u);
            v.w = ((s4 & 8u) ? 0x3F80u : 0u) | ((l4 & 8u) ? 0x3F800000u : 0u);
            __stcs(&o[q], v);
        }
    } else {
        // 8 pixels x 2 channels x u8 = 16 B per store
        uint4 *o = reinterpret_cast<uint4 *>(out) + a * (long long)(W * H / 8);
        for (int q = threadIdx.x; q < W * H / 8; q += blockDim.x) {
            const unsigned p = (unsigned)q * 8u;
            const uint32_t s8 = (sm[p >> 5] >> (p & 31u)) & 0xffu, l8 = (lm[p >> 5] >> (p & 31u)) & 0xffu;
            uint32_t r[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t sa = (s8 >> (2 * j)) & 1u, la = (l8 >> (2 * j)) & 1u;
                const uint32_t sb = (s8 >> (2 * j + 1)) & 1u, lb = (l8 >> (2 * j + 1)) & 1u;
                r[j] = sa | (la << 8) | (sb << 16) | (lb << 24);
            }
            __stcs(&o[q], make_uint4(r[0], r[1], r[2], r[3]));
        }
    }
}

int ofb_raster_bits_launch(const ofb_arenas *h, void *out_dev, cudaStream_t st) {
    return ofb_raster(h, out_dev, OFB_MAP_BITS, (void *)st);
}

extern "C" int ofb_raster(const ofb_arenas *h, void *out_dev, int format, void *stream) {
    if (!h || !out_dev || format < OFB_MAP_BITS || format > OFB_MAP_U8) {
        ofb_set_error("ofb_raster: bad argument");
        return OFB_E_ARG;
    }
    const size_t smem = (size_t)(h->lay.W * h->lay.H / 32) * 2 * sizeof(uint32_t);
    OFB_CUDA_CHECK(cudaSetDevice(h->device));
    static thread_local SmemAttrCache attr = {};
    if (smem > 48 * 1024) OFB_CUDA_CHECK(attr.ensure(k_raster, (int)smem));
    k_raster<<<(unsigned)h->n_arenas, 256, smem, (cudaStream_t)stream>>>(h->state, h->lay, out_dev, format, h->n_arenas);
    OFB_CUDA_CHECK(cudaGetLastError());
    return OFB_OK;
}

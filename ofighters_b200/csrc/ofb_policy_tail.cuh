// Geometry of the fused pointer-head tail (ofb_policy_tail.cu) shared with the host-side operand packer (ofb_policy.cu).
#pragma once
#define TL_P 26                         // M rows per image row: 25 blocks of 4 input pixels + 1 dummy (halo slot)
#define TL_TILES 21                     // 128-row tiles per ship: 21 * 128 = 2688 >= 100 * 26
#define TL_RING 448                     // ring rows per plane: 64 mirror rows + 3 tiles
#define TL_MARGIN 64
#define TL_UP4_LAG 32                   // upconv4's tile t covers the M rows [128 t - 32, 128 t + 96)
#define TL_A3_SLOTS 184                 // staged input entries per plane and tile: 128 + 2 * 26 + 1, rounded to 8
#define TL_UP2_PLANE 2752               // entries (16 B) per plane of upconv2's output in the pairs layout: 102 * 26 + tail
#define TL_UP2_ITEM_BYTES (3 * TL_UP2_PLANE * 16)
// the weight blob (bytes): upconv3 B operand [3 u][2 ks][2 chunks][128 n][16 B]; its top / bottom variants
// [2 sets][2 u][2 ks][2 chunks][64 n][16 B]; upconv4 B operand dy 0: [5 ks][2][48 n], dy 1, 2: [5][2][80], dy 3: [5][2][48];
// its top / bottom variants [2 sets][2 dy][5 ks][2][16 n]; upconv3's bias operand; floats: upconv3 bias [8], corner weights [4][2][2][8]
#define TL_OFF_B3 0
#define TL_OFF_B3V 24576
#define TL_OFF_B4 (TL_OFF_B3V + 16384)
#define TL_OFF_B4V (TL_OFF_B4 + 40960)
#define TL_OFF_BIAS3 (TL_OFF_B4V + 10240) /* upconv3's bias as a B operand [2 chunks][128 n][16 B]: K lane 0 = bias, the rest 0 */
#define TL_OFF_AUX (TL_OFF_BIAS3 + 4096)
#define TL_AUX_FLOATS (8 + 128)
#define TL_WBYTES (TL_OFF_AUX + TL_AUX_FLOATS * 4)
// CTA pairs (cta_group::2): each CTA of a pair holds HALF of every B operand's N columns -- rank r the columns [r N/2, (r + 1) N/2) --
// in the same block order, so every operand offset above is halved; the floats follow at TL_OFF_AUX / 2.  Blob: [2 ranks][TL_WBYTES_H]
#define TL_WBYTES_H (TL_OFF_AUX / 2 + TL_AUX_FLOATS * 4)

// Layer table of the packed weight blob.  Order and shapes mirror XFeatModel.__init__ (model.py:33-111).
// Each layer stores  W[tap][cin][cout] (tap = ky*KS+kx, cout fastest)  followed by  bias[cout],
// with eval-mode BatchNorm(affine=False, eps=1e-5) folded in:  W' = W/sqrt(var+eps),  b' = (b-mean)/sqrt(var+eps).
// accelerated_features_b200/weights.py writes exactly this layout.
#pragma once
#include <stddef.h>

namespace xf {

enum LayerId {
  L_B1_0 = 0, L_B1_1, L_B1_2, L_B1_3, L_SKIP1,
  L_B2_0, L_B2_1,
  L_B3_0, L_B3_1, L_B3_2,
  L_B4_0, L_B4_1, L_B4_2,
  L_B5_0, L_B5_1, L_B5_2, L_B5_3,
  L_FU_0, L_FU_1, L_FU_2,
  L_HH_0, L_HH_1, L_HH_2,
  L_KH_0, L_KH_1, L_KH_2, L_KH_3,
  L_FM_0, L_FM_1, L_FM_2, L_FM_3, L_FM_4,
  L_COUNT
};

struct LayerSpec {
  int cin, cout, ks, stride, relu;
};

// clang-format off
static const LayerSpec kLayers[L_COUNT] = {
  {  1,   4, 3, 1, 1}, {  4,   8, 3, 2, 1}, {  8,   8, 3, 1, 1}, {  8,  24, 3, 2, 1}, {  1,  24, 1, 1, 0},
  { 24,  24, 3, 1, 1}, { 24,  24, 3, 1, 1},
  { 24,  64, 3, 2, 1}, { 64,  64, 3, 1, 1}, { 64,  64, 1, 1, 1},
  { 64,  64, 3, 2, 1}, { 64,  64, 3, 1, 1}, { 64,  64, 3, 1, 1},
  { 64, 128, 3, 2, 1}, {128, 128, 3, 1, 1}, {128, 128, 3, 1, 1}, {128,  64, 1, 1, 1},
  { 64,  64, 3, 1, 1}, { 64,  64, 3, 1, 1}, { 64,  64, 1, 1, 0},
  { 64,  64, 1, 1, 1}, { 64,  64, 1, 1, 1}, { 64,   1, 1, 1, 0},
  { 64,  64, 1, 1, 1}, { 64,  64, 1, 1, 1}, { 64,  64, 1, 1, 1}, { 64,  65, 1, 1, 0},
  {128, 512, 1, 1, 1}, {512, 512, 1, 1, 1}, {512, 512, 1, 1, 1}, {512, 512, 1, 1, 1}, {512,  64, 1, 1, 0},
};
// clang-format on

struct LayerTable {
  size_t w_off[L_COUNT];
  size_t b_off[L_COUNT];
  size_t total;
};

static inline LayerTable make_layer_table() {
  LayerTable t;
  size_t off = 0;
  for (int i = 0; i < L_COUNT; ++i) {
    const LayerSpec& s = kLayers[i];
    t.w_off[i] = off;
    off += (size_t)s.ks * s.ks * s.cin * s.cout;
    off = (off + 3) & ~(size_t)3;  // keep every array 16-byte aligned
    t.b_off[i] = off;
    off += (size_t)s.cout;
    off = (off + 3) & ~(size_t)3;
  }
  t.total = off;
  return t;
}

}  // namespace xf

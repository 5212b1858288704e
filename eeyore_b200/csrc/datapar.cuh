// Shared constants of the data-parallel path (BASELINE config 5, MLP 16-64-64-1): flat theta layout of
// eeyore/models/model.py:44-55 (W0 [64,16], b0 [64], W1 [64,64], b1 [64], W2 [1,64], b2 [1]).
#pragma once
namespace eb {
constexpr int DP_D0 = 16, DP_H = 64;
constexpr int DP_R = 128;            // rows per tile
constexpr int DP_THREADS = 256;
constexpr int DP_OFF_B0 = DP_H * DP_D0;                 // 1024
constexpr int DP_OFF_W1 = (DP_D0 + 1) * DP_H;           // 1088
constexpr int DP_OFF_B1 = DP_OFF_W1 + DP_H * DP_H;      // 5184
constexpr int DP_OFF_W2 = DP_OFF_B1 + DP_H;             // 5248
constexpr int DP_OFF_B2 = DP_OFF_W2 + DP_H;             // 5312
constexpr int DP_P = DP_OFF_B2 + 1;                     // 5313
}  // namespace eb

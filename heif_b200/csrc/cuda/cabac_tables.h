// Host-side construction of dev::CabacTabs (uploaded once per context, copied to shared memory by the
// CABAC kernel).  Table 9-45 / 9-46 values are the standard's (the reference carries the same numbers at
// src/cabac/arithmetic.rs:177-255); initType-0 init values as in src/cabac/syntax_element.rs:90-242
// restricted to the Main / Main Still Picture elements.
#pragma once
#include <cstring>

#include "cabac_parse.cuh"

namespace heic {
namespace dev {

inline void build_cabac_tabs(CabacTabs& t) {
  static const uint8_t kRangeTabLps[64][4] = {
      {128, 176, 208, 240}, {128, 167, 197, 227}, {128, 158, 187, 216}, {123, 150, 178, 205},
      {116, 142, 169, 195}, {111, 135, 160, 185}, {105, 128, 152, 175}, {100, 122, 144, 166},
      {95, 116, 137, 158},  {90, 110, 130, 150},  {85, 104, 123, 142},  {81, 99, 117, 135},
      {77, 94, 111, 128},   {73, 89, 105, 122},   {69, 85, 100, 116},   {66, 80, 95, 110},
      {62, 76, 90, 104},    {59, 72, 86, 99},     {56, 69, 81, 94},     {53, 65, 77, 89},
      {51, 62, 73, 85},     {48, 59, 69, 80},     {46, 56, 66, 76},     {43, 53, 63, 72},
      {41, 50, 59, 69},     {39, 48, 56, 65},     {37, 45, 54, 62},     {35, 43, 51, 59},
      {33, 41, 48, 56},     {32, 39, 46, 53},     {30, 37, 43, 50},     {29, 35, 41, 48},
      {27, 33, 39, 45},     {26, 31, 37, 43},     {24, 30, 35, 41},     {23, 28, 33, 39},
      {22, 27, 32, 37},     {21, 26, 30, 35},     {20, 24, 29, 33},     {19, 23, 27, 31},
      {18, 22, 26, 30},     {17, 21, 25, 28},     {16, 20, 23, 27},     {15, 19, 22, 25},
      {14, 18, 21, 24},     {14, 17, 20, 23},     {13, 16, 19, 22},     {12, 15, 18, 21},
      {12, 14, 17, 20},     {11, 14, 16, 19},     {11, 13, 15, 18},     {10, 12, 15, 17},
      {10, 12, 14, 16},     {9, 11, 13, 15},      {9, 11, 12, 14},      {8, 10, 12, 14},
      {8, 9, 11, 13},       {7, 9, 11, 12},       {7, 9, 10, 12},       {7, 8, 10, 11},
      {6, 8, 9, 11},        {6, 7, 9, 10},        {6, 7, 8, 9},         {2, 2, 2, 2}};
  static const uint8_t kTransIdxLps[64] = {0,  0,  1,  2,  2,  4,  4,  5,  6,  7,  8,  9,  9,  11, 11, 12,
                                           13, 13, 15, 15, 16, 16, 18, 18, 19, 19, 21, 21, 22, 22, 23, 24,
                                           24, 25, 26, 26, 27, 27, 28, 29, 29, 30, 30, 30, 31, 32, 32, 33,
                                           33, 33, 34, 34, 35, 35, 35, 36, 36, 36, 37, 37, 37, 38, 38, 63};
  static const uint8_t kInitValues[NUM_CTX] = {
      153,                                                                                     // sao_merge
      200,                                                                                     // sao_type_idx
      139, 141, 157,                                                                           // split_cu_flag
      154,                                                                                     // cu_transquant_bypass
      184,                                                                                     // part_mode
      184,                                                                                     // prev_intra_luma_pred
      63,                                                                                      // intra_chroma_pred_mode
      153, 138, 138,                                                                           // split_transform_flag
      111, 141,                                                                                // cbf_luma
      94,  138, 182, 154,                                                                      // cbf_cb / cbf_cr
      154, 154,                                                                                // cu_qp_delta_abs
      139, 139,                                                                                // transform_skip_flag
      110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79,  108, 123, 63,  // last x prefix
      110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79,  108, 123, 63,  // last y prefix
      91,  171, 134, 141,                                                                      // coded_sub_block_flag
      111, 111, 125, 110, 110, 94,  124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
      107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,  // sig_coeff
      140, 92,  137, 138, 140, 152, 138, 139, 153, 74,  149, 92,  139, 107, 122, 152, 140, 179, 166, 182, 140,
      227, 122, 197,                                                                           // greater1
      138, 153, 136, 167, 152, 152};                                                           // greater2
  static const uint8_t kSigMap4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};

  std::memset(&t, 0, sizeof t);
  for (int s = 0; s < 128; s++) {
    int p = s >> 1, mps = s & 1;
    t.st[s].lps = (uint32_t)kRangeTabLps[p][0] | ((uint32_t)kRangeTabLps[p][1] << 8) |
                  ((uint32_t)kRangeTabLps[p][2] << 16) | ((uint32_t)kRangeTabLps[p][3] << 24);
    int p_mps = p < 62 ? p + 1 : p;  // transIdxMps
    int mps_after_lps = p == 0 ? 1 - mps : mps;
    t.st[s].next = (uint32_t)((p_mps << 1) | mps) | ((uint32_t)((kTransIdxLps[p] << 1) | mps_after_lps) << 8);
  }
  // 6.5.3 up-right diagonal scans
  auto diag = [](int lg, uint8_t* fwd, uint8_t* inv) {
    int n = 1 << lg, i = 0, x = 0, y = 0;
    bool stop = false;
    while (!stop) {
      while (y >= 0) {
        if (x < n && y < n) {
          fwd[i] = (uint8_t)(x | (y << lg));
          inv[(y << lg) | x] = (uint8_t)i;
          i++;
        }
        y--;
        x++;
      }
      y = x;
      x = 0;
      if (i >= n * n) stop = true;
    }
  };
  diag(1, t.diag2, t.inv_diag2);
  diag(2, t.diag4, t.inv_diag4);
  diag(3, t.diag8, t.inv_diag8);
  std::memcpy(t.sig_map4, kSigMap4, 16);
  for (int pat = 0; pat < 4; pat++)
    for (int yp = 0; yp < 4; yp++)
      for (int xp = 0; xp < 4; xp++) {
        int c;
        if (pat == 0) c = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
        else if (pat == 1) c = (yp == 0) ? 2 : (yp == 1) ? 1 : 0;
        else if (pat == 2) c = (xp == 0) ? 2 : (xp == 1) ? 1 : 0;
        else c = 2;
        t.sig_pat[pat][(yp << 2) | xp] = (uint8_t)c;
      }
  for (int scan = 0; scan < 3; scan++) {
    auto pos = [&](int k) {  // scan position k -> (xP, yP) of a 4x4 scan (6.5.3-6.5.5)
      int x, y;
      if (scan == 0) x = t.diag4[k] & 3, y = t.diag4[k] >> 2;
      else if (scan == 1) x = k & 3, y = k >> 2;
      else x = k >> 2, y = k & 3;
      return (y << 2) | x;
    };
    uint64_t w = 0;
    for (int k = 0; k < 16; k++) w |= (uint64_t)kSigMap4[pos(k)] << (4 * k);
    t.sig_nib[scan] = w;
    for (int pat = 0; pat < 4; pat++) {
      w = 0;
      for (int k = 0; k < 16; k++) w |= (uint64_t)t.sig_pat[pat][pos(k)] << (4 * k);
      t.sig_nib[3 + scan * 4 + pat] = w;
    }
  }
  std::memcpy(t.init_value, kInitValues, NUM_CTX);
}

}  // namespace dev
}  // namespace heic

/*
 * hevc_oracle.c — plain-C CPU restatement of the HEIC reconstruction path.
 * TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
 *
 * Two kinds of code live here:
 *  (1) stages the reference implements — restated from it, function by function:
 *        CABAC engine init / decision / bypass / terminate   reference src/cabac/arithmetic.rs:23-169
 *        context initialisation (9.3.2.2)                     src/cabac/arithmetic.rs:40-78
 *        initType-0 init values                               src/cabac/syntax_element.rs:90-242
 *        binarisations (FL, TR, EGk, coeff_abs_level_remaining, cu_qp_delta_abs,
 *        intra_chroma_pred_mode, sao_type_idx, last_sig_coeff_prefix, part_mode)
 *                                                             src/cabac/decoder.rs:23-284
 *        CTU loop skeleton                                    src/hevc/slice.rs:206-247
 *  (2) stages the reference ends in todo!() for (slice.rs:249-255) — SAO syntax, coding quadtree,
 *      residual coding, scaling, inverse transforms, intra prediction, deblocking, SAO — restated
 *      from ITU-T H.265 (v1 clauses cited inline).
 *
 * PARITY PINNING: (1) is pinned by the reference's own vectors (Table 9-39/9-41 bin strings,
 * tests/test_oracle_reference_vectors.py).  (2) is NOT pinned by the reference (it has nothing to
 * pin against); it is pinned by FFmpeg's independent HEVC decoder on all 48 tiles of
 * halfmoonbay.heic and on synthetic streams (tests/golden/, oracle/ffmpeg_oracle.py).
 */
#include "hevc_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Tables                                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* Table 9-46 rangeTabLps and Table 9-45 transIdxLps / transIdxMps (arithmetic.rs:177-255). */
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
static const uint8_t kTransIdxMps[64] = {1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16,
                                         17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32,
                                         33, 34, 35, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48,
                                         49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 62, 63};

/* Context layout (one flat array; the reference keys a HashMap by (table, idx)). */
enum {
  CTX_SAO_MERGE = 0,        /* Table 9-5  */
  CTX_SAO_TYPE = 1,         /* Table 9-6  */
  CTX_SPLIT_CU = 2,         /* Table 9-7, 3 */
  CTX_CU_TQ_BYPASS = 5,     /* Table 9-8  */
  CTX_PART_MODE = 6,        /* Table 9-11 */
  CTX_PREV_INTRA = 7,       /* Table 9-12 */
  CTX_CHROMA_PRED = 8,      /* Table 9-13 */
  CTX_SPLIT_TRANSFORM = 9,  /* Table 9-20, 3 */
  CTX_CBF_LUMA = 12,        /* Table 9-21, 2 */
  CTX_CBF_CHROMA = 14,      /* Table 9-22, 4 */
  CTX_CU_QP_DELTA = 18,     /* Table 9-24, 2 */
  CTX_TSKIP = 20,           /* Table 9-25, luma + chroma */
  CTX_LAST_X = 22,          /* Table 9-26, 18 */
  CTX_LAST_Y = 40,          /* Table 9-27, 18 */
  CTX_CSBF = 58,            /* Table 9-28, 4 */
  CTX_SIG = 62,             /* Table 9-29, 42 */
  CTX_GT1 = 104,            /* Table 9-30, 24 */
  CTX_GT2 = 128,            /* Table 9-31, 6 */
  NUM_CTX = 134
};

/* initType 0 (I slices) init values; same numbers as syntax_element.rs:90-242 for the
 * Main-Still-Picture elements (the three suspect RExt slots of SURVEY H14 are not used). */
static const uint8_t kInitValues[NUM_CTX] = {
    153,                                                                  /* sao_merge_*_flag */
    200,                                                                  /* sao_type_idx_* */
    139, 141, 157,                                                        /* split_cu_flag */
    154,                                                                  /* cu_transquant_bypass_flag */
    184,                                                                  /* part_mode */
    184,                                                                  /* prev_intra_luma_pred_flag */
    63,                                                                   /* intra_chroma_pred_mode */
    153, 138, 138,                                                        /* split_transform_flag */
    111, 141,                                                             /* cbf_luma */
    94,  138, 182, 154,                                                   /* cbf_cb / cbf_cr */
    154, 154,                                                             /* cu_qp_delta_abs */
    139, 139,                                                             /* transform_skip_flag luma, chroma */
    110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63, /* last x */
    110, 110, 124, 125, 140, 153, 125, 127, 140, 109, 111, 143, 127, 111, 79, 108, 123, 63, /* last y */
    91,  171, 134, 141,                                                   /* coded_sub_block_flag */
    111, 111, 125, 110, 110, 94,  124, 108, 124, 107, 125, 141, 179, 153, 125, 107, 125, 141, 179, 153, 125,
    107, 125, 141, 179, 153, 125, 140, 139, 182, 182, 152, 136, 152, 136, 153, 136, 139, 111, 136, 139, 111,
    140, 92,  137, 138, 140, 152, 138, 139, 153, 74,  149, 92,  139, 107, 122, 152, 140, 179, 166, 182, 140,
    227, 122, 197,                                                        /* greater1 */
    138, 153, 136, 167, 152, 152};                                        /* greater2 */

static const uint8_t kSigCtxIdxMap4x4[16] = {0, 1, 4, 5, 2, 3, 4, 5, 6, 6, 8, 8, 7, 7, 8, 8};
static const int8_t kIntraPredAngle[35] = {0,  0,  32,  26,  21,  17,  13,  9,   5,   2,   0,   -2, -5, -9, -13, -17, -21, -26,
                                           -32, -26, -21, -17, -13, -9, -5, -2, 0,   2,   5,   9,  13, 17, 21,  26,  32};
static const int16_t kInvAngle[15] = {-4096, -1638, -910, -630, -482, -390, -315, -256,
                                      -315,  -390,  -482, -630, -910, -1638, -4096}; /* modes 11..25 */
static const uint8_t kLevelScale[6] = {40, 45, 51, 57, 64, 72};
static const uint8_t kBetaTable[52] = {0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  0,  6,  7,
                                       8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 20, 22, 24, 26, 28, 30, 32,
                                       34, 36, 38, 40, 42, 44, 46, 48, 50, 52, 54, 56, 58, 60, 62, 64};
static const uint8_t kTcTable[54] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,  0,  0,  0,
                                     1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3,  3,  3,  4,
                                     4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 13, 14, 16, 18, 20, 22, 24};
static const uint8_t kChromaQpTable[14] = {29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37}; /* qPi 30..43 */
/* 8.6.4.2: column 0..32 of a quarter cosine; transMatrix[k][n] = C[(k*(2n+1)) mod 128] with sign folding. */
static const int8_t kDctQuarter[33] = {64, 90, 90, 90, 89, 88, 87, 85, 83, 82, 80, 78, 75, 73, 70, 67, 64,
                                       61, 57, 54, 50, 46, 43, 38, 36, 31, 25, 22, 18, 13, 9,  4,  0};
static const int8_t kDst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

static int8_t g_dct32[32][32];
static uint8_t g_scan[3][4][64][2]; /* [scanIdx][log2-? index: 0=2x2,1=4x4,2=8x8][pos][x,y] */
static int g_tables_ready = 0;

static int dct_coef(int k, int n) {
  int m = (k * (2 * n + 1)) & 127;
  if (m <= 32) return kDctQuarter[m];
  if (m <= 64) return -kDctQuarter[64 - m];
  if (m < 96) return -kDctQuarter[m - 64];
  return kDctQuarter[128 - m];
}

static void build_scan(int scan_idx, int log2, uint8_t out[64][2]) {
  int n = 1 << log2, i = 0;
  if (scan_idx == 0) { /* 6.5.3 up-right diagonal */
    int x = 0, y = 0, stop = 0;
    while (!stop) {
      while (y >= 0) {
        if (x < n && y < n) {
          out[i][0] = (uint8_t)x;
          out[i][1] = (uint8_t)y;
          i++;
        }
        y--;
        x++;
      }
      y = x;
      x = 0;
      if (i >= n * n) stop = 1;
    }
  } else if (scan_idx == 1) { /* 6.5.4 horizontal */
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++, i++) out[i][0] = (uint8_t)x, out[i][1] = (uint8_t)y;
  } else { /* 6.5.5 vertical */
    for (int x = 0; x < n; x++)
      for (int y = 0; y < n; y++, i++) out[i][0] = (uint8_t)x, out[i][1] = (uint8_t)y;
  }
}

static void init_tables(void) {
  if (g_tables_ready) return;
  for (int k = 0; k < 32; k++)
    for (int n = 0; n < 32; n++) g_dct32[k][n] = (int8_t)dct_coef(k, n);
  for (int s = 0; s < 3; s++)
    for (int l = 0; l < 4; l++)
      if (l >= 1) build_scan(s, l, g_scan[s][l]);
      else g_scan[s][0][0][0] = g_scan[s][0][0][1] = 0;
  g_tables_ready = 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Decoder state                                                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint8_t type[3], band_pos[3], eo_class[3];
  int8_t offset[3][4];
} SaoParams;

typedef struct {
  const heic_sps* sps;
  const heic_pps* pps;
  const heic_slice_header* sh;
  hevc_oracle_out* out;
  int parse_only;
  int w, h, wc, hc, chroma;
  int log2_ctb, ctb_size, log2_min_cb, log2_min_tb, log2_max_tb, wctb, hctb, ctb4;
  int w4, h4, w8, h8; /* maps cover whole CTBs */
  int slice_qp, log2_min_cu_qp_delta_size, max_trafo_depth_intra;
  /* CABAC engine (arithmetic.rs:13-20) */
  const uint8_t* data;
  uint32_t rbsp_len_;
  uint32_t byte_pos, bit_pos, end;
  uint32_t range, offset;
  uint8_t ctx[NUM_CTX], ctx_wpp[NUM_CTX]; /* pStateIdx<<1 | valMps */
  uint32_t bins;
  const uint8_t* test_bins; /* unit tests only: bins come from this list instead of the arithmetic decoder */
  int test_n, test_i;
  /* per-picture maps */
  uint8_t* ct_depth; /* per 8x8 */
  uint8_t* ipm;      /* luma IntraPredModeY per 4x4 */
  uint8_t* decoded4; /* reconstructed flag per 4x4 (availability, 6.4.1) */
  uint8_t* qp_map;   /* QpY per 8x8 */
  uint8_t* edge;     /* per 4x4: bit0 = TU edge on the left, bit1 = TU edge on top */
  SaoParams* sao;
  uint8_t* plane[3];
  int stride[3];
  uint8_t scaling[4][3][32 * 32]; /* ScalingFactor[sizeId][cIdx][y*n+x] (intra) */
  /* QP state 8.6.1 */
  int is_cu_qp_delta_coded, cu_qp_delta_val, qp_y, last_qp_y, qp_y_pred, first_qg_in_row, qg_x, qg_y;
  /* current CU */
  int cu_x, cu_y, cu_log2, part_nxn, chroma_mode;
  int err;
} Dec;

#define FAIL(d, code, ...)                                                   \
  do {                                                                       \
    if (!(d)->err) {                                                         \
      (d)->err = (code);                                                     \
      snprintf((d)->out->error, sizeof((d)->out->error), __VA_ARGS__);       \
    }                                                                        \
  } while (0)

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip8(int v) { return clip3(0, 255, v); }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* z-order index of the 4x4 block (x4, y4) inside its CTB */
static inline uint32_t zorder4(int x4, int y4) {
  uint32_t z = 0;
  for (int b = 0; b < 4; b++) z |= (uint32_t)((x4 >> b) & 1) << (2 * b) | (uint32_t)((y4 >> b) & 1) << (2 * b + 1);
  return z;
}
static uint32_t tu_index(const Dec* d, int x, int y) { /* luma pixel position */
  int rx = x >> d->log2_ctb, ry = y >> d->log2_ctb;
  int m = d->ctb4 - 1;
  return (uint32_t)(ry * d->wctb + rx) * (uint32_t)(d->ctb4 * d->ctb4) + zorder4((x >> 2) & m, (y >> 2) & m);
}
static uint32_t coeff_offset(const Dec* d, int x, int y, int c_idx) { /* luma position of the TU origin */
  if (c_idx == 0) return tu_index(d, x, y) * 16u;
  int rx = x >> d->log2_ctb, ry = y >> d->log2_ctb;
  int c4 = d->ctb4 >> 1, m = c4 - 1;
  return ((uint32_t)(ry * d->wctb + rx) * (uint32_t)(c4 * c4) + zorder4((x >> 3) & m, (y >> 3) & m)) * 16u;
}

uint32_t hevc_oracle_tu_map_len(const heic_sps* sps) {
  int log2_ctb = sps->log2_min_luma_coding_block_size_minus3 + 3 + sps->log2_diff_max_min_luma_coding_block_size;
  int cs = 1 << log2_ctb;
  uint32_t n = ((sps->pic_width_in_luma_samples + cs - 1) >> log2_ctb) * ((sps->pic_height_in_luma_samples + cs - 1) >> log2_ctb);
  return n * (uint32_t)((cs / 4) * (cs / 4));
}
uint32_t hevc_oracle_coeff_len(const heic_sps* sps, int c_idx) {
  uint32_t n = hevc_oracle_tu_map_len(sps) * 16u;
  return c_idx == 0 ? n : n / 4;
}

/* ------------------------------------------------------------------------------------------ */
/* CABAC engine — arithmetic.rs:23-169                                                         */
/* ------------------------------------------------------------------------------------------ */

static inline uint32_t read_bit(Dec* d) { /* rbsp_reader.rs:123-136; zeros past the substream end */
  uint32_t bit = 0;
  if (d->byte_pos < d->end) bit = (d->data[d->byte_pos] >> (7 - d->bit_pos)) & 1u;
  if (++d->bit_pos == 8) {
    d->bit_pos = 0;
    d->byte_pos++;
  }
  return bit;
}

/* 9.3.2.5 (arithmetic.rs:23-38) */
static void cabac_init_engine(Dec* d, uint32_t start, uint32_t end) {
  d->byte_pos = start;
  d->bit_pos = 0;
  d->end = end;
  d->range = 510;
  d->offset = 0;
  for (int i = 0; i < 9; i++) d->offset = (d->offset << 1) | read_bit(d);
  if (d->offset == 510 || d->offset == 511) FAIL(d, HEIC_E_BITSTREAM, "Invalid ivlOffset value");
}

/* 9.3.2.2 (arithmetic.rs:40-78) */
void hevc_oracle_context_init(int slice_qp, uint8_t* state) {
  for (int i = 0; i < NUM_CTX; i++) {
    int v = kInitValues[i];
    int slope_idx = v >> 4, offset_idx = v & 15;
    int m = slope_idx * 5 - 45, n = (offset_idx << 3) - 16;
    int pre = clip3(1, 126, ((m * clip3(0, 51, slice_qp)) >> 4) + n);
    int val_mps = pre > 63;
    int p_state = val_mps ? pre - 64 : 63 - pre;
    state[i] = (uint8_t)((p_state << 1) | val_mps);
  }
}

static inline void renorm(Dec* d) { /* arithmetic.rs:137-144 */
  while (d->range < 256) {
    d->range <<= 1;
    d->offset = (d->offset << 1) | read_bit(d);
  }
}

/* 9.3.4.3.2 (arithmetic.rs:97-135) */
static int test_bin(Dec* d) { return d->test_i < d->test_n ? d->test_bins[d->test_i++] : 0; }

static int decode_decision(Dec* d, int ctx_idx) {
  if (d->test_bins) return test_bin(d);
  uint8_t s = d->ctx[ctx_idx];
  uint32_t p = s >> 1, mps = s & 1;
  uint32_t q = (d->range >> 6) & 3;
  uint32_t lps = kRangeTabLps[p][q];
  int bin;
  d->bins++;
  d->range -= lps;
  if (d->offset >= d->range) {
    bin = !mps;
    d->offset -= d->range;
    d->range = lps;
    if (p == 0) mps = 1 - mps;
    p = kTransIdxLps[p];
  } else {
    bin = (int)mps;
    p = kTransIdxMps[p];
  }
  d->ctx[ctx_idx] = (uint8_t)((p << 1) | mps);
  renorm(d);
  return bin;
}

/* 9.3.4.3.4 (arithmetic.rs:146-157) */
static int decode_bypass(Dec* d) {
  if (d->test_bins) return test_bin(d);
  d->bins++;
  d->offset = (d->offset << 1) | read_bit(d);
  if (d->offset >= d->range) {
    d->offset -= d->range;
    return 1;
  }
  return 0;
}

/* 9.3.4.3.5 (arithmetic.rs:159-169) */
static int decode_terminate(Dec* d) {
  d->bins++;
  d->range -= 2;
  if (d->offset >= d->range) return 1;
  renorm(d);
  return 0;
}

/* Binarisations — cabac/decoder.rs:152-284 */
static uint32_t decode_fl_bypass(Dec* d, int n_bits) { /* decoder.rs:152-164 */
  uint32_t v = 0;
  for (int i = 0; i < n_bits; i++) v = (v << 1) | (uint32_t)decode_bypass(d);
  return v;
}
static uint32_t decode_tr_bypass(Dec* d, uint32_t c_max) { /* decoder.rs:166-190 with cRiceParam 0 */
  uint32_t v = 0;
  while (v < c_max && decode_bypass(d)) v++;
  return v;
}
static uint32_t decode_egk_bypass(Dec* d, int k) { /* decoder.rs:206-222, 32-bit (Appendix B #11) */
  int ones = 0;
  while (decode_bypass(d)) {
    if (++ones > 31) {
      FAIL(d, HEIC_E_BITSTREAM, "EGk prefix too long");
      return 0;
    }
  }
  uint32_t suffix = decode_fl_bypass(d, ones + k);
  return (((1u << ones) - 1u) << k) + suffix;
}
/* decoder.rs:224-261: TR prefix (cMax = 4 << rice) + EG(rice+1) escape. */
static uint32_t decode_coeff_abs_level_remaining(Dec* d, int rice) {
  uint32_t c_max = 4u << rice;
  uint32_t prefix = 0;
  while (prefix < 4 && decode_bypass(d)) prefix++;
  if (prefix < 4) return (prefix << rice) + decode_fl_bypass(d, rice);
  return c_max + decode_egk_bypass(d, rice + 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Scaling factors 7.4.5                                                                       */
/* ------------------------------------------------------------------------------------------ */
static void build_scaling_factors(Dec* d) {
  const heic_scaling_list* sl = NULL;
  if (d->sps->scaling_list_enabled_flag) {
    if (d->pps->pps_scaling_list_data_present_flag) sl = &d->pps->scaling_list;
    else if (d->sps->sps_scaling_list_data_present_flag) sl = &d->sps->scaling_list;
  }
  for (int size_id = 0; size_id < 4; size_id++) {
    int n = 4 << size_id;
    for (int c = 0; c < 3; c++) {
      uint8_t* f = d->scaling[size_id][c];
      if (!d->sps->scaling_list_enabled_flag) {
        memset(f, 16, (size_t)n * n);
        continue;
      }
      int matrix_id = (size_id == 3) ? 0 : c; /* 32x32: only luma exists in 4:2:0 */
      uint8_t deflist[64];
      const uint8_t* list;
      int dc = 16;
      if (sl) {
        list = sl->list[size_id][matrix_id];
        if (size_id >= 2) dc = sl->dc[size_id - 2][matrix_id];
      } else { /* Table 7-5 / 7-6 defaults */
        static const uint8_t kIntra8[64] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18,
                                            17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
                                            24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31,
                                            29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115};
        if (size_id == 0) memset(deflist, 16, 64);
        else memcpy(deflist, kIntra8, 64);
        list = deflist;
      }
      if (size_id == 0) {
        for (int i = 0; i < 16; i++) f[g_scan[0][2][i][1] * 4 + g_scan[0][2][i][0]] = list[i];
      } else {
        int rep = n / 8;
        for (int i = 0; i < 64; i++) {
          int x = g_scan[0][3][i][0], y = g_scan[0][3][i][1];
          for (int j = 0; j < rep; j++)
            for (int k = 0; k < rep; k++) f[(y * rep + j) * n + x * rep + k] = list[i];
        }
        if (size_id >= 2) f[0] = (uint8_t)dc;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Inverse transforms 8.6.4.2 (8-bit: first-stage shift 7, second 12)                          */
/* ------------------------------------------------------------------------------------------ */
void hevc_oracle_idct(const int16_t* coeff, int16_t* resid, int log2_size, int dst) {
  init_tables();
  int n = 1 << log2_size, step = 32 >> log2_size;
  int tmp[32 * 32];
  for (int x = 0; x < n; x++) /* columns */
    for (int i = 0; i < n; i++) {
      int s = 0;
      for (int j = 0; j < n; j++) {
        int m = dst ? kDst4[j][i] : g_dct32[j * step][i];
        s += m * coeff[j * n + x];
      }
      tmp[i * n + x] = clip3(-32768, 32767, (s + 64) >> 7);
    }
  for (int y = 0; y < n; y++) /* rows */
    for (int i = 0; i < n; i++) {
      int s = 0;
      for (int j = 0; j < n; j++) {
        int m = dst ? kDst4[j][i] : g_dct32[j * step][i];
        s += m * tmp[y * n + j];
      }
      resid[y * n + i] = (int16_t)clip3(-32768, 32767, (s + 2048) >> 12);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Intra prediction 8.4.4.2                                                                    */
/* ------------------------------------------------------------------------------------------ */
static int avail4(const Dec* d, int xl, int yl) { /* luma position; 6.4.1 via the reconstructed map */
  if (xl < 0 || yl < 0 || xl >= d->w || yl >= d->h) return 0;
  return d->decoded4[(yl >> 2) * d->w4 + (xl >> 2)];
}

static void intra_predict(Dec* d, int x0, int y0, int log2, int c_idx, int mode, uint8_t* pred /* n*n */) {
  int n = 1 << log2;
  int sub = (c_idx && d->chroma) ? 1 : 0; /* 4:2:0 */
  const uint8_t* pl = d->plane[c_idx];
  int st = d->stride[c_idx];
  /* p[-1][-1 .. 2n-1] -> left[0 .. 2n], p[-1 .. 2n-1][-1] -> top[0 .. 2n]; index 0 is the corner */
  int left_raw[65], top_raw[65], fl[65], ft[65];
  uint8_t la[65], ta[65];
  int *left = left_raw, *top = top_raw;
  int any = 0;
  for (int i = 0; i <= 2 * n; i++) {
    int yy = y0 + i - 1, xx = x0 - 1;
    la[i] = (uint8_t)avail4(d, xx << sub, yy << sub);
    if (la[i]) left[i] = pl[yy * st + xx], any = 1;
    xx = x0 + i - 1, yy = y0 - 1;
    ta[i] = (uint8_t)avail4(d, xx << sub, yy << sub);
    if (ta[i]) top[i] = pl[yy * st + xx], any = 1;
  }
  /* 8.4.4.2.2 substitution */
  if (!any) {
    for (int i = 0; i <= 2 * n; i++) left[i] = top[i] = 128;
  } else {
    if (!la[2 * n]) {
      int found = -1, v = 128;
      for (int i = 2 * n - 1; i >= 0 && found < 0; i--)
        if (la[i]) found = i, v = left[i];
      if (found < 0)
        for (int i = 1; i <= 2 * n; i++)
          if (ta[i]) {
            v = top[i];
            break;
          }
      left[2 * n] = v;
    }
    for (int i = 2 * n - 1; i >= 0; i--)
      if (!la[i]) left[i] = left[i + 1];
    top[0] = left[0];
    for (int i = 1; i <= 2 * n; i++)
      if (!ta[i]) top[i] = top[i - 1];
  }
  /* 8.4.4.2.3 filtering (luma only in 4:2:0) */
  if (c_idx == 0 && mode != 1 && n != 4) {
    int d1 = iabs(mode - 26), d2 = iabs(mode - 10);
    int min_dist = d1 < d2 ? d1 : d2;
    int thr = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (min_dist > thr) {
      int strong = d->sps->strong_intra_smoothing_enabled_flag && n == 32 &&
                   iabs(top[0] + top[64] - 2 * top[32]) < 8 && iabs(left[0] + left[64] - 2 * left[32]) < 8;
      if (strong) {
        fl[0] = ft[0] = top[0];
        for (int i = 1; i < 64; i++) {
          fl[i] = ((64 - i) * left[0] + i * left[64] + 32) >> 6;
          ft[i] = ((64 - i) * top[0] + i * top[64] + 32) >> 6;
        }
        fl[64] = left[64];
        ft[64] = top[64];
      } else {
        fl[0] = ft[0] = (left[1] + 2 * top[0] + top[1] + 2) >> 2;
        for (int i = 1; i < 2 * n; i++) {
          fl[i] = (left[i + 1] + 2 * left[i] + left[i - 1] + 2) >> 2;
          ft[i] = (top[i + 1] + 2 * top[i] + top[i - 1] + 2) >> 2;
        }
        fl[2 * n] = left[2 * n];
        ft[2 * n] = top[2 * n];
      }
      left = fl;
      top = ft;
    }
  }
  /* left[1+y] = p[-1][y], top[1+x] = p[x][-1] */
  if (mode == 0) { /* 8.4.4.2.4 planar */
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
        pred[y * n + x] = (uint8_t)(((n - 1 - x) * left[1 + y] + (x + 1) * top[1 + n] + (n - 1 - y) * top[1 + x] +
                                     (y + 1) * left[1 + n] + n) >> (log2 + 1));
  } else if (mode == 1) { /* 8.4.4.2.5 DC */
    int s = n;
    for (int i = 0; i < n; i++) s += left[1 + i] + top[1 + i];
    int dc = s >> (log2 + 1);
    for (int i = 0; i < n * n; i++) pred[i] = (uint8_t)dc;
    if (c_idx == 0 && n < 32) {
      pred[0] = (uint8_t)((left[1] + 2 * dc + top[1] + 2) >> 2);
      for (int x = 1; x < n; x++) pred[x] = (uint8_t)((top[1 + x] + 3 * dc + 2) >> 2);
      for (int y = 1; y < n; y++) pred[y * n] = (uint8_t)((left[1 + y] + 3 * dc + 2) >> 2);
    }
  } else { /* 8.4.4.2.6 angular */
    int angle = kIntraPredAngle[mode];
    int ref_buf[3 * 32 + 2];
    int* ref = ref_buf + 32; /* ref[-n .. 2n] */
    int vertical = mode >= 18;
    const int* main_ = vertical ? top : left;
    const int* side = vertical ? left : top;
    for (int i = 0; i <= n; i++) ref[i] = main_[i];
    if (angle < 0) {
      int last = (n * angle) >> 5;
      if (last < -1) {
        int inv = kInvAngle[mode - 11];
        for (int i = -1; i >= last; i--) ref[i] = side[(i * inv + 128) >> 8];
      }
    } else {
      for (int i = n + 1; i <= 2 * n; i++) ref[i] = main_[i];
    }
    for (int j = 0; j < n; j++) {   /* j: distance from the main reference (y for vertical modes) */
      int idx = ((j + 1) * angle) >> 5, fact = ((j + 1) * angle) & 31;
      for (int i = 0; i < n; i++) { /* i: position along the main reference */
        int v = fact ? ((32 - fact) * ref[i + idx + 1] + fact * ref[i + idx + 2] + 16) >> 5 : ref[i + idx + 1];
        if (vertical) pred[j * n + i] = (uint8_t)v;
        else pred[i * n + j] = (uint8_t)v;
      }
    }
    if (c_idx == 0 && n < 32) {
      if (mode == 26)
        for (int y = 0; y < n; y++) pred[y * n] = (uint8_t)clip8(top[1] + ((left[1 + y] - left[0]) >> 1));
      else if (mode == 10)
        for (int x = 0; x < n; x++) pred[x] = (uint8_t)clip8(left[1] + ((top[1 + x] - top[0]) >> 1));
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Residual coding 7.3.8.11 + 9.3.4.2.x                                                        */
/* ------------------------------------------------------------------------------------------ */
static int decode_last_sig_coeff_prefix(Dec* d, int ctx_base, int c_idx, int log2) { /* decoder.rs:109-130 */
  int ctx_offset, ctx_shift;
  if (c_idx == 0) {
    ctx_offset = 3 * (log2 - 2) + ((log2 - 1) >> 2);
    ctx_shift = (log2 + 1) >> 2;
  } else {
    ctx_offset = 15;
    ctx_shift = log2 - 2;
  }
  int c_max = (log2 << 1) - 1, v = 0;
  while (v < c_max && decode_decision(d, ctx_base + (v >> ctx_shift) + ctx_offset)) v++;
  return v;
}

/* Parses one transform block; writes TransCoeffLevel into lvl[n*n] (raster).  Returns tskip flag. */
static int residual_coding(Dec* d, int log2, int c_idx, int pred_mode, int16_t* lvl) {
  int n = 1 << log2;
  memset(lvl, 0, sizeof(int16_t) * (size_t)n * n);
  int tskip = 0;
  if (d->pps->transform_skip_enabled_flag && log2 <= 2) tskip = decode_decision(d, CTX_TSKIP + (c_idx ? 1 : 0));
  int last_x = decode_last_sig_coeff_prefix(d, CTX_LAST_X, c_idx, log2);
  int last_y = decode_last_sig_coeff_prefix(d, CTX_LAST_Y, c_idx, log2);
  if (last_x > 3) {
    int nb = (last_x >> 1) - 1;
    last_x = (1 << nb) * (2 + (last_x & 1)) + (int)decode_fl_bypass(d, nb);
  }
  if (last_y > 3) {
    int nb = (last_y >> 1) - 1;
    last_y = (1 << nb) * (2 + (last_y & 1)) + (int)decode_fl_bypass(d, nb);
  }
  int scan_idx = 0;
  if (log2 == 2 || (log2 == 3 && c_idx == 0)) {
    if (pred_mode >= 6 && pred_mode <= 14) scan_idx = 2;
    else if (pred_mode >= 22 && pred_mode <= 30) scan_idx = 1;
  }
  if (scan_idx == 2) {
    int t = last_x;
    last_x = last_y;
    last_y = t;
  }
  const uint8_t(*sb_scan)[2] = g_scan[scan_idx][log2 - 2];
  const uint8_t(*pos_scan)[2] = g_scan[scan_idx][2];
  int last_sub_block = (1 << ((log2 << 1) - 4)) - 1, last_scan_pos = 16;
  {
    int xc, yc, guard = 0;
    do {
      if (last_scan_pos == 0) {
        last_scan_pos = 16;
        last_sub_block--;
      }
      last_scan_pos--;
      if (last_sub_block < 0 || ++guard > 1024) {
        FAIL(d, HEIC_E_BITSTREAM, "last significant coefficient outside the transform block");
        return tskip;
      }
      xc = (sb_scan[last_sub_block][0] << 2) + pos_scan[last_scan_pos][0];
      yc = (sb_scan[last_sub_block][1] << 2) + pos_scan[last_scan_pos][1];
    } while (xc != last_x || yc != last_y);
  }
  uint8_t csbf[8][8];
  memset(csbf, 0, sizeof csbf);
  int sb_w = 1 << (log2 - 2);
  int greater1_ctx = 1, first_sub_block = 1;
  for (int i = last_sub_block; i >= 0 && !d->err; i--) {
    int xs = sb_scan[i][0], ys = sb_scan[i][1];
    int infer_sb_dc = 0;
    int csbf_ctx = 0;
    if (xs < sb_w - 1) csbf_ctx |= csbf[ys][xs + 1];
    if (ys < sb_w - 1) csbf_ctx |= csbf[ys + 1][xs];
    if (i < last_sub_block && i > 0) {
      csbf[ys][xs] = (uint8_t)decode_decision(d, CTX_CSBF + (c_idx ? 2 : 0) + csbf_ctx);
      infer_sb_dc = 1;
    } else {
      csbf[ys][xs] = 1;
    }
    uint8_t sig[16];
    memset(sig, 0, sizeof sig);
    int n_start = 15;
    if (i == last_sub_block) {
      n_start = last_scan_pos - 1;
      sig[last_scan_pos] = 1;
    }
    int prev_csbf = 0;
    if (xs < sb_w - 1) prev_csbf += csbf[ys][xs + 1];
    if (ys < sb_w - 1) prev_csbf += csbf[ys + 1][xs] << 1;
    for (int k = n_start; k >= 0; k--) {
      int xp = pos_scan[k][0], yp = pos_scan[k][1];
      int xc = (xs << 2) + xp, yc = (ys << 2) + yp;
      if (csbf[ys][xs] && (k > 0 || !infer_sb_dc)) {
        int sig_ctx;
        if (log2 == 2) {
          sig_ctx = kSigCtxIdxMap4x4[(yc << 2) + xc];
        } else if (xc + yc == 0) {
          sig_ctx = 0;
        } else {
          if (prev_csbf == 0) sig_ctx = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
          else if (prev_csbf == 1) sig_ctx = (yp == 0) ? 2 : (yp == 1) ? 1 : 0;
          else if (prev_csbf == 2) sig_ctx = (xp == 0) ? 2 : (xp == 1) ? 1 : 0;
          else sig_ctx = 2;
          if (c_idx == 0) {
            if (xs > 0 || ys > 0) sig_ctx += 3;
            sig_ctx += (log2 == 3) ? (scan_idx == 0 ? 9 : 15) : 21;
          } else {
            sig_ctx += (log2 == 3) ? 9 : 12;
          }
        }
        sig[k] = (uint8_t)decode_decision(d, CTX_SIG + (c_idx ? 27 : 0) + sig_ctx);
        if (sig[k]) infer_sb_dc = 0;
      } else if (k == 0 && infer_sb_dc && csbf[ys][xs]) {
        sig[0] = 1;
      }
    }
    /* greater1 / greater2 (9.3.4.2.6, 9.3.4.2.7) */
    int first_sig = 16, last_sig = -1, num_g1 = 0, last_g1_pos = -1;
    uint8_t g1[16], g2[16];
    memset(g1, 0, sizeof g1);
    memset(g2, 0, sizeof g2);
    int n_sig = 0;
    for (int k = 15; k >= 0; k--) n_sig += sig[k];
    if (!n_sig) continue;
    int ctx_set = (i > 0 && c_idx == 0) ? 2 : 0;
    if (!first_sub_block && greater1_ctx == 0) ctx_set++;
    first_sub_block = 0;
    greater1_ctx = 1;
    for (int k = 15; k >= 0; k--) {
      if (!sig[k]) continue;
      if (num_g1 < 8) {
        g1[k] = (uint8_t)decode_decision(d, CTX_GT1 + (c_idx ? 16 : 0) + (ctx_set << 2) + greater1_ctx);
        num_g1++;
        if (g1[k]) {
          greater1_ctx = 0;
          if (last_g1_pos < 0) last_g1_pos = k;
        } else if (greater1_ctx > 0 && greater1_ctx < 3) {
          greater1_ctx++;
        }
      }
      if (last_sig < 0) last_sig = k;
      first_sig = k;
    }
    int sign_hidden = d->pps->sign_data_hiding_enabled_flag && (last_sig - first_sig > 3);
    if (last_g1_pos >= 0) g2[last_g1_pos] = (uint8_t)decode_decision(d, CTX_GT2 + (c_idx ? 4 : 0) + ctx_set);
    uint8_t sign[16];
    memset(sign, 0, sizeof sign);
    for (int k = 15; k >= 0; k--)
      if (sig[k] && (!sign_hidden || k != first_sig)) sign[k] = (uint8_t)decode_bypass(d);
    int num_sig = 0, sum_abs = 0, rice = 0;
    for (int k = 15; k >= 0; k--) {
      if (!sig[k]) continue;
      int base = 1 + g1[k] + g2[k];
      int abs_level = base;
      if (base == ((num_sig < 8) ? ((k == last_g1_pos) ? 3 : 2) : 1)) {
        uint32_t rem = decode_coeff_abs_level_remaining(d, rice);
        if (rem > 32768u) {
          FAIL(d, HEIC_E_BITSTREAM, "coeff_abs_level_remaining out of range");
          return tskip;
        }
        abs_level = base + (int)rem;
        if (abs_level > 3 * (1 << rice)) rice = rice < 4 ? rice + 1 : 4; /* decoder.rs:230-236 */
      }
      int v = sign[k] ? -abs_level : abs_level;
      if (sign_hidden) {
        sum_abs += abs_level;
        if (k == first_sig && (sum_abs & 1)) v = -v;
      }
      int xc = (xs << 2) + pos_scan[k][0], yc = (ys << 2) + pos_scan[k][1];
      lvl[yc * n + xc] = (int16_t)clip3(-32768, 32767, v);
      num_sig++;
    }
  }
  return tskip;
}

/* 8.6.2-8.6.4: scaling + transform of one block -> residual */
static void dequant_transform(Dec* d, const int16_t* lvl, int16_t* res, int log2, int c_idx, int qp, int tskip) {
  int n = 1 << log2;
  int16_t coef[32 * 32];
  const uint8_t* m = d->scaling[log2 - 2][c_idx];
  int bd_shift = 8 + log2 - 5;
  int64_t scale = (int64_t)kLevelScale[qp % 6] << (qp / 6);
  for (int i = 0; i < n * n; i++) {
    if (!lvl[i]) {
      coef[i] = 0;
      continue;
    }
    int mm = (d->sps->scaling_list_enabled_flag && !(tskip && n > 4)) ? m[i] : 16;
    int64_t v = ((int64_t)lvl[i] * mm * scale + ((int64_t)1 << (bd_shift - 1))) >> bd_shift;
    coef[i] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
  }
  if (tskip) {
    for (int i = 0; i < n * n; i++) res[i] = (int16_t)((((int)coef[i] << 7) + 2048) >> 12);
  } else {
    hevc_oracle_idct(coef, res, log2, c_idx == 0 && log2 == 2);
  }
}

static int chroma_qp(const Dec* d, int qp_y, int c_idx) {
  int off = c_idx == 1 ? d->pps->pps_cb_qp_offset + d->sh->slice_cb_qp_offset
                       : d->pps->pps_cr_qp_offset + d->sh->slice_cr_qp_offset;
  int qpi = clip3(0, 57, qp_y + off);
  if (qpi < 30) return qpi;
  if (qpi >= 43) return qpi - 6;
  return kChromaQpTable[qpi - 30];
}

/* ------------------------------------------------------------------------------------------ */
/* Syntax: SAO, quadtree, CU, TU                                                               */
/* ------------------------------------------------------------------------------------------ */
static void parse_sao(Dec* d, int rx, int ry) { /* 7.3.8.3 — todo!() at slice.rs:249-251 */
  SaoParams* p = &d->sao[ry * d->wctb + rx];
  memset(p, 0, sizeof *p);
  int merge_left = 0, merge_up = 0;
  if (rx > 0) merge_left = decode_decision(d, CTX_SAO_MERGE);
  if (ry > 0 && !merge_left) merge_up = decode_decision(d, CTX_SAO_MERGE);
  if (merge_left) {
    *p = d->sao[ry * d->wctb + rx - 1];
    return;
  }
  if (merge_up) {
    *p = d->sao[(ry - 1) * d->wctb + rx];
    return;
  }
  int n_comp = d->chroma ? 3 : 1;
  for (int c = 0; c < n_comp; c++) {
    if (!((c == 0 && d->sh->slice_sao_luma_flag) || (c > 0 && d->sh->slice_sao_chroma_flag))) continue;
    if (c == 2) {
      p->type[2] = p->type[1];
    } else { /* sao_type_idx: TR cMax 2, bin0 ctx, bin1 bypass (decoder.rs:93-100) */
      int t = 0;
      if (decode_decision(d, CTX_SAO_TYPE)) t = decode_bypass(d) ? 2 : 1;
      p->type[c] = (uint8_t)t;
    }
    if (!p->type[c]) continue;
    int abs_v[4];
    for (int i = 0; i < 4; i++) abs_v[i] = (int)decode_tr_bypass(d, 7);
    if (p->type[c] == 1) {
      for (int i = 0; i < 4; i++) {
        int neg = abs_v[i] ? decode_bypass(d) : 0;
        p->offset[c][i] = (int8_t)(neg ? -abs_v[i] : abs_v[i]);
      }
      p->band_pos[c] = (uint8_t)decode_fl_bypass(d, 5);
    } else {
      if (c == 0) p->eo_class[0] = (uint8_t)decode_fl_bypass(d, 2);
      else if (c == 1) p->eo_class[1] = (uint8_t)decode_fl_bypass(d, 2);
      else p->eo_class[2] = p->eo_class[1];
      p->offset[c][0] = (int8_t)abs_v[0];
      p->offset[c][1] = (int8_t)abs_v[1];
      p->offset[c][2] = (int8_t)-abs_v[2];
      p->offset[c][3] = (int8_t)-abs_v[3];
    }
  }
}

static void set_qp_pred(Dec* d, int x_qg, int y_qg) { /* 8.6.1, called at each quantisation-group start */
  int qp_prev = d->first_qg_in_row ? d->slice_qp : d->last_qp_y;
  d->first_qg_in_row = 0;
  int ctb_mask = d->ctb_size - 1;
  int qa = qp_prev, qb = qp_prev;
  if (x_qg & ctb_mask) qa = d->qp_map[(y_qg >> 3) * d->w8 + ((x_qg - 1) >> 3)];
  if (y_qg & ctb_mask) qb = d->qp_map[((y_qg - 1) >> 3) * d->w8 + (x_qg >> 3)];
  d->qp_y_pred = (qa + qb + 1) >> 1;
  d->qp_y = d->qp_y_pred;
}

static void reconstruct_block(Dec* d, int xl, int yl, int log2, int c_idx, int mode, int cbf, const int16_t* res) {
  /* (xl, yl): luma position of the TU; block is n x n in component c_idx */
  int n = 1 << log2;
  int sub = (c_idx && d->chroma) ? 1 : 0;
  int x0 = xl >> sub, y0 = yl >> sub;
  uint8_t pred[32 * 32];
  intra_predict(d, x0, y0, log2, c_idx, mode, pred);
  uint8_t* pl = d->plane[c_idx];
  int st = d->stride[c_idx];
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++) pl[(y0 + y) * st + x0 + x] = (uint8_t)(cbf ? clip8(pred[y * n + x] + res[y * n + x]) : pred[y * n + x]);
}

static void store_block16(int16_t* dst, uint32_t off, const int16_t* src, int n) {
  if (dst) memcpy(dst + off, src, sizeof(int16_t) * (size_t)n * n);
}

/* 7.3.8.10 transform_unit (+ reconstruction 8.4.4.1 order: luma, then Cb, Cr) */
static void transform_unit(Dec* d, int x0, int y0, int x_base, int y_base, int log2, int blk_idx, int cbf_luma,
                           int cbf_cb, int cbf_cr) {
  int luma_mode = d->ipm[(y0 >> 2) * d->w4 + (x0 >> 2)];
  int has_chroma = d->chroma && (log2 > 2 || blk_idx == 3);
  int xc = log2 > 2 ? x0 : x_base, yc = log2 > 2 ? y0 : y_base;
  int log2c = log2 > 2 ? log2 - 1 : 2;
  /* 7.3.8.10: cbfChroma uses the (possibly parent-inherited) chroma cbfs even for 4x4 luma blocks 0..2 */
  int any_cbf = cbf_luma || cbf_cb || cbf_cr;
  if (!has_chroma) cbf_cb = cbf_cr = 0;
  int16_t lvl[3][32 * 32], res[32 * 32];
  int tskip[3] = {0, 0, 0};
  if (any_cbf) {
    if (d->pps->cu_qp_delta_enabled_flag && !d->is_cu_qp_delta_coded) {
      /* cu_qp_delta_abs: TR5 with ctx (bin0: 0, bins1-4: 1) + EG0 bypass (decoder.rs:263-284) */
      int v = 0;
      while (v < 5 && decode_decision(d, CTX_CU_QP_DELTA + (v ? 1 : 0))) v++;
      if (v == 5) v += (int)decode_egk_bypass(d, 0);
      int neg = v ? decode_bypass(d) : 0;
      d->is_cu_qp_delta_coded = 1;
      d->cu_qp_delta_val = neg ? -v : v;
      if (d->cu_qp_delta_val < -26 || d->cu_qp_delta_val > 25) FAIL(d, HEIC_E_BITSTREAM, "CuQpDeltaVal out of range");
      d->qp_y = ((d->qp_y_pred + d->cu_qp_delta_val + 52) % 52);
    }
  }
  if (d->err) return;
  if (cbf_luma) tskip[0] = residual_coding(d, log2, 0, luma_mode, lvl[0]);
  if (cbf_cb) tskip[1] = residual_coding(d, log2c, 1, d->chroma_mode, lvl[1]);
  if (cbf_cr) tskip[2] = residual_coding(d, log2c, 2, d->chroma_mode, lvl[2]);
  if (d->err) return;

  uint32_t ti = tu_index(d, x0, y0);
  if (d->out->tu_map) {
    d->out->tu_map[ti] = 1u | ((uint32_t)(log2 - 2) << 1) | ((uint32_t)cbf_luma << 3) | ((uint32_t)cbf_cb << 4) |
                         ((uint32_t)cbf_cr << 5) | ((uint32_t)has_chroma << 6) | ((uint32_t)luma_mode << 7) |
                         ((uint32_t)d->chroma_mode << 13) | ((uint32_t)d->qp_y << 19) | ((uint32_t)tskip[0] << 25) |
                         ((uint32_t)tskip[1] << 26) | ((uint32_t)tskip[2] << 27);
  }
  /* TU edges for deblocking (8.7.2.3) */
  for (int i = 0; i < (1 << log2) >> 2; i++) {
    d->edge[((y0 >> 2) + i) * d->w4 + (x0 >> 2)] |= 1;
    d->edge[(y0 >> 2) * d->w4 + (x0 >> 2) + i] |= 2;
  }
  int n = 1 << log2;
  uint32_t off_y = coeff_offset(d, x0, y0, 0);
  if (cbf_luma) store_block16(d->out->level[0], off_y, lvl[0], n);
  if (!d->parse_only) {
    if (cbf_luma) {
      dequant_transform(d, lvl[0], res, log2, 0, d->qp_y, tskip[0]);
      store_block16(d->out->resid[0], off_y, res, n);
    }
    reconstruct_block(d, x0, y0, log2, 0, luma_mode, cbf_luma, res);
  }
  for (int y = 0; y < n >> 2; y++)
    for (int x = 0; x < n >> 2; x++) d->decoded4[((y0 >> 2) + y) * d->w4 + (x0 >> 2) + x] = 1;
  if (has_chroma) {
    int nc = 1 << log2c;
    uint32_t off_c = coeff_offset(d, xc, yc, 1);
    for (int c = 1; c <= 2; c++) {
      int cbf = c == 1 ? cbf_cb : cbf_cr;
      if (cbf) store_block16(d->out->level[c], off_c, lvl[c], nc);
      if (!d->parse_only) {
        if (cbf) {
          dequant_transform(d, lvl[c], res, log2c, c, chroma_qp(d, d->qp_y, c), tskip[c]);
          store_block16(d->out->resid[c], off_c, res, nc);
        }
        reconstruct_block(d, xc, yc, log2c, c, d->chroma_mode, cbf, res);
      }
    }
  }
}

/* 7.3.8.8 transform_tree */
static void transform_tree(Dec* d, int x0, int y0, int x_base, int y_base, int log2, int depth, int blk_idx,
                           int parent_cbf_cb, int parent_cbf_cr) {
  if (d->err) return;
  int intra_split = d->part_nxn;
  int max_depth = d->max_trafo_depth_intra + intra_split;
  int split;
  if (log2 <= d->log2_max_tb && log2 > d->log2_min_tb && depth < max_depth && !(intra_split && depth == 0))
    split = decode_decision(d, CTX_SPLIT_TRANSFORM + 5 - log2);
  else
    split = (log2 > d->log2_max_tb) || (intra_split && depth == 0);
  int cbf_cb = 0, cbf_cr = 0;
  if (d->chroma) {
    if (log2 > 2) {
      if (depth == 0 || parent_cbf_cb) cbf_cb = decode_decision(d, CTX_CBF_CHROMA + depth);
      if (depth == 0 || parent_cbf_cr) cbf_cr = decode_decision(d, CTX_CBF_CHROMA + depth);
    } else { /* inferred from the parent when trafoDepth > 0 && log2TrafoSize == 2 */
      cbf_cb = depth > 0 ? parent_cbf_cb : 0;
      cbf_cr = depth > 0 ? parent_cbf_cr : 0;
    }
  }
  if (split) {
    int h = 1 << (log2 - 1);
    transform_tree(d, x0, y0, x0, y0, log2 - 1, depth + 1, 0, cbf_cb, cbf_cr);
    transform_tree(d, x0 + h, y0, x0, y0, log2 - 1, depth + 1, 1, cbf_cb, cbf_cr);
    transform_tree(d, x0, y0 + h, x0, y0, log2 - 1, depth + 1, 2, cbf_cb, cbf_cr);
    transform_tree(d, x0 + h, y0 + h, x0, y0, log2 - 1, depth + 1, 3, cbf_cb, cbf_cr);
  } else {
    int cbf_luma = decode_decision(d, CTX_CBF_LUMA + (depth == 0 ? 1 : 0)); /* always present for intra */
    transform_unit(d, x0, y0, x_base, y_base, log2, blk_idx, cbf_luma, cbf_cb, cbf_cr);
  }
}

/* 8.4.2 luma intra prediction mode */
static int derive_luma_mode(Dec* d, int x, int y, int prev_flag, int mpm_idx, int rem) {
  int cand_a = 1, cand_b = 1; /* DC when unavailable */
  if (x > 0) cand_a = d->ipm[(y >> 2) * d->w4 + ((x - 1) >> 2)];
  if (y > 0 && ((y - 1) >> d->log2_ctb) == (y >> d->log2_ctb)) cand_b = d->ipm[((y - 1) >> 2) * d->w4 + (x >> 2)];
  int c[3];
  if (cand_a == cand_b) {
    if (cand_a < 2) {
      c[0] = 0;
      c[1] = 1;
      c[2] = 26;
    } else {
      c[0] = cand_a;
      c[1] = 2 + ((cand_a + 29) % 32);
      c[2] = 2 + ((cand_a - 2 + 1) % 32);
    }
  } else {
    c[0] = cand_a;
    c[1] = cand_b;
    if (cand_a != 0 && cand_b != 0) c[2] = 0;
    else if (cand_a != 1 && cand_b != 1) c[2] = 1;
    else c[2] = 26;
  }
  if (prev_flag) return c[mpm_idx];
  if (c[0] > c[1]) { int t = c[0]; c[0] = c[1]; c[1] = t; }
  if (c[0] > c[2]) { int t = c[0]; c[0] = c[2]; c[2] = t; }
  if (c[1] > c[2]) { int t = c[1]; c[1] = c[2]; c[2] = t; }
  int mode = rem;
  for (int i = 0; i < 3; i++)
    if (mode >= c[i]) mode++;
  return mode;
}

/* intra_chroma_pred_mode binarisation, Table 9-41 (decoder.rs:23-35,192-204): "0" -> 4, "1" + FL(2) -> 0..3 */
static int decode_intra_chroma_pred_mode_idx(Dec* d) {
  if (!decode_decision(d, CTX_CHROMA_PRED)) return 4;
  return (int)decode_fl_bypass(d, 2);
}

/* 7.3.8.5 coding_unit (I slice) */
static void coding_unit(Dec* d, int x0, int y0, int log2) {
  int n = 1 << log2;
  d->cu_x = x0;
  d->cu_y = y0;
  d->cu_log2 = log2;
  d->part_nxn = 0;
  if (d->pps->cu_qp_delta_enabled_flag) { /* quantisation group of this CU (8.6.1) */
    int mask = (1 << d->log2_min_cu_qp_delta_size) - 1;
    int x_qg = x0 & ~mask, y_qg = y0 & ~mask;
    if (x_qg != d->qg_x || y_qg != d->qg_y) {
      d->qg_x = x_qg;
      d->qg_y = y_qg;
      set_qp_pred(d, x_qg, y_qg);
    }
    d->qp_y = (d->qp_y_pred + d->cu_qp_delta_val + 52) % 52;
  }
  if (log2 == d->log2_min_cb) d->part_nxn = !decode_decision(d, CTX_PART_MODE); /* decoder.rs:136-149 */
  if (d->part_nxn && log2 == 3 && d->log2_min_tb >= 3) {
    FAIL(d, HEIC_E_BITSTREAM, "PART_NxN with 8x8 CU requires 4x4 transform blocks");
    return;
  }
  int n_pu = d->part_nxn ? 2 : 1, pb = n / n_pu;
  int prev[4], k = 0;
  for (int j = 0; j < n_pu; j++)
    for (int i = 0; i < n_pu; i++) prev[k++] = decode_decision(d, CTX_PREV_INTRA);
  k = 0;
  for (int j = 0; j < n_pu; j++)
    for (int i = 0; i < n_pu; i++, k++) {
      int mpm_idx = 0, rem = 0;
      if (prev[k]) mpm_idx = (int)decode_tr_bypass(d, 2);
      else rem = (int)decode_fl_bypass(d, 5);
      int px = x0 + i * pb, py = y0 + j * pb;
      int mode = derive_luma_mode(d, px, py, prev[k], mpm_idx, rem);
      for (int yy = 0; yy < pb >> 2; yy++) memset(&d->ipm[((py >> 2) + yy) * d->w4 + (px >> 2)], mode, (size_t)(pb >> 2));
    }
  d->chroma_mode = 0;
  if (d->chroma) { /* intra_chroma_pred_mode (decoder.rs:23-35,192-204) + 8.4.3 */
    int idx = decode_intra_chroma_pred_mode_idx(d);
    int luma = d->ipm[(y0 >> 2) * d->w4 + (x0 >> 2)];
    static const uint8_t kMap[4] = {0, 26, 10, 1};
    if (idx == 4) d->chroma_mode = luma;
    else d->chroma_mode = kMap[idx] == luma ? 34 : kMap[idx];
  }
  transform_tree(d, x0, y0, x0, y0, log2, 0, 0, 0, 0);
  /* QpY of the CU (8.6.1) -> qp_map; CuQpDeltaVal decoded inside this CU applies to the whole CU */
  for (int yy = y0 >> 3; yy < (y0 + n) >> 3; yy++) memset(&d->qp_map[yy * d->w8 + (x0 >> 3)], d->qp_y, (size_t)(n >> 3));
  d->last_qp_y = d->qp_y;
}

/* 7.3.8.4 coding_quadtree — todo!() at slice.rs:253-255 */
static void coding_quadtree(Dec* d, int x0, int y0, int log2, int depth) {
  if (d->err) return;
  int n = 1 << log2, split;
  if (x0 + n <= d->w && y0 + n <= d->h && log2 > d->log2_min_cb) {
    int inc = 0;
    if (x0 > 0 && d->ct_depth[(y0 >> 3) * d->w8 + ((x0 - 1) >> 3)] > depth) inc++;
    if (y0 > 0 && d->ct_depth[((y0 - 1) >> 3) * d->w8 + (x0 >> 3)] > depth) inc++;
    split = decode_decision(d, CTX_SPLIT_CU + inc);
  } else {
    split = log2 > d->log2_min_cb;
  }
  if (d->pps->cu_qp_delta_enabled_flag && log2 >= d->log2_min_cu_qp_delta_size) {
    d->is_cu_qp_delta_coded = 0;
    d->cu_qp_delta_val = 0;
  }
  if (split) {
    int h = n >> 1;
    coding_quadtree(d, x0, y0, log2 - 1, depth + 1);
    if (x0 + h < d->w) coding_quadtree(d, x0 + h, y0, log2 - 1, depth + 1);
    if (y0 + h < d->h) coding_quadtree(d, x0, y0 + h, log2 - 1, depth + 1);
    if (x0 + h < d->w && y0 + h < d->h) coding_quadtree(d, x0 + h, y0 + h, log2 - 1, depth + 1);
  } else {
    for (int yy = y0 >> 3; yy < (y0 + n) >> 3; yy++) memset(&d->ct_depth[yy * d->w8 + (x0 >> 3)], depth, (size_t)(n >> 3));
    coding_unit(d, x0, y0, log2);
  }
}

/* 7.3.8.1 slice_segment_data — skeleton of slice.rs:206-247 with the WPP handling it lacks
 * (engine re-init at each entry point 9.3.2.5, context sync/storage 9.3.2.2/9.3.2.4). */
static void slice_segment_data(Dec* d) {
  const heic_slice_header* sh = d->sh;
  int wpp = d->pps->entropy_coding_sync_enabled_flag;
  uint32_t base = sh->slice_data_byte_offset;
  int n_ctb = d->wctb * d->hctb;
  for (int addr = 0; addr < n_ctb && !d->err; addr++) {
    int rx = addr % d->wctb, ry = addr / d->wctb;
    if (addr == 0 || (wpp && rx == 0)) {
      uint32_t start = base + (wpp ? sh->substream_offset[ry] : 0);
      uint32_t end = (wpp && (uint32_t)ry < sh->num_entry_point_offsets) ? base + sh->substream_offset[ry + 1] : d->rbsp_len_;
      if (start > d->rbsp_len_ || end > d->rbsp_len_ || start > end) {
        FAIL(d, HEIC_E_BITSTREAM, "substream %d outside the slice data", ry);
        return;
      }
      cabac_init_engine(d, start, end);
      if (addr == 0 || d->wctb == 1) hevc_oracle_context_init(d->slice_qp, d->ctx);
      else memcpy(d->ctx, d->ctx_wpp, NUM_CTX);
      d->first_qg_in_row = 1;
    }
    if (!d->pps->cu_qp_delta_enabled_flag) d->qp_y = d->slice_qp;
    if (sh->slice_sao_luma_flag || sh->slice_sao_chroma_flag) parse_sao(d, rx, ry);
    coding_quadtree(d, rx << d->log2_ctb, ry << d->log2_ctb, d->log2_ctb, 0);
    if (d->err) return;
    d->out->ctus++;
    if (wpp && rx == 1) memcpy(d->ctx_wpp, d->ctx, NUM_CTX);
    int end_of_slice = decode_terminate(d);
    if (end_of_slice) {
      if (addr != n_ctb - 1) FAIL(d, HEIC_E_BITSTREAM, "end_of_slice_segment_flag before the last CTU (addr %d)", addr);
      return;
    }
    if (addr == n_ctb - 1) {
      FAIL(d, HEIC_E_BITSTREAM, "missing end_of_slice_segment_flag");
      return;
    }
    if (wpp && rx == d->wctb - 1) {
      if (!decode_terminate(d)) {
        FAIL(d, HEIC_E_BITSTREAM, "end_of_subset_one_bit is 0");
        return;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Deblocking 8.7.2                                                                            */
/* ------------------------------------------------------------------------------------------ */
static void deblock_luma_edge(Dec* d, uint8_t* px, int step_across, int step_along, int qp_p, int qp_q) {
  /* px points at q0 of line 0; p_i = px[-(i+1)*step_across], q_i = px[i*step_across] */
  int qpl = (qp_p + qp_q + 1) >> 1;
  int beta = kBetaTable[clip3(0, 51, qpl + (d->sh->slice_beta_offset_div2 << 1))];
  int tc = kTcTable[clip3(0, 53, qpl + 2 + (d->sh->slice_tc_offset_div2 << 1))];
#define P(i, l) ((int)px[-((i) + 1) * step_across + (l) * step_along])
#define Q(i, l) ((int)px[(i) * step_across + (l) * step_along])
  int dp0 = iabs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = iabs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
  int dq0 = iabs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = iabs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
  int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
  if (dpq0 + dpq3 >= beta) return;
  int s0 = 2 * dpq0 < (beta >> 2) && iabs(P(3, 0) - P(0, 0)) + iabs(Q(0, 0) - Q(3, 0)) < (beta >> 3) &&
           iabs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
  int s3 = 2 * dpq3 < (beta >> 2) && iabs(P(3, 3) - P(0, 3)) + iabs(Q(0, 3) - Q(3, 3)) < (beta >> 3) &&
           iabs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
  int strong = s0 && s3;
  int dep = dp < ((beta + (beta >> 1)) >> 3), deq = dq < ((beta + (beta >> 1)) >> 3);
  for (int l = 0; l < 4; l++) {
    int p0 = P(0, l), p1 = P(1, l), p2 = P(2, l), p3 = P(3, l);
    int q0 = Q(0, l), q1 = Q(1, l), q2 = Q(2, l), q3 = Q(3, l);
    uint8_t* line = px + l * step_along;
    if (strong) {
      line[-1 * step_across] = (uint8_t)clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      line[-2 * step_across] = (uint8_t)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      line[-3 * step_across] = (uint8_t)clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      line[0] = (uint8_t)clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      line[step_across] = (uint8_t)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      line[2 * step_across] = (uint8_t)clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    } else {
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (iabs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        line[-1 * step_across] = (uint8_t)clip8(p0 + delta);
        line[0] = (uint8_t)clip8(q0 - delta);
        if (dep) {
          int dlt = clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1);
          line[-2 * step_across] = (uint8_t)clip8(p1 + dlt);
        }
        if (deq) {
          int dlt = clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1);
          line[step_across] = (uint8_t)clip8(q1 + dlt);
        }
      }
    }
  }
#undef P
#undef Q
}

static void deblock_chroma_edge(Dec* d, uint8_t* px, int step_across, int step_along, int qp_p, int qp_q, int c_idx) {
  int off = c_idx == 1 ? d->pps->pps_cb_qp_offset : d->pps->pps_cr_qp_offset; /* cQpPicOffset: PPS only */
  int qpi = ((qp_p + qp_q + 1) >> 1) + off;
  int qpc = qpi < 30 ? qpi : (qpi >= 43 ? qpi - 6 : kChromaQpTable[qpi - 30]);
  int tc = kTcTable[clip3(0, 53, qpc + 2 + (d->sh->slice_tc_offset_div2 << 1))];
  if (!tc) return;
  for (int l = 0; l < 4; l++) {
    uint8_t* line = px + l * step_along;
    int p0 = line[-step_across], p1 = line[-2 * step_across], q0 = line[0], q1 = line[step_across];
    int delta = clip3(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
    line[-step_across] = (uint8_t)clip8(p0 + delta);
    line[0] = (uint8_t)clip8(q0 - delta);
  }
}

static void deblock_picture(Dec* d) {
  if (d->sh->slice_deblocking_filter_disabled_flag) return;
  for (int dir = 0; dir < 2; dir++) { /* 0: vertical edges over the whole picture, then 1: horizontal */
    /* luma: 8x8 grid, 4-line segments */
    for (int y4 = 0; y4 < d->h >> 2; y4++)
      for (int x4 = 0; x4 < d->w >> 2; x4++) {
        int x = x4 << 2, y = y4 << 2;
        if (dir == 0) {
          if (x == 0 || (x & 7) || !(d->edge[y4 * d->w4 + x4] & 1)) continue;
          int qp_p = d->qp_map[(y >> 3) * d->w8 + ((x - 1) >> 3)], qp_q = d->qp_map[(y >> 3) * d->w8 + (x >> 3)];
          deblock_luma_edge(d, d->plane[0] + y * d->stride[0] + x, 1, d->stride[0], qp_p, qp_q);
        } else {
          if (y == 0 || (y & 7) || !(d->edge[y4 * d->w4 + x4] & 2)) continue;
          int qp_p = d->qp_map[((y - 1) >> 3) * d->w8 + (x >> 3)], qp_q = d->qp_map[(y >> 3) * d->w8 + (x >> 3)];
          deblock_luma_edge(d, d->plane[0] + y * d->stride[0] + x, d->stride[0], 1, qp_p, qp_q);
        }
      }
    if (!d->chroma) continue;
    /* chroma 4:2:0: edges on the 8-sample chroma grid (16 luma), 4 chroma lines (8 luma) per segment */
    for (int c = 1; c <= 2; c++)
      for (int y8 = 0; y8 < d->h >> 3; y8++)
        for (int x8 = 0; x8 < d->w >> 3; x8++) {
          int x = x8 << 3, y = y8 << 3; /* luma */
          if (dir == 0) {
            if (x == 0 || (x & 15) || !(d->edge[(y >> 2) * d->w4 + (x >> 2)] & 1)) continue;
            int qp_p = d->qp_map[y8 * d->w8 + x8 - 1], qp_q = d->qp_map[y8 * d->w8 + x8];
            deblock_chroma_edge(d, d->plane[c] + (y >> 1) * d->stride[c] + (x >> 1), 1, d->stride[c], qp_p, qp_q, c);
          } else {
            if (y == 0 || (y & 15) || !(d->edge[(y >> 2) * d->w4 + (x >> 2)] & 2)) continue;
            int qp_p = d->qp_map[(y8 - 1) * d->w8 + x8], qp_q = d->qp_map[y8 * d->w8 + x8];
            deblock_chroma_edge(d, d->plane[c] + (y >> 1) * d->stride[c] + (x >> 1), d->stride[c], 1, qp_p, qp_q, c);
          }
        }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SAO 8.7.3                                                                                   */
/* ------------------------------------------------------------------------------------------ */
static void sao_picture(Dec* d, uint8_t* const src[3], uint8_t* const dst[3]) {
  static const int8_t kHPos[4][2] = {{-1, 1}, {0, 0}, {-1, 1}, {1, -1}};
  static const int8_t kVPos[4][2] = {{0, 0}, {-1, 1}, {-1, 1}, {-1, 1}};
  int n_comp = d->chroma ? 3 : 1;
  for (int c = 0; c < n_comp; c++) {
    int w = c ? d->wc : d->w, h = c ? d->hc : d->h, st = d->stride[c];
    int cs = c ? d->ctb_size >> 1 : d->ctb_size;
    memcpy(dst[c], src[c], (size_t)st * h);
    if (!((c == 0 && d->sh->slice_sao_luma_flag) || (c > 0 && d->sh->slice_sao_chroma_flag))) continue;
    for (int ry = 0; ry < d->hctb; ry++)
      for (int rx = 0; rx < d->wctb; rx++) {
        const SaoParams* p = &d->sao[ry * d->wctb + rx];
        if (!p->type[c]) continue;
        int x_end = (rx + 1) * cs < w ? (rx + 1) * cs : w, y_end = (ry + 1) * cs < h ? (ry + 1) * cs : h;
        int off[5] = {0, p->offset[c][0], p->offset[c][1], p->offset[c][2], p->offset[c][3]};
        if (p->type[c] == 1) {
          int band_table[32];
          memset(band_table, 0, sizeof band_table);
          for (int k = 0; k < 4; k++) band_table[(k + p->band_pos[c]) & 31] = k + 1;
          for (int y = ry * cs; y < y_end; y++)
            for (int x = rx * cs; x < x_end; x++) dst[c][y * st + x] = (uint8_t)clip8(src[c][y * st + x] + off[band_table[src[c][y * st + x] >> 3]]);
        } else {
          int cl = p->eo_class[c];
          for (int y = ry * cs; y < y_end; y++)
            for (int x = rx * cs; x < x_end; x++) {
              int xa = x + kHPos[cl][0], ya = y + kVPos[cl][0], xb = x + kHPos[cl][1], yb = y + kVPos[cl][1];
              if (xa < 0 || ya < 0 || xa >= w || ya >= h || xb < 0 || yb < 0 || xb >= w || yb >= h) continue;
              int v = src[c][y * st + x], a = src[c][ya * st + xa], b = src[c][yb * st + xb];
              int e = 2 + (v > a) - (v < a) + (v > b) - (v < b);
              if (e <= 2) e = (e == 2) ? 0 : e + 1;
              dst[c][y * st + x] = (uint8_t)clip8(v + off[e]);
            }
        }
      }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* Entry points                                                                                */
/* ------------------------------------------------------------------------------------------ */
static int decode_impl(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh, const uint8_t* rbsp,
                       uint32_t rbsp_len, hevc_oracle_out* out, int parse_only) {
  init_tables();
  Dec* d = (Dec*)calloc(1, sizeof(Dec));
  if (!d) return HEIC_E_NOMEM;
  d->sps = sps;
  d->pps = pps;
  d->sh = sh;
  d->out = out;
  d->parse_only = parse_only;
  out->bins = out->ctus = 0;
  out->error[0] = 0;
  int rc = 0;
#define UNSUPPORTED(cond, msg)               \
  if (cond) {                                \
    FAIL(d, HEIC_E_UNSUPPORTED, "%s", msg);  \
    goto done;                               \
  }
  UNSUPPORTED(sps->chroma_format_idc > 1, "only 4:2:0 and 4:0:0 are supported");
  UNSUPPORTED(sps->bit_depth_luma_minus8 || sps->bit_depth_chroma_minus8, "only 8-bit is supported");
  UNSUPPORTED(sps->pcm_enabled_flag, "PCM is not supported");
  UNSUPPORTED(pps->transquant_bypass_enabled_flag, "cu_transquant_bypass is not supported");
  UNSUPPORTED(pps->tiles_enabled_flag, "HEVC tiles are not supported (HEIF grid tiles are separate pictures)");
  UNSUPPORTED(sh->slice_type != 2, "only I slices are supported");
  d->w = (int)sps->pic_width_in_luma_samples;
  d->h = (int)sps->pic_height_in_luma_samples;
  d->chroma = sps->chroma_format_idc == 1;
  d->wc = d->w >> 1;
  d->hc = d->h >> 1;
  d->log2_min_cb = (int)sps->log2_min_luma_coding_block_size_minus3 + 3;
  d->log2_ctb = d->log2_min_cb + (int)sps->log2_diff_max_min_luma_coding_block_size;
  d->ctb_size = 1 << d->log2_ctb;
  d->log2_min_tb = (int)sps->log2_min_luma_transform_block_size_minus2 + 2;
  d->log2_max_tb = d->log2_min_tb + (int)sps->log2_diff_max_min_luma_transform_block_size;
  d->wctb = (d->w + d->ctb_size - 1) >> d->log2_ctb;
  d->hctb = (d->h + d->ctb_size - 1) >> d->log2_ctb;
  d->ctb4 = d->ctb_size >> 2;
  d->w4 = d->wctb * d->ctb4;
  d->h4 = d->hctb * d->ctb4;
  d->w8 = d->w4 >> 1;
  d->h8 = d->h4 >> 1;
  d->slice_qp = 26 + pps->init_qp_minus26 + sh->slice_qp_delta;
  d->log2_min_cu_qp_delta_size = d->log2_ctb - (int)pps->diff_cu_qp_delta_depth;
  d->max_trafo_depth_intra = (int)sps->max_transform_hierarchy_depth_intra;
  d->data = rbsp;
  d->rbsp_len_ = rbsp_len;
  UNSUPPORTED(pps->entropy_coding_sync_enabled_flag && (int)sh->num_entry_point_offsets != d->hctb - 1,
              "WPP picture without one entry point per CTB row");
  d->ct_depth = (uint8_t*)calloc((size_t)d->w8 * d->h8, 1);
  d->ipm = (uint8_t*)malloc((size_t)d->w4 * d->h4);
  d->decoded4 = (uint8_t*)calloc((size_t)d->w4 * d->h4, 1);
  d->qp_map = (uint8_t*)calloc((size_t)d->w8 * d->h8, 1);
  d->edge = (uint8_t*)calloc((size_t)d->w4 * d->h4, 1);
  d->sao = (SaoParams*)calloc((size_t)d->wctb * d->hctb, sizeof(SaoParams));
  if (!d->ct_depth || !d->ipm || !d->decoded4 || !d->qp_map || !d->edge || !d->sao) {
    FAIL(d, HEIC_E_NOMEM, "out of memory");
    goto done;
  }
  memset(d->ipm, 1, (size_t)d->w4 * d->h4);
  for (int c = 0; c < 3; c++) {
    d->stride[c] = c ? d->wc : d->w;
    size_t sz = (size_t)d->stride[c] * (c ? d->hc : d->h);
    d->plane[c] = (uint8_t*)calloc(sz ? sz : 1, 1);
    if (!d->plane[c]) {
      FAIL(d, HEIC_E_NOMEM, "out of memory");
      goto done;
    }
  }
  build_scaling_factors(d);
  if (out->tu_map) memset(out->tu_map, 0, sizeof(uint32_t) * hevc_oracle_tu_map_len(sps));
  for (int c = 0; c < 3; c++) {
    if (out->level[c]) memset(out->level[c], 0, sizeof(int16_t) * hevc_oracle_coeff_len(sps, c));
    if (out->resid[c]) memset(out->resid[c], 0, sizeof(int16_t) * hevc_oracle_coeff_len(sps, c));
  }
  d->qp_y = d->last_qp_y = d->slice_qp;
  d->qg_x = d->qg_y = -1;

  slice_segment_data(d);
  out->bins = d->bins;
  if (d->err) goto done;

  if (out->qp_map)
    for (int y = 0; y < d->h >> 3; y++) memcpy(out->qp_map + y * (d->w >> 3), d->qp_map + y * d->w8, (size_t)(d->w >> 3));
  if (out->sao)
    for (int i = 0; i < d->wctb * d->hctb; i++) {
      const SaoParams* p = &d->sao[i];
      for (int c = 0; c < 3; c++) {
        uint32_t v = p->type[c] | ((uint32_t)(p->type[c] == 1 ? p->band_pos[c] : p->eo_class[c]) << 2);
        for (int k = 0; k < 4; k++) v |= ((uint32_t)(p->offset[c][k] & 15)) << (8 + 4 * k);
        out->sao[i * 4 + c] = v;
      }
      out->sao[i * 4 + 3] = 0;
    }
  if (parse_only) goto done;
  int n_comp = d->chroma ? 3 : 1;
  for (int c = 0; c < n_comp; c++)
    if (out->recon[c]) memcpy(out->recon[c], d->plane[c], (size_t)d->stride[c] * (c ? d->hc : d->h));
  deblock_picture(d);
  for (int c = 0; c < n_comp; c++)
    if (out->deblocked[c]) memcpy(out->deblocked[c], d->plane[c], (size_t)d->stride[c] * (c ? d->hc : d->h));
  {
    uint8_t* dst[3] = {out->plane[0], out->plane[1], out->plane[2]};
    if (!dst[0] || (d->chroma && (!dst[1] || !dst[2]))) {
      FAIL(d, HEIC_E_INVALID_ARG, "output planes missing");
      goto done;
    }
    sao_picture(d, d->plane, dst);
  }
done:
  rc = d->err;
  free(d->ct_depth);
  free(d->ipm);
  free(d->decoded4);
  free(d->qp_map);
  free(d->edge);
  free(d->sao);
  for (int c = 0; c < 3; c++) free(d->plane[c]);
  free(d);
  return rc;
}

int hevc_oracle_decode_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                               const uint8_t* rbsp, uint32_t rbsp_len, hevc_oracle_out* out) {
  return decode_impl(sps, pps, sh, rbsp, rbsp_len, out, 0);
}
int hevc_oracle_parse_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                              const uint8_t* rbsp, uint32_t rbsp_len, hevc_oracle_out* out) {
  return decode_impl(sps, pps, sh, rbsp, rbsp_len, out, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* Unit-test hooks: run one binarisation on a caller-supplied bin string, exactly as the reference's   */
/* tests feed decode_truncated_rice / decode_intra_chroma_pred_mode_bins from a bin iterator           */
/* (cabac/decoder.rs:286-374).  kind 0: TR prefix with cRiceParam 0 and cMax = arg (Table 9-39);        */
/* 1: intra_chroma_pred_mode (Table 9-41); 2: EGk with k = arg; 3: coeff_abs_level_remaining with        */
/* cRiceParam = arg; 4: FL with arg bits.  Returns the decoded value; *consumed = bins read.             */
/* ------------------------------------------------------------------------------------------ */
uint32_t hevc_oracle_test_binarization(int kind, uint32_t arg, const uint8_t* bins, int n_bins, int* consumed) {
  Dec* d = (Dec*)calloc(1, sizeof(Dec));
  hevc_oracle_out out;
  memset(&out, 0, sizeof out);
  static const uint8_t kNone = 0;
  d->out = &out;
  d->test_bins = n_bins ? bins : &kNone;
  d->test_n = n_bins;
  uint32_t v = 0;
  switch (kind) {
    case 0: v = decode_tr_bypass(d, arg); break;
    case 1: v = (uint32_t)decode_intra_chroma_pred_mode_idx(d); break;
    case 2: v = decode_egk_bypass(d, (int)arg); break;
    case 3: v = decode_coeff_abs_level_remaining(d, (int)arg); break;
    default: v = decode_fl_bypass(d, (int)arg); break;
  }
  if (consumed) *consumed = d->test_i;
  free(d);
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* Colour conversion + grid stitch: the frozen integer definition.             */
/*   nearest-neighbour chroma (each 2x2 luma quad shares one Cb/Cr sample), 8.8 fixed point.   */
/*   full range, BT.601 (matrix_coeffs 5/6, and 2 = unspecified):                              */
/*     R = clip8(Y + ((359*d + 128) >> 8));  G = clip8(Y + ((-88*c - 183*d + 128) >> 8));      */
/*     B = clip8(Y + ((454*c + 128) >> 8));            c = Cb - 128, d = Cr - 128              */
/*   full range, BT.709 (matrix_coeffs 1): 403 / -48,-120 / 475                                */
/*   limited range: Y' = (298*(Y-16)) first, then BT.601: 409 / -100,-208 / 516 and            */
/*     BT.709: 459 / -55,-136 / 541, all >> 8 with +128 rounding.                              */
/* ------------------------------------------------------------------------------------------ */
static void color_coeffs(uint32_t full_range, uint32_t mc, int k[6]) {
  /* k = {y_mul, y_sub, rv, gu, gv, bu} */
  int bt709 = mc == 1;
  if (full_range) {
    k[0] = 256; k[1] = 0;
    if (bt709) { k[2] = 403; k[3] = -48; k[4] = -120; k[5] = 475; }
    else { k[2] = 359; k[3] = -88; k[4] = -183; k[5] = 454; }
  } else {
    k[0] = 298; k[1] = 16;
    if (bt709) { k[2] = 459; k[3] = -55; k[4] = -136; k[5] = 541; }
    else { k[2] = 409; k[3] = -100; k[4] = -208; k[5] = 516; }
  }
}

void hevc_oracle_color_stitch(const uint8_t* planes, uint32_t grid_rows, uint32_t grid_cols, uint32_t tile_w,
                              uint32_t tile_h, uint32_t out_w, uint32_t out_h, uint32_t full_range,
                              uint32_t matrix_coeffs, uint8_t* rgb, uint64_t pitch) {
  int k[6];
  color_coeffs(full_range, matrix_coeffs, k);
  uint32_t cw = tile_w / 2, ch = tile_h / 2;
  size_t tile_stride = (size_t)tile_w * tile_h + 2 * (size_t)cw * ch;
  (void)grid_rows;
  for (uint32_t y = 0; y < out_h; y++)
    for (uint32_t x = 0; x < out_w; x++) {
      uint32_t tr = y / tile_h, tc = x / tile_w, ly = y % tile_h, lx = x % tile_w;
      const uint8_t* t = planes + (size_t)(tr * grid_cols + tc) * tile_stride;
      int Y = t[(size_t)ly * tile_w + lx];
      int c = t[(size_t)tile_w * tile_h + (size_t)(ly / 2) * cw + lx / 2] - 128;
      int d = t[(size_t)tile_w * tile_h + (size_t)cw * ch + (size_t)(ly / 2) * cw + lx / 2] - 128;
      int yy = k[0] * (Y - k[1]);
      uint8_t* o = rgb + (size_t)y * pitch + (size_t)x * 3;
      o[0] = (uint8_t)clip8((yy + k[2] * d + 128) >> 8);
      o[1] = (uint8_t)clip8((yy + k[3] * c + k[4] * d + 128) >> 8);
      o[2] = (uint8_t)clip8((yy + k[5] * c + 128) >> 8);
    }
}

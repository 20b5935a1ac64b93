/*
 * hevc_oracle.h — CPU oracle for the HEIC reconstruction path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in heif_b200/ may include, link or call this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.  See oracle/README.md.
 */
#ifndef HEVC_ORACLE_H
#define HEVC_ORACLE_H
#include <stdint.h>

#include "heic_b200.h" /* the POD parameter-set / slice-header structs are the shared interface */

#ifdef __cplusplus
extern "C" {
#endif

/* All pointers optional (NULL = not wanted) except plane[].  Layouts match DESIGN.md
 * "Data layout in HBM" so each GPU stage can be compared buffer-for-buffer. */
typedef struct hevc_oracle_out {
  uint8_t* plane[3];        /* final (post-SAO) planes: w*h, (w/2)*(h/2) x2 (4:2:0) */
  uint8_t* recon[3];        /* pre-deblock reconstruction */
  uint8_t* deblocked[3];    /* post-deblock, pre-SAO */
  uint32_t* tu_map;         /* n_ctb * ctb4^2 words */
  int16_t* level[3];        /* CABAC output (TransCoeffLevel), z-order TU-contiguous layout */
  int16_t* resid[3];        /* residual after dequant + inverse transform, same layout */
  uint8_t* qp_map;          /* QpY per 8x8 luma block, raster (w/8)*(h/8) */
  uint32_t* sao;            /* 4 words per CTB */
  uint32_t bins;            /* out: number of CABAC bins decoded (decision + bypass + terminate) */
  uint32_t ctus;            /* out */
  char error[160];          /* out: message when the return value is negative */
} hevc_oracle_out;

/* Decodes one picture (one slice segment, intra only).  Returns 0 or a negative heic_status. */
int hevc_oracle_decode_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                               const uint8_t* rbsp, uint32_t rbsp_len, hevc_oracle_out* out);

/* Stops after entropy decoding (tu_map, level, qp_map, sao are produced; no reconstruction). */
int hevc_oracle_parse_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                              const uint8_t* rbsp, uint32_t rbsp_len, hevc_oracle_out* out);

/* Row C1 of SURVEY section 8: YCbCr 4:2:0 -> RGB8 + grid stitch + crop, the frozen integer definition.
 * planes: per tile Y (tile_w*tile_h) then Cb, Cr ((tile_w/2)*(tile_h/2)), tiles row-major. */
void hevc_oracle_color_stitch(const uint8_t* planes, uint32_t grid_rows, uint32_t grid_cols,
                              uint32_t tile_w, uint32_t tile_h, uint32_t out_w, uint32_t out_h,
                              uint32_t full_range, uint32_t matrix_coeffs, uint8_t* rgb, uint64_t pitch);

/* Stage helpers exposed for unit tests */
void hevc_oracle_idct(const int16_t* coeff, int16_t* resid, int log2_size, int dst);   /* 8.6.4.2, 8-bit */
void hevc_oracle_context_init(int slice_qp, uint8_t* state /* [HEVC_ORACLE_NUM_CTX] = pStateIdx<<1 | valMps */);
#define HEVC_ORACLE_NUM_CTX 134
/* Binarisations on a caller-supplied bin string (the shape of the reference's own tests, cabac/decoder.rs:286-374). */
uint32_t hevc_oracle_test_binarization(int kind, uint32_t arg, const uint8_t* bins, int n_bins, int* consumed);
uint32_t hevc_oracle_tu_map_len(const heic_sps* sps);
uint32_t hevc_oracle_coeff_len(const heic_sps* sps, int c_idx);

#ifdef __cplusplus
}
#endif
#endif

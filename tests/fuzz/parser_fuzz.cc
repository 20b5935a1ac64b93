// TEST INFRASTRUCTURE: AddressSanitizer / UBSan fuzz of the DEVICE slice-data parser (heif_b200/csrc/cuda/cabac_parse.cuh)
// in its host build (tests/emul/cabac_emul.cc).  compute-sanitizer cannot be run on the GPU pool, so this is the memory
// safety check of the CUDA syntax walker on corrupt input: every output arena is an exact-size heap block, the slice data
// and the slice header of a real tile are mutated, and the parser must either finish or flag the tile — never write or
// read out of bounds, never hang.  See tests/test_host_fuzz.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "heic_b200.h"

extern "C" int emul_parse_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh, const uint8_t* rbsp,
                                  uint32_t len, uint32_t* tu_map, int16_t* lvl0, int16_t* lvl1, int16_t* lvl2, uint8_t* qp_out,
                                  uint32_t* sao, uint32_t* bins, uint32_t* ctus, int per_row_threads);

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  FILE* fp = fopen(argv[1], "rb");
  if (!fp) return 2;
  std::vector<uint8_t> data;
  {
    uint8_t buf[65536];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, fp)) > 0) data.insert(data.end(), buf, buf + n);
  }
  fclose(fp);
  const int seed0 = atoi(argv[2]), n = atoi(argv[3]);
  heic_b200_file* f = nullptr;
  if (heic_b200_file_open(data.data(), data.size(), &f) != 0) return 2;
  const heic_image_desc* img = heic_b200_file_primary_image(f);
  if (!img || !img->n_tiles) return 2;
  const uint32_t w = img->sps.pic_width_in_luma_samples, h = img->sps.pic_height_in_luma_samples;
  const size_t n_tu = (size_t)((w + 63) / 64 * 64 / 4) * ((h + 63) / 64 * 64 / 4);  // 4x4 blocks of the CTB-padded picture
  const size_t n_ctb_words = (size_t)((w + 15) / 16) * ((h + 15) / 16) * 4;          // enough for any CTB size >= 16
  long ok = 0, flagged = 0;
  for (int seed = seed0; seed < seed0 + n; seed++) {
    std::mt19937 rng(seed);
    const heic_tile_desc& td = img->tiles[rng() % img->n_tiles];
    heic_slice_header sh = td.header;
    std::vector<uint8_t> v(td.rbsp, td.rbsp + td.rbsp_len);
    const uint32_t d0 = sh.slice_data_byte_offset < v.size() ? sh.slice_data_byte_offset : 0;
    const int mode = seed == seed0 ? -1 : (int)(rng() % 7);  // the first one unmutated
    heic_pps pps = img->pps;
    heic_sps sps = img->sps;
    if (mode == 0) {  // a few bit flips in the slice data
      const int k = 1 + rng() % 4;
      for (int i = 0; i < k; i++) v[d0 + rng() % (v.size() - d0)] ^= 1u << (rng() % 8);
    } else if (mode == 1) {  // garbage from some point on
      for (size_t i = d0 + rng() % (v.size() - d0); i < v.size(); i++) v[i] = rng() & 255;
    } else if (mode == 2) {  // truncation
      v.resize(d0 + rng() % (v.size() - d0));
    } else if (mode == 3) {  // all ones / all zeros: longest possible escapes and prefixes
      const uint8_t fill = rng() % 2 ? 0xff : 0x00;
      for (size_t i = d0 + rng() % (v.size() - d0); i < v.size(); i++) v[i] = fill;
    } else if (mode == 4) {  // entry points moved (kept inside the data; the library validates the range on the host)
      for (uint32_t k = 1; k <= sh.num_entry_point_offsets && k <= HEIC_MAX_ENTRY_POINTS; k++)
        if (rng() % 4 == 0) {
          const int64_t moved = (int64_t)sh.substream_offset[k] + (int64_t)(rng() % 9) - 4;
          if (moved > (int64_t)sh.substream_offset[k - 1] && moved < (int64_t)(v.size() - d0)) sh.substream_offset[k] = (uint32_t)moved;
        }
    } else if (mode == 5) {  // slice QP / offsets at their extremes
      sh.slice_qp_delta = (int32_t)(rng() % 104) - 52;
      if (rng() % 2) v[d0 + rng() % (v.size() - d0)] ^= 0x10;
    } else if (mode == 6) {  // coding tools the fixture does not use: the real stream parsed under other flags is garbage
      // that reaches the transform-skip, sign-hiding, cu_qp_delta, strong-smoothing and other-depth paths of the walker
      if (rng() % 2) pps.sign_data_hiding_enabled_flag ^= 1;
      if (rng() % 2) pps.transform_skip_enabled_flag ^= 1;
      if (rng() % 2) { pps.cu_qp_delta_enabled_flag ^= 1; pps.diff_cu_qp_delta_depth = rng() % 3; }
      if (rng() % 2) pps.entropy_coding_sync_enabled_flag ^= 1;
      if (rng() % 4 == 0) sps.max_transform_hierarchy_depth_intra = rng() % 3;
      if (rng() % 4 == 0) sh.slice_sao_luma_flag ^= 1;
      if (rng() % 4 == 0) sh.slice_sao_chroma_flag ^= 1;
    }
    uint8_t* rbsp = (uint8_t*)malloc(v.size() ? v.size() : 1);
    memcpy(rbsp, v.data(), v.size());
    uint32_t* tu = (uint32_t*)malloc(n_tu * 4);
    int16_t* l0 = (int16_t*)malloc(n_tu * 16 * 2);
    int16_t* l1 = (int16_t*)malloc(n_tu * 4 * 2);
    int16_t* l2 = (int16_t*)malloc(n_tu * 4 * 2);
    uint8_t* qp = (uint8_t*)malloc((size_t)(w / 8) * (h / 8) + 1);
    uint32_t* sao = (uint32_t*)calloc(n_ctb_words, 4);
    uint32_t bins = 0, ctus = 0;
    const int rc = emul_parse_picture(&sps, &pps, &sh, rbsp, (uint32_t)v.size(), tu, l0, l1, l2, qp, sao, &bins, &ctus,
                                      (int)(rng() % 2));
    if (mode == -1 && rc != 0) {
      printf("the unmutated tile does not parse (%d)\n", rc);
      return 2;
    }
    (rc == 0 ? ok : flagged)++;
    free(rbsp); free(tu); free(l0); free(l1); free(l2); free(qp); free(sao);
  }
  heic_b200_file_close(f);
  printf("done ok=%ld flagged=%ld\n", ok, flagged);
  return 0;
}

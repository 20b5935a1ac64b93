// TEST INFRASTRUCTURE: AddressSanitizer / UBSan fuzz of the host-side parse layer (container, parameter sets, slice headers)
// through the file API of include/heic_b200.h: mutates a HEIC file (byte flips, truncation, size-field attacks) and walks
// every accessor.  A finding aborts the process; see tests/test_host_fuzz.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>
#include "heic_b200.h"
int main(int argc, char** argv) {
  FILE* fp = fopen(argv[1], "rb");
  std::vector<uint8_t> data;
  { uint8_t buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, fp)) > 0) data.insert(data.end(), buf, buf + n); }
  fclose(fp);
  int seed0 = atoi(argv[2]), n = atoi(argv[3]);
  long ok = 0, err = 0, ps_ok = 0, ps_err = 0;
  // the pristine parameter sets and the first tile's slice NAL, for the direct parser fuzz below
  std::vector<uint8_t> sps_nal, pps_nal, tile_nal;
  heic_sps sps0;
  heic_pps pps0;
  memset(&sps0, 0, sizeof sps0);
  memset(&pps0, 0, sizeof pps0);
  {
    heic_b200_file* f = nullptr;
    if (heic_b200_file_open(data.data(), data.size(), &f) != 0) { printf("fixture does not open\n"); return 2; }
    const uint8_t* p; size_t l;
    if (heic_b200_file_parameter_set_nal(f, -1, 33, &p, &l) == 0) sps_nal.assign(p + 2, p + l);
    if (heic_b200_file_parameter_set_nal(f, -1, 34, &p, &l) == 0) pps_nal.assign(p + 2, p + l);
    if (heic_b200_file_tile_nal(f, -1, 0, &p, &l) == 0) tile_nal.assign(p + 2, p + (l < 66 ? l : 66));
    heic_b200_file_close(f);
    std::vector<uint8_t> u(sps_nal.size() + 1);
    int64_t ul = heic_b200_remove_emulation_prevention(sps_nal.data(), sps_nal.size(), u.data(), nullptr, 0, nullptr);
    if (ul < 0 || heic_b200_parse_sps(u.data(), (size_t)ul, &sps0) != 0) { printf("fixture SPS does not parse\n"); return 2; }
    u.assign(pps_nal.size() + 1, 0);
    ul = heic_b200_remove_emulation_prevention(pps_nal.data(), pps_nal.size(), u.data(), nullptr, 0, nullptr);
    if (ul < 0 || heic_b200_parse_pps(u.data(), (size_t)ul, &pps0) != 0) { printf("fixture PPS does not parse\n"); return 2; }
  }
  auto mutate = [](std::vector<uint8_t> v, std::mt19937& rng) {
    if (v.empty()) return v;
    int k = 1 + rng() % 4;
    for (int i = 0; i < k; i++) {
      if (rng() % 2) v[rng() % v.size()] ^= 1u << (rng() % 8);
      else v[rng() % v.size()] = rng() & 255;
    }
    if (rng() % 4 == 0) v.resize(rng() % (v.size() + 1));
    return v;
  };
  auto heap_copy = [](const std::vector<uint8_t>& v) {  // exact-size heap block: ASan sees any over-read
    uint8_t* h = (uint8_t*)malloc(v.size() ? v.size() : 1);
    memcpy(h, v.data(), v.size());
    return h;
  };
  for (int seed = seed0; seed < seed0 + n; seed++) {
    std::mt19937 rng(seed);
    std::vector<uint8_t> b = data;
    int mode = rng() % 5;
    if (mode == 0) { int k = 1 + rng() % 8; for (int i = 0; i < k; i++) b[rng() % 3700] = rng() & 255; }
    else if (mode == 1) { b.resize(rng() % 10 < 7 ? rng() % 4000 : 4000 + rng() % (b.size() - 4000)); }
    else if (mode == 2) { int k = 1 + rng() % 8; for (int i = 0; i < k; i++) b[3600 + rng() % 4000] = rng() & 255; }
    else if (mode == 3) { size_t p = rng() % 3696; uint32_t vals[] = {0, 1, 7, 8, 0xffffffffu, 0x7fffffffu, 0x80000000u, (uint32_t)rng()};
      uint32_t v = vals[rng() % 8]; b[p] = v >> 24; b[p + 1] = v >> 16; b[p + 2] = v >> 8; b[p + 3] = v; }
    else { int k = 1 + rng() % 3; for (int i = 0; i < k; i++) { size_t p = rng() % 3700; b[p] ^= 1u << (rng() % 8); } }
    // exact-size heap copy so that ASan sees any over-read
    uint8_t* heap = (uint8_t*)malloc(b.size() ? b.size() : 1);
    memcpy(heap, b.data(), b.size());
    heic_b200_file* f = nullptr;
    int rc = heic_b200_file_open(heap, b.size(), &f);
    if (rc == 0 && f) {
      ok++;
      heic_file_info info;
      heic_b200_file_info(f, &info);
      const heic_image_desc* imgs[2] = {heic_b200_file_primary_image(f), heic_b200_file_primary_image_raw(f)};
      unsigned long long sum = 0;
      for (const heic_image_desc* d : imgs) if (d) sum += d->n_tiles;
      uint32_t na = heic_b200_file_aux_image_count(f);
      for (uint32_t i = 0; i < na; i++) { const heic_image_desc* d = heic_b200_file_aux_image(f, i); if (d) sum += d->n_tiles; d = heic_b200_file_aux_image_raw(f, i); if (d) sum += d->n_tiles; }
      for (int image = -1; image < (int)na; image++) {
        for (uint32_t t : {32u, 33u, 34u}) { const uint8_t* p; size_t l; if (heic_b200_file_parameter_set_nal(f, image, t, &p, &l) == 0) for (size_t i = 0; i < l; i++) sum += p[i]; }
        for (uint32_t t = 0; t < 64; t++) { const uint8_t* p; size_t l; if (heic_b200_file_tile_nal(f, image, t, &p, &l) == 0) { for (size_t i = 0; i < l; i += 97) sum += p[i]; if (l) sum += p[l - 1]; } }
      }
      if (sum == 0xdeadbeef) printf("x");
      heic_b200_file_close(f);
    } else err++;
    free(heap);
    // ---- the parameter-set and slice-header parsers directly, on mutated RBSPs ----
    {
      std::vector<uint8_t> v = mutate(sps_nal, rng);
      uint8_t* h = heap_copy(v);
      heic_sps o;
      (heic_b200_parse_sps(h, v.size(), &o) == 0 ? ps_ok : ps_err)++;
      free(h);
      v = mutate(pps_nal, rng);
      h = heap_copy(v);
      heic_pps q;
      (heic_b200_parse_pps(h, v.size(), &q) == 0 ? ps_ok : ps_err)++;
      free(h);
      v = mutate(tile_nal, rng);
      h = heap_copy(v);
      heic_slice_header sh;
      (heic_b200_parse_slice_header_raw(h, v.size(), 19 + rng() % 2, &sps0, &pps0, &sh) == 0 ? ps_ok : ps_err)++;
      std::vector<uint8_t> u(v.size() + 1);
      std::vector<uint32_t> epb(8);
      size_t n_epb = 0;
      int64_t ul = heic_b200_remove_emulation_prevention(h, v.size(), u.data(), epb.data(), epb.size(), &n_epb);
      if (ul >= 0) {
        uint8_t* h2 = (uint8_t*)malloc(ul ? (size_t)ul : 1);
        memcpy(h2, u.data(), (size_t)ul);
        (heic_b200_parse_slice_header(h2, (size_t)ul, 19, &sps0, &pps0, epb.data(), n_epb < epb.size() ? n_epb : epb.size(), &sh) == 0 ? ps_ok : ps_err)++;
        free(h2);
      }
      free(h);
    }
  }
  printf("parsers ok=%ld err=%ld\n", ps_ok, ps_err);
  printf("done ok=%ld err=%ld\n", ok, err);
  return 0;
}

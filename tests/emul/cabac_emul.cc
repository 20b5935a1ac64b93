// TEST INFRASTRUCTURE ONLY — compiles the per-thread device parser (heif_b200/csrc/cuda/cabac_parse.cuh)
// for the host so its syntax logic can be checked against the CPU oracle without a GPU.  Each WPP row
// gets a fresh Parser (as each GPU thread does); rows run one after another, so the wavefront waits are
// trivially satisfied.  Never linked into libheic_b200.so.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../heif_b200/csrc/cuda/cabac_tables.h"
#include "../../heif_b200/csrc/cuda/host_params.h"

using namespace heic;
using namespace heic::dev;

namespace {
struct SeqSync {
  static constexpr int kSaveStride = 1;
  std::vector<uint8_t> save;
  SeqSync() : save(NUM_CTX_PAD) {}
  bool wait(int, int) { return true; }
  void publish(int, int) {}
  uint8_t* save_area(int) { return save.data(); }
  void abort(int) {}
};
}  // namespace

extern "C" int emul_parse_picture(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                                  const uint8_t* rbsp, uint32_t len, uint32_t* tu_map, int16_t* lvl0, int16_t* lvl1,
                                  int16_t* lvl2, uint8_t* qp_out, uint32_t* sao, uint32_t* bins, uint32_t* ctus,
                                  int per_row_threads) {
  try {
    PicParams pp;
    make_pic_params(*sps, *pps, pp);
    heic_tile_desc td;
    std::memset(&td, 0, sizeof td);
    td.rbsp = rbsp;
    td.rbsp_len = len;
    td.header = *sh;
    TileParams tp;
    std::memset(&tp, 0, sizeof tp);
    make_tile_params(pp, *pps, td, tp);
    static CabacTabs tabs;
    build_cabac_tabs(tabs);
    std::vector<uint8_t> ipm((size_t)pp.w4 * pp.h4, 0), ctd((size_t)pp.w8 * pp.h8, 0), qp((size_t)pp.w8 * pp.h8, 0);
    std::vector<uint8_t> ctx(NUM_CTX_PAD);
    std::memset(tu_map, 0, sizeof(uint32_t) * (size_t)pp.n_tu);
    std::memset(lvl0, 0, sizeof(int16_t) * (size_t)pp.n_tu * 16);
    if (pp.chroma) {
      std::memset(lvl1, 0, sizeof(int16_t) * (size_t)pp.n_tu * 4);
      std::memset(lvl2, 0, sizeof(int16_t) * (size_t)pp.n_tu * 4);
    }
    SeqSync sync;
    const int n_slots = (pp.wpp && per_row_threads) ? pp.hctb : 1;
    *bins = *ctus = 0;
    int err = 0;
    for (int slot = 0; slot < n_slots && !err; slot++) {
      Parser<1> P;
      std::memset(&P, 0, sizeof P);
      P.T = &tabs;
      P.ctx = ctx.data();
      P.pp = &pp;
      P.tp = &tp;
      P.tu_map = tu_map;
      P.coeff0 = lvl0;
      P.coeff1 = lvl1;
      P.coeff2 = lvl2;
      P.ipm = ipm.data();
      P.ct_depth = ctd.data();
      P.qp_map = qp.data();
      P.sao = sao;
      P.e.data = rbsp;  // the host build of Engine::load_word reads byte-wise and never past `end`
      *ctus += parse_rows<1>(P, sh->substream_offset, slot, n_slots, sync);
      *bins += P.e.bins;
      err = P.err;
    }
    for (int y = 0; y < pp.h >> 3; y++) std::memcpy(qp_out + (size_t)y * (pp.w >> 3), qp.data() + (size_t)y * pp.w8, (size_t)(pp.w >> 3));
    return err;
  } catch (const Error& e) {
    std::fprintf(stderr, "emul: %s\n", e.what());
    return e.code;
  }
}

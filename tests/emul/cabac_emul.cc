// TEST INFRASTRUCTURE ONLY — compiles the per-thread device parser (heif_b200/csrc/cuda/cabac_parse.cuh)
// for the host so its syntax logic can be checked against the CPU oracle without a GPU.  Each WPP row
// gets a fresh Parser (as each GPU thread does); rows run one after another, so the wavefront waits are
// trivially satisfied.  Never linked into libheic_b200.so.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../tools/experiments/cabac_fsm/cabac_fsm.cuh"
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

// ---- the state-machine walker (cabac_fsm.cuh), one Fsm per row slot of one column, stepped round-robin -----------------
namespace {
// What the threads of one column share on the device (shared memory there): per-slot progress, the abort flag, and the
// 4-entry ring through which the column's threads agree on the tile of every use.
struct ColumnShared {
  uint32_t progress[16] = {};
  uint32_t aborted = 0;  // use + 1 of the tile that failed
  uint32_t slot_use[4] = {}, slot_tile[4] = {}, slot_readers[4] = {};
  const uint32_t* queue = nullptr;
  uint32_t queue_len = 0, queue_pos = 0;
  int n_slots = 1;
  TileStatusDev* status = nullptr;
};
struct FsmEnvHost {
  const CabacTabs* T = nullptr;
  uint8_t ctx[NUM_CTX_PAD] = {};
  uint32_t cold[CW_COUNT] = {};
  const Arenas* A = nullptr;
  ColumnShared* col = nullptr;
  int my_slot = 0;
  uint32_t cur_use = 0;

  const CabacTabs* tabs() const { return T; }
  uint32_t ld_ctx(int idx) const { return ctx[idx]; }
  void st_ctx(int idx, uint32_t v) { ctx[idx] = (uint8_t)v; }
  uint32_t& cw(int j) { return cold[j]; }
  const Arenas* arenas() const { return A; }
  int slot() const { return my_slot; }
  int n_slots() const { return col->n_slots; }
  static uint32_t key(uint32_t use, int row, int n) { return ((use * 1024u + (uint32_t)row) << 10) | (uint32_t)n; }
  uint32_t acquire_tile(uint32_t use) {
    cur_use = use;
    const int k = (int)(use & 3u);
    const uint32_t tag = (use + 1u) << 1;
    if ((col->slot_use[k] & ~1u) == tag) {
      if (!(col->slot_use[k] & 1u)) return TILE_RETRY;
      col->slot_readers[k]++;
      return col->slot_tile[k];
    }
    if (use >= 4 && col->slot_readers[k] < (uint32_t)col->n_slots) return TILE_RETRY;  // a straggler has not read use - 4 yet
    col->slot_readers[k] = 0;
    col->slot_use[k] = tag;
    uint32_t t = TILE_NONE;
    while (col->queue_pos < col->queue_len && (t = col->queue[col->queue_pos++]) == TILE_NONE) {}
    col->slot_tile[k] = t;
    col->slot_use[k] = tag | 1u;
    col->slot_readers[k] = 1;
    return t;
  }
  uint32_t wait_key(int row, int need) const { return key(cur_use, row, need); }
  int wait_ready(uint32_t k) const {
    const int up = my_slot == 0 ? col->n_slots - 1 : my_slot - 1;
    if (col->progress[up] >= k) return 1;
    return col->aborted == cur_use + 1u ? 2 : 0;
  }
  void publish(int row, int n) { col->progress[row % col->n_slots] = key(cur_use, row, n); }
  void abort_tile(int code) {
    if (code != -100 && col->status[cold[CW_TILE]].code == 0) col->status[cold[CW_TILE]].code = code;
    col->aborted = cur_use + 1u;
  }
  void finish_tile(uint32_t tile, uint32_t bins, uint32_t ctus) {
    col->status[tile].bins += bins;
    col->status[tile].ctus += ctus;
  }
};
}  // namespace

// n_slots threads share the rows of the picture (row r -> slot r % n_slots); `repeat` decodes the picture that many times in a
// row through the column's tile ring (exercises the hand-over protocol).  `steps` (optional) receives the iterations of the
// busiest thread.
extern "C" int emul_parse_picture_fsm(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* sh,
                                      const uint8_t* rbsp, uint32_t len, uint32_t* tu_map, int16_t* lvl0, int16_t* lvl1,
                                      int16_t* lvl2, uint8_t* qp_out, uint32_t* sao, uint32_t* bins, uint32_t* ctus, int n_slots,
                                      int repeat, uint64_t* steps) {
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
    if (!pp.wpp) n_slots = 1;
    if (n_slots < 1) n_slots = 1;
    if (n_slots > 16) n_slots = 16;
    std::vector<uint8_t> ipm((size_t)pp.w4 * pp.h4, 0), ctd((size_t)pp.w8 * pp.h8, 0), qp((size_t)pp.w8 * pp.h8, 0);
    std::vector<uint8_t> wpp_save((size_t)pp.hctb * NUM_CTX_PAD, 0);
    std::vector<int16_t> coeff((size_t)pp.n_tu * 24, 0);
    std::vector<uint32_t> subs(sh->substream_offset, sh->substream_offset + tp.n_sub);
    std::memset(tu_map, 0, sizeof(uint32_t) * (size_t)pp.n_tu);
    std::memset(sao, 0, sizeof(uint32_t) * (size_t)pp.wctb * pp.hctb * 4);
    tp.coeff_off[0] = 0;
    tp.coeff_off[1] = (uint64_t)pp.n_tu * 16;
    tp.coeff_off[2] = (uint64_t)pp.n_tu * 20;
    TileStatusDev status;
    std::memset(&status, 0, sizeof status);
    Arenas A;
    std::memset(&A, 0, sizeof A);
    A.bitstream = rbsp;
    A.substreams = subs.data();
    A.pics = &pp;
    A.tiles = &tp;
    A.tu_map = tu_map;
    A.coeff = coeff.data();
    A.ipm = ipm.data();
    A.ct_depth = ctd.data();
    A.qp_map = qp.data();
    A.sao = sao;
    A.wpp_save = wpp_save.data();
    A.status = &status;
    A.n_tiles = 1;
    std::vector<uint32_t> queue((size_t)(repeat < 1 ? 1 : repeat), 0u);
    ColumnShared col;
    col.queue = queue.data();
    col.queue_len = (uint32_t)queue.size();
    col.n_slots = n_slots;
    col.status = &status;
    std::vector<Fsm<FsmEnvHost>> th((size_t)n_slots);
    std::vector<char> alive((size_t)n_slots, 1);
    std::vector<uint64_t> n_steps((size_t)n_slots, 0);
    for (int s = 0; s < n_slots; s++) {
      th[s].env.T = &tabs;
      th[s].env.A = &A;
      th[s].env.col = &col;
      th[s].env.my_slot = s;
      th[s].init();
    }
    // the arenas are decoded into `repeat` times; the last pass is what the caller compares (every pass must clear first)
    int live = n_slots;
    uint64_t guard = 0;
    uint32_t last_use_seen = 0;
    while (live > 0) {
      for (int s = n_slots - 1; s >= 0; s--) {  // bottom rows first: they meet unmet dependencies and must idle
        if (!alive[s]) continue;
        // a new use begins: clear the outputs the way the pipeline's memsets do
        if (th[s].st == S_TILE_NEXT && th[s].env.cold[CW_USE] > last_use_seen && th[s].env.cold[CW_USE] < (uint32_t)queue.size()) {
          bool all = true;
          for (int q = 0; q < n_slots; q++) all = all && (!alive[q] || (th[q].st == S_TILE_NEXT && th[q].env.cold[CW_USE] == th[s].env.cold[CW_USE]));
          if (!all) continue;  // let the others finish the previous pass before its outputs are cleared
          last_use_seen = th[s].env.cold[CW_USE];
          std::memset(tu_map, 0, sizeof(uint32_t) * (size_t)pp.n_tu);
          std::fill(coeff.begin(), coeff.end(), (int16_t)0);
          std::memset(sao, 0, sizeof(uint32_t) * (size_t)pp.wctb * pp.hctb * 4);
          std::memset(&status, 0, sizeof status);
        }
        if (!th[s].step()) {
          alive[s] = 0;
          live--;
        }
        n_steps[s]++;
        if (++guard > (1ull << 34)) return -99;  // a hang in the protocol
      }
    }
    std::memcpy(lvl0, coeff.data(), sizeof(int16_t) * (size_t)pp.n_tu * 16);
    if (pp.chroma) {
      std::memcpy(lvl1, coeff.data() + (size_t)pp.n_tu * 16, sizeof(int16_t) * (size_t)pp.n_tu * 4);
      std::memcpy(lvl2, coeff.data() + (size_t)pp.n_tu * 20, sizeof(int16_t) * (size_t)pp.n_tu * 4);
    }
    *bins = status.bins;
    *ctus = status.ctus;
    if (steps) {
      *steps = 0;
      for (uint64_t v : n_steps) *steps = v > *steps ? v : *steps;
    }
    for (int y = 0; y < pp.h >> 3; y++) std::memcpy(qp_out + (size_t)y * (pp.w >> 3), qp.data() + (size_t)y * pp.w8, (size_t)(pp.w >> 3));
    return status.code;
  } catch (const Error& e) {
    std::fprintf(stderr, "emul: %s\n", e.what());
    return e.code;
  }
}

// ---- SIMT model of one CTA of the state-machine kernel (tools/fsm_model.py): `n_cols` columns x `n_slots` row slots decode
//      the pictures of `queue` (indices into the picture arrays) exactly as cabac_fsm_kernel.cu schedules them -- static first
//      entries, then a shared counter -- with every warp stepped once per round.  Reports, per warp-iteration, how many lanes
//      did useful work and how many different states (switch bodies) the warp had to run. ----
extern "C" int emul_fsm_simulate(const heic_sps* sps, const heic_pps* pps, const heic_slice_header* const* headers,
                                 const uint8_t* const* rbsps, const uint32_t* lens, uint32_t n_pics, const uint32_t* queue,
                                 uint32_t queue_len, int n_cols, int n_slots, uint64_t* out /* [8 + S_COUNT * 2] */) {
  try {
    PicParams pp;
    make_pic_params(*sps, *pps, pp);
    static CabacTabs tabs;
    build_cabac_tabs(tabs);
    std::vector<TileParams> tps(n_pics);
    std::vector<uint32_t> subs;
    std::vector<uint8_t> bitstream;
    for (uint32_t p = 0; p < n_pics; p++) {
      heic_tile_desc td;
      std::memset(&td, 0, sizeof td);
      td.rbsp = rbsps[p];
      td.rbsp_len = lens[p];
      td.header = *headers[p];
      std::memset(&tps[p], 0, sizeof(TileParams));
      make_tile_params(pp, *pps, td, tps[p]);
      tps[p].sub_first = (uint32_t)subs.size();
      for (uint32_t k = 0; k < tps[p].n_sub; k++) subs.push_back(headers[p]->substream_offset[k]);
      tps[p].bs_off = (uint32_t)bitstream.size();
      bitstream.insert(bitstream.end(), rbsps[p], rbsps[p] + lens[p]);
      bitstream.resize((bitstream.size() + 31) & ~(size_t)15, 0);
    }
    // one scratch output arena per COLUMN (the model does not check outputs; columns must not share neighbour maps)
    const size_t n_tu = (size_t)pp.n_tu, m4 = (size_t)pp.w4 * pp.h4, m8 = (size_t)pp.w8 * pp.h8;
    std::vector<uint32_t> tu(n_tu * n_cols), sao((size_t)pp.wctb * pp.hctb * 4 * n_cols);
    std::vector<int16_t> coeff(n_tu * 24 * n_cols);
    std::vector<uint8_t> ipm(m4 * n_cols), ctd(m8 * n_cols), qp(m8 * n_cols), save((size_t)pp.hctb * NUM_CTX_PAD * n_cols);
    // every (column, picture) pair needs its own TileParams (offsets into the column's arena)
    std::vector<TileParams> col_tps((size_t)n_cols * n_pics);
    for (int c = 0; c < n_cols; c++)
      for (uint32_t p = 0; p < n_pics; p++) {
        TileParams t = tps[p];
        t.tu_off = n_tu * c;
        t.coeff_off[0] = n_tu * 24 * c;
        t.coeff_off[1] = t.coeff_off[0] + n_tu * 16;
        t.coeff_off[2] = t.coeff_off[0] + n_tu * 20;
        t.map4_off = m4 * c;
        t.map8_off = m8 * c;
        t.sao_off = (size_t)pp.wctb * pp.hctb * 4 * c;
        t.wpp_off = (size_t)pp.hctb * NUM_CTX_PAD * c;
        col_tps[(size_t)c * n_pics + p] = t;
      }
    std::vector<TileStatusDev> status((size_t)n_cols * n_pics);
    std::vector<Arenas> arenas((size_t)n_cols);
    std::vector<ColumnShared> cols((size_t)n_cols);
    // the queue of every column: the shared counter is emulated by handing each column the entries in claim order
    std::vector<std::vector<uint32_t>> col_queue((size_t)n_cols);
    uint32_t next = 0;
    for (int c = 0; c < n_cols && next < queue_len; c++) col_queue[c].push_back(queue[next++]);
    for (int c = 0; c < n_cols; c++) {
      Arenas& A = arenas[c];
      std::memset(&A, 0, sizeof A);
      A.bitstream = bitstream.data();
      A.substreams = subs.data();
      A.pics = &pp;
      A.tiles = col_tps.data() + (size_t)c * n_pics;
      A.tu_map = tu.data();
      A.coeff = coeff.data();
      A.ipm = ipm.data();
      A.ct_depth = ctd.data();
      A.qp_map = qp.data();
      A.sao = sao.data();
      A.wpp_save = save.data();
      A.status = status.data() + (size_t)c * n_pics;
      cols[c].n_slots = n_slots;
      cols[c].status = A.status;
    }
    std::vector<Fsm<FsmEnvHost>> th((size_t)n_cols * n_slots);
    std::vector<char> alive(th.size(), 1);
    for (int s = 0; s < n_slots; s++)
      for (int c = 0; c < n_cols; c++) {
        Fsm<FsmEnvHost>& F = th[(size_t)s * n_cols + c];
        F.env.T = &tabs;
        F.env.A = &arenas[c];
        F.env.col = &cols[c];
        F.env.my_slot = s;
        F.init();
      }
    uint64_t warp_iters = 0, lane_busy = 0, lane_wait = 0, lane_dead = 0, states_sum = 0, lane_engine = 0;
    std::vector<uint64_t> st_present(S_COUNT, 0), st_lanes(S_COUNT, 0);
    size_t live = th.size();
    while (live > 0) {
      for (int s = 0; s < n_slots; s++) {  // one round: every warp one iteration
        bool any = false;
        uint64_t seen = 0;
        for (int c = 0; c < n_cols; c++) {
          const size_t id = (size_t)s * n_cols + c;
          if (!alive[id]) {
            lane_dead++;
            continue;
          }
          any = true;
          Fsm<FsmEnvHost>& F = th[id];
          // dynamic claims: feed the column's private queue from the shared one when it asks for a new use
          ColumnShared& col = cols[c];
          if (F.st == S_TILE_NEXT && col.queue_pos >= col_queue[c].size() && F.env.cold[CW_USE] >= col_queue[c].size() && next < queue_len)
            col_queue[c].push_back(queue[next++]);
          col.queue = col_queue[c].data();
          col.queue_len = (uint32_t)col_queue[c].size();
          const uint32_t st_before = F.st;
          if (F.op != OP_NONE) lane_engine++;
          if (st_before == S_WAIT || (st_before == S_TILE_NEXT && F.op == OP_NONE && false)) lane_wait++;
          else lane_busy++;
          seen |= 1ull << st_before;
          st_lanes[st_before]++;
          if (!F.step()) {
            alive[id] = 0;
            live--;
          }
        }
        if (any) {
          warp_iters++;
          states_sum += (uint64_t)__builtin_popcountll(seen);
          for (uint32_t k = 0; k < S_COUNT; k++)
            if ((seen >> k) & 1u) st_present[k]++;
        }
      }
    }
    uint64_t bins = 0;
    for (const TileStatusDev& t : status) bins += t.bins;
    out[0] = warp_iters;
    out[1] = lane_busy;
    out[2] = lane_wait;
    out[3] = lane_dead;
    out[4] = states_sum;
    out[5] = bins;
    out[6] = lane_engine;
    out[7] = S_COUNT;
    for (uint32_t k = 0; k < S_COUNT; k++) {
      out[8 + 2 * k] = st_present[k];
      out[9 + 2 * k] = st_lanes[k];
    }
    return 0;
  } catch (const Error& e) {
    std::fprintf(stderr, "emul: %s\n", e.what());
    return e.code;
  }
}

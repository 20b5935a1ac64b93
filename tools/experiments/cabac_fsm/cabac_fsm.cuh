// slice_segment_data() as a per-thread STATE MACHINE whose iteration is one arithmetic-decoder operation.
//
// The nested walker of cabac_parse.cuh is what a CPU would write: loops inside loops, a decoder call wherever the syntax
// needs a bin.  With 32 different pictures in the 32 lanes of a warp that form serialises: lanes that sit in different
// branches take turns, every loop level costs its slowest lane, and all rows of a warp wait for the slowest tile's
// wavefront.  Here every thread instead carries its position in the syntax as a state, and one iteration of the loop is
//
//     1. ONE engine operation for every lane (a context-coded bin, up to seven bypass bins, a coeff_abs_level_remaining,
//        a truncated-unary bypass value, or the terminate bin) -- the same few dozen instructions for all 32 lanes,
//        whatever syntax element each of them is in;
//     2. a `switch` on the lane's state that consumes the result and prepares the next request.
//
// Lanes never wait for each other: a lane whose WPP dependency (the row above two CTUs ahead, 9.3.2.2) is not met yet
// simply idles in its state while the others go on, and a thread that has finished its rows of a tile fetches the
// column's next tile on its own (cabac_fsm_kernel.cu).  Only the switch bodies of the states that are present in a warp at
// the same time serialise, and each of them is short by construction: nothing in a state loops over data (the only
// loops are the rare per-row context copies and per-CU map fills).
//
// Bins are consumed in exactly the order of the nested walker (tests/emul runs both against the oracle).
// Reference interfaces replaced: the same as cabac_parse.cuh (src/cabac/*, src/hevc/slice.rs:206-255).
#pragma once
#include "../../../heif_b200/csrc/cuda/cabac_parse.cuh"

namespace heic {
namespace dev {

enum : uint32_t { OP_NONE = 0, OP_DEC = 1, OP_BYP = 2, OP_CALR = 3, OP_TR = 4, OP_TERM = 5 };

// result markers of OP_CALR: the prefix ran into the Exp-Golomb escape / one suffix bit is still missing
enum : uint32_t { CALR_ESCAPE = 0x80000000u, CALR_ONE_MORE = 0x40000000u };

enum : uint32_t { TILE_NONE = 0xffffffffu, TILE_RETRY = 0xfffffffeu };

// Cold per-thread state (touched by the header states only) lives in words the environment provides: shared memory on the
// device (one word column per thread), a plain array on the host.
enum {
  CW_TILE = 0,   // current tile
  CW_USE,        // tiles this column has taken so far
  CW_POS,        // rx | ry << 10 | z << 20 (coding quadtree position, minimum-CB units) | cq_log2 << 27
  CW_CTUS,       // CTUs decoded in the current tile
  CW_QP,         // qp_y | last_qp_y << 6 | qp_y_pred << 12 | (cu_qp_delta_val + 64) << 18 | coded << 25 | first_qg_in_row << 26
  CW_QG,         // qg_x | qg_y << 16 (0xffff: none yet)
  CW_CU,         // cu_x | cu_y << 14 | cu_log2 << 28 | part_nxn << 31
  CW_PU,         // IntraPredModeY of the four prediction blocks
  CW_CU2,        // chroma_mode | prev flags << 6 | k << 10 | n_pu << 13 | comp << 16 (component of the residual block)
  CW_TT,         // tz | cb_mask << 9 | cr_mask << 13 | tt_log2 << 17 | split << 20 | cbf_luma << 21 | cbf_cb << 22 | cbf_cr << 23 |
                 // has_chroma << 24 | ts << 25 (transform_skip_flag x 3) | cu_qp_delta prefix << 29
  CW_EG,         // Exp-Golomb helper: ones | k << 6 | left << 10 | ret << 16
  CW_EGV,        // its value so far
  CW_SAO,        // abs4 | i << 16 | c << 19 | type_c << 21 | class_c << 23 | type << 25 | merge_left << 27
  CW_SAOV,       // parameter word under construction
  CW_TP_LO, CW_TP_HI, CW_PP_LO, CW_PP_HI,  // TileParams* / PicParams* of the current tile
  CW_COUNT
};

enum FsmState : uint32_t {
  S_TILE_NEXT = 0, S_CTU_BEGIN, S_WAIT, S_SAO_ML, S_SAO_MU, S_SAO_T1, S_SAO_T2, S_SAO_OFF, S_SAO_SIGN, S_SAO_BAND, S_SAO_CLASS,
  S_CQT_SPLIT, S_PART, S_PREV, S_PU_MODE, S_CHROMA1, S_CHROMA2, S_TT_SPLIT, S_TT_CB, S_TT_CR, S_TT_LUMA, S_DQP_PREFIX,
  S_DQP_SIGN, S_DQP_EG, S_EG_PRE, S_EG_SUF, S_RC_TSKIP, S_RC_LAST_PRE, S_RC_LAST_SUF, S_CSBF, S_SIG, S_DC, S_GT1, S_GT2, S_SIGN,
  S_LEVEL, S_LEVEL_ESC, S_LEVEL_ONE, S_EOS, S_EOSUB, S_COUNT
};

// Env: see FsmEnvHost (tests/emul) and the device environment in cabac_fsm_kernel.cu.  It provides
//   const CabacTabs* tabs();  uint32_t ld_ctx(int idx);  void st_ctx(int idx, uint32_t v);
//   uint32_t& cw(int j);                         cold word j of this thread
//   const Arenas* arenas();
//   int slot();  int n_slots();                  this thread's row slot, row slots per column
//   uint32_t acquire_tile(uint32_t use);         tile index, TILE_RETRY (try again next iteration) or TILE_NONE (queue empty)
//   uint32_t wait_key(int row, int need);        what `row` (the one above this thread's) must have published ...
//   int wait_ready(uint32_t key);                ... 0: not yet, 1: it has, 2: the tile was aborted
//   void publish(int row, int n_ctus);
//   void abort_tile(int code);                   records the failure (code -100: silently) and releases waiting rows
//   void finish_tile(uint32_t tile, uint32_t bins, uint32_t ctus);
template <class Env>
struct Fsm {
  Engine e;
  Env env;
  uint32_t st, op, arg;
  // ---- hot state: the residual block / sub-block being decoded ----
  int16_t* out;       // TransCoeffLevel of the current block (raster n x n)
  uint64_t csbf;      // coded_sub_block_flag, bit ys * 8 + xs
  uint64_t nib;       // sigCtx of the 16 scan positions of the sub-block
  uint32_t rcw;       // block constants: log2 | c_idx << 3 | scan_idx << 5 | last_sub_block << 7 | last_scan_pos << 13 |
                      //   prev_g1_zero << 17 (greater1Ctx of the previous sub-block ended at 0) | first_sub_block << 18 |
                      //   tskip << 19 | pred_mode << 20 (until the scan is chosen; then bit 20 = sign_data_hiding_enabled) |
                      //   sig_off << 26 (6 bits: the block's sigCtx offset)
  uint32_t sbw;       // sub-block: i (signed 8 bits) | xs << 8 | ys << 11 | infer_dc << 14 | k << 15 | add << 19 (8 bits) |
                      //   ctx_set << 27 | greater1_ctx << 29
  uint32_t sig_g1;    // sig | g1 << 16
  uint32_t m;         // coefficients still to visit in the greater1 / level pass
  uint32_t lvw;       // level pass: rice | num_sig << 3 | sum_parity << 8 | (last_g1_pos + 1) << 9 | g2 << 14 | sign_hidden << 15 |
                      //   first_sig << 16 | cur_k << 20 | base << 24 | num_gt1 << 27
  uint32_t sign_bits; // coeff_sign_flags of the sub-block, next one in bit 31 (also scratch of the last-position prefixes)

  HEIC_HD const PicParams* pp() { return reinterpret_cast<const PicParams*>(((uint64_t)env.cw(CW_PP_HI) << 32) | env.cw(CW_PP_LO)); }
  HEIC_HD const TileParams* tp() { return reinterpret_cast<const TileParams*>(((uint64_t)env.cw(CW_TP_HI) << 32) | env.cw(CW_TP_LO)); }
  HEIC_HD static uint32_t bits(uint32_t w, int lo, int n) { return (w >> lo) & ((1u << n) - 1u); }
  HEIC_HD static void put(uint32_t& w, int lo, int n, uint32_t v) { w = (w & ~(((1u << n) - 1u) << lo)) | ((v & ((1u << n) - 1u)) << lo); }

  HEIC_HD void init() {
    st = S_TILE_NEXT;
    op = OP_NONE;
    arg = 0;
    out = nullptr;
    csbf = nib = 0;
    rcw = sbw = sig_g1 = m = lvw = sign_bits = 0;
    e.bins = 0;
    e.range = 510;
    e.val = 0;
    e.nbits = 0;
    e.pos = e.end = 0;
    e.look0 = e.look1 = 0;
    e.data = nullptr;
    for (int j = 0; j < CW_COUNT; j++) env.cw(j) = 0;
  }

  // ---- scan helpers (6.5.3-6.5.5), as in Parser ----
  HEIC_HD uint32_t scan_xy(int scan_idx, int lg, int i) {
    if (lg == 0) return 0;
    if (scan_idx == 0) {
      uint32_t v;
      if (lg == 2) {
        v = env.tabs()->diag4[i];
        return (v & 3u) | ((v >> 2) << 4);
      }
      if (lg == 3) {
        v = env.tabs()->diag8[i];
        return (v & 7u) | ((v >> 3) << 4);
      }
      v = env.tabs()->diag2[i];
      return (v & 1u) | ((v >> 1) << 4);
    }
    uint32_t a = (uint32_t)i & ((1u << lg) - 1u), b = (uint32_t)i >> lg;
    return scan_idx == 1 ? (a | (b << 4)) : (b | (a << 4));
  }
  HEIC_HD int scan_inv(int scan_idx, int lg, int x, int y) {
    if (lg == 0) return 0;
    if (scan_idx == 0) {
      if (lg == 2) return env.tabs()->inv_diag4[(y << 2) | x];
      if (lg == 3) return env.tabs()->inv_diag8[(y << 3) | x];
      return env.tabs()->inv_diag2[(y << 1) | x];
    }
    return scan_idx == 1 ? ((y << lg) | x) : ((x << lg) | y);
  }
  // context of bin v of last_sig_coeff_{x,y}_prefix (decoder.rs:109-130)
  HEIC_HD static int last_ctx(int d, int c_idx, int log2, int v) {
    int ctx_offset, ctx_shift;
    if (c_idx == 0) {
      ctx_offset = 3 * (log2 - 2) + ((log2 - 1) >> 2);
      ctx_shift = (log2 + 1) >> 2;
    } else {
      ctx_offset = 15;
      ctx_shift = log2 - 2;
    }
    return (d ? CTX_LAST_Y : CTX_LAST_X) + (v >> ctx_shift) + ctx_offset;
  }
  // 8.4.2 (as Parser::derive_luma_mode)
  HEIC_HD int derive_luma_mode(const PicParams* P, const uint8_t* ipm, int x, int y, int prev_flag, int mpm_idx, int rem) {
    int cand_a = 1, cand_b = 1;
    if (x > 0) cand_a = ipm[(y >> 2) * P->w4 + ((x - 1) >> 2)];
    if (y > 0 && ((y - 1) >> P->log2_ctb) == (y >> P->log2_ctb)) cand_b = ipm[((y - 1) >> 2) * P->w4 + (x >> 2)];
    int c0, c1, c2;
    if (cand_a == cand_b) {
      if (cand_a < 2) {
        c0 = 0;
        c1 = 1;
        c2 = 26;
      } else {
        c0 = cand_a;
        c1 = 2 + ((cand_a + 29) & 31);
        c2 = 2 + ((cand_a - 2 + 1) & 31);
      }
    } else {
      c0 = cand_a;
      c1 = cand_b;
      if (cand_a != 0 && cand_b != 0) c2 = 0;
      else if (cand_a != 1 && cand_b != 1) c2 = 1;
      else c2 = 26;
    }
    if (prev_flag) return mpm_idx == 0 ? c0 : (mpm_idx == 1 ? c1 : c2);
    int t;
    if (c0 > c1) t = c0, c0 = c1, c1 = t;
    if (c0 > c2) t = c0, c0 = c2, c2 = t;
    if (c1 > c2) t = c1, c1 = c2, c2 = t;
    int mode = rem;
    if (mode >= c0) mode++;
    if (mode >= c1) mode++;
    if (mode >= c2) mode++;
    return mode;
  }

  // ---- the engine operation of this iteration (9.3.4.3; arithmetic.rs:97-169) ----
  HEIC_HD uint32_t engine_op() {
    uint32_t res = 0;
    if (op == OP_DEC) {
      uint32_t s = env.ld_ctx((int)arg);
      res = (uint32_t)e.decision(env.tabs(), s);
      env.st_ctx((int)arg, s);
    } else if (op == OP_BYP) {
      res = e.bypass_bins((int)arg);
    } else if (op == OP_CALR) {
      // coeff_abs_level_remaining (decoder.rs:224-261): prefix (up to four ones and the closing zero) and, when they fit,
      // the rice suffix bits from one division; the Exp-Golomb escape and the one case where a suffix bit does not fit
      // are reported to the state machine
      const int rice = (int)arg;
      const int k = 4 + (rice < 3 ? rice : 3);
      const uint32_t X = e.val >> (22 - k), Q = e.bypass_quot(X);
      const uint32_t top4 = Q >> (k - 4);
      const int prefix = HEIC_CLZ(~(top4 << 28));
      if (prefix >= 4) {
        e.bypass_take(X, Q, k, 4);
        res = CALR_ESCAPE;
      } else {
        const int j = prefix + 1 + rice;
        if (j <= k) res = ((uint32_t)prefix << rice) + (e.bypass_take(X, Q, k, j) & ((1u << rice) - 1u));
        else res = CALR_ONE_MORE | (((uint32_t)prefix << rice) + ((e.bypass_take(X, Q, k, k) & 7u) << 1));
      }
    } else if (op == OP_TR) {  // truncated unary, bypass coded, cMax = arg <= 7 (decoder.rs:166-190 with cRiceParam 0)
      const int k = (int)arg;
      const uint32_t X = e.val >> (22 - k), Q = e.bypass_quot(X);
      const int ones = HEIC_CLZ(~(Q << (32 - k)));  // leading ones of the k-bin quotient (Q has k bits)
      res = (uint32_t)(ones < k ? ones : k);
      e.bypass_take(X, Q, k, ones < k ? ones + 1 : k);
    } else if (op == OP_TERM) {
      res = (uint32_t)e.terminate();
    }
    return res;
  }

  HEIC_HD void tile_failed(int code) {
    env.abort_tile(code);
  }

  // One iteration.  Returns false when this thread has nothing left to do.
  HEIC_HD bool step() {
    const uint32_t res = engine_op();
    // scratch (declared up front: the state bodies are entered by `switch` and by `goto`)
    const PicParams* P;
    const TileParams* T;
    const Arenas* A;
    uint32_t w, pos, qpw, tt, cu, cu2, sao, v, sxy;
    int rx, ry, z, log2, x0, y0, n, depth, inc, lvl, tzb, k, i, c, xs, ys, right, below, coded, n_start, prev_csbf, add, ctx0,
        ctx_set, g1ctx, kk, num, last_g1, n_sign, take, got, base, abs_level, rice, num_sig, first_sig, last_sig, sign_hidden,
        scan_idx, c_idx, lg_sb, sb_w, last_sb, last_pos, pv, nb, d, full, last_x, last_y, type, cl, a, o, mode, pb, n_pu, px, py,
        b4, x4, y4, j, idx, luma, mm, qa, qb, qp_prev, ctb_mask, mask, x_qg, y_qg, qp_y, need, ready, wctb, hctb, addr, ones, left,
        cbf_cb, cbf_cr, cbf_luma, par_cb, par_cr, split, max_depth, intra_split, tlog2, has_chroma, comp, cbf, pm, log2c, dv, neg,
        pu_idx, pb_shift, b8, x8, y8, yy, xx, ctb4, any_cbf;
    uint64_t bit;
    uint32_t* sp;
    uint8_t* bp;
    const uint8_t* cbp;

    switch (st) {
      // =====================================================================================================
      // tile / row / CTU level (parse_rows of cabac_parse.cuh)
      // =====================================================================================================
      case S_TILE_NEXT: {
        w = env.acquire_tile(env.cw(CW_USE));
        if (w == TILE_NONE) return false;
        if (w == TILE_RETRY) goto idle;
        A = env.arenas();
        T = A->tiles + w;
        P = A->pics + T->pic;
        env.cw(CW_TILE) = w;
        env.cw(CW_TP_LO) = (uint32_t)(uintptr_t)T;
        env.cw(CW_TP_HI) = (uint32_t)((uint64_t)(uintptr_t)T >> 32);
        env.cw(CW_PP_LO) = (uint32_t)(uintptr_t)P;
        env.cw(CW_PP_HI) = (uint32_t)((uint64_t)(uintptr_t)P >> 32);
        env.cw(CW_CTUS) = 0;
        e.data = A->bitstream + T->bs_off;
        e.bins = 0;
        qp_y = T->slice_qp;
        env.cw(CW_QP) = (uint32_t)qp_y | ((uint32_t)qp_y << 6) | ((uint32_t)qp_y << 12) | (64u << 18) | (1u << 26);
        env.cw(CW_QG) = 0xffffffffu;
        ry = env.slot();
        if (ry >= P->hctb) goto tile_done;  // fewer CTB rows than row slots: nothing for this thread in this tile
        env.cw(CW_POS) = (uint32_t)ry << 10;
        goto ctu_begin;
      }
      case S_CTU_BEGIN:
      ctu_begin: {
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        wctb = P->wctb;
        if (P->wpp && ry > 0) {
          need = rx == 0 ? (wctb < 2 ? wctb : 2) : rx + 1;
          m = env.wait_key(ry - 1, need);  // (the residual registers are free between CTUs)
          ready = env.wait_ready(m);
          if (ready == 2) goto tile_done;  // another row of this tile failed and recorded its code
          if (ready == 0) {
            st = S_WAIT;
            goto idle;
          }
        }
        goto ctu_go;
      }
      case S_WAIT:  // the row above has not finished the CTUs this one depends on (9.3.2.2: two ahead); a few instructions per poll
        ready = env.wait_ready(m);
        if (ready == 0) goto idle;
        if (ready == 2) goto tile_done;
      ctu_go: {
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        wctb = P->wctb;
        if (rx == 0 && (ry == 0 || P->wpp)) {
          A = env.arenas();
          const uint32_t* sub = A->substreams + T->sub_first;
          const uint32_t start = T->data_off + (P->wpp ? sub[ry] : 0u);
          const uint32_t stop = (P->wpp && ry + 1 < (int)T->n_sub) ? T->data_off + sub[ry + 1] : T->bs_len;
          e.init(e.data, start, stop);
          if (e.offset_is_illegal()) goto fail3;  // arithmetic.rs:33-36
          if (ry == 0 || wctb == 1) {
            const int slice_qp = T->slice_qp;
            for (j = 0; j < NUM_CTX; j++) env.st_ctx(j, context_init_state(env.tabs()->init_value[j], slice_qp));
          } else {
            cbp = A->wpp_save + T->wpp_off + (size_t)(ry - 1) * NUM_CTX_PAD;
            for (j = 0; j < NUM_CTX; j++) env.st_ctx(j, cbp[j]);
          }
          env.cw(CW_QP) |= 1u << 26;  // first_qg_in_row
        }
        // ---- coding_tree_unit(rx, ry) ----
        if (!P->cu_qp_delta_enabled) put(env.cw(CW_QP), 0, 6, (uint32_t)T->slice_qp);
        if (T->sao_luma || T->sao_chroma) {
          // 7.3.8.3 sao() (todo!() at slice.rs:249-251)
          env.cw(CW_SAO) = 0;
          if (rx > 0) {
            op = OP_DEC, arg = CTX_SAO_MERGE, st = S_SAO_ML;
            return true;
          }
          goto sao_after_ml;
        }
        goto ctu_body;
      }
      case S_SAO_ML:
        if (res) env.cw(CW_SAO) |= 1u << 27;
      sao_after_ml:
        ry = (int)bits(env.cw(CW_POS), 10, 10);
        if (ry > 0 && !(env.cw(CW_SAO) >> 27)) {
          op = OP_DEC, arg = CTX_SAO_MERGE, st = S_SAO_MU;
          return true;
        }
        goto sao_after_mu0;
      case S_SAO_MU:
        if (res) env.cw(CW_SAO) |= 1u << 28;
      sao_after_mu0: {
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        sp = env.arenas()->sao + T->sao_off + (size_t)(ry * P->wctb + rx) * 4;
        sao = env.cw(CW_SAO);
        if (sao >> 27) {
          const uint32_t* q = (sao >> 27) & 1u ? sp - 4 : sp - (size_t)P->wctb * 4;
          sp[0] = q[0];
          sp[1] = q[1];
          sp[2] = q[2];
          sp[3] = 0;
          goto ctu_body;
        }
        sp[0] = sp[1] = sp[2] = sp[3] = 0;
        goto sao_comp;
      }
      sao_comp: {  // next colour component (the state word's c)
        P = pp();
        T = tp();
        sao = env.cw(CW_SAO);
        c = (int)bits(sao, 19, 2);
        if (c >= (P->chroma ? 3 : 1)) goto ctu_body;
        if (!((c == 0 && T->sao_luma) || (c > 0 && T->sao_chroma))) {
          put(env.cw(CW_SAO), 19, 2, (uint32_t)c + 1);
          goto sao_comp;
        }
        if (c == 2) {
          put(env.cw(CW_SAO), 25, 2, bits(sao, 21, 2));
          goto sao_have_type;
        }
        op = OP_DEC, arg = CTX_SAO_TYPE, st = S_SAO_T1;  // sao_type_idx: TR cMax 2, bin 0 context coded, bin 1 bypass
        return true;
      }
      case S_SAO_T1:
        if (res) {
          op = OP_BYP, arg = 1, st = S_SAO_T2;
          return true;
        }
        put(env.cw(CW_SAO), 25, 2, 0);
        if (bits(env.cw(CW_SAO), 19, 2) == 1) put(env.cw(CW_SAO), 21, 2, 0);
        goto sao_have_type;
      case S_SAO_T2:
        put(env.cw(CW_SAO), 25, 2, res ? 2u : 1u);
        if (bits(env.cw(CW_SAO), 19, 2) == 1) put(env.cw(CW_SAO), 21, 2, res ? 2u : 1u);
      sao_have_type:
        sao = env.cw(CW_SAO);
        if (!bits(sao, 25, 2)) {
          put(env.cw(CW_SAO), 19, 2, bits(sao, 19, 2) + 1);
          goto sao_comp;
        }
        put(env.cw(CW_SAO), 0, 19, 0);  // abs4 = 0, i = 0
        op = OP_TR, arg = 7, st = S_SAO_OFF;
        return true;
      case S_SAO_OFF: {
        sao = env.cw(CW_SAO);
        i = (int)bits(sao, 16, 3);
        sao |= res << (4 * i);
        i++;
        put(sao, 16, 3, (uint32_t)i);
        env.cw(CW_SAO) = sao;
        if (i < 4) {
          op = OP_TR, arg = 7, st = S_SAO_OFF;
          return true;
        }
        type = (int)bits(sao, 25, 2);
        env.cw(CW_SAOV) = (uint32_t)type;
        if (type == 1) {
          put(env.cw(CW_SAO), 16, 3, 0);
          goto sao_sign;
        }
        c = (int)bits(sao, 19, 2);
        if (c < 2) {
          op = OP_BYP, arg = 2, st = S_SAO_CLASS;
          return true;
        }
        cl = (int)bits(sao, 23, 2);
        goto sao_eo;
      }
      sao_sign: {  // band offset: a sign for every non-zero offset, then the band position
        sao = env.cw(CW_SAO);
        i = (int)bits(sao, 16, 3);
        while (i < 4 && ((sao >> (4 * i)) & 15u) == 0) i++;
        put(env.cw(CW_SAO), 16, 3, (uint32_t)i);
        if (i < 4) {
          op = OP_BYP, arg = 1, st = S_SAO_SIGN;
        } else {
          op = OP_BYP, arg = 5, st = S_SAO_BAND;
        }
        return true;
      }
      case S_SAO_SIGN: {
        sao = env.cw(CW_SAO);
        i = (int)bits(sao, 16, 3);
        a = (int)((sao >> (4 * i)) & 15u);
        o = res ? -a : a;
        env.cw(CW_SAOV) |= ((uint32_t)o & 15u) << (8 + 4 * i);
        put(env.cw(CW_SAO), 16, 3, (uint32_t)i + 1);
        goto sao_sign;
      }
      case S_SAO_BAND:
        env.cw(CW_SAOV) |= res << 2;  // sao_band_position
        goto sao_store;
      case S_SAO_CLASS:
        cl = (int)res;
        if (bits(env.cw(CW_SAO), 19, 2) == 1) put(env.cw(CW_SAO), 23, 2, (uint32_t)cl);
      sao_eo: {
        sao = env.cw(CW_SAO);
        const uint32_t abs4 = sao & 0xffffu;
        v = env.cw(CW_SAOV);
        v |= (uint32_t)cl << 2;
        v |= (abs4 & 15u) << 8;
        v |= ((abs4 >> 4) & 15u) << 12;
        v |= ((uint32_t)(-(int)((abs4 >> 8) & 15u)) & 15u) << 16;
        v |= ((uint32_t)(-(int)((abs4 >> 12) & 15u)) & 15u) << 20;
        env.cw(CW_SAOV) = v;
      }
      sao_store: {
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        sao = env.cw(CW_SAO);
        c = (int)bits(sao, 19, 2);
        env.arenas()->sao[T->sao_off + (size_t)(ry * P->wctb + rx) * 4 + c] = env.cw(CW_SAOV);
        put(env.cw(CW_SAO), 19, 2, (uint32_t)c + 1);
        goto sao_comp;
      }

      // =====================================================================================================
      // 7.3.8.4 coding_quadtree (todo!() at slice.rs:253-255): z-order walk over the CTB's minimum-size blocks
      // =====================================================================================================
      ctu_body:
        put(env.cw(CW_POS), 20, 7, 0);  // z = 0
      cqt_next: {
        P = pp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        z = (int)bits(pos, 20, 7);
        const int log2_ctb = P->log2_ctb, log2_min_cb = P->log2_min_cb;
        const int n_min = 1 << (2 * (log2_ctb - log2_min_cb));
        for (;;) {
          if (z >= n_min) {
            put(env.cw(CW_POS), 20, 7, (uint32_t)z);
            goto ctu_end;
          }
          lvl = log2_ctb - log2_min_cb;
          if (z) {
            tzb = (31 - HEIC_CLZ((uint32_t)z & (0u - (uint32_t)z))) >> 1;
            if (tzb < lvl) lvl = tzb;
          }
          x0 = (rx << log2_ctb) + (int)(compact1by1((uint32_t)z) << log2_min_cb);
          y0 = (ry << log2_ctb) + (int)(compact1by1((uint32_t)z >> 1) << log2_min_cb);
          if (x0 < P->w && y0 < P->h) break;
          z += 1 << (2 * lvl);  // quadrant entirely outside the picture: not coded
        }
        log2 = log2_min_cb + lvl;
        put(pos, 20, 7, (uint32_t)z);
        put(pos, 27, 3, (uint32_t)log2);
        env.cw(CW_POS) = pos;
        env.cw(CW_CU) = (uint32_t)x0 | ((uint32_t)y0 << 14);
        goto cqt_split;
      }
      cqt_split: {  // split_cu_flag at the block of size cq_log2 whose origin is (cu_x, cu_y)
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        log2 = (int)bits(pos, 27, 3);
        cu = env.cw(CW_CU);
        x0 = (int)bits(cu, 0, 14);
        y0 = (int)bits(cu, 14, 14);
        n = 1 << log2;
        depth = P->log2_ctb - log2;
        if (x0 + n <= P->w && y0 + n <= P->h && log2 > P->log2_min_cb) {
          cbp = env.arenas()->ct_depth + T->map8_off;
          inc = 0;
          if (x0 > 0 && cbp[(y0 >> 3) * P->w8 + ((x0 - 1) >> 3)] > depth) inc++;
          if (y0 > 0 && cbp[((y0 - 1) >> 3) * P->w8 + (x0 >> 3)] > depth) inc++;
          op = OP_DEC, arg = (uint32_t)(CTX_SPLIT_CU + inc), st = S_CQT_SPLIT;
          return true;
        }
        split = log2 > P->log2_min_cb;
        goto cqt_have_split;
      }
      case S_CQT_SPLIT:
        split = (int)res;
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        log2 = (int)bits(pos, 27, 3);
      cqt_have_split: {
        if (P->cu_qp_delta_enabled && log2 >= P->log2_min_cu_qp_delta_size) {
          qpw = env.cw(CW_QP);
          put(qpw, 18, 7, 64);  // CuQpDeltaVal = 0
          put(qpw, 25, 1, 0);   // IsCuQpDeltaCoded = 0
          env.cw(CW_QP) = qpw;
        }
        if (split) {
          put(env.cw(CW_POS), 27, 3, (uint32_t)log2 - 1);
          goto cqt_split;
        }
        cu = env.cw(CW_CU);
        x0 = (int)bits(cu, 0, 14);
        y0 = (int)bits(cu, 14, 14);
        {
          bp = env.arenas()->ct_depth + T->map8_off;
          depth = P->log2_ctb - log2;
          b8 = 1 << (log2 - 3);
          x8 = x0 >> 3;
          y8 = y0 >> 3;
          for (j = 0; j < b8; j++) bp[(y8 + j) * P->w8 + x8 + b8 - 1] = (uint8_t)depth;
          for (j = 0; j < b8 - 1; j++) bp[(y8 + b8 - 1) * P->w8 + x8 + j] = (uint8_t)depth;
        }
        // ---- 7.3.8.5 coding_unit: prediction part ----
        cu |= (uint32_t)log2 << 28;  // part_nxn (bit 31) = 0
        env.cw(CW_CU) = cu;
        if (P->cu_qp_delta_enabled) {
          mask = (1 << P->log2_min_cu_qp_delta_size) - 1;
          x_qg = x0 & ~mask;
          y_qg = y0 & ~mask;
          qpw = env.cw(CW_QP);
          if (env.cw(CW_QG) != ((uint32_t)x_qg | ((uint32_t)y_qg << 16))) {
            env.cw(CW_QG) = (uint32_t)x_qg | ((uint32_t)y_qg << 16);
            // 8.6.1 set_qp_pred
            qp_prev = bits(qpw, 26, 1) ? T->slice_qp : (int)bits(qpw, 6, 6);
            put(qpw, 26, 1, 0);
            ctb_mask = (1 << P->log2_ctb) - 1;
            qa = qb = qp_prev;
            cbp = env.arenas()->qp_map + T->map8_off;
            if (x_qg & ctb_mask) qa = cbp[(y_qg >> 3) * P->w8 + ((x_qg - 1) >> 3)];
            if (y_qg & ctb_mask) qb = cbp[((y_qg - 1) >> 3) * P->w8 + (x_qg >> 3)];
            put(qpw, 12, 6, (uint32_t)((qa + qb + 1) >> 1));
          }
          qp_y = ((int)bits(qpw, 12, 6) + ((int)bits(qpw, 18, 7) - 64) + 52) % 52;
          put(qpw, 0, 6, (uint32_t)qp_y);
          env.cw(CW_QP) = qpw;
        }
        if (log2 == P->log2_min_cb) {
          op = OP_DEC, arg = CTX_PART_MODE, st = S_PART;  // decoder.rs:136-149
          return true;
        }
        goto cu_after_part;
      }
      case S_PART:
        P = pp();
        if (!res) {
          env.cw(CW_CU) |= 1u << 31;
          if (bits(env.cw(CW_CU), 28, 3) == 3 && P->log2_min_tb >= 3) goto fail3;
        }
      cu_after_part:
        n_pu = (env.cw(CW_CU) >> 31) ? 4 : 1;
        env.cw(CW_CU2) = (uint32_t)n_pu << 13;  // chroma_mode = 0, prev = 0, k = 0
        env.cw(CW_PU) = 0;
        op = OP_DEC, arg = CTX_PREV_INTRA, st = S_PREV;  // prev_intra_luma_pred_flag of every prediction block first
        return true;
      case S_PREV: {
        cu2 = env.cw(CW_CU2);
        k = (int)bits(cu2, 10, 3);
        cu2 |= res << (6 + k);
        k++;
        n_pu = (int)bits(cu2, 13, 3);
        if (k < n_pu) {
          put(cu2, 10, 3, (uint32_t)k);
          env.cw(CW_CU2) = cu2;
          op = OP_DEC, arg = CTX_PREV_INTRA, st = S_PREV;
          return true;
        }
        put(cu2, 10, 3, 0);
        env.cw(CW_CU2) = cu2;
        goto cu_pu;
      }
      cu_pu:
        cu2 = env.cw(CW_CU2);
        k = (int)bits(cu2, 10, 3);
        if ((cu2 >> (6 + k)) & 1u) {
          op = OP_TR, arg = 2, st = S_PU_MODE;  // mpm_idx
        } else {
          op = OP_BYP, arg = 5, st = S_PU_MODE;  // rem_intra_luma_pred_mode
        }
        return true;
      case S_PU_MODE: {
        P = pp();
        T = tp();
        cu = env.cw(CW_CU);
        cu2 = env.cw(CW_CU2);
        k = (int)bits(cu2, 10, 3);
        n_pu = (int)bits(cu2, 13, 3);
        x0 = (int)bits(cu, 0, 14);
        y0 = (int)bits(cu, 14, 14);
        log2 = (int)bits(cu, 28, 3);
        n = 1 << log2;
        pb = n_pu == 4 ? n >> 1 : n;
        const int prev_k = (int)((cu2 >> (6 + k)) & 1u);
        px = x0 + (k & 1) * pb;
        py = y0 + (k >> 1) * pb;
        bp = env.arenas()->ipm + T->map4_off;
        mode = derive_luma_mode(P, bp, px, py, prev_k, prev_k ? (int)res : 0, prev_k ? 0 : (int)res);
        env.cw(CW_PU) |= (uint32_t)mode << (8 * k);
        // neighbours only ever read the right column and the bottom row of a prediction block
        b4 = pb >> 2;
        x4 = px >> 2;
        y4 = py >> 2;
        for (j = 0; j < b4; j++) bp[(y4 + j) * P->w4 + x4 + b4 - 1] = (uint8_t)mode;
        for (j = 0; j < b4 - 1; j++) bp[(y4 + b4 - 1) * P->w4 + x4 + j] = (uint8_t)mode;
        k++;
        if (k < n_pu) {
          put(cu2, 10, 3, (uint32_t)k);
          env.cw(CW_CU2) = cu2;
          goto cu_pu;
        }
        if (n_pu == 1) env.cw(CW_PU) *= 0x01010101u;
        if (P->chroma) {
          op = OP_DEC, arg = CTX_CHROMA_PRED, st = S_CHROMA1;  // intra_chroma_pred_mode (decoder.rs:23-35,192-204) + 8.4.3
          return true;
        }
        goto tt_begin;
      }
      case S_CHROMA1:
        if (res) {
          op = OP_BYP, arg = 2, st = S_CHROMA2;
          return true;
        }
        idx = 4;
        goto cu_chroma_done;
      case S_CHROMA2:
        idx = (int)res;
      cu_chroma_done:
        luma = (int)(env.cw(CW_PU) & 0xffu);
        if (idx == 4) {
          mode = luma;
        } else {
          mm = idx == 0 ? 0 : (idx == 1 ? 26 : (idx == 2 ? 10 : 1));
          mode = (mm == luma) ? 34 : mm;
        }
        put(env.cw(CW_CU2), 0, 6, (uint32_t)mode);
        goto tt_begin;

      // =====================================================================================================
      // 7.3.8.8 transform_tree, z-order walk over the CU's 4x4 blocks; 7.3.8.10 transform_unit
      // =====================================================================================================
      tt_begin:
        env.cw(CW_TT) = 0;
      tt_next: {
        cu = env.cw(CW_CU);
        tt = env.cw(CW_TT);
        log2 = (int)bits(cu, 28, 3);  // cu_log2
        const int tz = (int)bits(tt, 0, 9);
        if (tz >= (1 << (2 * (log2 - 2)))) goto cu_end;
        lvl = log2 - 2;  // largest block whose origin is tz
        if (tz) {
          tzb = (31 - HEIC_CLZ((uint32_t)tz & (0u - (uint32_t)tz))) >> 1;
          if (tzb < lvl) lvl = tzb;
        }
        put(env.cw(CW_TT), 17, 3, (uint32_t)lvl + 2);
        goto tt_level;
      }
      tt_level: {
        P = pp();
        cu = env.cw(CW_CU);
        tt = env.cw(CW_TT);
        tlog2 = (int)bits(tt, 17, 3);
        depth = (int)bits(cu, 28, 3) - tlog2;
        intra_split = (int)(cu >> 31);
        max_depth = P->max_trafo_depth_intra + intra_split;
        if (tlog2 <= P->log2_max_tb && tlog2 > P->log2_min_tb && depth < max_depth && !(intra_split && depth == 0)) {
          op = OP_DEC, arg = (uint32_t)(CTX_SPLIT_TRANSFORM + 5 - tlog2), st = S_TT_SPLIT;
          return true;
        }
        split = (tlog2 > P->log2_max_tb) || (intra_split && depth == 0);
        goto tt_have_split;
      }
      case S_TT_SPLIT:
        split = (int)res;
        P = pp();
        cu = env.cw(CW_CU);
        tt = env.cw(CW_TT);
        tlog2 = (int)bits(tt, 17, 3);
        depth = (int)bits(cu, 28, 3) - tlog2;
      tt_have_split: {
        put(tt, 20, 1, (uint32_t)split);
        put(tt, 22, 2, 0);  // cbf_cb = cbf_cr = 0
        env.cw(CW_TT) = tt;
        if (P->chroma) {
          par_cb = depth ? (int)((bits(tt, 9, 4) >> (depth - 1)) & 1u) : 1;
          if (tlog2 > 2) {
            if (par_cb) {
              op = OP_DEC, arg = (uint32_t)(CTX_CBF_CHROMA + depth), st = S_TT_CB;
              return true;
            }
          } else if (depth && par_cb) {
            env.cw(CW_TT) |= 1u << 22;  // inferred from the parent when log2TrafoSize == 2
          }
        }
        goto tt_cr;
      }
      case S_TT_CB:
        if (res) env.cw(CW_TT) |= 1u << 22;
      tt_cr: {
        P = pp();
        cu = env.cw(CW_CU);
        tt = env.cw(CW_TT);
        tlog2 = (int)bits(tt, 17, 3);
        depth = (int)bits(cu, 28, 3) - tlog2;
        if (P->chroma) {
          par_cr = depth ? (int)((bits(tt, 13, 4) >> (depth - 1)) & 1u) : 1;
          if (tlog2 > 2) {
            if (par_cr) {
              op = OP_DEC, arg = (uint32_t)(CTX_CBF_CHROMA + depth), st = S_TT_CR;
              return true;
            }
          } else if (depth && par_cr) {
            env.cw(CW_TT) |= 1u << 23;
          }
        }
        goto tt_after_cbf;
      }
      case S_TT_CR:
        if (res) env.cw(CW_TT) |= 1u << 23;
        cu = env.cw(CW_CU);
        tt = env.cw(CW_TT);
        tlog2 = (int)bits(tt, 17, 3);
        depth = (int)bits(cu, 28, 3) - tlog2;
      tt_after_cbf: {
        tt = env.cw(CW_TT);
        cbf_cb = (int)bits(tt, 22, 1);
        cbf_cr = (int)bits(tt, 23, 1);
        // cb_mask / cr_mask: bit d = cbf of the current node at trafoDepth d
        tt = (tt & ~((1u << (9 + depth)) | (1u << (13 + depth)))) | ((uint32_t)cbf_cb << (9 + depth)) | ((uint32_t)cbf_cr << (13 + depth));
        if (bits(tt, 20, 1)) {  // split: one level down, same origin
          put(tt, 17, 3, (uint32_t)tlog2 - 1);
          env.cw(CW_TT) = tt;
          goto tt_level;
        }
        env.cw(CW_TT) = tt;
        op = OP_DEC, arg = (uint32_t)(CTX_CBF_LUMA + (depth == 0 ? 1 : 0)), st = S_TT_LUMA;  // always present for intra CUs
        return true;
      }
      case S_TT_LUMA: {
        // ---- transform_unit: tu_begin ----
        P = pp();
        tt = env.cw(CW_TT);
        put(tt, 21, 1, res);
        tlog2 = (int)bits(tt, 17, 3);
        const int tz = (int)bits(tt, 0, 9);
        has_chroma = P->chroma && (tlog2 > 2 || (tz & 3) == 3);
        cbf_luma = (int)res;
        cbf_cb = (int)bits(tt, 22, 1);
        cbf_cr = (int)bits(tt, 23, 1);
        any_cbf = cbf_luma | cbf_cb | cbf_cr;  // 7.3.8.10: parent-inherited chroma cbfs count for blkIdx 0..2 too
        if (!has_chroma) put(tt, 22, 2, 0);
        put(tt, 24, 1, (uint32_t)has_chroma);
        put(tt, 25, 7, 0);  // ts = 0, comp = 0, dqp_v = 0
        env.cw(CW_TT) = tt;
        if (any_cbf && P->cu_qp_delta_enabled && !bits(env.cw(CW_QP), 25, 1)) {
          // cu_qp_delta_abs: prefix TR cMax 5 (bin 0 ctx 0, bins 1-4 ctx 1) + EG0 suffix (decoder.rs:263-284)
          op = OP_DEC, arg = CTX_CU_QP_DELTA, st = S_DQP_PREFIX;
          return true;
        }
        goto rc_comp;
      }
      case S_DQP_PREFIX: {
        tt = env.cw(CW_TT);
        dv = (int)bits(tt, 29, 3);
        if (res) {
          dv++;
          put(tt, 29, 3, (uint32_t)dv);
          env.cw(CW_TT) = tt;
          if (dv < 5) {
            op = OP_DEC, arg = CTX_CU_QP_DELTA + 1, st = S_DQP_PREFIX;
            return true;
          }
          env.cw(CW_EG) = (0u << 6) | ((uint32_t)S_DQP_EG << 16);  // EG0, back to S_DQP_EG
          env.cw(CW_EGV) = 0;
          op = OP_BYP, arg = 1, st = S_EG_PRE;
          return true;
        }
        env.cw(CW_EGV) = (uint32_t)dv;
        goto dqp_have_abs;
      }
      case S_DQP_EG:
        env.cw(CW_EGV) += 5;
      dqp_have_abs:
        if (env.cw(CW_EGV)) {
          op = OP_BYP, arg = 1, st = S_DQP_SIGN;
          return true;
        }
        neg = 0;
        goto dqp_done;
      case S_DQP_SIGN:
        neg = (int)res;
      dqp_done: {
        if (env.cw(CW_EGV) > 26u) goto fail3;
        dv = (int)env.cw(CW_EGV);
        dv = neg ? -dv : dv;
        if (dv < -26 || dv > 25) goto fail3;
        qpw = env.cw(CW_QP);
        put(qpw, 25, 1, 1);
        put(qpw, 18, 7, (uint32_t)(dv + 64));
        put(qpw, 0, 6, (uint32_t)(((int)bits(qpw, 12, 6) + dv + 52) % 52));
        env.cw(CW_QP) = qpw;
        goto rc_comp;
      }
      // ---- Exp-Golomb of order k, bypass coded (decoder.rs:206-222 with 32-bit arithmetic): unary prefix bin by bin,
      //      suffix in chunks of up to seven bins; returns to the state in CW_EG with the value in CW_EGV ----
      case S_EG_PRE: {
        w = env.cw(CW_EG);
        ones = (int)bits(w, 0, 6);
        if (res) {
          if (++ones > 31) goto fail3;
          put(w, 0, 6, (uint32_t)ones);
          env.cw(CW_EG) = w;
          op = OP_BYP, arg = 1, st = S_EG_PRE;
          return true;
        }
        k = (int)bits(w, 6, 4);
        left = ones + k;
        env.cw(CW_EGV) = ((1u << ones) - 1u) << k;  // + suffix
        put(w, 10, 6, (uint32_t)left);
        env.cw(CW_EG) = w;
        sign_bits = 0;  // suffix accumulator (free between the sign and the level pass of a sub-block)
        goto eg_suffix;
      }
      eg_suffix: {
        w = env.cw(CW_EG);
        left = (int)bits(w, 10, 6);
        if (left > 0) {
          take = left < 7 ? left : 7;
          op = OP_BYP, arg = (uint32_t)take, st = S_EG_SUF;
          return true;
        }
        env.cw(CW_EGV) += sign_bits;
        st = bits(w, 16, 8);
        op = OP_NONE;
        return true;  // the return state consumes no result
      }
      case S_EG_SUF: {
        w = env.cw(CW_EG);
        left = (int)bits(w, 10, 6);
        take = left < 7 ? left : 7;
        sign_bits = (sign_bits << take) | res;
        put(w, 10, 6, (uint32_t)(left - take));
        env.cw(CW_EG) = w;
        goto eg_suffix;
      }

      // =====================================================================================================
      // 7.3.8.11 residual_coding + 9.3.4.2.4-7 of the transform unit's components, one after the other
      // =====================================================================================================
      rc_comp: {
        P = pp();
        T = tp();
        tt = env.cw(CW_TT);
        comp = (int)bits(env.cw(CW_CU2), 16, 2);
        for (;;) {
          if (comp == 3) goto tu_end;
          cbf = comp == 0 ? (int)bits(tt, 21, 1) : (comp == 1 ? (int)bits(tt, 22, 1) : (int)bits(tt, 23, 1));
          if (cbf) break;
          comp++;
        }
        put(env.cw(CW_CU2), 16, 2, (uint32_t)comp);
        cu = env.cw(CW_CU);
        tlog2 = (int)bits(tt, 17, 3);
        const int tz = (int)bits(tt, 0, 9);
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        z = (int)bits(pos, 20, 7);
        ctb4 = 1 << (P->log2_ctb - 2);
        const uint32_t ctb_addr = (uint32_t)(ry * P->wctb + rx);
        const uint32_t z4 = ((uint32_t)z << (2 * (P->log2_min_cb - 2))) + (uint32_t)tz;
        const uint32_t ti = ctb_addr * (uint32_t)(ctb4 * ctb4) + z4;
        A = env.arenas();
        if (comp == 0) {
          out = A->coeff + T->coeff_off[0] + (size_t)ti * 16;
          log2 = tlog2;
          // IntraPredModeY of the prediction block this TU lies in
          pb_shift = (cu >> 31) ? (int)bits(cu, 28, 3) - 1 : (int)bits(cu, 28, 3);
          x0 = (int)(compact1by1((uint32_t)tz) << 2);
          y0 = (int)(compact1by1((uint32_t)tz >> 1) << 2);
          pu_idx = ((x0 >> pb_shift) & 1) | (((y0 >> pb_shift) & 1) << 1);
          pm = (int)((env.cw(CW_PU) >> (8 * pu_idx)) & 0xffu);
        } else {
          out = A->coeff + T->coeff_off[comp] + ((size_t)ctb_addr * (uint32_t)((ctb4 * ctb4) >> 2) + (z4 >> 2)) * 16;
          log2 = tlog2 > 2 ? tlog2 - 1 : 2;
          pm = (int)bits(env.cw(CW_CU2), 0, 6);
        }
        rcw = (uint32_t)log2 | ((uint32_t)comp << 3) | ((uint32_t)pm << 20);
        if (P->tskip_enabled && log2 <= 2) {
          op = OP_DEC, arg = (uint32_t)(CTX_TSKIP + (comp ? 1 : 0)), st = S_RC_TSKIP;
          return true;
        }
        goto rc_last;
      }
      case S_RC_TSKIP:
        rcw |= res << 19;
      rc_last:
        // last_sig_coeff_{x,y}_prefix: both prefixes first, then both suffixes (7.3.8.11).  sign_bits = x | y << 8 | d << 16
        sign_bits = 0;
        op = OP_DEC, arg = (uint32_t)last_ctx(0, (int)bits(rcw, 3, 2), (int)bits(rcw, 0, 3), 0), st = S_RC_LAST_PRE;
        return true;
      case S_RC_LAST_PRE: {
        log2 = (int)bits(rcw, 0, 3);
        c_idx = (int)bits(rcw, 3, 2);
        d = (int)bits(sign_bits, 16, 1);
        v = bits(sign_bits, 8 * d, 8);
        if (res) {
          v++;
          put(sign_bits, 8 * d, 8, v);
          if ((int)v < (log2 << 1) - 1) {
            op = OP_DEC, arg = (uint32_t)last_ctx(d, c_idx, log2, (int)v), st = S_RC_LAST_PRE;
            return true;
          }
        }
        if (d == 0) {
          sign_bits |= 1u << 16;
          op = OP_DEC, arg = (uint32_t)last_ctx(1, c_idx, log2, 0), st = S_RC_LAST_PRE;
          return true;
        }
        put(sign_bits, 16, 2, 0);  // d = 0 for the suffix pass
        goto rc_suffix;
      }
      rc_suffix: {
        d = (int)bits(sign_bits, 16, 2);
        while (d < 2) {
          pv = (int)bits(sign_bits, 8 * d, 8);
          if (pv > 3) {
            put(sign_bits, 16, 2, (uint32_t)d);
            op = OP_BYP, arg = (uint32_t)((pv >> 1) - 1), st = S_RC_LAST_SUF;
            return true;
          }
          d++;
        }
        goto rc_setup;
      }
      case S_RC_LAST_SUF: {
        d = (int)bits(sign_bits, 16, 2);
        pv = (int)bits(sign_bits, 8 * d, 8);
        nb = (pv >> 1) - 1;
        full = (1 << nb) * (2 + (pv & 1)) + (int)res;
        put(sign_bits, 8 * d, 8, (uint32_t)full);
        put(sign_bits, 16, 2, (uint32_t)d + 1);
        goto rc_suffix;
      }
      rc_setup: {
        log2 = (int)bits(rcw, 0, 3);
        c_idx = (int)bits(rcw, 3, 2);
        pm = (int)bits(rcw, 20, 6);
        last_x = (int)bits(sign_bits, 0, 8);
        last_y = (int)bits(sign_bits, 8, 8);
        scan_idx = 0;
        if (log2 == 2 || (log2 == 3 && c_idx == 0)) {
          if (pm >= 6 && pm <= 14) scan_idx = 2;
          else if (pm >= 22 && pm <= 30) scan_idx = 1;
        }
        if (scan_idx == 2) {
          k = last_x;
          last_x = last_y;
          last_y = k;
        }
        lg_sb = log2 - 2;
        last_sb = scan_inv(scan_idx, lg_sb, last_x >> 2, last_y >> 2);
        last_pos = scan_inv(scan_idx, 2, last_x & 3, last_y & 3);
        // sigCtx offset for the non-4x4, non-DC case (9.3.4.2.5), before the luma "not the first sub-block" + 3
        const int sig_off = log2 == 2 ? 0 : (c_idx == 0 ? ((log2 == 3) ? (scan_idx == 0 ? 9 : 15) : 21) : ((log2 == 3) ? 9 : 12));
        // (the prediction mode in bits 20..25 has done its duty: bit 20 now carries sign_data_hiding_enabled_flag)
        rcw = (rcw & 0x0008001fu) | ((uint32_t)scan_idx << 5) | ((uint32_t)last_sb << 7) | ((uint32_t)last_pos << 13) | (1u << 18) |
              ((uint32_t)(pp()->sign_hiding ? 1 : 0) << 20) | ((uint32_t)sig_off << 26);
        csbf = 0;
        sbw = (uint32_t)last_sb & 0xffu;
        goto sb_next;
      }
      sb_next: {  // sub-block i of the block (scan order, counting down); i < 0: the block is done
        i = (int)(int8_t)(sbw & 0xffu);
        if (i < 0) {
          // rc_done: transform_skip_flag of this component, then the next component
          env.cw(CW_TT) |= bits(rcw, 19, 1) << (25 + bits(rcw, 3, 2));
          put(env.cw(CW_CU2), 16, 2, bits(rcw, 3, 2) + 1);
          goto rc_comp;
        }
        log2 = (int)bits(rcw, 0, 3);
        c_idx = (int)bits(rcw, 3, 2);
        scan_idx = (int)bits(rcw, 5, 2);
        last_sb = (int)bits(rcw, 7, 6);
        lg_sb = log2 - 2;
        sb_w = 1 << lg_sb;
        sxy = scan_xy(scan_idx, lg_sb, i);
        xs = (int)(sxy & 15u);
        ys = (int)(sxy >> 4);
        right = (xs < sb_w - 1) ? (int)((csbf >> (ys * 8 + xs + 1)) & 1u) : 0;
        below = (ys < sb_w - 1) ? (int)((csbf >> ((ys + 1) * 8 + xs)) & 1u) : 0;
        // sigCtx (9.3.4.2.5) of all 16 scan positions of this sub-block as one word of nibbles + one additive term
        nib = env.tabs()->sig_nib[log2 == 2 ? scan_idx : 3 + scan_idx * 4 + (right | (below << 1))];
        add = CTX_SIG + (c_idx ? 27 : 0) + (log2 == 2 ? 0 : ((c_idx == 0 && (xs | ys)) ? 3 : 0) + (int)bits(rcw, 26, 6));
        sbw = ((uint32_t)i & 0xffu) | ((uint32_t)xs << 8) | ((uint32_t)ys << 11) | ((uint32_t)add << 19);
        if (i < last_sb && i > 0) {
          sbw |= 1u << 14;  // infer_sb_dc
          op = OP_DEC, arg = (uint32_t)(CTX_CSBF + (c_idx ? 2 : 0) + (right | below)), st = S_CSBF;
          return true;
        }
        goto sb_coded;
      }
      case S_CSBF:
        if (!res) {
          sbw = (sbw & ~0xffu) | (((sbw & 0xffu) - 1u) & 0xffu);
          goto sb_next;
        }
      sb_coded: {
        xs = (int)bits(sbw, 8, 3);
        ys = (int)bits(sbw, 11, 3);
        csbf |= (uint64_t)1 << (ys * 8 + xs);
        i = (int)(int8_t)(sbw & 0xffu);
        n_start = 15;
        sig_g1 = 0;
        if (i == (int)bits(rcw, 7, 6)) {
          last_pos = (int)bits(rcw, 13, 4);
          n_start = last_pos - 1;
          sig_g1 = 1u << last_pos;
        }
        if (n_start > 0) {
          put(sbw, 15, 4, (uint32_t)n_start);
          op = OP_DEC, arg = bits(sbw, 19, 8) + (uint32_t)((nib >> (4 * n_start)) & 15u), st = S_SIG;
          return true;
        }
        if (n_start == 0) goto sb_dc;
        goto sb_after_sig;
      }
      case S_SIG: {
        k = (int)bits(sbw, 15, 4);
        sig_g1 |= res << k;
        k--;
        if (k > 0) {
          put(sbw, 15, 4, (uint32_t)k);
          arg = bits(sbw, 19, 8) + (uint32_t)((nib >> (4 * k)) & 15u);  // op and state stay
          return true;
        }
      }
      sb_dc:
        if (bits(sbw, 14, 1) && sig_g1 == 0) {
          sig_g1 = 1u;  // inferred DC of a coded sub-block with no other significant coefficient
          goto sb_after_sig;
        }
        // the DC coefficient of the whole block has its own context (sigCtx 0)
        log2 = (int)bits(rcw, 0, 3);
        ctx0 = (log2 > 2 && (sbw & 0xffu) == 0) ? CTX_SIG + (bits(rcw, 3, 2) ? 27 : 0) : (int)bits(sbw, 19, 8) + (int)(nib & 15u);
        op = OP_DEC, arg = (uint32_t)ctx0, st = S_DC;
        return true;
      case S_DC:
        sig_g1 |= res;
      sb_after_sig: {
        if (!sig_g1) {
          sbw = (sbw & ~0xffu) | (((sbw & 0xffu) - 1u) & 0xffu);
          goto sb_next;
        }
        // 9.3.4.2.6 / 9.3.4.2.7: up to 8 greater1 flags, one greater2 flag
        c_idx = (int)bits(rcw, 3, 2);
        ctx_set = ((sbw & 0xffu) != 0 && c_idx == 0) ? 2 : 0;
        if (!bits(rcw, 18, 1) && bits(rcw, 17, 1)) ctx_set++;  // !first_sub_block && previous greater1Ctx == 0
        rcw &= ~(1u << 18);
        put(sbw, 27, 2, (uint32_t)ctx_set);
        put(sbw, 29, 2, 1);  // greater1Ctx = 1
        m = sig_g1;
        lvw = 0;
        op = OP_DEC, arg = (uint32_t)(CTX_GT1 + (c_idx ? 16 : 0) + (ctx_set << 2) + 1), st = S_GT1;
        return true;
      }
      case S_GT1: {
        kk = 31 - HEIC_CLZ(m);
        m &= ~(1u << kk);
        num = (int)bits(lvw, 27, 4) + 1;
        put(lvw, 27, 4, (uint32_t)num);
        g1ctx = (int)bits(sbw, 29, 2);
        if (res) {
          sig_g1 |= 1u << (16 + kk);
          g1ctx = 0;
          if (!bits(lvw, 9, 5)) put(lvw, 9, 5, (uint32_t)kk + 1);  // last_g1_pos: the first coefficient with the flag set
        } else if (g1ctx > 0 && g1ctx < 3) {
          g1ctx++;
        }
        put(sbw, 29, 2, (uint32_t)g1ctx);
        c_idx = (int)bits(rcw, 3, 2);
        if (m && num < 8) {
          arg = (uint32_t)(CTX_GT1 + (c_idx ? 16 : 0) + ((int)bits(sbw, 27, 2) << 2) + g1ctx);
          return true;
        }
        put(rcw, 17, 1, g1ctx == 0 ? 1u : 0u);
        if (bits(lvw, 9, 5)) {
          op = OP_DEC, arg = (uint32_t)(CTX_GT2 + (c_idx ? 4 : 0) + (int)bits(sbw, 27, 2)), st = S_GT2;
          return true;
        }
        goto sb_signs;
      }
      case S_GT2:
        lvw |= res << 14;
      sb_signs: {
        // coeff_sign_flag: one bypass bin per coefficient in scan order (none for the hidden one, which comes last)
        const uint32_t sig = sig_g1 & 0xffffu;
        last_sig = 31 - HEIC_CLZ(sig);
        first_sig = 31 - HEIC_CLZ(sig & (0u - sig));
        sign_hidden = bits(rcw, 20, 1) && (last_sig - first_sig > 3);
        put(lvw, 15, 1, (uint32_t)sign_hidden);
        put(lvw, 16, 4, (uint32_t)first_sig);
        n_sign = HEIC_POPC(sig) - (sign_hidden ? 1 : 0);
        sign_bits = 0;
        m = (uint32_t)n_sign;  // sign bins still to read
        if (n_sign) {
          take = n_sign < 7 ? n_sign : 7;
          op = OP_BYP, arg = (uint32_t)take, st = S_SIGN;
          return true;
        }
        goto sb_levels;
      }
      case S_SIGN: {
        take = (int)m < 7 ? (int)m : 7;
        sign_bits = (sign_bits << take) | res;
        m -= (uint32_t)take;
        if (m) {
          arg = m < 7 ? m : 7;
          return true;
        }
        n_sign = HEIC_POPC(sig_g1 & 0xffffu) - (int)bits(lvw, 15, 1);
        sign_bits <<= 32 - n_sign;
      }
      sb_levels:
        m = sig_g1 & 0xffffu;
        put(lvw, 0, 9, 0);  // rice = 0, num_sig = 0, sum parity = 0
        goto lv_next;
      lv_next: {  // choose the next coefficient (m != 0) and ask for its coeff_abs_level_remaining if it has one
        kk = 31 - HEIC_CLZ(m);
        m &= ~(1u << kk);
        num_sig = (int)bits(lvw, 3, 5);
        last_g1 = (int)bits(lvw, 9, 5) - 1;
        base = 1 + (int)((sig_g1 >> (16 + kk)) & 1u) + ((kk == last_g1) ? (int)bits(lvw, 14, 1) : 0);
        put(lvw, 20, 4, (uint32_t)kk);
        put(lvw, 24, 2, (uint32_t)base);
        st = S_LEVEL;
        if (base == ((num_sig < 8) ? ((kk == last_g1) ? 3 : 2) : 1)) {
          op = OP_CALR, arg = bits(lvw, 0, 3);
        } else {
          op = OP_NONE;
          lvw |= 1u << 26;  // no remaining level for this coefficient
        }
        return true;
      }
      case S_LEVEL: {
        base = (int)bits(lvw, 24, 2);
        abs_level = base;
        if (!bits(lvw, 26, 1)) {
          if (res & CALR_ESCAPE) {
            rice = (int)bits(lvw, 0, 3);
            env.cw(CW_EG) = ((uint32_t)(rice + 1) << 6) | ((uint32_t)S_LEVEL_ESC << 16);
            env.cw(CW_EGV) = 0;
            // the suffix accumulator of the Exp-Golomb states is sign_bits: park the sign bins in the cold word
            env.cw(CW_SAOV) = sign_bits;
            op = OP_BYP, arg = 1, st = S_EG_PRE;
            return true;
          }
          if (res & CALR_ONE_MORE) {
            env.cw(CW_EGV) = res & ~CALR_ONE_MORE;
            op = OP_BYP, arg = 1, st = S_LEVEL_ONE;
            return true;
          }
          v = res;
          goto lv_have_rem;
        }
        lvw &= ~(1u << 26);
        goto lv_store;
      }
      case S_LEVEL_ONE:
        v = env.cw(CW_EGV) | res;
        base = (int)bits(lvw, 24, 2);
        goto lv_have_rem;
      case S_LEVEL_ESC:
        sign_bits = env.cw(CW_SAOV);
        v = (4u << bits(lvw, 0, 3)) + env.cw(CW_EGV);
        base = (int)bits(lvw, 24, 2);
      lv_have_rem: {
        if (v > 32768u) goto fail3;
        rice = (int)bits(lvw, 0, 3);
        abs_level = base + (int)v;
        if (abs_level > 3 * (1 << rice)) put(lvw, 0, 3, (uint32_t)(rice < 4 ? rice + 1 : 4));  // decoder.rs:230-236
      }
      lv_store: {
        kk = (int)bits(lvw, 20, 4);
        int val = (sign_bits >> 31) ? -abs_level : abs_level;
        sign_bits <<= 1;
        if (bits(lvw, 15, 1)) {  // sign data hiding: the parity of the sum decides the sign of the first coefficient
          lvw ^= ((uint32_t)abs_level & 1u) << 8;
          if (kk == (int)bits(lvw, 16, 4) && bits(lvw, 8, 1)) val = -val;
        }
        log2 = (int)bits(rcw, 0, 3);
        sxy = scan_xy((int)bits(rcw, 5, 2), 2, kk);
        const int xc = ((int)bits(sbw, 8, 3) << 2) + (int)(sxy & 15u), yc = ((int)bits(sbw, 11, 3) << 2) + (int)(sxy >> 4);
        out[(yc << log2) + xc] = (int16_t)clip3i(-32768, 32767, val);
        put(lvw, 3, 5, bits(lvw, 3, 5) + 1);
        if (m) goto lv_next;
        sbw = (sbw & ~0xffu) | (((sbw & 0xffu) - 1u) & 0xffu);
        goto sb_next;
      }

      // =====================================================================================================
      tu_end: {
        P = pp();
        T = tp();
        tt = env.cw(CW_TT);
        cu = env.cw(CW_CU);
        cu2 = env.cw(CW_CU2);
        put(env.cw(CW_CU2), 16, 2, 0);  // comp = 0 for the next transform unit
        tlog2 = (int)bits(tt, 17, 3);
        const int tz = (int)bits(tt, 0, 9);
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        z = (int)bits(pos, 20, 7);
        ctb4 = 1 << (P->log2_ctb - 2);
        const uint32_t ti = (uint32_t)(ry * P->wctb + rx) * (uint32_t)(ctb4 * ctb4) + ((uint32_t)z << (2 * (P->log2_min_cb - 2))) + (uint32_t)tz;
        pb_shift = (cu >> 31) ? (int)bits(cu, 28, 3) - 1 : (int)bits(cu, 28, 3);
        x0 = (int)(compact1by1((uint32_t)tz) << 2);
        y0 = (int)(compact1by1((uint32_t)tz >> 1) << 2);
        pu_idx = ((x0 >> pb_shift) & 1) | (((y0 >> pb_shift) & 1) << 1);
        const uint32_t luma_mode = (env.cw(CW_PU) >> (8 * pu_idx)) & 0xffu;
        env.arenas()->tu_map[T->tu_off + ti] =
            1u | ((uint32_t)(tlog2 - 2) << 1) | (bits(tt, 21, 1) << 3) | (bits(tt, 22, 1) << 4) | (bits(tt, 23, 1) << 5) |
            (bits(tt, 24, 1) << 6) | (luma_mode << 7) | (bits(cu2, 0, 6) << 13) | (bits(env.cw(CW_QP), 0, 6) << 19) |
            (bits(tt, 25, 3) << 25);
        put(env.cw(CW_TT), 0, 9, (uint32_t)tz + (1u << (2 * (tlog2 - 2))));
        goto tt_next;
      }
      cu_end: {
        // QpY of the CU (8.6.1): CuQpDeltaVal decoded anywhere inside the CU applies to all of it
        P = pp();
        T = tp();
        cu = env.cw(CW_CU);
        x0 = (int)bits(cu, 0, 14);
        y0 = (int)bits(cu, 14, 14);
        log2 = (int)bits(cu, 28, 3);
        n = 1 << log2;
        qpw = env.cw(CW_QP);
        qp_y = (int)bits(qpw, 0, 6);
        bp = env.arenas()->qp_map + T->map8_off;
        for (yy = y0 >> 3; yy < (y0 + n) >> 3; yy++)
          for (xx = x0 >> 3; xx < (x0 + n) >> 3; xx++) bp[yy * P->w8 + xx] = (uint8_t)qp_y;
        put(qpw, 6, 6, (uint32_t)qp_y);  // last_qp_y
        env.cw(CW_QP) = qpw;
        pos = env.cw(CW_POS);
        put(pos, 20, 7, bits(pos, 20, 7) + (1u << (2 * (log2 - P->log2_min_cb))));
        env.cw(CW_POS) = pos;
        goto cqt_next;
      }
      ctu_end: {
        P = pp();
        T = tp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        env.cw(CW_CTUS)++;
        if (P->wpp && rx == 1 && ry + 1 < P->hctb) {  // 9.3.2.2: the contexts after the second CTU seed the next row
          bp = env.arenas()->wpp_save + T->wpp_off + (size_t)ry * NUM_CTX_PAD;
          for (j = 0; j < NUM_CTX; j++) bp[j] = (uint8_t)env.ld_ctx(j);
        }
        op = OP_TERM, arg = 0, st = S_EOS;  // end_of_slice_segment_flag (slice.rs:214)
        return true;
      }
      case S_EOS: {
        P = pp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        addr = ry * P->wctb + rx;
        if ((int)res != (addr == P->wctb * P->hctb - 1)) goto fail3;
        if (P->wpp && rx == P->wctb - 1 && !res) {
          op = OP_TERM, arg = 0, st = S_EOSUB;  // end_of_subset_one_bit (slice.rs:222-227)
          return true;
        }
        goto ctu_published;
      }
      case S_EOSUB:
        if (!res) goto fail3;
      ctu_published: {
        P = pp();
        pos = env.cw(CW_POS);
        rx = (int)bits(pos, 0, 10);
        ry = (int)bits(pos, 10, 10);
        env.publish(ry, rx + 1);
        rx++;
        if (rx >= P->wctb) {
          rx = 0;
          ry += env.n_slots();
          if (ry >= P->hctb) goto tile_done;
        }
        env.cw(CW_POS) = (uint32_t)rx | ((uint32_t)ry << 10);
        goto ctu_begin;
      }
      fail3:
        tile_failed(-3);
      tile_done:
        env.finish_tile(env.cw(CW_TILE), e.bins, env.cw(CW_CTUS));
        env.cw(CW_USE)++;
        st = S_TILE_NEXT;
      idle:
        op = OP_NONE;
        return true;
      default:
        return false;
    }
  }
};

}  // namespace dev
}  // namespace heic

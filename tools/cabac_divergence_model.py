"""Where do 32 different tiles in one CABAC warp lose their time?  A host-side model, no GPU needed.

The device walks CTU -> coding unit -> transform unit -> component -> 4x4 sub-block as nested loops; under SIMT every
loop iteration costs the warp the MAXIMUM over the lanes that have that iteration (a lane without one idles).  This script
takes the syntax of the fixture's 48 tiles from the CPU oracle (TEST INFRASTRUCTURE), deals 32 of them into a warp the way
bench.py's `distinct_tiles_per_warp` does, and evaluates that recursion with a per-sub-block cost in approximate bins,
once as it is and once with one level "flattened" (the lanes' iterations of that level re-aligned perfectly: sum per lane
first, maximum afterwards).  The ratios say which level is worth attacking; absolute numbers are a model.

    python tools/cabac_divergence_model.py [tiles_per_warp=32]
    python tools/cabac_divergence_model.py sweep      # model of bench.py's k = 1..32 different tiles per warp

Calibration against the B200 (CABAC stage, 592 images, before the bypass change): measured 62 / 87 / 123 / 179 / 226 /
255 ms for k = 1 / 2 / 4 / 8 / 16 / 32; the model's ratios 1 / 1.70 / 2.69 / 3.83 / 5.12 / 6.43 fit them as
26.5 ms + 35.5 ms x ratio (within 10 %), i.e. about 40 % of the converged time does not scale with divergence.  For the
position-paired coding-unit loop the model gives 5.37 -> 217 ms; measured 218 ms.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import heif_b200 as H  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

DIAG4 = [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0), (0, 3), (1, 2), (2, 1), (3, 0), (1, 3), (2, 2), (3, 1), (2, 3), (3, 2), (3, 3)]  # (x, y)


def diag(n):
    out = []
    for s in range(2 * n - 1):
        for x in range(n):
            y = s - x
            if 0 <= y < n:
                out.append((x, y))
    # up-right diagonal scan (6.5.3) walks each anti-diagonal from bottom-left to top-right
    res, i = [], 0
    for s in range(2 * n - 1):
        cnt = sum(1 for x in range(n) if 0 <= s - x < n)
        seg = out[i:i + cnt]
        res += sorted(seg, key=lambda p: -p[1])
        i += cnt
    return res


SB_SCAN = {1: diag(1), 2: diag(2), 4: diag(4), 8: diag(8)}


def block_cost(blk):
    """blk: n x n int16 levels (raster).  Returns the per-sub-block costs in scan order from the last coded one down to 0:
    (context-coded bins, bypass steps)."""
    n = blk.shape[0]
    nsb = n // 4
    scan = SB_SCAN[nsb]
    sb_nz = [(blk[4 * y:4 * y + 4, 4 * x:4 * x + 4] != 0) for (x, y) in scan]
    coded = [m.any() for m in sb_nz]
    if not any(coded):
        return []
    last = max(i for i, c in enumerate(coded) if c)
    costs = []
    for i in range(last, -1, -1):
        m = sb_nz[i]
        if not coded[i]:
            costs.append((1, 0))  # coded_sub_block_flag = 0
            continue
        sub = blk[4 * scan[i][1]:4 * scan[i][1] + 4, 4 * scan[i][0]:4 * scan[i][0] + 4]
        pos = [k for k, (x, y) in enumerate(DIAG4) if sub[y, x] != 0]
        nnz = len(pos)
        n_sig = (max(pos) if i == last else 16)  # sig_coeff_flags decoded
        ctx = (0 if i in (0, last) else 1) + n_sig + min(8, nnz) + 1
        big = int((np.abs(sub) > 2).sum())
        byp = (nnz + 6) // 7 + big  # sign bins in groups of 7, one step per remaining level
        costs.append((ctx, byp))
    return costs


def tile_structure(img, t):
    """-> list over CTUs of list over CUs of list over TUs of list over components of list over sub-blocks of cost."""
    td = img.tiles[t]
    r = O.decode_picture(img.sps, img.pps, td.header, (td.rbsp, td.rbsp_len), parse_only=True)
    tu, lv = r["tu_map"], r["level"]
    ctb4sq = 64  # CTB 32 -> 8 x 8 blocks of 4 x 4
    ctus, cu_pos = [], []
    for c in range(len(tu) // ctb4sq):
        cus, pos, z = [], [], 0
        while z < ctb4sq:
            w = int(tu[c * ctb4sq + z])
            if not (w & 1):
                z += 1
                continue
            lg = ((w >> 1) & 3) + 2
            n = 1 << lg
            if lg == 2:  # four 4x4 TUs of an NxN 8x8 coding unit
                tus = []
                for q in range(4):
                    wq = int(tu[c * ctb4sq + z + q])
                    comps = []
                    if wq & 8:
                        comps.append(block_cost(lv[0][(c * ctb4sq + z + q) * 16:(c * ctb4sq + z + q) * 16 + 16].reshape(4, 4)))
                    if wq & 64:
                        for ci, bit in ((1, 16), (2, 32)):
                            if wq & bit:
                                o = (c * ctb4sq + z) * 4
                                comps.append(block_cost(lv[ci][o:o + 16].reshape(4, 4)))
                    tus.append(comps)
                cus.append(tus)
                pos.append(z)
                z += 4
                continue
            comps = []
            o = (c * ctb4sq + z) * 16
            if w & 8:
                comps.append(block_cost(lv[0][o:o + n * n].reshape(n, n)))
            for ci, bit in ((1, 16), (2, 32)):
                if w & bit:
                    oc = (c * ctb4sq + z) * 4
                    comps.append(block_cost(lv[ci][oc:oc + n * n // 4].reshape(n // 2, n // 2)))
            cus.append([comps])
            pos.append(z)
            z += (n // 4) ** 2
        ctus.append(cus)
        cu_pos.append(pos)
    return ctus, cu_pos


W_CTX, W_BYP, W_CU, W_TU, W_COMP = 1.0, 1.0, 8.0, 4.0, 6.0  # header bins per coding unit / transform unit / component


def leaf(c):
    return W_CTX * c[0] + W_BYP * c[1]


def lane_total(node, depth):
    """cost of one lane alone"""
    if depth == 4:
        return leaf(node)
    head = (W_CU, W_TU, W_COMP, 0.0)[depth - 1] if depth >= 1 else 0.0
    return head * (1 if depth >= 1 else 0) + sum(lane_total(ch, depth + 1) for ch in node)


def joint(nodes, depth, flat_level):
    """SIMT cost of the lanes in `nodes` walking level `depth` together (0: CUs of a CTU ... 3: sub-blocks, 4: leaf)."""
    if not nodes:
        return 0.0
    if depth == 4:
        # phases of a sub-block reconverge one by one: context-coded run, then the bypass part
        return W_CTX * max(n[0] for n in nodes) + W_BYP * max(n[1] for n in nodes)
    head = (0.0, W_CU, W_TU, W_COMP)[depth]
    if depth == flat_level:
        return head + max(lane_total(n, depth) - (head if depth >= 1 else 0.0) for n in nodes)
    total = head
    for j in range(max(len(n) for n in nodes)):
        total += joint([n[j] for n in nodes if len(n) > j], depth + 1, flat_level)
    return total


def joint_by_position(nodes, positions, flat_level):
    """The coding-unit loop paired by POSITION instead of by index: one warp-uniform pass over the CTB's minimum-size
    block positions; at each one the lanes whose next coding unit starts there decode it, the others idle."""
    total = 0.0
    for z in sorted({p for ps in positions for p in ps}):
        here = [n[ps.index(z)] for n, ps in zip(nodes, positions) if z in ps]
        total += joint(here, 1, flat_level)
    return total


def sweep():
    f = H.HeicFile(open(os.path.join(ROOT, "tests", "golden", "halfmoonbay.heic"), "rb").read())
    img = f.primary
    order = sorted(range(img.n_tiles), key=lambda t: -img.tiles[t].rbsp_len)
    S, P = {}, {}
    for t in order:
        S[t], P[t] = tile_structure(img, t)
    n_ctu = len(S[order[0]])
    uniform = sum(sum(lane_total(S[t][c], 0) for c in range(n_ctu)) for t in order)  # one warp step per tile
    for by_pos in (False, True):
        for k in (1, 2, 4, 8, 16, 32):
            m = 0.0
            for b0 in range(0, img.n_tiles, k):
                lanes = order[b0:b0 + k]
                if by_pos:
                    j = sum(joint_by_position([S[t][c] for t in lanes], [P[t][c] for t in lanes], -1) for c in range(n_ctu))
                else:
                    j = sum(joint([S[t][c] for t in lanes], 0, -1) for c in range(n_ctu))
                m += j * len(lanes)
            print(f"{'by position' if by_pos else 'by index   '}  k = {k:2d}: {m / uniform:5.2f} x the converged cost"
                  f"  -> {26.5 + 35.5 * m / uniform:6.1f} ms with the B200 fit")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        return sweep()
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    f = H.HeicFile(open(os.path.join(ROOT, "tests", "golden", "halfmoonbay.heic"), "rb").read())
    img = f.primary
    order = sorted(range(img.n_tiles), key=lambda t: -img.tiles[t].rbsp_len)
    print(f"parsing {img.n_tiles} tiles with the oracle ...", flush=True)
    S, P = {}, {}
    for t in order:
        S[t], P[t] = tile_structure(img, t)
    names = ["as the device walks it (nested)", "coding-unit loop flattened", "transform-unit loop flattened",
             "component loop flattened", "sub-block loop flattened"]
    rows = []
    for w0 in range(0, img.n_tiles - k + 1, max(1, k // 2)):
        lanes = order[w0:w0 + k]
        n_ctu = len(S[lanes[0]])
        alone = [sum(lane_total(S[t][c], 0) for c in range(n_ctu)) for t in lanes]
        per_ctu_max = sum(max(lane_total(S[t][c], 0) for t in lanes) for c in range(n_ctu))
        res = [sum(joint([S[t][c] for t in lanes], 0, fl) for c in range(n_ctu)) for fl in (-1, 0, 1, 2, 3)]
        by_pos = sum(joint_by_position([S[t][c] for t in lanes], [P[t][c] for t in lanes], -1) for c in range(n_ctu))
        by_pos_tu = sum(joint_by_position([S[t][c] for t in lanes], [P[t][c] for t in lanes], 1) for c in range(n_ctu))
        rows.append((lanes, alone, per_ctu_max, res, by_pos))
        print(f"warp of tiles ranked {w0}..{w0 + k - 1} by size: mean lane {np.mean(alone):9.0f}  heaviest lane {max(alone):9.0f}  "
              f"sum of per-CTU maxima {per_ctu_max:9.0f}")
        for nm, v in zip(names, res):
            print(f"    {nm:38s} {v:10.0f}   x{v / np.mean(alone):5.2f} of the mean lane, x{v / per_ctu_max:5.2f} of the per-CTU maxima")
        for nm, v in (("coding units paired by position", by_pos), ("  ... and everything below a CU flattened", by_pos_tu)):
            print(f"    {nm:38s} {v:10.0f}   x{v / np.mean(alone):5.2f} of the mean lane, x{v / per_ctu_max:5.2f} of the per-CTU maxima")
    return rows


if __name__ == "__main__":
    main()

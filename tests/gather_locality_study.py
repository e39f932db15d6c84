"""Diagnostic (not a test): how many distinct 128-byte lines does one warp-wide neighbour gather touch, and how would other
neighbour-list ORDERS or particle orders change that?

Why: ncu shows every neighbour sweep bound by L1 data-stage wavefronts, ~9.8 per scattered 16-byte gather = about one per
distinct cache line touched. That number can be computed exactly on the CPU from a realistic particle
state, so list orders can be compared without a GPU:

    python tests/gather_locality_study.py [n_side=50] [steps=40]

A dam break of n_side^3 particles is advanced `steps` steps with the CPU port (oracle/ — which is why this file lives under
tests/), the LINEAR_CELL sorted order and neighbour lists are rebuilt with the host harness of tests/cpp/list_build_host.cu,
and for every warp (32 consecutive sorted particles) and list slot k the distinct lines (index >> 3 for float4 arrays) and
32-byte sectors (index >> 1) among the active lanes' k-th neighbours are counted.
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

H = np.float32(0.1)


def state_after(n_side, steps):
    from akuaengine_b200 import scenes
    from oracle import PortOracle, param_block
    p, bmin, bmax = scenes.dam_break(n_side)
    o = PortOracle(p, param_block())
    for _ in range(steps):
        o.step(0.0083, bmin, bmax)
    pos = o.particles["position"].copy()
    o.close()
    return pos, bmin, bmax


def lists_for(pos, bmin, bmax, max_n=128):
    import test_list_build_host as t
    import subprocess, shutil
    if not t.OUT.exists():
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        t.OUT.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler",
                        "-fPIC,-ffp-contract=off", "-shared", "-o", str(t.OUT), str(t.SRC)], check=True)
    lib = C.CDLL(str(t.OUT))
    fn = lib.akua_test_list_build_host
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return t._run(fn, pos, bmin, bmax, 0.1, max_n, 0)


def cell_coords(p):
    return np.floor(p / H).astype(np.int64)


def warp_metrics(lst, cnt, index_map=None, lane_perm=None):
    """lst: (n, K) neighbour indices in slot order (entries >= cnt ignored); index_map: neighbour index -> storage slot
    (for alternative particle orders); lane_perm: order in which particles are assigned to lanes.
    Returns dict with gather instructions (warp slots), distinct lines and sectors summed over all warp slots."""
    n, K = lst.shape
    order = np.arange(n) if lane_perm is None else lane_perm
    pad = (-n) % 32
    order = np.concatenate([order, np.full(pad, -1)])
    W = len(order) // 32
    order = order.reshape(W, 32)
    valid_lane = order >= 0
    o = np.where(valid_lane, order, 0)
    c = np.where(valid_lane, cnt[o], 0)                        # (W, 32)
    kmax = c.max(axis=1)                                        # slots a warp executes
    slots_total = int(kmax.sum())
    padded_groups = int((((kmax + 3) // 4) * 4).sum())
    lines_total = sectors_total = 0
    active_total = 0
    per_k = []
    for k in range(int(kmax.max())):
        act = c > k                                             # (W, 32)
        j = np.where(act, lst[o, min(k, K - 1)].astype(np.int64), 0)
        if index_map is not None:
            j = index_map[j]
        for shift, name in ((3, "lines"), (1, "sectors")):
            v = np.where(act, j >> shift, -1)
            v.sort(axis=1)
            d = (np.diff(v, axis=1) != 0).sum(axis=1) + 1      # distinct values incl. the -1 bucket
            d = d - (v[:, 0] == -1)                            # drop the inactive bucket
            d = np.where(act.any(axis=1), d, 0)
            if name == "lines":
                lines_total += int(d.sum()); per_k.append(float(d[act.any(axis=1)].mean()))
            else:
                sectors_total += int(d.sum())
        active_total += int(act.sum())
    return {"warp_gather_slots": slots_total, "slots_padded_to_4": padded_groups, "lines": lines_total, "sectors": sectors_total,
            "lines_per_slot": lines_total / slots_total, "sectors_per_slot": sectors_total / slots_total,
            "active_lanes_per_slot": active_total / slots_total, "lines_by_k": [round(x, 2) for x in per_k[:40:4]]}


def reorder_lists(lst, cnt, keyfn):
    """Re-sorts every particle's list by keyfn(i_index_array, j_index_array, rank_array) (stable)."""
    n, K = lst.shape
    out = lst.copy()
    ks = np.arange(K)[None, :]
    act = ks < cnt[:, None]
    i = np.broadcast_to(np.arange(n)[:, None], (n, K))
    key = keyfn(i, lst.astype(np.int64), act)
    key = np.where(act, key, np.iinfo(np.int64).max)
    idx = np.argsort(key, axis=1, kind="stable")
    return np.take_along_axis(out, idx, axis=1)


def main():
    n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    pos, bmin, bmax = state_after(n_side, steps)
    order, lst, cnt = lists_for(pos, bmin, bmax)
    p = pos[order]
    n = len(p)
    cells = cell_coords(p)
    print(json.dumps({"particles": n, "steps": steps, "mean_neighbours": float(cnt.mean()), "max": int(cnt.max())}))
    res = {}
    res["O0 current: rows (dx,dy) ascending, index ascending"] = warp_metrics(lst, cnt)

    # row of a neighbour relative to its owner and the hit rank inside that row
    def row_of(i, j):
        d = cells[j] - cells[i]
        return (np.clip(d[..., 0], -1, 1) + 1) * 3 + (np.clip(d[..., 1], -1, 1) + 1)
    n_, K = lst.shape
    ii = np.broadcast_to(np.arange(n_)[:, None], (n_, K))
    act = np.arange(K)[None, :] < cnt[:, None]
    rows = np.where(act, row_of(ii, np.where(act, lst, 0).astype(np.int64)), 99)
    # rank within row (lists are row-major already, so rank = position - first position of that row)
    first = np.full((n_, 10), K, np.int64)
    for r in range(9):
        m = rows == r
        pos_first = np.where(m.any(axis=1), m.argmax(axis=1), K)
        first[:, r] = pos_first
    rank = np.arange(K)[None, :] - np.take_along_axis(first, np.minimum(rows, 9), axis=1)

    res["O2 round-robin over rows (t-th hit of every row, then t+1)"] = warp_metrics(
        reorder_lists(lst, cnt, lambda i, j, a: rank * 16 + rows), cnt)
    res["O3 by z cell of the neighbour, then row"] = warp_metrics(
        reorder_lists(lst, cnt, lambda i, j, a: (cells[np.where(a, j, 0)][..., 2] - cells[i][..., 2] + 1) * 16 + rows), cnt)
    res["O4 descending index"] = warp_metrics(reorder_lists(lst, cnt, lambda i, j, a: -j), cnt)

    # padded row lockstep: every row padded to the warp's maximum hit count in that row
    hits = np.stack([(rows == r).sum(axis=1) for r in range(9)], axis=1)          # (n, 9)
    padn = (-n) % 32
    hw = np.concatenate([hits, np.zeros((padn, 9), hits.dtype)]).reshape(-1, 32, 9)
    res["P1 rows padded to the warp maximum (slots only)"] = {"warp_gather_slots": int(hw.max(axis=1).sum()),
                                                             "vs_current_slots": float(hw.max(axis=1).sum() / res[next(iter(res))]["warp_gather_slots"])}
    hq = np.concatenate([hits, np.zeros(((-n) % 8, 9), hits.dtype)]).reshape(-1, 8, 9)
    res["P2 rows padded to the quarter-warp maximum (slots only)"] = {"quarter_slots_mean": float(hq.max(axis=1).sum(axis=1).mean()),
                                                                     "list_len_mean": float(cnt.mean())}

    # Morton order of cells instead of x-major / z-fastest (storage order AND lane assignment change)
    def part1by2(v):
        v = v & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    cc = cells - cells.min(axis=0)
    morton = part1by2(cc[:, 0]) | (part1by2(cc[:, 1]) << 1) | (part1by2(cc[:, 2]) << 2)
    perm = np.argsort(morton, kind="stable")                     # new slot s holds old particle perm[s]
    inv = np.empty(n, np.int64); inv[perm] = np.arange(n)
    res["M0 Morton cell order, lists in ascending NEW index"] = warp_metrics(
        reorder_lists(lst, cnt, lambda i, j, a: inv[np.where(a, j, 0)]), cnt, index_map=inv, lane_perm=perm)
    # 2x2x2 blocks of cells (z fastest inside the block): a warp = 4 cells of one block
    blk = ((cc[:, 0] >> 1) * 4096 + (cc[:, 1] >> 1)) * 4096 + (cc[:, 2] >> 1)
    sub = (cc[:, 0] & 1) * 4 + (cc[:, 1] & 1) * 2 + (cc[:, 2] & 1)
    perm2 = np.lexsort((np.arange(n), sub, blk))
    inv2 = np.empty(n, np.int64); inv2[perm2] = np.arange(n)
    res["M1 2x2x2 cell blocks, lists in ascending NEW index"] = warp_metrics(
        reorder_lists(lst, cnt, lambda i, j, a: inv2[np.where(a, j, 0)]), cnt, index_map=inv2, lane_perm=perm2)
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()

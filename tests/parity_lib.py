"""Shared helpers for the parity tests: teacher-forced phase comparison of the CUDA solver against an oracle.

"Teacher forcing" (SURVEY.md §4): integer structures are only well defined when both implementations start a phase from
bit-identical state, so each phase is run from the ORACLE's pre-phase state (uploaded as AoS-108) and compared after.
"""
from __future__ import annotations

import numpy as np

from akuaengine_b200 import DBG, KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PARTICLE_DTYPE, PBFConfig, LambdaCorrParams, PBFSolver

H = 0.1


def ids_of(p):
    """Particle ids stashed in color.x by the fixtures."""
    return p["color"][:, 0].astype(np.int64)


def rel_err(a, b, scale=None):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if scale is None:
        scale = max(np.abs(b).max(), 1e-30)
    return float(np.abs(a - b).max() / scale)


def neighbour_sets_by_id(nbr_arr, nbr_cnt, ids):
    """-> dict id -> sorted tuple of neighbour ids (multiset: duplicates kept)."""
    out = {}
    for i in range(len(nbr_cnt)):
        out[int(ids[i])] = tuple(sorted(int(ids[j]) for j in nbr_arr[i, :nbr_cnt[i]]))
    return out


def ragged_to_padded(flat, cnt, max_n):
    arr = np.zeros((len(cnt), max_n), np.uint32)
    off = 0
    for i, c in enumerate(cnt):
        arr[i, :c] = flat[off:off + c]
        off += c
    return arr


def make_solver(n, params, key_mode, **kw):
    cfg = PBFConfig(restDensity=float(params[0]), particle_spacing=float(params[1]), smoothRadius=float(params[2]),
                    spatialHashCellSize=float(params[3]), relaxation=float(params[4]), vorticityEpsilon=float(params[5]),
                    viscosity=float(params[6]), maxNeighbours=int(params[7]), solverIterations=int(params[8]),
                    gravity=[float(x) for x in params[9:12]])
    corr = LambdaCorrParams(k=float(params[12]), n=float(params[13]), delta_q=float(params[14]))
    return PBFSolver(n, cfg, corr, key_mode=key_mode, **kw)


def phase_report(trace: dict, key_mode: int, fast_math: bool = True) -> dict:
    """Runs every phase of one step teacher-forced from `trace` (a fixture of tests/golden/make_golden.py or the same
    structure produced live from an oracle) and returns a dict of error measures."""
    init = trace["init"]
    n = len(init)
    dt = float(trace["dt"])
    bmin, bmax = trace["box_min"], trace["box_max"]
    params = trace["params"]
    iters = int(trace["iters"])
    max_n = int(params[7])
    rep = {}
    s = make_solver(n, params, key_mode, fast_math=fast_math)

    # ---- K1 predict: x* must be bit-exact
    s.upload_particles(init)
    s.predictNewPosition(dt)
    xs = s.debug(DBG.XSTAR)[:, :3]
    rep["predict_xstar_bitexact"] = bool(np.array_equal(xs.view(np.uint32), trace["after_predict"]["new_position"].view(np.uint32)))

    # ---- neighbour phase from the oracle's post-predict state
    s.upload_particles(trace["after_predict"])
    s.findParticleNeighbours(bmin, bmax)
    got = s.download_particles()
    want = trace["after_neighbours"]
    cnt = s.debug(DBG.NBR_COUNT)
    lst = s.debug(DBG.NBR_LIST)
    want_cnt = trace["nbr_count"]
    want_lst = ragged_to_padded(trace["nbr_flat"], want_cnt, max_n)
    ids_got = s.debug(DBG.ID).astype(np.int64)          # sorted slot -> upload slot
    ids_want_sorted = ids_of(want)
    ids_upload = ids_of(trace["after_predict"])
    if key_mode == KEY_REFERENCE_HASH:
        # keys of the unsorted particles == hash the reference computed (compare through particle ids)
        keys_unsorted = s.debug(DBG.KEYS_UNSORTED)
        ref_hash_by_id = np.zeros(n, np.uint32)
        ref_hash_by_id[ids_want_sorted] = want["hash"]
        rep["keys_bitexact"] = bool(np.array_equal(keys_unsorted, ref_hash_by_id[ids_upload]))
        rep["sorted_keys_bitexact"] = bool(np.array_equal(s.debug(DBG.KEYS_SORTED), want["hash"]))
        rep["permutation_bitexact"] = bool(np.array_equal(ids_upload[ids_got], ids_want_sorted))
        # bucket-start table == K3 applied to the reference's sorted hashes
        table = s.debug(DBG.BUCKET_START)
        exp = np.full(len(table), 0xFFFFFFFF, np.uint32)
        hs = want["hash"]
        first = np.ones(n, bool)
        first[1:] = hs[1:] != hs[:-1]
        exp[hs[first]] = np.nonzero(first)[0].astype(np.uint32)
        rep["bucket_table_bitexact"] = bool(np.array_equal(table, exp))
        rep["nbr_count_bitexact"] = bool(np.array_equal(cnt, want_cnt))
        mask = np.arange(max_n)[None, :] < want_cnt[:, None]
        rep["nbr_list_bitexact"] = bool(np.array_equal(lst[mask], want_lst[mask])) if rep["nbr_count_bitexact"] else False
        rep["sorted_state_bitexact"] = bool(np.array_equal(got["new_position"].view(np.uint32), want["new_position"].view(np.uint32))
                                            and np.array_equal(got["position"].view(np.uint32), want["position"].view(np.uint32))
                                            and np.array_equal(got["velocity"].view(np.uint32), want["velocity"].view(np.uint32)))
    else:
        keys_sorted = s.debug(DBG.KEYS_SORTED)
        rep["sorted_keys_monotone"] = bool(np.all(keys_sorted[1:] >= keys_sorted[:-1]))
        perm = s.debug(DBG.PERM).astype(np.int64)
        rep["permutation_is_stable_sort"] = bool(np.array_equal(perm, np.argsort(s.debug(DBG.KEYS_UNSORTED), kind="stable")))
    # neighbour SETS as sets of particle ids (both modes)
    sets_got = neighbour_sets_by_id(lst, cnt, ids_upload[ids_got])
    sets_want = neighbour_sets_by_id(want_lst, want_cnt, ids_want_sorted)
    rep["nbr_sets_equal"] = sets_got == sets_want
    rep["nbr_mean"] = float(cnt.mean())
    rep["nbr_max"] = int(cnt.max())

    # ---- constraint solve from the oracle's post-neighbour state. In REFERENCE_HASH mode our sorted order equals the
    # oracle's, so the solver continues from its own lists (verified identical above). In LINEAR_CELL mode compare by id.
    def by_id(p, field):
        out = np.zeros_like(p[field])
        out[ids_of(p)] = p[field]
        return out

    def ours_by_id(arr):
        out = np.zeros_like(arr)
        out[ids_upload[s.debug(DBG.ID).astype(np.int64)]] = arr
        return out

    s.runConstraintSolver(iters, bmin, bmax)
    got = s.download_particles()
    want = trace["after_solve"]
    rep["solve_xstar_rel_h"] = rel_err(ours_by_id(got["new_position"]), by_id(want, "new_position"), H)
    rep["solve_density_rel"] = rel_err(ours_by_id(got["density"]), by_id(want, "density"))
    rep["solve_lambda_rel"] = rel_err(ours_by_id(got["lambda"]), by_id(want, "lambda"))
    rep["solve_dp_rel_h"] = rel_err(ours_by_id(got["position_delta"]), by_id(want, "position_delta"), H)

    # ---- K9, K10 from the oracle's post-solve state. Upload resets ids/order, so re-run the neighbour phase on the
    # uploaded state (the lists depend only on x* ordering, which the upload preserves for REFERENCE_HASH).
    def forced(state_before):
        s.upload_particles(state_before)

    forced(trace["after_solve"])
    s.updatePositionAndVelocity(dt)
    got = s.download_particles()
    want = trace["after_update"]
    vscale = H / dt
    rep["update_pos_bitexact"] = bool(np.array_equal(got["position"].view(np.uint32), want["position"].view(np.uint32)))
    rep["update_vel_rel"] = rel_err(got["velocity"], want["velocity"], vscale)
    forced(trace["after_update"])
    s.applyBoundaryVelocityDamping(bmin, bmax)
    got = s.download_particles()
    want = trace["after_damping"]
    rep["damping_vel_rel"] = rel_err(got["velocity"], want["velocity"], vscale)
    rep["damping_touched"] = int(np.any(trace["after_damping"]["velocity"] != trace["after_update"]["velocity"], axis=1).sum())

    # ---- K11-K13 from the oracle's post-damping state: needs neighbour lists for that ordering. The post-damping state
    # is in the oracle's sorted order; its new_position changed since the lists were built, so rebuild lists from the
    # PRE-solve x* (after_neighbours state), then swap in the post-damping fields.
    pre = trace["after_neighbours"]
    s.upload_particles(pre)
    s.findParticleNeighbours(bmin, bmax)          # lists frozen from the initial x*, as in the reference
    order = s.debug(DBG.ID).astype(np.int64)      # our sorted slot -> slot in `pre` (identity in REFERENCE_HASH mode)
    post = trace["after_damping"][order]          # same particles, our order, post-damping fields
    # keep our lists, replace the per-particle state
    s.upload_particles(np.ascontiguousarray(post))
    s.applyVorticityAndViscosity(dt)
    got = s.download_particles()
    want = trace["after_vv"][order]
    rep["vv_vel_rel"] = rel_err(got["velocity"], want["velocity"], vscale)
    wscale = max(float(np.abs(want["vorticity"]).max()), 1e-30)
    rep["vv_vorticity_rel"] = rel_err(got["vorticity"], want["vorticity"], wscale)
    s.close()
    return rep


def trajectory_report(init, bmin, bmax, params, dt, golden: dict, steps=(1, 10), key_mode=KEY_LINEAR_CELL,
                      fast_math=True) -> dict:
    """Free-running steps of the CUDA solver compared with golden step snapshots (matched by particle id)."""
    n = len(init)
    s = make_solver(n, params, key_mode, fast_math=fast_math)
    s.upload_particles(init)
    ids_upload = ids_of(init)
    rep = {}
    for k in range(1, max(steps) + 1):
        s.step(dt, bmin, bmax)
        if k in steps:
            pos4, vel4, pid = s.download()
            ids = ids_upload[pid.astype(np.int64)]
            o_got = np.argsort(ids)
            o_want = np.argsort(golden[f"step{k}_id"])
            dp = np.abs(pos4[o_got, :3].astype(np.float64) - golden[f"step{k}_position"][o_want])
            dv = np.abs(vel4[o_got, :3].astype(np.float64) - golden[f"step{k}_velocity"][o_want])
            rep[f"step{k}_pos_max_rel_h"] = float(dp.max() / H)
            rep[f"step{k}_pos_rms_rel_h"] = float(np.sqrt((dp ** 2).mean()) / H)
            rep[f"step{k}_vel_max_rel"] = float(dv.max() / (H / dt))
            rep[f"step{k}_vel_rms_rel"] = float(np.sqrt((dv ** 2).mean()) / (H / dt))
            rho = vel4[o_got, 3]
            rep[f"step{k}_density_rel"] = rel_err(rho, golden[f"step{k}_density"][o_want])
    s.close()
    return rep

"""GPU check of the opt-in two-phase ("mask") neighbour-list build (options.list_build, akuaengine_b200/csrc/list_build.cuh).

k_build_neighbours_mask<4|8> must leave exactly the lists, counts and — after a run — exactly the particle state that the
default k_build_neighbours<KEY_LINEAR> leaves: it visits the same candidates in the same order (the order
kernel_find_neighbours does, src/CUDA/NeighbourSearchCUDA.cu:72-130). The same per-particle function is checked on the
CPU in tests/test_list_build_host.py. This file sorts last on purpose: the variants are not the default.
"""
import numpy as np
import pytest

from akuaengine_b200 import (DBG, KEY_LINEAR_CELL, LIST_BUILD_MASK4, LIST_BUILD_MASK8, LIST_BUILD_SCAN, PBFConfig, PBFSolver,
                             scenes)

pytestmark = pytest.mark.gpu

VARIANTS = [LIST_BUILD_SCAN, LIST_BUILD_MASK4, LIST_BUILD_MASK8]


def _scene(name):
    if name == "dam break 20^3":
        init, bmin, bmax = scenes.dam_break(20)
        return init, bmin, bmax, 128
    rng = np.random.default_rng(11)
    bmin, bmax = np.array([1.5, 0, 1.5], np.float32), np.array([4.5, 4, 4.5], np.float32)
    if name == "dense blob (cap 128 bites, rows > 32 candidates)":
        pos = (np.array([3.0, 2.0, 3.0]) + rng.normal(0, 0.07, (6000, 3))).astype(np.float32)
        return scenes.particles_from_positions(pos), bmin, bmax, 128
    if name == "jittered block, cap 20":
        pos = scenes._lattice(16, 14, 15, np.array([1.52, 0.02, 1.52], np.float32))
        pos = (pos + rng.uniform(-0.02, 0.02, pos.shape)).astype(np.float32)
        return scenes.particles_from_positions(pos), bmin, bmax, 20
    if name == "cloud partly outside the grid":
        pos = rng.uniform([1.0, -0.5, 1.0], [3.0, 1.0, 3.0], (20000, 3)).astype(np.float32)
        return scenes.particles_from_positions(pos), bmin, bmax, 128
    raise KeyError(name)


@pytest.mark.parametrize("name", ["dam break 20^3", "dense blob (cap 128 bites, rows > 32 candidates)",
                                  "jittered block, cap 20", "cloud partly outside the grid"])
def test_list_build_variants_identical(name):
    init, bmin, bmax, cap = _scene(name)
    cfg = PBFConfig()
    cfg.maxNeighbours = cap
    lists, states = {}, {}
    for v in VARIANTS:
        for use_graph in (False, True):
            s = PBFSolver(len(init), config=cfg, key_mode=KEY_LINEAR_CELL, list_build=v, use_graph=use_graph)
            s.upload_particles(init)
            s.step(0.0083, bmin, bmax)
            cnt = s.debug(DBG.NBR_COUNT).copy()
            lst = s.debug(DBG.NBR_LIST).copy()
            mask = np.arange(lst.shape[1])[None, :] < cnt[:, None]
            lists[(v, use_graph)] = (cnt, np.where(mask, lst, 0xffffffff))
            for _ in range(4):
                s.step(0.0083, bmin, bmax)
            states[(v, use_graph)] = s.download_particles().tobytes()
            s.close()
    ref_cnt, ref_lst = lists[(LIST_BUILD_SCAN, False)]
    assert ref_cnt.max() <= cap
    for k, (cnt, lst) in lists.items():
        assert np.array_equal(cnt, ref_cnt), ("neighbour counts differ", k)
        assert np.array_equal(lst, ref_lst), ("neighbour lists differ", k)
    for k, st in states.items():
        assert st == states[(LIST_BUILD_SCAN, False)], ("state after 5 steps differs", k)

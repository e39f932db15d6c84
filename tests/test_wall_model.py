"""The OPT-IN boundary-handling upgrade (akua_pbf_options::wall_model = 1, akuaengine_b200/csrc/wall_model.cuh),
run on the CPU: the product's host solver and kernels compiled by g++ against the SIMT emulator of tests/emu (the same build
tests/test_emu_slab.py uses), driven through the real C ABI.

The reference has no counterpart (its author's note at src/CUDA/ConstraintSolverCUDA.cu:132-135: "Later we will use virtual
particles"), so there is no oracle: the closed forms are checked against numerical quadrature of the reference's own kernels
(SmoothingKernelsCUDA.h:16-28), the default mode is checked to be untouched, and the effect is checked on a resting block.
"""
import numpy as np
import pytest

from akuaengine_b200 import DBG, PBFSolver, scenes
from test_emu_slab import emulib  # noqa: F401  (fixture: builds / loads tests/emu/_build/libakua_pbf_emu.so)

H, RHO0, DT = np.float32(0.1), np.float32(7600.0), 0.0083
PI_REF = 3.14  # the reference's pi (SmoothingKernelsCUDA.h:20,27)
C6 = 315.0 / (64.0 * PI_REF * float(H) ** 9)
CS = -45.0 / (PI_REF * float(H) ** 6)


def quad_half_space(d, n=400):
    """Numerical integrals over the half-space at distance d below a particle: of W_poly6 (density factor F) and of the normal
    component of grad W_spiky (magnitude S), in cylindrical coordinates (t = depth, s = radius), midpoint rule."""
    h = float(H)
    if d >= h:
        return 0.0, 0.0
    t = d + (np.arange(n) + 0.5) * (h - d) / n
    F = S = 0.0
    for tk in t:
        smax = np.sqrt(max(h * h - tk * tk, 0.0))
        s = (np.arange(n) + 0.5) * smax / n
        r2 = tk * tk + s * s
        r = np.sqrt(r2)
        w = C6 * np.maximum(h * h - r2, 0.0) ** 3
        g = abs(CS) * np.maximum(h - r, 0.0) ** 2 * (tk / r)
        F += (w * 2 * np.pi * s).sum() * (smax / n)
        S += (g * 2 * np.pi * s).sum() * (smax / n)
    return F * (h - d) / n, S * (h - d) / n


def one_particle(lib, d, wall_model):
    """A single particle at height d above the floor of a roomy box (no other wall within h), one solver iteration, no gravity."""
    p = scenes.particles_from_positions(np.array([[2.0, d, 2.0]], np.float32))
    p["new_position"] = p["position"]      # the phase-level operators work on the predicted position x*
    bmin, bmax = np.array([0, 0, 0], np.float32), np.array([4, 4, 4], np.float32)
    s = PBFSolver(1, lib=lib, use_graph=False, wall_model=wall_model, fast_math=False)
    s.upload_particles(p)
    s.setGravity(np.zeros(3, np.float32))
    s.findParticleNeighbours(bmin, bmax)
    s.runConstraintSolver(1, bmin, bmax)
    rho, lam = float(s.debug(DBG.DENSITY)[0]), float(s.debug(DBG.LAMBDA)[0])
    dp = s.download_particles()["position_delta"][0].astype(np.float64)
    s.close()
    return rho, lam, dp


@pytest.mark.parametrize("d", [0.0, 0.02, 0.05, 0.08, 0.0999, 0.15])
def test_wall_density_and_gradient_match_quadrature(emulib, d):
    rho, lam, dp = one_particle(emulib, d, 1)
    F, S = quad_half_space(d)
    self_w = C6 * float(H) ** 6                      # m W(0), m = 1
    assert rho == pytest.approx(self_w + float(RHO0) * F, rel=2e-3, abs=1e-2)
    # lambda = -C / (|grad_i C|^2 + relaxation), grad_i C = (1 / rho0) * (-rho0 S n) ; delta-p = (1 / rho0) lambda (-rho0 S n) = -lambda S n
    C = rho / float(RHO0) - 1.0
    assert lam == pytest.approx(-C / (S * S + 600.0), rel=5e-3, abs=1e-9)
    assert dp[0] == 0.0 and dp[2] == 0.0
    assert dp[1] == pytest.approx(-lam * S, rel=5e-3, abs=1e-9)
    # an isolated particle is under-dense: the constraint draws it TOWARDS the virtual fluid, never through the wall clamp
    if d < float(H):
        assert dp[1] <= 0.0
    # half a kernel's worth of fluid at the wall itself: F(0) = 1/2 (up to the reference's pi = 3.14)
    if d == 0.0:
        assert F == pytest.approx(0.5 * np.pi / PI_REF, rel=2e-3)


def test_reference_mode_ignores_the_walls(emulib):
    rho, lam, dp = one_particle(emulib, 0.02, 0)
    assert rho == pytest.approx(C6 * float(H) ** 6, rel=1e-6)        # only the self term, like the reference
    assert np.all(dp == 0.0)


def _resting_block(lib, wall_model, steps=25):
    """10 x 8 x 10 lattice block sitting on the floor of a box that hugs it on four sides: every face but the top is a wall."""
    pos = scenes._lattice(10, 8, 10, np.array([0.05, 0.05, 0.05], np.float32))
    p = scenes.particles_from_positions(pos)
    bmin, bmax = np.array([0.0, 0.0, 0.0], np.float32), np.array([0.6, 1.0, 0.6], np.float32)
    s = PBFSolver(len(p), lib=lib, use_graph=False, wall_model=wall_model)
    s.upload_particles(p)
    for _ in range(steps):
        s.step(DT, bmin, bmax)
    pos4, vel4, pid = s.download()
    rho = s.debug(DBG.DENSITY)
    s.close()
    o = np.argsort(pid)
    return pos4[o, :3], rho[o]


def test_virtual_fluid_walls_stop_the_crowding_at_the_floor(emulib):
    """Without a boundary model the bottom of a resting block misses half its neighbourhood, reads a density deficit, and the
    constraint fills it the only way it can: by pulling more fluid down until two lattice layers sit within h / 2 of the floor,
    held out of the wall by the soft clamp alone. With virtual fluid behind the walls the deficit is gone and the layers keep
    their distance."""
    x0, rho0 = _resting_block(emulib, 0)
    x1, rho1 = _resting_block(emulib, 1)
    assert np.isfinite(x1).all() and np.isfinite(rho1).all()
    bmin, bmax = np.array([0.0, 0.0, 0.0]), np.array([0.6, 1.0, 0.6])
    for x in (x0, x1):                      # nobody leaves the box in either mode (the soft clamp stays in place)
        assert np.all(x > bmin - 0.03) and np.all(x < bmax + 0.03)
    crowd0, crowd1 = int((x0[:, 1] < 0.06).sum()), int((x1[:, 1] < 0.06).sum())
    assert crowd0 > 180 and crowd1 < 0.8 * crowd0, (crowd0, crowd1)       # one lattice layer is 100 particles
    # mean height of the eight lattice layers (ids are x-outer, y, z-inner): the first gap nearly closes without the wall model
    lay0, lay1 = x0[:, 1].reshape(10, 8, 10).mean(axis=(0, 2)), x1[:, 1].reshape(10, 8, 10).mean(axis=(0, 2))
    assert lay1[1] - lay1[0] > 1.3 * (lay0[1] - lay0[0]), (lay0, lay1)
    assert lay1[-1] > lay0[-1]                                              # the block stands taller: less fluid squeezed into the floor
    # the bulk is left alone: density-constraint error of the whole block within a factor of 1.3 of the reference mode's
    assert np.abs(rho1 / RHO0 - 1.0).mean() < 1.3 * np.abs(rho0 / RHO0 - 1.0).mean()


def test_wall_model_is_rejected_in_slab_mode(emulib):
    from akuaengine_b200 import AkuaError
    s = PBFSolver(64, lib=emulib, use_graph=False, wall_model=1, capacity_factor=2.0)
    with pytest.raises(AkuaError):
        s.comm_init(0, 1, PBFSolver.comm_unique_id(emulib))
    s.close()


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="opt-in feature written after the round's GPU minutes were spent: its first run on hardware is "
                                        "this one — recorded (XPASS / XFAIL), not gating. The CPU-emulated tests above are the gate.")
def test_wall_model_on_the_gpu_matches_the_closed_form_and_leaves_the_default_alone():
    d = 0.03
    rho, lam, dp = one_particle(None, d, 1)
    F, S = quad_half_space(d)
    assert rho == pytest.approx(C6 * float(H) ** 6 + float(RHO0) * F, rel=2e-3)
    assert dp[1] == pytest.approx(-lam * S, rel=5e-3, abs=1e-9)
    rho0, _, dp0 = one_particle(None, d, 0)
    assert rho0 == pytest.approx(C6 * float(H) ** 6, rel=1e-6) and np.all(dp0 == 0.0)
    x0, r0 = _resting_block(None, 0)
    x1, r1 = _resting_block(None, 1)
    assert int((x1[:, 1] < 0.06).sum()) < 0.8 * int((x0[:, 1] < 0.06).sum())

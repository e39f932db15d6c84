"""Closed-form unit checks of the CPU port's building blocks (hash, SPH kernels, collision clamp, damping), against
float64 evaluations of the formulas as the reference writes them."""
import numpy as np

from akuaengine_b200 import PARTICLE_DTYPE
from oracle import PortOracle, param_block

F = np.float32
H = 0.1
PI = 3.14  # the reference's pi (SmoothingKernelsCUDA.h:20,27)


def two_particles(d, m1=1.0, m2=2.0):
    p = np.zeros(2, PARTICLE_DTYPE)
    p["position"][0] = [2.0, 2.0, 2.0]
    p["position"][1] = [2.0 + d, 2.0, 2.0]
    p["new_position"] = p["position"]
    p["mass"] = [m1, m2]
    return p


def poly6(d2):
    return 0.0 if d2 > H * H else 315.0 / (64.0 * PI * H ** 9) * (H * H - d2) ** 3


def spiky(r):
    return 0.0 if (r > H or r < 1e-5) else -45.0 / (PI * H ** 6) * (H - r) ** 2


def test_hash_matches_reference_formula():
    # NeighbourSearchCUDA.cu:15-27 with tableSize = maxNeighbours * n (PBFSolver.cpp:15)
    rng = np.random.default_rng(0)
    n = 64
    p = np.zeros(n, PARTICLE_DTYPE)
    p["new_position"] = rng.uniform(-3, 7, (n, 3)).astype(F)
    p["position"] = p["new_position"]
    o = PortOracle(p, param_block(gravity=(0, 0, 0)))
    o.findParticleNeighbours()
    q = o.particles
    cell = np.floor(q["new_position"] / F(H)).astype(np.int64)
    M = 1 << 32
    hx = (cell[:, 0] * 73856093) % M; hy = (cell[:, 1] * 19349663) % M; hz = (cell[:, 2] * 83492791) % M
    want = ((hx ^ hy ^ hz) % (128 * n)).astype(np.uint32)
    assert np.array_equal(q["hash"], want)
    assert np.all(np.diff(q["hash"].astype(np.int64)) >= 0)  # sorted by hash


def test_density_lambda_two_particles():
    d = 0.04
    o = PortOracle(two_particles(d), param_block())
    o.findParticleNeighbours()
    arr, cnt = o.neighbours()
    assert list(cnt) == [1, 1]
    o.runConstraintSolver(1, [0, 0, 0], [10, 10, 10])
    q = o.particles
    i0 = int(np.argmin(q["mass"]))  # the m=1 particle
    rho = 1.0 * poly6(0.0) + 2.0 * poly6(d * d)
    assert abs(q["density"][i0] / rho - 1) < 1e-5
    g = 2.0 * spiky(d) / 7600.0  # |grad_pi C| = |grad_pj C| for a single neighbour
    lam = -(rho / 7600.0 - 1.0) / (2 * g * g + 600.0)
    assert abs(q["lambda"][i0] / lam - 1) < 1e-4


def test_neighbour_threshold_is_strict_and_self_is_skipped():
    p = two_particles(0.2)  # farther than h: no neighbours
    o = PortOracle(p, param_block())
    o.findParticleNeighbours()
    assert list(o.neighbours()[1]) == [0, 0]
    p = two_particles(0.0)  # coincident: neighbours of each other (self is skipped by index, not by distance)
    o = PortOracle(p, param_block())
    o.findParticleNeighbours()
    assert list(o.neighbours()[1]) == [1, 1]


def test_collision_clamp_and_damping():
    # handle_particle_collision (ConstraintSolverCUDA.cu:136-157): x < min+0.025 -> x += 0.5*(min+0.025-x)
    p = np.zeros(1, PARTICLE_DTYPE)
    p["position"][0] = [1.0, 0.01, 1.0]; p["new_position"] = p["position"]; p["mass"] = 1
    o = PortOracle(p, param_block())
    o.findParticleNeighbours()
    o.runConstraintSolver(1, [0, 0, 0], [2, 2, 2])
    y = o.particles["new_position"][0, 1]
    assert abs(y - (0.01 + 0.5 * (0.025 - 0.01))) < 1e-7
    # resolve_collision (IntegrationCUDA.cu:51-73) with restitution 0, friction 0.95: approaching the floor
    p["position"][0] = [1.0, 0.01, 1.0]; p["velocity"][0] = [1.0, -2.0, 3.0]
    o = PortOracle(p, param_block())
    o.applyBoundaryVelocityDamping([0, 0, 0], [2, 2, 2])
    v = o.particles["velocity"][0]
    assert v[1] == 0.0 and abs(v[0] - (1 - F(0.95)) * 1.0) < 1e-7 and abs(v[2] - (1 - F(0.95)) * 3.0) < 1e-7
    # moving away from the wall: untouched
    p["velocity"][0] = [1.0, 2.0, 3.0]
    o = PortOracle(p, param_block())
    o.applyBoundaryVelocityDamping([0, 0, 0], [2, 2, 2])
    assert np.array_equal(o.particles["velocity"][0], F([1, 2, 3]))


def test_neighbour_cap_and_empty_input():
    n = 200  # all in one cell: every particle sees 199 candidates, capped at maxNeighbours = 128
    rng = np.random.default_rng(1)
    p = np.zeros(n, PARTICLE_DTYPE)
    p["position"] = (2.0 + rng.uniform(0.01, 0.04, (n, 3))).astype(F); p["new_position"] = p["position"]; p["mass"] = 1
    o = PortOracle(p, param_block())
    o.findParticleNeighbours()
    arr, cnt = o.neighbours()
    assert np.all(cnt == 128)
    # survivors are the first 128 candidates in bucket order (self skipped)
    assert list(arr[0, :3]) == [1, 2, 3] and list(arr[150, :3]) == [0, 1, 2]
    o0 = PortOracle(np.zeros(0, PARTICLE_DTYPE), param_block(maxNeighbours=128))
    o0.predictNewPosition(0.01)  # n = 0 must not crash

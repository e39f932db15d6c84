"""Deterministic synthetic scenes for the PBF step.

dam_break follows the reference's only scene, Application::prepareDamBreak (src/Application/Application.cpp:162-192):
a cubic lattice of spacing 0.05 with positions fl(min + fl(i * spacing)), x outermost / z innermost, mass 1, v = 0,
color blue, size 50, inside the box BOX_MIN/BOX_MAX (:14-15). Larger particle counts scale the scene constants by
s = n_side / 30 so the block keeps its place in the box.
"""
from __future__ import annotations

import numpy as np

from . import PARTICLE_DTYPE

F = np.float32


def _lattice(nx, ny, nz, origin, spacing=0.05):
    sp = F(spacing)
    xs = F(origin[0]) + np.arange(nx, dtype=F) * sp
    ys = F(origin[1]) + np.arange(ny, dtype=F) * sp
    zs = F(origin[2]) + np.arange(nz, dtype=F) * sp
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")  # x outermost, z innermost (Application.cpp:173-181)
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(F)


def particles_from_positions(pos: np.ndarray, mass: float = 1.0) -> np.ndarray:
    p = np.zeros(len(pos), dtype=PARTICLE_DTYPE)
    p["position"] = pos
    p["mass"] = F(mass)
    p["color"] = np.array([0.0, 0.0, 1.0, 1.0], dtype=F)  # Application.cpp:186
    p["size"] = F(50.0)                                    # Application.cpp:187
    return p


def dam_break(n_side: int = 30):
    """README scene at n_side=30 (27 000 particles); n_side=100 -> 1 M; 252 -> 16 M. Returns (particles, boxMin, boxMax)."""
    s = F(n_side / 30.0)
    # every scene constant of Application.cpp:14-15,162 is scaled by s, so the block keeps its place in the box and the
    # box keeps its distance from the world origin (the reference's hash double-counts neighbours in cells
    # with two zero coordinates, so scenes stay clear of the origin planes in x and z like the README scene does)
    box_min = (np.array([1.5, 0.0, 1.5], dtype=F) * s).astype(F)
    box_max = (np.array([4.5, 4.0, 4.5], dtype=F) * s).astype(F)
    origin = (np.array([2.0, 1.0, 2.0], dtype=F) * s).astype(F)
    pos = _lattice(n_side, n_side, n_side, origin)
    return particles_from_positions(pos), box_min, box_max


def dam_break_wide_positions(n_side: int, mult: int):
    """Weak-scaling extension of dam_break: the scene is `mult` times as long in x (box and block), i.e. a
    (mult*n_side) x n_side x n_side lattice — `mult` x the particles of dam_break(n_side), for x-slab runs on `mult` GPUs.
    Returns (positions float32 [N,3], boxMin, boxMax); mult = 1 is exactly dam_break(n_side)."""
    s = F(n_side / 30.0)
    box_min = (np.array([1.5, 0.0, 1.5], dtype=F) * s).astype(F)
    box_max = (np.array([4.5, 4.0, 4.5], dtype=F) * s).astype(F)
    box_max[0] = F(box_min[0] + (box_max[0] - box_min[0]) * F(mult))
    origin = (np.array([2.0, 1.0, 2.0], dtype=F) * s).astype(F)
    return _lattice(n_side * mult, n_side, n_side, origin), box_min, box_max


def lattice_x_columns(nx: int, origin_x, h: float = 0.1, spacing: float = 0.05):
    """Absolute x cell (floor of the float32 division, as the kernels compute it) of each lattice x index."""
    xs = F(origin_x) + np.arange(nx, dtype=F) * F(spacing)
    return np.floor(xs / F(h)).astype(np.int64)


def lattice_slab(nx: int, ny: int, nz: int, origin, ix0: int, ix1: int, spacing: float = 0.05):
    """Positions and global lattice ids of the x-index range [ix0, ix1) of an nx*ny*nz lattice (same values and the
    same x-outer / z-inner numbering as the full lattice), so each rank can generate only its own slab."""
    sp = F(spacing)
    xs = (F(origin[0]) + np.arange(nx, dtype=F) * sp)[ix0:ix1]
    ys = F(origin[1]) + np.arange(ny, dtype=F) * sp
    zs = F(origin[2]) + np.arange(nz, dtype=F) * sp
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    pos = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(F)
    ids = (np.arange(ix0 * ny * nz, ix1 * ny * nz, dtype=np.int64)).astype(np.uint32)
    return pos, ids


def dam_break_wide_layout(n_side: int, mult: int):
    """(lattice dims, origin, boxMin, boxMax) of dam_break_wide_positions without generating it."""
    s = F(n_side / 30.0)
    box_min = (np.array([1.5, 0.0, 1.5], dtype=F) * s).astype(F)
    box_max = (np.array([4.5, 4.0, 4.5], dtype=F) * s).astype(F)
    box_max[0] = F(box_min[0] + (box_max[0] - box_min[0]) * F(mult))
    origin = (np.array([2.0, 1.0, 2.0], dtype=F) * s).astype(F)
    return (n_side * mult, n_side, n_side), origin, box_min, box_max


def tank_layout(nx: int, ny: int, nz: int):
    sp = 0.05
    L = np.array([nx, ny, nz], dtype=np.float64) * sp
    off = np.array([1.5, 0.0, 1.5])
    box_min = off.astype(F)
    box_max = (off + np.array([1.25 * L[0], 2.0 * L[1], 1.05 * L[2]])).astype(F)
    origin = (off + 0.05).astype(F)
    return (nx, ny, nz), origin, box_min, box_max


def tank(nx: int, ny: int, nz: int):
    """Tank-slosh scene (config 4): lattice filling the lower part of a box of size (1.25 Lx, 2 Ly, 1.05 Lz);
    the slosh is driven through setGravity (see tank_gravity)."""
    sp = 0.05
    L = np.array([nx, ny, nz], dtype=np.float64) * sp
    off = np.array([1.5, 0.0, 1.5])  # keep clear of the origin planes in x and z, like the README box
    box_min = off.astype(F)
    box_max = (off + np.array([1.25 * L[0], 2.0 * L[1], 1.05 * L[2]])).astype(F)
    origin = (off + 0.05).astype(F)
    pos = _lattice(nx, ny, nz, origin)
    return particles_from_positions(pos), box_min, box_max


def tank_gravity(theta_deg: float = 15.0):
    t = np.deg2rad(theta_deg)
    return np.array([9.8 * np.sin(t), -9.8 * np.cos(t), 0.0], dtype=F)


def _splitmix64(state):
    state = (state + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = state
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return state, z ^ (z >> np.uint64(31))


def _uniform01(n, seed):
    """splitmix64 -> float in [0,1), vectorised (counter mode: stream element i uses state seed + (i+1)*gamma)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float64) / float(1 << 24)).astype(F)


def uniform_cloud(n: int, seed: int = 42, per_cell: float = 8.0, h: float = 0.1):
    """Neighbour-search microbench input: i.i.d. uniform positions at `per_cell` particles per h-cell (lattice density)."""
    side = (n / per_cell) ** (1.0 / 3.0) * h
    u = _uniform01(3 * n, seed).reshape(n, 3)
    pos = (u * F(side)).astype(F)
    return particles_from_positions(pos), np.zeros(3, dtype=F), np.full(3, side, dtype=F)


def clustered_cloud(n: int, blobs: int = 64, sigma_cells: float = 4.0, h: float = 0.1, per_cell: float = 8.0):
    """64 isotropic Gaussian blobs (sigma = 4h), centres uniform (seed 43), points (seed 44) clipped to the box,
    same mean density as uniform_cloud."""
    side = (n / per_cell) ** (1.0 / 3.0) * h
    centres = _uniform01(3 * blobs, 43).reshape(blobs, 3).astype(np.float64) * side
    u = _uniform01(7 * n, 44).reshape(n, 7).astype(np.float64)
    which = np.minimum((u[:, 0] * blobs).astype(np.int64), blobs - 1)
    # Box-Muller
    r1 = np.sqrt(-2.0 * np.log(np.maximum(u[:, 1], 1e-12)))
    r2 = np.sqrt(-2.0 * np.log(np.maximum(u[:, 3], 1e-12)))
    g = np.stack([r1 * np.cos(2 * np.pi * u[:, 2]), r1 * np.sin(2 * np.pi * u[:, 2]), r2 * np.cos(2 * np.pi * u[:, 4])], 1)
    pos = centres[which] + g * (sigma_cells * h)
    pos = np.clip(pos, 0.0, side * (1 - 1e-6)).astype(F)
    return particles_from_positions(pos), np.zeros(3, dtype=F), np.full(3, side, dtype=F)

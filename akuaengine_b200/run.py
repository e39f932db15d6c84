"""Headless scene runner: replaces Application::prepareDamBreak + the frame loop of
Application::run (src/Application/Application.cpp:37-70,126-204) by a JSON scene description.

    python -m akuaengine_b200.run scenes_json/config1_dambreak_27k.json [--steps N] [--checkpoint out.akpbf] [--resume in.akpbf]

Scene JSON keys (all optional except "scene"):
  "scene": "dam_break" | "tank" | "uniform_cloud" | "clustered_cloud";  "n_side" / "dims" / "n": size
  "dt": 0.0083, "steps": 100, "solver_iterations": 4, "key_mode": "linear" | "hash", "fast_math": true
  "config": {PBFConfig fields}, "corr": {LambdaCorrParams fields}
  "gravity_schedule": [[t0, gx, gy, gz], ...]   piecewise-constant gravity over simulated time (setGravity is the
                                                  reference's only runtime control, PBFSolver.h:18)
  "report_every": 10                              density-constraint error / timing report interval
"""
from __future__ import annotations

import argparse
import json
import sys
import time

import numpy as np

from . import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, LambdaCorrParams, PBFConfig, PBFSolver, scenes


def build_scene(desc: dict):
    kind = desc.get("scene", "dam_break")
    if kind == "dam_break":
        return scenes.dam_break(int(desc.get("n_side", 30)))
    if kind == "tank":
        nx, ny, nz = desc.get("dims", [40, 20, 20])
        return scenes.tank(nx, ny, nz)
    if kind == "uniform_cloud":
        return scenes.uniform_cloud(int(desc.get("n", 100000)), seed=int(desc.get("seed", 42)))
    if kind == "clustered_cloud":
        return scenes.clustered_cloud(int(desc.get("n", 100000)))
    raise ValueError(f"unknown scene {kind!r}")


def gravity_at(schedule, t, default):
    g = default
    for entry in schedule:
        if t >= entry[0]:
            g = entry[1:4]
    return g


def run(desc: dict, steps=None, checkpoint=None, resume=None, out=sys.stdout) -> dict:
    particles, bmin, bmax = build_scene(desc)
    cfg = PBFConfig(**desc.get("config", {}))
    if "solver_iterations" in desc:
        cfg.solverIterations = int(desc["solver_iterations"])
    corr = LambdaCorrParams(**desc.get("corr", {}))
    mode = KEY_LINEAR_CELL if desc.get("key_mode", "linear") == "linear" else KEY_REFERENCE_HASH
    solver = PBFSolver(len(particles), cfg, corr, key_mode=mode, fast_math=bool(desc.get("fast_math", True)))
    solver.upload_particles(particles)
    if resume:
        solver.load_checkpoint(resume)
    dt = float(desc.get("dt", 0.0083))
    steps = int(steps if steps is not None else desc.get("steps", 100))
    every = int(desc.get("report_every", 10))
    schedule = desc.get("gravity_schedule", [])
    default_g = list(cfg.gravity)
    t0 = solver.counters()["steps"] * dt
    last_g = None
    report = []
    solver.sync()
    wall = time.perf_counter()
    for k in range(steps):
        g = gravity_at(schedule, t0 + k * dt, default_g)
        if g != last_g:
            solver.setGravity(g)
            last_g = g
        solver.step(dt, bmin, bmax)
        if every and (k + 1) % every == 0:
            mean_err, max_err = solver.density_error()
            row = {"step": solver.counters()["steps"], "t": round(t0 + (k + 1) * dt, 6), "density_err_mean": mean_err,
                   "density_err_max": max_err}
            report.append(row)
            print(json.dumps(row), file=out)
    solver.sync()
    wall = time.perf_counter() - wall
    n = solver.n
    summary = {"particles": n, "steps": steps, "ms_per_step": wall / max(steps, 1) * 1e3,
               "particle_iterations_per_s": n * cfg.solverIterations * steps / wall if wall > 0 else None,
               "counters": solver.counters()}
    print(json.dumps(summary), file=out)
    if checkpoint:
        solver.save_checkpoint(checkpoint)
    solver.close()
    return {"report": report, "summary": summary}


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("scene_json")
    ap.add_argument("--steps", type=int)
    ap.add_argument("--checkpoint")
    ap.add_argument("--resume")
    a = ap.parse_args()
    with open(a.scene_json) as f:
        desc = json.load(f)
    run(desc, steps=a.steps, checkpoint=a.checkpoint, resume=a.resume)


if __name__ == "__main__":
    main()

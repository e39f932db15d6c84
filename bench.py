#!/usr/bin/env python
"""bench.py — headline benchmark of the PBF simulation step (BASELINE.json metric: particle-iterations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--n-side S] [--key-mode linear|hash]

A "step" is one call of the hot path (PBFSolver::step: predict, neighbour search, `solverIterations` constraint
iterations, commit, damping, vorticity confinement, XSPH) over the whole particle set of the named scene.
metric = N_particles * solverIterations * K / seconds, whole steps (all phases), aggregate over all ranks.

  value     : state resident in HBM when the timed region starts; K steps, CUDA events on the solver's stream.
  e2e       : the same K steps through the reference-facing C-ABI calls with HOST buffers: every step uploads the
              AoS-108 particle buffer from pinned host memory, steps, and downloads the AoS-108 buffer back.
  roofline  : dominant kernel (constraint pass B: delta-p + apply + collision): algorithmic bytes per launch (36 B per
              particle) / average launch duration measured live with CUDA events on the solver's stream.
  cpu_baseline : the host-C++ restatement (oracle/, kind "port") timed on this box's cores on a bounded sample.
  --impl reference : the UNMODIFIED reference kernels rebuilt headless (oracle/_ref, the reference has no CPU path);
              if that library is not present, the CPU port on a reduced sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

# The contract is ONE JSON line on stdout. Libraries loaded below (NCCL prints its version banner to stdout on some boxes)
# must not add to it: file descriptor 1 is pointed at stderr for the whole run and the JSON line is written to the saved
# descriptor.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(obj) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

DT = 0.0083
ITERS = 4
PASS_B_BYTES = 36   # R x* 16 + lambda 4 -> W x* 16 (phase D pass 2)
PASS_A_BYTES = 20   # R x* 16 -> W lambda 4
STEP_BYTES = 460 + 56 * ITERS


def measured_peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.idx = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_ours(args):
    import torch
    from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PARTICLE_DTYPE, PBFSolver, PinnedBuffer, scenes

    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU/oracle arm)")
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    key_mode = KEY_LINEAR_CELL if args.key_mode == "linear" else KEY_REFERENCE_HASH
    gravity = None
    if args.workload == "tank":
        # config 4: tank slosh; --tank-per-gpu => weak scaling (nx grows with the GPU count), --tank-total => strong
        if args.tank_total:
            dims = [int(v) for v in args.tank_total.split(",")]
            scaling = "strong"
        else:
            dims = [int(v) for v in args.tank_per_gpu.split(",")]
            dims[0] *= world
            scaling = "weak"
        (nx, ny, nz), origin, bmin, bmax = scenes.tank_layout(*dims)
        gravity = scenes.tank_gravity(15.0)
        scene_name = f"tank slosh {nx}x{ny}x{nz} lattice, gravity tilted 15 deg (SURVEY.md §8d config 4)"
    else:
        (nx, ny, nz), origin, bmin, bmax = scenes.dam_break_wide_layout(args.n_side, world)
        scaling = "weak"
        scene_name = (f"dam break {args.n_side}^3 lattice (SURVEY.md §8d config {'2' if args.n_side == 100 else 'n/a'})"
                      + (f", {world}x as long in x for {world} GPUs" if world > 1 else ""))
    n_total = nx * ny * nz
    if world == 1:
        pos, ids = scenes.lattice_slab(nx, ny, nz, origin, 0, nx)
        particles = scenes.particles_from_positions(pos)
        del pos
        n = len(particles)
        solver = PBFSolver(n, key_mode=key_mode, device=local, fast_math=bool(args.fast_math))
        solver.upload_particles(particles)
    else:
        # x-slab partition (one slab per GPU); ghost planes + migration go over NCCL/NVLink inside akua_pbf_step
        # (csrc/pbf_slab.inl). Each rank generates only its own slab of the lattice.
        from akuaengine_b200.slab import partition_columns, broadcast_unique_id
        cols1d = scenes.lattice_x_columns(nx, origin[0])
        col_min = int(cols1d.min())
        hist = np.bincount(cols1d - col_min).astype(np.int64) * (ny * nz)
        bounds = partition_columns(hist, world)
        lo, hi = col_min + int(bounds[rank]), col_min + int(bounds[rank + 1])
        sel = np.nonzero((cols1d >= lo) & (cols1d < hi))[0]
        pos, ids = scenes.lattice_slab(nx, ny, nz, origin, int(sel[0]), int(sel[-1]) + 1)
        particles = scenes.particles_from_positions(pos)
        del pos
        n = len(particles)
        solver = PBFSolver(max(n, n_total // world), key_mode=key_mode, device=local, fast_math=bool(args.fast_math),
                           capacity_factor=1.6)
        solver.comm_init(rank, world, broadcast_unique_id(dist, rank))
        solver.set_slab(lo, hi)
        solver.upload_particles(particles)
        solver.upload_ids(ids)
    del particles
    if gravity is not None:
        solver.setGravity(gravity)
    stream = torch.cuda.ExternalStream(solver.stream_ptr(), device=local)

    def barrier():
        solver.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        solver.step(DT, bmin, bmax)
    barrier()
    c0 = solver.counters()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        solver.step(DT, bmin, bmax)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    c1 = solver.counters()
    launches = c1["kernel_launches"] - c0["kernel_launches"]

    # ---- dominant-kernel timing, live, CUDA events around every pass-A / pass-B launch on the solver's stream ----
    solver.enable_timing(True)
    pa, pb, ph = [], [], None
    for _ in range(min(args.steps, 10)):
        solver.step(DT, bmin, bmax)
        t = solver.last_step_timing()
        pa.append(t["pass_a_sum"] / max(t["timed_iterations"], 1)); pb.append(t["pass_b_sum"] / max(t["timed_iterations"], 1))
        ph = t
    solver.enable_timing(False)
    pass_a_ms, pass_b_ms = float(np.mean(pa)), float(np.mean(pb))
    mean_err, max_err = solver.density_error()

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------
    cap = int(max(n, solver.n) * 1.6) + 1024
    pin = PinnedBuffer((cap,), PARTICLE_DTYPE)
    pin_ids = np.empty(cap, np.uint32)

    def e2e_step():
        m = solver.n                      # owned count can change through migration in slab mode
        solver.upload_particles(pin.array[:m])
        if world > 1:
            solver.upload_ids(pin_ids[:m])
        solver.step(DT, bmin, bmax)
        m = solver.n
        solver.download_particles(pin.array[:m])
        if world > 1:
            pin_ids[:m] = solver.debug(3)
        return m

    m = solver.n
    solver.download_particles(pin.array[:m])
    pin_ids[:m] = solver.debug(3)
    for _ in range(3):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    t0 = time.perf_counter()
    e0.record(stream)
    moved = 0
    for _ in range(e2e_steps):
        moved += e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    finite = bool(np.isfinite(pin.array["position"][:solver.n]).all())
    pin.free()

    # ---- max over ranks -----------------------------------------------------------------------------------------
    ms_step = ms_total / args.steps
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms = float(t[0]), float(t[1])
    total_particles = n_total
    n_owned_end = solver.n
    slab_stats = solver.slab_stats() if world > 1 else None
    value = total_particles * ITERS / (ms_step * 1e-3)
    e2e_value = total_particles * ITERS / (e2e_ms * 1e-3)

    peak, peak_src = measured_peaks()
    achieved = PASS_B_BYTES * n / (pass_b_ms * 1e-3) / 1e9
    out = {
        "metric": "particle-iterations/sec", "value": value, "unit": "particle-iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{scene_name}: {n_total} particles, dt={DT}, {ITERS} solver iterations, artificial pressure + "
                               "vorticity confinement + XSPH, box " + str([float(x) for x in bmin]) + "-" + str([float(x) for x in bmax]),
                   "particles_total": n_total, "particles_rank0": n, "key_mode": args.key_mode, "fast_math": bool(args.fast_math),
                   "list_build": {"0": "scan", "1": "mask4", "2": "mask8"}.get(os.environ.get("AKUA_LIST_BUILD", "0"), "scan"),
                   "l2": "no explicit flush: per-step working set (neighbour lists ~100 B/particle + 7 float4 arrays) "
                         f"= ~{(100 + 7 * 16 + 24) * n / 1e6:.0f} MB vs 126 MB L2",
                   "parallelism": "single GPU" if world == 1 else
                   f"{world} x-slabs (one per GPU), 1-cell ghost planes + per-step migration over NVLink inside akua_pbf_step "
                   "(CUDA-IPC P2P stores / copy-engine pushes; NCCL send/recv as fallback)"},
        "e2e": {"value": e2e_value, "unit": "particle-iterations/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": 108 * n_total, "d2h_bytes_per_step": 108 * n_total,
                "what": "akua_pbf_upload_aos108(pinned host) + akua_pbf_step + akua_pbf_download_aos108(pinned host) per step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_delta_apply (constraint pass B)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": PASS_B_BYTES, "launch_ms": pass_b_ms,
                     "launch_ms_source": "CUDA events recorded on the solver's stream around every pass-A / pass-B launch of "
                                         "10 further steps run right after the timed region (event pairs per launch "
                                         "would perturb the graph-replayed timed region itself)",
                     "pass_a": {"kernel": "k_density_lambda", "launch_ms": pass_a_ms,
                                "achieved": PASS_A_BYTES * n / (pass_a_ms * 1e-3) / 1e9},
                     "whole_step": {"algorithmic_bytes_per_particle": STEP_BYTES,
                                    "achieved": STEP_BYTES * n / (ms_step * 1e-3) / 1e9,
                                    "frac": STEP_BYTES * n / (ms_step * 1e-3) / 1e9 / peak}},
        "phases_ms": ph,
        "density_error": {"mean": mean_err, "max": max_err},
        "finite": finite,
    }
    if world > 1:
        out["slab_rank0"] = {"owned_start": n, "owned_end": n_owned_end, **slab_stats}
    prof = REPO / "profiles" / "r01_ncu_pass_b_traffic.json"
    if prof.exists():
        try:
            out["roofline"]["traffic"] = json.loads(prof.read_text()).get(str(n))
        except Exception:
            pass
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args)
        emit(out)
    solver.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(args, n_side=None, steps=8):
    """The CPU port (oracle/pbf_oracle.cpp, OpenMP) on a bounded sample of the same workload."""
    from akuaengine_b200 import scenes
    from oracle import PortOracle, param_block
    n_side = n_side or args.n_side
    particles, bmin, bmax = scenes.dam_break(n_side)
    o = PortOracle(particles, param_block())
    o.step(DT, bmin, bmax)  # warm (first touch of the 128*N table)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT, bmin, bmax)
    dt = time.perf_counter() - t0
    val = len(particles) * ITERS * steps / dt
    cores = o.threads
    o.close()
    return {"value": val, "unit": "particle-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps (after 1 warm-up step from rest) of the {len(particles)}-particle dam break, "
                      f"oracle/pbf_oracle.cpp with OpenMP on {cores} threads", "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path. Its solver is CUDA-only (no CPU path exists), so
    this runs the UNMODIFIED reference kernels rebuilt headless (oracle/_ref/libakua_ref.so) through PBFSolver::step on
    cuda:0, driven by one host thread. Falls back to the CPU port when that library did not travel."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from akuaengine_b200 import scenes
    from oracle import REF_LIB, PortOracle, RefOracle, param_block
    n_side = args.n_side
    use_ref = REF_LIB.exists()
    if use_ref:
        try:
            particles, bmin, bmax = scenes.dam_break(n_side)
            o = RefOracle(particles, param_block())
            kind, cores = "reference", 1
            where = "unmodified reference kernels (oracle/_ref) on cuda:0, 1 host thread; the reference has no CPU path"
        except Exception as e:  # no GPU
            use_ref = False
            why = str(e)
    if not use_ref:
        n_side = min(n_side, 50)  # bounded sample so K steps finish within minutes on CPU
        particles, bmin, bmax = scenes.dam_break(n_side)
        o = PortOracle(particles, param_block())
        kind, cores = "port", o.threads
        where = f"CPU port (oracle/pbf_oracle.cpp, OpenMP {cores} threads) on a reduced {n_side}^3 sample"
    n = len(particles)
    steps, warm = args.steps, args.warmup
    if use_ref:
        # the reference round-trips 2 KB/particle/step over PCIe: keep the whole run within a few minutes
        o.step(DT, bmin, bmax)
        t0 = time.perf_counter(); o.step(DT, bmin, bmax); one = time.perf_counter() - t0
        budget = 150.0
        if one * (steps + warm) > budget:
            warm = max(1, min(warm, int(0.2 * budget / one)))
            steps = max(2, min(steps, int(0.8 * budget / one)))
    for _ in range(warm):
        o.step(DT, bmin, bmax)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT, bmin, bmax)
    dt = time.perf_counter() - t0
    value = n * ITERS * steps / dt
    sample = f"{steps} timed steps after {warm} warm-up steps of the {n}-particle dam break; {where}"
    out = {"impl": "reference", "metric": "particle-iterations/sec", "value": value, "unit": "particle-iterations/s",
           "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": f"dam break {n} particles ({n_side}^3 lattice), dt={DT}, {ITERS} solver iterations",
                      "requested_steps": args.steps, "requested_warmup": args.warmup},
           "cpu_baseline": {"value": value, "unit": "particle-iterations/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "particle-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)
    o.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-side", type=int, default=100, help="lattice side: 30 -> 27 K (config 1), 100 -> 1 M (config 2), 252 -> 16 M (config 3)")
    ap.add_argument("--key-mode", default="linear", choices=["linear", "hash"])
    ap.add_argument("--workload", default="dam", choices=["dam", "tank"])
    ap.add_argument("--tank-per-gpu", default="200,200,200", help="tank lattice per GPU (weak scaling): nx,ny,nz")
    ap.add_argument("--tank-total", default="", help="tank lattice in total (strong scaling), e.g. 400,400,400 = 64 M")
    ap.add_argument("--fast-math", type=int, default=1, help="1: rsqrt-based spiky gradient (default); 0: IEEE sqrt/div")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

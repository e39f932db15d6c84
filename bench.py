#!/usr/bin/env python
"""bench.py — headline benchmark of the PBF simulation step (BASELINE.json metric: particle-iterations / second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Default workload = BASELINE.json config 4, weak scaling: tank slosh, a 200^3 lattice (8 M particles) PER GPU, x-slab
partitioned: 8 M on 1 GPU ... 64 M on 8 GPUs. (`--workload dam --n-side 100|252` = configs 2 / 3; the default run reports
them in an `extra` block at N = 1.)

A "step" is one call of the hot path (PBFSolver::step: predict, neighbour search, `solverIterations` constraint
iterations, commit, damping, vorticity confinement, XSPH) over the whole particle set.
metric = N_particles * solverIterations * K / seconds, whole steps (all phases), aggregate over all ranks.

Protocol (BASELINE.md §3 / SURVEY.md §8d): `--settle` steps (default 300) so that the fluid is disordered, W warm-up steps,
then `--windows` (default 5) timed windows of EXACTLY K steps each, every window bracketed by a barrier +
synchronize, CUDA events on the solver's stream, max over ranks; the MEDIAN window is the value, all windows are listed.

  value     : state resident in HBM when the timed region starts.
  e2e       : the same steps through the reference-facing C-ABI calls with HOST buffers: every step uploads the
              AoS-108 particle buffer from pinned host memory, steps, and downloads the AoS-108 buffer back.
  roofline  : constraint pass B (delta-p + apply + collision): algorithmic bytes per launch (36 B per particle,
              SURVEY.md §8d) / average launch duration measured live with CUDA events on the solver's stream.
  cpu_baseline : the host-C++ restatement (oracle/, kind "port") timed on this box's cores on a bounded sample.
  --impl reference : the UNMODIFIED reference kernels rebuilt headless (oracle/_ref; the reference has no CPU path) on
              the same workload as our N = 1 arm; if that library is not present, the CPU port on a reduced sample.
"""
from __future__ import annotations

import argparse
import hashlib
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
KERNEL_SOURCES = ["akuaengine_b200/csrc/pbf_kernels.cuh"]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peaks():
    f = REPO / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# The kernels the committed ncu captures describe (roofline.traffic / dram_frac_ncu): pass B and pass A of the default single-GPU step.
PROFILED_KERNELS = ["k_delta_applyILb1ELb0ELb1ELb1ELb0EE", "k_density_lambdaILb1ELb0EE"]


def kernel_source_hash() -> str:
    """Identity of the profiled kernels: a hash of their SASS (opcodes and operands of every instruction, addresses and encodings
    dropped) read from the built library with cuobjdump — an ncu capture stays valid exactly as long as the machine code of the
    kernels it describes is unchanged, whatever else in the source file moved. Falls back to a hash of the source file."""
    try:
        import re
        lib = REPO / "akuaengine_b200" / "libakua_pbf.so"
        txt = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True, timeout=120, check=True).stdout
        h = hashlib.sha256()
        found = 0
        for blk in re.split(r"\n\s*Function : ", txt)[1:]:
            name = blk.split("\n", 1)[0].strip()
            if not any(k in name for k in PROFILED_KERNELS):
                continue
            found += 1
            h.update(name.encode())
            for line in blk.split("\n"):
                m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(.*?);", line)
                if m:
                    h.update(m.group(1).strip().encode())
        if found == len(PROFILED_KERNELS):
            return "sass:" + h.hexdigest()[:16]
    except Exception:
        pass
    h = hashlib.sha256()
    for rel in KERNEL_SOURCES:
        h.update((REPO / rel).read_bytes())
    return "src:" + h.hexdigest()[:16]


def ncu_traffic(n_rank: int):
    """DRAM bytes per pass-B launch and its DRAM throughput from the committed `ncu --set full` captures
    (profiles/r02_ncu_traffic.json, written by tools/ncu_traffic_json.py): the capture nearest in size, its bytes PER PARTICLE
    scaled to this rank's particle count (beyond the L2 the sweeps' traffic per particle does not depend on n). Only trusted
    while the kernel sources it was captured from are unchanged."""
    f = REPO / "profiles" / "r02_ncu_traffic.json"
    if not f.exists():
        return None, None, "no committed ncu capture"
    try:
        d = json.loads(f.read_text())
    except Exception:
        return None, None, "unreadable profiles/r02_ncu_traffic.json"
    stale = d.get("kernel_source_sha") != kernel_source_hash()
    best = None
    for label, cap in d.get("captures", {}).items():
        row = cap.get("kernels", {}).get("k_delta_apply")
        if not row or not row.get("dram_bytes_per_particle"):
            continue
        dist = abs(np.log(max(cap["particles"], 1) / max(n_rank, 1)))
        if best is None or dist < best[0]:
            best = (dist, label, cap["particles"], row)
    if best is None or best[0] > np.log(2.5):
        return None, None, f"no capture within 2.5x of {n_rank} particles per GPU"
    _, label, n_cap, row = best
    return (int(row["dram_bytes_per_particle"] * n_rank), row.get("dram_pct_of_peak"),
            f"ncu --set full, {label} ({n_cap} particles): {row['dram_bytes_per_particle']:.1f} B per particle x {n_rank} particles"
            + ("; STALE: the profiled kernels changed since this capture" if stale else "; SASS of the profiled kernels unchanged since the capture"))


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
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            try:
                pw.append(float(f[3]))
            except ValueError:
                pass
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def bind_to_gpu_numa_node(local: int):
    """Pins this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is allocated, so
    that the e2e staging buffers of N ranks are spread over the host's memory controllers instead of all landing on node 0.
    Pure /sys reads; a no-op where the topology is not exposed."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = Path(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node")
        node = int(path.read_text().strip())
        if node < 0:
            return {"numa_node": None}
        cpus = Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        ids &= os.sched_getaffinity(0)
        if ids:
            os.sched_setaffinity(0, ids)
            return {"numa_node": node, "cpus": len(ids)}
        return {"numa_node": node, "cpus": 0}
    except Exception as e:  # no sysfs, no permission: run unbound
        return {"numa_node": None, "why": str(e)[:80]}


def nvlink_counters(local: int):
    """Cumulative NVLink payload counters of this rank's GPU, all links summed: (tx_bytes, rx_bytes), from NVML's
    NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX / _RX fields (KiB units). None where NVML does not expose them."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0".encode())
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xffffffff),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xffffffff)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                raise RuntimeError("field not supported")
            t = v.valueType
            raw = {0: v.value.dVal, 1: v.value.uiVal, 2: v.value.ulVal, 3: v.value.ullVal, 4: v.value.sllVal}.get(t, v.value.ullVal)
            out.append(int(raw) * 1024)
        return tuple(out)
    except Exception:
        pass
    try:   # the same counters through nvidia-smi (per link, KiB): "Link 0: Data Tx: 123 KiB"
        import re
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True, timeout=20).stdout
        tx = sum(int(v) for v in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
        rx = sum(int(v) for v in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
        if "Data Tx" not in txt:
            return None
        return tx * 1024, rx * 1024
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------- workload
def workload(args, world: int):
    """(lattice dims, origin, boxMin, boxMax, gravity, scaling, config dict). The config dict is the same object for our
    arm and for the reference arm (which always runs the world = 1 instance)."""
    from akuaengine_b200 import scenes
    gravity = None
    if args.workload == "tank":
        if args.tank_total:
            dims = [int(v) for v in args.tank_total.split(",")]
            scaling = "strong"
        else:
            dims = [int(v) for v in args.tank_per_gpu.split(",")]
            dims[0] *= world
            scaling = "weak"
        (nx, ny, nz), origin, bmin, bmax = scenes.tank_layout(*dims)
        gravity = scenes.tank_gravity(15.0)
        name = f"tank slosh {nx}x{ny}x{nz} lattice, gravity tilted 15 deg (BASELINE.json config 4, SURVEY.md §8d)"
    else:
        (nx, ny, nz), origin, bmin, bmax = scenes.dam_break_wide_layout(args.n_side, world)
        scaling = "weak"
        cfgno = {30: "1", 100: "2", 252: "3"}.get(args.n_side, "n/a")
        name = (f"dam break {args.n_side}^3 lattice (BASELINE.json config {cfgno})"
                + (f", {world}x as long in x for {world} GPUs" if world > 1 else ""))
    n_total = nx * ny * nz
    config = {
        "workload": f"{name}: {n_total} particles, dt={DT}, {ITERS} solver iterations, artificial pressure + vorticity "
                    f"confinement + XSPH, box {[round(float(x), 4) for x in bmin]}-{[round(float(x), 4) for x in bmax]}",
        "particles_total": n_total, "dt": DT, "solver_iterations": ITERS,
        "l2": "no explicit flush: the per-step working set (neighbour lists ~100 B/particle + 7 float4 arrays = "
              f"~{(100 + 7 * 16 + 24) * (n_total // world) / 1e6:.0f} MB per GPU) exceeds the 126 MB L2",
        "parallelism": "single GPU" if world == 1 else f"{world} x-slabs, one per GPU",
    }
    return (nx, ny, nz), origin, bmin, bmax, gravity, scaling, config


def make_solver(args, world, rank, local, dist, dims, origin, gravity, capacity_factor=2.5):
    """Creates the solver with this rank's share of the lattice uploaded. Returns (solver, n_local)."""
    from akuaengine_b200 import KEY_LINEAR_CELL, KEY_REFERENCE_HASH, PBFSolver, scenes
    nx, ny, nz = dims
    key_mode = KEY_LINEAR_CELL if args.key_mode == "linear" else KEY_REFERENCE_HASH
    if world == 1:
        pos, ids = scenes.lattice_slab(nx, ny, nz, origin, 0, nx)
        particles = scenes.particles_from_positions(pos)
        del pos
        n = len(particles)
        solver = PBFSolver(n, key_mode=key_mode, device=local, fast_math=bool(args.fast_math))
        solver.upload_particles(particles)
    else:
        # x-slab partition (one slab per GPU); ghost planes + migration go over NVLink inside akua_pbf_step
        # (csrc/pbf_slab.inl). Each rank generates only its own slab of the lattice.
        from akuaengine_b200.slab import broadcast_unique_id, partition_columns
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
        solver = PBFSolver(max(n, (nx * ny * nz) // world), key_mode=key_mode, device=local, fast_math=bool(args.fast_math),
                           capacity_factor=capacity_factor)
        solver.comm_init(rank, world, broadcast_unique_id(dist, rank))
        solver.set_slab(lo, hi)
        solver.upload_particles(particles)
        solver.upload_ids(ids)
    del particles
    if gravity is not None:
        solver.setGravity(gravity)
    return solver, n


def id_checksums(ids: np.ndarray):
    """Order-independent fingerprint of a multiset of particle ids: count, sum, sum of squares (mod 2^64), xor."""
    a = ids.astype(np.uint64)
    with np.errstate(over="ignore"):
        return np.array([len(a), int(a.sum(dtype=np.uint64)), int((a * a).sum(dtype=np.uint64)),
                         int(np.bitwise_xor.reduce(a)) if len(a) else 0], dtype=np.uint64)


def timed_windows(solver, stream, barrier, bmin, bmax, steps, windows, rebalance=0, owned_log=None, step_log=None):
    """`rebalance` = k > 0: multi-GPU runs call akua_pbf_rebalance every k steps, INSIDE the timed region (a sloshing scene shifts
    the load between slabs by several per cent within twenty steps; a production run re-balances at this rate and pays for
    it: one stream synchronisation and a small all-reduce per call, a graph re-capture only when a slab leaves its window)."""
    import torch
    out = []
    done = 0   # steps since the start of the timed region: the re-balancing cadence runs through the windows
    for _ in range(windows):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        marks = []
        for k in range(steps):
            solver.step(DT, bmin, bmax)
            done += 1
            if rebalance and done % rebalance == 0:
                solver.rebalance_async()   # no host synchronisation inside the timed region
            if step_log is not None:   # one event per step: shows a one-off cost (graph re-capture after a re-balance) as what it is
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                marks.append(ev)
        e1.record(stream)
        barrier()
        if step_log is not None:
            prev, row = e0, []
            for ev in marks:
                row.append(round(prev.elapsed_time(ev), 4))
                prev = ev
            step_log.append(row)
        if owned_log is not None:
            owned_log.append(int(solver.n))
        # events on the solver's stream; the host clock only matters when a host-side wait (the re-balancing collective) was not
        # covered by them
        out.append(max(e0.elapsed_time(e1), 0.0))
    return out


def run_ours(args):
    import torch
    from akuaengine_b200 import PARTICLE_DTYPE, PinnedBuffer

    rank, world, local = dist_env()
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU/oracle arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    dims, origin, bmin, bmax, gravity, scaling, config = workload(args, world)
    n_total = dims[0] * dims[1] * dims[2]

    # ---- multi-GPU self-check on a small instance of the same scene: N-slab result against a 1-GPU replica ----------
    mgpu_check = None
    if world > 1 and not args.no_selfcheck:
        from akuaengine_b200.slab import slab_selfcheck
        mgpu_check = slab_selfcheck(dist, rank, world, local, steps=24)
        log(f"[rank {rank}] slab self-check: {mgpu_check}")

    solver, n = make_solver(args, world, rank, local, dist, dims, origin, gravity)
    stream = torch.cuda.ExternalStream(solver.stream_ptr(), device=local)

    def barrier():
        solver.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident timing -------------------------------------------------------------------------------
    t_settle = time.perf_counter()
    for k in range(args.settle):
        solver.step(DT, bmin, bmax)
        if world > 1 and args.rebalance_every and (k + 1) % args.rebalance_every == 0:
            solver.rebalance_async()
    barrier()
    t_settle = time.perf_counter() - t_settle
    for _ in range(args.warmup):
        solver.step(DT, bmin, bmax)
    barrier()
    c0 = solver.counters()
    sampler = ClockSampler(local)
    sampler.start()
    owned_log = []
    nv0 = nvlink_counters(local) if world > 1 else None
    st0 = solver.slab_stats() if world > 1 else None
    wt0 = solver.slab_wait_stats() if world > 1 else None
    step_log = []
    win_ms = timed_windows(solver, stream, barrier, bmin, bmax, args.steps, args.windows,
                           rebalance=args.rebalance_every if world > 1 else 0, owned_log=owned_log, step_log=step_log)
    clocks = sampler.stop()
    nv1 = nvlink_counters(local) if world > 1 else None
    st1 = solver.slab_stats() if world > 1 else None
    wt1 = solver.slab_wait_stats() if world > 1 else None
    c1 = solver.counters()
    launches = (c1["kernel_launches"] - c0["kernel_launches"]) // args.windows
    graph_replays = (c1["graph_replays"] - c0["graph_replays"]) // args.windows

    # ---- optional launch timeline of the step right after the timed region (one file per rank) ----
    if args.trace:
        barrier()
        for _ in range(2):
            solver.trace_next_step(f"{args.trace}.rank{rank}.jsonl")
            solver.step(DT, bmin, bmax)
        barrier()

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
    n_rank = solver.n

    # ---- particle conservation across ranks (ids are a permutation of 0..n_total-1) ------------------------------
    conservation = None
    if world > 1:
        ids_now = solver.debug(3)
        fp = torch.from_numpy(id_checksums(ids_now).view(np.int64)).cuda()
        allfp = [torch.zeros_like(fp) for _ in range(world)]
        dist.all_gather(allfp, fp)
        got = np.stack([t.cpu().numpy().view(np.uint64) for t in allfp])
        with np.errstate(over="ignore"):
            tot = np.array([got[:, 0].sum(dtype=np.uint64), got[:, 1].sum(dtype=np.uint64), got[:, 2].sum(dtype=np.uint64),
                            np.bitwise_xor.reduce(got[:, 3])], dtype=np.uint64)
        # expected fingerprint of 0..n_total-1, in chunks to bound memory
        exp = np.zeros(4, np.uint64)
        with np.errstate(over="ignore"):
            for a in range(0, n_total, 1 << 24):
                c = id_checksums(np.arange(a, min(n_total, a + (1 << 24)), dtype=np.uint64))
                exp[0] += c[0]; exp[1] += c[1]; exp[2] += c[2]; exp[3] ^= c[3]
        errs = torch.tensor([mean_err * n_rank, max_err, float(n_rank)], device="cuda", dtype=torch.float64)
        allerr = [torch.zeros_like(errs) for _ in range(world)]
        dist.all_gather(allerr, errs)
        allerr = np.stack([t.cpu().numpy() for t in allerr])
        conservation = {"ids_are_a_permutation": bool(np.array_equal(tot, exp)), "owned_per_rank": [int(x) for x in got[:, 0]],
                        "density_error_mean_global": float(allerr[:, 0].sum() / max(allerr[:, 2].sum(), 1.0)),
                        "density_error_max_global": float(allerr[:, 1].max()),
                        "check": "count, sum, sum of squares and xor of all ranks' particle ids equal those of 0..n-1"}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------------
    cap = int(max(n, n_rank) * 2.0) + 1024
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / e2e_steps
    finite = bool(np.isfinite(pin.array["position"][:solver.n]).all())
    pin.free()

    # ---- max over ranks -----------------------------------------------------------------------------------------
    win = np.array(win_ms, dtype=np.float64) / args.steps
    if world > 1:
        t = torch.tensor(list(win) + [e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        win, e2e_ms = t[:-1].cpu().numpy(), float(t[-1])
        fin = torch.tensor([1 if finite else 0], device="cuda")
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
        finite = bool(int(fin))
        ol = torch.tensor(owned_log, device="cuda", dtype=torch.int64)
        allol = [torch.zeros_like(ol) for _ in range(world)]
        dist.all_gather(allol, ol)
        owned_log = np.stack([t.cpu().numpy() for t in allol], axis=1).tolist()   # [window][rank]
    ms_step = float(np.median(win))
    slab_stats = solver.slab_stats() if world > 1 else None
    value = n_total * ITERS / (ms_step * 1e-3)
    e2e_value = n_total * ITERS / (e2e_ms * 1e-3)

    peak, peak_src = measured_peaks()
    achieved = PASS_B_BYTES * n_rank / (pass_b_ms * 1e-3) / 1e9
    traffic, dram_pct, traffic_src = ncu_traffic(n_rank)
    out = {
        "metric": "particle-iterations/sec", "value": value, "unit": "particle-iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "protocol": {"settle_steps": args.settle, "settle_wall_s": round(t_settle, 2), "windows": args.windows,
                     "window_ms_per_step": [round(float(x), 5) for x in win], "value_is": "median window",
                     "owned_per_rank_after_each_window": owned_log if world > 1 else None,
                     "rank0_step_ms_in_each_window": step_log,
                     "simulated_time_at_start_s": round((args.settle + args.warmup) * DT, 3),
                     "rebalance": (f"akua_pbf_rebalance_async every {args.rebalance_every} steps, in the settle phase and inside the timed windows"
                                   if world > 1 and args.rebalance_every > 0 else "none")},
        "impl_config": {"key_mode": args.key_mode, "fast_math": bool(args.fast_math), "particles_rank0": int(n_rank),
                        "list_build": os.environ.get("AKUA_LIST_BUILD", "default"),
                        "cuda_graph_replays_per_window": int(graph_replays),
                        "transport": (slab_stats or {}).get("transport", "none (single GPU)"), "numa": numa},
        "e2e": {"value": e2e_value, "unit": "particle-iterations/s", "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": 108 * n_total, "d2h_bytes_per_step": 108 * n_total,
                "what": "akua_pbf_upload_aos108(pinned host) + akua_pbf_step + akua_pbf_download_aos108(pinned host) per step"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_delta_apply (constraint pass B)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "dram_frac_ncu": (dram_pct / 100.0) if dram_pct is not None else None,
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_particle": PASS_B_BYTES, "launch_ms": pass_b_ms,
                     "launch_ms_source": "CUDA events recorded on the solver's stream around every pass-A / pass-B launch of "
                                         "10 further steps run right after the timed region (event pairs per launch "
                                         "would perturb the graph-replayed timed region itself)",
                     "pass_a": {"kernel": "k_density_lambda", "launch_ms": pass_a_ms,
                                "achieved": PASS_A_BYTES * n_rank / (pass_a_ms * 1e-3) / 1e9},
                     "whole_step": {"algorithmic_bytes_per_particle": STEP_BYTES,
                                    "achieved": STEP_BYTES * n_total / world / (ms_step * 1e-3) / 1e9,
                                    "frac": STEP_BYTES * n_total / world / (ms_step * 1e-3) / 1e9 / peak}},
        "phases_ms": ph,
        "density_error": {"mean": mean_err, "max": max_err},
        "finite": finite,
    }
    if world > 1:
        # NVLink halo traffic against link bandwidth (north-star): what the solver counts per step (boundary-plane pushes +
        # migration records, device-side counter) beside the NVML link counters of rank 0's GPU over the same timed windows
        nsteps = args.steps * args.windows
        halo_bytes_step = (st1["bytes_sent"] - st0["bytes_sent"]) / max(nsteps, 1)
        link_peak = 900.0   # GB/s per direction per GPU (NVLink 5, 18 links x 50 GB/s)
        out["nvlink"] = {
            "rank": 0, "halo_bytes_sent_per_step_counted": halo_bytes_step,
            "exchanges_per_step": (st1["exchanges"] - st0["exchanges"]) / max(nsteps, 1),
            "nvml_tx_bytes_per_step": ((nv1[0] - nv0[0]) / max(nsteps, 1)) if nv0 and nv1 else None,
            "nvml_rx_bytes_per_step": ((nv1[1] - nv0[1]) / max(nsteps, 1)) if nv0 and nv1 else None,
            "link_peak_gbs_per_direction": link_peak,
            "link_utilisation_over_step": halo_bytes_step / (ms_step * 1e-3) / 1e9 / link_peak,
            "note": "end ranks have one neighbour, interior ranks two (twice the bytes); the pushes are P2P stores issued by the "
                    "boundary CTAs of each sweep while the interior CTAs of the same launch compute"}
        # where the ranks waited for each other during the timed windows (device clock): idle in the count exchange = load
        # imbalance; wait of the first boundary CTA for ghost planes = halo latency that was not hidden behind the interior
        dsteps = max(wt1["device_steps"] - wt0["device_steps"], 1)
        waits = torch.tensor([(wt1["plan_wait_ns"] - wt0["plan_wait_ns"]) / dsteps / 1e3, (wt1["halo_wait_ns"] - wt0["halo_wait_ns"]) / dsteps / 1e3],
                             device="cuda", dtype=torch.float64)
        allw = [torch.zeros_like(waits) for _ in range(world)]
        dist.all_gather(allw, waits)
        allw = np.stack([t.cpu().numpy() for t in allw])
        out["sync_waits"] = {"count_exchange_idle_us_per_step_by_rank": [round(float(x), 1) for x in allw[:, 0]],
                             "ghost_plane_wait_us_per_step_by_rank": [round(float(x), 1) for x in allw[:, 1]],
                             "exchanges_per_step": 12, "rebalances_that_moved_a_boundary": int(wt1["rebalances_moved"] - wt0["rebalances_moved"]),
                             "what": "device-clock time per step a rank idled for its neighbours' count message (k_slab_plan: a rank ahead of "
                                     "its neighbours waits here) and that the first boundary CTA of its 11 sweeps + the ghost-key kernel "
                                     "waited for ghost planes (halo_wait)"}
        out["slab_rank0"] = {"owned_start": n, "owned_end": int(n_rank), **slab_stats}
        out["mgpu_check"] = {"small_scene_vs_single_gpu": mgpu_check, "conservation": conservation}
    solver.close()
    if rank == 0:
        if world == 1 and not args.no_extra:
            out["extra"] = extra_configs(args, local)
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(args)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extra_configs(args, local):
    """BASELINE.json configs 2 and 3 (dam break 1 M / 16 M on one GPU), device-resident, one window each."""
    import torch
    from akuaengine_b200 import PBFSolver, scenes
    res = {}
    for name, side, settle, steps in (("config2_dam_break_1m", 100, 300, 50), ("config3_dam_break_16m", 252, 60, 10)):
        try:
            p, bmin, bmax = scenes.dam_break(side)
            s = PBFSolver(len(p), device=local, fast_math=bool(args.fast_math))
            s.upload_particles(p)
            n = len(p)
            del p
            stream = torch.cuda.ExternalStream(s.stream_ptr(), device=local)
            for _ in range(settle):
                s.step(DT, bmin, bmax)
            s.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                s.step(DT, bmin, bmax)
            e1.record(stream)
            s.sync()
            ms = e0.elapsed_time(e1) / steps
            me, mx = s.density_error()
            res[name] = {"particles": n, "ms_per_step": ms, "value": n * ITERS / (ms * 1e-3), "unit": "particle-iterations/s",
                         "settle_steps": settle, "steps": steps, "density_error_mean": me}
            s.close()
        except Exception as e:  # never lose the headline line to an extra
            res[name] = {"error": str(e)[:200]}
    # BASELINE.json config 5: neighbour-search microbench (key + sort + ranges + reorder, and the list build apart) at 1 M and
    # 16 M particles, uniform vs clustered (the full 1 M - 128 M table: profiles/r02_neighbour_search_microbench.jsonl)
    try:
        from akuaengine_b200 import DBG
        rows = []
        for m in (1, 16):
            n = m * 1_000_000
            for kind in ("uniform", "clustered"):
                p, bmin, bmax = scenes.uniform_cloud(n, seed=42) if kind == "uniform" else scenes.clustered_cloud(n)
                pos = p["position"].copy()
                del p
                s = PBFSolver(n, device=local)
                s.upload(pos)
                s.enable_timing(True)
                t = []
                for _ in range(4):
                    s.findParticleNeighbours(bmin, bmax)
                    t.append(s.last_step_timing())
                med = {k: float(np.median([r[k] for r in t[1:]])) for k in ("predict_key", "sort", "reorder_ranges", "neighbour_lists")}
                search = med["predict_key"] + med["sort"] + med["reorder_ranges"]
                passes = s.counters()["sort_passes_last"]
                cnt = s.debug(DBG.NBR_COUNT)
                rows.append({"n": n, "density": kind, "search_ms": search, "list_build_ms": med["neighbour_lists"], "ms": med,
                             "search_particles_per_s": n / (search * 1e-3), "sort_passes": passes,
                             "search_algorithmic_GBps": (24 + 16 * passes + 104) * n / (search * 1e-3) / 1e9,
                             "neighbours_mean": float(cnt.mean()), "particles_at_the_128_cap": int((cnt >= 128).sum())})
                s.close()
        res["config5_neighbour_search"] = rows
    except Exception as e:
        res["config5_neighbour_search"] = {"error": str(e)[:200]}
    return res


def cpu_baseline(args, steps=6):
    """The CPU port (oracle/pbf_oracle.cpp, OpenMP) on a bounded sample of the same workload family."""
    from akuaengine_b200 import scenes
    from oracle import PortOracle, param_block
    if args.workload == "tank":
        particles, bmin, bmax = scenes.tank(100, 100, 100)
        g = scenes.tank_gravity(15.0)
        what = "100x100x100 tank (1/8 of the per-GPU workload)"
    else:
        particles, bmin, bmax = scenes.dam_break(min(args.n_side, 100))
        g = None
        what = f"{min(args.n_side, 100)}^3 dam break"
    o = PortOracle(particles, param_block())
    if g is not None:
        o.setGravity(g)
    o.step(DT, bmin, bmax)  # warm (first touch of the 128*N table)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT, bmin, bmax)
    dt = time.perf_counter() - t0
    val = len(particles) * ITERS * steps / dt
    cores = o.threads
    o.close()
    return {"value": val, "unit": "particle-iterations/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps (after 1 warm-up step from rest) of a {len(particles)}-particle {what}, "
                      f"oracle/pbf_oracle.cpp with OpenMP on {cores} threads", "ms_per_step": dt / steps * 1e3}


def run_reference(args):
    """Reference arm: the reference's own implementation of the path. Its solver is CUDA-only (no CPU path exists), so
    this runs the UNMODIFIED reference kernels rebuilt headless (oracle/_ref/libakua_ref.so) through PBFSolver::step on
    cuda:0, driven by one host thread, on the SAME workload as our N = 1 arm (the reference is single-GPU and its
    `int tableSize = 128 * N` caps it at 16.7 M particles). Falls back to the CPU port when that library did not travel."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from akuaengine_b200 import scenes
    from oracle import REF_LIB, PortOracle, RefOracle, param_block
    dims, origin, bmin, bmax, gravity, scaling, config = workload(args, 1)
    use_ref = REF_LIB.exists()
    why = ""
    if use_ref:
        try:
            pos, _ = scenes.lattice_slab(*dims, origin, 0, dims[0])
            particles = scenes.particles_from_positions(pos)
            del pos
            o = RefOracle(particles, param_block())
            kind, cores = "reference", 1
            where = "unmodified reference kernels (oracle/_ref) on cuda:0, 1 host thread; the reference has no CPU path"
        except Exception as e:  # no GPU
            use_ref = False
            why = str(e)
    if not use_ref:
        # bounded sample so K steps finish within minutes on CPU: 1/64 (tank) of the workload
        if args.workload == "tank":
            sdims = [max(8, d // 4) for d in dims]
            (nx, ny, nz), origin, bmin, bmax = scenes.tank_layout(*sdims)
            pos, _ = scenes.lattice_slab(nx, ny, nz, origin, 0, nx)
            particles = scenes.particles_from_positions(pos)
        else:
            particles, bmin, bmax = scenes.dam_break(min(args.n_side, 50))
        o = PortOracle(particles, param_block())
        kind, cores = "port", o.threads
        where = (f"CPU port (oracle/pbf_oracle.cpp, OpenMP {cores} threads) on a reduced {len(particles)}-particle sample"
                 + (f" [{why[:80]}]" if why else ""))
    if gravity is not None:
        o.setGravity(gravity)
    n = len(particles)
    del particles
    steps, warm = args.steps, args.warmup
    if use_ref:
        # the reference round-trips 2 KB/particle/step over PCIe: keep the whole run within a few minutes
        o.step(DT, bmin, bmax)
        t0 = time.perf_counter(); o.step(DT, bmin, bmax); one = time.perf_counter() - t0
        budget = 170.0
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
    sample = f"{steps} timed steps after {warm + (2 if use_ref else 0)} warm-up steps from rest, {n} particles; {where}"
    out = {"impl": "reference", "metric": "particle-iterations/sec", "value": value, "unit": "particle-iterations/s",
           "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config if n == config["particles_total"] else dict(config, workload=config["workload"] + f" [REDUCED SAMPLE: {n} particles]", particles_total=n),
           "protocol": {"requested_steps": args.steps, "requested_warmup": args.warmup,
                        "note": "starts from the lattice at rest (the reference cannot afford our arm's settle steps: "
                                f"{dt / steps:.2f} s per step)"},
           "cpu_baseline": {"value": value, "unit": "particle-iterations/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "particle-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)
    o.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50, help="steps per timed window (BASELINE.md §3: >= 50)")
    ap.add_argument("--warmup", type=int, default=20, help="untimed steps right before the timed windows (BASELINE.md §3: >= 20)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tank", choices=["dam", "tank"])
    ap.add_argument("--tank-per-gpu", default="200,200,200", help="tank lattice per GPU (weak scaling): nx,ny,nz")
    ap.add_argument("--tank-total", default="", help="tank lattice in total (strong scaling), e.g. 400,400,400 = 64 M")
    ap.add_argument("--n-side", type=int, default=100, help="--workload dam: 30 -> 27 K (config 1), 100 -> 1 M (config 2), 252 -> 16 M (config 3)")
    ap.add_argument("--key-mode", default="linear", choices=["linear", "hash"])
    ap.add_argument("--settle", type=int, default=300, help="untimed steps before the warm-up, so that the fluid is disordered")
    ap.add_argument("--windows", type=int, default=5, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--rebalance-every", type=int, default=5, help="multi-GPU: akua_pbf_rebalance every k steps (settle phase and timed windows)")
    ap.add_argument("--fast-math", type=int, default=1, help="1: rsqrt-based spiky gradient (default); 0: IEEE sqrt/div")
    ap.add_argument("--trace", default="", help="write a launch timeline of two steps after the timed region to PATH.rank<r>.jsonl")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs 2 / 3 extra block")
    ap.add_argument("--no-selfcheck", action="store_true", help="multi-GPU: skip the small-scene check against one GPU")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    try:
        run_ours(args)
    except BaseException:
        # A rank that fails must not linger: its peers are waiting for it inside kernels and collectives, and the interpreter's
        # orderly shutdown (destructors synchronising streams whose kernels wait for those peers) can take minutes. Report and
        # leave at once, so that the launcher sees the failure and ends the other ranks.
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            os._exit(1)
        raise


if __name__ == "__main__":
    main()

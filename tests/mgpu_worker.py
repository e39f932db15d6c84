"""Multi-GPU worker (launched by torchrun, one rank per GPU): steps a scene with the x-slab solver and checks it against
a single-GPU run of the same library on rank 0. Exit code 0 = all checks passed.
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_worker.py [--scene dam|tank] [--steps 8]
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))

from akuaengine_b200 import PBFSolver, scenes  # noqa: E402
from akuaengine_b200.slab import setup_slab_solver  # noqa: E402

H, DT = 0.1, 0.0083


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="dam")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--side", type=int, default=40)
    ap.add_argument("--vx", type=float, default=0.0, help="initial x velocity of every particle (forces migration)")
    ap.add_argument("--rebalance-every", type=int, default=0)
    ap.add_argument("--skew", type=float, default=0.0, help="start from deliberately unbalanced slabs (fraction moved to rank 0)")
    ap.add_argument("--canonical", action="store_true", help="options.canonical_order on both sides: the result must be bit-identical")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.scene == "dam":
        particles, bmin, bmax = scenes.dam_break(args.side)
    else:
        particles, bmin, bmax = scenes.tank(2 * args.side, args.side // 2, args.side)
    n = len(particles)
    particles["velocity"][:, 0] = np.float32(args.vx)
    # a payload that must follow its particle through sorts and migrations (the reference sorts whole structs,
    # src/CUDA/NeighbourSearchCUDA.cu:167-170; the renderer reads color@88 / size@104, src/Rendering/Renderer.cpp:201-213)
    particles["color"][:, 0] = (np.arange(n) % 251).astype(np.float32)
    particles["size"] = (np.arange(n) % 17 + 1).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    if args.scene == "tank":
        g = scenes.tank_gravity(15.0)
    solver = setup_slab_solver(particles, ids, dist, rank, world, local, H, capacity_factor=2.0, skew=args.skew,
                               canonical_order=args.canonical)
    if args.scene == "tank":
        solver.setGravity(g)
    counts0 = solver.n
    start = torch.tensor([counts0], device="cuda", dtype=torch.int64)
    all_start = [torch.zeros_like(start) for _ in range(world)]
    dist.all_gather(all_start, start)
    all_start = [int(t) for t in all_start]
    for k in range(args.steps):
        solver.step(DT, bmin, bmax)
        if args.rebalance_every and (k + 1) % args.rebalance_every == 0:
            solver.rebalance()
    pos4, vel4, pid = solver.download()
    aos = solver.download_particles()
    st = solver.slab_stats()
    graph_replays = solver.counters()["graph_replays"]
    # gather everything on rank 0
    owned = torch.tensor([solver.n], device="cuda", dtype=torch.int64)
    all_owned = [torch.zeros_like(owned) for _ in range(world)]
    dist.all_gather(all_owned, owned)
    all_owned = [int(t) for t in all_owned]
    migt = torch.tensor([st["migrated_in"]], device="cuda", dtype=torch.int64)
    all_mig = [torch.zeros_like(migt) for _ in range(world)]
    dist.all_gather(all_mig, migt)
    mx = max(all_owned)
    pack = torch.zeros((mx, 11), dtype=torch.float64, device="cuda")
    pack[:solver.n, 0:4] = torch.from_numpy(pos4.astype(np.float64)).cuda()
    pack[:solver.n, 4:8] = torch.from_numpy(vel4.astype(np.float64)).cuda()
    pack[:solver.n, 8] = torch.from_numpy(pid.astype(np.float64)).cuda()
    pack[:solver.n, 9] = torch.from_numpy(aos["color"][:, 0].astype(np.float64)).cuda()
    pack[:solver.n, 10] = torch.from_numpy(aos["size"].astype(np.float64)).cuda()
    gathered = [torch.zeros_like(pack) for _ in range(world)]
    dist.all_gather(gathered, pack)
    ok = True
    if rank == 0:
        allp = np.concatenate([g[:c].cpu().numpy() for g, c in zip(gathered, all_owned)])
        got_ids = allp[:, 8].astype(np.int64)
        print(f"ranks own {all_owned} (start {counts0} on rank 0), stats rank0 {st}, graph replays rank0 {graph_replays}")
        if sum(all_owned) != n or not np.array_equal(np.sort(got_ids), np.arange(n)):
            print("FAIL: particles not conserved / ids not a permutation"); ok = False
        ref = PBFSolver(n, device=local, canonical_order=args.canonical)
        ref.upload_particles(particles)
        if args.scene == "tank":
            ref.setGravity(g)
        for _ in range(args.steps):
            ref.step(DT, bmin, bmax)
        rp, rv, rid = ref.download()
        o1 = np.argsort(got_ids); o2 = np.argsort(rid)
        d = np.abs(allp[o1, 0:3] - rp[o2, :3]).max(axis=1) / H
        dp, p99, rms = d.max(), np.quantile(d, 0.99), np.sqrt((d * d).mean())
        dv = np.abs(allp[o1, 4:7] - rv[o2, :3]).max() / (H / DT)
        drho = np.abs(allp[o1, 7] - rv[o2, 3]).max() / 7600.0
        mig_total = sum(int(t) for t in all_mig)
        # Without migration the N-GPU run is bit-identical to one GPU (same sort order, same list order). With migration the
        # neighbour ORDER differs (arrivals are appended), float sums differ in the last bit, and wall contacts amplify that for
        # a few particles (one GPU in its two key modes diverges just as much): statistical tolerance, like slab_selfcheck.
        # default order: a coarse sanity bound only (one GPU in its two key modes — same neighbour sets, different order —
        # differs by 3e-3 h p99 / 1.4e-3 h rms / 0.09 h max after 24 steps of this kind of scene, profiles/r02_order_sensitivity.json);
        # the exact statement is the --canonical run
        # (8 ranks, vx = 1.5, 20 steps on B200s: p99 1.7e-2 h, rms 6.6e-3 h, max 0.41 h — profiles/r02_c15_worker8_default_order.log)
        tol = 3e-2
        print(f"slab({world}) vs single GPU after {args.steps} steps: dpos/h max={dp:.3e} p99={p99:.3e} rms={rms:.3e} "
              f"dvel/(h/dt)={dv:.3e} drho/rho0={drho:.3e} migrated={mig_total}")
        if args.canonical or (mig_total == 0 and not args.rebalance_every):
            # canonical order: a cell's particles are ordered by id on every rank, so neighbour order (and every float sum) is
            # that of the single-GPU run whatever migrated or was re-balanced
            if not (dp == 0.0 and dv == 0.0 and drho == 0.0):
                print("FAIL: the slab result must be bit-identical to the single-GPU result (canonical order, or nothing migrated)"); ok = False
        elif not (p99 < tol and rms < 1.5e-2 and dp < 1.0):
            print("FAIL: slab result differs from the single-GPU result"); ok = False
        if not (np.array_equal(allp[o1, 9], particles["color"][:, 0].astype(np.float64))
                and np.array_equal(allp[o1, 10], particles["size"].astype(np.float64))):
            print("FAIL: the render payload (color / size) did not follow its particle"); ok = False
        if args.rebalance_every and world > 1:
            imb = max(all_owned) / (sum(all_owned) / world)
            imb0 = max(all_start) / (sum(all_start) / world)
            print(f"imbalance (heaviest slab / mean, particle counts): {imb0:.3f} at the start, {imb:.3f} after rebalancing")
            # two slabs re-balance within a few calls; with more ranks a skewed start drains rank by rank (a boundary only
            # moves inside the two slabs it separates per call), and the balance is by WORK, not by particle count
            # (work-balanced slabs are not count-balanced: a slab of deeper, denser fluid holds fewer particles)
            if (imb > 1.3) if world == 2 else (imb > 1.3 and imb >= imb0):
                print("FAIL: slabs not balanced after rebalancing"); ok = False
        if world > 1 and st["exchanges"] == 0:
            print("FAIL: no exchanges happened"); ok = False
        mig = sum(int(t) for t in all_mig)
        print("particles migrated between ranks:", mig)
        if args.vx != 0.0 and world > 1 and mig == 0:
            print("FAIL: expected migration with an initial x velocity"); ok = False
        ref.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    solver.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) == 1 else 1)


if __name__ == "__main__":
    main()

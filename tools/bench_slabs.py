#!/usr/bin/env python
"""BASELINE config 5: one large scene split into x-slabs across the GPUs of a box, halo exchange over peer memory (NVLink; `--nccl`: the
first transport, dense halos through torch.distributed P2P with one Python round trip per phase).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_slabs.py [--particles 8000000] [--grid 256] [--substeps 32]
  python tools/bench_slabs.py --check           # (under torchrun) small scene, compares against a single-handle run on rank 0

Strong scaling: the scene is fixed, every rank owns the particles of its slab.  Prints one JSON line on rank 0:
particle-substeps/s forward+backward, max over ranks of the device time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def slab_record(n, n_grid, S, rank, ws, local, reps=3, sort_every=16, check=False, migrate_every=0, sphere=False, drift=0.0, peer=True):
    """One strong-scaling measurement of the slab decomposition (the process group must exist when ws > 1).  Returns the record on
    rank 0 (None elsewhere); with check=True the forward states are compared with a single handle on rank 0."""
    import types
    import torch
    import torch.distributed as dist
    import scenes
    from harness import sim_cfg, rel_l2
    from softmac_b200.slabs import DistSlab, DistMigratingSlab
    args = types.SimpleNamespace(n=n, n_grid=n_grid, substeps=S, reps=reps, sort_every=sort_every, check=check, migrate_every=migrate_every,
                                 sphere=sphere, drift=drift)
    S = args.substeps
    dt = 2e-4 * 64 / args.n_grid * 0.5 if args.n_grid > 64 else 2e-4          # 1e-4 at 128^3, 5e-5 at 256^3 (stability: DESIGN.md section 5)
    cfg = sim_cfg(args.n, n_grid=args.n_grid, max_steps=S + 2, dt=dt)
    st = scenes.cube_state(args.n)                                              # same on every rank (np.random.seed(0))
    if args.check:
        st[:, 3:6] = 0.5 * np.random.default_rng(1).normal(size=(args.n, 3)).astype(np.float32)
    if args.drift:
        st[:, 3] += np.float32(args.drift)
    seed = st[:, :3] - st[:, :3].mean(0)
    E = args.migrate_every
    make_prims, s13 = None, None
    if args.sphere:
        from softmac_b200.engine import Primitives, Mesh
        tab = scenes.sphere_table()
        s13 = np.concatenate([[0.5, 0.3 - 0.39 / 2 - 0.05, 0.5], [1, 0, 0, 0], [0.0, 0.2, 0.0], [0, 0, 0.3]])     # just under the cube, moving up into it

        def make_prims():
            m = Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]), cfg=dict(friction=0.5), max_timesteps=S + 2)
            p = Primitives(primitives=[m], max_timesteps=S + 2)
            p.initialize()
            return p
    if ws > 1 and E:
        sl = DistMigratingSlab(cfg, st, E, env_dt=5 * dt, sort_every=args.sort_every, make_primitives=make_prims)
    elif ws > 1:
        sl = DistSlab(cfg, st, env_dt=5 * dt, sort_every=args.sort_every, make_primitives=make_prims, peer=peer)
    if args.sphere and ws > 1:
        if E:
            sl.set_primitive_state(0, 0, S + 2, s13); sl.clear_ext_f()
        else:
            sl.primitives[0].set_all_states(0, s13, f_end=S + 2); sl.primitives[0].clear_ext_f()
    from softmac_b200.engine import MPMSimulator
    if ws == 1:
        sim = MPMSimulator(cfg, (), env_dt=5 * dt, sort_every=args.sort_every)
        sim.reset(st)

        class _One:
            def step(self, s0, c): sim.step(s0, c)
            def step_grad(self, s1, c): sim.step_grad(s1, c)
            def add_x_grad(self, f, g): sim.add_x_grad(f, g)
        sl, the_sim = _One(), sim
        sl.sim = sim
    migrating = ws > 1 and E > 0
    if migrating:
        sim = sl.r.epochs[0].sim                    # first epoch: its spare frame E + 1 keeps the initial state
        sim.copyframe(0, E + 1)

        seed_dev = torch.zeros((args.n, 24), dtype=torch.float32, device="cuda")      # the dense seed, resident like the states
        seed_dev[:, :3] = torch.as_tensor(seed.astype(np.float32), device="cuda")

        def step():
            sl.rewind()
            sl.clear_all_gradients()
            sim.copyframe(E + 1, 0)
            sl.step(0, S)
            sl.add_state_grad_dev(S, seed_dev)          # the owners of the last frame differ from those at reset
            sl.step_grad(S, S)
    else:
        sim = sl.sim
        sim.copyframe(0, S + 1)
        sl.add_x_grad(S, seed)

        def step():
            sim.copyframe(S + 1, 0)
            sl.step(0, S)
            sl.step_grad(S, S)

    times = []
    for r in range(args.reps + 2):
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize(); sim.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize(); sim.synchronize()
        if ws > 1:
            dist.barrier()
        if r >= 2:
            times.append(time.perf_counter() - t0)
        if r == 0 and ws > 1 and not migrating:         # a neighbour that never answers: stop before the timed repetitions pile up timeouts
            bad = torch.tensor([float(sl.halo_status()["timeouts"])], dtype=torch.float64, device="cuda")
            dist.all_reduce(bad)
            if bad.item() > 0:
                return {"error": "halo exchange timed out on %d exchanges in the first pass" % int(bad.item()), "n_gpus": ws} if rank == 0 else None
    t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda")
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    T = float(t.item())
    out = {"workload": f"slab decomposition (config 5): {args.n} particles, {args.n_grid}^3, {S} substeps fwd + {S} bwd", "n_gpus": ws,
           "ms_per_step": T * 1e3, "particle_substeps_per_s_fwd_bwd": args.n * S / T, "scaling": "strong", "local_particles": int(sim.n_particles),
           "counters": sim.counters()}
    if ws > 1 and not migrating:
        hs = sl.halo_status()
        out["transport"] = ("peer memory: non-empty halo blocks pushed into the neighbour's receive slot over NVLink, flag / wait / add kernels on the "
                            "simulator's stream inside smx_step (one native call per rank and pass)") if sl.peer else "torch.distributed P2P (NCCL), dense 2-column halos, one Python round trip per phase"
        out["halo"] = hs
        if hs["timeouts"]:
            out["error"] = "a halo exchange timed out: the numbers of this record are invalid"
    if migrating:
        out["migrate_every"] = E
        out["migrated_particles_per_step"] = sl.migrated()
        out["local_particles_last_epoch"] = int(sl.r.epochs[-1].n)
    if args.check and ws > 1:
        got = sl.gather_state(S)
        if args.sphere:                 # the timed loop ran forward + backward several times: one clean forward for the wrench
            if migrating:
                sl.rewind(); sl.clear_all_gradients(); sim.copyframe(E + 1, 0)
            else:
                sim.copyframe(S + 1, 0)
            sl.clear_ext_f() if migrating else sl.primitives[0].clear_ext_f()
            sl.step(0, S)
            wrench = sl.ext_f(0)
            got = sl.gather_state(S)
        if rank == 0:
            pr = make_prims() if args.sphere else ()
            ref = MPMSimulator(cfg, pr, env_dt=5 * dt, sort_every=args.sort_every, device=local)
            if args.sphere:
                pr[0].set_all_states(0, s13, f_end=S + 2)
            ref.reset(st)
            if args.sphere:
                pr[0].clear_ext_f()
            ref.step(0, S)
            r = ref.get_state(S)
            if args.sphere:
                fe = pr[0].get_ext_f()
                out["check_wrench_rel_l2"] = rel_l2(wrench, fe); out["wrench_norm"] = float(np.linalg.norm(fe))
                assert out["wrench_norm"] > 0 and out["check_wrench_rel_l2"] <= 1e-3, out
            out["check_rel_l2_x"] = rel_l2(got[:, :3], r[:, :3]); out["check_rel_l2_v"] = rel_l2(got[:, 3:6], r[:, 3:6])
            out["check_rel_l2_F"] = rel_l2(got[:, 6:15], r[:, 6:15])
            assert out["check_rel_l2_x"] <= 1e-6 and out["check_rel_l2_v"] <= 5e-5 and out["check_rel_l2_F"] <= 5e-5, out
    return out if rank == 0 else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--particles", dest="n", type=int, default=8_000_000)
    ap.add_argument("--grid", dest="n_grid", type=int, default=256)
    ap.add_argument("--substeps", type=int, default=32)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sort-every", type=int, default=16)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--migrate-every", type=int, default=0, help="re-establish particle ownership every E substeps (particle migration "
                    "between slab ranks over NCCL P2P); 0: ownership fixed at reset")
    ap.add_argument("--sphere", action="store_true", help="(with --check) a sphere primitive on the slab boundary: forecast contact, wrench summed over ranks")
    ap.add_argument("--nccl", action="store_true", help="halo exchange through torch.distributed P2P (the first transport) instead of peer memory")
    ap.add_argument("--drift", type=float, default=0.0, help="add this x-velocity (m/s) to every particle so that material streams through the slab boundaries")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from softmac_b200 import rollouts
    rank, ws, local = rollouts.init()
    torch.cuda.set_device(local)
    if args.check:
        args.n, args.n_grid, args.substeps = 200_000, 64, 8
    out = slab_record(args.n, args.n_grid, args.substeps, rank, ws, local, reps=args.reps, sort_every=args.sort_every, check=args.check,
                      migrate_every=args.migrate_every, sphere=args.sphere, drift=args.drift, peer=not args.nccl)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if ws > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

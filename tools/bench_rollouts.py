#!/usr/bin/env python
"""BASELINE config 4: N independent grip-like rollouts sharded over the GPUs of one box, several rollouts per GPU
batched in one handle, action-gradient all-reduce (mean) at the end of the episode.

  python tools/bench_rollouts.py [--rollouts 64] [--particles 10000] [--env-steps 40] [--substeps 5] [--reps 3]
  torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_rollouts.py ...

Scene (the structure of demo_grip, softmac/config/demo_grip_config.py + demo_grip.py:117-124, with analytic sphere SDFs
for the two fingers and the stand-in rigid integrator instead of Jade): a 10k-particle plasticine block on a 64^3 grid
squeezed by two prismatic fingers driven by force actions 0.3*[1,-1]*(1 + 0.1 xi_k); loss = |x - target|^2 at the last
frame.  Prints one JSON line (rank 0).
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rollouts", type=int, default=64)
    ap.add_argument("--particles", dest="n", type=int, default=10000)
    ap.add_argument("--env-steps", type=int, default=40)
    ap.add_argument("--substeps", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sort-every", type=int, default=None, help="re-bin + re-sort period in substeps (default: the simulator's max(substeps, 4))")
    ap.add_argument("--device-rigid", action="store_true", help="run the (affine) rigid bridge on the GPU: no host round trip per env step")
    args = ap.parse_args()
    import torch
    import scenes
    from harness import sim_cfg
    from softmac_b200 import rollouts
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    from softmac_b200.engine.batched_env import BatchedTaichiEnv
    from softmac_b200.engine.rigid_simulator import RigidSimulator
    from softmac_b200.config import CfgNode
    rank, ws, local = rollouts.init()
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    mine = rollouts.shard(args.rollouts, rank, ws)
    B, n, S, K = len(mine), args.n, args.substeps, args.env_steps
    dt, n_grid = 2e-4, 64
    max_steps = K * S + S + 2
    rng = np.random.default_rng(0)
    x = ((rng.random((n, 3)) * 2 - 1) * 0.08 + np.array([0.5, 0.12, 0.5])).astype(np.float32).astype(np.float64)
    tab = scenes.sphere_table(radius=0.06, dx=0.005, margin=0.03)
    ms = [Mesh(sdf=dict(sdf=tab["sdf"], normal=tab["normal"], position=(tab["lower"], tab["upper"]), dx=tab["dx"]), cfg=dict(friction=0.3), max_timesteps=max_steps)
          for _ in range(2)]
    prims = Primitives(primitives=ms, max_timesteps=max_steps)
    sim = MPMSimulator(sim_cfg(n, n_grid=n_grid, max_steps=max_steps, dt=dt), prims, env_dt=dt * S, n_batch=B, device=local, sort_every=args.sort_every)
    bodies = [dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 - 0.138, 0.12, 0.5), mass=1.0, gravity=False),
              dict(joint="prismatic", axis=(1, 0, 0), origin=(0.5 + 0.138, 0.12, 0.5), mass=1.0, gravity=False)]
    rcfg = CfgNode(gravity=(0., 0., 0.), init_state=(0., 0., 0.3, -0.3), bodies=bodies)
    env = BatchedTaichiEnv(sim, prims, lambda b, views: RigidSimulator(rcfg, views, substeps=S, env_dt=dt * S), x, device_rigid=args.device_rigid)
    target = x * np.array([0.8, 1.1, 1.0]) + np.array([0.1, 0.0, 0.0])
    acts = np.stack([np.tile(0.3 * np.array([1.0, -1.0]) * (1 + 0.1 * np.random.default_rng(k).normal()) * 100, (K, 1)) for k in mine])   # (B, K, 2)

    def episode():
        env.reset()
        sim.clear_all_gradients()
        for k in range(K):
            env.step(acts[:, k])
        xs = sim.get_x(K * S).reshape(B, n, 3)
        d = xs - target
        sim.add_x_grad(K * S, d.reshape(B * n, 3))
        g = env.backward()                                  # (B, K, 2)
        return 0.5 * (d * d).sum(axis=(1, 2)), g

    times = []
    for r in range(args.reps + 1):
        if ws > 1:
            torch.distributed.barrier()
        sim.synchronize()
        t0 = time.perf_counter()
        loss, g = episode()
        gmean = rollouts.allreduce_gradients(g, args.rollouts)
        sim.synchronize()
        if ws > 1:
            torch.distributed.barrier()
        if r > 0:
            times.append(time.perf_counter() - t0)
    t = torch.tensor([float(np.median(times))], dtype=torch.float64, device="cuda" if torch.cuda.is_available() else "cpu")
    if ws > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    if rank == 0:
        T = float(t.item())
        print(json.dumps({"workload": "grip-like rollouts (config 4)", "rollouts": args.rollouts, "n_gpus": ws, "rollouts_per_gpu": B, "device_rigid": bool(args.device_rigid), "n_particles": n,
                          "env_steps": K, "substeps": S, "episode_s": T, "rollouts_per_s": args.rollouts / T,
                          "particle_substeps_per_s_fwd_bwd": args.rollouts * n * K * S / T, "grad_norm": float(np.linalg.norm(gmean)),
                          "loss_mean_local": float(np.mean(loss)), "counters": sim.counters()}), flush=True)
    if ws > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

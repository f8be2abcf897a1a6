#!/usr/bin/env python
"""bench.py -- forward+backward particle-substeps/s of the MLS-MPM substep loop (BASELINE.json metric).

Workload (SURVEY.md 8d, config 3 "cube-1M"): 1,000,000 particles from the reference generator
(Shapes.add_box, np.random.seed(0)) in a 0.390625^3 box at (0.5, 0.30, 0.5), 128^3 grid, co-rotated plastic
material, E = 3e3, nu = 0.2, gravity -9.8, sticky floor, mixed (forecast) contact model, dt = 1e-4 (SURVEY.md quotes
2e-4 "as grip", but at dx = 1/128 the reference scheme itself is unstable there: explicit-MPM limit dt < dx/sqrt(E)
= 1.4e-4; the f64 oracle blows up after ~20 substeps at 2e-4 -- see DESIGN.md).
Variant A: no primitive.  Variant B (--variant B): one static rigid sphere SDF under the cube.
One "step" = S substeps forward (smx_substep) followed by S substeps backward (smx_substep_grad) with a
dense seed on x at the last frame; value = N_gpus * n * S / step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--substeps S] [--n N] [--variant A|B]

N > 1 is launched by torchrun (one rank per GPU, NCCL): every rank runs an independent rollout (weak scaling)
and the per-rollout gradient summary is all-reduced at the end of every step, as in BASELINE config 4.
--impl reference times the CPU oracle (oracle/mpm_oracle.c: the f64 OpenMP restatement of the Taichi kernels;
Taichi itself cannot be installed here) on a bounded sample of the same workload, on rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle-substeps/sec fwd+bwd (1M p, 128^3)"
UNIT = "particle-substeps/s"
DT = 1e-4                   # see module docstring
ALGO_BYTES_STEP = 480.0     # 192 B forward + 288 B backward per particle-substep (fp32 storage)
# algorithmic HBM bytes per particle of each particle kernel (DESIGN.md "Kernels"): frame components read + written
# DRAM traffic per launch of the particle kernels from the ncu --set full capture of this round
# (profiles/r1c_ncu_full_particle_kernels.csv: dram__bytes_read.sum + dram__bytes_write.sum; k_g2p from r1_v3_...)
NCU_TRAFFIC = {"k_p2g_grad": 226.6e6 + 64.4e6, "k_p2g": 79.7e6 + 119.1e6, "k_g2p": 14.6e6 + 5.9e6, "k_g2p_grad": 85.1e6 + 6.4e6}
KERNEL_BYTES = {"k_p2g": 96 + 36, "k_g2p": 12 + 60, "k_g2p_grad": 12 + 60 + 12, "k_p2g_grad": 96 + 36 + 12 + 96,
                "k_p2g(recompute)": 96}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--substeps", type=int, default=64, help="substeps forward (and backward) per step")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--n-grid", type=int, default=128)
    ap.add_argument("--variant", default="A", choices=["A", "B"])
    ap.add_argument("--init", default="rest", choices=["rest", "stressed"], help="rest: the reference generator (v=0, F=I, C=0); stressed: same x with F = I + 0.01 N(0,1), v = 0.3 N(0,1), C = N(0,1) (every particle needs full SVD sweeps and clips plastically)")
    ap.add_argument("--batch", type=int, default=1, help="independent rollouts batched in one handle (per GPU)")
    ap.add_argument("--sort-every", type=int, default=32)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e arm on one handle only (no overlap of host I/O with the kernels of the other handle)")
    ap.add_argument("--cpu-sample-substeps", type=int, default=2)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def workload_cfg(args, max_steps):
    from harness import sim_cfg
    return sim_cfg(args.n, n_grid=args.n_grid, max_steps=max_steps, dt=DT, E=3e3, nu=0.2, gravity=(0., -9.8, 0.),
                   ground_friction=20., material_model=0, ptype=0, collision_type=2)


# variant B: a static sphere (radius 0.10) whose top sits 1.7 mm under the bottom face of the cube (y = 0.1047): inside the 5 mm band in
# which the forecast contact model is active, but not penetrating (a penetrating start ejects particles at sdf/dt = 350 m/s in one substep)
VARIANT_B_POSE = [0.5, 0.003, 0.5, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.]


def variant_b_table():
    import scenes
    return scenes.sphere_table(radius=0.10, dx=0.005, margin=0.03)


def make_inputs(args, rank):
    import scenes
    st = scenes.cube_state(args.n, seed=rank)           # rank r: np.random.seed(r) (rank 0 = the reference generator's seed)
    if getattr(args, "init", "rest") == "stressed":
        rng = np.random.default_rng(1000 + rank)
        st[:, 3:6] = 0.3 * rng.normal(size=(args.n, 3))
        st[:, 6:15] += 0.01 * rng.normal(size=(args.n, 9))
        st[:, 15:24] = rng.normal(size=(args.n, 9))
        st = st.astype(np.float32).astype(np.float64)
    seed = st[:, :3] - st[:, :3].mean(0)                # x_bar[S] = x[S] - mean (any fixed dense seed)
    return st, np.ascontiguousarray(seed)


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle (kind "port") on all host threads, bounded sample per step."""
    if rank != 0:
        return
    import scenes
    from oracle import mpm_oracle as mo
    mo.set_num_threads(os.cpu_count())          # all host threads (torchrun exports OMP_NUM_THREADS=1)
    S = args.cpu_sample_substeps
    sim = mo.OracleSim(args.n, n_grid=args.n_grid, max_steps=S + 1, dt=DT, E=3e3, nu=0.2, gravity=(0., -9.8, 0.),
                       ground_friction=20., material_model=0, ptype=0, collision_type=2, substeps=5)
    if args.variant == "B":
        t = variant_b_table()
        sim.add_primitive(t["sdf"], t["normal"], t["lower"], t["upper"], t["dx"], friction=0.5, softness=666.)
        for f in range(S + 1):
            sim.set_primitive_state(0, f, np.array(VARIANT_B_POSE))
    st, seed = make_inputs(args, 0)
    g24 = np.zeros((args.n, 24)); g24[:, :3] = seed

    def step():
        sim.set_frame(0, st)
        for f in range(S):
            sim.substep(f)
        sim.clear_grads(); sim.add_frame_grad(S, g24)
        for f in range(S - 1, -1, -1):
            sim.substep_grad(f)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = args.n * S / dt
    cores = mo.num_threads()
    sample = f"{S} substeps forward + {S} backward of the {args.n}-particle workload per step (f64, OpenMP, {cores} threads)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, S),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "restated ti.cpu-equivalent: Taichi 1.4.1 is not installable in this image"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def config_dict(args, S):
    return {"workload": f"cube-{args.n} variant {args.variant}: {args.n} particles, {args.n_grid}^3 grid, corotated plastic, mixed contact, "
                        f"{S} substeps forward + {S} backward per step",
            "n_particles": args.n, "rollouts_per_gpu": args.batch, "n_grid": args.n_grid, "substeps_per_step": S, "variant": args.variant, "init": args.init, "dt": DT,
            "sort_every": args.sort_every, "parallelism": f"{args.gpus} independent rollout(s), one per GPU, gradient all-reduce" if args.gpus > 1 else "single GPU",
            "cache": "inputs larger than L2: 96 MB per particle frame, one fresh frame per substep"}


def cpu_baseline(args):
    import scenes  # noqa: F401
    from oracle import mpm_oracle as mo
    mo.set_num_threads(os.cpu_count())
    S = args.cpu_sample_substeps
    sim = mo.OracleSim(args.n, n_grid=args.n_grid, max_steps=S + 1, dt=DT, E=3e3, nu=0.2, gravity=(0., -9.8, 0.),
                       ground_friction=20., material_model=0, ptype=0, collision_type=2, substeps=5)
    st, seed = make_inputs(args, 0)
    g24 = np.zeros((args.n, 24)); g24[:, :3] = seed
    reps, times = 3, []
    for r in range(reps + 1):
        sim.set_frame(0, st)
        t0 = time.perf_counter()
        for f in range(S):
            sim.substep(f)
        sim.clear_grads(); sim.add_frame_grad(S, g24)
        for f in range(S - 1, -1, -1):
            sim.substep_grad(f)
        if r > 0:
            times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    cores = mo.num_threads()
    return {"value": args.n * S / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{S} substeps forward + {S} backward at {args.n} particles, median of {reps} (variant A, f64 OpenMP restatement of the Taichi kernels)"}


def P2(args, S):
    """A second primitive container for the second handle of the pipelined end-to-end arm (variant B)."""
    from softmac_b200.engine import Primitives, Mesh
    t = variant_b_table()
    m = Mesh(sdf=dict(sdf=t["sdf"], normal=t["normal"], position=(t["lower"], t["upper"]), dx=t["dx"]), cfg=dict(friction=0.5), max_timesteps=S + 2)
    m.softness[None] = 666.
    m.set_all_states(0, np.array(VARIANT_B_POSE), f_end=S + 2)
    return Primitives(primitives=[m], max_timesteps=S + 2)


# ------------------------------------------------------------------------------------------------------------
def run_cuda(args, rank, world, local_rank):
    import torch
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # NCCL prints its version banner (and any NCCL_DEBUG output) on stdout: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S = args.substeps
    cfg = workload_cfg(args, S + 2)
    prims = []
    if args.variant == "B":
        t = variant_b_table()
        m = Mesh(sdf=dict(sdf=t["sdf"], normal=t["normal"], position=(t["lower"], t["upper"]), dx=t["dx"]), cfg=dict(friction=0.5), max_timesteps=S + 2)
        m.softness[None] = 666.
        prims.append(m)
    P = Primitives(primitives=prims, max_timesteps=S + 2)
    sim = MPMSimulator(cfg, P, env_dt=5 * DT, device=local_rank, sort_every=args.sort_every, flags=args.flags, n_batch=args.batch)
    for p in prims:
        p.set_all_states(0, np.array(VARIANT_B_POSE), f_end=S + 2)
    st, seed = make_inputs(args, rank)
    if args.batch > 1:
        st, seed = np.tile(st, (args.batch, 1)), np.tile(seed, (args.batch, 1))
    gsum = torch.zeros(16, device="cuda")

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        sim.synchronize()

    # ---- device-resident arm: inputs already in HBM when the timed region starts -------------------------------
    sim.reset(st)
    sim.copyframe(0, S + 1)             # pristine copy of the initial state, stays in HBM
    sim.add_x_grad(S, seed)             # seed buffer resident on the device

    def device_step():
        sim.copyframe(S + 1, 0)
        sim.step(0, S)
        sim.step_grad(S, S)
        if dist:
            dist.all_reduce(gsum)       # gradient all-reduce across rollouts (BASELINE config 4)

    for _ in range(args.warmup):
        device_step()
    barrier()
    l0 = sim.launch_count()
    clocks = ClockSampler(local_rank); clocks.start()
    sim.timer_start()
    for _ in range(args.steps):
        device_step()
    ms = sim.timer_stop()
    barrier()
    clk = clocks.stop()
    launches = sim.launch_count() - l0
    t_dev = torch.tensor([ms / args.steps], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_step = float(t_dev.item())
    value = world * args.batch * args.n * S / (ms_step * 1e-3)
    counters = sim.counters()

    # ---- per-kernel durations (CUDA events on the simulator's stream), one extra step --------------------------
    roof = kernel_roofline(sim, args, S)

    # ---- end-to-end arm: host buffers in, host result out, through the public API ------------------------------
    e2e = None
    if not args.no_e2e:
        reps = max(2, min(args.steps, 3))
        times = []
        for r in range(reps + 1):
            barrier()
            t0 = time.perf_counter()
            sim.reset(st)                       # H2D: n*24 fp32 from pinned staging
            sim.clear_all_gradients()
            sim.add_x_grad(S, seed)             # H2D: seed
            sim.step(0, S)
            sim.step_grad(S, S)
            xg, vg = sim.get_grad(0)            # D2H: the reference's own read-out, MPMSimulator.get_grad(f) -> (x_bar, v_bar)
            if dist:
                dist.all_reduce(gsum)
            barrier()
            if r > 0:
                times.append(time.perf_counter() - t0)
        t = torch.tensor([float(np.median(times))], device="cuda", dtype=torch.float64)
        if dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e = {"value": world * args.batch * args.n * S / float(t.item()), "unit": UNIT, "h2d_bytes_per_step": args.batch * (args.n * 24 * 4 + args.n * 3 * 4),
               "d2h_bytes_per_step": args.batch * args.n * 6 * 4, "checksum": float(np.abs(xg).sum() + np.abs(vg).sum()),
               "api": "reset(state (n,24) f64) + add_x_grad + step + step_grad + get_grad(0) -> (x_bar, v_bar) f64"}
        if not args.no_e2e_pipeline:
            # The same calls, every step with its own host -> device upload and device -> host read-back, on TWO handles (two streams):
            # while the kernels of step k run on one handle, the host converts / uploads the inputs of step k+1 into the other and reads
            # back the result of step k-1 -- independent rollouts, the way config 4 runs them.
            sim2 = MPMSimulator(cfg, Primitives(primitives=[], max_timesteps=S + 2) if not prims else P2(args, S), env_dt=5 * DT, device=local_rank,
                                sort_every=args.sort_every, flags=args.flags, n_batch=args.batch)
            pair = [sim, sim2]

            def start(h):                           # upload + forward substeps (asynchronous)
                h.reset(st)
                h.clear_all_gradients()
                h.add_x_grad(S, seed)
                h.step(0, S)

            def finish(h):                          # backward substeps (waits for the forward: one look at the checkpoint counter)
                h.step_grad(S, S)

            def sync_all():
                barrier(); sim2.synchronize()
            K = max(10, 2 * reps)
            chk = 0.0
            for r in range(2):                      # one warm-up pass, one timed pass of K steps
                sync_all()
                t0 = time.perf_counter()
                start(pair[0]); finish(pair[0])
                for k in range(1, K):
                    start(pair[k % 2])                              # host conversion + H2D of step k overlap the backward of step k-1
                    xg, vg = pair[(k - 1) % 2].get_grad(0)          # D2H + conversion of step k-1 overlap the forward of step k
                    finish(pair[k % 2])
                xg, vg = pair[(K - 1) % 2].get_grad(0)
                if dist:
                    dist.all_reduce(gsum)
                sync_all()
                tp = (time.perf_counter() - t0) / K
                chk = float(np.abs(xg).sum() + np.abs(vg).sum())
            t = torch.tensor([tp], device="cuda", dtype=torch.float64)
            if dist:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e["sequential_value"] = e2e["value"]
            e2e["value"] = world * args.batch * args.n * S / float(t.item())
            e2e["mode"] = (f"two handles, software-pipelined over {K} steps: upload of step k+1 and read-back of step k-1 overlap the kernels of step k; "
                           "every step still uploads its own inputs and reads back its own result; sequential_value = one handle, one step at a time")
            e2e["pipelined_checksum"] = chk
            del sim2

    if rank == 0:
        peak, peak_src = peaks()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, S), "clocks": clk, "gpu_launches": int(launches),
                "step_hbm_frac": ALGO_BYTES_STEP * value / world / (peak * 1e9), "counters": counters}
        if roof:
            roof.update({"peak": peak, "peak_source": peak_src, "frac": roof["achieved"] / peak})
            line["roofline"] = roof
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def kernel_roofline(sim, args, S):
    """Times each kernel class of one forward and one backward substep with CUDA events on the launching stream
    (smx_timer_*), by running the substep loop of the step with profiling splits.  Returns the dominant kernel."""
    try:
        from softmac_b200 import _capi
        L = _capi.lib()
        if not hasattr(L, "smx_profile_substep"):
            return None
    except Exception:
        return None
    import ctypes as C
    names = (C.c_char_p * 32)()
    ms = (C.c_float * 32)()
    cnt = C.c_int(0)
    tot = {}
    L.smx_profile_substep.argtypes = [_capi.vp, C.c_int32, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.smx_profile_substep.restype = C.c_int
    nprof = min(8, S)

    def prof(f, mode):
        _capi.check(L.smx_profile_substep(sim._h, f, mode, names, ms, C.byref(cnt)))
        for i in range(cnt.value):
            key = names[i].decode() + ("(recompute)" if mode == 1 and names[i].decode() in ("k_p2g", "k_grid_op", "k_contact") else "")
            tot.setdefault(key, []).append(ms[i])

    sim.copyframe(S + 1, 0)
    for f in range(S):
        if S // 2 <= f < S // 2 + nprof or S <= nprof:
            prof(f, 0)
        else:
            sim.substep(f)
    for f in range(S - 1, -1, -1):
        if S // 2 <= f < S // 2 + nprof or S <= nprof:
            prof(f, 1)
        else:
            sim.substep_grad(f)
    avg = {k: float(np.mean(v)) for k, v in tot.items()}
    total = sum(avg.values())
    top = max((k for k in avg if k in KERNEL_BYTES), key=lambda k: avg[k])
    achieved = KERNEL_BYTES[top] * args.n * args.batch / (avg[top] * 1e-3) / 1e9
    traffic = NCU_TRAFFIC.get(top) if (args.n == 1_000_000 and args.batch == 1) else None
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "unit": "GB/s", "traffic": traffic,
            "traffic_source": "ncu --set full capture of this kernel at this size (profiles/r1c_ncu_full_particle_kernels.csv), bytes per launch" if traffic else None,
            "algorithmic_bytes_per_launch": KERNEL_BYTES[top] * args.n * args.batch, "avg_launch_ms": avg[top],
            "share_of_substep_pair": avg[top] / total if total > 0 else None,
            "kernel_ms": avg, "how": "CUDA events around each launch on the simulator stream, 8 forward + 8 backward substeps after the timed region"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

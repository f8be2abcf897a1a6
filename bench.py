#!/usr/bin/env python
"""bench.py -- forward+backward particle-substeps/s of the MLS-MPM substep loop (BASELINE.json metric).

Workload (SURVEY.md 8d, config 3 "cube-1M"): 1,000,000 particles from the reference generator
(Shapes.add_box, np.random.seed(0)) in a 0.390625^3 box at (0.5, 0.30, 0.5), 128^3 grid, co-rotated plastic
material, E = 3e3, nu = 0.2, gravity -9.8, sticky floor, mixed (forecast) contact model, dt = 1e-4 (SURVEY.md quotes
2e-4 "as grip", but at dx = 1/128 the reference scheme itself is unstable there: explicit-MPM limit dt < dx/sqrt(E)
= 1.4e-4; the f64 oracle blows up after ~20 substeps at 2e-4 -- see DESIGN.md).
One "step" = S substeps forward (smx_step) followed by S substeps backward (smx_step_grad) with a dense seed on x at the
last frame; value = N_gpus * n * S / step time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl cuda|reference] [--substeps S] [--n N] [--variant A|B]

The one JSON line carries, measured in the same run:
  value / ms_per_step  headline arm: variant A (no primitive), rest state, inputs resident in HBM, CUDA events on the simulator stream
  parity               CUDA vs the f64 oracle on the SAME 1M-particle inputs (2 substeps forward + the frame-0 adjoint): rel-L2 of
                       x / v / F / C per substep and the adjoint cosine; the process exits with rc 3 when a north_star tolerance
                       (1e-4 / 0.999) is missed
  stressed, variant_B  the harder arms (every particle needs full Jacobi sweeps and clips plastically; a sphere under the cube:
                       forecast contact + wrench reduction every substep), each with its own value, step_hbm_frac and parity
  roofline             dominant kernel of the fused hot path, timed live with CUDA events; traffic from the checked-in ncu capture
  e2e                  the same step through the public API with HOST buffers (float32 pinned once; the f64 arm of the reference's own types
                       alongside), H2D / D2H inside the timed region
  cpu_baseline         the oracle (f64 OpenMP port of the Taichi kernels) on all host threads, bounded sample
N > 1 (torchrun, one rank per GPU, NCCL): every rank runs an independent rollout of the headline workload (weak scaling) and the
per-rollout gradient summary (sums and norms of x.grad / v.grad of frame 0, computed on the device) is all-reduced at the end of
every step; then two sub-records that exercise BASELINE configs 4 and 5 on the same ranks:
  rollouts64           64 demo_grip rollouts, 64/N per rank, action gradients all-reduced (strong scaling)
  slab_8M              8M particles / 256^3 split into x-slabs over the N ranks, halo exchange over NCCL (strong scaling), plus a
                       200k-particle correctness check against a single handle
--impl reference times the CPU oracle (Taichi itself cannot be installed here) on a bounded sample of the same workload, rank 0.
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
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle-substeps/sec fwd+bwd (1M p, 128^3)"
UNIT = "particle-substeps/s"
DT = 1e-4                   # see module docstring
ALGO_BYTES_STEP = 480.0     # 192 B forward + 288 B backward per particle-substep (fp32 storage)
# algorithmic HBM bytes per particle of each kernel class of the fused hot path (DESIGN.md "Kernels"): the frame components the
# kernel must read + write in its fused role (SVD records, grid checkpoints and prefetches are NOT algorithmic)
KERNEL_BYTES = {
    "k_g2p2g": 12 + 60 + 36 + 36,                   # G2P of f-1 (x in, x v C out) + P2G of f (F in, F out; x v C from registers)
    "k_p2g": 96 + 36,                               # first substep of a call / after a re-sort
    "k_g2p": 12 + 60,                               # last substep of a call / before a re-sort
    "k_g2p_grad": 12 + 60 + 12,                     # first adjoint substep of a call
    "k_p2g_grad": 96 + 36 + 12 + 96,                # last adjoint substep of a call
    "k_p2g_grad+g2p_grad": 96 + 36 + 12 + 36 + 12 + 12,   # frame, F adjoint in, x partial in, F adjoint out, x of f-1 in, x partial out
}
NCU_JSON = os.path.join(ROOT, "profiles", "r2_ncu_kernels.json")     # per-kernel summary of the checked-in ncu --set full capture


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--substeps", type=int, default=64, help="substeps forward (and backward) per step")
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--n-grid", type=int, default=128)
    ap.add_argument("--variant", default="A", choices=["A", "B"], help="headline arm (the other arms are sub-records)")
    ap.add_argument("--init", default="rest", choices=["rest", "stressed"], help="rest: the reference generator (v=0, F=I, C=0); stressed: same x with F = I + 0.01 N(0,1), v = 0.3 N(0,1), C = N(0,1) (every particle needs full SVD sweeps and clips plastically)")
    ap.add_argument("--batch", type=int, default=1, help="independent rollouts batched in one handle (per GPU)")
    ap.add_argument("--sort-every", type=int, default=32)
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-pipeline", action="store_true", help="e2e arm on one handle only (no overlap of host I/O with the kernels of the other handle)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison on the bench inputs")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the stressed / variant_B (and, N > 1, rollouts64 / slab_8M) sub-records")
    ap.add_argument("--cpu-sample-substeps", type=int, default=2)
    ap.add_argument("--slab-particles", type=int, default=8_000_000)
    ap.add_argument("--slab-grid", type=int, default=256)
    ap.add_argument("--slab-substeps", type=int, default=16)
    ap.add_argument("--rollouts", type=int, default=64)
    ap.add_argument("--rollout-env-steps", type=int, default=400)
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


def make_inputs(args, rank, init=None):
    import scenes
    st = scenes.cube_state(args.n, seed=rank)           # rank r: np.random.seed(r) (rank 0 = the reference generator's seed)
    if (init or getattr(args, "init", "rest")) == "stressed":
        rng = np.random.default_rng(1000 + rank)
        st[:, 3:6] = 0.3 * rng.normal(size=(args.n, 3))
        st[:, 6:15] += 0.01 * rng.normal(size=(args.n, 9))
        st[:, 15:24] = rng.normal(size=(args.n, 9))
        st = st.astype(np.float32).astype(np.float64)
    seed = st[:, :3] - st[:, :3].mean(0)                # x_bar[S] = x[S] - mean (any fixed dense seed)
    return st, np.ascontiguousarray(seed)


def config_dict(args, S, variant=None, init=None):
    variant, init = variant or args.variant, init or args.init
    return {"workload": f"cube-{args.n} variant {variant}: {args.n} particles, {args.n_grid}^3 grid, corotated plastic, mixed contact, "
                        f"{S} substeps forward + {S} backward per step",
            "n_particles": args.n, "rollouts_per_gpu": args.batch, "n_grid": args.n_grid, "substeps_per_step": S, "variant": variant, "init": init, "dt": DT,
            "sort_every": args.sort_every, "parallelism": f"{args.gpus} independent rollout(s), one per GPU, gradient all-reduce" if args.gpus > 1 else "single GPU",
            "cache": "inputs larger than L2: 96 MB per particle frame, one fresh frame per substep"}


# ------------------------------------------------------------------------------------------------------------
# the oracle (test infrastructure): CPU baseline and parity checker.  Never on the product path.
# ------------------------------------------------------------------------------------------------------------
def oracle_sim(args, S, variant):
    from oracle import mpm_oracle as mo
    mo.set_num_threads(os.cpu_count())          # all host threads (torchrun exports OMP_NUM_THREADS=1)
    sim = mo.OracleSim(args.n, n_grid=args.n_grid, max_steps=S + 1, dt=DT, E=3e3, nu=0.2, gravity=(0., -9.8, 0.),
                       ground_friction=20., material_model=0, ptype=0, collision_type=2, substeps=5)
    if variant == "B":
        t = variant_b_table()
        t32 = {k: np.asarray(t[k], dtype=np.float32).astype(np.float64) for k in ("sdf", "normal", "lower", "upper")}
        sim.add_primitive(t32["sdf"], t32["normal"], t32["lower"], t32["upper"], t["dx"], friction=0.5, softness=666.)
        for f in range(S + 1):
            sim.set_primitive_state(0, f, np.array(VARIANT_B_POSE))
    return sim, mo


def oracle_step(sim, S, st, g24):
    sim.set_frame(0, st)
    for f in range(S):
        sim.substep(f)
    sim.clear_grads(); sim.add_frame_grad(S, g24)
    for f in range(S - 1, -1, -1):
        sim.substep_grad(f)


def run_reference(args, rank, world):
    """CPU arm: the oracle (kind "port") on all host threads, bounded sample per step."""
    if rank != 0:
        return
    S = args.cpu_sample_substeps
    sim, mo = oracle_sim(args, S, args.variant)
    st, seed = make_inputs(args, 0)
    g24 = np.zeros((args.n, 24)); g24[:, :3] = seed
    for _ in range(args.warmup):
        oracle_step(sim, S, st, g24)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(sim, S, st, g24)
    dt = (time.perf_counter() - t0) / args.steps
    value = args.n * S / dt
    cores = mo.num_threads()
    sample = f"{S} substeps forward + {S} backward of the {args.n}-particle workload per step (f64, OpenMP, {cores} threads)"
    cfg = config_dict(args, S)
    cfg["sample_of"] = f"{args.substeps}+{args.substeps} substeps per step of the cuda arm (throughput-normalised bounded sample of the same workload)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "restated ti.cpu-equivalent: Taichi 1.4.1 is not installable in this image"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parity_and_cpu(sim, args, variant, init, prims, want_cpu_time):
    """CUDA vs oracle on the bench inputs themselves: S_c substeps forward through smx_step (the fused launches of the timed path),
    the adjoint of frame 0 through smx_step_grad; per-substep rel-L2 of x / v / F / C and the adjoint cosine.  The same oracle run
    is the cpu_baseline sample (timed, median of 3) when want_cpu_time."""
    from harness import rel_l2, cosine
    S = args.cpu_sample_substeps
    orc, mo = oracle_sim(args, S, variant)
    st, seed = make_inputs(args, 0, init)
    g24 = np.zeros((args.n, 24)); g24[:, :3] = seed
    times = []
    for r in range(4 if want_cpu_time else 1):
        t0 = time.perf_counter()
        oracle_step(orc, S, st, g24)
        if r > 0:
            times.append(time.perf_counter() - t0)
    sim.reset(st)
    for p in prims:
        p.set_all_states(0, np.array(VARIANT_B_POSE), f_end=S + 1)
    sim.clear_all_gradients()
    sim.add_x_grad(S, seed)
    sim.step(0, S)
    sim.step_grad(S, S)
    cols = dict(x=slice(0, 3), v=slice(3, 6), F=slice(6, 15), C=slice(15, 24))
    worst = {k: 0.0 for k in cols}
    dx = 1.0 / args.n_grid
    for f in range(1, S + 1):
        ref, got = orc.get_frame(f), sim.get_state(f)
        # quantities whose oracle norm is ~0 are measured against their natural scale (C of a rigid translation: |v| / dx; v of a
        # body at rest: g dt), as in tests/test_cuda_parity.py
        floors = dict(x=0.0, v=float(np.sqrt(args.n)) * 9.8 * DT, F=0.0, C=float(np.linalg.norm(ref[:, cols["v"]])) / dx)
        for k, sl in cols.items():
            worst[k] = max(worst[k], rel_l2(got[:, sl], ref[:, sl], floor=floors[k]))
    go, gg = orc.get_frame_grad(0), sim.get_state_grad(0)
    rec = {"substeps": S, "n_particles": args.n, "x": worst["x"], "v": worst["v"], "F": worst["F"], "C": worst["C"],
           "adjoint_cosine": cosine(gg, go), "adjoint_rel_l2": rel_l2(gg, go), "adjoint_norm_oracle": float(np.linalg.norm(go)),
           "tolerance": "north_star: per-substep rel-L2 <= 1e-4 on x/v/F/C, gradient cosine >= 0.999",
           "oracle": "oracle/mpm_oracle.c (f64 restatement; parity to Taichi itself is unpinned, see DESIGN.md section 3)"}
    rec["pass"] = bool(max(worst.values()) <= 1e-4 and rec["adjoint_cosine"] >= 0.999 and rec["adjoint_norm_oracle"] > 0)
    cpu = None
    if want_cpu_time:
        dt = float(np.median(times))
        cpu = {"value": args.n * S / dt, "unit": UNIT, "cores": mo.num_threads(), "kind": "port",
               "sample": f"{S} substeps forward + {S} backward at {args.n} particles, median of 3 (variant {variant}, f64 OpenMP restatement of the Taichi kernels)"}
    del orc
    return rec, cpu


# ------------------------------------------------------------------------------------------------------------
def build_sim(args, S, variant, local_rank):
    from softmac_b200.engine import MPMSimulator, Primitives, Mesh
    cfg = workload_cfg(args, S + 2)
    prims = []
    if variant == "B":
        t = variant_b_table()
        m = Mesh(sdf=dict(sdf=t["sdf"], normal=t["normal"], position=(t["lower"], t["upper"]), dx=t["dx"]), cfg=dict(friction=0.5), max_timesteps=S + 2)
        m.softness[None] = 666.
        prims.append(m)
    P = Primitives(primitives=prims, max_timesteps=S + 2)
    sim = MPMSimulator(cfg, P, env_dt=5 * DT, device=local_rank, sort_every=args.sort_every, flags=args.flags, n_batch=args.batch)
    for p in prims:
        p.set_all_states(0, np.array(VARIANT_B_POSE), f_end=S + 2)
    return sim, prims, cfg


def device_arm(sim, args, S, st, seed, steps, warmup, local_rank, dist=None, sample_clocks=False):
    """Inputs resident in HBM when the timed region starts; CUDA events on the simulator's stream; max over ranks."""
    import torch
    gsum = torch.zeros(16, device="cuda")
    ext = torch.cuda.ExternalStream(sim.stream_ptr(), device=torch.device("cuda", local_rank)) if dist else None

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        sim.synchronize()

    sim.reset(st)
    sim.copyframe(0, S + 1)             # pristine copy of the initial state, stays in HBM
    sim.clear_all_gradients()
    sim.add_x_grad(S, seed)             # seed buffer resident on the device

    def step():
        sim.copyframe(S + 1, 0)
        sim.step(0, S)
        sim.step_grad(S, S)
        if dist:
            # gradient all-reduce across the independent rollouts (BASELINE config 4): the summary of this rollout's frame-0 adjoint is
            # reduced on the device by the library, then summed over the ranks by NCCL -- stream-ordered, no host synchronisation
            sim.grad_summary_dev(0, gsum.data_ptr())
            torch.cuda.current_stream().wait_stream(ext)
            dist.all_reduce(gsum)
            ext.wait_stream(torch.cuda.current_stream())

    for _ in range(warmup):
        step()
    barrier()
    l0 = sim.launch_count()
    clocks = ClockSampler(local_rank) if sample_clocks else None
    if clocks:
        clocks.start()
    sim.timer_start()
    for _ in range(steps):
        step()
    ms = sim.timer_stop()
    barrier()
    clk = clocks.stop() if clocks else None
    launches = sim.launch_count() - l0
    t_dev = torch.tensor([ms / steps], device="cuda", dtype=torch.float64)
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    return float(t_dev.item()), int(launches), clk, [float(v) for v in gsum.tolist()]


def sub_record(args, S, variant, init, local_rank, peak):
    """A harder arm measured in the same run (device-resident, fewer steps), with its own parity check."""
    sim, prims, _ = build_sim(args, S, variant, local_rank)
    st, seed = make_inputs(args, 0, init)
    if args.batch > 1:
        st, seed = np.tile(st, (args.batch, 1)), np.tile(seed, (args.batch, 1))
    steps, warmup = max(3, args.steps // 2), max(3, min(args.warmup, 3))
    ms_step, launches, _, _ = device_arm(sim, args, S, st, seed, steps, warmup, local_rank)
    value = args.batch * args.n * S / (ms_step * 1e-3)
    rec = {"value": value, "unit": UNIT, "ms_per_step": ms_step, "steps": steps, "warmup": warmup, "step_hbm_frac": ALGO_BYTES_STEP * value / (peak * 1e9),
           "gpu_launches": launches, "config": config_dict(args, S, variant, init), "counters": sim.counters()}
    if not args.no_parity and args.batch == 1:
        rec["parity"], _ = parity_and_cpu(sim, args, variant, init, prims, False)
    del sim
    return rec


def run_cuda(args, rank, world, local_rank):
    # stdout carries exactly ONE line (the JSON record): whatever libraries print there (NCCL's version banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    from softmac_b200.engine import MPMSimulator, Primitives
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        # NCCL prints its version banner (and any NCCL_DEBUG output) on stdout: keep stdout for the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_debug.%h.%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    S = args.substeps
    peak, peak_src = peaks()
    sim, prims, cfg = build_sim(args, S, args.variant, local_rank)
    st, seed = make_inputs(args, rank)
    if args.batch > 1:
        st, seed = np.tile(st, (args.batch, 1)), np.tile(seed, (args.batch, 1))

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        sim.synchronize()

    # ---- device-resident arm: inputs already in HBM when the timed region starts -------------------------------
    ms_step, launches, clk, gsum = device_arm(sim, args, S, st, seed, args.steps, args.warmup, local_rank, dist, sample_clocks=True)
    value = world * args.batch * args.n * S / (ms_step * 1e-3)
    counters = sim.counters()

    # ---- per-kernel durations of the fused hot path (CUDA events on the simulator's stream), one extra step ----
    roof = kernel_roofline(sim, args, S)

    # ---- end-to-end arm: host buffers in, host result out, through the public API ------------------------------
    e2e = None
    if not args.no_e2e:
        reps = max(2, min(args.steps, 3))
        sim2 = prims2 = None
        if not args.no_e2e_pipeline:
            sim2, prims2, _ = build_sim(args, S, args.variant, local_rank)

        def reduce_max(x):
            t = torch.tensor([float(x)], device="cuda", dtype=torch.float64)
            if dist:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        def e2e_arm(st_h, seed_h, outs):
            """One step = reset(state) + add_x_grad(seed) + step + step_grad + get_grad(0) with HOST buffers: sequential on one handle,
            then software-pipelined on two handles (two streams).  outs: None (f64: fresh arrays) or per-handle fp32 (xg, vg)."""
            def read(h, i):
                return h.get_grad(0) if outs is None else h.get_grad(0, out=outs[i])
            times = []
            for r in range(reps + 1):
                barrier()
                t0 = time.perf_counter()
                sim.reset(st_h)                     # H2D: n*24 fp32 (f64 input: converted on the host threads into pinned staging first)
                sim.clear_all_gradients()
                sim.add_x_grad(S, seed_h)           # H2D: seed
                sim.step(0, S)
                sim.step_grad(S, S)
                xg, vg = read(sim, 0)               # D2H: the reference's own read-out, MPMSimulator.get_grad(f) -> (x_bar, v_bar)
                barrier()
                if r > 0:
                    times.append(time.perf_counter() - t0)
            out = {"sequential_s": reduce_max(np.median(times)), "checksum": float(np.abs(xg).sum(dtype=np.float64) + np.abs(vg).sum(dtype=np.float64))}
            if sim2 is None:
                return out
            # The same calls, every step with its own host -> device upload and device -> host read-back, on TWO handles (two streams):
            # while the kernels of step k run on one handle, the host uploads the inputs of step k+1 into the other and reads back the
            # result of step k-1 -- independent rollouts, the way config 4 runs them.
            pair = [sim, sim2]

            def start(h):                           # upload + forward substeps (asynchronous)
                h.reset(st_h)
                h.clear_all_gradients()
                h.add_x_grad(S, seed_h)
                h.step(0, S)

            def finish(h):                          # backward substeps (waits for the forward: one look at the checkpoint counter)
                h.step_grad(S, S)

            def sync_all():
                barrier(); sim2.synchronize()
            K = max(10, 2 * reps)
            for r in range(2):                      # one warm-up pass, one timed pass of K steps
                sync_all()
                t0 = time.perf_counter()
                start(pair[0]); finish(pair[0])
                for k in range(1, K):
                    start(pair[k % 2])                              # H2D of step k overlaps the backward of step k-1
                    xg, vg = read(pair[(k - 1) % 2], (k - 1) % 2)   # D2H of step k-1 overlaps the forward of step k
                    finish(pair[k % 2])
                xg, vg = read(pair[(K - 1) % 2], (K - 1) % 2)
                sync_all()
                tp = (time.perf_counter() - t0) / K
            out.update({"pipelined_s": reduce_max(tp), "pipelined_steps": K,
                        "pipelined_checksum": float(np.abs(xg).sum(dtype=np.float64) + np.abs(vg).sum(dtype=np.float64))})
            return out

        units = world * args.batch * args.n * S
        a64 = e2e_arm(st, seed, None)
        # float32 host buffers pinned once by the caller (MPMSimulator.pin): the rows travel as they are, no conversion pass on the host
        st32, seed32 = np.ascontiguousarray(st, dtype=np.float32), np.ascontiguousarray(seed, dtype=np.float32)
        outs = [(np.empty_like(seed32), np.empty_like(seed32)) for _ in range(2)]
        pinned = [st32, seed32] + [a for o in outs for a in o]
        sim.pin(*pinned)
        a32 = e2e_arm(st32, seed32, outs)
        sim.unpin(*pinned)
        best = a32.get("pipelined_s", a32["sequential_s"])
        e2e = {"value": units / best, "unit": UNIT, "h2d_bytes_per_step": args.batch * (args.n * 24 * 4 + args.n * 3 * 4),
               "d2h_bytes_per_step": args.batch * args.n * 6 * 4, "checksum": a32["checksum"],
               "api": "reset(state (n,24) float32, pinned once) + add_x_grad(seed float32) + step + step_grad + get_grad(0, out=(x_bar, v_bar) float32): "
                      "MPMSimulator's fp32 host entry points (smx_reset_f32 / smx_add_x_grad_f32 / smx_get_grad_f32), host buffers in, host result out",
               "sequential_value": units / a32["sequential_s"],
               "mode": ("two handles, software-pipelined over %d steps: upload of step k+1 and read-back of step k-1 overlap the kernels of step k; every step "
                        "still uploads its own inputs and reads back its own result; sequential_value = one handle, one step at a time" % a32["pipelined_steps"])
                       if "pipelined_s" in a32 else "one handle, one step at a time",
               "f64": {"api": "the reference's own types: reset(state (n,24) float64) + add_x_grad(float64) + ... + get_grad(0) -> float64 (converted on the host threads)",
                       "value": units / a64.get("pipelined_s", a64["sequential_s"]), "sequential_value": units / a64["sequential_s"], "checksum": a64["checksum"]}}
        if "pipelined_checksum" in a32:
            e2e["pipelined_checksum"] = a32["pipelined_checksum"]
        del sim2, prims2

    # ---- parity on the bench inputs + CPU baseline (rank 0; the oracle is the checker, never the thing measured) ----
    parity = cpu = None
    if rank == 0 and args.batch == 1 and not (args.no_parity and args.no_cpu_baseline):
        parity, cpu = parity_and_cpu(sim, args, args.variant, args.init, prims, not args.no_cpu_baseline)
        if args.no_parity:
            parity = None
    del sim

    # ---- harder arms of the same workload, same run (rank 0 only: single-GPU numbers) ----------------------------
    subs = {}
    if rank == 0 and not args.no_subrecords:
        if not (args.init == "stressed" and args.variant == "A"):
            subs["stressed"] = sub_record(args, S, "A", "stressed", local_rank, peak)
        if args.variant != "B":
            subs["variant_B"] = sub_record(args, S, "B", "rest", local_rank, peak)
    if dist:
        dist.barrier()

    # ---- BASELINE configs 4 and 5 on the same ranks (N > 1) ----------------------------------------------------
    if dist and not args.no_subrecords:
        import bench_demo
        import bench_slabs
        try:
            rec = bench_demo.rollouts_record(bench_demo.scene("grip"), "grip", args.rollout_env_steps, args.rollouts, rank, world, local_rank, reps=2, sort_every=25)
        except Exception as e:                      # noqa: BLE001  (a failing sub-record must not take the headline line with it)
            rec = {"error": repr(e)[:300]}
        if rank == 0:
            subs["rollouts64"] = rec
        dist.barrier()
        try:
            rec = bench_slabs.slab_record(args.slab_particles, args.slab_grid, args.slab_substeps, rank, world, local_rank, reps=2, sort_every=16)
            chk = bench_slabs.slab_record(200_000, 64, 8, rank, world, local_rank, reps=1, sort_every=16, check=True)
            if rank == 0:
                rec["check_200k"] = {k: v for k, v in chk.items() if k.startswith("check_") or k in ("n_gpus", "workload")}
        except Exception as e:                      # noqa: BLE001
            rec = {"error": repr(e)[:300]}
        if rank == 0:
            subs["slab_8M"] = rec
        dist.barrier()

    ok = True
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, S), "clocks": clk, "gpu_launches": int(launches),
                "step_hbm_frac": ALGO_BYTES_STEP * value / world / (peak * 1e9), "counters": counters}
        if dist:
            line["allreduce"] = {"what": "sum over ranks of [sum x.grad(3), sum v.grad(3), |x.grad|^2, |v.grad|^2, particles] of frame 0 of every rank's rollout, "
                                         "reduced on the device (smx_grad_summary_dev) and all-reduced with NCCL every step", "last": gsum[:9]}
        if roof:
            roof.update({"peak": peak, "peak_source": peak_src, "frac": roof["achieved"] / peak})
            line["roofline"] = roof
        if e2e:
            line["e2e"] = e2e
        if parity:
            line["parity"] = parity
        if cpu:
            line["cpu_baseline"] = cpu
        line.update(subs)
        fails = [k for k in ("parity",) if line.get(k) and not line[k]["pass"]] + \
                [k for k in ("stressed", "variant_B") if line.get(k, {}).get("parity") and not line[k]["parity"]["pass"]]
        if fails:
            line["parity_failed"] = fails
            ok = False
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if dist:
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def ncu_traffic():
    """{kernel class: DRAM bytes per launch, ...} from the checked-in ncu --set full capture of the shipped library."""
    if not os.path.exists(NCU_JSON):
        return {}, None
    d = json.load(open(NCU_JSON))
    return d.get("kernels", {}), d.get("source")


def kernel_roofline(sim, args, S):
    """Times each kernel class of the FUSED hot path (smx_step / smx_step_grad, what the timed region runs) with CUDA events on the
    launching stream, one extra step after the timed region.  Returns the dominant kernel with its algorithmic bytes."""
    try:
        sim.copyframe(S + 1, 0)
        fw = sim.profile_step(0, S, False)
        bw = sim.profile_step(S, S, True)
    except Exception:                               # noqa: BLE001
        return None
    tot = {}
    for d in (fw, bw):
        for k, (ms, n) in d.items():
            a = tot.setdefault(k, [0.0, 0])
            a[0] += ms; a[1] += n
    avg = {k: v[0] / max(v[1], 1) for k, v in tot.items()}
    total_ms = sum(v[0] for v in tot.values())
    cand = [k for k in tot if k in KERNEL_BYTES]
    if not cand:
        return None
    top = max(cand, key=lambda k: tot[k][0])        # largest share of the step
    achieved = KERNEL_BYTES[top] * args.n * args.batch / (avg[top] * 1e-3) / 1e9
    ncu, ncu_src = ncu_traffic()
    tr = ncu.get(top) if (args.n == 1_000_000 and args.batch == 1) else None
    return {"bound": "hbm", "kernel": top, "achieved": achieved, "unit": "GB/s",
            "traffic": tr["dram_bytes"] if tr else None,
            "traffic_source": (ncu_src + " (dram__bytes_read.sum + dram__bytes_write.sum per launch)") if tr else None,
            "l2_red_sectors_per_launch": tr.get("lts_red_sectors") if tr else None,
            "algorithmic_bytes_per_launch": KERNEL_BYTES[top] * args.n * args.batch, "avg_launch_ms": avg[top],
            "share_of_step": tot[top][0] / total_ms if total_ms > 0 else None,
            "kernel_ms": avg, "kernel_launches": {k: v[1] for k, v in tot.items()}, "kernel_total_ms": {k: v[0] for k, v in tot.items()},
            "kernel_frac": {k: KERNEL_BYTES[k] * args.n * args.batch / (avg[k] * 1e-3) / 1e9 / peaks()[0] for k in cand},
            "how": f"CUDA events after every launch of one smx_step(0, {S}) + smx_step_grad({S}, {S}) on the simulator stream, after the timed region"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_cuda(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

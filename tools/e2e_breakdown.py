"""Wall-clock breakdown of the end-to-end step of bench.py (host buffers in, host gradient out)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench


def main():
    args = bench.parse()
    from softmac_b200.engine import MPMSimulator, Primitives
    S = args.substeps
    cfg = bench.workload_cfg(args, S + 2)
    sim = MPMSimulator(cfg, Primitives(primitives=[], max_timesteps=S + 2), env_dt=5 * bench.DT, device=0, sort_every=args.sort_every, flags=args.flags)
    st, seed = bench.make_inputs(args, 0)
    names = ["reset", "clear_grads", "add_x_grad", "step", "step_grad", "get_grad", "get_state_grad"]
    acc = {k: [] for k in names}
    for r in range(4):
        t = [time.perf_counter()]
        sim.reset(st); sim.synchronize(); t.append(time.perf_counter())
        sim.clear_all_gradients(); sim.synchronize(); t.append(time.perf_counter())
        sim.add_x_grad(S, seed); sim.synchronize(); t.append(time.perf_counter())
        sim.step(0, S); sim.synchronize(); t.append(time.perf_counter())
        sim.step_grad(S, S); sim.synchronize(); t.append(time.perf_counter())
        xg, vg = sim.get_grad(0); t.append(time.perf_counter())
        g0 = sim.get_state_grad(0); t.append(time.perf_counter())
        if r:
            for k, a, b in zip(names, t[:-1], t[1:]):
                acc[k].append((b - a) * 1e3)
    tot = 0
    for k in names:
        m = float(np.median(acc[k])); tot += m
        print(f"{k:16s} {m:8.2f} ms")
    tot -= float(np.median(acc["get_state_grad"]))       # the bench reads back through get_grad
    print(f"{'total':16s} {tot:8.2f} ms  -> {args.n * S / tot / 1e6:.3f} G particle-substeps/s (with get_grad as the read-back)")


if __name__ == "__main__":
    main()

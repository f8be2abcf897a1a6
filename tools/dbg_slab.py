import os, sys, numpy as np
os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import scenes
from harness import sim_cfg, rel_l2
from softmac_b200.engine import MPMSimulator
from softmac_b200.slabs import SlabCluster
rng = np.random.default_rng(31)
n, steps, n_grid = 20000, 7, 64
st = scenes.blob_state(n, rng, center=(0.5, 0.3, 0.5), width=0.5, vel=0.5, Fdev=0.003, Cdev=0.5)
st[:, 1] = 0.3 + (st[:, 1] - 0.3) * 0.3
st = st.astype(np.float32).astype(np.float64)
for sort_every in (3, 100):
    for mode in ("legacy", "peer-substep", "peer-step"):
        cfg = sim_cfg(n, n_grid=n_grid, max_steps=steps + 2)
        ref = MPMSimulator(cfg, (), env_dt=1e-3, sort_every=sort_every); ref.reset(st)
        clu = SlabCluster(cfg, 2, st, peer=(mode != "legacy"), env_dt=1e-3, sort_every=sort_every)
        errs = []
        if mode == "peer-step":
            ref.step(0, steps); clu.step(0, steps)
            errs.append(rel_l2(clu.get_state(steps)[:, :3], ref.get_state(steps)[:, :3]))
        else:
            for f in range(steps):
                ref.substep(f); clu.substep(f)
                errs.append(rel_l2(clu.get_state(f + 1)[:, :3], ref.get_state(f + 1)[:, :3]))
        a, b = clu.get_state(steps), ref.get_state(steps)
        bad = np.nonzero(np.abs(a[:, :3] - b[:, :3]).max(1) > 1e-6)[0]
        print(sort_every, mode, ["%.1e" % e for e in errs], "bad", len(bad), "x of bad", np.round(b[bad[:6], 0], 3) if len(bad) else "", "bound", clu.bounds[1] * 4 / n_grid,
              [r.halo_status() for r in clu.ranks] if mode != "legacy" else "")

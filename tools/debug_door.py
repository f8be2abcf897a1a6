"""GPU debugging aid: the door-like episode of tests/test_coupling_episode.py, oracle backend vs CUDA, one loss term at a time."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import test_coupling_episode as T
from harness import rel_l2, cosine

np.set_printoptions(precision=4, linewidth=200)
env_steps = int(os.environ.get("STEPS", 12))
actions = np.tile([[45.0, 30.0, -600.0]], (env_steps, 1, 1)) * (1 + 0.05 * np.arange(env_steps))[:, None, None]
frames = [env_steps, env_steps - 3]
for w in ((1.0, 0.1, 5.0), (1.0, 0.0, 0.0), (0.0, 0.1, 0.0), (0.0, 0.0, 5.0)):
    res = {}
    for backend in ("oracle", "cuda"):
        env = T.build_door(backend, env_steps=env_steps)
        env.loss.weight = w
        res[backend] = T.run_door(env, actions, frames)
        env.loss.weight = w
    (lo, go, ro), (lg, gg, rg) = res["oracle"], res["cuda"]
    go, gg = np.asarray(go).reshape(env_steps, 3), np.asarray(gg).reshape(env_steps, 3)
    print("weights", w, "loss", lo, lg, "rigid", ro[:2], rg[:2], "cos", cosine(gg, go), "rel", rel_l2(gg, go))
    print(np.concatenate([go, gg, gg - go], axis=1))

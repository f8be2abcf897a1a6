#!/usr/bin/env python
"""Prints the measured CUDA-vs-oracle errors of the parity scenes (run on a GPU box; output kept in profiles/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import scenes  # noqa: E402
from harness import Pair, rel_l2, cosine, prim_states_for  # noqa: E402

COLS = dict(x=slice(0, 3), v=slice(3, 6), F=slice(6, 15), C=slice(15, 24))
NAMES = {(0, 0): "corotated plastic", (1, 0): "corotated elastic", (2, 0): "corotated liquid", (1, 1): "neo-Hookean elastic", (2, 1): "neo-Hookean liquid"}


def scene(rng, n=4000, P=2, **kw):
    pair = Pair(n, tables=[scenes.sphere_table() for _ in range(P)], prim_params=[(0.4 + 0.3 * i, 666.) for i in range(P)], **kw)
    for i, s13 in enumerate(prim_states_for(rng, P, (0.5, 0.3, 0.5))):
        pair.set_prim_state(i, 0, pair.cfg.max_steps, s13)
    pair.reset(scenes.blob_state(n, rng))
    pair.clear_ext_f()
    return pair


def main():
    print("# CUDA (fp32) vs f64 oracle, one substep forward + adjoint, 4000 particles, 2 sphere primitives in contact, 32^3 grid")
    print("# columns: relative L2 error of x, v, F, C after the substep | relative L2 / cosine of the adjoint of frame 0 | wrench rel L2 | primitive-state adjoint cosine")
    rows = [(ctype, cname, ptype, model, mname, None) for ctype, cname in ((2, "mixed (forecast)"), (0, "grid"), (1, "particle"))
            for (ptype, model), mname in NAMES.items() if ctype == 2 or (ptype, model) == (0, 0)]
    rows.append((2, "mixed (forecast)", 0, 0, "corotated von Mises", 100.0))      # soft_cloth's flow rule; about half of the blob yields
    for ctype, cname, ptype, model, mname, ys in rows:
        if True:
            rng = np.random.default_rng(1000 + 10 * ctype + 3 * model + ptype)
            pair = scene(rng, ptype=ptype, material_model=model, collision_type=ctype, yield_stress=ys)
            pair.substep(0)
            ref, got = pair.orc.get_frame(1), pair.gpu.get_state(1)
            fwd = " ".join(f"{k}={rel_l2(got[:, s], ref[:, s]):.1e}" for k, s in COLS.items())
            wr = max(rel_l2(pair.prims[i].get_ext_f(), pair.orc.get_ext_f(i)) for i in range(pair.P))
            cot = rng.normal(size=(pair.n, 24)).astype(np.float32).astype(np.float64)
            ext = [rng.normal(size=6).astype(np.float32).astype(np.float64) for _ in range(pair.P)]
            pair.orc.clear_grads(); pair.gpu.clear_all_gradients()
            pair.orc.add_frame_grad(1, cot); pair.gpu.add_state_grad(1, cot)
            for i in range(pair.P):
                pair.orc.set_ext_f_grad(i, ext[i])
            pair.orc.substep_grad(0); pair.gpu.substep_grad(0, ext_f_grad=ext)
            go, gg = pair.orc.get_frame_grad(0), pair.gpu.get_state_grad(0)
            pc = min(cosine(pair.prims[i].get_all_states_grad(0), pair.orc.get_primitive_state_grad(i, 0)) for i in range(pair.P))
            print(f"{cname:17s} {mname:20s} | {fwd} | adj rel={rel_l2(gg, go):.1e} cos={cosine(gg, go):.8f} | wrench={wr:.1e} | prim-adj cos={pc:.8f}")


if __name__ == "__main__":
    main()

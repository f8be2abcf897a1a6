"""world_size-2 gloo test (CPU) of the N>1 host logic: rollout sharding + gradient all-reduce."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import sys, json
    import numpy as np
    sys.path.insert(0, %r)
    from softmac_b200 import rollouts
    rank, ws, _ = rollouts.init(backend="gloo")
    N = 7          # not divisible by 2: uneven shards
    def fn(k):
        rng = np.random.default_rng(k)
        return float(k), rng.normal(size=(5, 2))
    g, losses = rollouts.run_rollouts(N, fn, (5, 2), device="cpu")
    open(sys.argv[1] + "/out%%d.json" %% rank, "w").write(json.dumps({"rank": rank, "mine": sorted(losses), "g": g.tolist()}))
""") % ROOT


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_covers_every_rollout_once():
    from softmac_b200.rollouts import shard
    for n, w in ((64, 8), (7, 2), (3, 8), (0, 4)):
        ids = sorted(k for r in range(w) for k in shard(n, r, w))
        assert ids == list(range(n))
        sizes = [len(shard(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_allreduce_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    outs = [json.loads((tmp_path / f"out{k}.json").read_text()) for k in range(2)]
    expect = np.mean([np.random.default_rng(k).normal(size=(5, 2)) for k in range(7)], axis=0)
    for o in outs:
        assert np.allclose(o["g"], expect, atol=1e-12)
    assert sorted(outs[0]["mine"] + outs[1]["mine"]) == list(range(7))

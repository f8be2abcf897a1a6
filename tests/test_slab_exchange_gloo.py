"""world_size-3 gloo test (CPU) of the particle hand-over between slab ranks (softmac_b200/slabs.py: exchange_rows /
exchange_back): variable-size payloads to both x-neighbours, global ids carried bit-exactly, empty messages, and the
backward hand-over returning adjoint rows to the rank that sent the particles."""
import json
import os
import subprocess
import sys
import textwrap

import numpy as np

from test_rollouts_gloo import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import sys, json
    import numpy as np, torch
    import torch.distributed as dist
    sys.path.insert(0, %r)
    from softmac_b200 import rollouts
    from softmac_b200.slabs import exchange_rows, exchange_back
    rank, ws, _ = rollouts.init(backend="gloo")
    # rank r sends (r + 1) rows down and 2 * r rows up (rank 0 sends nothing up); row content and ids identify the sender
    def rows(k, tag):
        return torch.full((k, 24), float(tag)) + torch.arange(k).reshape(-1, 1), torch.arange(k, dtype=torch.int64) + 1000000 * tag + 16777217
    send_lo, send_hi = rows(rank + 1, 10 * rank + 1), rows(2 * rank, 10 * rank + 2)
    recv_lo, recv_hi = exchange_rows(dist, rank, ws, send_lo, send_hi)
    # adjoints of what was received travel back, scaled so the origin can be checked
    back_lo, back_hi = exchange_back(dist, rank, ws, 2 * recv_lo[0], 3 * recv_hi[0],
                                     send_lo[0].shape[0] if rank > 0 else 0, send_hi[0].shape[0] if rank < ws - 1 else 0)
    out = {"rank": rank, "recv_lo": recv_lo[0].tolist(), "gid_lo": recv_lo[1].tolist(), "recv_hi": recv_hi[0].tolist(), "gid_hi": recv_hi[1].tolist(),
           "back_lo": back_lo.tolist(), "back_hi": back_hi.tolist(), "sent_lo": send_lo[0].tolist(), "sent_hi": send_hi[0].tolist()}
    open(sys.argv[1] + "/out%%d.json" %% rank, "w").write(json.dumps(out))
    dist.destroy_process_group()
""") % ROOT


def test_row_exchange_three_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=3", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), str(script), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    o = [json.loads((tmp_path / f"out{k}.json").read_text()) for k in range(3)]

    def expect(k, tag):
        return (np.full((k, 24), float(tag)) + np.arange(k).reshape(-1, 1)).tolist(), (np.arange(k) + 1000000 * tag + 16777217).tolist()
    for rk in range(3):
        # from the lower neighbour: what rank rk-1 sent UP (2 * (rk - 1) rows); from the upper one: what rk+1 sent DOWN (rk + 2 rows)
        lo = expect(2 * (rk - 1), 10 * (rk - 1) + 2) if rk > 0 else ([], [])
        hi = expect(rk + 2, 10 * (rk + 1) + 1) if rk < 2 else ([], [])
        assert o[rk]["recv_lo"] == lo[0] and o[rk]["gid_lo"] == lo[1]          # ids above 2^24 survive (int32 bit pattern, not a float value)
        assert o[rk]["recv_hi"] == hi[0] and o[rk]["gid_hi"] == hi[1]
        # the adjoints come back to the sender: rows sent down return scaled by 3 (they were the receiver's "from hi" part), rows sent up by 2
        assert o[rk]["back_lo"] == ((3 * np.array(o[rk]["sent_lo"])).tolist() if rk > 0 else [])
        assert o[rk]["back_hi"] == ((2 * np.array(o[rk]["sent_hi"])).tolist() if rk < 2 and len(o[rk]["sent_hi"]) else [])

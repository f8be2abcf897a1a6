"""Independent rollouts across GPUs: the only way the reference's workload shards without a data-path exchange
(BASELINE config 4: "64 independent demo_grip rollouts sharded over 8 x B200 with gradient allreduce").

One process per GPU (torchrun); every rank owns whole rollouts -- no collective inside the substep loop -- and the
per-rollout action gradients (tiny: steps x action_dim) are reduced once per optimisation step over NCCL (or gloo in
the CPU tests).  torch is used only here, for ``torch.distributed``; the simulator itself never sees a torch type.
"""
import os

import numpy as np


def world():
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def shard(n_rollouts, rank, world_size):
    """Rollout ids owned by `rank`: round-robin, so every rank gets floor or ceil of n/world (k mod world == rank)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    return list(range(rank, n_rollouts, world_size))


def init(backend=None, device_index=None):
    """Initialise torch.distributed when launched under torchrun; no-op for a single process."""
    import torch.distributed as dist
    rank, ws, local = world()
    if ws == 1 or dist.is_initialized():
        return rank, ws, local
    import torch
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {}
    if backend == "nccl":
        torch.cuda.set_device(local if device_index is None else device_index)
        kw["device_id"] = torch.device("cuda", local if device_index is None else device_index)
    dist.init_process_group(backend, **kw)
    return rank, ws, local


def allreduce_gradients(local_grads, n_rollouts, device=None):
    """Mean over ALL rollouts of the per-rollout gradients.

    local_grads: array (n_local, ...) of this rank's rollouts (n_local may be 0).  Returns a numpy array of shape (...).
    A sum all-reduce of (sum of local gradients); the division by the global count happens after the reduce, so ranks
    with fewer rollouts are weighted correctly."""
    import torch
    import torch.distributed as dist
    g = np.asarray(local_grads, dtype=np.float64)
    tot = g.sum(axis=0) if g.shape[0] > 0 else np.zeros(g.shape[1:], dtype=np.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if device is None:
            device = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.as_tensor(tot, dtype=torch.float64, device=device).contiguous()
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot = t.cpu().numpy()
    return tot / float(n_rollouts)


def run_rollouts(n_rollouts, rollout_fn, grad_shape, device=None):
    """Runs `rollout_fn(k) -> (loss, action_grad)` for every rollout k owned by this rank and returns
    (mean gradient over all rollouts, {k: loss} of the local ones)."""
    rank, ws, _ = world()
    mine = shard(n_rollouts, rank, ws)
    grads, losses = [], {}
    for k in mine:
        loss, g = rollout_fn(k)
        g = np.asarray(g, dtype=np.float64)
        if g.shape != tuple(grad_shape):
            raise ValueError(f"rollout {k}: gradient shape {g.shape}, expected {tuple(grad_shape)}")
        grads.append(g)
        losses[k] = float(loss)
    local = np.stack(grads) if grads else np.zeros((0,) + tuple(grad_shape))
    return allreduce_gradients(local, n_rollouts, device=device), losses

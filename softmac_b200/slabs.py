"""Spatial slab decomposition of one large scene across the GPUs of a box (BASELINE config 5).

The grid is cut along x into slabs of whole 4-node block columns; every rank owns the particles whose base cell lies in
its slab and runs the ordinary substep kernels on them, on a grid array with the GLOBAL index space (only the active
blocks of the slab are ever touched).  The block-major grid layout makes an x-block column one contiguous range of
nb^2 KB, so the halo of a boundary -- the last column of the left rank and the first column of the right rank -- is a
single contiguous 2-column range that is exchanged with one send/recv pair and summed on both sides:

    forward  : begin (clear + P2G)       -> sum g_in halo    -> end (grid update, G2P)
    adjoint  : begin (restore, G2P adj)  -> sum gg_out halo  -> end (grid adjoint, P2G adj)

Both ranks then hold identical, complete values on the shared columns (a + b == b + a) and update them redundantly, so
no second (broadcast) exchange is needed.  Transport (``peer=True``, the default of ``DistSlab``): PEER MEMORY -- every rank
owns one allocation of receive slots that its x-neighbours write into directly over NVLink (CUDA IPC between the processes);
the library pushes the non-empty halo blocks, flags, waits and adds on the simulator's stream inside ``smx_step`` /
``smx_step_grad`` (``smx_slab_halo_*``), so a whole ``step(count)`` is ONE native call per rank with no host round trip, no
NCCL call on the data path and the cross-substep fusion (G2P2G, fused adjoint pair, deferred grid records) intact;
``torch.distributed`` only carries the IPC handles at set-up and the read-out reductions.  ``peer=False`` keeps the first
transport: the dense 2-column range through ``torch.distributed`` P2P (NCCL) with one Python round trip per phase, or plain
tensor adds for several ranks emulated in one process on one stream.

With the forecast contact model two more halo operations are needed: the contact kernel scatters velocity corrections
into g_out, so after it every rank adds the neighbour's part of (g_out - g_mix) on the halo; in the adjoint the contact
kernel gathers gg_out (complete after the first halo sum) and scatters into gg_mix, whose halo is then summed too.
Wrenches and primitive-state adjoints are per-rank partial sums and are reduced when read.

Particle migration (``MigratingSlabCluster`` / ``DistSlab(migrate_every=E)``): ownership is re-established every E substeps.
A rank is then a SEQUENCE of simulator handles, one per epoch of E substeps; at an epoch boundary the rows of the last frame
are taken on the device (``smx_get_state_dev``), the particles whose base cell left the slab are sent to the x-neighbour
(24 floats + the global id per particle), and the next epoch's handle is reset from [staying | received-from-lo |
received-from-hi] rows without a host round trip (``smx_reset_dev``).  In the backward pass the adjoint of the first frame of
an epoch is split the same way, the received parts travel back to the rank that sent those particles, and the reassembled
rows become a loss seed of the previous epoch's last frame (``smx_add_state_grad_dev``): the chain rule across a migration is
a permutation.  Without migration (``SlabCluster`` / ``DistSlab``) ownership is fixed at reset and a particle may drift up to
two cells out of its slab before the counters flag it.  Migration with primitives is not built (variant A of config 5 has none).
"""
import ctypes as C

import numpy as np

from ._capi import lib, check, vp, SMX_FLAG_EXTERNAL_STREAM


class _DevArray:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def choose_bounds(x, n_ranks, n_grid):
    """Block-column boundaries [b_0 = 0, ..., b_R = nb] with ~equal particle counts and >= 2 columns per slab."""
    nb = n_grid // 4
    base = np.clip((np.asarray(x, dtype=np.float32) * np.float32(n_grid) - np.float32(0.5)).astype(np.int32), 0, n_grid - 3)
    hist = np.bincount(base >> 2, minlength=nb)
    cum = np.concatenate([[0], np.cumsum(hist)])
    bounds = [0]
    for r in range(1, n_ranks):
        b = int(np.searchsorted(cum, cum[-1] * r / n_ranks))
        b = max(b, bounds[-1] + 2)
        bounds.append(b)
    bounds.append(nb)
    for r in range(n_ranks - 1, 0, -1):         # keep >= 2 columns at the high end too
        bounds[r] = min(bounds[r], bounds[r + 1] - 2)
    assert all(bounds[r + 1] - bounds[r] >= 2 for r in range(n_ranks)), "grid too small for this many slabs"
    return bounds


class SlabRank:
    """One rank: a simulator over the particles of its slab."""

    def __init__(self, cfg, rank, bounds, state, device=0, use_torch_stream=True, primitives=(), **sim_kw):
        import copy
        import torch
        from .engine.mpm_simulator import MPMSimulator
        self.rank, self.n_ranks, self.bounds = rank, len(bounds) - 1, list(bounds)
        self.lo, self.hi = bounds[rank], bounds[rank + 1]
        n_grid = int(128 * cfg.quality * 0.5)
        self.nb = n_grid // 4
        st = np.asarray(state, dtype=np.float64)
        base = np.clip((st[:, 0].astype(np.float32) * np.float32(n_grid) - np.float32(0.5)).astype(np.int32), 0, n_grid - 3) >> 2
        self.ids = np.nonzero((base >= self.lo) & (base < self.hi))[0]
        c = copy.deepcopy(cfg)
        c.n_particles = len(self.ids)
        self.device = device
        # run on torch's current stream so that the halo sums / NCCL P2P (torch ops) are ordered with the kernels
        stream = torch.cuda.current_stream(device).cuda_stream if use_torch_stream else None
        flags = sim_kw.pop("flags", 0) | (SMX_FLAG_EXTERNAL_STREAM if use_torch_stream else 0)
        self.primitives = primitives
        self.sim = MPMSimulator(c, primitives, device=device, stream=stream or None, flags=flags, **sim_kw)
        check(lib().smx_set_slab(self.sim._h, self.lo, self.hi, int(rank > 0), int(rank < self.n_ranks - 1)))
        self.sim.reset(st[self.ids] if st.shape[1] == 24 else st[self.ids, :3])
        self._views = {}

    def halo(self, which, side):
        """float32 torch view of the 2-column halo range of grid array `which` at the 'lo' or 'hi' boundary."""
        import torch
        key = (which, side)
        if key not in self._views:
            p, n = vp(), C.c_int64()
            check(lib().smx_grid_dev(self.sim._h, which, C.byref(p), C.byref(n)))
            full = torch.as_tensor(_DevArray(p.value, n.value * 4), device=f"cuda:{self.device}")
            col = self.nb * self.nb * 64 * 4
            b = self.lo if side == "lo" else self.hi
            self._views[key] = full[(b - 1) * col:(b + 1) * col]
        return self._views[key]

    # halo exchange over peer memory --------------------------------------------------------------------------------
    def halo_export(self):
        """(device pointer, 64-byte CUDA IPC handle) of this rank's halo allocation."""
        base, h = vp(), (C.c_ubyte * 64)()
        check(lib().smx_slab_halo_export(self.sim._h, C.byref(base), C.cast(h, vp)))
        return int(base.value), bytes(h)

    def halo_connect(self, side, base=None, handle=None):
        """side 0: the rank below, side 1: the rank above; `base` (same process) or `handle` (another process)."""
        hb = (C.c_ubyte * 64).from_buffer_copy(handle) if handle is not None else None
        check(lib().smx_slab_halo_connect(self.sim._h, int(side), vp(base) if base is not None else None, C.cast(hb, vp) if hb is not None else None))

    def halo_status(self):
        out = (C.c_int64 * 3)()
        check(lib().smx_slab_halo_status(self.sim._h, out))
        return dict(timeouts=int(out[0]), exchanges=int(out[1]), halo_bytes=int(out[2]))

    # phases ------------------------------------------------------------------------------------------------------
    def begin(self, f):
        check(lib().smx_substep_begin(self.sim._h, int(f)))

    def end(self, f):
        check(lib().smx_substep_end(self.sim._h, int(f)))

    def grad_begin(self, f):
        check(lib().smx_substep_grad_begin(self.sim._h, int(f)))

    def grad_end(self, f):
        check(lib().smx_substep_grad_end(self.sim._h, int(f)))

    def mid(self, f):
        check(lib().smx_substep_mid(self.sim._h, int(f)))

    def grad_mid(self, f):
        check(lib().smx_substep_grad_mid(self.sim._h, int(f)))

    def has_contact(self):
        return self.sim.collision_type == 2 and any(self.sim.primitives_contact)

    def scatter_to_global(self, local, n_global):
        out = np.zeros((n_global,) + local.shape[1:])
        out[self.ids] = local
        return out


class SlabCluster:
    """All ranks emulated in ONE process on one device (tests / single-GPU use): same phases, halos summed directly."""

    def __init__(self, cfg, n_ranks, state, device=0, make_primitives=None, peer=False, **sim_kw):
        """make_primitives() -> a fresh Primitives container (every rank needs its own, bound to its handle).
        peer: the ranks exchange their halos through the library's peer-memory path (smx_slab_halo_*), every rank on a stream of its
        own and driven by a host thread of its own -- the single-device rehearsal of one process per GPU."""
        n_grid = int(128 * cfg.quality * 0.5)
        self.n = len(state)
        self.peer = bool(peer)
        self.bounds = choose_bounds(np.asarray(state)[:, 0], n_ranks, n_grid)
        self.ranks = [SlabRank(cfg, r, self.bounds, state, device=device, primitives=make_primitives() if make_primitives else (),
                               use_torch_stream=not self.peer, **sim_kw) for r in range(n_ranks)]
        if self.peer:
            import os
            if os.environ.get("CUDA_MODULE_LOADING", "").upper() != "EAGER":
                raise RuntimeError("SlabCluster(peer=True) emulates the ranks in one process: set CUDA_MODULE_LOADING=EAGER before CUDA is initialised "
                                   "(a rank that waits on a device flag would otherwise block the lazy load of a kernel its neighbour launches for the "
                                   "first time); one process per GPU (DistSlab) does not need it")
            bases = [r.halo_export()[0] for r in self.ranks]
            for i, r in enumerate(self.ranks):
                if i > 0:
                    r.halo_connect(0, base=bases[i - 1])
                if i < n_ranks - 1:
                    r.halo_connect(1, base=bases[i + 1])

    def _threads(self, fn):
        """fn(rank) on one host thread per rank (the native calls release the GIL): the ranks enqueue concurrently, as separate processes would."""
        import threading
        errs = []

        def run(r):
            try:
                fn(r)
            except Exception as e:      # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=run, args=(r,)) for r in self.ranks]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        if errs:
            raise errs[0]

    def step(self, s0, count):
        if self.peer:
            self._threads(lambda r: r.sim.step(s0, count))
        else:
            for f in range(s0, s0 + count):
                self.substep(f)

    def step_grad(self, s1, count):
        if self.peer:
            self._threads(lambda r: r.sim.step_grad(s1, count))
        else:
            for f in range(s1 - 1, s1 - 1 - count, -1):
                self.substep_grad(f)

    def _exchange(self, which):
        for r in range(len(self.ranks) - 1):
            a, b = self.ranks[r].halo(which, "hi"), self.ranks[r + 1].halo(which, "lo")
            t = a + b
            a.copy_(t); b.copy_(t)

    def _exchange_contact(self):
        for r in range(len(self.ranks) - 1):
            L, R = self.ranks[r], self.ranks[r + 1]
            da = L.halo(1, "hi") - L.halo(2, "hi")          # own contact scatter = g_out - g_mix
            db = R.halo(1, "lo") - R.halo(2, "lo")
            L.halo(1, "hi").add_(db); R.halo(1, "lo").add_(da)

    def substep(self, f):
        if self.peer:
            return self._threads(lambda r: r.sim.substep(f))
        contact = self.ranks[0].has_contact()
        for r in self.ranks:
            r.begin(f)
        self._exchange(0)
        if contact:
            for r in self.ranks:
                r.mid(f)
            self._exchange_contact()
        for r in self.ranks:
            r.end(f)

    def substep_grad(self, f):
        if self.peer:
            return self._threads(lambda r: r.sim.substep_grad(f))
        contact = self.ranks[0].has_contact()
        for r in self.ranks:
            r.grad_begin(f)
        self._exchange(3)
        if contact:
            for r in self.ranks:
                r.grad_mid(f)
            self._exchange(4)
        for r in self.ranks:
            r.grad_end(f)

    # reductions over ranks of the per-rank partial sums
    def set_primitive_state(self, i, f0, f1, s13):
        for r in self.ranks:
            r.primitives[i].set_all_states(f0, s13, f_end=f1)

    def clear_ext_f(self):
        for r in self.ranks:
            for p in r.primitives:
                p.clear_ext_f()

    def ext_f(self, i):
        return sum(r.primitives[i].get_ext_f() for r in self.ranks)

    def set_ext_f_grad(self, i, g):
        for r in self.ranks:
            r.primitives[i].set_ext_f_grad(g)

    def primitive_state_grad(self, i, f0, f1):
        return sum(r.primitives[i].get_all_states_grad(f0, f_end=f1) for r in self.ranks)

    def get_state(self, f):
        out = np.zeros((self.n, 24))
        for r in self.ranks:
            out[r.ids] = r.sim.get_state(f)
        return out

    def add_x_grad(self, f, g):
        for r in self.ranks:
            r.sim.add_x_grad(f, np.asarray(g)[r.ids])

    def get_state_grad(self, f):
        out = np.zeros((self.n, 24))
        for r in self.ranks:
            out[r.ids] = r.sim.get_state_grad(f)
        return out

    def counters(self):
        return [r.sim.counters() for r in self.ranks]


class DistSlab:
    """One rank per process (torchrun, NCCL): this process's slab + P2P halo exchange with its x-neighbours."""

    def __init__(self, cfg, state, device=None, make_primitives=None, peer=True, **sim_kw):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.cuda.current_device() if device is None else device
        n_grid = int(128 * cfg.quality * 0.5)
        self.n = len(state)
        self.peer = bool(peer) and self.world > 1
        self.bounds = choose_bounds(np.asarray(state)[:, 0], self.world, n_grid)     # deterministic: same on every rank
        self.r = SlabRank(cfg, self.rank, self.bounds, state, device=self.device, primitives=make_primitives() if make_primitives else (),
                          use_torch_stream=not self.peer, **sim_kw)
        self.primitives = self.r.primitives
        self.sim = self.r.sim
        self._tmp = {}
        if self.peer:
            # one all-gather of the 64-byte IPC handles; from here on the halo sums are peer-memory kernels on the simulator's stream
            handles = [None] * self.world
            dist.all_gather_object(handles, self.r.halo_export()[1])
            if self.rank > 0:
                self.r.halo_connect(0, handle=handles[self.rank - 1])
            if self.rank < self.world - 1:
                self.r.halo_connect(1, handle=handles[self.rank + 1])
            dist.barrier()

    def _exchange(self, which):
        import torch
        dist, ops, pend = self.dist, [], []
        for side, peer in (("lo", self.rank - 1), ("hi", self.rank + 1)):
            if 0 <= peer < self.world:
                v = self.r.halo(which, side)
                t = self._tmp.setdefault((which, side), torch.empty_like(v))
                ops += [dist.P2POp(dist.isend, v, peer), dist.P2POp(dist.irecv, t, peer)]
                pend.append((v, t))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for v, t in pend:
                v.add_(t)

    def _exchange_contact(self):
        import torch
        dist, ops, pend = self.dist, [], []
        for side, peer in (("lo", self.rank - 1), ("hi", self.rank + 1)):
            if 0 <= peer < self.world:
                own = self.r.halo(1, side) - self.r.halo(2, side)        # own contact scatter = g_out - g_mix
                t = self._tmp.setdefault(("c", side), torch.empty_like(own))
                ops += [dist.P2POp(dist.isend, own, peer), dist.P2POp(dist.irecv, t, peer)]
                pend.append((self.r.halo(1, side), t))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for v, t in pend:
                v.add_(t)

    def substep(self, f):
        if self.peer:
            return self.sim.substep(f)
        self.r.begin(f)
        self._exchange(0)
        if self.r.has_contact():
            self.r.mid(f)
            self._exchange_contact()
        self.r.end(f)

    def substep_grad(self, f):
        if self.peer:
            return self.sim.substep_grad(f)
        self.r.grad_begin(f)
        self._exchange(3)
        if self.r.has_contact():
            self.r.grad_mid(f)
            self._exchange(4)
        self.r.grad_end(f)

    def _allreduce(self, a):
        import torch
        t = torch.as_tensor(np.asarray(a, dtype=np.float64), device=f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def ext_f(self, i):
        """Contact wrench on primitive i summed over the ranks (each rank accumulates its own particles' reactions)."""
        return self._allreduce(self.primitives[i].get_ext_f())

    def primitive_state_grad(self, i, f0, f1):
        return self._allreduce(self.primitives[i].get_all_states_grad(f0, f_end=f1))

    def step(self, s0, count):
        if self.peer:
            return self.sim.step(s0, count)     # ONE native call: substeps, halo sums and fusion all inside smx_step
        for f in range(s0, s0 + count):
            self.substep(f)

    def step_grad(self, s1, count):
        if self.peer:
            return self.sim.step_grad(s1, count)
        for f in range(s1 - 1, s1 - 1 - count, -1):
            self.substep_grad(f)

    def halo_status(self):
        return self.r.halo_status() if self.peer else dict(timeouts=0, exchanges=0, halo_bytes=0)

    def add_x_grad(self, f, g_global):
        self.sim.add_x_grad(f, np.asarray(g_global)[self.r.ids])

    def gather_state(self, f):
        """Global (n, 24) state on every rank (all-reduce of the scattered local rows; for tests and read-out)."""
        import torch
        out = torch.zeros((self.n, 24), dtype=torch.float64, device=f"cuda:{self.device}")
        out[torch.as_tensor(self.r.ids, device=out.device)] = torch.as_tensor(self.sim.get_state(f), device=out.device)
        self.dist.all_reduce(out)
        return out.cpu().numpy()


# ------------------------------------------------------------------------------------------------------------------------
# particle migration: a rank as a sequence of epoch handles
# ------------------------------------------------------------------------------------------------------------------------
class _Epoch:
    """The particles one rank owns during the substeps [f0, f0 + E): a simulator handle reset from device rows."""

    def __init__(self, cfg, rank, bounds, rows, gid, f0, epoch_len, device, use_torch_stream, sim_kw, make_primitives=None):
        import copy
        import torch
        from .engine.mpm_simulator import MPMSimulator
        self.f0, self.E, self.n, self.gid, self.device = f0, epoch_len, int(rows.shape[0]), gid, device
        self.primitives = make_primitives() if make_primitives else ()
        if self.n == 0:
            raise RuntimeError(f"slab rank {rank} owns no particles in the epoch starting at substep {f0}")
        c = copy.deepcopy(cfg)
        c.n_particles, c.max_steps = self.n, epoch_len + 2
        kw = dict(sim_kw)
        stream = torch.cuda.current_stream(device).cuda_stream if use_torch_stream else None
        flags = kw.pop("flags", 0) | (SMX_FLAG_EXTERNAL_STREAM if use_torch_stream else 0)
        self.sim = MPMSimulator(c, self.primitives, device=device, stream=stream or None, flags=flags, **kw)
        n_ranks = len(bounds) - 1
        self.lo, self.hi, self.nb = bounds[rank], bounds[rank + 1], int(128 * cfg.quality * 0.5) // 4
        check(lib().smx_set_slab(self.sim._h, self.lo, self.hi, int(rank > 0), int(rank < n_ranks - 1)))
        rows = rows.contiguous()
        check(lib().smx_reset_dev(self.sim._h, rows.data_ptr()))
        self.parts = None           # (idx_stay, idx_lo, idx_hi, n_recv_lo, n_recv_hi): how this epoch's rows came out of the previous one
        self._views = {}

    halo = SlabRank.halo
    has_contact = SlabRank.has_contact

    def load_primitive_states(self, series):
        """series[i]: {global frame: s13}; this handle holds the local frames [0, E + 2) = global [f0, f0 + E + 2)."""
        for i, frames in enumerate(series):
            for g, s13 in frames.items():
                if self.f0 <= g < self.f0 + self.E + 2:
                    self.primitives[i].set_all_states(g - self.f0, s13)
            self.primitives[i].clear_ext_f()

    def rows_dev(self, local_f, grad=False):
        import torch
        out = torch.empty((self.n, 24), dtype=torch.float32, device=f"cuda:{self.device}")
        fn = lib().smx_get_state_grad_dev if grad else lib().smx_get_state_dev
        check(fn(self.sim._h, int(local_f), out.data_ptr()))
        return out


class MigratingSlabRank:
    def __init__(self, cfg, rank, bounds, state, migrate_every, device=0, use_torch_stream=True, make_primitives=None, **sim_kw):
        import torch
        self.cfg, self.rank, self.bounds, self.E = cfg, rank, list(bounds), int(migrate_every)
        self.make_primitives = make_primitives
        self.series = []            # per primitive: {global frame: s13}, replayed into every new epoch handle
        self.ext_carry = None       # wrench accumulated by earlier epoch handles since the last clear_ext_f
        self.n_ranks, self.device, self.use_torch_stream, self.sim_kw = len(bounds) - 1, device, use_torch_stream, sim_kw
        self.n_grid = int(128 * cfg.quality * 0.5)
        st = np.asarray(state, dtype=np.float64)
        if st.shape[1] == 3:
            full = np.zeros((len(st), 24)); full[:, :3] = st; full[:, 6] = full[:, 10] = full[:, 14] = 1
            st = full
        col = np.clip((st[:, 0].astype(np.float32) * np.float32(self.n_grid) - np.float32(0.5)).astype(np.int32), 0, self.n_grid - 3) >> 2
        ids = np.nonzero((col >= bounds[rank]) & (col < bounds[rank + 1]))[0]
        dev = f"cuda:{device}"
        rows = torch.as_tensor(st[ids].astype(np.float32), device=dev)
        self.epochs = [_Epoch(cfg, rank, bounds, rows, torch.as_tensor(ids, device=dev), 0, self.E, device, use_torch_stream, sim_kw, make_primitives)]
        self.series = [dict() for _ in self.epochs[0].primitives]
        self.ext_carry = np.zeros((len(self.series), 6))
        self.migrated = 0
        self._pool = {}             # particle count -> idle epoch handles (a repeated rollout re-uses them: no cudaMalloc / cudaFree)

    def rewind(self):
        """Back to the first epoch (its frames are kept); the later handles go to the pool."""
        for ep in self.epochs[1:]:
            self._pool.setdefault(ep.n, []).append(ep)
        del self.epochs[1:]
        self.migrated = 0

    def _new_epoch(self, rows, gid, f0):
        idle = self._pool.get(int(rows.shape[0]))
        if idle:
            ep = idle.pop()
            ep.f0, ep.gid = f0, gid
            ep.sim.clear_all_gradients()
            check(lib().smx_reset_dev(ep.sim._h, rows.contiguous().data_ptr()))
        else:
            ep = _Epoch(self.cfg, self.rank, self.bounds, rows, gid, f0, self.E, self.device, self.use_torch_stream, self.sim_kw, self.make_primitives)
        if self.series:
            for i in range(len(self.series)):               # the wrench accumulated so far stays with this rank
                self.ext_carry[i] += self.epochs[-1].primitives[i].get_ext_f()
            ep.load_primitive_states(self.series)
        return ep

    # primitive plumbing (global frames) --------------------------------------------------------------------------
    def set_primitive_state(self, i, f0, f1, s13):
        s13 = np.asarray(s13, dtype=np.float64)
        for g in range(f0, f1):
            self.series[i][g] = s13
        for ep in self.epochs:
            a, b = max(f0, ep.f0), min(f1, ep.f0 + self.E + 2)
            if a < b:
                ep.primitives[i].set_all_states(a - ep.f0, s13, f_end=b - ep.f0)

    def clear_ext_f(self):
        self.ext_carry[:] = 0
        for p in self.epochs[-1].primitives:
            p.clear_ext_f()

    def ext_f(self, i):
        return self.ext_carry[i] + self.epochs[-1].primitives[i].get_ext_f()

    def primitive_state_grad(self, i, f0, f1):
        """Sum over the global frames [f0, f1) of this rank's part of the primitive-state adjoint (frame g is collided against by
        substep g, which ran in epoch g // E)."""
        out = np.zeros(13)
        for e, ep in enumerate(self.epochs):
            a, b = max(f0, e * self.E), min(f1, (e + 1) * self.E)
            if a < b:
                out += ep.primitives[i].get_all_states_grad(a - ep.f0, f_end=b - ep.f0)
        return out

    # epoch of substep f (input frame f) / of frame f (the latest epoch that holds it)
    def epoch_of_substep(self, f):
        return self.epochs[f // self.E]

    def epoch_of_frame(self, f):
        return self.epochs[min(f // self.E, len(self.epochs) - 1)]

    def needs_epoch(self, f):
        return f // self.E >= len(self.epochs)

    # forward migration, phase 1: rows of the last frame of the current epoch, split by the slab the base cell lies in
    def split_last(self):
        import torch
        ep = self.epochs[-1]
        rows = ep.rows_dev(self.E)
        col = torch.clamp((rows[:, 0] * float(self.n_grid) - 0.5).to(torch.int32), 0, self.n_grid - 3) >> 2
        go_lo = (col < ep.lo) if self.rank > 0 else torch.zeros_like(col, dtype=torch.bool)
        go_hi = (col >= ep.hi) if self.rank < self.n_ranks - 1 else torch.zeros_like(col, dtype=torch.bool)
        idx = torch.arange(ep.n, device=rows.device)
        self._pend = (rows, idx[~(go_lo | go_hi)], idx[go_lo], idx[go_hi])
        rows_lo, rows_hi = rows[self._pend[2]], rows[self._pend[3]]
        return (rows_lo, ep.gid[self._pend[2]]), (rows_hi, ep.gid[self._pend[3]])

    # phase 2: the next epoch from [staying | received from lo | received from hi]
    def start_epoch(self, recv_lo, recv_hi):
        import torch
        ep = self.epochs[-1]
        rows, i_stay, i_lo, i_hi = self._pend
        new_rows = torch.cat([rows[i_stay], recv_lo[0], recv_hi[0]])
        new_gid = torch.cat([ep.gid[i_stay], recv_lo[1], recv_hi[1]])
        ne = self._new_epoch(new_rows, new_gid, ep.f0 + self.E)
        ne.parts = (i_stay, i_lo, i_hi, int(recv_lo[0].shape[0]), int(recv_hi[0].shape[0]))
        self.migrated += int(i_lo.numel() + i_hi.numel())
        self.epochs.append(ne)
        self._pend = None

    # backward migration, phase 1: adjoint of the first frame of epoch e, parts that go back to the neighbours
    def split_first_grad(self, e):
        ep = self.epochs[e]
        g = ep.rows_dev(0, grad=True)
        i_stay, i_lo, i_hi, n_rlo, n_rhi = ep.parts
        ns = int(i_stay.numel())
        self._pend_b = (g[:ns], e)
        return g[ns:ns + n_rlo], g[ns + n_rlo:ns + n_rlo + n_rhi]

    # phase 2: seed of the last frame of epoch e - 1
    def seed_previous(self, back_lo, back_hi):
        import torch
        g_stay, e = self._pend_b
        i_stay, i_lo, i_hi, _, _ = self.epochs[e].parts
        prev = self.epochs[e - 1]
        G = torch.zeros((prev.n, 24), dtype=torch.float32, device=g_stay.device)
        G[i_stay] = g_stay
        G[i_lo] = back_lo
        G[i_hi] = back_hi
        check(lib().smx_add_state_grad_dev(prev.sim._h, self.E, G.contiguous().data_ptr()))
        self._pend_b = None


class _MigratingBase:
    """Epoch bookkeeping shared by the emulated cluster and the NCCL rank."""

    def _local(self, f):
        return f % self.E


class MigratingSlabCluster(_MigratingBase):
    """All ranks in ONE process on one device, with particle migration every `migrate_every` substeps (tests / single GPU)."""

    def __init__(self, cfg, n_ranks, state, migrate_every, device=0, make_primitives=None, **sim_kw):
        n_grid = int(128 * cfg.quality * 0.5)
        self.n, self.E = len(state), int(migrate_every)
        self.bounds = choose_bounds(np.asarray(state)[:, 0], n_ranks, n_grid)
        self.ranks = [MigratingSlabRank(cfg, r, self.bounds, state, migrate_every, device=device, make_primitives=make_primitives, **sim_kw)
                      for r in range(n_ranks)]
        self._bwd_epoch = None

    def _exchange_contact(self, eps):
        for r in range(len(eps) - 1):
            L, R = eps[r], eps[r + 1]
            da = L.halo(1, "hi") - L.halo(2, "hi")          # own contact scatter = g_out - g_mix
            db = R.halo(1, "lo") - R.halo(2, "lo")
            L.halo(1, "hi").add_(db); R.halo(1, "lo").add_(da)

    def set_primitive_state(self, i, f0, f1, s13):
        for r in self.ranks:
            r.set_primitive_state(i, f0, f1, s13)

    def clear_ext_f(self):
        for r in self.ranks:
            r.clear_ext_f()

    def ext_f(self, i):
        return sum(r.ext_f(i) for r in self.ranks)

    def set_ext_f_grad(self, i, g):
        for r in self.ranks:
            for ep in r.epochs:
                ep.primitives[i].set_ext_f_grad(g)

    def primitive_state_grad(self, i, f0, f1):
        return sum(r.primitive_state_grad(i, f0, f1) for r in self.ranks)

    def _exchange(self, eps, which):
        for r in range(len(eps) - 1):
            a, b = eps[r].halo(which, "hi"), eps[r + 1].halo(which, "lo")
            t = a + b
            a.copy_(t); b.copy_(t)

    def _migrate(self):
        import torch
        sends = [r.split_last() for r in self.ranks]
        R = len(self.ranks)
        dev = sends[0][0][0].device
        empty = (torch.empty((0, 24), dtype=torch.float32, device=dev), torch.empty((0,), dtype=torch.int64, device=dev))
        for r, rk in enumerate(self.ranks):
            rk.start_epoch(sends[r - 1][1] if r > 0 else empty, sends[r + 1][0] if r < R - 1 else empty)

    def substep(self, f):
        if self.ranks[0].needs_epoch(f):
            self._migrate()
        eps = [r.epoch_of_substep(f) for r in self.ranks]
        lf = self._local(f)
        for ep in eps:
            check(lib().smx_substep_begin(ep.sim._h, lf))
        self._exchange(eps, 0)
        if eps[0].has_contact():
            for ep in eps:
                check(lib().smx_substep_mid(ep.sim._h, lf))
            self._exchange_contact(eps)
        for ep in eps:
            check(lib().smx_substep_end(ep.sim._h, lf))
        self._bwd_epoch = None

    def substep_grad(self, f):
        e = f // self.E
        if self._bwd_epoch is not None and e == self._bwd_epoch - 1:        # crossing an epoch boundary backwards
            R = len(self.ranks)
            parts = [r.split_first_grad(e + 1) for r in self.ranks]          # (to lo neighbour, to hi neighbour)
            for r, rk in enumerate(self.ranks):
                # what rank r sent to r-1 came back as r-1's "received from hi" part, and symmetrically
                back_lo = parts[r - 1][1] if r > 0 else parts[r][0][:0]
                back_hi = parts[r + 1][0] if r < R - 1 else parts[r][0][:0]
                rk.seed_previous(back_lo, back_hi)
        self._bwd_epoch = e
        eps = [r.epochs[e] for r in self.ranks]
        lf = self._local(f)
        for ep in eps:
            check(lib().smx_substep_grad_begin(ep.sim._h, lf))
        self._exchange(eps, 3)
        if eps[0].has_contact():
            for ep in eps:
                check(lib().smx_substep_grad_mid(ep.sim._h, lf))
            self._exchange(eps, 4)
        for ep in eps:
            check(lib().smx_substep_grad_end(ep.sim._h, lf))

    def _gather(self, f, grad):
        out = np.zeros((self.n, 24))
        for r in self.ranks:
            ep = r.epoch_of_frame(f)
            lf = f - ep.f0
            out[ep.gid.cpu().numpy()] = ep.sim.get_state_grad(lf) if grad else ep.sim.get_state(lf)
        return out

    def get_state(self, f):
        return self._gather(f, False)

    def get_state_grad(self, f):
        return self._gather(f, True)

    def add_x_grad(self, f, g):
        for r in self.ranks:
            ep = r.epoch_of_frame(f)
            ep.sim.add_x_grad(f - ep.f0, np.asarray(g)[ep.gid.cpu().numpy()])

    def clear_all_gradients(self):
        for r in self.ranks:
            for ep in r.epochs:
                ep.sim.clear_all_gradients()
        self._bwd_epoch = None

    def counters(self):
        return [ep.sim.counters() for r in self.ranks for ep in r.epochs]

    def migrated(self):
        return sum(r.migrated for r in self.ranks)


def exchange_rows(dist, rank, world, send_lo, send_hi):
    """Variable-size hand-over of particle rows to the two x-neighbours.  send_* = (rows (k, 24) float32, gid (k,) int64);
    returns (recv_lo, recv_hi) in the same form.  One message of counts, then one payload per neighbour: 24 floats + the global
    id (int32 bit pattern in a 25th float column) per particle."""
    import torch
    dev = send_lo[0].device
    peers = [(rank - 1, send_lo), (rank + 1, send_hi)]
    cnt_out = {p: torch.tensor([snd[0].shape[0]], dtype=torch.int64, device=dev) for p, snd in peers if 0 <= p < world}
    cnt_in = {p: torch.zeros(1, dtype=torch.int64, device=dev) for p in cnt_out}
    ops = []
    for p in cnt_out:
        ops += [dist.P2POp(dist.isend, cnt_out[p], p), dist.P2POp(dist.irecv, cnt_in[p], p)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    pay_out, pay_in, ops = {}, {}, []
    for p, snd in peers:
        if not (0 <= p < world):
            continue
        k_out, k_in = int(snd[0].shape[0]), int(cnt_in[p].item())
        if k_out:
            pay_out[p] = torch.cat([snd[0].to(torch.float32), snd[1].to(torch.int32).view(torch.float32).reshape(-1, 1)], dim=1).contiguous()
            ops.append(dist.P2POp(dist.isend, pay_out[p], p))
        if k_in:
            pay_in[p] = torch.empty((k_in, 25), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, pay_in[p], p))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def unpack(p):
        if p not in pay_in:
            return torch.empty((0, 24), dtype=torch.float32, device=dev), torch.empty((0,), dtype=torch.int64, device=dev)
        t = pay_in[p]
        return t[:, :24].contiguous(), t[:, 24].contiguous().view(torch.int32).to(torch.int64)
    return unpack(rank - 1), unpack(rank + 1)


def exchange_back(dist, rank, world, to_lo, to_hi, n_from_lo, n_from_hi):
    """Backward counterpart: the adjoint rows of the particles received from a neighbour travel back to it; the sizes are known
    on both sides from the forward hand-over (n_from_* = what this rank SENT to that neighbour then)."""
    import torch
    dev, ops, got = to_lo.device, [], {}
    for p, snd, k_in in ((rank - 1, to_lo, n_from_lo), (rank + 1, to_hi, n_from_hi)):
        if not (0 <= p < world):
            continue
        if snd.shape[0]:
            ops.append(dist.P2POp(dist.isend, snd.contiguous(), p))
        if k_in:
            got[p] = torch.empty((k_in, 24), dtype=torch.float32, device=dev)
            ops.append(dist.P2POp(dist.irecv, got[p], p))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    empty = torch.empty((0, 24), dtype=torch.float32, device=dev)
    return got.get(rank - 1, empty), got.get(rank + 1, empty)


class DistMigratingSlab(_MigratingBase):
    """One rank per process (torchrun, NCCL) with particle migration every `migrate_every` substeps."""

    def __init__(self, cfg, state, migrate_every, device=None, make_primitives=None, **sim_kw):
        import torch
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.cuda.current_device() if device is None else device
        n_grid = int(128 * cfg.quality * 0.5)
        self.n, self.E = len(state), int(migrate_every)
        self.bounds = choose_bounds(np.asarray(state)[:, 0], self.world, n_grid)
        self.r = MigratingSlabRank(cfg, self.rank, self.bounds, state, migrate_every, device=self.device, make_primitives=make_primitives, **sim_kw)
        self._tmp = {}
        self._bwd_epoch = None

    def _exchange_contact(self, ep):
        import torch
        dist, ops, pend = self.dist, [], []
        for side, peer in (("lo", self.rank - 1), ("hi", self.rank + 1)):
            if 0 <= peer < self.world:
                own = ep.halo(1, side) - ep.halo(2, side)            # own contact scatter = g_out - g_mix
                t = self._tmp.setdefault(("c", side), torch.empty_like(own))
                ops += [dist.P2POp(dist.isend, own, peer), dist.P2POp(dist.irecv, t, peer)]
                pend.append((ep.halo(1, side), t))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for v, t in pend:
                v.add_(t)

    def _allreduce(self, a):
        import torch
        t = torch.as_tensor(np.asarray(a, dtype=np.float64), device=f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    # primitives: states replicated on every rank, wrench and primitive-state adjoints are sums over ranks (and epochs)
    def set_primitive_state(self, i, f0, f1, s13):
        self.r.set_primitive_state(i, f0, f1, s13)

    def clear_ext_f(self):
        self.r.clear_ext_f()

    def ext_f(self, i):
        return self._allreduce(self.r.ext_f(i))

    def set_ext_f_grad(self, i, g):
        for ep in self.r.epochs:
            ep.primitives[i].set_ext_f_grad(g)

    def primitive_state_grad(self, i, f0, f1):
        return self._allreduce(self.r.primitive_state_grad(i, f0, f1))

    def _exchange(self, ep, which):
        import torch
        dist, ops, pend = self.dist, [], []
        for side, peer in (("lo", self.rank - 1), ("hi", self.rank + 1)):
            if 0 <= peer < self.world:
                v = ep.halo(which, side)
                t = self._tmp.setdefault((which, side), torch.empty_like(v))
                ops += [dist.P2POp(dist.isend, v, peer), dist.P2POp(dist.irecv, t, peer)]
                pend.append((v, t))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
            for v, t in pend:
                v.add_(t)

    def substep(self, f):
        if self.r.needs_epoch(f):
            send_lo, send_hi = self.r.split_last()
            self.r.start_epoch(*exchange_rows(self.dist, self.rank, self.world, send_lo, send_hi))
        ep, lf = self.r.epoch_of_substep(f), self._local(f)
        check(lib().smx_substep_begin(ep.sim._h, lf))
        self._exchange(ep, 0)
        if ep.has_contact():
            check(lib().smx_substep_mid(ep.sim._h, lf))
            self._exchange_contact(ep)
        check(lib().smx_substep_end(ep.sim._h, lf))
        self._bwd_epoch = None

    def substep_grad(self, f):
        e = f // self.E
        if self._bwd_epoch is not None and e == self._bwd_epoch - 1:
            to_lo, to_hi = self.r.split_first_grad(e + 1)
            _, i_lo, i_hi, _, _ = self.r.epochs[e + 1].parts
            self.r.seed_previous(*exchange_back(self.dist, self.rank, self.world, to_lo, to_hi, int(i_lo.numel()), int(i_hi.numel())))
        self._bwd_epoch = e
        ep, lf = self.r.epochs[e], self._local(f)
        check(lib().smx_substep_grad_begin(ep.sim._h, lf))
        self._exchange(ep, 3)
        if ep.has_contact():
            check(lib().smx_substep_grad_mid(ep.sim._h, lf))
            self._exchange(ep, 4)
        check(lib().smx_substep_grad_end(ep.sim._h, lf))

    def step(self, s0, count):
        for f in range(s0, s0 + count):
            self.substep(f)

    def step_grad(self, s1, count):
        for f in range(s1 - 1, s1 - 1 - count, -1):
            self.substep_grad(f)

    def add_x_grad(self, f, g_global):
        ep = self.r.epoch_of_frame(f)
        ep.sim.add_x_grad(f - ep.f0, np.asarray(g_global)[ep.gid.cpu().numpy()])

    def add_state_grad_dev(self, f, g_global_dev):
        """Loss seed of frame f from a device tensor (n_global, 24) float32 in global particle order: gathered by the current
        owners' ids and accumulated on the device (no host round trip)."""
        ep = self.r.epoch_of_frame(f)
        G = g_global_dev[ep.gid].contiguous()
        check(lib().smx_add_state_grad_dev(ep.sim._h, int(f - ep.f0), G.data_ptr()))

    def clear_all_gradients(self):
        for ep in self.r.epochs:
            ep.sim.clear_all_gradients()
        self._bwd_epoch = None

    def rewind(self):
        self.r.rewind()
        self._bwd_epoch = None

    def gather_state(self, f):
        import torch
        ep = self.r.epoch_of_frame(f)
        out = torch.zeros((self.n, 24), dtype=torch.float64, device=f"cuda:{self.device}")
        out[ep.gid] = torch.as_tensor(ep.sim.get_state(f - ep.f0), device=out.device)
        self.dist.all_reduce(out)
        return out.cpu().numpy()

    def migrated(self):
        import torch
        t = torch.tensor([float(self.r.migrated)], dtype=torch.float64, device=f"cuda:{self.device}")
        self.dist.all_reduce(t)
        return int(t.item())

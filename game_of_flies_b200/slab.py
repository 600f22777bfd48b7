"""Slab decomposition of the periodic box over several GPUs.

For N beyond one GPU's sweet spot the box is cut along z into slabs of whole bin
layers.  Each slab is one `SlabEngine` (one GPU, normally one process): it owns the
particles whose bin layer lies in [layer_lo, layer_hi) and, per step,

  1. bins its particles and hands the ones that left to the neighbouring slabs
     (particle migration: fixed-size buffers of records position+type, velocity+id,
     acceleration, the record count in a header word);
  2. sorts by cell, then exchanges its two boundary layers with the neighbours
     (halo-cell exchange: cell-ordered positions, velocities and per-cell counts,
     contiguous slices of the sorted arrays, so nothing is packed);
  3. computes forces for the owned particles over the local grid (owned layers +
     one halo layer each side) and integrates.

Two transports move the buffers:

* `P2PRunner` (default on one node): the slabs map each other's arenas (CUDA IPC
  between processes) and the exchange steps are kernels that STORE into the
  neighbour's memory over NVLink and raise tagged flags there; one-block kernels on
  the consumer's stream wait for the flags.  A step is one C call that enqueues a
  fixed sequence of launches: no host synchronisation, no library collective, no
  size ever returns to the host (csrc/gof_slab_p2p.cuh).
* `SlabRunner` sequences the same phases from Python and moves the buffers with
  NCCL send/recv when every process drives one slab (`DistComm`; gloo in the CPU
  tests), or with plain device copies when several slabs live in one process
  (`LocalComm`).  Its only collectives are tiny: one all-gather of four words per
  slab and step (the halo message sizes) and, with the adaptive timestep, a MAX
  all-reduce of one word; the halo traffic runs under the interior forces.

The physics is that of the single-GPU engine, bit for bit, on either transport: the
same cells, the same order inside a cell (by x, ties by particle id), the same
summation order.
"""
from __future__ import annotations

import contextlib
import os
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib


def split_layers(num_layers: int, world: int):
    """Contiguous, near-equal ranges of z-layers, one per slab."""
    if world < 2:
        raise ValueError("slab decomposition needs at least 2 slabs")
    if num_layers < 2 * world or num_layers < 4:
        raise ValueError(f"{num_layers} bin layers along z cannot be cut into {world} slabs "
                         "(every slab needs >= 2 layers and the box >= 4)")
    bounds = [(r * num_layers) // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


class SlabEngine:
    """One z-slab on one CUDA device (gof_slab_* entry points of include/gof_b200.h)."""

    REC = 12  # floats per migration record

    def __init__(self, parameter_matrix, limits, r_max, layer_lo, layer_hi, capacity,
                 migrate_capacity=None, halo_capacity=None, device=None,
                 wrap_mode=_lib.WRAP_REFERENCE):
        from .device import _require_cuda
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        self.index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        pm = np.ascontiguousarray(parameter_matrix, dtype=np.float64)
        self.num_types = int(pm.shape[1])
        self.has_boids = bool(np.any(pm[-1] == 1))  # Boids alignment reads the neighbours' velocities
        self.grid = _lib.make_grid(limits, r_max)
        self.layer_lo, self.layer_hi = int(layer_lo), int(layer_hi)
        self.nlz = self.layer_hi - self.layer_lo + 2
        self.per_layer = int(self.grid.num_bin_x) * int(self.grid.num_bin_y)
        self.capacity = int(capacity)
        owned = self.layer_hi - self.layer_lo
        mean_layer = max(1, self.capacity // max(owned, 1))
        self.halo_capacity = int(halo_capacity or 4 * mean_layer + 1024)
        # Migration buffers travel whole, so their size is part of the protocol: it must be the
        # SAME on every slab (do not derive it from a per-slab quantity).  8192 records per
        # direction and step is ~25% of a 32k-particle boundary layer; a step moves a particle
        # by at most 0.4 (adaptive) so the expected flux is two orders of magnitude lower.
        self.migrate_capacity = int(migrate_capacity or 8192)
        self.wrap_mode = int(wrap_mode)
        self._handle = C.c_void_p()
        _lib.check(self.lib.gof_slab_create(C.byref(self._handle), self.index, self.capacity, self.num_types,
                                            C.byref(self.grid), pm.ctypes.data, self.layer_lo, self.layer_hi,
                                            self.migrate_capacity, self.halo_capacity))
        nbytes = int(self.lib.gof_engine_arena_bytes(self._handle))
        with torch.cuda.device(self.index):
            self.arena = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
        self._arena_ptr = (self.arena.data_ptr() + 255) // 256 * 256
        _lib.check(self.lib.gof_engine_bind_arena(self._handle, self._arena_ptr, nbytes))
        ptrs = (C.c_void_p * 11)()
        _lib.check(self.lib.gof_slab_pointers(self._handle, ptrs, 11))
        self._ptr = [int(p or 0) for p in ptrs]
        self.n_live = 0
        self.lo_pop = self.hi_pop = 0

    # ---- helpers ----------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.index).cuda_stream

    def _view(self, ptr: int, count: int, dtype) -> torch.Tensor:
        """`count` elements of `dtype` at device address `ptr` (inside the arena) as a tensor."""
        item = torch.empty((), dtype=dtype).element_size()
        off = ptr - self.arena.data_ptr()
        return self.arena[off: off + count * item].view(dtype)

    def close(self):
        if getattr(self, "_handle", None) is not None and self._handle.value:
            torch.cuda.synchronize(self.index)
            self.lib.gof_engine_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state -------------------------------------------------------------------------------
    def upload(self, pos, types, ids, vel=None, acc=None, n=None):
        """pos / vel / acc: three fp32 arrays each; numpy arrays, or (pinned) CPU / CUDA tensors of
        which the first `n` entries are used (no staging copy on the host)."""
        from .engine import _any_ptr

        def f(a, dt):
            if a is None or isinstance(a, torch.Tensor):
                return a
            return np.ascontiguousarray(a, dtype=dt)
        p = [f(a, np.float32) for a in pos]
        v = [f(a, np.float32) for a in vel] if vel is not None else [None] * 3
        a_ = [f(a, np.float32) for a in acc] if acc is not None else [None] * 3
        t, i = f(types, np.int32), f(ids, np.int32)
        count = int(n) if n is not None else int(t.numel() if isinstance(t, torch.Tensor) else t.size)
        _lib.check(self.lib.gof_slab_upload(self._handle, count, *[_any_ptr(x) for x in p], *[_any_ptr(x) for x in v],
                                            *[_any_ptr(x) for x in a_], _any_ptr(t), _any_ptr(i), self._stream()))
        torch.cuda.current_stream(self.index).synchronize()

    def count(self) -> int:
        """Live particles of this slab (synchronises; raises a pending device-side error)."""
        n = C.c_int()
        _lib.check(self.lib.gof_slab_check(self._handle, C.byref(n), self._stream()))
        return int(n.value)

    def download(self):
        n = self.count()
        out = {k: np.empty((3, n), dtype=np.float32) for k in ("pos", "vel", "acc")}
        ids = np.empty(n, dtype=np.int32)
        types = np.empty(n, dtype=np.int32)
        ptrs = [out[k][c].ctypes.data for k in ("pos", "vel", "acc") for c in range(3)]
        _lib.check(self.lib.gof_slab_download(self._handle, *ptrs, ids.ctypes.data, types.ctypes.data, self._stream()))
        torch.cuda.current_stream(self.index).synchronize()
        out["ids"] = ids
        out["types"] = types
        return out

    def download_into(self, pos, vel, acc, ids, types) -> int:
        """The state in rest order into caller-owned buffers of at least `capacity` entries (pinned
        tensors: no allocation, no staging); returns the particle count (synchronises)."""
        from .engine import _any_ptr
        n = self.count()
        ptrs = [_any_ptr(x) for grp in (pos, vel, acc) for x in grp]
        _lib.check(self.lib.gof_slab_download(self._handle, *ptrs, _any_ptr(ids), _any_ptr(types), self._stream()))
        torch.cuda.current_stream(self.index).synchronize()
        return n

    def phase_times(self):
        """Per-step phase times of the peer-memory transport since timing(True), in ms: prefix (hash,
        migration, sort), interior forces, halo + boundary forces (side stream), whole step."""
        out = (C.c_float * 4)()
        k = C.c_int()
        _lib.check(self.lib.gof_slab_phase_times(self._handle, out, C.byref(k)))
        steps = max(int(k.value), 1)
        return {"prefix": out[0] / steps, "interior_forces": out[1] / steps, "halo_and_boundary_forces": out[2] / steps,
                "step": out[3] / steps, "steps": int(k.value)}

    def pair_stats(self) -> dict:
        n = self.count()  # (also refreshes the library's host-side count from the device words)
        out = (C.c_ulonglong * 3)()
        _lib.check(self.lib.gof_engine_pair_stats(self._handle, self.wrap_mode, out, self._stream()))
        return {"n": max(n, 1), "candidates": int(out[0]),
                "evaluated": int(out[1]), "in_range": int(out[2])}

    def set_parameters(self, parameter_matrix):
        pm = np.ascontiguousarray(parameter_matrix, dtype=np.float64)
        _lib.check(self.lib.gof_engine_set_parameters(self._handle, pm.ctypes.data, self._stream()))

    def timing(self, enable: bool):
        _lib.check(self.lib.gof_engine_timing(self._handle, 1 if enable else 0))

    def force_time(self):
        ms, k = C.c_float(), C.c_int()
        _lib.check(self.lib.gof_engine_force_time(self._handle, C.byref(ms), C.byref(k)))
        return float(ms.value), int(k.value)

    # ---- phases ------------------------------------------------------------------------------
    # Phases 1 and 2 only enqueue work.  Migration buffers travel whole (fixed size, the record
    # count is their header word) so no size has to reach the host before the sort; the only
    # words the host reads per step are phase 2's, all-gathered from device memory.
    def phase_hash(self):
        _lib.check(self.lib.gof_slab_phase_hash_async(self._handle, self._stream()))

    def _mig_words(self) -> int:
        return 4 * (1 + 3 * self.migrate_capacity)

    def mig_send(self, direction: int) -> torch.Tensor:
        """The whole buffer (header + records) going to the slab below (0) / above (1)."""
        return self._view(self._ptr[direction], self._mig_words(), torch.float32)

    def mig_recv(self, source: int) -> torch.Tensor:
        """Where the buffer coming from the slab below (0) / above (1) is received."""
        return self._view(self._ptr[2 + source], self._mig_words(), torch.float32)

    def speed_word(self) -> torch.Tensor:
        """max |v|^2 of the resident particles as its int32 bit pattern (orders like the float)."""
        return self._view(self._ptr[7 + int(self.lib.gof_slab_parity(self._handle))], 1, torch.int32)

    def phase_sort(self, timestep=None):
        ts = -1.0 if timestep is None else float(timestep)
        _lib.check(self.lib.gof_slab_phase_sort_async(self._handle, ts, self._stream()))

    def sort_words(self) -> torch.Tensor:
        """int32[4] on the device after phase_sort: live, bottom layer, top layer population, error flag."""
        return self._view(self._ptr[10] + 16, 4, torch.int32)

    def commit(self, words):
        _lib.check(self.lib.gof_slab_commit(self._handle, int(words[0]), int(words[1]), int(words[2]), int(words[3])))
        self.n_live, self.lo_pop, self.hi_pop = int(words[0]), int(words[1]), int(words[2])

    def _rows(self, first: int, n: int):
        return (self._view(self._ptr[4] + 16 * first, 4 * n, torch.float32),
                self._view(self._ptr[5] + 16 * first, 4 * n, torch.float32))

    def _counts(self, layer: int) -> torch.Tensor:
        return self._view(self._ptr[6] + 4 * layer * self.per_layer, self.per_layer, torch.int32)

    def _layer_msgs(self, first: int, n: int, layer: int):
        pos, vel = self._rows(first, n)
        return [pos, vel, self._counts(layer)] if self.has_boids else [pos, self._counts(layer)]

    def halo_send(self, direction: int):
        """Boundary layer handed to the slab below (0) / above (1): cell-ordered positions (and
        velocities when some species pair is of the Boids kind) plus the per-cell counts."""
        if direction == 0:
            return self._layer_msgs(0, self.lo_pop, 1)
        return self._layer_msgs(self.n_live - self.hi_pop, self.hi_pop, self.nlz - 2)

    def halo_recv(self, source: int, n: int, n_before: int = 0):
        """Where the layer received from the slab below (0) / above (1) goes: behind the owned
        particles (the upper halo behind the lower one: pass its size as `n_before`)."""
        return self._layer_msgs(self.n_live + n_before, n, 0 if source == 0 else self.nlz - 1)

    def phase_force(self, n_halo_lo: int, n_halo_hi: int):
        """Everything in one go (no overlap)."""
        _lib.check(self.lib.gof_slab_phase_force(self._handle, int(n_halo_lo), int(n_halo_hi), self.wrap_mode,
                                                 self._stream()))

    def phase_force_interior(self):
        """The owned layers that touch no halo; enqueued before the sizes are known on the host."""
        _lib.check(self.lib.gof_slab_phase_force_interior(self._handle, self.wrap_mode, self._stream()))

    def phase_force_boundary(self, n_halo_lo: int, n_halo_hi: int):
        _lib.check(self.lib.gof_slab_phase_force_boundary(self._handle, int(n_halo_lo), int(n_halo_hi),
                                                          self.wrap_mode, self._stream()))

    def side_stream(self):
        """A second stream for the halo traffic that runs under the interior forces."""
        if getattr(self, "_side", None) is None:
            # HIGH priority: its few blocks (size gather, halo copies, boundary layers) must be
            # scheduled ahead of the thousands of pending blocks of the interior kernel, otherwise
            # they only start when that kernel drains and nothing overlaps
            self._side = torch.cuda.Stream(device=self.index, priority=-1)
        return self._side


@dataclass
class Transfer:
    src: int                 # global slab index of the sender
    dst: int                 # global slab index of the receiver
    send: object = None      # callable -> list of tensors (evaluated only where src is local)
    recv: object = None      # callable -> list of tensors (evaluated only where dst is local)


class LocalComm:
    """All slabs live in this process: transfers are device copies."""

    def __init__(self, world_slabs: int):
        self.world_slabs = world_slabs

    def owner(self, slab: int) -> int:
        return 0

    def rank(self) -> int:
        return 0

    def gather_words(self, local: dict) -> dict:
        """{slab: int32[4] tensor} -> {slab: [4 ints]}; one synchronisation for all of them."""
        keys = sorted(local)
        vals = torch.stack([local[k] for k in keys]).cpu().tolist()
        return dict(zip(keys, vals))

    def allreduce_max(self, words):
        if len(words) > 1:
            m = torch.stack([w.reshape(()) for w in words]).max()
            for w in words:
                w.fill_(m)

    def allreduce_max_begin(self, words):
        self.allreduce_max(words)
        return lambda: None

    def exchange(self, transfers):
        for t in transfers:
            for s, d in zip(t.send(), t.recv()):
                if d.numel():
                    d.copy_(s, non_blocking=True)


def nccl_options():
    """Process-group options for the slab runner: NCCL's own stream at high priority, so that
    the small halo / size messages are not queued behind the interior force kernel's blocks."""
    import torch.distributed as dist
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = True
    return opts


class DistComm:
    """One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests).
    Slab s belongs to rank s * world / world_slabs (normally one slab per rank).  Create the
    NCCL group with `pg_options=nccl_options()`."""

    def __init__(self, world_slabs: int, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.world_slabs = world_slabs
        self.world = dist.get_world_size()
        self._rank = dist.get_rank()
        self.device = device
        assert world_slabs % self.world == 0
        self.per_rank = world_slabs // self.world

    def owner(self, slab: int) -> int:
        return slab // self.per_rank

    def rank(self) -> int:
        return self._rank

    def gather_words(self, local: dict) -> dict:
        """{slab: int32[4] tensor} of this rank's slabs -> {slab: [4 ints]} for every slab.
        The words are all-gathered from where the kernels left them (device memory under
        NCCL), then read back once."""
        base = self._rank * self.per_rank
        mine = torch.stack([local[base + k] for k in range(self.per_rank)]).reshape(-1).contiguous()
        out = torch.empty(self.world * mine.numel(), dtype=mine.dtype, device=mine.device)
        if self.dist.get_backend() == "gloo":
            parts = list(out.chunk(self.world))
            self.dist.all_gather(parts, mine)
        else:
            self.dist.all_gather_into_tensor(out, mine)
        vals = out.cpu().reshape(self.world_slabs, 4).tolist()
        return {s: vals[s] for s in range(self.world_slabs)}

    def allreduce_max(self, words):
        self.allreduce_max_begin(words)()

    def allreduce_max_begin(self, words):
        """Start the MAX all-reduce of the slabs' speed words; the returned callable makes the
        current stream wait for it.  The words are final once the previous step's forces are
        enqueued, so the reduction can run under this step's hash pass."""
        if len(words) > 1:
            m = torch.stack([w.reshape(()) for w in words]).max()
            for w in words:
                w.fill_(m)
        work = self.dist.all_reduce(words[0], op=self.dist.ReduceOp.MAX, async_op=True)

        def finish():
            work.wait()
            for w in words[1:]:
                w.copy_(words[0])
        return finish

    def exchange(self, transfers):
        """Every rank walks the same global list; NCCL matches the send/recv of a pair of
        ranks in posting order, so the order of the list is the protocol."""
        ops = []
        for t in transfers:
            so, do = self.owner(t.src), self.owner(t.dst)
            if so == self._rank and do == self._rank:
                for s, d in zip(t.send(), t.recv()):
                    if d.numel():
                        d.copy_(s, non_blocking=True)
            elif so == self._rank:
                ops += [self.dist.P2POp(self.dist.isend, s, do) for s in t.send() if s.numel()]
            elif do == self._rank:
                ops += [self.dist.P2POp(self.dist.irecv, d, so) for d in t.recv() if d.numel()]
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()


class SlabRunner:
    """Drives the three phases of a step over this process's slabs."""

    def __init__(self, slabs: dict, world_slabs: int, comm, overlap: bool = True):
        self.overlap = overlap            # run the halo traffic under the interior forces
        self.slabs = dict(slabs)          # {global slab index: engine}
        self.world_slabs = int(world_slabs)
        self.comm = comm
        self.messages = 0                 # tensors handed to the transport so far (diagnostics)
        self.transport = "copies / send-recv sequenced from Python, sizes gathered per step"
        # Capacities are part of the protocol: every rank must be able to check EVERY slab's words
        # (a sender can pass its own check with a layer its receiver cannot hold), so they are
        # gathered once here: {slab: [capacity, halo_capacity, migrate_capacity, 0]}.
        dev = next(iter(self.slabs.values())).device if self.slabs and hasattr(next(iter(self.slabs.values())), "device") else None
        mk = lambda e: torch.tensor([int(getattr(e, "capacity", 1 << 30)), int(getattr(e, "halo_capacity", 1 << 30)),
                                     int(getattr(e, "migrate_capacity", 0)), 0], dtype=torch.int32,
                                    device=dev if dev is not None else "cpu")
        self.caps = self.comm.gather_words({s: mk(e) for s, e in self.slabs.items()})
        migs = {c[2] for c in self.caps.values()}
        if len(migs) > 1:
            raise ValueError(f"SlabRunner: migrate_capacity must be the same on every slab, got {sorted(migs)}")
        # GOF_SLAB_TIMING=1: device time of the serial part of a step (hash + migration + sort) and
        # of the forces (with the halo traffic underneath), summed over the steps, in `phase_ms`
        self.phase_ms = {"prefix": 0.0, "forces": 0.0, "steps": 0} if os.environ.get("GOF_SLAB_TIMING") else None

    def _below(self, s):
        return (s - 1) % self.world_slabs

    def _above(self, s):
        return (s + 1) % self.world_slabs

    def step(self, n_steps: int = 1, timestep=None):
        W = self.world_slabs
        for _ in range(n_steps):
            marks = None
            if self.phase_ms is not None:
                e0 = next(iter(self.slabs.values()))
                ms = torch.cuda.current_stream(e0.index)
                marks = [ms.record_event(torch.cuda.Event(enable_timing=True))]
            # ---- 0. the global maximum speed of the state as it stands (adaptive timestep): the
            # reduction runs on the transport's stream under the hash pass
            finish_speed = None
            if timestep is None:
                finish_speed = self.comm.allreduce_max_begin([e.speed_word() for e in self.slabs.values()])
            # ---- 1. bin, pack leavers; the fixed-size migration buffers go to the neighbours ---
            for e in self.slabs.values():
                e.phase_hash()
            transfers = []
            for s in range(W):  # downward buffers first, then upward: one global order
                transfers.append(Transfer(s, self._below(s),
                                          send=lambda s=s: [self.slabs[s].mig_send(0)],
                                          recv=lambda s=s: [self.slabs[self._below(s)].mig_recv(1)]))
            for s in range(W):
                transfers.append(Transfer(s, self._above(s),
                                          send=lambda s=s: [self.slabs[s].mig_send(1)],
                                          recv=lambda s=s: [self.slabs[self._above(s)].mig_recv(0)]))
            self.comm.exchange(transfers)
            if finish_speed is not None:
                finish_speed()
            # ---- 2. append arrivals (counts are read on the device), sort ---------------------
            for e in self.slabs.values():
                e.phase_sort(timestep)
            if marks is not None:
                marks.append(ms.record_event(torch.cuda.Event(enable_timing=True)))
            # ---- 3a. interior forces start now; sizes + halo layers move underneath ------------
            overlap = self.overlap and all(hasattr(e, "side_stream") for e in self.slabs.values())
            if overlap:
                main = {s: torch.cuda.current_stream(e.index) for s, e in self.slabs.items()}
                sorted_ev = {s: main[s].record_event() for s in self.slabs}
                for e in self.slabs.values():
                    e.phase_force_interior()
                first = next(iter(self.slabs.values()))
                # one side stream: the local slabs of a runner share a device (asserted below)
                assert all(e.index == first.index for e in self.slabs.values()), \
                    "SlabRunner(overlap=True): the slabs of one process must live on one device"
                side = first.side_stream()
                for s in self.slabs:
                    side.wait_event(sorted_ev[s])
                ctx = torch.cuda.stream(side)
            else:
                ctx = contextlib.nullcontext()
            with ctx:
                pops = self.comm.gather_words({s: e.sort_words() for s, e in self.slabs.items()})
                # Every rank looks at EVERY slab's words and raises the same error, before any halo
                # byte moves: a failure seen by one rank only would leave its peers in send/recv.
                self._check_words(pops)
                for s, e in self.slabs.items():
                    e.commit(pops[s])
                transfers = []
                for s in range(W):  # every slab's bottom layer becomes the upper halo of the slab below
                    n = pops[s][1]
                    d = self._below(s)
                    # the receiver stores the upper halo behind its lower halo (= top layer of ITS lower slab)
                    n_before = pops[self._below(d)][2]
                    transfers.append(Transfer(s, d,
                                              send=lambda s=s: self.slabs[s].halo_send(0),
                                              recv=lambda d=d, n=n, nb=n_before: self.slabs[d].halo_recv(1, n, nb)))
                for s in range(W):  # every slab's top layer becomes the lower halo of the slab above
                    n = pops[s][2]
                    d = self._above(s)
                    transfers.append(Transfer(s, d,
                                              send=lambda s=s: self.slabs[s].halo_send(1),
                                              recv=lambda d=d, n=n: self.slabs[d].halo_recv(0, n, 0)))
                self.comm.exchange(transfers)
                # ---- 3b. the boundary layers once their halo is in: on the side stream too, so
                # that their (few) blocks fill in next to the interior kernel instead of after it
                for s, e in self.slabs.items():
                    if overlap:
                        e.phase_force_boundary(pops[self._below(s)][2], pops[self._above(s)][1])
                    else:
                        e.phase_force(pops[self._below(s)][2], pops[self._above(s)][1])
                if overlap:
                    done_ev = side.record_event()
            if overlap:
                for s in self.slabs:
                    main[s].wait_event(done_ev)
            self.messages += 4 * W
            if marks is not None:
                marks.append(ms.record_event(torch.cuda.Event(enable_timing=True)))
                marks[2].synchronize()
                self.phase_ms["prefix"] += marks[0].elapsed_time(marks[1])
                self.phase_ms["forces"] += marks[1].elapsed_time(marks[2])
                self.phase_ms["steps"] += 1

    _ERRORS = {1: "a particle moved more than one layer in a step, or arrived in the wrong slab",
               2: "migration buffer overflow (raise migrate_capacity)",
               5: "particle capacity exceeded"}

    def _check_words(self, pops):
        """pops: {slab: [live, bottom layer, top layer, error]} of ALL slabs (identical on every rank)."""
        for s in range(self.world_slabs):
            live, lo, hi, err = pops[s]
            if err:
                raise _lib.GofError(_lib.E_STATE if err == 1 else _lib.E_NOMEM,
                                    f"slab {s}: {self._ERRORS.get(err, f'device-side error {err}')}")
            if live > self.caps[s][0]:
                raise _lib.GofError(_lib.E_NOMEM, f"slab {s}: particle capacity exceeded ({live} > {self.caps[s][0]})")
            below, above = self._below(s), self._above(s)
            if lo > self.caps[below][1]:
                raise _lib.GofError(_lib.E_NOMEM, f"slab {s}: bottom layer of {lo} particles does not fit slab {below}'s "
                                                  f"halo_capacity {self.caps[below][1]}")
            if hi > self.caps[above][1]:
                raise _lib.GofError(_lib.E_NOMEM, f"slab {s}: top layer of {hi} particles does not fit slab {above}'s "
                                                  f"halo_capacity {self.caps[above][1]}")

    def close(self):
        pass

    def gather_state(self, num_particles: int):
        """This process's particles scattered into (3, N) arrays by id (others stay NaN)."""
        out = {k: np.full((3, num_particles), np.nan, dtype=np.float32) for k in ("pos", "vel", "acc")}
        owner = np.full(num_particles, -1, dtype=np.int32)
        for s, e in self.slabs.items():
            d = e.download()
            for k in ("pos", "vel", "acc"):
                out[k][:, d["ids"]] = d[k]
            owner[d["ids"]] = s
        out["owner"] = owner
        return out


class P2PUnavailable(RuntimeError):
    """The peer-memory transport cannot be set up on this job (raised on EVERY rank alike)."""


class P2PRunner:
    """The slabs of a job exchange through peer memory (gof_slab_connect / gof_slab_step_p2p).

    `slabs`: {global slab index: SlabEngine} of this process.  With torch.distributed initialised
    (one slab per rank is the normal case) the arenas are shared with CUDA IPC: every rank exports
    a handle, all-gathers handles and window descriptions once, and maps its peers.  Without it all
    slabs must be local (tests on one GPU).  A step is ONE C call; nothing is read back.

    Set-up failures (no IPC between the ranks' devices, a driver that refuses the mapping) are
    agreed on collectively after every stage, so that all ranks raise P2PUnavailable together and a
    caller can fall back to the send/recv transport (make_runner does)."""

    transport = "p2p (peer-memory stores over NVLink, flag waits on the stream, no host sync)"

    def __init__(self, slabs: dict, world_slabs: int, group_ready: bool | None = None):
        self.slabs = dict(slabs)
        self.world_slabs = int(world_slabs)
        self.lib = _lib.load()
        self.phase_ms = None
        self._opened = []
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized() if group_ready is None else group_ready
        self._distributed = distributed
        first = next(iter(self.slabs.values()))
        self.index = first.index
        if any(e.index != self.index for e in self.slabs.values()):
            raise ValueError("P2PRunner: the local slabs must share one device")
        W = self.world_slabs
        if len(self.slabs) < W and not distributed:
            raise ValueError("P2PRunner: slabs missing and torch.distributed is not initialised")

        def agree(error):
            """All ranks learn whether ANY of them failed the stage; everybody raises then."""
            if not distributed:
                if error is not None:
                    raise P2PUnavailable(str(error))
                return
            flag = torch.tensor([0 if error is None else 1], dtype=torch.int32, device=first.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX)
            if int(flag.item()):
                self._unmap()
                raise P2PUnavailable(str(error) if error is not None else "another rank could not set it up")

        # ---- stage 1: describe and export the local arenas
        infos, bases, mine, err = {}, {}, {}, None
        try:
            for s, e in self.slabs.items():
                w = (C.c_int64 * 12)()
                _lib.check(self.lib.gof_slab_window(e._handle, w, 12))
                infos[s] = list(w)
                bases[s] = e._arena_ptr
                if len(self.slabs) < W:
                    h = (C.c_ubyte * 64)()
                    off = C.c_int64()
                    with torch.cuda.device(self.index):
                        _lib.check(self.lib.gof_ipc_export(e._arena_ptr, h, C.byref(off)))
                    mine[s] = (bytes(h), int(off.value), infos[s])
        except Exception as exc:  # noqa: BLE001
            err = exc
        agree(err)
        # ---- stage 2: map the peers
        err = None
        if len(self.slabs) < W:
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, mine)
            try:
                for part in gathered:
                    for s, (h, off, info) in part.items():
                        infos[s] = info
                        if s in self.slabs:
                            continue
                        base = C.c_void_p()
                        buf = (C.c_ubyte * 64).from_buffer_copy(h)
                        with torch.cuda.device(self.index):
                            _lib.check(self.lib.gof_ipc_open(buf, C.byref(base)))
                        self._opened.append(base.value)
                        bases[s] = base.value + off
                if sorted(bases) != list(range(W)):
                    raise ValueError(f"need every slab 0..{W - 1}, have {sorted(bases)}")
            except Exception as exc:  # noqa: BLE001
                err = exc
            agree(err)
        # ---- stage 3: connect (zeroes this slab's flag block), then nobody stores before everybody is through
        err = None
        try:
            base_arr = (C.c_void_p * W)(*[bases[s] for s in range(W)])
            info_arr = (C.c_int64 * (12 * W))(*[x for s in range(W) for x in infos[s]])
            with torch.cuda.device(self.index):
                for s, e in self.slabs.items():
                    _lib.check(self.lib.gof_slab_connect(e._handle, s, W, base_arr, info_arr))
                torch.cuda.synchronize(self.index)
        except Exception as exc:  # noqa: BLE001
            err = exc
        agree(err)
        if distributed:
            dist.barrier()
        self._handles = (C.c_void_p * len(self.slabs))(*[self.slabs[s]._handle for s in sorted(self.slabs)])
        self._side = torch.cuda.Stream(device=self.index, priority=-1)

    def step(self, n_steps: int = 1, timestep=None):
        first = next(iter(self.slabs.values()))
        ts = -1.0 if timestep is None else float(timestep)
        with torch.cuda.device(self.index):
            _lib.check(self.lib.gof_slab_step_p2p(self._handles, len(self.slabs), int(n_steps), ts, first.wrap_mode,
                                                  torch.cuda.current_stream(self.index).cuda_stream,
                                                  self._side.cuda_stream))

    def gather_state(self, num_particles: int):
        return SlabRunner.gather_state(self, num_particles)

    def _unmap(self):
        for b in self._opened:
            self.lib.gof_ipc_close(b)
        self._opened = []

    def close(self):
        """Unmap the peers (every rank must have finished stepping: barrier first)."""
        if self._opened:
            torch.cuda.synchronize(self.index)
            if self._distributed:
                import torch.distributed as dist
                dist.barrier()
            self._unmap()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_runner(eng, world: int, device=None, transport: str = "auto"):
    """The runner of a one-slab-per-rank job (torch.distributed initialised): the peer-memory
    transport; `transport == "nccl"` (or GOF_SLAB_TRANSPORT=nccl) selects NCCL send/recv, and
    "auto" falls back to it when peer memory cannot be set up (all ranks decide together)."""
    import torch.distributed as dist
    choice = os.environ.get("GOF_SLAB_TRANSPORT", transport)
    rank = dist.get_rank()
    if choice in ("auto", "p2p"):
        try:
            return P2PRunner({rank: eng}, world)
        except P2PUnavailable as exc:
            if choice == "p2p":
                raise
            if rank == 0:
                import warnings
                warnings.warn(f"peer-memory transport unavailable ({exc}): using NCCL send/recv", RuntimeWarning)
    runner = SlabRunner({rank: eng}, world, DistComm(world, device=device))
    runner.transport = "nccl (send/recv of halo layers and migration records, sizes all-gathered per step)"
    return runner


def assign_to_slabs(pos_z, grid_bin_size_z: float, num_layers: int, layers):
    """Slab index of every particle from its z bin (host helper for the initial distribution)."""
    cz = np.minimum(np.floor(np.asarray(pos_z, dtype=np.float64) / grid_bin_size_z).astype(np.int64),
                    num_layers - 1)
    slab = np.empty(cz.shape, dtype=np.int32)
    for s, (lo, hi) in enumerate(layers):
        slab[(cz >= lo) & (cz < hi)] = s
    return slab

"""A numpy stand-in for game_of_flies_b200.slab.SlabEngine (test infrastructure).

It implements the same phase interface in float64 on the CPU so that the host-side
slab logic (SlabRunner + DistComm: neighbour bookkeeping, message sizes and order,
migration, halo placement, the all-reduced adaptive timestep) can be exercised with
two gloo ranks and no GPU.  Forces are brute force over (owned + received halo)
particles with the minimum image convention, Clusters kind only: if a halo were
incomplete or misplaced the result would differ from the single-domain oracle.
"""
import numpy as np
import torch


class FakeSlabEngine:
    REC = 12
    CAP = 64  # records per migration buffer (fixed-size messages, count in the header)

    def __init__(self, parameter_matrix, limits, r_max, layer_lo, layer_hi, **_):
        self.pm = np.asarray(parameter_matrix, dtype=np.float64)
        self.L = np.asarray(limits, dtype=np.float64)
        self.r_max = float(r_max)
        self.nb = np.floor(self.L / self.r_max).astype(int)
        self.bs = self.L / self.nb
        self.lo, self.hi = int(layer_lo), int(layer_hi)
        self.nlz = self.hi - self.lo + 2
        self.per_layer = int(self.nb[0] * self.nb[1])
        self.pos = np.zeros((3, 0)); self.vel = np.zeros((3, 0)); self.acc = np.zeros((3, 0))
        self.types = np.zeros(0, dtype=np.int64); self.ids = np.zeros(0, dtype=np.int64)
        self._speed = torch.zeros(1, dtype=torch.float64)
        self._send = [None, None]
        self._recv = [None, None]
        self._halo = [None, None]
        self._words = torch.zeros(4, dtype=torch.int32)
        self.n_live = self.lo_pop = self.hi_pop = 0

    # ---- state ----
    def upload(self, pos, types, ids, vel=None, acc=None):
        n = len(types)
        self.pos = np.array(pos, dtype=np.float64).reshape(3, n)
        self.vel = np.zeros((3, n)) if vel is None else np.array(vel, dtype=np.float64).reshape(3, n)
        self.acc = np.zeros((3, n)) if acc is None else np.array(acc, dtype=np.float64).reshape(3, n)
        self.types = np.array(types, dtype=np.int64)
        self.ids = np.array(ids, dtype=np.int64)
        self._speed[0] = float((self.vel ** 2).sum(0).max(initial=0.0))

    def download(self):
        return dict(pos=self.pos.copy(), vel=self.vel.copy(), acc=self.acc.copy(), ids=self.ids.copy())

    # ---- helpers ----
    def _cells(self, pos):
        c = np.minimum(np.floor(pos / self.bs[:, None]).astype(int), self.nb[:, None] - 1)
        c = np.maximum(c, 0)
        lz = (c[2] - (self.lo - 1)) % self.nb[2]
        return c[0] + self.nb[0] * (c[1] + self.nb[1] * lz), lz

    def _records(self, mask):
        n = int(mask.sum())
        assert n <= self.CAP, "migration buffer overflow in the fake engine"
        buf = np.zeros(4 + self.CAP * 12)
        buf[0] = n
        rec = buf[4:4 + n * 12].reshape(n, 12)
        rec[:, 0:3] = self.pos[:, mask].T; rec[:, 3] = self.types[mask]
        rec[:, 4:7] = self.vel[:, mask].T; rec[:, 7] = self.ids[mask]
        rec[:, 8:11] = self.acc[:, mask].T
        return torch.from_numpy(buf)

    # ---- phases ----
    def phase_hash(self):
        _, lz = self._cells(self.pos)
        down, up = lz == 0, lz == self.nlz - 1
        assert not np.any((lz > self.nlz - 1)), "particle moved more than one layer"
        self._send = [self._records(down), self._records(up)]
        keep = ~(down | up)
        self.pos, self.vel, self.acc = self.pos[:, keep], self.vel[:, keep], self.acc[:, keep]
        self.types, self.ids = self.types[keep], self.ids[keep]

    def sort_words(self):
        return self._words

    def commit(self, words):
        assert words[3] == 0 and (self.n_live, self.lo_pop, self.hi_pop) == tuple(words[:3])

    def mig_send(self, direction):
        return self._send[direction]

    def mig_recv(self, source):
        self._recv[source] = torch.zeros(4 + self.CAP * 12, dtype=torch.float64)
        return self._recv[source]

    def speed_word(self):
        return self._speed

    def phase_sort(self, timestep=None):
        for src in (0, 1):
            buf = self._recv[src].numpy()
            n = int(buf[0])
            if n:
                rec = buf[4:4 + n * 12].reshape(n, 12)
                self.pos = np.concatenate([self.pos, rec[:, 0:3].T], axis=1)
                self.vel = np.concatenate([self.vel, rec[:, 4:7].T], axis=1)
                self.acc = np.concatenate([self.acc, rec[:, 8:11].T], axis=1)
                self.types = np.concatenate([self.types, rec[:, 3].astype(np.int64)])
                self.ids = np.concatenate([self.ids, rec[:, 7].astype(np.int64)])
        if timestep is None:
            t = np.sqrt(float(self._speed[0]))
            timestep = 0.4 / t if t > 1e-6 else 0.1
        self.dt = timestep
        # integrator.step1
        self.vel = self.vel + self.acc * 0.5 * timestep
        sq = (self.vel ** 2).sum(0)
        fast = sq > 400.0
        self.vel[:, fast] *= 20.0 / np.sqrt(sq[fast])
        cell, lz = self._cells(self.pos)
        assert np.all((lz >= 1) & (lz <= self.nlz - 2)), "a particle is not in an owned layer after migration"
        order = np.lexsort((self.ids, cell))
        self.pos, self.vel, self.acc = self.pos[:, order], self.vel[:, order], self.acc[:, order]
        self.types, self.ids = self.types[order], self.ids[order]
        self.cell, lz = cell[order], lz[order]
        self.counts = np.bincount(self.cell, minlength=self.per_layer * self.nlz)
        self.n_live = len(self.ids)
        self.lo_pop = int((lz == 1).sum())
        self.hi_pop = int((lz == self.nlz - 2).sum())
        self._words = torch.tensor([self.n_live, self.lo_pop, self.hi_pop, 0], dtype=torch.int32)

    def _layer(self, first, n, layer):
        pos = np.zeros((n, 4)); pos[:, :3] = self.pos[:, first:first + n].T; pos[:, 3] = self.types[first:first + n]
        vel = np.zeros((n, 4)); vel[:, :3] = self.vel[:, first:first + n].T; vel[:, 3] = self.ids[first:first + n]
        cnt = self.counts[layer * self.per_layer:(layer + 1) * self.per_layer].astype(np.int64)
        return [torch.from_numpy(pos.reshape(-1).copy()), torch.from_numpy(vel.reshape(-1).copy()),
                torch.from_numpy(cnt.copy())]

    def halo_send(self, direction):
        if direction == 0:
            return self._layer(0, self.lo_pop, 1)
        return self._layer(self.n_live - self.hi_pop, self.hi_pop, self.nlz - 2)

    def halo_recv(self, source, n, n_before=0):
        self._halo[source] = [torch.zeros(4 * n, dtype=torch.float64), torch.zeros(4 * n, dtype=torch.float64),
                              torch.zeros(self.per_layer, dtype=torch.int64)]
        return self._halo[source]

    def phase_force(self, n_halo_lo, n_halo_hi):
        hp = [self._halo[k][0].numpy().reshape(-1, 4) for k in (0, 1)]
        assert hp[0].shape[0] == n_halo_lo and hp[1].shape[0] == n_halo_hi
        assert int(self._halo[0][2].sum()) == n_halo_lo and int(self._halo[1][2].sum()) == n_halo_hi
        allpos = np.concatenate([self.pos.T, hp[0][:, :3], hp[1][:, :3]], axis=0)
        alltyp = np.concatenate([self.types, hp[0][:, 3].astype(np.int64), hp[1][:, 3].astype(np.int64)])
        acc = np.zeros_like(self.pos)
        pm = self.pm
        for i in range(self.n_live):
            d = self.pos[:, i][None, :] - allpos
            d -= self.L[None, :] * np.round(d / self.L[None, :])
            r2 = (d * d).sum(1)
            m = (r2 > 1e-10) & (r2 < self.r_max ** 2)
            r = np.sqrt(r2[m]); tj = alltyp[m]; ti = self.types[i]
            a0, a1, a2, a3 = (pm[k, ti, tj] for k in range(4))
            f = np.where(r < a0, -a2 / a0 * r + a2,
                         np.where(r < (a0 + a1) / 2, 2 * a3 / (a1 - a0) * (r - a0), 2 * a3 / (a1 - a0) * (a1 - r)))
            acc[:, i] = -(f / r * d[m].T).sum(1)
        dt = self.dt
        self.acc = acc
        self.pos = self.pos + (self.vel + 0.5 * acc * dt) * dt
        self.vel = (self.vel + acc * 0.5 * dt) * 0.95
        for k in range(3):
            L = self.L[k]
            self.pos[k] = np.where(self.pos[k] < 0, self.pos[k] + L, np.where(self.pos[k] > L, self.pos[k] - L, self.pos[k]))
        self._speed[0] = float((self.vel ** 2).sum(0).max(initial=0.0))

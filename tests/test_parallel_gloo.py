"""CPU tests of the multi-GPU host logic (world_size 2 and 3, gloo backend):
slab partition, migration + ghost import message framing, dt all-reduce.
The device ops are replaced by a numpy test double with the same interface as
pysph_b200.parallel.DeviceHaloOps; the GPU version of the same check lives in
tests/test_gpu_multi.py."""
import os
import socket

import numpy as np
import pytest

from pysph_b200 import geometry as geo
from pysph_b200.parallel import (HALO_FIELDS, MIGRATE_FIELDS,
                                 SlabParallelManager, balanced_cuts,
                                 dam_break_column_weights)

F64 = ['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'h', 'm',
       'x0', 'y0', 'z0', 'u0', 'v0', 'w0', 'rho0']


class NumpyHaloOps(object):
    """Same protocol as DeviceHaloOps, on host numpy arrays / CPU tensors."""

    def __init__(self, arrays):
        import torch
        self.torch = torch
        self.arrays = arrays          # list of dict field -> np.ndarray (+ gid)
        self.nreal = [a['x'].size for a in arrays]
        self.narr = len(arrays)

    def new_buffer(self, n):
        return self.torch.zeros(max(int(n), 1), dtype=self.torch.float64)

    def new_counts(self, values=None, n=0):
        t = self.torch
        return t.tensor(values, dtype=t.int64) if values is not None \
            else t.zeros(n, dtype=t.int64)

    def n_real(self, a):
        return self.nreal[a]

    def drop_ghosts(self, a):
        for k in self.arrays[a]:
            self.arrays[a][k] = self.arrays[a][k][:self.nreal[a]]

    def pack(self, a, slot, lo, hi, buf, off):
        A = self.arrays[a]
        x = A['x'][:self.nreal[a]]
        sel = np.where((x >= lo) & (x < hi))[0]
        self.saved = getattr(self, 'saved', {})
        self.saved[(a, slot)] = sel
        return self._pack_idx(a, sel, buf, off)

    def _pack_idx(self, a, sel, buf, off):
        A = self.arrays[a]
        n = sel.size
        for f, name in enumerate(F64[:HALO_FIELDS]):
            buf[off + f * n: off + (f + 1) * n] = self.torch.from_numpy(A[name][sel])
        return n

    def pack_selected(self, a, slot, buf, off):
        return self._pack_idx(a, self.saved[(a, slot)], buf, off)

    def overwrite(self, a, ghost_first, buf, off, n):
        if not n:
            return
        A = self.arrays[a]
        blk = buf[off:off + n * HALO_FIELDS].numpy().reshape(HALO_FIELDS, n)
        lo = self.nreal[a] + ghost_first
        for f, name in enumerate(F64[:HALO_FIELDS]):
            A[name][lo:lo + n] = blk[f]

    def drift(self):
        return getattr(self, 'fake_drift', (0.0, 1.0))

    def read_later(self, t):          # deferred protocol: the host reads it later
        return lambda: float(t[0])

    def keep_build(self, strict=True):
        self.kept = getattr(self, 'kept', 0) + 1
        return True

    def append(self, a, buf, off, n, nfields, as_real):
        if not n:
            return
        A = self.arrays[a]
        blk = buf[off:off + n * nfields].numpy().reshape(nfields, n)
        for name in F64 + ['gid']:
            if name in F64[:HALO_FIELDS] or nfields == MIGRATE_FIELDS:
                idx = (F64 + ['gid']).index(name)
                new = blk[idx]
            else:
                new = np.full(n, -1.0) if name == 'gid' else np.zeros(n)
            A[name] = np.concatenate([A[name], new])
        if as_real:
            assert A['x'].size - n == self.nreal[a]
            self.nreal[a] += n

    def column_weights(self, x0, width, nbins, weights):
        out = np.zeros(nbins)
        for a, A in enumerate(self.arrays):
            b = np.clip(np.floor((A['x'][:self.nreal[a]] - x0) / width).astype(int),
                        0, nbins - 1)
            out += weights[a] * np.bincount(b, minlength=nbins)
        return self.torch.from_numpy(out)

    def migrate_out(self, a, lo, hi, buf, off):
        A = self.arrays[a]
        x = A['x']
        out = []
        for sel in (np.where(x < lo)[0], np.where(x >= hi)[0]):
            n = sel.size
            for f, name in enumerate(F64 + ['gid']):
                buf[off + f * n: off + (f + 1) * n] = \
                    self.torch.from_numpy(A[name][sel].astype(float))
            off += n * MIGRATE_FIELDS
            out.append(n)
        keep = (x >= lo) & (x < hi)
        for k in A:
            A[k] = A[k][keep]
        self.nreal[a] = int(keep.sum())
        return out[0], out[1]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _global_particles(seed=0):
    rs = np.random.RandomState(seed)
    arrays = []
    for n in (400, 150):
        a = dict((k, rs.uniform(0, 1, n)) for k in F64)
        a['x'] = rs.uniform(0, 3.0, n)
        a['gid'] = np.arange(n, dtype=float) + 1000 * len(arrays)
        arrays.append(a)
    return arrays


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cuts = [-np.inf] + [3.0 * k / world for k in range(1, world)] + [np.inf]
        halo = 0.2
        glob = _global_particles()
        # start from a WRONG ownership (round robin) so migration has work to do:
        # every particle must first travel to its owner through neighbours only,
        # so give each rank the particles of its own and adjacent slabs
        mine = []
        for a in glob:
            owner = np.searchsorted(cuts, a['x'], side='right') - 1
            sel = (np.abs(owner - rank) <= 1) & ((np.arange(a['x'].size) % 2 == rank % 2) |
                                                 (owner == rank))
            # keep exactly one copy globally: particle i goes to owner, except
            # odd-indexed ones that start on the left/right neighbour
            start = np.where(np.arange(a['x'].size) % 3 == 0,
                             np.clip(owner + 1, 0, world - 1), owner)
            sel = start == rank
            mine.append(dict((k, v[sel].copy()) for k, v in a.items()))
        ops = NumpyHaloOps(mine)
        pm = SlabParallelManager(ops, rank, world, cuts, halo, dist=dist)
        pm.update()
        res = []
        for ai, a in enumerate(glob):
            owner = np.searchsorted(cuts, a['x'], side='right') - 1
            own_gids = np.sort(a['gid'][owner == rank])
            got = ops.arrays[ai]
            nr = ops.nreal[ai]
            ok_real = np.array_equal(np.sort(got['gid'][:nr]), own_gids)
            # every property travelled with its particle
            order = np.argsort(got['gid'][:nr])
            src = np.argsort(a['gid'])
            src = src[np.isin(a['gid'][src], own_gids)]
            ok_props = all(np.array_equal(got[k][:nr][order], a[k][src]) for k in F64)
            # ghosts: neighbours' particles within halo of my cut planes
            lo, hi = cuts[rank], cuts[rank + 1]
            want = a['x'][((a['x'] >= lo - halo) & (a['x'] < lo)) |
                          ((a['x'] >= hi) & (a['x'] < hi + halo))]
            ok_ghost = np.array_equal(np.sort(got['x'][nr:]), np.sort(want))
            # ghost payload: rho of each ghost matches the global particle
            gx = got['x'][nr:]
            ok_payload = True
            for xg, rg in zip(gx, got['rho'][nr:]):
                j = np.where(a['x'] == xg)[0][0]
                ok_payload &= (a['rho'][j] == rg)
            res.append((ok_real, ok_props, ok_ghost, bool(ok_payload)))
        # a second update while the build is valid only REFRESHES the ghost values
        n_before = [ops.arrays[i]['x'].size for i in range(2)]
        for ai in range(2):
            nr = ops.nreal[ai]
            ops.arrays[ai]['rho'][:nr] += 1.0 + rank      # owners move on
        pm.update()
        idem = n_before == [ops.arrays[i]['x'].size for i in range(2)]
        idem &= pm.n_refresh == 1 and pm.n_full == 1
        for ai, a in enumerate(glob):
            got = ops.arrays[ai]
            nr = ops.nreal[ai]
            owner = np.searchsorted(cuts, a['x'], side='right') - 1
            for xg, rg in zip(got['x'][nr:], got['rho'][nr:]):
                j = np.where(a['x'] == xg)[0][0]
                idem &= bool(rg == a['rho'][j] + 1.0 + owner[j])
        # drift beyond the skin on ONE rank forces the full path on all ranks
        ops.fake_drift = (2.0, 1.0) if rank == 0 else (0.1, 1.0)
        pm.update()
        idem &= pm.n_full == 2 and n_before == [ops.arrays[i]['x'].size for i in range(2)]
        # deferred protocol (what the integrator uses): the refresh is applied on
        # the assumption "valid"; confirm() reads the all-reduced answer afterwards
        ops.fake_drift = (0.1, 1.0)
        for ai in range(2):
            ops.arrays[ai]['rho'][:ops.nreal[ai]] += 2.0
        nf, nr_ = pm.n_full, pm.n_refresh
        pm.update(deferred=True)
        idem &= pm._pending is not None and pm.confirm() is False
        idem &= (pm.n_full, pm.n_refresh) == (nf, nr_ + 1)
        for ai, a in enumerate(glob):
            got = ops.arrays[ai]
            nr = ops.nreal[ai]
            owner = np.searchsorted(cuts, a['x'], side='right') - 1
            for xg, rg in zip(got['x'][nr:], got['rho'][nr:]):
                j = np.where(a['x'] == xg)[0][0]
                idem &= bool(rg == a['rho'][j] + 3.0 + owner[j])
        # ... and a "no" from the LAST rank makes every rank run the full path in confirm()
        ops.fake_drift = (2.0, 1.0) if rank == world - 1 else (0.1, 1.0)
        pm.update(deferred=True)
        idem &= pm.n_full == nf                     # nothing happened yet
        idem &= pm.confirm() is True
        idem &= (pm.n_full, pm.n_refresh, pm.n_deferred_failed) == (nf + 1, nr_ + 1, 1)
        idem &= n_before == [ops.arrays[i]['x'].size for i in range(2)]
        idem &= pm.confirm() is False               # nothing pending any more
        dtmin = pm.update_time_steps(0.1 * (rank + 1))
        q.put((rank, res, idem, dtmin))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_slab_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res, idem, dtmin in out:
        assert all(all(r) for r in res), (rank, res)
        assert idem
        assert abs(dtmin - 0.1) < 1e-15


def _recut_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # equal-width slabs over [0, 3) but the particles crowd towards x = 0
        cuts = [-np.inf] + [3.0 * k / world for k in range(1, world)] + [np.inf]
        halo = 0.1
        rs = np.random.RandomState(5)
        glob = []
        for n in (3000, 900):
            a = dict((k, rs.uniform(0, 1, n)) for k in F64)
            a['x'] = 3.0 * rs.uniform(0, 1, n) ** 2.0
            a['gid'] = np.arange(n, dtype=float) + 10000 * len(glob)
            glob.append(a)
        weights = [1.0, 0.5]
        mine = []
        for a in glob:
            owner = np.searchsorted(cuts, a['x'], side='right') - 1
            mine.append(dict((k, v[owner == rank].copy()) for k, v in a.items()))
        ops = NumpyHaloOps(mine)
        pm = SlabParallelManager(ops, rank, world, cuts, halo, dist=dist, lb_freq=3,
                                 lb_columns=(0.0, 0.05, 60), lb_weights=weights)
        history = []
        ok = True

        def load():
            return sum(wt * ops.nreal[i] for i, wt in enumerate(weights))

        def check(cuts_now):
            good = True
            for ai, a in enumerate(glob):
                owner = np.searchsorted(cuts_now, a['x'], side='right') - 1
                nr = ops.nreal[ai]
                got = ops.arrays[ai]
                good &= np.array_equal(np.sort(got['gid'][:nr]), np.sort(a['gid'][owner == rank]))
                order = np.argsort(got['gid'][:nr])
                src = np.argsort(a['gid'])
                src = src[owner[src] == rank]
                good &= all(np.array_equal(got[k][:nr][order], a[k][src]) for k in F64)
                lo, hi = cuts_now[rank], cuts_now[rank + 1]
                want = a['x'][((a['x'] >= lo - halo) & (a['x'] < lo)) |
                              ((a['x'] >= hi) & (a['x'] < hi + halo))]
                good &= np.array_equal(np.sort(got['x'][nr:]), np.sort(want))
            return bool(good)

        pm.update()                                  # first build: the given planes
        ok &= pm.n_recut == 0 and check(cuts)
        history.append(load())
        ops.fake_drift = (2.0, 1.0)                  # every update takes the full path
        for it in range(12):
            before = pm.n_recut
            old = list(pm.cuts)
            pm.update()
            # a re-cut happens on a full update once lb_freq evaluations have passed
            ok &= (pm.n_recut > before) <= (it % 3 == 1)
            # planes stay on column boundaries, in order, clear of the old neighbours
            for k in range(1, world):
                ok &= abs(pm.cuts[k] / 0.05 - round(pm.cuts[k] / 0.05)) < 1e-9
                ok &= pm.cuts[k] - pm.cuts[k - 1] >= 2 * halo
                if k > 1:
                    ok &= pm.cuts[k] >= old[k - 1] + halo - 1e-12
                if k < world - 1:
                    ok &= pm.cuts[k] <= old[k + 1] - halo + 1e-12
            ok &= check(pm.cuts)
            history.append(load())
        q.put((rank, bool(ok), history, pm.cuts, pm.n_recut))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4])
def test_slab_recut_gloo(world):
    """lb_freq re-cut (parallel_manager.pyx:512-530): the planes move towards equal
    weighted counts through neighbour-only migration, nothing is lost or duplicated."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_recut_worker, args=(r, world, port, q))
             for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] for o in out), [o[1] for o in out]
    assert all(o[3] == out[0][3] for o in out)              # same planes everywhere
    assert out[0][4] >= (2 if world > 2 else 1)
    first = [o[2][0] for o in out]
    last = [o[2][-1] for o in out]
    assert abs(sum(first) - sum(last)) < 1e-9               # weighted total conserved
    imb0 = max(first) / (sum(first) / world)
    imb1 = max(last) / (sum(last) / world)
    assert imb0 > 1.4 and imb1 < 1.15, (imb0, imb1)


def test_partition_matches_geometry():
    dx = 0.05
    xs, w = dam_break_column_weights(dx, solid_weight=1.0)
    pas = geo.dam_break_3d_particles(dx=dx)
    total = sum(pa.get_number_of_particles() for pa in pas)
    assert abs(w.sum() - total) < 1e-9
    # per-column counts agree with the generated lattice
    allx = np.concatenate([pa.x for pa in pas])
    for i in (0, 1, 5, 20, len(xs) - 1):
        assert abs(w[i] - np.count_nonzero(np.abs(allx - xs[i]) < 1e-9)) < 1e-9
    for nparts in (2, 4, 8):
        xs, w = dam_break_column_weights(dx, solid_weight=0.3)
        cuts = balanced_cuts(xs, w, nparts, dx)
        assert len(cuts) == nparts + 1 and cuts[0] == -np.inf and cuts[-1] == np.inf
        assert all(cuts[k] < cuts[k + 1] for k in range(nparts))
        loads = [w[(xs >= cuts[k]) & (xs < cuts[k + 1])].sum() for k in range(nparts)]
        assert max(loads) <= 1.35 * (sum(loads) / nparts)
        # slab-restricted generation reproduces the global particle set
        n = 0
        for k in range(nparts):
            part = geo.dam_break_3d_particles(dx=dx, xrange=(cuts[k], cuts[k + 1]))
            n += sum(pa.get_number_of_particles() for pa in part)
        assert n == total

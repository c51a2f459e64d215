"""B200NNPS: the NNPS facade of the B200 backend.

Python-visible surface of pysph/base/nnps_base.pxd:279-371 that the hot path
uses: ``update()``, ``update_domain()``, ``set_context()``,
``get_nearest_particles()``, ``set_in_parallel()``,
``spatially_order_particles()`` and the attributes ``cell_size, hmin,
radius_scale, dim, xmin, xmax, ncells_per_dim, n_cells, particles``.

The build is a device-side counting sort of ALL arrays into one global cell
grid (not a linked list per array, linked_list_nnps.pyx:235-286); cell size,
bounds padding and the 2^28-cell guard follow the reference exactly
(nnps_base.pyx:942-978, :1520-1575, linked_list_nnps.pyx:293-343).
"""
import ctypes as C

import numpy as np

from . import _lib
from .kernels import kernel_id


class B200NNPS(object):
    def __init__(self, dim, particles, radius_scale=2.0, backend=None,
                 kernel=None, domain=None, cache=False, sort_gids=False):
        from .backend import B200Backend
        self.dim = dim
        self.particles = list(particles)
        self.narrays = len(self.particles)
        self.backend = backend or B200Backend(self.particles)
        self.ctx = self.backend.ctx
        if kernel is not None:
            self.set_kernel(kernel)
        self.radius_scale = radius_scale if kernel is None \
            else kernel.radius_scale
        self.domain = domain
        if domain is not None and (getattr(domain, 'is_periodic', False) or
                                   getattr(domain, 'is_mirror', False)):
            from .domain import DomainManager
            if not isinstance(domain, DomainManager):   # a PySPH DomainManager
                m = getattr(domain, 'manager', domain)
                domain = DomainManager(
                    m.xmin, m.xmax, m.ymin, m.ymax, m.zmin, m.zmax,
                    m.periodic_in_x, m.periodic_in_y, m.periodic_in_z,
                    n_layers=getattr(m, 'n_layers', 2.0),
                    mirror_in_x=getattr(m, 'mirror_in_x', False),
                    mirror_in_y=getattr(m, 'mirror_in_y', False),
                    mirror_in_z=getattr(m, 'mirror_in_z', False))
            domain.apply(self.ctx)
        self.in_parallel = False
        self.src_index = self.dst_index = 0
        self._grid = None
        # like the reference constructor (linked_list_nnps.pyx:84-88)
        self.update_domain()
        self.update()

    def set_kernel(self, kernel):
        self.ctx.call('b200sph_set_kernel', kernel_id(kernel), int(kernel.dim))
        self.radius_scale = kernel.radius_scale

    # -- reference protocol ---------------------------------------------------
    def update_domain(self, *args, **kw):
        self.ctx.call('b200sph_update_domain')

    def update(self, deferred=False):
        """deferred=True (used by the integrator only): when the persistent
        neighbour lists are reused, the measurement that proves them valid is
        enqueued but not awaited; ``confirm()`` must follow the evaluation."""
        self.ctx.call('b200sph_nnps_update_deferred' if deferred
                      else 'b200sph_nnps_update')
        self._grid = None

    def confirm(self):
        """True if the evaluation since the last deferred update must be
        repeated after a plain ``update()`` (the lists turned out stale)."""
        redo = C.c_int(0)
        self.ctx.call('b200sph_nnps_confirm', C.byref(redo))
        return bool(redo.value)

    def set_context(self, src_index, dst_index):
        self.src_index, self.dst_index = src_index, dst_index

    def set_in_parallel(self, in_parallel):
        self.in_parallel = bool(in_parallel)

    def spatially_order_particles(self, pa_index):
        # Solver.reorder_particles (solver.py:296-302) exists to restore memory
        # locality; here every build already sorts, so this is a no-op.
        return None

    def get_nearest_particles(self, src_index, dst_index, d_idx, nbrs=None,
                              cap=1 << 16):
        """Neighbours of destination d_idx (ascending source indices)."""
        buf = np.empty(cap, dtype=np.uint32)
        n = self.ctx.call('b200sph_get_neighbors', dst_index, src_index,
                          int(d_idx), buf.ctypes.data, cap)
        if n > cap:
            return self.get_nearest_particles(src_index, dst_index, d_idx,
                                              nbrs, cap=int(n))
        out = buf[:n].copy()
        if nbrs is not None and hasattr(nbrs, 'resize'):
            nbrs.resize(int(n))
            nbrs.set_data(out) if hasattr(nbrs, 'set_data') else None
        return out

    # -- grid attributes ------------------------------------------------------
    def _info(self):
        if self._grid is None:
            g = _lib.GridInfo()
            self.ctx.call('b200sph_get_grid', C.byref(g))
            self._grid = g
        return self._grid

    cell_size = property(lambda self: self._info().cell_size)
    hmin = property(lambda self: self._info().hmin)
    xmin = property(lambda self: np.array(self._info().xmin[:]))
    xmax = property(lambda self: np.array(self._info().xmax[:]))
    ncells_per_dim = property(lambda self: np.array(self._info().ncells[:]))
    n_cells = property(lambda self: self._info().n_cells)

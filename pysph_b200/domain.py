"""DomainManager descriptor: periodic boxes for the B200 NNPS.

Mirrors the constructor of ``pysph.base.nnps_base.DomainManager``
(nnps_base.pyx:226-347).  On the B200 backend a periodic axis wraps positions
into the box at every ``update_domain`` (``_box_wrap_periodic``, :699-743) and
particles interact with the periodic images of the others, but the images are
NOT materialised as ``tag = Ghost`` particles (``_create_ghosts_periodic``,
:744-940): the cell grid tiles the axis exactly and neighbour cells wrap, see
``k_list_build<true>`` in csrc/b200sph.cu; ``n_layers`` plays no role there.

A mirror axis (``mirror_in_x`` ..., nnps_base.pyx:329-335, ``_create_ghosts_mirror``
:506-689) DOES materialise its images as ``tag = Ghost`` particles appended after the
real ones -- every particle within ``n_layers`` cells of a plane, reflected, with the
normal velocity component negated, corner images in the reference's order.  The
reference re-selects them at every ``update_domain``; here the selection is made when the
neighbour lists are built (its width covers the list skin) and the images' values are
refreshed before every evaluation (``b200sph_set_mirror``).  WCSPH arrays, one GPU.
"""
import ctypes as C


class DomainManager(object):
    def __init__(self, xmin=-1000., xmax=1000., ymin=0., ymax=0., zmin=0., zmax=0.,
                 periodic_in_x=False, periodic_in_y=False, periodic_in_z=False,
                 n_layers=2.0, backend=None, props=None, mirror_in_x=False,
                 mirror_in_y=False, mirror_in_z=False):
        if xmax < xmin or ymax < ymin or zmax < zmin:
            raise ValueError("Invalid domain limits!")     # nnps_base.pyx:352-355
        self.xmin, self.xmax = float(xmin), float(xmax)
        self.ymin, self.ymax = float(ymin), float(ymax)
        self.zmin, self.zmax = float(zmin), float(zmax)
        self.periodic_in_x = bool(periodic_in_x)
        self.periodic_in_y = bool(periodic_in_y)
        self.periodic_in_z = bool(periodic_in_z)
        self.is_periodic = (self.periodic_in_x or self.periodic_in_y or
                            self.periodic_in_z)
        self.mirror_in_x = bool(mirror_in_x)
        self.mirror_in_y = bool(mirror_in_y)
        self.mirror_in_z = bool(mirror_in_z)
        self.is_mirror = self.mirror_in_x or self.mirror_in_y or self.mirror_in_z
        self.n_layers = n_layers
        self.manager = self

    def apply(self, ctx):
        lo = (C.c_double * 3)(self.xmin, self.ymin, self.zmin)
        hi = (C.c_double * 3)(self.xmax, self.ymax, self.zmax)
        per = (C.c_int * 3)(int(self.periodic_in_x), int(self.periodic_in_y),
                            int(self.periodic_in_z))
        ctx.call('b200sph_set_domain', lo, hi, per)
        if self.is_mirror:
            mir = (C.c_int * 3)(int(self.mirror_in_x), int(self.mirror_in_y),
                                int(self.mirror_in_z))
            ctx.call('b200sph_set_mirror', mir, float(self.n_layers))

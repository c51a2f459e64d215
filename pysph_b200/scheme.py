"""WCSPHScheme mirror: which loops run, in which order.

Restates the equation assembly of pysph/sph/scheme.py:388-506
(``WCSPHScheme.get_equations``) for the options the hot path covers
(``hg_correction``, ``update_h``, ``summation_density``,
``tensile_correction``, ``nu != 0`` -> ``LaminarViscosity``); ``delta_sph`` (gradient
correction + density diffusion) has no kernels and raises.
"""
from .equations import (Group, SummationDensity, TaitEOS, TaitEOSHGCorrection,
                        ContinuityEquation, MomentumEquation, XSPHCorrection,
                        LaminarViscosity, UpdateSmoothingLengthFerrari)


class WCSPHScheme(object):
    def __init__(self, fluids, solids, dim, rho0, c0, h0, hdx, gamma=7.0,
                 gx=0.0, gy=0.0, gz=0.0, alpha=0.1, beta=0.0, delta=0.1,
                 nu=0.0, tensile_correction=False, hg_correction=False,
                 update_h=False, delta_sph=False, summation_density=False):
        if delta_sph:
            raise NotImplementedError(
                'B200 backend: delta_sph (gradient correction + density diffusion, '
                'scheme.py:431-447) has no kernels')
        self.nu = nu
        self.fluids = list(fluids)
        self.solids = list(solids)
        self.dim = dim
        self.rho0 = rho0
        self.c0 = c0
        self.h0 = h0
        self.hdx = hdx
        self.gamma = gamma
        self.gx, self.gy, self.gz = gx, gy, gz
        self.alpha = alpha
        self.beta = beta
        self.tensile_correction = tensile_correction
        self.hg_correction = hg_correction
        self.update_h = update_h
        self.summation_density = summation_density

    def get_timestep(self, cfl=0.5):
        # scheme.py:357-358
        return cfl * self.h0 / self.c0

    def get_equations(self):
        equations = []
        all_ = self.fluids + self.solids
        if self.summation_density:
            equations.append(Group(
                [SummationDensity(dest=f, sources=all_) for f in self.fluids],
                real=False))
        g1 = [TaitEOS(dest=f, sources=None, rho0=self.rho0, c0=self.c0,
                      gamma=self.gamma) for f in self.fluids]
        for s in self.solids:
            cls = TaitEOSHGCorrection if self.hg_correction else TaitEOS
            g1.append(cls(dest=s, sources=None, rho0=self.rho0, c0=self.c0,
                          gamma=self.gamma))
        equations.append(Group(g1, real=False))

        g2 = [ContinuityEquation(dest=s, sources=self.fluids)
              for s in self.solids]
        for f in self.fluids:
            if not self.summation_density:
                g2.append(ContinuityEquation(dest=f, sources=all_))
            g2.append(MomentumEquation(
                dest=f, sources=all_, c0=self.c0, alpha=self.alpha,
                beta=self.beta, gx=self.gx, gy=self.gy, gz=self.gz,
                tensile_correction=self.tensile_correction))
            g2.append(XSPHCorrection(dest=f, sources=[f]))
            if abs(self.nu) > 1e-14:            # scheme.py:486-496: g2.insert(-1, eq)
                g2.insert(-1, LaminarViscosity(dest=f, sources=all_, nu=self.nu))
        equations.append(Group(g2))

        if self.update_h:
            equations.append(Group(
                [UpdateSmoothingLengthFerrari(dest=f, sources=None,
                                              dim=self.dim, hdx=self.hdx)
                 for f in self.fluids], real=False))
        return equations

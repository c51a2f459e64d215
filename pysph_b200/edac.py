"""EDAC scheme, internal-flow (transport-velocity) branch, with or without solid walls.

Mirrors pysph/sph/wc/edac.py for the Taylor-Green configuration (BASELINE configs[3],
pysph/examples/taylor_green.py:190-203) and for internal flows between walls (cavity,
Poiseuille / Couette): ``ComputeAveragePressure`` (:62-79), ``SolidWallPressureBC`` (:136-166),
``SourceNumberDensity`` (:177-183), ``SetWallVelocity`` (:186-230), ``EDACEquation`` (:354-386),
``MomentumEquationPressureGradient`` (:389-488), ``EDACTVFStep`` (:491-540) and ``EDACScheme``
(:543-971, ``get_equations`` for ``pb != 0`` -- internal flows with the transport velocity -- and
for ``pb == 0`` -- the external-flow branch :882-971 with ``EDACStep`` (:82-133), the
number-density ``MomentumEquation`` (:301-352), ``XSPHCorrection`` and ``ClampWallPressure``
(:169-174)).  Inviscid solids and inlet / outlet raise NotImplementedError.
"""
from .equations import Equation, Group, XSPHCorrection
from .integrator import IntegratorStep
from .transport_velocity import (MomentumEquationArtificialStress,
                                 MomentumEquationArtificialViscosity,
                                 MomentumEquationViscosity, SolidWallNoSlipBC,
                                 SummationDensity, VolumeSummation)


class ComputeAveragePressure(Equation):
    """wc/edac.py:62-79"""


class SolidWallPressureBC(Equation):
    """wc/edac.py:136-166: the fluid pressure extrapolated to a wall particle (Adami, Hu)."""

    def __init__(self, dest, sources, gx=0.0, gy=0.0, gz=0.0):
        self.gx = gx
        self.gy = gy
        self.gz = gz
        super(SolidWallPressureBC, self).__init__(dest, sources)


class ClampWallPressure(Equation):
    """wc/edac.py:169-174: p = max(p, 0) on a wall, behind SolidWallPressureBC"""


class MomentumEquation(Equation):
    """wc/edac.py:301-352: the pressure gradient in the number-density form of Hu and Adams
    (NOT wc/basic.py's MomentumEquation: recognised by its module)"""

    def __init__(self, dest, sources, c0, gx=0.0, gy=0.0, gz=0.0, tdamp=0.0):
        self.gx = gx
        self.gy = gy
        self.gz = gz
        self.c0 = c0
        self.tdamp = tdamp
        super(MomentumEquation, self).__init__(dest, sources)


class SourceNumberDensity(Equation):
    """wc/edac.py:177-183: wij = sum over the fluid neighbours of W"""


class SetWallVelocity(Equation):
    """wc/edac.py:186-230: uf = the fluid velocity extrapolated to the wall, ug = 2 u - uf"""


class EDACEquation(Equation):
    """wc/edac.py:354-386"""

    def __init__(self, dest, sources, cs, nu, rho0):
        self.cs = cs
        self.nu = nu
        self.rho0 = rho0
        super(EDACEquation, self).__init__(dest, sources)


class MomentumEquationPressureGradient(Equation):
    """wc/edac.py:389-488 (the variant that subtracts the average pressure)"""

    def __init__(self, dest, sources, pb, gx=0., gy=0., gz=0., tdamp=0.0):
        self.pb = pb
        self.gx = gx
        self.gy = gy
        self.gz = gz
        self.tdamp = tdamp
        super(MomentumEquationPressureGradient, self).__init__(dest, sources)


class EDACTVFStep(IntegratorStep):
    """wc/edac.py:491-540 (device kernel: k_stage_tvf)"""


class EDACStep(IntegratorStep):
    """wc/edac.py:82-133: the stepper without transport velocity (device kernel: k_stage_tvf, ext)"""


class EDACScheme(object):
    def __init__(self, fluids, solids, dim, c0, nu, rho0, pb=0.0, gx=0.0, gy=0.0,
                 gz=0.0, tdamp=0.0, eps=0.0, h=0.0, edac_alpha=0.5, alpha=0.0,
                 bql=True, clamp_p=False, inlet_outlet_manager=None,
                 inviscid_solids=None):
        self.c0 = c0
        self.nu = nu
        self.rho0 = rho0
        self.gx, self.gy, self.gz = gx, gy, gz
        self.tdamp = tdamp
        self.dim = dim
        self.eps = eps
        self.fluids = list(fluids)
        self.solids = list(solids)
        self.pb = pb
        self.bql = bql
        self.clamp_p = clamp_p
        self.edac_alpha = edac_alpha
        self.alpha = alpha
        self.h = h
        self.inlet_outlet_manager = inlet_outlet_manager
        self.inviscid_solids = [] if inviscid_solids is None else inviscid_solids
        self.attributes_changed()

    def attributes_changed(self):                      # wc/edac.py:651-655
        if self.pb is not None:
            self.use_tvf = abs(self.pb) > 1e-14
        if self.h is not None and self.c0 is not None:
            self.art_nu = self.edac_alpha * self.h * self.c0 / 8

    def configure(self, **kw):
        for k, v in kw.items():
            if not hasattr(self, k):
                raise RuntimeError('Parameter {param} not defined for {scheme}.'
                                   .format(param=k, scheme=self.__class__.__name__))
            setattr(self, k, v)
        self.attributes_changed()

    def _get_edac_nu(self):                            # wc/edac.py:766-774
        return self.art_nu if self.art_nu > 0 else self.nu

    def get_steppers(self):                            # wc/edac.py:682-687
        cls = EDACTVFStep if self.use_tvf else EDACStep
        return dict((f, cls()) for f in self.fluids)

    def get_equations(self):                           # wc/edac.py:704-708, :776-880
        if self.inviscid_solids or self.inlet_outlet_manager is not None:
            raise NotImplementedError('B200 backend: EDAC with inviscid solids / inlet-outlet')
        if not self.use_tvf:
            return self._get_external_flow_equations()
        edac_nu = self._get_edac_nu()
        all_ = self.fluids + self.solids
        has_solids = len(self.solids) > 0
        group1, avg_p_group = [], []
        for fluid in self.fluids:
            group1.append(SummationDensity(dest=fluid, sources=all_))
            if self.bql:
                eq = ComputeAveragePressure(dest=fluid, sources=all_)
                (avg_p_group if has_solids else group1).append(eq)
        for solid in self.solids:                      # :815-822
            group1.extend([
                SourceNumberDensity(dest=solid, sources=self.fluids),
                VolumeSummation(dest=solid, sources=all_),
                SolidWallPressureBC(dest=solid, sources=self.fluids, gx=self.gx, gy=self.gy,
                                    gz=self.gz),
                SetWallVelocity(dest=solid, sources=self.fluids)])
        group2 = []
        for fluid in self.fluids:
            group2.append(MomentumEquationPressureGradient(
                dest=fluid, sources=all_, pb=self.pb, gx=self.gx, gy=self.gy,
                gz=self.gz, tdamp=self.tdamp))
            if self.alpha > 0.0:
                group2.append(MomentumEquationArtificialViscosity(
                    dest=fluid, sources=self.fluids + self.solids, alpha=self.alpha,
                    c0=self.c0))
            if self.nu > 0.0:
                group2.append(MomentumEquationViscosity(
                    dest=fluid, sources=self.fluids, nu=self.nu))
            if has_solids and self.nu > 0.0:
                group2.append(SolidWallNoSlipBC(dest=fluid, sources=self.solids, nu=self.nu))
            group2.extend([
                MomentumEquationArtificialStress(dest=fluid, sources=self.fluids),
                EDACEquation(dest=fluid, sources=all_, nu=edac_nu, cs=self.c0,
                             rho0=self.rho0)])
        groups = [Group(equations=group1, real=False)]
        if self.bql and has_solids:
            # the average pressure *after* the wall pressure is set up (:840-842)
            groups.append(Group(equations=avg_p_group, real=True))
        groups.append(Group(equations=group2))
        return groups

    def _get_external_flow_equations(self):            # wc/edac.py:882-971
        edac_nu = self._get_edac_nu()
        all_ = self.fluids + self.solids
        group1 = [SummationDensity(dest=fluid, sources=all_) for fluid in self.fluids]
        for solid in self.solids:
            group1.extend([
                SourceNumberDensity(dest=solid, sources=self.fluids),
                VolumeSummation(dest=solid, sources=all_),
                SolidWallPressureBC(dest=solid, sources=self.fluids, gx=self.gx, gy=self.gy,
                                    gz=self.gz),
                SetWallVelocity(dest=solid, sources=self.fluids)])
            if self.clamp_p:
                group1.append(ClampWallPressure(dest=solid, sources=None))
        group2 = []
        for fluid in self.fluids:
            group2.append(MomentumEquation(dest=fluid, sources=all_, gx=self.gx, gy=self.gy,
                                           gz=self.gz, c0=self.c0, tdamp=self.tdamp))
            if self.alpha > 0.0:
                group2.append(MomentumEquationArtificialViscosity(
                    dest=fluid, sources=self.fluids + self.solids, alpha=self.alpha,
                    c0=self.c0))
            if self.nu > 0.0:
                group2.append(MomentumEquationViscosity(
                    dest=fluid, sources=self.fluids, nu=self.nu))
            if len(self.solids) > 0 and self.nu > 0.0:
                group2.append(SolidWallNoSlipBC(dest=fluid, sources=self.solids, nu=self.nu))
            group2.extend([
                EDACEquation(dest=fluid, sources=all_, nu=edac_nu, cs=self.c0,
                             rho0=self.rho0),
                XSPHCorrection(dest=fluid, sources=[fluid], eps=self.eps)])
        return [Group(equations=group1, real=False), Group(equations=group2)]

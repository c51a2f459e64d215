"""pysph_b200 -- B200-native (sm_100a) evaluator for PySPH's WCSPH hot path.

The product is ``libb200sph.so`` (hand-written CUDA behind the C-ABI of
``include/b200sph.h``) plus the thin Python adapters in this package that
mirror the reference's plugin seam for this path:

==========================  ==================================================
reference (pypr/pysph)      here
==========================  ==================================================
compiled AccelerationEval   :class:`B200AccelerationEval`
compiled Integrator         :class:`PECIntegrator`, :class:`EPECIntegrator`
NNPS (LinkedListNNPS)       :class:`B200NNPS`
DeviceHelper (``pa.gpu``)   :class:`B200DeviceHelper` / :class:`B200Backend`
Equation / Group            :mod:`pysph_b200.equations`
WCSPHScheme.get_equations   :class:`pysph_b200.scheme.WCSPHScheme`
Solver step loop            :class:`B200Solver`
ParallelManager             :class:`pysph_b200.parallel.SlabParallelManager`
==========================  ==================================================

There is no CPU fallback: importing the adapters is cheap, but creating a
backend needs the built library and a CUDA device.
"""
from .particle_array import (ParticleArray, get_particle_array,
                             get_particle_array_wcsph, get_particle_array_edac,
                             get_particle_array_edac_wall, get_particle_array_edac_ext,
                             get_particle_array_elastic_dynamics)
from .kernels import CubicSpline, WendlandQuintic, QuinticSpline, Gaussian
from .equations import (Equation, Group, SummationDensity, ContinuityEquation, LaminarViscosity,
                        MonaghanArtificialViscosity, XSPHCorrection, TaitEOS,
                        TaitEOSHGCorrection, MomentumEquation,
                        UpdateSmoothingLengthFerrari)
from .scheme import WCSPHScheme
from .backend import B200Backend, B200DeviceHelper
from .domain import DomainManager
from .nnps import B200NNPS
from .acceleration_eval import B200AccelerationEval
from .integrator import (B200Integrator, PECIntegrator, EPECIntegrator,
                         WCSPHStep)
from .solver import B200Solver
from .edac import EDACScheme, EDACTVFStep, EDACStep
from .output import dump, load
from .solid_mech import ElasticSolidsScheme, SolidMechStep

__version__ = '0.1.0'


def make_wcsph_solver(particles, params, kernel, device=0, **solver_kw):
    """Build a ready-to-step WCSPH solver from the parameter dict returned by
    ``geometry.dam_break_3d_params`` / ``dam_break_2d_params``."""
    p = dict(params)
    integ_name = p.pop('integrator', 'EPEC')
    dt0 = p.pop('dt0')
    n_damp = p.pop('n_damp', 0)
    cfl = p.pop('cfl', 0.3)
    scheme = WCSPHScheme(**p)
    names = scheme.fluids + scheme.solids
    cls = {'EPEC': EPECIntegrator, 'PEC': PECIntegrator}[integ_name]
    integrator = cls(**dict((n, WCSPHStep()) for n in names))
    kw = dict(adaptive_timestep=True, cfl=cfl, n_damp=n_damp, tf=1e9,
              fixed_h=not p.get('update_h', False), device=device)
    kw.update(solver_kw)
    return B200Solver(particles, scheme.get_equations(), kernel, integrator,
                      dt=dt0, **kw)


def make_edac_solver(particles, scheme, kernel, dt, domain=None, integrator='PEC',
                     device=0, **solver_kw):
    """Fixed-dt solver for an :class:`EDACScheme` (the reference's configure_solver
    default is PECIntegrator + EDACTVFStep, wc/edac.py:657-702)."""
    cls = {'EPEC': EPECIntegrator, 'PEC': PECIntegrator}[integrator]
    integ = cls(**scheme.get_steppers())
    kw = dict(adaptive_timestep=False, tf=1e9, fixed_h=True, device=device,
              domain=domain)
    kw.update(solver_kw)
    return B200Solver(particles, scheme.get_equations(), kernel, integ, dt=dt, **kw)


def make_elastic_solver(particles, scheme, kernel, dt, integrator='EPEC', device=0,
                        **solver_kw):
    """Fixed-dt solver for an :class:`ElasticSolidsScheme` (configure_solver's default
    is EPECIntegrator + SolidMechStep, solid_mech/basic.py:653-684)."""
    cls = {'EPEC': EPECIntegrator, 'PEC': PECIntegrator}[integrator]
    integ = cls(**scheme.get_steppers())
    kw = dict(adaptive_timestep=False, tf=1e9, fixed_h=True, device=device)
    kw.update(solver_kw)
    return B200Solver(particles, scheme.get_equations(), kernel, integ, dt=dt, **kw)

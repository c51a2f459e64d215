"""Elastic dynamics (Gray, Monaghan & Swift 2001), mirroring pysph/sph/solid_mech/basic.py:
elastic solids, optionally with rigid ``solids`` as sources (BASELINE configs[4],
SURVEY.md 8f-2).

STATUS: the CUDA kernels behind these descriptors (``k_solid_pass1/2``,
``k_stage_solid``) were written after this round's GPU budget was spent.  They
compile for sm_100a and follow the oracle that is pinned to the reference
(tests/golden/solid_cases.json), but have not run on hardware yet:
tests/test_zz_gpu_solid_unvalidated.py is their first contact with a B200.
"""
from .equations import (ContinuityEquation, Equation, Group,
                        MonaghanArtificialViscosity, XSPHCorrection)
from .integrator import IntegratorStep


class IsothermalEOS(Equation):
    """solid_mech/basic.py:93-101"""


class VelocityGradient2D(Equation):
    """basic_equations.py:67-98"""


class VelocityGradient3D(Equation):
    """basic_equations.py:101-148"""


class MonaghanArtificialStress(Equation):
    """solid_mech/basic.py:104-242"""

    def __init__(self, dest, sources, eps=0.3):
        self.eps = eps
        super(MonaghanArtificialStress, self).__init__(dest, sources)


class MomentumEquationWithStress(Equation):
    """solid_mech/basic.py:245-387"""


class HookesDeviatoricStressRate(Equation):
    """solid_mech/basic.py:390-505"""


class SolidMechStep(IntegratorStep):
    """integrator_step.py:173-252 (device kernel: k_stage_solid)"""


class ElasticSolidsScheme(object):
    """solid_mech/basic.py:592-651; ``use_3d_gradient`` swaps VelocityGradient2D (what
    the reference scheme always emits) for VelocityGradient3D, which a 3-D run needs.
    ``ghost_group1`` makes group 1 a ``Group(real=False)``: with the slab decomposition
    importing two kernel supports of ghosts (parallel.make_rings_slab_solver) the ghosts'
    p and artificial stress are then this evaluation's, and N ranks reproduce one."""

    def __init__(self, elastic_solids, solids, dim, artificial_stress_eps=0.3,
                 xsph_eps=0.5, alpha=1.0, beta=1.0, use_3d_gradient=None,
                 ghost_group1=False):
        self.ghost_group1 = bool(ghost_group1)
        self.elastic_solids = list(elastic_solids)
        self.solids = list(solids)
        self.dim = dim
        self.alpha = alpha
        self.beta = beta
        self.xsph_eps = xsph_eps
        self.artificial_stress_eps = artificial_stress_eps
        self.use_3d_gradient = (dim == 3) if use_3d_gradient is None else use_3d_gradient

    def get_steppers(self):
        return dict((n, SolidMechStep()) for n in self.elastic_solids)

    def get_equations(self):
        all_ = self.solids + self.elastic_solids
        grad = VelocityGradient3D if self.use_3d_gradient else VelocityGradient2D
        g1, g2 = [], []
        for es in self.elastic_solids:
            g1.append(IsothermalEOS(es, sources=None))
            g1.append(grad(dest=es, sources=all_))
            g1.append(MonaghanArtificialStress(dest=es, sources=None,
                                               eps=self.artificial_stress_eps))
        for es in self.elastic_solids:
            g2.append(ContinuityEquation(dest=es, sources=all_))
            g2.append(MomentumEquationWithStress(dest=es, sources=all_))
            g2.append(MonaghanArtificialViscosity(dest=es, sources=all_,
                                                  alpha=self.alpha, beta=self.beta))
            g2.append(HookesDeviatoricStressRate(dest=es, sources=None))
            g2.append(XSPHCorrection(dest=es, sources=[es], eps=self.xsph_eps))
        return [Group(equations=g1, real=not self.ghost_group1), Group(g2)]

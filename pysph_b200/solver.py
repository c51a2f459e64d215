"""Time loop around the B200 integrator.

A small restatement of the parts of ``pysph.solver.solver.Solver`` that sit
directly on the hot path -- ``setup`` wiring (solver.py:186-229), the step loop
(:460-507) and the adaptive / damped time step (:647-688, :756-779) -- so that
the benchmark and the parity tests can run whole simulations without PySPH
installed.  With PySPH present the recommended route is INTEGRATION.md (the
reference Solver drives B200AccelerationEval / B200Integrator unchanged).
"""
import numpy as np

from .acceleration_eval import B200AccelerationEval
from .backend import B200Backend
from .nnps import B200NNPS


class B200Solver(object):
    def __init__(self, particles, equations, kernel, integrator, dt, tf=1.0,
                 adaptive_timestep=False, cfl=0.3, n_damp=0, fixed_h=False,
                 device=0, backend=None, capacity_factor=1.0,
                 extra_capacity=0, domain=None):
        self.particles = list(particles)
        self.kernel = kernel
        self.integrator = integrator
        self.dt = dt
        self.tf = tf
        self.t = 0.0
        self.count = 0
        self.adaptive_timestep = adaptive_timestep
        self.cfl = cfl
        self.n_damp = n_damp
        self._damping_factor = 1.0
        self.pm = None
        self.in_parallel = False

        self.backend = backend or B200Backend(
            self.particles, device=device, capacity_factor=capacity_factor,
            extra_capacity=extra_capacity)
        # Solver.setup, solver.py:186-229
        self.a_eval = B200AccelerationEval(self.particles, equations, kernel,
                                           backend=self.backend)
        self.nnps = B200NNPS(kernel.dim, self.particles, backend=self.backend,
                             kernel=kernel, domain=domain)
        self.a_eval.set_nnps(self.nnps)
        integrator.set_acceleration_evals([self.a_eval])
        integrator.set_nnps(self.nnps)
        integrator.set_fixed_h(fixed_h)
        self._initialised = False

    def set_parallel_manager(self, pm):
        dom = self.nnps.domain
        if pm is not None and dom is not None and getattr(dom, 'is_periodic', False):
            raise NotImplementedError(
                'B200 backend: periodic domains with the slab decomposition')
        self.pm = pm
        self.in_parallel = pm is not None
        self.integrator.set_parallel_manager(pm)

    # -- time step (solver.py:647-688, 756-779) -------------------------------
    def _compute_timestep(self):
        undamped = self.dt / self._damping_factor
        if not self.adaptive_timestep:
            return undamped
        dt = self.integrator.compute_time_step(undamped, self.cfl)
        if self.in_parallel:
            dt = self.pm.update_time_steps(1e20 if dt is None else dt)
        elif dt is None:
            dt = undamped
        return dt

    def _damp_timestep(self, dt):
        if self.count < self.n_damp and self.n_damp > 0:
            frac = (self.count + 1) / float(self.n_damp)
            self._damping_factor = 0.5 * (np.sin(np.pi * (-0.5 + frac)) + 1.0)
        else:
            self._damping_factor = 1.0
        return dt * self._damping_factor

    def _get_timestep(self):
        return self._damp_timestep(self._compute_timestep())

    # -- stepping -------------------------------------------------------------
    def initialise(self):
        if not self._initialised:
            if self.pm is not None:
                self.pm.update()
                self.nnps.update_domain()
                self.nnps.update()
            self.integrator.initial_acceleration(self.t, self.dt)  # solver.py:454
            self.dt = self._get_timestep()                          # solver.py:458
            self._initialised = True

    def step(self):
        """One iteration of the solve loop (solver.py:460-491)."""
        self.initialise()
        self.integrator.step(self.t, self.dt)
        self.t += self.dt
        self.count += 1
        self.dt = self._get_timestep()

    def solve(self, max_steps):
        self.initialise()
        while self.count < max_steps and (self.tf - self.t) > 1e-15:
            self.step()

    def pull(self, props=None):
        self.backend.pull_all(props)

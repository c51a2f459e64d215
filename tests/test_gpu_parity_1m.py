"""Field-level parity at BASELINE configs[1]'s FULL size: the 3-D dam break at dx = 0.00877
(1 224 945 particles, 81 M directed pairs per evaluation) against the fp64 oracle running
on every host core -- one evaluation of a perturbed state (random velocities and densities
so that every term of Continuity / Momentum + AV / XSPH is exercised) and then five
adaptive EPEC steps.  At this size the packed records (58 MB) and the neighbour lists
(0.5 GB) no longer sit in L1 / partly not in L2, the grid has 128 x 44 x 44 cells and the
lists are 120 entries long: what the dx = 0.05 parity tests cannot show.

Tolerances (fp32 pair arithmetic on cell-relative coordinates, fp64 integrated state):
pair count EQUAL to the oracle's; every acceleration-like field within 2e-5 of its largest
magnitude (au / av / aw: of max(|a|, g)); after 5 steps positions within 2e-6 h, velocities
within 2e-6 c0, density within 2e-7 rho0, t within 1e-6 relative.
"""
import os

import numpy as np
import pytest

from helpers import ACC_FIELDS, copy_arrays
from oracle import oracle as orc

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

DX = 0.00877
TOL_EVAL = 2e-5


def _threads():
    try:
        return max(1, min(len(os.sched_getaffinity(0)), 64))
    except Exception:
        return max(1, min(os.cpu_count() or 1, 64))


def test_dam_break_3d_full_size_eval_and_steps(gpu_device):
    import test_gpu_parity as P
    from pysph_b200 import geometry as geo
    pas = geo.dam_break_3d_particles(dx=DX)
    params = geo.dam_break_3d_params(DX)
    assert sum(pa.get_number_of_particles() for pa in pas) == 1224945
    rs = np.random.RandomState(20260923)
    f = pas[0]
    n = f.get_number_of_particles()
    for k in ('u', 'v', 'w'):
        f.properties[k][:] = rs.normal(scale=0.5, size=n)
    f.rho[:] *= 1.0 + 0.01 * rs.uniform(-1, 1, n)
    opas = copy_arrays(pas)
    s = P.make_solver(pas, P.scheme_params(params), 'CubicSpline')
    o = orc.WCSPHOracleSolver(opas, params, 'CubicSpline', threads=_threads())
    s.a_eval.count_pairs = True
    s.initialise()
    o.initialise()
    assert s.a_eval.last_pairs == o.pairs_last_eval
    assert s.a_eval.last_pairs > 80e6
    assert abs(s.dt - o.dt) <= 1e-6 * o.dt
    s.pull()
    worst = {}
    for pa, oa in zip(pas, opas):
        nr = oa.num_real_particles
        for fld in ACC_FIELDS:
            want = oa.properties[fld][:nr]
            got = pa.properties[fld][:nr]
            scale = max(np.max(np.abs(want)), 1e-30)
            if fld in ('au', 'av', 'aw'):
                scale = max(scale, 9.81)
            err = np.max(np.abs(got - want)) / scale
            worst[(pa.name, fld)] = err
            assert err <= TOL_EVAL, (pa.name, fld, err)
    s.a_eval.count_pairs = False
    for _ in range(5):
        s.step()
        o.step()
    s.pull()
    assert abs(s.t - o.t) <= 1e-6 * o.t
    P._compare_state(pas, opas, tol_pos=2e-6, tol_vel=2e-6, tol_rho=2e-7,
                     h0=params['h0'], c0=params['c0'], rho0=params['rho0'])
    print('full-size parity: worst single-evaluation error %.2e (%s)'
          % (max(worst.values()), max(worst, key=worst.get)))

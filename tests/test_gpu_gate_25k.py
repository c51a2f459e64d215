"""BASELINE configs[0] at the size it is quoted at: the 2-D dam break with `--dx 0.01`,
20 301 fluid + 4 852 boundary particles, WendlandQuintic, PEC, update_h, HG correction,
adaptive damped dt -- including the reference's quirk that the particles keep the
module-level h = 0.039 and m = 0.9 (dam_break_2d.py:35,45-47,230-232):
~183 neighbours per particle and 9x over-heavy particles, so the run is a parity gate, not
physics -- it blows up (dt -> 1e-8, in the oracle and on the device alike) after ~33 steps,
and the comparison stops at 30.  tests/test_gpu_parity.py::test_dam_break_2d_gate runs the
same path at the example's default dx = 0.03.  First passed on a B200 in the driver's
round-1 run (GPUTEST_r01.json)."""
import numpy as np
import pytest

from helpers import copy_arrays
from oracle import oracle as orc

pytestmark = [pytest.mark.gpu,
              pytest.mark.timeout(300)]


def test_dam_break_2d_gate_25k(gpu_device):
    from pysph_b200 import geometry as geo
    import test_gpu_parity as P
    pas = geo.dam_break_2d_particles(dx=0.01)
    params = geo.dam_break_2d_params(dx=0.01)
    assert [pa.get_number_of_particles() for pa in pas] == [20301, 4852]
    opas = copy_arrays(pas)
    s = P.make_solver(pas, P.scheme_params(params), 'WendlandQuintic')
    o = orc.WCSPHOracleSolver(opas, params, 'WendlandQuintic', threads=4)
    s.a_eval.count_pairs = True
    s.initialise()
    o.initialise()
    assert s.a_eval.last_pairs == o.pairs_last_eval == 3726345
    for _ in range(30):
        s.step()
        o.step()
    s.pull()
    assert abs(s.t - o.t) <= 1e-5 * o.t
    P._compare_state(pas, opas, tol_pos=1e-6, tol_vel=1e-6, tol_rho=1e-6,
                     h0=params['h0'], c0=params['c0'], rho0=params['rho0'])
    assert np.allclose(pas[0].h, opas[0].h, rtol=1e-6)

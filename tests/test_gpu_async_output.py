"""Asynchronous output path (SURVEY.md 8f-4 "async D2H of only output props at pfreq"):
`b200sph_snapshot_take / fetch / release` and `B200Solver.solve(pfreq=..., asynchronous=True)`.
The ordering between the time loop's stream and the copy stream is what only a GPU can
show; the same comparison also runs on the host emulation of the library
(tests/test_library_on_cpu.py::test_async_output_small).  First passed on a B200 in the
driver's round-1 run (GPUTEST_r01.json)."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.timeout(300)]


def test_async_dumps_equal_sync_dumps(gpu_device, tmp_path):
    import pysph_b200 as pb
    from pysph_b200 import geometry as geo, output
    dx = 0.04                      # ~14 k fluid particles: the loop runs ahead of the writer
    files = {}
    for mode, asyn in (('async', True), ('sync', False)):
        pas = geo.dam_break_3d_particles(dx=dx)
        s = pb.make_wcsph_solver(pas, geo.dam_break_3d_params(dx), pb.CubicSpline(dim=3))
        d = tmp_path / mode
        s.solve(40, pfreq=5, output_directory=str(d), fname='db', asynchronous=asyn)
        names = sorted(os.listdir(str(d)))
        assert names == ['db_%05d.npz' % k for k in range(0, 41, 5)]
        files[mode] = dict((f, output.load(str(d / f))) for f in names)
    for f, want in files['sync'].items():
        got = files['async'][f]
        for k in ('t', 'dt', 'count'):
            assert float(got['solver_data'][k]) == float(want['solver_data'][k]), (f, k)
        for name, pa in want['arrays'].items():
            q = got['arrays'][name]
            for k in pa.output_property_arrays:
                # the same deterministic run: every dump is bitwise the synchronous one, i.e.
                # the snapshot was taken at the right place in the stream and not overwritten
                assert np.array_equal(q.properties[k], pa.properties[k]), (f, name, k)
    a, b = files['async']['db_00035.npz'], files['async']['db_00040.npz']
    assert np.max(np.abs(a['arrays']['fluid'].z - b['arrays']['fluid'].z)) > 0

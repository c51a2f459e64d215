"""Output / restart files: pysph_b200.output against a fixture that
the REFERENCE's own pysph/solver/output.py wrote (tests/golden/ref_dump.npz, made by
oracle/gen_golden.py, which also checks that the reference's load() reads our dump)."""
import os

import numpy as np
import pytest

from helpers import GOLD, load_golden
import pysph_b200 as pb
from pysph_b200 import output


def _structure(d, depth=0):
    """nested key / type skeleton of a dumped dictionary"""
    if isinstance(d, dict):
        return dict((k, _structure(v, depth + 1)) for k, v in sorted(d.items()))
    if isinstance(d, np.ndarray):
        return ('ndarray', str(d.dtype))
    if isinstance(d, (list, tuple)):
        return 'list'
    return type(d).__name__


def test_load_reads_the_reference_dump():
    exp = load_golden('ref_dump_expect.json')
    r = output.load(os.path.join(GOLD, 'ref_dump.npz'))
    sd = r['solver_data']
    assert (float(sd['dt']), float(sd['t']), int(sd['count'])) == \
        (exp['solver_data']['dt'], exp['solver_data']['t'], exp['solver_data']['count'])
    assert sorted(r['arrays']) == ['boundary', 'fluid']
    for name, props in exp['arrays'].items():
        pa = r['arrays'][name]
        assert pa.output_property_arrays == exp['output_property_arrays'][name]
        assert sorted(pa.properties) == exp['all_properties'][name]
        n = len(props['x'])
        assert pa.get_number_of_particles() == n          # only_real=True: ghosts dropped
        for k, v in props.items():
            assert np.array_equal(pa.properties[k], np.array(v)), (name, k)
        assert pa.gid.dtype == np.uint32 and pa.tag.dtype == np.int32
        assert np.all(pa.au == 0.0)                       # not written: default
    assert np.array_equal(r['arrays']['boundary'].constants['total_mass'], [2.8])


def test_dump_writes_the_reference_layout(tmp_path):
    ref = np.load(os.path.join(GOLD, 'ref_dump.npz'), allow_pickle=True)
    back = output.load(os.path.join(GOLD, 'ref_dump.npz'))
    pas = [back['arrays']['fluid'], back['arrays']['boundary']]
    f = output.dump(str(tmp_path / 'mine_00300.npz'), pas, back['solver_data'])
    assert f.endswith('mine_00300.npz')
    mine = np.load(f, allow_pickle=True)
    assert sorted(mine.files) == sorted(ref.files) == ['particles', 'solver_data', 'version']
    assert int(mine['version']) == int(ref['version']) == 2
    for key in ('particles', 'solver_data'):
        a, b = mine[key].reshape(1)[0], ref[key].reshape(1)[0]
        assert _structure(a) == _structure(b), key
    a, b = mine['particles'].reshape(1)[0], ref['particles'].reshape(1)[0]
    for name in a:
        for k, v in b[name]['arrays'].items():
            assert np.array_equal(a[name]['arrays'][k], v), (name, k)
        assert a[name]['properties']['gid'] == b[name]['properties']['gid']
    # '.hdf5' without h5py falls back to npz (output.py:403-412)
    g = output.dump(str(tmp_path / 'x.hdf5'), pas, back['solver_data'], compress=True)
    assert g.endswith('x.npz') and output.load(g)['arrays']['fluid'].x.size == 7
    with pytest.raises(RuntimeError):
        output.load(str(tmp_path / 'missing.npz'))


def test_detailed_output_and_ghosts(tmp_path):
    pa = pb.get_particle_array_wcsph(name='f', x=np.arange(6.0), h=1.0, m=1.0)
    pa.set_num_real_particles(4)
    pa.au[:] = 3.0
    f = output.dump(str(tmp_path / 'd'), [pa], {'dt': 1.0, 't': 2.0, 'count': 3},
                    detailed_output=True, only_real=False)
    q = output.load(f)['arrays']['f']
    assert q.get_number_of_particles() == 6 and np.all(q.au == 3.0)
    assert set(q.properties) == set(pa.properties)


@pytest.mark.gpu
def test_dump_and_restart_on_device(gpu_device, tmp_path):
    """Solver.dump_output / load_output (solver.py:520-624): only the output properties
    of the real particles leave the device, the file holds t, count and the UNDAMPED dt
    (solver.py:747-753).  A restart is what the reference does with the file: a fresh
    solver, initial_acceleration, a new damped adaptive dt (:454-458) -- so, like in the
    reference, it is not bitwise the uninterrupted run (TaitEOSHGCorrection clamps the
    solids' density again, wc/basic.py:119-120); it is compared with the oracle restarted
    from the same file the same way, inside the damping window."""
    from pysph_b200 import geometry as geo
    from oracle import oracle as orc
    dx = 0.05
    params = geo.dam_break_3d_params(dx)
    assert params['n_damp'] > 20
    pas = geo.dam_break_3d_particles(dx=dx)
    s = pb.make_wcsph_solver(pas, dict(params), pb.CubicSpline(dim=3))
    for _ in range(12):
        s.step()
    f = s.dump_output(str(tmp_path), 'db')
    assert os.path.basename(f) == 'db_00012.npz'
    data = output.load(f)
    sd = data['solver_data']
    assert int(sd['count']) == 12 and float(sd['t']) == s.t
    assert 0.0 < s._damping_factor < 0.5
    assert abs(float(sd['dt']) - s.dt / s._damping_factor) <= 1e-14 * float(sd['dt'])
    assert data['arrays']['fluid'].get_number_of_particles() == pas[0].num_real_particles
    # only the output properties travelled (the loader adds the defaults back as zeros)
    assert set(pas[0].output_property_arrays) <= set(data['arrays']['fluid'].properties)
    assert 'au' not in data['arrays']['fluid'].properties or \
        not np.any(data['arrays']['fluid'].properties['au'])
    # the restarted run ...
    pas2 = geo.dam_break_3d_particles(dx=dx)
    s2 = pb.make_wcsph_solver(pas2, dict(params), pb.CubicSpline(dim=3))
    s2.load_output(f)
    assert (s2.count, s2.t, s2.dt) == (12, float(sd['t']), float(sd['dt']))
    # ... and the oracle, started from the same file the same way
    opas = [data['arrays'][pa.name] for pa in pas]
    for q, pa in zip(opas, pas):
        for k in pa.properties:
            if k not in q.properties:
                q.add_property(k)
    o = orc.WCSPHOracleSolver(opas, dict(params), 'CubicSpline', threads=4)
    o.t, o.dt, o.count = float(sd['t']), float(sd['dt']), 12
    o.initialise()
    s2.initialise()
    assert abs(s2.dt - o.dt) <= 1e-6 * o.dt
    for _ in range(8):
        s2.step()
        o.step()
    s2.pull()
    assert s2.count == 20 and abs(s2.t - o.t) <= 1e-5 * o.t
    h0, c0, rho0 = params['h0'], params['c0'], params['rho0']
    for a, b in zip(pas2, opas):
        n = b.get_number_of_particles()
        for k, tol in (('x', 2e-6 * h0), ('y', 2e-6 * h0), ('z', 2e-6 * h0), ('u', 2e-6 * c0),
                       ('v', 2e-6 * c0), ('w', 2e-6 * c0), ('rho', 2e-7 * rho0)):
            assert np.max(np.abs(a.properties[k][:n] - b.properties[k][:n])) <= tol, \
                (a.name, k)

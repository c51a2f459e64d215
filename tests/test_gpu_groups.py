"""Group features of the AccelerationEval loop nest on the device path: condition, iterate /
min_iterations / max_iterations, pre / post, start_idx / stop_idx.

Mirrors pysph/sph/tests/test_acceleration_eval.py:318-666 (the reference exercises them with
throw-away Python equations; here the equations are the ones that have CUDA kernels, on the
reference's own 1-D fixture of ten particles and on a golden 3-array WCSPH case)."""
import numpy as np
import pytest

from helpers import ACC_FIELDS, arrays_from_dict, load_golden, wcsph_params_from_case

pytestmark = pytest.mark.gpu


def fixture_1d(equations):
    import pysph_b200 as pb
    g = load_golden('density_1d.json')
    pa = pb.get_particle_array_wcsph(name='fluid', x=np.array(g['x']),
                                     h=np.array(g['h']), m=np.array(g['m']))
    kernel = pb.CubicSpline(dim=1)
    ae = pb.B200AccelerationEval([pa], equations, kernel)
    nn = pb.B200NNPS(1, [pa], backend=ae.backend, kernel=kernel)
    ae.set_nnps(nn)
    return pa, ae, np.array(g['rho'])


def set_rho(pa, ae, value):
    pa.rho[:] = value
    ae.backend.push_all()
    ae.nnps.update()


def test_group_honors_condition(gpu_device):
    # test_acceleration_eval.py:617-666
    import pysph_b200 as pb
    calls = []

    def cond(t, dt):
        calls.append((t, dt))
        return False

    sd = lambda: pb.SummationDensity(dest='fluid', sources=['fluid'])
    pa, ae, rho = fixture_1d([
        pb.Group(equations=[sd()], condition=cond),
        pb.Group(equations=[pb.Group(equations=[sd()], condition=cond)]),
    ])
    set_rho(pa, ae, 1.0)
    ae.compute(0.0, 0.1)
    ae.backend.pull_all(['rho'])
    assert calls == [(0.0, 0.1), (0.0, 0.1)]
    assert np.all(pa.rho == 1.0)            # neither group ran
    # ... and a condition that holds runs the group
    pa, ae, rho = fixture_1d([pb.Group(equations=[sd()], condition=lambda t, dt: t > 0.5)])
    set_rho(pa, ae, 1.0)
    ae.compute(0.0, 0.1)
    ae.backend.pull_all(['rho'])
    assert np.all(pa.rho == 1.0)
    ae.compute(1.0, 0.1)
    ae.backend.pull_all(['rho'])
    assert np.allclose(pa.rho, rho, rtol=2e-6)


def test_iterated_groups(gpu_device):
    # test_acceleration_eval.py:355-392 and acceleration_eval_cython_helper.py:320-340
    import pysph_b200 as pb
    runs = []

    def NeverConverged(**kw):          # equations are recognised by class name: patch an instance
        eq = pb.SummationDensity(**kw)
        eq.converged = lambda: -1.0
        return eq

    # converged() > 0 from the start: min_iterations decides
    pa, ae, rho = fixture_1d([pb.Group(
        equations=[pb.SummationDensity(dest='fluid', sources=['fluid'])],
        iterate=True, min_iterations=3, max_iterations=7, pre=lambda: runs.append('a'))])
    ae.compute(0.0, 0.1)
    assert runs == ['a'] * 3
    # never converged: max_iterations decides; iterate on the mother group repeats its sub-groups
    del runs[:]
    pa, ae, rho = fixture_1d([pb.Group(equations=[
        pb.Group(equations=[NeverConverged(dest='fluid', sources=['fluid'])],
                 pre=lambda: runs.append('s1')),
        pb.Group(equations=[pb.SummationDensity(dest='fluid', sources=['fluid'])],
                 pre=lambda: runs.append('s2'), iterate=True, max_iterations=9),
    ], iterate=True, max_iterations=4)])
    set_rho(pa, ae, 1.0)
    ae.compute(0.0, 0.1)
    assert runs == ['s1', 's2'] * 4        # the sub-group's own `iterate` is not looked at
    ae.backend.pull_all(['rho'])
    assert np.allclose(pa.rho, rho, rtol=2e-6)
    # a group that is not iterated runs once
    del runs[:]
    pa, ae, rho = fixture_1d([pb.Group(
        equations=[NeverConverged(dest='fluid', sources=['fluid'])],
        max_iterations=5, pre=lambda: runs.append('x'))])
    ae.compute(0.0, 0.1)
    assert runs == ['x']


def test_pre_post_order(gpu_device):
    # test_acceleration_eval.py:494-556: pre before anything of the group, post after everything,
    # for a plain group and for a mother group
    import pysph_b200 as pb
    log = []
    sd = lambda: pb.SummationDensity(dest='fluid', sources=['fluid'])
    pa, ae, rho = fixture_1d([
        pb.Group(equations=[sd()], pre=lambda: log.append('pre1'), post=lambda: log.append('post1')),
        pb.Group(equations=[
            pb.Group(equations=[sd()], pre=lambda: log.append('pre_sub'),
                     post=lambda: log.append('post_sub'))],
            pre=lambda: log.append('pre2'), post=lambda: log.append('post2')),
    ])

    # pre / post see and may change the particle arrays (the reference's tests do): a pre that
    # doubles the masses on the device doubles the density the group computes
    def double_mass():
        ae.backend.pull_all(['m'])
        pa.m *= 2.0
        ae.backend.push_all()
        ae.nnps.update()
    ae.compute(0.0, 0.1)
    assert log == ['pre1', 'post1', 'pre2', 'pre_sub', 'post_sub', 'post2']
    pa2, ae2, rho = fixture_1d([pb.Group(equations=[sd()])])
    ae2.ops.insert(0, ('call', lambda: None))
    ae = ae2
    pa = pa2
    ae.ops[0] = ('call', double_mass)
    ae.compute(0.0, 0.1)
    ae.backend.pull_all(['rho'])
    assert np.allclose(pa.rho, 2.0 * rho, rtol=2e-6)


@pytest.mark.parametrize('as_str', [False, True])
def test_start_stop_idx(gpu_device, as_str):
    # test_acceleration_eval.py:558-615
    import pysph_b200 as pb
    g = load_golden('density_1d.json')
    pa = pb.get_particle_array_wcsph(name='fluid', x=np.array(g['x']),
                                     h=np.array(g['h']), m=np.array(g['m']))
    if as_str:
        pa.add_constant('start', 1)
        pa.add_constant('stop', 3)
        kw = dict(start_idx='start', stop_idx='stop')
        lo, hi = 1, 3
    else:
        kw = dict(start_idx=1, stop_idx=2)
        lo, hi = 1, 2
    kernel = pb.CubicSpline(dim=1)
    ae = pb.B200AccelerationEval(
        [pa], [pb.Group(equations=[pb.SummationDensity(dest='fluid', sources=['fluid'])], **kw)],
        kernel)
    nn = pb.B200NNPS(1, [pa], backend=ae.backend, kernel=kernel)
    ae.set_nnps(nn)
    set_rho(pa, ae, 1.0)
    ae.count_pairs = True
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['rho'])
    expect = np.ones(10)
    expect[lo:hi] = np.array(g['rho'])[lo:hi]
    assert np.allclose(pa.rho, expect, rtol=2e-6)
    assert np.all(pa.rho[:lo] == 1.0) and np.all(pa.rho[hi:] == 1.0)
    assert ae.last_pairs == sum(g['nbr_counts'][lo:hi])
    if as_str:
        # the named constant is read at every compute (helper:265-278)
        pa.constants['stop'][0] = 5
        set_rho(pa, ae, 1.0)
        ae.compute(0.1, 0.1)
        ae.backend.pull_all(['rho'])
        assert np.allclose(pa.rho[1:5], np.array(g['rho'])[1:5], rtol=2e-6)
        assert np.all(pa.rho[5:] == 1.0) and pa.rho[0] == 1.0
    # the range is one-shot: an unrestricted group afterwards touches everything
    ae2 = pb.B200AccelerationEval(
        [pa], [pb.SummationDensity(dest='fluid', sources=['fluid'])], kernel, backend=ae.backend)
    ae2.set_nnps(nn)
    nn.update()
    ae2.compute(0.1, 0.1)
    ae.backend.pull_all(['rho'])
    assert np.allclose(pa.rho, g['rho'], rtol=2e-6)


def test_start_stop_idx_three_arrays(gpu_device):
    """The WCSPH Group (continuity + momentum + XSPH, three arrays) restricted to a window of
    destinations: inside the window == the unrestricted evaluation, outside untouched."""
    import pysph_b200 as pb
    case = load_golden('wcsph_cases.json')[0]
    p = wcsph_params_from_case(case)
    kernel = getattr(pb, case['kernel'])(dim=p['dim'])

    def scheme_groups(**kw):
        s = pb.WCSPHScheme(p['fluids'], p['solids'], dim=p['dim'], rho0=p['rho0'], c0=p['c0'],
                           h0=p['h0'], hdx=p['hdx'], gamma=p['gamma'], gx=p.get('gx', 0.0),
                           gy=p.get('gy', 0.0), gz=p.get('gz', 0.0), alpha=p['alpha'],
                           beta=p['beta'], tensile_correction=p.get('tensile_correction', False),
                           hg_correction=p.get('hg_correction', False),
                           update_h=False, summation_density=False)
        groups = s.get_equations()
        if kw:
            last = groups[-1]
            groups[-1] = pb.Group(equations=last.equations, real=last.real, **kw)
        return groups

    def run(groups, sentinel):
        pas = arrays_from_dict(case['inputs'])
        ae = pb.B200AccelerationEval(pas, groups, kernel)
        nn = pb.B200NNPS(p['dim'], pas, backend=ae.backend, kernel=kernel)
        ae.set_nnps(nn)
        for pa in pas:
            for f in ACC_FIELDS:
                pa.properties[f][:] = sentinel
        ae.backend.push_all()
        nn.update()
        ae.compute(0.0, 0.0)
        ae.backend.pull_all()
        return pas

    full = run(scheme_groups(), 0.0)
    lo, hi = 3, 11
    part = run(scheme_groups(start_idx=lo, stop_idx=hi), 7.0)
    for a, b in zip(full, part):
        n = a.get_number_of_particles()
        for f in ('arho', 'au', 'av', 'aw', 'ax', 'ay', 'az'):
            fa, fb = a.properties[f], b.properties[f]
            scale = max(np.max(np.abs(fa)), 1e-30)
            w = slice(lo, min(hi, n))
            if np.max(np.abs(fa)) > 0.0:
                assert np.max(np.abs(fa[w] - fb[w])) <= 2e-6 * scale, (a.name, f)
            outside = np.r_[fb[:lo], fb[hi:]]
            # destinations outside the window keep what they held
            touched = np.max(np.abs(fa)) > 0.0
            if touched:
                assert np.all(outside == 7.0), (a.name, f)

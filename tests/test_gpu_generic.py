"""Generic-equation fallback: user Equation bodies translated (pysph_b200/codegen.py), compiled
at run time (NVRTC on the GPU box; for the host when the emulated library is loaded) and run
over the persistent neighbour lists.

The equations and the expected numbers are the reference's own tests of its code generator on
ten particles in 1-D (pysph/sph/tests/test_acceleration_eval.py:138-236, 294-666, 668-675);
the last tests check the pre-computed pair symbols against the hand-written kernels on the
golden 3-array WCSPH case."""
import numpy as np
import pytest

from helpers import arrays_from_dict, load_golden, wcsph_params_from_case

pytestmark = pytest.mark.gpu

EXPECT = np.asarray([3., 4., 5., 5., 5., 5., 5., 5., 4., 3.])


def _equations():
    import pysph_b200 as pb

    class SimpleEquation(pb.Equation):            # test_acceleration_eval.py:138-159
        def __init__(self, dest, sources):
            super(SimpleEquation, self).__init__(dest, sources)
            self.count = 0

        def initialize(self, d_idx, d_u, d_au):
            d_u[d_idx] = 0.0
            d_au[d_idx] = 0.0

        def loop(self, d_idx, d_au, s_idx, s_m):
            d_au[d_idx] += s_m[s_idx]

        def post_loop(self, d_idx, d_u, d_au):
            d_u[d_idx] = d_au[d_idx]

        def converged(self):
            self.count += 1
            result = self.count - 1
            if result > 0:
                self.count = 0
            return result

    class MixedTypeEquation(pb.Equation):         # :162-172
        def initialize(self, d_idx, d_u, d_au, d_pid, d_tag):
            d_u[d_idx] = 0.0 + d_pid[d_idx]
            d_au[d_idx] = 0.0 + d_tag[d_idx]

        def loop(self, d_idx, d_au, s_idx, s_m, s_pid, s_tag):
            d_au[d_idx] += s_m[s_idx] + s_pid[s_idx] + s_tag[s_idx]

        def post_loop(self, d_idx, d_u, d_au, d_pid):
            d_u[d_idx] = d_au[d_idx] + d_pid[d_idx]

    class DumbEquation(pb.Equation):              # :221-230 (without its reduce)
        def initialize(self, d_idx, d_au):
            d_au[d_idx] += 1

        def loop(self, d_idx, d_au):
            d_au[d_idx] += 1

        def post_loop(self, d_idx, d_au):
            d_au[d_idx] += 1

    class EqWithTime(pb.Equation):                # :668-673
        def initialize(self, d_idx, d_au, t, dt):
            d_au[d_idx] = t + dt

        def loop(self, d_idx, d_au, s_idx, s_m, t, dt):
            d_au[d_idx] += t + dt

    class KernelSum(pb.Equation):
        """a property the device pool does not have + WIJ"""
        def initialize(self, d_idx, d_wsum):
            d_wsum[d_idx] = 0.0

        def loop(self, d_idx, d_wsum, s_idx, s_m, WIJ):
            d_wsum[d_idx] += s_m[s_idx] * WIJ

    return SimpleEquation, MixedTypeEquation, DumbEquation, EqWithTime, KernelSum


def make(equations, extra_props=()):
    import pysph_b200 as pb
    g = load_golden('density_1d.json')
    pa = pb.get_particle_array_wcsph(name='fluid', x=np.array(g['x']), h=np.array(g['h']),
                                     m=np.array(g['m']))
    for name in extra_props:
        pa.add_property(name)
    kernel = pb.CubicSpline(dim=1)
    ae = pb.B200AccelerationEval([pa], equations, kernel)
    nn = pb.B200NNPS(1, [pa], backend=ae.backend, kernel=kernel)
    ae.set_nnps(nn)
    return pa, ae, g


def test_simple_equation(gpu_device):
    # test_should_not_iterate_normal_group / test_accel_eval_should_work_on_gpu (:331-341, :719-731)
    SimpleEquation = _equations()[0]
    pa, ae, g = make([SimpleEquation(dest='fluid', sources=['fluid'])])
    assert ae.ops[0][0] == 'generic'
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u', 'au'])
    assert list(pa.u) == list(EXPECT) and list(pa.au) == list(EXPECT)


def test_iterated_generic_groups(gpu_device):
    # test_should_iterate_iterated_group / test_should_iterate_nested_groups (:355-392): the
    # equations' own converged() ends the iteration after the second pass
    import pysph_b200 as pb
    SimpleEquation = _equations()[0]
    pa, ae, g = make([pb.Group(equations=[SimpleEquation(dest='fluid', sources=['fluid']),
                                          SimpleEquation(dest='fluid', sources=['fluid'])],
                               iterate=True, max_iterations=10)])
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u'])
    assert list(pa.u) == list(EXPECT * 2)
    pa, ae, g = make([pb.Group(equations=[
        pb.Group(equations=[SimpleEquation(dest='fluid', sources=['fluid'])]),
        pb.Group(equations=[SimpleEquation(dest='fluid', sources=['fluid'])])],
        iterate=True, max_iterations=10)])
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u'])
    assert list(pa.u) == list(EXPECT)


def test_mixed_type_arrays_and_time(gpu_device):
    # test_should_work_with_non_double_arrays (:447-459); EqWithTime (:668-675, :803-815)
    _, MixedTypeEquation, _, EqWithTime, _ = _equations()
    pa, ae, g = make([MixedTypeEquation(dest='fluid', sources=['fluid'])])
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u'])
    assert list(pa.u) == list(EXPECT)
    pa, ae, g = make([EqWithTime(dest='fluid', sources=['fluid'])])
    ae.compute(0.25, 0.5)
    ae.backend.pull_all(['au'])
    assert np.allclose(pa.au, 0.75 * (1.0 + EXPECT), rtol=1e-7)


def test_generic_group_controls(gpu_device):
    # condition (:617-666), start_idx / stop_idx (:558-615), pre / post (:494-531)
    import pysph_b200 as pb
    SimpleEquation, _, DumbEquation, _, _ = _equations()
    calls = []

    def cond(t, dt):
        calls.append((t, dt))
        return False
    pa, ae, g = make([
        pb.Group(equations=[DumbEquation(dest='fluid', sources=['fluid'])], condition=cond),
        pb.Group(equations=[pb.Group(equations=[DumbEquation(dest='fluid', sources=['fluid'])],
                                     condition=cond)]),
        pb.Group(equations=[DumbEquation(dest='fluid', sources=['fluid'])])])
    ae.compute(0.0, 0.1)
    ae.backend.pull_all(['au'])
    expect = np.ones(10) * 7
    expect[0] = expect[-1] = 5
    expect[1] = expect[-2] = 6
    assert calls == [(0.0, 0.1), (0.0, 0.1)] and list(pa.au) == list(expect)

    pa, ae, g = make([pb.Group(equations=[SimpleEquation(dest='fluid', sources=['fluid'])],
                               start_idx=1, stop_idx=2)])
    pa.u[:] = 1.0
    pa.au[:] = 1.0
    ae.backend.push_all()
    ae.nnps.update()
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u', 'au'])
    expect = np.ones(10)
    expect[1] = 4.0
    assert list(pa.u) == list(expect) and list(pa.au) == list(expect)

    def pre():
        ae.backend.pull_all(['m'])
        pa.m += 1.0
        ae.backend.push_all()
        ae.nnps.update()

    def post():
        ae.backend.pull_all(['u'])
        pa.u += 1.0
        ae.backend.push_all()
        ae.nnps.update()
    pa, ae, g = make([pb.Group(equations=[SimpleEquation(dest='fluid', sources=['fluid'])],
                               pre=pre, post=post)])
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['u'])
    assert list(pa.u) == list(2.0 * EXPECT + 1.0)       # [7, 9, 11, ...]


def test_user_property_and_kernel_symbol(gpu_device):
    KernelSum = _equations()[4]
    pa, ae, g = make([KernelSum(dest='fluid', sources=['fluid'])], extra_props=['wsum'])
    assert ae.backend.user_props == ['wsum']
    ae.compute(0.0, 0.1)
    ae.backend.pull_all(['wsum'])
    # SummationDensity of the reference's fixture (test_acceleration_eval.py:294-303, :341)
    assert np.allclose(pa.wsum, g['rho'], rtol=1e-6)
    assert np.allclose(pa.wsum, [7.357] + [9.0] * 8 + [7.357], atol=1e-2)


def test_untranslatable_bodies_fail_at_setup(gpu_device):
    import pysph_b200 as pb

    class WithReduce(pb.Equation):              # (loop_all: the neighbour array itself)
        def initialize(self, d_idx, d_au):
            d_au[d_idx] = 0.0

        def loop_all(self, d_idx, d_au, NBRS, N_NBRS):
            pass

    class WithWhile(pb.Equation):
        def initialize(self, d_idx, d_au):
            while d_au[d_idx] < 3.0:
                d_au[d_idx] += 1.0

    class WritesSource(pb.Equation):
        def loop(self, d_idx, s_idx, s_au):
            s_au[s_idx] = 1.0

    for cls in (WithReduce, WithWhile, WritesSource):
        with pytest.raises(NotImplementedError):
            make([cls(dest='fluid', sources=['fluid'])])


def _generic_wcsph():
    """Continuity + Momentum (no tensile correction) + XSPH written as user equations: the
    bodies of basic_equations.py:180-192, wc/basic.py:204-269, basic_equations.py:285-300."""
    import pysph_b200 as pb

    class MyContinuity(pb.Equation):
        def initialize(self, d_idx, d_arho):
            d_arho[d_idx] = 0.0

        def loop(self, d_idx, d_arho, s_idx, s_m, DWIJ, VIJ):
            vijdotdwij = DWIJ[0] * VIJ[0] + DWIJ[1] * VIJ[1] + DWIJ[2] * VIJ[2]
            d_arho[d_idx] += s_m[s_idx] * vijdotdwij

    class MyMomentum(pb.Equation):
        def __init__(self, dest, sources, c0, alpha, beta, gx, gy, gz):
            self.c0, self.alpha, self.beta = c0, alpha, beta
            self.gx, self.gy, self.gz = gx, gy, gz
            super(MyMomentum, self).__init__(dest, sources)

        def initialize(self, d_idx, d_au, d_av, d_aw, d_dt_cfl):
            d_au[d_idx] = 0.0
            d_av[d_idx] = 0.0
            d_aw[d_idx] = 0.0
            d_dt_cfl[d_idx] = 0.0

        def loop(self, d_idx, s_idx, d_rho, d_cs, d_p, d_au, d_av, d_aw, s_m, s_rho, s_cs, s_p,
                 VIJ, XIJ, HIJ, R2IJ, RHOIJ1, EPS, DWIJ, d_dt_cfl):
            rhoi21 = 1.0 / (d_rho[d_idx] * d_rho[d_idx])
            rhoj21 = 1.0 / (s_rho[s_idx] * s_rho[s_idx])
            vijdotxij = VIJ[0] * XIJ[0] + VIJ[1] * XIJ[1] + VIJ[2] * XIJ[2]
            piij = 0.0
            if vijdotxij < 0:
                cij = 0.5 * (d_cs[d_idx] + s_cs[s_idx])
                muij = (HIJ * vijdotxij) / (R2IJ + EPS)
                piij = -self.alpha * cij * muij + self.beta * muij * muij
                piij = piij * RHOIJ1
            _dt_cfl = 0.0
            if R2IJ > 1e-12:
                _dt_cfl = abs(HIJ * vijdotxij / R2IJ) + self.c0
                d_dt_cfl[d_idx] = max(_dt_cfl, d_dt_cfl[d_idx])
            tmp = d_p[d_idx] * rhoi21 + s_p[s_idx] * rhoj21
            d_au[d_idx] += -s_m[s_idx] * (tmp + piij) * DWIJ[0]
            d_av[d_idx] += -s_m[s_idx] * (tmp + piij) * DWIJ[1]
            d_aw[d_idx] += -s_m[s_idx] * (tmp + piij) * DWIJ[2]

        def post_loop(self, d_idx, d_au, d_av, d_aw, d_dt_force):
            d_au[d_idx] += self.gx
            d_av[d_idx] += self.gy
            d_aw[d_idx] += self.gz
            acc2 = d_au[d_idx] * d_au[d_idx] + d_av[d_idx] * d_av[d_idx] + d_aw[d_idx] * d_aw[d_idx]
            d_dt_force[d_idx] = acc2

    class MyXSPH(pb.Equation):
        def __init__(self, dest, sources, eps):
            self.eps = eps
            super(MyXSPH, self).__init__(dest, sources)

        def initialize(self, d_idx, d_ax, d_ay, d_az):
            d_ax[d_idx] = 0.0
            d_ay[d_idx] = 0.0
            d_az[d_idx] = 0.0

        def loop(self, s_idx, d_idx, s_m, d_ax, d_ay, d_az, WIJ, RHOIJ1, VIJ):
            tmp = -self.eps * s_m[s_idx] * WIJ * RHOIJ1
            d_ax[d_idx] += tmp * VIJ[0]
            d_ay[d_idx] += tmp * VIJ[1]
            d_az[d_idx] += tmp * VIJ[2]

        def post_loop(self, d_idx, d_ax, d_ay, d_az, d_u, d_v, d_w):
            d_ax[d_idx] += d_u[d_idx]
            d_ay[d_idx] += d_v[d_idx]
            d_az[d_idx] += d_w[d_idx]

    return MyContinuity, MyMomentum, MyXSPH


@pytest.mark.parametrize('idx', [0, 1, 2, 5])
def test_generic_wcsph_group_equals_golden(gpu_device, idx):
    """The WCSPH Group written as USER equations (every pre-computed pair symbol the reference
    bodies use: XIJ VIJ R2IJ HIJ RHOIJ1 EPS WIJ DWIJ), three arrays, against the outputs the
    reference's own bodies produced (tests/golden/wcsph_cases.json).  fp64 bodies on fp32
    cell-relative positions: tolerance 2e-5 of each field's largest magnitude."""
    import pysph_b200 as pb
    case = load_golden('wcsph_cases.json')[idx]
    p = wcsph_params_from_case(case)
    if p.get('tensile_correction') or p.get('summation_density'):
        pytest.skip('the generic restatement has no tensile correction / summation density')
    MyContinuity, MyMomentum, MyXSPH = _generic_wcsph()
    kernel = getattr(pb, case['kernel'])(dim=p['dim'])
    solids = p['solids']
    # the scheme's own Groups (scheme.py:414-483), its pair equations replaced one by one
    scheme = pb.WCSPHScheme(p['fluids'], solids, dim=p['dim'], rho0=p['rho0'], c0=p['c0'],
                            h0=p['h0'], hdx=p['hdx'], gamma=p['gamma'], gx=p.get('gx', 0.0),
                            gy=p.get('gy', 0.0), gz=p.get('gz', 0.0), alpha=p['alpha'],
                            beta=p['beta'], tensile_correction=False,
                            hg_correction=p.get('hg_correction', False), update_h=False)
    groups = scheme.get_equations()
    eos, g2 = groups[0].equations, []
    for e in groups[1].equations:
        name = e.__class__.__name__
        if name == 'ContinuityEquation':
            g2.append(MyContinuity(dest=e.dest, sources=e.sources))
        elif name == 'MomentumEquation':
            g2.append(MyMomentum(dest=e.dest, sources=e.sources, c0=e.c0, alpha=e.alpha,
                                 beta=e.beta, gx=e.gx, gy=e.gy, gz=e.gz))
        else:
            assert name == 'XSPHCorrection'
            g2.append(MyXSPH(dest=e.dest, sources=e.sources, eps=e.eps))
    pas = arrays_from_dict(case['inputs'])
    ae = pb.B200AccelerationEval(pas, [pb.Group(equations=eos, real=False),
                                       pb.Group(equations=g2)], kernel)
    assert [o[0] for o in ae.ops][-1] == 'generic'
    nn = pb.B200NNPS(p['dim'], pas, backend=ae.backend, kernel=kernel)
    ae.set_nnps(nn)
    ae.compute(0.0, 0.0)
    ae.backend.pull_all()
    for pa in pas:
        ref = case['outputs'][pa.name]
        nr = ref['_n_real']
        fields = ['arho'] if pa.name in solids else \
            ['arho', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'dt_cfl', 'dt_force']
        for f in fields:
            want = np.array(ref[f])[:nr]
            got = pa.properties[f][:nr]
            scale = np.max(np.abs(want))
            if scale == 0.0:
                assert np.max(np.abs(got)) == 0.0, (pa.name, f)
                continue
            assert np.max(np.abs(got - want)) <= 2e-5 * scale, (pa.name, f, np.max(np.abs(got - want)) / scale)


def test_generic_kernel_symbols_vs_reference_kernels(gpu_device):
    """WIJ and DWIJ of the generated code, every kernel and dimension, on two particles against
    the values of the reference's COMPILED kernels (tests/golden/kernels.json)."""
    import pysph_b200 as pb

    class Probe(pb.Equation):
        def initialize(self, d_idx, d_wsum, d_gx, d_gy, d_gz):
            d_wsum[d_idx] = 0.0
            d_gx[d_idx] = 0.0
            d_gy[d_idx] = 0.0
            d_gz[d_idx] = 0.0

        def loop(self, d_idx, s_idx, s_m, d_wsum, d_gx, d_gy, d_gz, WIJ, DWIJ):
            d_wsum[d_idx] += s_m[s_idx] * WIJ
            d_gx[d_idx] += s_m[s_idx] * DWIJ[0]
            d_gy[d_idx] += s_m[s_idx] * DWIJ[1]
            d_gz[d_idx] += s_m[s_idx] * DWIJ[2]

    checked = 0
    for entry in load_golden('kernels.json'):
        name, dim = entry['kernel'], entry['dim']
        kernel = getattr(pb, name)(dim=dim)
        for c in entry['cases'][1::4]:
            if c['rij'] >= kernel.radius_scale * c['h'] or c['rij'] < 1e-9:
                continue
            pa = pb.get_particle_array_wcsph(
                name='f', x=np.array([c['xij'][0], 0.0]), y=np.array([c['xij'][1], 0.0]),
                z=np.array([c['xij'][2], 0.0]), h=np.full(2, c['h']), m=np.array([0.0, 1.0]))
            for k in ('wsum', 'gx', 'gy', 'gz'):
                pa.add_property(k)
            ae = pb.B200AccelerationEval([pa], [Probe(dest='f', sources=['f'])], kernel)
            nn = pb.B200NNPS(dim, [pa], backend=ae.backend, kernel=kernel)
            ae.set_nnps(nn)
            ae.compute(0.0, 0.0)
            ae.backend.pull_all(['wsum', 'gx', 'gy', 'gz'])
            scale = entry['fac'] / c['h'] ** dim
            # positions reach the kernel as fp32 cell-relative coordinates
            assert abs(pa.wsum[0] - c['w']) <= 3e-6 * max(scale, abs(c['w'])), (name, dim, c)
            g = np.array([pa.gx[0], pa.gy[0], pa.gz[0]])
            assert np.max(np.abs(g - np.array(c['grad']))) <= 3e-6 * max(scale / c['h'], np.max(np.abs(c['grad']))), (name, dim, c)
            checked += 1
    assert checked >= 20


def test_host_callbacks_and_helpers(gpu_device):
    # test_should_call_py_initialize (:425-445), test_should_run_reduce (:394-406),
    # test_should_handle_repeated_helper_functions (:482-508)
    import pysph_b200 as pb
    from pysph_b200.reduce_array import serial_reduce_array

    class PyInit(pb.Equation):
        def py_initialize(self, dst, t, dt):
            self.called_with = t, dt
            if dst.gpu:
                dst.gpu.pull('au')
            dst.au[:] = 1.0
            if dst.gpu:
                dst.gpu.push('au')

        def initialize(self, d_idx, d_au):
            d_au[d_idx] += 1.0

    class SimpleReduction(pb.Equation):
        def initialize(self, d_idx, d_au):
            d_au[d_idx] = 0.0

        def reduce(self, dst, t, dt):
            dst.gpu.pull('m')
            dst.total_mass[0] = serial_reduce_array(dst.m, op='sum')

    def helper(x=1.0):
        return x * 1.5

    class SillyEquation2(pb.Equation):
        def initialize(self, d_idx, d_au, d_m):
            d_au[d_idx] += helper(d_m[d_idx])

        def _get_helpers_(self):
            return [helper]

    eq = PyInit(dest='fluid', sources=None)
    pa, ae, g = make([eq])
    ae.compute(1.0, 0.1)
    ae.backend.pull_all(['au'])
    assert np.all(pa.au == 2.0) and eq.called_with == (1.0, 0.1)

    pa, ae, g = make([SimpleReduction(dest='fluid', sources=['fluid'])])
    pa.add_constant('total_mass', 0.0)
    ae.compute(0.1, 0.1)
    assert abs(pa.total_mass[0] - np.sum(pa.m)) < 1e-14

    pa, ae, g = make([SillyEquation2(dest='fluid', sources=['fluid']),
                      SillyEquation2(dest='fluid', sources=['fluid'])])
    pa.au[:] = 0.0
    ae.backend.push_all()
    ae.nnps.update()
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['au'])
    assert list(pa.au) == [3.0] * 10


def test_mixed_and_aliased_groups(gpu_device):
    """A Group that mixes an equation the library knows BY NAME with an unknown one is translated
    as a whole when every equation carries its Python body (the reference's objects do); a
    generated kernel may not touch a property that a hand-written scheme keeps elsewhere."""
    import pysph_b200 as pb
    SimpleEquation = _equations()[0]

    class SummationDensity(pb.Equation):          # a name the library has a kernel for, with a body
        def initialize(self, d_idx, d_rho):
            d_rho[d_idx] = 0.0

        def loop(self, d_idx, d_rho, s_idx, s_m, WIJ):
            d_rho[d_idx] += s_m[s_idx] * WIJ

    pa, ae, g = make([pb.Group(equations=[SummationDensity(dest='fluid', sources=['fluid']),
                                          SimpleEquation(dest='fluid', sources=['fluid'])])])
    assert [o[0] for o in ae.ops] == ['generic']
    ae.compute(0.1, 0.1)
    ae.backend.pull_all(['rho', 'u'])
    assert np.allclose(pa.rho, g['rho'], rtol=1e-6) and list(pa.u) == list(EXPECT)
    # this package's descriptor of the same equation has no body: the mix is refused
    with pytest.raises(NotImplementedError):
        make([pb.Group(equations=[pb.SummationDensity(dest='fluid', sources=['fluid']),
                                  SimpleEquation(dest='fluid', sources=['fluid'])])])

    class TouchesPressure(pb.Equation):
        def initialize(self, d_idx, d_p):
            d_p[d_idx] = 1.0

    # on an EDAC fluid p is the EVOLVED pressure (a different device array): refused at set-up
    n = 10
    fl = pb.get_particle_array_edac(name='fluid', x=np.linspace(0, 1, n), h=np.full(n, 0.12),
                                    m=np.ones(n))
    with pytest.raises(NotImplementedError):
        pb.B200AccelerationEval([fl], [TouchesPressure(dest='fluid', sources=None)],
                                pb.CubicSpline(dim=1))

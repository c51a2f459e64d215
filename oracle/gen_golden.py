"""Generate the golden fixtures in tests/golden/ FROM THE REFERENCE ITSELF.

TEST INFRASTRUCTURE.  Runs only in the build container (needs /root/reference);
the fixtures it writes are committed so that the tests can run on the GPU box,
where the reference does not exist.

What is executed here is the reference's own code, unmodified, loaded by path:

* pysph/base/kernels.py                      (pure Python kernels)
* oracle/_ref/c_kernels*.so                  (the reference's compiled kernels,
                                              built by oracle/build_ref.py)
* pysph/sph/wc/basic.py, pysph/sph/basic_equations.py  -- the equation classes,
  with ``pysph.sph.equation`` stubbed by a 6-line ``Equation`` base (the real
  module imports compyle/mako which are not installed) and ``pow/abs/max/sqrt``
  injected (the transpiler normally supplies them)
* pysph/sph/integrator_step.py               (WCSPHStep)

Only the loop DRIVER is ours (brute-force neighbours with the criterion of
pysph/base/nnps_base.pyx:1365, precomputed symbols per
pysph/sph/equation.py:188-297, loop nest per acceleration_eval_cython.mako).

    python oracle/gen_golden.py          # rewrites tests/golden/*.json
"""
import importlib.util
import inspect
import json
import math
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('PYSPH_REFERENCE', '/root/reference')
GOLD = os.path.join(ROOT, 'tests', 'golden')


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    stub = types.ModuleType('pysph.sph.equation')

    class Equation(object):
        def __init__(self, dest, sources):
            self.dest = dest
            self.sources = sources
            self.no_source = sources is None

    class Group(object):
        def __init__(self, equations, real=True, **kw):
            self.equations = equations
            self.real = real

    stub.Equation = Equation
    stub.Group = Group
    for pkg in ('pysph', 'pysph.sph', 'pysph.sph.wc', 'pysph.base'):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    sys.modules['pysph.sph.equation'] = stub
    # what wc/edac.py imports besides the equations: a Scheme base and two helpers
    # it only uses in methods that are not called here
    utils = types.ModuleType('pysph.base.utils')
    utils.get_particle_array = None
    utils.DEFAULT_PROPS = set()
    sys.modules['pysph.base.utils'] = utils
    scheme = types.ModuleType('pysph.sph.scheme')
    scheme.Scheme = type('Scheme', (object,), {})
    scheme.add_bool_argument = None
    sys.modules['pysph.sph.scheme'] = scheme
    kernels = _load('pysph.base.kernels',
                    os.path.join(REF, 'pysph/base/kernels.py'))
    basic = _load('pysph.sph.basic_equations',
                  os.path.join(REF, 'pysph/sph/basic_equations.py'))
    wc = _load('pysph.sph.wc.basic', os.path.join(REF, 'pysph/sph/wc/basic.py'))
    steps = _load('pysph.sph.integrator_step',
                  os.path.join(REF, 'pysph/sph/integrator_step.py'))
    for m in (basic, wc, steps):
        m.pow, m.abs, m.max, m.sqrt = pow, abs, max, math.sqrt
    sys.path.insert(0, os.path.join(HERE, '_ref'))
    import c_kernels
    return kernels, basic, wc, steps, c_kernels


def load_reference_edac():
    """wc/transport_velocity.py and wc/edac.py, unmodified (after load_reference)."""
    tvf = _load('pysph.sph.wc.transport_velocity',
                os.path.join(REF, 'pysph/sph/wc/transport_velocity.py'))
    edac = _load('pysph.sph.wc.edac', os.path.join(REF, 'pysph/sph/wc/edac.py'))
    return tvf, edac


def call(method, env):
    """Call an equation method with the arguments its signature asks for."""
    names = [a for a in inspect.signature(method).parameters]
    return method(*[env[a] for a in names])


# ---------------------------------------------------------------------------
def gen_kernels(kernels, c_kernels):
    out = []
    rs = np.random.RandomState(7)
    combos = [('CubicSpline', (1, 2, 3)), ('WendlandQuintic', (2, 3)),
              ('QuinticSpline', (1, 2, 3)), ('Gaussian', (1, 2, 3))]
    for name, dims in combos:
        for dim in dims:
            k = getattr(kernels, name)(dim=dim)
            ck = getattr(c_kernels, name)(**k.__dict__)
            wrap = getattr(c_kernels, name + 'Wrapper')(ck)
            cases = []
            h_vals = [1.0, 0.0114, 0.039, 0.5]
            q_vals = [0.0, 1e-13, 0.25, 0.5, 0.999, 1.0, 1.001, 1.5, 1.999, 2.0,
                      2.001, 2.5, 2.999, 3.0, 3.5] + list(rs.uniform(0, 3.2, 8))
            for h in h_vals:
                for q in q_vals:
                    d = rs.normal(size=3)
                    if dim < 3:
                        d[2] = 0.0
                    if dim < 2:
                        d[1] = 0.0
                    d = d / np.linalg.norm(d) * q * h
                    rij = float(np.sqrt(np.dot(d, d)))
                    grad = [0.0, 0.0, 0.0]
                    k.gradient(list(d), rij, h, grad)
                    w_py = k.kernel(list(d), rij, h)
                    w_c = wrap.kernel(d[0], d[1], d[2], 0.0, 0.0, 0.0, h)
                    g_c = wrap.gradient(d[0], d[1], d[2], 0.0, 0.0, 0.0, h)
                    # the reference's compiled and Python kernels agree
                    assert abs(w_py - w_c) <= 1e-12 * max(1.0, abs(w_py)), (name, dim, q)
                    assert np.allclose(grad, g_c, rtol=1e-11,
                                       atol=1e-12 * k.fac / h ** (dim + 1)), (name, dim, q)
                    cases.append(dict(xij=list(map(float, d)), rij=rij, h=h,
                                      w=float(w_c), grad=list(map(float, g_c))))
            out.append(dict(kernel=name, dim=dim, fac=k.fac,
                            radius_scale=k.radius_scale,
                            deltap=k.get_deltap(), cases=cases))
    return out


def pair_symbols(kernel, d, s, di, si):
    XIJ = [d['x'][di] - s['x'][si], d['y'][di] - s['y'][si],
           d['z'][di] - s['z'][si]]
    VIJ = [d['u'][di] - s['u'][si], d['v'][di] - s['v'][si],
           d['w'][di] - s['w'][si]]
    R2IJ = XIJ[0] * XIJ[0] + XIJ[1] * XIJ[1] + XIJ[2] * XIJ[2]
    RIJ = math.sqrt(R2IJ)
    HIJ = 0.5 * (d['h'][di] + s['h'][si])
    EPS = 0.01 * HIJ * HIJ
    RHOIJ = 0.5 * (d['rho'][di] + s['rho'][si])
    # only requested by equations that read rho; guard the unused 0/0 case
    RHOIJ1 = 1.0 / RHOIJ if RHOIJ != 0.0 else float('inf')
    WIJ = kernel.kernel(XIJ, RIJ, HIJ)
    DWIJ = [0.0, 0.0, 0.0]
    kernel.gradient(XIJ, RIJ, HIJ, DWIJ)
    WDP = kernel.kernel(XIJ, kernel.get_deltap() * HIJ, HIJ)
    return dict(XIJ=XIJ, VIJ=VIJ, R2IJ=R2IJ, RIJ=RIJ, HIJ=HIJ, EPS=EPS,
                RHOIJ=RHOIJ, RHOIJ1=RHOIJ1, WIJ=WIJ, DWIJ=DWIJ, WDP=WDP)


def evaluate_reference(kernel, arrays, groups, t=0.0):
    """Our driver (mako:10-154) around the reference's loop bodies.
    arrays: dict name -> dict prop -> list;  groups: list of (real, [eq])."""
    k2 = kernel.radius_scale
    for real, eqs in groups:
        dests = []
        for e in eqs:
            if e.dest not in dests:
                dests.append(e.dest)
        for dname in dests:
            d = arrays[dname]
            npd = d['_n_real'] if real else len(d['x'])
            deqs = [e for e in eqs if e.dest == dname]
            denv = dict(('d_' + k, v) for k, v in d.items() if k[0] != '_')
            denv['t'] = t
            for di in range(npd):
                for e in deqs:
                    if hasattr(e, 'initialize'):
                        call(e.initialize, dict(denv, d_idx=di))
            for e in deqs:
                if e.sources is None and hasattr(e, 'loop'):
                    for di in range(npd):
                        call(e.loop, dict(denv, d_idx=di))
            srcs = []
            for e in deqs:
                for s in (e.sources or []):
                    if s not in srcs:
                        srcs.append(s)
            for sname in srcs:
                s = arrays[sname]
                senv = dict(('s_' + k, v) for k, v in s.items() if k[0] != '_')
                seqs = [e for e in deqs if e.sources and sname in e.sources]
                for di in range(npd):
                    hi = k2 * d['h'][di]
                    for si in range(len(s['x'])):
                        hj = k2 * s['h'][si]
                        xij2 = ((d['x'][di] - s['x'][si]) ** 2 +
                                (d['y'][di] - s['y'][si]) ** 2 +
                                (d['z'][di] - s['z'][si]) ** 2)
                        if xij2 < hi * hi or xij2 < hj * hj:
                            env = dict(denv)
                            env.update(senv)
                            env.update(pair_symbols(kernel, d, s, di, si))
                            env.update(d_idx=di, s_idx=si)
                            for e in seqs:
                                call(e.loop, env)
            for di in range(npd):
                for e in deqs:
                    if hasattr(e, 'post_loop'):
                        call(e.post_loop, dict(denv, d_idx=di))


PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'cs', 'arho', 'au',
         'av', 'aw', 'ax', 'ay', 'az', 'dt_cfl', 'dt_force']


def make_array(rs, n, lo, hi, dx, dim, rho0, ghost=0, vel=1.0, hvar=0.0):
    pts = rs.uniform(lo, hi, size=(n, 3))
    if dim < 3:
        pts[:, 2] = 0.0
    if dim < 2:
        pts[:, 1] = 0.0
    a = dict((p, [0.0] * n) for p in PROPS)
    a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
    v = rs.normal(scale=vel, size=(n, 3))
    if dim < 3:
        v[:, 2] = 0.0
    if dim < 2:
        v[:, 1] = 0.0
    a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
    a['h'] = list(map(float, 1.3 * dx * (1.0 + hvar * rs.uniform(-1, 1, n))))
    a['m'] = [float(rho0 * dx ** dim)] * n
    a['rho'] = list(map(float, rho0 * (1.0 + 0.02 * rs.uniform(-1, 1, n))))
    a['_n_real'] = n - ghost
    return a


def gen_wcsph_case(kernels, basic, wc, kernel_name, dim, seed, hvar=0.0,
                   tensile=False, summation_density=False):
    """A small 3-array WCSPH evaluation through the reference's bodies."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho0, c0, gamma = 1000.0, 32.85, 7.0
    nf, nb, no = 90, 60, 12
    if dim == 3:
        lo_f, hi_f = [0, 0, 0], [0.5, 0.4, 0.4]
        lo_b, hi_b = [-0.15, -0.1, -0.15], [0.65, 0.5, 0.0]
        lo_o, hi_o = [0.5, 0.1, 0.0], [0.6, 0.3, 0.2]
    else:
        lo_f, hi_f = [0, 0, 0], [1.0, 0.8, 0.0]
        lo_b, hi_b = [-0.2, -0.2, 0], [1.2, 0.0, 0.0]
        lo_o, hi_o = [1.0, 0.0, 0.0], [1.2, 0.4, 0.0]
    arrays = dict(
        fluid=make_array(rs, nf, lo_f, hi_f, dx, dim, rho0, ghost=10, hvar=hvar),
        boundary=make_array(rs, nb, lo_b, hi_b, dx, dim, rho0, ghost=5,
                            vel=0.0, hvar=hvar),
        obstacle=make_array(rs, no, lo_o, hi_o, dx, dim, rho0, vel=0.0,
                            hvar=hvar))
    inputs = json.loads(json.dumps(arrays))
    fluids, solids = ['fluid'], ['boundary', 'obstacle']
    all_ = fluids + solids
    params = dict(rho0=rho0, c0=c0, gamma=gamma, alpha=0.25, beta=0.1,
                  gx=0.3, gy=-0.2 if dim > 1 else 0.0,
                  gz=-9.81 if dim == 3 else 0.0, tensile_correction=tensile,
                  hg_correction=True, summation_density=summation_density,
                  dim=dim, hdx=1.3, h0=1.3 * dx)
    groups = []
    if summation_density:
        groups.append((False, [basic.SummationDensity(dest=f, sources=all_)
                               for f in fluids]))
    g1 = [wc.TaitEOS(dest=f, sources=None, rho0=rho0, c0=c0, gamma=gamma)
          for f in fluids]
    g1 += [wc.TaitEOSHGCorrection(dest=s, sources=None, rho0=rho0, c0=c0,
                                  gamma=gamma) for s in solids]
    groups.append((False, g1))
    g2 = [basic.ContinuityEquation(dest=s, sources=fluids) for s in solids]
    for f in fluids:
        if not summation_density:
            g2.append(basic.ContinuityEquation(dest=f, sources=all_))
        g2.append(wc.MomentumEquation(
            dest=f, sources=all_, c0=c0, alpha=params['alpha'],
            beta=params['beta'], gx=params['gx'], gy=params['gy'],
            gz=params['gz'], tensile_correction=tensile))
        g2.append(basic.XSPHCorrection(dest=f, sources=[f]))
    groups.append((True, g2))
    evaluate_reference(kernel, arrays, groups)
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs,
                outputs=arrays)


def gen_laminar_case(kernels, basic, wc, kernel_name, dim, seed, nu=0.05, hvar=0.0):
    """WCSPHScheme(nu != 0) (scheme.py:486-496): the WCSPH Group with LaminarViscosity
    (wc/viscosity.py:5-27) inserted before XSPHCorrection, through the reference's bodies."""
    visc = _load('pysph.sph.wc.viscosity', os.path.join(REF, 'pysph/sph/wc/viscosity.py'))
    case = gen_wcsph_case(kernels, basic, wc, kernel_name, dim, seed, hvar=hvar)
    arrays = json.loads(json.dumps(case['inputs']))
    kernel = getattr(kernels, kernel_name)(dim=dim)
    p = case['params']
    fluids, solids = ['fluid'], ['boundary', 'obstacle']
    all_ = fluids + solids
    g1 = [wc.TaitEOS(dest=f, sources=None, rho0=p['rho0'], c0=p['c0'], gamma=p['gamma']) for f in fluids]
    g1 += [wc.TaitEOSHGCorrection(dest=s_, sources=None, rho0=p['rho0'], c0=p['c0'],
                                  gamma=p['gamma']) for s_ in solids]
    g2 = [basic.ContinuityEquation(dest=s_, sources=fluids) for s_ in solids]
    for f in fluids:
        g2.append(basic.ContinuityEquation(dest=f, sources=all_))
        g2.append(wc.MomentumEquation(dest=f, sources=all_, c0=p['c0'], alpha=p['alpha'],
                                      beta=p['beta'], gx=p['gx'], gy=p['gy'], gz=p['gz'],
                                      tensile_correction=False))
        g2.append(visc.LaminarViscosity(dest=f, sources=all_, nu=nu))
        g2.append(basic.XSPHCorrection(dest=f, sources=[f]))
    evaluate_reference(kernel, arrays, [(False, g1), (True, g2)])
    params = dict(p, nu=nu, eta=0.01)
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=case['inputs'], outputs=arrays)


def gen_monaghan_av_case(kernels, basic, wc, kernel_name, dim, seed, hvar=0.0):
    """The stand-alone artificial viscosity (basic_equations.py:195-257): a Group in which
    the fluid's velocity change comes from MonaghanArtificialViscosity alone (no
    MomentumEquation) next to ContinuityEquation and XSPHCorrection -- the branch of the
    device's pair body that the WCSPH scheme never takes."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho0, c0, gamma = 1000.0, 32.85, 7.0
    if dim == 3:
        lo_f, hi_f = [0, 0, 0], [0.5, 0.4, 0.4]
        lo_b, hi_b = [-0.15, -0.1, -0.15], [0.65, 0.5, 0.0]
    else:
        lo_f, hi_f = [0, 0, 0], [1.0, 0.8, 0.0]
        lo_b, hi_b = [-0.2, -0.2, 0], [1.2, 0.0, 0.0]
    arrays = dict(
        fluid=make_array(rs, 90, lo_f, hi_f, dx, dim, rho0, ghost=10, hvar=hvar),
        boundary=make_array(rs, 60, lo_b, hi_b, dx, dim, rho0, ghost=5, vel=0.3,
                            hvar=hvar))
    inputs = json.loads(json.dumps(arrays))
    all_ = ['fluid', 'boundary']
    params = dict(rho0=rho0, c0=c0, gamma=gamma, alpha=0.7, beta=1.3, dim=dim,
                  eps_xsph=0.4, names=all_)
    g1 = [wc.TaitEOS(dest=a, sources=None, rho0=rho0, c0=c0, gamma=gamma) for a in all_]
    g2 = [basic.ContinuityEquation(dest='fluid', sources=all_),
          basic.MonaghanArtificialViscosity(dest='fluid', sources=all_,
                                            alpha=params['alpha'], beta=params['beta']),
          basic.XSPHCorrection(dest='fluid', sources=['fluid'], eps=params['eps_xsph'])]
    evaluate_reference(kernel, arrays, [(False, g1), (True, g2)])
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs,
                outputs=arrays)


EDAC_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'au', 'av', 'aw',
              'uhat', 'vhat', 'what', 'auhat', 'avhat', 'awhat', 'ap', 'V', 'pavg',
              'nnbr']


def gen_edac_case(kernels, edac, kernel_name, dim, seed, nfluids=1, alpha=0.0,
                  nu=0.01, bql=True, hvar=0.0, gx=0.0, tdamp=0.0, t=0.0):
    """One evaluation of EDACScheme(fluids, solids=[], pb != 0).get_equations() --
    the reference's own scheme method AND equation bodies -- on random particles."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho0, c0 = 1.0, 10.0
    p0 = rho0 * c0 * c0
    names = ['fluid', 'fluid2'][:nfluids]
    arrays = {}
    for k, name in enumerate(names):
        n = 70 if k == 0 else 30
        hi = [0.5, 0.5, 0.4 if dim == 3 else 0.0]
        pts = rs.uniform(0.0, 1.0, size=(n, 3)) * np.array(hi)
        a = dict((q, [0.0] * n) for q in EDAC_PROPS)
        a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
        v = rs.normal(size=(n, 3))
        if dim < 3:
            v[:, 2] = 0.0
        a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
        vh = v + 0.05 * rs.normal(size=(n, 3)) * (np.arange(3) < dim)
        a['uhat'], a['vhat'], a['what'] = (list(map(float, vh[:, i])) for i in range(3))
        a['h'] = list(map(float, 1.0 * dx * (1.0 + hvar * rs.uniform(-1, 1, n))))
        a['m'] = [float(rho0 * dx ** dim * (1.0 + 0.5 * k))] * n
        a['p'] = list(map(float, rs.normal(scale=2.0, size=n)))
        a['_n_real'] = n - (6 if k == 0 else 0)     # some trailing ghosts
        arrays[name] = a
    inputs = json.loads(json.dumps(arrays))
    h0 = dx
    scheme = edac.EDACScheme(names, [], dim=dim, c0=c0, nu=nu, rho0=rho0, pb=p0,
                             gx=gx, tdamp=tdamp, h=h0, alpha=alpha, bql=bql)
    eqs = scheme.get_equations()
    groups = [(g.real, g.equations) for g in eqs]
    evaluate_reference(kernel, arrays, groups, t=t)
    params = dict(dim=dim, c0=c0, rho0=rho0, nu=nu, pb=p0, h=h0, alpha=alpha,
                  edac_alpha=0.5, bql=bql, gx=gx, gy=0.0, gz=0.0, tdamp=tdamp, t=t,
                  fluids=names,
                  groups=[[type(e).__name__ for e in g.equations] for g in eqs],
                  group_real=[bool(g.real) for g in eqs])
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs,
                outputs=arrays)


EDAC_WALL_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'au', 'av', 'aw',
                   'V', 'wij', 'ax', 'ay', 'az', 'uf', 'vf', 'wf', 'ug', 'vg', 'wg']


def gen_edac_wall_case(kernels, edac, kernel_name, dim, seed, alpha=0.0, nu=0.01, bql=True,
                       gx=0.0, gy=0.0, tdamp=0.0, t=0.0, moving=False):
    """One evaluation of EDACScheme(fluids=['fluid'], solids=['wall'], pb != 0).get_equations()
    (wc/edac.py:776-880 with solids: SourceNumberDensity, VolumeSummation, SolidWallPressureBC,
    SetWallVelocity in group 1, the average pressure in a group of its own, SolidWallNoSlipBC
    in group 2) through the reference's own scheme method and equation bodies: a slab of fluid
    over a two-layer wall, random perturbations, some trailing ghosts in both arrays."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho0, c0 = 1.0, 10.0
    p0 = rho0 * c0 * c0
    arrays = {}
    for name, n, ylo, yhi, props in (('fluid', 80, 0.0, 0.45, EDAC_PROPS),
                                     ('wall', 45, -0.25, 0.0, EDAC_WALL_PROPS)):
        ext = np.array([0.5, 1.0, 0.4 if dim == 3 else 0.0])
        pts = rs.uniform(0.0, 1.0, size=(n, 3)) * ext
        pts[:, 1] = ylo + (yhi - ylo) * rs.uniform(0.0, 1.0, n)
        a = dict((q, [0.0] * n) for q in props)
        a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
        a['h'] = [float(1.0 * dx)] * n
        a['m'] = [float(rho0 * dx ** dim)] * n
        if name == 'fluid':
            v = rs.normal(size=(n, 3)) * (np.arange(3) < dim)
            vh = v + 0.05 * rs.normal(size=(n, 3)) * (np.arange(3) < dim)
            a['uhat'], a['vhat'], a['what'] = (list(map(float, vh[:, i])) for i in range(3))
            a['p'] = list(map(float, rs.normal(scale=2.0, size=n)))
            a['_n_real'] = n - 6
        else:
            # prescribed wall velocity and acceleration (SetWallVelocity, SolidWallPressureBC)
            v = (0.3 * rs.normal(size=(n, 3)) if moving else np.zeros((n, 3))) * (np.arange(3) < dim)
            acc = (0.5 * rs.normal(size=(n, 3)) if moving else np.zeros((n, 3))) * (np.arange(3) < dim)
            a['au'], a['av'], a['aw'] = (list(map(float, acc[:, i])) for i in range(3))
            a['rho'] = list(map(float, rho0 * (1.0 + 0.05 * rs.uniform(-1, 1, n))))
            a['p'] = list(map(float, rs.normal(scale=2.0, size=n)))   # overwritten by the BC
            a['_n_real'] = n - 4
        a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
        arrays[name] = a
    inputs = json.loads(json.dumps(arrays))
    scheme = edac.EDACScheme(['fluid'], ['wall'], dim=dim, c0=c0, nu=nu, rho0=rho0, pb=p0,
                             gx=gx, gy=gy, tdamp=tdamp, h=dx, alpha=alpha, bql=bql)
    eqs = scheme.get_equations()
    groups = [(g.real, g.equations) for g in eqs]
    evaluate_reference(kernel, arrays, groups, t=t)
    params = dict(dim=dim, c0=c0, rho0=rho0, nu=nu, pb=p0, h=dx, alpha=alpha, edac_alpha=0.5,
                  bql=bql, gx=gx, gy=gy, gz=0.0, tdamp=tdamp, t=t, fluids=['fluid'],
                  solids=['wall'],
                  groups=[[type(e).__name__ for e in g.equations] for g in eqs],
                  group_real=[bool(g.real) for g in eqs],
                  sources=[[list(e.sources or []) for e in g.equations] for g in eqs])
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs, outputs=arrays)


EDAC_EXT_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'au', 'av', 'aw', 'ax', 'ay',
                  'az', 'ap', 'V']
EDAC_EXT_WALL_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'au', 'av', 'aw', 'ap',
                       'p0', 'V', 'wij', 'ax', 'ay', 'az', 'uf', 'vf', 'wf', 'ug', 'vg', 'wg']


def gen_edac_ext_case(kernels, edac, kernel_name, dim, seed, alpha=0.0, nu=0.01, eps=0.5,
                      clamp_p=False, gy=0.0, tdamp=0.0, t=0.0, moving=False, walls=True,
                      nfluids=1):
    """One evaluation of EDACScheme(fluids, solids, pb=0).get_equations(): the EXTERNAL-flow
    branch (wc/edac.py:882-971) -- SummationDensity, the wall equations [+ ClampWallPressure],
    then the number-density MomentumEquation (:301-352), [artificial / physical viscosity,
    no-slip], EDACEquation and XSPHCorrection -- through the reference's own scheme method and
    equation bodies."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho0, c0 = 1.0, 10.0
    arrays = {}
    fluids = ['fluid', 'fluid2'][:nfluids]
    spec = [(f, 80 if k == 0 else 30, 0.0, 0.45, EDAC_EXT_PROPS) for k, f in enumerate(fluids)]
    if walls:
        spec.append(('wall', 45, -0.25, 0.0, EDAC_EXT_WALL_PROPS))
    for name, n, ylo, yhi, props in spec:
        ext = np.array([0.5, 1.0, 0.4 if dim == 3 else 0.0])
        pts = rs.uniform(0.0, 1.0, size=(n, 3)) * ext
        pts[:, 1] = ylo + (yhi - ylo) * rs.uniform(0.0, 1.0, n)
        a = dict((q, [0.0] * n) for q in props)
        a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
        a['h'] = [float(1.0 * dx)] * n
        a['m'] = [float(rho0 * dx ** dim * (1.0 + 0.5 * (name == 'fluid2')))] * n
        a['p'] = list(map(float, rs.normal(scale=2.0, size=n)))
        if name != 'wall':
            v = rs.normal(size=(n, 3)) * (np.arange(3) < dim)
            a['_n_real'] = n - (6 if name == 'fluid' else 0)
        else:
            v = (0.3 * rs.normal(size=(n, 3)) if moving else np.zeros((n, 3))) * (np.arange(3) < dim)
            acc = (0.5 * rs.normal(size=(n, 3)) if moving else np.zeros((n, 3))) * (np.arange(3) < dim)
            a['au'], a['av'], a['aw'] = (list(map(float, acc[:, i])) for i in range(3))
            a['rho'] = list(map(float, rho0 * (1.0 + 0.05 * rs.uniform(-1, 1, n))))
            a['_n_real'] = n - 4
        a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
        arrays[name] = a
    inputs = json.loads(json.dumps(arrays))
    scheme = edac.EDACScheme(fluids, ['wall'] if walls else [], dim=dim, c0=c0, nu=nu, rho0=rho0,
                             pb=0.0, gy=gy, tdamp=tdamp, h=dx, alpha=alpha, eps=eps,
                             clamp_p=clamp_p)
    eqs = scheme.get_equations()
    groups = [(g.real, g.equations) for g in eqs]
    evaluate_reference(kernel, arrays, groups, t=t)
    params = dict(dim=dim, c0=c0, rho0=rho0, nu=nu, pb=0.0, h=dx, alpha=alpha, edac_alpha=0.5,
                  eps=eps, clamp_p=clamp_p, gx=0.0, gy=gy, gz=0.0, tdamp=tdamp, t=t,
                  fluids=fluids, solids=['wall'] if walls else [],
                  groups=[[type(e).__name__ for e in g.equations] for g in eqs],
                  group_real=[bool(g.real) for g in eqs],
                  sources=[[list(e.sources or []) for e in g.equations] for g in eqs])
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs, outputs=arrays)


def gen_edac_ext_stepper(edac):
    """EDACStep (wc/edac.py:82-133) on random data: initialize, stage1, stage2."""
    rs = np.random.RandomState(13)
    n = 12
    names = ['x', 'y', 'z', 'u', 'v', 'w', 'p', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'ap',
             'x0', 'y0', 'z0', 'u0', 'v0', 'w0', 'p0']
    st = edac.EDACStep()
    out = {}
    dt = 0.013
    for which in ('initialize', 'stage1', 'stage2'):
        a = dict((k, list(map(float, rs.normal(size=n)))) for k in names)
        before = json.loads(json.dumps(a))
        env = dict(('d_' + k, v) for k, v in a.items())
        env['dt'] = dt
        for i in range(n):
            call(getattr(st, which), dict(env, d_idx=i))
        out[which] = dict(inputs=before, outputs=a)
    out['dt'] = dt
    return out


def gen_edac_ext_cases(kernels, edac):
    return [
        gen_edac_ext_case(kernels, edac, 'QuinticSpline', 2, 501, gy=-1.0),
        gen_edac_ext_case(kernels, edac, 'CubicSpline', 3, 502, alpha=0.2, moving=True, clamp_p=True),
        gen_edac_ext_case(kernels, edac, 'WendlandQuintic', 3, 503, nu=0.0, alpha=0.1, eps=0.0,
                          gy=-1.0, tdamp=1.0, t=0.3),
        gen_edac_ext_case(kernels, edac, 'QuinticSpline', 2, 504, walls=False, nfluids=2, eps=0.3),
    ]


def gen_edac_stepper(edac):
    rs = np.random.RandomState(12)
    n = 7
    st = edac.EDACTVFStep()
    names = ['x', 'y', 'z', 'u', 'v', 'w', 'p', 'x0', 'y0', 'z0', 'u0', 'v0', 'w0',
             'p0', 'au', 'av', 'aw', 'auhat', 'avhat', 'awhat', 'uhat', 'vhat',
             'what', 'ap']
    a = dict((k, list(map(float, rs.normal(size=n)))) for k in names)
    inputs = json.loads(json.dumps(a))
    dt = 0.0123
    res = {}
    for which, meth in (('initialize', st.initialize), ('stage1', st.stage1),
                        ('stage2', st.stage2)):
        b = json.loads(json.dumps(inputs))
        env = dict(('d_' + k, v) for k, v in b.items())
        for i in range(n):
            call(meth, dict(env, d_idx=i, dt=dt))
        res[which] = b
    return dict(dt=dt, inputs=inputs, outputs=res)


SOLID_PROPS = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'p', 'cs', 'e', 'arho',
               'au', 'av', 'aw', 'ax', 'ay', 'az', 'ae'] + \
    ['v%d%d' % (i, j) for i in range(3) for j in range(3)] + \
    [pre + k for pre in ('s', 'as', 'r') for k in ('00', '01', '02', '11', '12', '22')]


class _Ref(float):
    """V[0] as the reference body writes it: a float that remembers where it lives, so
    that cython.address(V[0]) can hand the whole vector to the compiled routine."""
    def __new__(cls, value, parent):
        o = float.__new__(cls, value)
        o.parent = parent
        return o


class _Mat(list):
    def __getitem__(self, i):
        v = list.__getitem__(self, i)
        return _Ref(v, self) if isinstance(v, (int, float)) else v


def load_reference_solid():
    """solid_mech/basic.py, unmodified.  MonaghanArtificialStress.loop is written for
    Cython (declare('matrix'), cython.address, cimported eigen_decomposition): the four
    names are provided here, the last two backed by the reference's OWN compiled
    pysph/base/linalg3.pyx (oracle/_ref/linalg3*.so)."""
    import linalg3
    mod = _load('pysph.sph.solid_mech.basic',
                os.path.join(REF, 'pysph/sph/solid_mech/basic.py'))

    def declare(spec):
        if '(3,3)' in spec.replace(' ', ''):
            return [_Mat([0.0, 0.0, 0.0]) for _ in range(3)]
        return _Mat([0.0, 0.0, 0.0])

    def eigen_decomposition(S, R, v_addr):
        d, v = linalg3.py_eigen_decompose_eispack(np.array(S, dtype=float))
        for i in range(3):
            list.__setitem__(v_addr, i, float(d[i]))
            for j in range(3):
                list.__setitem__(R[i], j, float(v[i, j]))

    def transform_diag_inv(rd_addr, R, Rab):
        res = linalg3.py_transform_diag_inv(np.array(list(rd_addr), dtype=float),
                                            np.array(R, dtype=float))
        for i in range(3):
            for j in range(3):
                list.__setitem__(Rab[i], j, float(res[i, j]))

    cython = types.SimpleNamespace(address=lambda ref: ref.parent)
    mod.declare, mod.cython = declare, cython
    mod.eigen_decomposition, mod.transform_diag_inv = eigen_decomposition, transform_diag_inv
    mod.pow = pow
    return mod


def gen_solid_case(kernels, solid, kernel_name, dim, seed, wdeltap=True, two=False,
                   grad3d=False, wall=False):
    """One evaluation of ElasticSolidsScheme(...).get_equations() (the reference's scheme
    method and equation bodies) on random particles with random stresses.  wall=True adds a
    rigid `solids` array: a source of every pair equation, a destination of none -- it
    contributes with whatever p / s / r it carries (solid_mech/basic.py:613-648)."""
    rs = np.random.RandomState(seed)
    kernel = getattr(kernels, kernel_name)(dim=dim)
    dx = 0.1
    rho_ref, E, nu = 1.2, 1e3, 0.3975
    G = E / (2.0 * (1.0 + nu))
    c0 = math.sqrt(E / (3 * (1.0 - 2 * nu)) / rho_ref)      # get_speed_of_sound, :19-29
    names = ['ring', 'ring2'][:2 if two else 1]
    arrays, consts = {}, {}
    for k, name in enumerate(names):
        n = 60 if k == 0 else 25
        hi = [0.5, 0.5, 0.4 if dim == 3 else 0.0]
        pts = rs.uniform(0.0, 1.0, size=(n, 3)) * np.array(hi) + 0.3 * k
        a = dict((q, [0.0] * n) for q in SOLID_PROPS)
        a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
        v = rs.normal(size=(n, 3)) * (np.arange(3) < dim)
        a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
        a['h'] = [1.3 * dx] * n
        a['m'] = [float(rho_ref * dx ** dim)] * n
        a['rho'] = list(map(float, rho_ref * (1 + 0.05 * rs.uniform(-1, 1, n))))
        a['cs'] = [c0] * n
        for key in ('s00', 's01', 's11') + (('s02', 's12', 's22') if dim == 3 else ()):
            a[key] = list(map(float, 20.0 * rs.normal(size=n)))
        a['_n_real'] = n - (5 if k == 0 else 0)
        arrays[name] = a
        wd = kernel.kernel([0, 0, 0], dx, 1.3 * dx) if wdeltap else -1.0
        consts[name] = dict(wdeltap=[wd], n=[4.0], G=[G], rho_ref=[rho_ref],
                            c0_ref=[c0])
    solids = []
    if wall:
        n = 30
        hi = [0.5, 0.5, 0.4 if dim == 3 else 0.0]
        pts = rs.uniform(0.0, 1.0, size=(n, 3)) * np.array(hi) + np.array([0.25, 0.0, 0.0])
        a = dict((q, [0.0] * n) for q in SOLID_PROPS)
        a['x'], a['y'], a['z'] = (list(map(float, pts[:, i])) for i in range(3))
        v = 0.3 * rs.normal(size=(n, 3)) * (np.arange(3) < dim)
        a['u'], a['v'], a['w'] = (list(map(float, v[:, i])) for i in range(3))
        a['h'] = [1.2 * dx] * n
        a['m'] = [float(1.5 * rho_ref * dx ** dim)] * n
        a['rho'] = list(map(float, 1.5 * rho_ref * (1 + 0.02 * rs.uniform(-1, 1, n))))
        a['cs'] = [1.3 * c0] * n
        a['p'] = list(map(float, 15.0 * rs.normal(size=n)))
        for key in ('s00', 's01', 's11') + (('s02', 's12', 's22') if dim == 3 else ()):
            a[key] = list(map(float, 10.0 * rs.normal(size=n)))
        for key in ('r00', 'r01', 'r11') + (('r02', 'r12', 'r22') if dim == 3 else ()):
            a[key] = list(map(float, 0.5 * rs.normal(size=n)))
        a['_n_real'] = n
        arrays['wall'] = a
        consts['wall'] = dict(wdeltap=[-1.0], n=[4.0], G=[0.0], rho_ref=[1.5 * rho_ref],
                              c0_ref=[1.3 * c0])
        solids = ['wall']
    inputs = json.loads(json.dumps(arrays))
    scheme = solid.ElasticSolidsScheme(names, solids, dim=dim, artificial_stress_eps=0.3,
                                       xsph_eps=0.5, alpha=1.0, beta=1.5)
    elastic = list(names)
    names = names + solids
    eqs = scheme.get_equations()
    if grad3d:
        # the scheme always emits VelocityGradient2D (solid_mech/basic.py:620-623); a 3-D
        # run needs the reference's VelocityGradient3D (basic_equations.py:101-148) in its place
        basic = sys.modules['pysph.sph.basic_equations']
        g1 = eqs[0].equations
        for k, e in enumerate(g1):
            if type(e).__name__ == 'VelocityGradient2D':
                g1[k] = basic.VelocityGradient3D(dest=e.dest, sources=e.sources)
    # array constants are seen by the bodies as d_<name> / s_<name>
    for name in names:
        for ck, cv in consts[name].items():
            arrays[name][ck] = cv
    groups = [(g.real, g.equations) for g in eqs]
    evaluate_reference(kernel, arrays, groups)
    for name in names:
        for ck in consts[name]:
            del arrays[name][ck]
    params = dict(dim=dim, eps=0.3, eps_xsph=0.5, alpha=1.0, beta=1.5, names=names,
                  elastic=elastic, solids=solids,
                  constants=consts, grad3d=bool(grad3d),
                  groups=[[type(e).__name__ for e in g.equations] for g in eqs],
                  group_real=[bool(g.real) for g in eqs])
    return dict(kernel=kernel_name, dim=dim, params=params, inputs=inputs,
                outputs=arrays)


def gen_solid_stepper(steps):
    rs = np.random.RandomState(13)
    n = 6
    st = steps.SolidMechStep()
    sym = ['00', '01', '02', '11', '12', '22']
    names = ['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'e', 'x0', 'y0', 'z0', 'u0', 'v0', 'w0',
             'rho0', 'e0', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'arho', 'ae'] + \
        ['s' + k for k in sym] + ['s' + k + '0' for k in sym] + ['as' + k for k in sym]
    a = dict((k, list(map(float, rs.normal(size=n)))) for k in names)
    inputs = json.loads(json.dumps(a))
    dt = 0.0123
    res = {}
    for which, meth in (('initialize', st.initialize), ('stage1', st.stage1),
                        ('stage2', st.stage2)):
        b = json.loads(json.dumps(inputs))
        env = dict(('d_' + k, v) for k, v in b.items())
        for i in range(n):
            call(meth, dict(env, d_idx=i, dt=dt))
        res[which] = b
    return dict(dt=dt, inputs=inputs, outputs=res)


def gen_output_fixture():
    """pysph/solver/output.py (unmodified, loaded by path) writes the fixture and reads
    back a file written by pysph_b200.output.  Its three imports are satisfied by the
    stand-in ParticleArray of pysph_b200 and by pysph_b200.output.get_particles_info
    (a restatement of pysph/base/utils.py:466-497, the one part that needs cyarray)."""
    import tempfile
    sys.path.insert(0, ROOT)
    from pysph_b200 import output as ours
    from pysph_b200 import particle_array as ppa
    m_pa = types.ModuleType('pysph.base.particle_array')
    m_pa.ParticleArray = ppa.ParticleArray
    sys.modules['pysph.base.particle_array'] = m_pa
    utils = sys.modules['pysph.base.utils']
    utils.get_particles_info = ours.get_particles_info
    utils.get_particle_array = ppa.get_particle_array
    sys.modules['pysph'].has_h5py = lambda: False
    ref = _load('pysph.solver.output', os.path.join(REF, 'pysph/solver/output.py'))

    rs = np.random.RandomState(5)

    def arrays():
        f = ppa.get_particle_array_wcsph(
            name='fluid', x=rs.uniform(size=9), y=rs.uniform(size=9),
            u=rs.normal(size=9), rho=1000.0 + rs.normal(size=9), h=0.013, m=0.7,
            p=rs.normal(size=9))
        f.set_num_real_particles(7)                   # 2 trailing ghosts
        f.gid[:] = np.arange(9) + 100
        f.tag[7:] = 1
        b = ppa.get_particle_array_wcsph(name='boundary', x=rs.uniform(size=4),
                                         h=0.013, m=0.7, rho=1000.0)
        b.add_constant('total_mass', [2.8])
        return [f, b]
    pas = arrays()
    solver_data = {'dt': 1.25e-4, 't': 0.0375, 'count': 300}
    # (a) the reference's dump() writes the committed fixture
    ref.dump(os.path.join(GOLD, 'ref_dump'), pas, solver_data, detailed_output=False,
             only_real=True)
    assert os.path.exists(os.path.join(GOLD, 'ref_dump.npz'))
    # (b) the reference's load() reads OUR dump
    tmp = tempfile.mkdtemp()
    mine = ours.dump(os.path.join(tmp, 'ours_00300'), pas, solver_data)
    back = ref.load(mine)
    assert dict(back['solver_data']) == solver_data
    for pa in pas:
        q = back['arrays'][pa.name]
        assert q.get_number_of_particles() == pa.get_number_of_particles(real=True)
        for k in pa.output_property_arrays:
            assert np.array_equal(q.properties[k],
                                  pa.properties[k][:pa.num_real_particles]), k
    expect = dict((pa.name, dict((k, list(map(float, pa.properties[k][:pa.num_real_particles])))
                                 for k in pa.output_property_arrays)) for pa in pas)
    return dict(solver_data=solver_data, arrays=expect,
                output_property_arrays=dict((pa.name, pa.output_property_arrays)
                                            for pa in pas),
                all_properties=dict((pa.name, sorted(pa.properties)) for pa in pas),
                constants={'boundary': {'total_mass': [2.8]}})


def gen_steppers(steps):
    rs = np.random.RandomState(11)
    n = 7
    st = steps.WCSPHStep()
    names = ['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'x0', 'y0', 'z0', 'u0', 'v0',
             'w0', 'rho0', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'arho']
    a = dict((k, list(map(float, rs.normal(size=n)))) for k in names)
    inputs = json.loads(json.dumps(a))
    dt = 0.0123
    res = {}
    for which, meth in (('initialize', st.initialize), ('stage1', st.stage1),
                        ('stage2', st.stage2)):
        b = json.loads(json.dumps(inputs))
        env = dict(('d_' + k, v) for k, v in b.items())
        for i in range(n):
            call(meth, dict(env, d_idx=i, dt=dt))
        res[which] = b
    return dict(dt=dt, inputs=inputs, outputs=res)


def gen_eos(wc):
    rs = np.random.RandomState(3)
    rho = list(map(float, 1000.0 * (1 + 0.03 * rs.uniform(-1, 1, 16))))
    out = {}
    for name, cls, kw in (('TaitEOS', wc.TaitEOS, dict(p0=12.5)),
                          ('TaitEOSHGCorrection', wc.TaitEOSHGCorrection, {})):
        e = cls(dest='f', sources=None, rho0=1000.0, c0=32.85, gamma=7.0, **kw)
        r = list(rho)
        p = [0.0] * len(r)
        cs = [0.0] * len(r)
        for i in range(len(r)):
            e.loop(i, r, p, cs)
        out[name] = dict(rho_in=rho, rho_out=r, p=p, cs=cs, rho0=1000.0,
                         c0=32.85, gamma=7.0, p0=kw.get('p0', 0.0))
    f = wc.UpdateSmoothingLengthFerrari(dest='f', sources=None, dim=2, hdx=1.3)
    h = [0.0] * len(rho)
    m = [0.9] * len(rho)
    for i in range(len(rho)):
        f.loop(i, rho, h, m)
    out['UpdateSmoothingLengthFerrari'] = dict(rho=rho, m=m, h=h, dim=2, hdx=1.3)
    return out


def gen_density_1d(kernels, basic):
    """test_acceleration_eval.py:294-303: 10 points on [0,1], m=1, h=1.05dx."""
    n = 10
    dx = 1.0 / (n - 1)
    x = list(map(float, np.linspace(0, 1, n)))
    a = dict((p, [0.0] * n) for p in PROPS)
    a['x'] = x
    a['m'] = [1.0] * n
    a['h'] = [1.05 * dx] * n
    a['_n_real'] = n
    kernel = kernels.CubicSpline(dim=1)
    eq = basic.SummationDensity(dest='fluid', sources=['fluid'])
    arrays = dict(fluid=a)
    evaluate_reference(kernel, arrays, [(True, [eq])])
    k2 = 2.0
    counts = []
    for i in range(n):
        c = 0
        for j in range(n):
            r2 = (x[i] - x[j]) ** 2
            if r2 < (k2 * a['h'][i]) ** 2 or r2 < (k2 * a['h'][j]) ** 2:
                c += 1
        counts.append(c)
    return dict(x=x, h=a['h'], m=a['m'], rho=a['rho'], nbr_counts=counts)


def gen_laminar_cases(kernels, basic, wc):
    return [gen_laminar_case(kernels, basic, wc, 'CubicSpline', 3, 121),
            gen_laminar_case(kernels, basic, wc, 'WendlandQuintic', 2, 122, nu=0.2, hvar=0.1)]


def gen_edac_wall_cases(kernels, edac):
    return [
        gen_edac_wall_case(kernels, edac, 'QuinticSpline', 2, 401, gy=-1.0),
        gen_edac_wall_case(kernels, edac, 'QuinticSpline', 3, 402, alpha=0.2, moving=True, gx=0.4),
        gen_edac_wall_case(kernels, edac, 'CubicSpline', 3, 403, nu=0.0, alpha=0.1, bql=False,
                           gy=-1.0, tdamp=1.0, t=0.3),
        gen_edac_wall_case(kernels, edac, 'WendlandQuintic', 2, 404, moving=True),
    ]


def main():
    if sys.argv[1:] == ['laminar']:
        kernels, basic, wc, steps, c_kernels = load_reference()
        with open(os.path.join(GOLD, 'laminar_cases.json'), 'w') as f:
            json.dump(gen_laminar_cases(kernels, basic, wc), f)
        print('wrote laminar_cases.json', os.path.getsize(os.path.join(GOLD, 'laminar_cases.json')), 'bytes')
        return
    if sys.argv[1:] == ['edac_ext']:            # only the files added last
        kernels, basic, wc, steps, c_kernels = load_reference()
        tvf, edac = load_reference_edac()
        edac.sin, edac.M_PI = math.sin, math.pi
        for name, obj in (('edac_ext_cases.json', gen_edac_ext_cases(kernels, edac)),
                          ('edac_ext_stepper.json', gen_edac_ext_stepper(edac))):
            with open(os.path.join(GOLD, name), 'w') as f:
                json.dump(obj, f)
            print('wrote', name, os.path.getsize(os.path.join(GOLD, name)), 'bytes')
        return
    if sys.argv[1:] == ['edac_walls']:          # only the file added last
        kernels, basic, wc, steps, c_kernels = load_reference()
        tvf, edac = load_reference_edac()
        with open(os.path.join(GOLD, 'edac_wall_cases.json'), 'w') as f:
            json.dump(gen_edac_wall_cases(kernels, edac), f)
        print('wrote edac_wall_cases.json', os.path.getsize(os.path.join(GOLD, 'edac_wall_cases.json')), 'bytes')
        return
    kernels, basic, wc, steps, c_kernels = load_reference()
    os.makedirs(GOLD, exist_ok=True)

    def dump(name, obj):
        with open(os.path.join(GOLD, name), 'w') as f:
            json.dump(obj, f)
        print('wrote', name, os.path.getsize(os.path.join(GOLD, name)), 'bytes')

    dump('kernels.json', gen_kernels(kernels, c_kernels))
    dump('eos.json', gen_eos(wc))
    dump('steppers.json', gen_steppers(steps))
    dump('density_1d.json', gen_density_1d(kernels, basic))
    cases = [
        gen_wcsph_case(kernels, basic, wc, 'CubicSpline', 3, 101),
        gen_wcsph_case(kernels, basic, wc, 'WendlandQuintic', 3, 102, hvar=0.15),
        gen_wcsph_case(kernels, basic, wc, 'WendlandQuintic', 2, 103),
        gen_wcsph_case(kernels, basic, wc, 'QuinticSpline', 3, 104, tensile=True),
        gen_wcsph_case(kernels, basic, wc, 'CubicSpline', 2, 105, hvar=0.1,
                       summation_density=True),
        gen_wcsph_case(kernels, basic, wc, 'Gaussian', 3, 106),
    ]
    dump('wcsph_cases.json', cases)
    dump('laminar_cases.json', gen_laminar_cases(kernels, basic, wc))
    dump('monaghan_av_cases.json', [
        gen_monaghan_av_case(kernels, basic, wc, 'CubicSpline', 3, 111),
        gen_monaghan_av_case(kernels, basic, wc, 'WendlandQuintic', 2, 112, hvar=0.1),
    ])
    tvf, edac = load_reference_edac()
    ecases = [
        gen_edac_case(kernels, edac, 'QuinticSpline', 2, 201),
        gen_edac_case(kernels, edac, 'QuinticSpline', 3, 202, hvar=0.1),
        gen_edac_case(kernels, edac, 'CubicSpline', 3, 203, nfluids=2, alpha=0.2,
                      gx=0.7, tdamp=1.0, t=0.3),
        gen_edac_case(kernels, edac, 'WendlandQuintic', 2, 204, nu=0.0, bql=False),
    ]
    dump('edac_cases.json', ecases)
    dump('edac_stepper.json', gen_edac_stepper(edac))
    dump('edac_wall_cases.json', gen_edac_wall_cases(kernels, edac))
    dump('edac_ext_cases.json', gen_edac_ext_cases(kernels, edac))
    dump('edac_ext_stepper.json', gen_edac_ext_stepper(edac))
    solid = load_reference_solid()
    scases = [
        gen_solid_case(kernels, solid, 'CubicSpline', 2, 301),
        gen_solid_case(kernels, solid, 'CubicSpline', 3, 302, two=True),
        gen_solid_case(kernels, solid, 'WendlandQuintic', 2, 303, wdeltap=False),
        gen_solid_case(kernels, solid, 'CubicSpline', 3, 304, grad3d=True),
        gen_solid_case(kernels, solid, 'CubicSpline', 2, 305, wall=True),
        gen_solid_case(kernels, solid, 'CubicSpline', 3, 306, two=True, grad3d=True, wall=True),
    ]
    dump('solid_cases.json', scases)
    dump('solid_stepper.json', gen_solid_stepper(steps))
    dump('ref_dump_expect.json', gen_output_fixture())


if __name__ == '__main__':
    main()

"""Translate equation Groups into the evaluator's op list.

This is the B200 counterpart of the reference's MegaGroup regrouping +
code generation (pysph/sph/acceleration_eval.py:94-162,
pysph/sph/acceleration_eval_cython.mako:10-154): instead of emitting Cython it
recognises each equation by *class name* and maps every Group onto

    ('eos', arr, hg, rho0, c0, gamma, p0, real_only)      no-source loop
    ('ferrari', arr, hdx, dim, real_only)                 no-source loop
    ('pair', PairProgram)                                  fused pair kernel
    ('update_nnps',)                                       Group(update_nnps=True)

Anything the CUDA library has no kernel for raises NotImplementedError here,
i.e. at setup time, never silently at run time.
"""
from . import _lib

PAIR_EQUATIONS = {
    'SummationDensity': _lib.EQ_SUMMATION_DENSITY,
    'ContinuityEquation': _lib.EQ_CONTINUITY,
    'MomentumEquation': _lib.EQ_MOMENTUM,
    'XSPHCorrection': _lib.EQ_XSPH,
    'MonaghanArtificialViscosity': _lib.EQ_MONAGHAN_AV,
}
NO_SOURCE_EQUATIONS = ('TaitEOS', 'TaitEOSHGCorrection',
                       'UpdateSmoothingLengthFerrari')

# properties each pair equation needs on dest / source arrays (checked like
# check_equation_array_properties, pysph/sph/acceleration_eval.py:32-73)
_REQUIRED = ('x', 'y', 'z', 'h', 'm', 'rho')


def _eq_name(eq):
    return eq.__class__.__name__


def _set_once(params, key, value, eq):
    if key in params and params[key] != value:
        raise NotImplementedError(
            'B200 backend: %s instances in one Group use different %s (%r vs '
            '%r); the fused kernel takes one value per Group'
            % (_eq_name(eq), key, params[key], value))
    params[key] = value


def _flatten(groups):
    """Group(subgroups) -> sequence of plain groups (mako:320-343)."""
    out = []
    for g in groups:
        if getattr(g, 'has_subgroups', False):
            if getattr(g, 'iterate', False):
                raise NotImplementedError(
                    'B200 backend: iterated groups are not supported')
            out.extend(_flatten(g.equations))
            if getattr(g, 'update_nnps', False):
                out.append('update_nnps')
        else:
            out.append(g)
    return out


def group_equations(equations):
    """A bare list of equations is one Group (equation.py:346-373)."""
    from .equations import Group
    if len(equations) == 0:
        return []
    if all(hasattr(e, 'equations') for e in equations):
        return list(equations)
    if any(hasattr(e, 'equations') for e in equations):
        raise ValueError('mix of Groups and Equations')
    return [Group(equations=list(equations))]


def build_program(groups, array_names, dim):
    """groups: list of Group objects (ours or PySPH's)."""
    index = dict((n, i) for i, n in enumerate(array_names))
    ops = []
    for g in _flatten(group_equations(groups)):
        if g == 'update_nnps':
            ops.append(('update_nnps',))
            continue
        for attr, default in (('iterate', False), ('condition', None),
                              ('pre', None), ('post', None), ('start_idx', 0),
                              ('stop_idx', None)):
            if getattr(g, attr, default) not in (default,):
                raise NotImplementedError(
                    'B200 backend: Group(%s=%r) is not supported'
                    % (attr, getattr(g, attr)))
        real_only = 1 if getattr(g, 'real', True) else 0
        pair_eqs = []
        nosrc_ops = []
        for eq in g.equations:
            name = _eq_name(eq)
            if eq.dest not in index:
                raise ValueError('equation %s: unknown destination array %r'
                                 % (name, eq.dest))
            d = index[eq.dest]
            if name in PAIR_EQUATIONS:
                if not eq.sources:
                    raise ValueError('%s needs sources' % name)
                pair_eqs.append(eq)
            elif name in ('TaitEOS', 'TaitEOSHGCorrection'):
                hg = 1 if name == 'TaitEOSHGCorrection' else 0
                nosrc_ops.append(('eos', d, hg, float(eq.rho0), float(eq.c0),
                                  float(eq.gamma),
                                  float(getattr(eq, 'p0', 0.0)), real_only))
            elif name == 'UpdateSmoothingLengthFerrari':
                nosrc_ops.append(('ferrari', d, float(eq.hdx),
                                  int(round(1.0 / eq.dim1)), real_only))
            else:
                raise NotImplementedError(
                    'B200 backend: no CUDA kernel for equation %r (supported: '
                    '%s)' % (name, ', '.join(sorted(list(PAIR_EQUATIONS) +
                                                    list(NO_SOURCE_EQUATIONS)))))
        if pair_eqs and nosrc_ops:
            raise NotImplementedError(
                'B200 backend: a Group mixing no-source equations and pair '
                'equations is not supported; put them in separate Groups '
                '(as WCSPHScheme does, scheme.py:414-483)')
        ops.extend(nosrc_ops)
        if pair_eqs:
            prog = _lib.PairProgram()
            params = {}
            all_bits = 0
            for eq in pair_eqs:
                name = _eq_name(eq)
                bit = PAIR_EQUATIONS[name]
                all_bits |= bit
                d = index[eq.dest]
                for s in eq.sources:
                    if s not in index:
                        raise ValueError('equation %s: unknown source array %r'
                                         % (name, s))
                    prog.eqmask[d][index[s]] |= bit
                if name == 'MomentumEquation':
                    for k in ('c0', 'alpha', 'beta', 'gx', 'gy', 'gz'):
                        _set_once(params, k, float(getattr(eq, k)), eq)
                    _set_once(params, 'tensile_correction',
                              int(bool(eq.tensile_correction)), eq)
                elif name == 'MonaghanArtificialViscosity':
                    for k in ('alpha', 'beta'):
                        _set_once(params, k, float(getattr(eq, k)), eq)
                elif name == 'XSPHCorrection':
                    _set_once(params, 'eps_xsph', float(eq.eps), eq)
            if (all_bits & _lib.EQ_SUMMATION_DENSITY) and (
                    all_bits & (_lib.EQ_MOMENTUM | _lib.EQ_XSPH |
                                _lib.EQ_MONAGHAN_AV)):
                raise NotImplementedError(
                    'B200 backend: SummationDensity (writes rho) cannot share a '
                    'Group with equations that read rho; the reference '
                    'evaluates destinations one after another there '
                    '(mako:20-135) -- use separate Groups')
            if (all_bits & _lib.EQ_MOMENTUM) and (all_bits & _lib.EQ_MONAGHAN_AV):
                raise NotImplementedError(
                    'B200 backend: MomentumEquation already contains the '
                    'artificial viscosity; combining it with '
                    'MonaghanArtificialViscosity in one Group is not supported')
            prog.real_only = real_only
            for k, v in params.items():
                setattr(prog, k, v)
            ops.append(('pair', prog))
        if getattr(g, 'update_nnps', False):
            ops.append(('update_nnps',))
    return ops

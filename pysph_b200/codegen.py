"""Generic-equation fallback: user ``Equation`` bodies compiled at run time with NVRTC.

The reference turns every ``Equation``'s Python methods into C through its code generator
(pysph/sph/equation.py:389-420 "Equation", :885-892; acceleration_eval_cython.mako:10-155
``do_group``).  The B200 backend has hand-written kernels for the equations of the WCSPH /
EDAC / elastic-dynamics paths; a Group made of equations it does NOT know (SURVEY.md 8f-4:
"compile arbitrary user Equation.loop bodies via NVRTC") takes this path instead:

* ``initialize / loop / post_loop`` are translated from their Python source (``ast``) into
  CUDA C -- the subset the reference's own translator accepts for these methods: arithmetic,
  comparisons, ``if / elif / else``, ``for i in range(..)``, calls of math functions,
  ``self.<number>`` attributes (baked in as literals), ``declare('matrix(n)')`` locals,
  ``d_<prop>[d_idx]`` / ``s_<prop>[s_idx]`` element access and the pre-computed pair symbols
  (equation.py:192-273: XIJ, RIJ, R2IJ, HIJ, VIJ, WIJ, DWIJ, RHOIJ, RHOIJ1, EPS, WI, WJ, DWI, DWJ);
* one kernel per (Group, destination array), thread per destination, walking the SAME persistent
  neighbour lists as ``k_pair_list`` with the same fp32 accept test on cell-relative positions
  (so the neighbour sets are the fast path's); arithmetic of the bodies in fp64;
* three launches per destination -- initialize, loop, post_loop -- and destinations in Group
  order, which is the reference's loop nest (a loop body may read what another particle's
  initialize, or an earlier destination's post_loop, wrote);
* properties the device pool does not have are created as fp64 "user" properties
  (``b200sph_user_property``) and travel with push / pull like any other.

* ``py_initialize(dst, t, dt)`` and ``reduce(dst, t, dt)`` are what they are in the reference:
  host calls with the destination ParticleArray (whose ``.gpu`` pushes / pulls), before the
  initialize launch and after the post_loop launch of that destination; the functions an
  equation lists in ``_get_helpers_()`` become ``__device__`` functions of the module.

Not translated (raises NotImplementedError at set-up): ``loop_all``, ``initialize_pair``,
array-valued constants, nested function definitions.
"""
import ast
import inspect
import math
import os
import textwrap

from . import _lib

F64_NAMES = ('x', 'y', 'z', 'u', 'v', 'w', 'rho', 'h', 'm',
             'x0', 'y0', 'z0', 'u0', 'v0', 'w0', 'rho0')
F32_NAMES = ('p', 'cs', 'arho', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'dt_cfl', 'dt_force')
U32_NAMES = ('gid', 'tag', 'pid')
MAX_USER = 16
USER_PROP0 = 110          # B200SPH_USER0

PAIR_SYMBOLS = ('XIJ', 'RIJ', 'R2IJ', 'HIJ', 'VIJ', 'WIJ', 'DWIJ', 'RHOIJ', 'RHOIJ1', 'EPS',
                'WI', 'WJ', 'DWI', 'DWJ')
MATH_CALLS = {'sqrt': 'sqrt', 'fabs': 'fabs', 'abs': 'fabs', 'exp': 'exp', 'log': 'log',
              'sin': 'sin', 'cos': 'cos', 'tan': 'tan', 'atan2': 'atan2', 'atan': 'atan',
              'asin': 'asin', 'acos': 'acos', 'sinh': 'sinh', 'cosh': 'cosh', 'tanh': 'tanh',
              'pow': 'pow', 'floor': 'floor', 'ceil': 'ceil', 'max': 'fmax', 'min': 'fmin',
              'log10': 'log10', 'erf': 'erf'}
MATH_CONSTS = {'M_PI': math.pi, 'pi': math.pi, 'M_1_PI': 1.0 / math.pi, 'M_2_SQRTPI': 2.0 / math.sqrt(math.pi),
               'M_PI_2': math.pi / 2, 'M_E': math.e, 'INFINITY': float('inf')}

# the struct every generated kernel takes by value: ONE text, shared with the library
ARGS_STRUCT = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'csrc',
                                'generic_args.h')).read()


class Translator(ast.NodeVisitor):
    """One method body -> C statements."""

    def __init__(self, eq, method, props, where, helpers=()):
        self.eq, self.method, self.props, self.where = eq, method, props, where
        self.helpers = tuple(helpers)          # names of the module's __device__ helper functions
        self.locals = {}          # name -> C declaration
        self.symbols = set()      # pair symbols used
        self.writes = set()       # property names stored to
        self.args = []

    def fail(self, node, msg):
        raise NotImplementedError('B200 generic equations: %s.%s line %d: %s' % (
            self.eq.__class__.__name__, self.method, getattr(node, 'lineno', 0), msg))

    # ---- expressions -------------------------------------------------------
    def expr(self, n, integer=False):
        if isinstance(n, ast.Constant):
            if isinstance(n.value, bool):
                return '1' if n.value else '0'
            if isinstance(n.value, int):
                return repr(n.value) if integer else '%d.0' % n.value
            if isinstance(n.value, float):
                return _lit(n.value)
            self.fail(n, 'constant %r' % (n.value,))
        if isinstance(n, ast.Name):
            if n.id in MATH_CONSTS:
                return _lit(MATH_CONSTS[n.id])
            if n.id in ('d_idx', 's_idx', 't', 'dt'):
                return n.id
            if n.id in PAIR_SYMBOLS:
                self.symbols.add(n.id)
                return n.id
            if n.id in ('True', 'False'):
                return '1' if n.id == 'True' else '0'
            if n.id in self.locals or n.id in self.args:
                return n.id
            self.fail(n, 'unknown name %r' % n.id)
        if isinstance(n, ast.Attribute):
            if isinstance(n.value, ast.Name) and n.value.id == 'self':
                v = getattr(self.eq, n.attr, None)
                if isinstance(v, bool):
                    return '1' if v else '0'
                if isinstance(v, int):
                    return repr(v) if integer else '%d.0' % v
                if isinstance(v, float):
                    return _lit(v)
                self.fail(n, 'self.%s is %r: only numbers can be baked into the kernel' % (n.attr, type(v).__name__))
            if isinstance(n.value, ast.Name) and n.value.id in ('math', 'np', 'numpy') and n.attr in MATH_CONSTS:
                return _lit(MATH_CONSTS[n.attr])
            self.fail(n, 'attribute access')
        if isinstance(n, ast.Subscript):
            return self.subscript(n, store=False)
        if isinstance(n, ast.BinOp):
            a, b = self.expr(n.left, integer), self.expr(n.right, integer)
            if isinstance(n.op, ast.Pow):
                return 'pow((double)(%s), (double)(%s))' % (a, b)
            if isinstance(n.op, ast.FloorDiv):
                return '((%s) / (%s))' % (a, b) if integer else 'floor((%s) / (%s))' % (a, b)
            if isinstance(n.op, ast.Mod):
                return '((%s) %% (%s))' % (a, b) if integer else 'fmod((%s), (%s))' % (a, b)
            ops = {ast.Add: '+', ast.Sub: '-', ast.Mult: '*', ast.Div: '/'}
            if type(n.op) not in ops:
                self.fail(n, 'operator %s' % type(n.op).__name__)
            return '(%s %s %s)' % (a, ops[type(n.op)], b)
        if isinstance(n, ast.UnaryOp):
            v = self.expr(n.operand, integer)
            if isinstance(n.op, ast.USub):
                return '(-%s)' % v
            if isinstance(n.op, ast.UAdd):
                return v
            if isinstance(n.op, ast.Not):
                return '(!(%s))' % v
            self.fail(n, 'unary operator')
        if isinstance(n, ast.Compare):
            ops = {ast.Lt: '<', ast.LtE: '<=', ast.Gt: '>', ast.GtE: '>=', ast.Eq: '==', ast.NotEq: '!='}
            parts, left = [], n.left
            for op, right in zip(n.ops, n.comparators):
                if type(op) not in ops:
                    self.fail(n, 'comparison %s' % type(op).__name__)
                parts.append('(%s %s %s)' % (self.expr(left), ops[type(op)], self.expr(right)))
                left = right
            return '(%s)' % ' && '.join(parts)
        if isinstance(n, ast.BoolOp):
            j = ' && ' if isinstance(n.op, ast.And) else ' || '
            return '(%s)' % j.join(self.expr(v) for v in n.values)
        if isinstance(n, ast.IfExp):
            return '((%s) ? (%s) : (%s))' % (self.expr(n.test), self.expr(n.body), self.expr(n.orelse))
        if isinstance(n, ast.Call):
            f = n.func
            name = f.id if isinstance(f, ast.Name) else (f.attr if isinstance(f, ast.Attribute) else None)
            if name in MATH_CALLS and not n.keywords:
                # float properties meet double locals: no overload ambiguity for the device compiler
                args = ['(double)(%s)' % self.expr(a) for a in n.args]
                if name in ('max', 'min') and len(args) > 2:
                    out = args[0]
                    for a in args[1:]:
                        out = '%s(%s, %s)' % (MATH_CALLS[name], out, a)
                    return out
                return '%s(%s)' % (MATH_CALLS[name], ', '.join(args))
            if name in self.helpers and isinstance(f, ast.Name) and not n.keywords:
                return '%s(%s)' % (name, ', '.join('(double)(%s)' % self.expr(a) for a in n.args))
            if name == 'float' and len(n.args) == 1:
                return '((double)(%s))' % self.expr(n.args[0])
            if name == 'int' and len(n.args) == 1:
                return '((long long)(%s))' % self.expr(n.args[0])
            self.fail(n, 'call of %r' % name)
        self.fail(n, 'expression %s' % type(n).__name__)

    def subscript(self, n, store):
        if not isinstance(n.value, ast.Name):
            self.fail(n, 'subscript of an expression')
        base = n.value.id
        idx = n.slice
        if isinstance(idx, ast.Index):        # python < 3.9
            idx = idx.value
        i = self.expr(idx, integer=True)
        if base in PAIR_SYMBOLS:
            self.symbols.add(base)
            return '%s[%s]' % (base, i)
        if base in self.locals:
            return '%s[%s]' % (base, i)
        if base[:2] in ('d_', 's_') and base in self.args:
            prop = base[2:]
            if store:
                if base[0] == 's':
                    self.fail(n, 'a loop body may not write to a source property')
                self.writes.add(prop)
            return '%s[%s]' % (self.props.pointer(prop), i)
        self.fail(n, 'subscript of %r' % base)

    # ---- statements --------------------------------------------------------
    def block(self, stmts, ind):
        out = []
        for s in stmts:
            out.extend(self.stmt(s, ind))
        return out

    def target(self, t):
        if isinstance(t, ast.Name):
            if t.id in self.args or t.id in PAIR_SYMBOLS:
                self.fail(t, 'assignment to the argument %r' % t.id)
            if t.id not in self.locals:
                self.locals[t.id] = 'double %s = 0.0;' % t.id
            return t.id
        if isinstance(t, ast.Subscript):
            return self.subscript(t, store=True)
        self.fail(t, 'assignment target %s' % type(t).__name__)

    def stmt(self, s, ind):
        p = '    ' * ind
        if isinstance(s, ast.Expr):
            if isinstance(s.value, ast.Constant):      # docstring
                return []
            self.fail(s, 'expression statement')
        if isinstance(s, ast.Pass):
            return []
        if isinstance(s, ast.Assign):
            if len(s.targets) != 1:
                self.fail(s, 'chained assignment')
            t = s.targets[0]
            v = s.value
            if isinstance(v, ast.Call) and isinstance(v.func, ast.Name) and v.func.id == 'declare':
                return self.declare(s, t, v)
            if isinstance(t, ast.Tuple):
                if not isinstance(v, ast.Tuple) or len(v.elts) != len(t.elts):
                    self.fail(s, 'tuple assignment')
                vals = [self.expr(e) for e in v.elts]
                tmp = ['const double _t%d_%d = %s;' % (s.lineno, k, e) for k, e in enumerate(vals)]
                return [p + x for x in tmp] + [p + '%s = _t%d_%d;' % (self.target(e), s.lineno, k)
                                               for k, e in enumerate(t.elts)]
            rhs = self.expr(v)
            return [p + '%s = %s;' % (self.target(t), rhs)]
        if isinstance(s, ast.AugAssign):
            ops = {ast.Add: '+=', ast.Sub: '-=', ast.Mult: '*=', ast.Div: '/='}
            if type(s.op) not in ops:
                self.fail(s, 'augmented assignment %s' % type(s.op).__name__)
            rhs = self.expr(s.value)
            return [p + '%s %s %s;' % (self.target(s.target), ops[type(s.op)], rhs)]
        if isinstance(s, ast.If):
            out = [p + 'if (%s) {' % self.expr(s.test)] + self.block(s.body, ind + 1)
            if s.orelse:
                out += [p + '} else {'] + self.block(s.orelse, ind + 1)
            return out + [p + '}']
        if isinstance(s, ast.For):
            if not (isinstance(s.target, ast.Name) and isinstance(s.iter, ast.Call) and
                    isinstance(s.iter.func, ast.Name) and s.iter.func.id == 'range' and not s.orelse):
                self.fail(s, 'only "for i in range(...)" loops')
            r = [self.expr(a, integer=True) for a in s.iter.args]
            lo, hi, st = ('0', r[0], '1') if len(r) == 1 else (r[0], r[1], r[2] if len(r) > 2 else '1')
            v = s.target.id
            self.locals[v] = 'long long %s = 0;' % v
            return [p + 'for (%s = %s; %s < %s; %s += %s) {' % (v, lo, v, hi, v, st)] + \
                self.block(s.body, ind + 1) + [p + '}']
        if isinstance(s, ast.Return):
            if self.where == 'helper':
                return [p + 'return %s;' % (self.expr(s.value) if s.value is not None else '0.0')]
            if s.value is not None:
                self.fail(s, 'return with a value')
            return [p + 'return;']         # the body is a lambda
        self.fail(s, 'statement %s' % type(s).__name__)

    def declare(self, s, t, v):
        """``x = declare('matrix(3)')``, ``i, j = declare('int', 2)`` (equation.py: the
        reference's type declarations for its C translation)."""
        if isinstance(t, ast.Tuple) and all(isinstance(e, ast.Name) for e in t.elts):
            names = [e.id for e in t.elts]
        elif isinstance(t, ast.Name):
            names = [t.id]
        else:
            self.fail(s, 'declare(...) target')
        if not (v.args and isinstance(v.args[0], ast.Constant) and isinstance(v.args[0].value, str)):
            self.fail(s, 'declare(...)')
        kind = v.args[0].value.replace(' ', '')
        for nme in names:
            if kind.startswith('matrix('):
                size = 1
                for d in kind[len('matrix('):-1].strip('()').split(','):
                    if d:
                        size *= int(d)
                self.locals[nme] = 'double %s[%d] = {0.0};' % (nme, size)
            elif kind in ('double', 'float'):
                self.locals[nme] = 'double %s = 0.0;' % nme
            elif kind in ('int', 'long', 'unsignedint'):
                self.locals[nme] = 'long long %s = 0;' % nme
            else:
                self.fail(s, 'declare(%r)' % kind)
        return []

    def translate(self):
        fn = getattr(self.eq, self.method)
        try:
            src = textwrap.dedent(inspect.getsource(fn))
        except (OSError, TypeError):
            raise NotImplementedError('B200 generic equations: no Python source for %s.%s' % (
                self.eq.__class__.__name__, self.method))
        node = ast.parse(src).body[0]
        self.args = [a.arg for a in node.args.args if a.arg != 'self']
        for a in self.args:
            if a in ('d_idx', 's_idx', 't', 'dt') or a in PAIR_SYMBOLS:
                if a in PAIR_SYMBOLS:
                    if self.where != 'loop':
                        self.fail(node, 'pair symbol %s outside loop()' % a)
                    self.symbols.add(a)
                if a == 's_idx' and self.where != 'loop':
                    self.fail(node, 's_idx outside loop()')
            elif a[:2] == 'd_' or (a[:2] == 's_' and self.where == 'loop'):
                self.props.pointer(a[2:])
            else:
                self.fail(node, 'argument %r is not a d_/s_ property, an index or a pair symbol' % a)
        body = self.block(node.body, 2)
        decl = ['        ' + d for d in self.locals.values()]
        return ['    [&]() {'] + decl + body + ['    }();']


def translate_helper(fn, helpers):
    """A plain Python function listed by Equation._get_helpers_() -> a __device__ function
    (double arguments, double result; default values kept)."""
    try:
        node = ast.parse(textwrap.dedent(inspect.getsource(fn))).body[0]
    except (OSError, TypeError):
        raise NotImplementedError('B200 generic equations: no Python source for helper %r' % fn)

    class helper(object):           # (names the function in the translator's error messages)
        pass
    t = Translator(helper(), fn.__name__, None, 'helper', helpers)
    args = [a.arg for a in node.args.args]
    t.args = args
    defaults = [None] * (len(args) - len(node.args.defaults)) + list(node.args.defaults)
    sig = []
    for a, d in zip(args, defaults):
        sig.append('double %s%s' % (a, '' if d is None else ' = ' + t.expr(d)))
    body = t.block(node.body, 1)
    decl = ['    ' + d for d in t.locals.values()]
    return '\n'.join(['__device__ static inline double %s(%s)' % (fn.__name__, ', '.join(sig)), '{'] +
                     decl + body + ['    return 0.0;', '}', ''])


def _lit(v):
    if v != v:
        return '(0.0/0.0)'
    if v in (float('inf'), float('-inf')):
        return '(%s1.0/0.0)' % ('-' if v < 0 else '')
    r = repr(float(v))
    return r if ('.' in r or 'e' in r or 'E' in r) else r + '.0'


class PropertyTable(object):
    """Property name -> pointer expression inside the kernel; names the pool does not have
    become user properties, in order of first mention."""

    def __init__(self, user_names=None):
        self.user = list(user_names or [])
        self.used = set()          # every property name a generated body mentions

    def pointer(self, name):
        self.used.add(name)
        if name in F64_NAMES:
            return 'a.f64[%d]' % F64_NAMES.index(name)
        if name in F32_NAMES:
            return 'a.f32[%d]' % F32_NAMES.index(name)
        if name in U32_NAMES:
            return 'a.u32[%d]' % U32_NAMES.index(name)
        if name not in self.user:
            if len(self.user) >= MAX_USER:
                raise NotImplementedError('B200 generic equations: more than %d properties beyond the '
                                          'built-in ones' % MAX_USER)
            self.user.append(name)
        return 'a.user[%d]' % self.user.index(name)


KERNEL_FUNCS = r'''
__device__ static inline double b2_w(const int K, const int DIM, const double fac, const double rij, const double h)
{
    const double h1 = 1.0 / h, q = rij * h1;
    double s = fac * h1;
    if (DIM > 1) s *= h1;
    if (DIM > 2) s *= h1;
    double v = 0.0;
    if (K == 0) {          /* CubicSpline kernels.py:69-136 */
        if (q > 2.0) v = 0.0;
        else if (q > 1.0) { const double t = 2.0 - q; v = 0.25 * t * t * t; }
        else v = 1.0 - 1.5 * q * q * (1.0 - 0.5 * q);
    } else if (K == 1) {   /* WendlandQuintic kernels.py:304-358 */
        if (q < 2.0) {
            double t = 1.0 - 0.5 * q; t *= t; t *= t;
            v = t * (2.0 * q + 1.0);
        }
    } else if (K == 2) {   /* QuinticSpline kernels.py:1088-1168 */
        const double t1 = 3.0 - q, t2 = 2.0 - q, t3 = 1.0 - q;
        const double p1 = t1 * t1 * t1 * t1 * t1, p2 = t2 * t2 * t2 * t2 * t2, p3 = t3 * t3 * t3 * t3 * t3;
        if (q > 3.0) v = 0.0;
        else if (q > 2.0) v = p1;
        else if (q > 1.0) v = p1 - 6.0 * p2;
        else v = p1 - 6.0 * p2 + 15.0 * p3;
    } else {               /* Gaussian kernels.py:830-905 */
        if (q < 3.0) v = exp(-q * q);
    }
    return s * v;
}
__device__ static inline double b2_dwdq(const int K, const int DIM, const double fac, const double rij, const double h)
{
    const double h1 = 1.0 / h, q = rij * h1;
    double s = fac * h1;
    if (DIM > 1) s *= h1;
    if (DIM > 2) s *= h1;
    double v = 0.0;
    if (!(rij > 1e-12)) return 0.0;
    if (K == 0) {
        if (q > 2.0) v = 0.0;
        else if (q > 1.0) { const double t = 2.0 - q; v = -0.75 * t * t; }
        else v = -3.0 * q * (1.0 - 0.75 * q);
    } else if (K == 1) {
        if (q < 2.0) { const double t = 1.0 - 0.5 * q; v = -5.0 * q * t * t * t; }
    } else if (K == 2) {
        const double t1 = 3.0 - q, t2 = 2.0 - q, t3 = 1.0 - q;
        const double p1 = -5.0 * t1 * t1 * t1 * t1, p2 = 30.0 * t2 * t2 * t2 * t2, p3 = -75.0 * t3 * t3 * t3 * t3;
        if (q > 3.0) v = 0.0;
        else if (q > 2.0) v = p1;
        else if (q > 1.0) v = p1 + p2;
        else v = p1 + p2 + p3;
    } else {
        if (q < 3.0) v = -2.0 * q * exp(-q * q);
    }
    return s * v;
}
__device__ static inline void b2_grad(const int K, const int DIM, const double fac, const double *xij, const double rij,
                                      const double h, double *g)
{
    double w = 0.0;
    if (rij > 1e-12) w = b2_dwdq(K, DIM, fac, rij, h) / (h * rij);
    g[0] = w * xij[0]; g[1] = w * xij[1]; g[2] = w * xij[2];
}
'''


class GenericGroup(object):
    """The kernels of one Group of untranslated equations."""

    def __init__(self, group, array_index, kernel, dim, props, uid):
        from .kernels import kernel_id
        self.group = group
        self.index = array_index
        self.real_only = 1 if getattr(group, 'real', True) else 0
        self.kid, self.dim, self.fac = kernel_id(kernel), int(dim), float(kernel.fac)
        self.props = props
        self.uid = uid
        for eq in group.equations:
            for bad in ('loop_all', 'initialize_pair'):
                if _defines(eq, bad):
                    raise NotImplementedError('B200 generic equations: %s.%s() is not translated' % (
                        eq.__class__.__name__, bad))
            if eq.dest not in array_index:
                raise ValueError('equation %s: unknown destination array %r' % (eq.__class__.__name__, eq.dest))
            for s in (eq.sources or []):
                if s not in array_index:
                    raise ValueError('equation %s: unknown source array %r' % (eq.__class__.__name__, s))
        # destinations in order of first mention (Group.data, equation.py:600-640)
        self.dests = []
        for eq in group.equations:
            if eq.dest not in self.dests:
                self.dests.append(eq.dest)
        self.kernels = []          # (name, dest array id, has_init, has_loop, has_post, src_mask)
        self.host_calls = {}       # dest array id -> (equations with py_initialize, with reduce)
        self.writes = set()
        # helper functions (Equation._get_helpers_, equation.py:389-420), once each
        self.helpers = []
        for eq in group.equations:
            for fn in (eq._get_helpers_() if hasattr(eq, '_get_helpers_') else []):
                if fn.__name__ not in [h.__name__ for h in self.helpers]:
                    self.helpers.append(fn)
        self.source = self._generate()

    def _generate(self):
        out = ['/* generated by pysph_b200.codegen for Group %s */' % getattr(self.group, 'name', '?'),
               '#define PT_GHOST 0x08', ARGS_STRUCT, KERNEL_FUNCS]
        hnames = [h.__name__ for h in self.helpers]
        for k, h in enumerate(self.helpers):
            out.append(translate_helper(h, hnames[:k + 1]))
        for d in self.dests:
            eqs = [e for e in self.group.equations if e.dest == d]
            name = 'b2g_%s' % d          # one module per Group: no clash between Groups
            init, loop, post = [], [], []
            symbols = set()
            src_mask = 0
            for e in eqs:
                if _defines(e, 'initialize'):
                    t = Translator(e, 'initialize', self.props, 'init', hnames)
                    init += ['    /* %s.initialize */' % e.__class__.__name__] + t.translate()
                    self.writes |= t.writes
                if _defines(e, 'loop'):
                    if not e.sources:
                        raise NotImplementedError('B200 generic equations: %s has a loop() but no sources'
                                                  % e.__class__.__name__)
                    t = Translator(e, 'loop', self.props, 'loop', hnames)
                    m = 0
                    for s in e.sources:
                        m |= 1 << self.index[s]
                    src_mask |= m
                    body = t.translate()
                    loop += ['    /* %s.loop */' % e.__class__.__name__,
                             '    if ((0x%xu >> tj) & 1u) {' % m] + body + ['    }']
                    symbols |= t.symbols
                    self.writes |= t.writes
                if _defines(e, 'post_loop'):
                    t = Translator(e, 'post_loop', self.props, 'post', hnames)
                    post += ['    /* %s.post_loop */' % e.__class__.__name__] + t.translate()
                    self.writes |= t.writes
            out.append(self._kernel(name, init, loop, post, symbols))
            self.kernels.append((name, self.index[d], bool(init), bool(loop), bool(post), src_mask))
            self.host_calls[self.index[d]] = ([e for e in eqs if _defines(e, 'py_initialize')],
                                              [e for e in eqs if _defines(e, 'reduce')])
        return '\n'.join(out) + '\n'

    def _kernel(self, name, init, loop, post, symbols):
        K, DIM, FAC = self.kid, self.dim, _lit(self.fac)
        sym = []
        if 'VIJ' in symbols:
            sym.append('        const double VIJ[3] = {a.f64[3][d_idx] - a.f64[3][s_idx], a.f64[4][d_idx] - a.f64[4][s_idx], '
                       'a.f64[5][d_idx] - a.f64[5][s_idx]};')
        if 'RHOIJ' in symbols or 'RHOIJ1' in symbols:
            sym.append('        const double RHOIJ = 0.5 * (a.f64[6][d_idx] + a.f64[6][s_idx]); const double RHOIJ1 = 1.0 / RHOIJ;')
        if 'EPS' in symbols:
            sym.append('        const double EPS = 0.01 * HIJ * HIJ;')
        if 'WIJ' in symbols:
            sym.append('        const double WIJ = b2_w(%d, %d, %s, RIJ, HIJ);' % (K, DIM, FAC))
        if 'DWIJ' in symbols:
            sym.append('        double DWIJ[3]; b2_grad(%d, %d, %s, XIJ, RIJ, HIJ, DWIJ);' % (K, DIM, FAC))
        if 'WI' in symbols:
            sym.append('        const double WI = b2_w(%d, %d, %s, RIJ, (double)Ai.w);' % (K, DIM, FAC))
        if 'WJ' in symbols:
            sym.append('        const double WJ = b2_w(%d, %d, %s, RIJ, (double)Aj.w);' % (K, DIM, FAC))
        if 'DWI' in symbols:
            sym.append('        double DWI[3]; b2_grad(%d, %d, %s, XIJ, RIJ, (double)Ai.w, DWI);' % (K, DIM, FAC))
        if 'DWJ' in symbols:
            sym.append('        double DWJ[3]; b2_grad(%d, %d, %s, XIJ, RIJ, (double)Aj.w, DWJ);' % (K, DIM, FAC))
        ind = lambda lines: ['    ' + l for l in lines]
        return '\n'.join([
            'extern "C" __global__ void %s(const b200sph_generic_args a)' % name,
            '{',
            '    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;',
            '    if (s >= a.n) return;',
            '    const int ti = a.stype[s];',
            '    if ((ti & 7) != a.dest_type) return;',
            '    if (a.real_only && (ti & PT_GHOST)) return;',
            '    const long long d_idx = a.perm[s];',
            '    if (d_idx - a.doff < a.dlo || d_idx - a.doff >= a.dhi) return;',
            '    const double t = a.t, dt = a.dt; (void)t; (void)dt;',
            '    if (a.phase == 0) {'] + ind(init) + [
            '        return;',
            '    }',
            '    if (a.phase == 2) {'] + ind(post) + [
            '        return;',
            '    }',
            '    const float4 Ai = a.AB[2 * s];',
            '    const float hi2 = a.k2 * Ai.w * Ai.w;',
            '    const unsigned int cnt = a.cnt[s];',
            '    const unsigned int *nxt = a.lst + ((size_t)(s >> 5) * (size_t)a.capg) * 32u + (unsigned int)(s & 31);',
            '    for (unsigned int k = 0; k < cnt; k++) {',
            '        const unsigned int e = nxt[(size_t)k * 32u];',
            '        const unsigned int j = e >> 6;',
            '        const float4 Aj = a.AB[2 * (size_t)j];',
            '        const float xf = Ai.x - Aj.x - (float)((int)(e & 3u) - 1) * a.cellx;',
            '        const float yf = Ai.y - Aj.y - (float)((int)((e >> 2) & 3u) - 1) * a.celly;',
            '        const float zf = Ai.z - Aj.z - (float)((int)((e >> 4) & 3u) - 1) * a.cellz;',
            '        const float r2f = xf * xf + yf * yf + zf * zf;',
            '        /* the exact accept test, linked_list_nnps.pyx:188 */',
            '        if (!((r2f < hi2) || (r2f < a.k2 * Aj.w * Aj.w))) continue;',
            '        const int tj = a.stype[j] & 7;',
            '        if (!((a.src_mask >> tj) & 1u)) continue;',
            '        const long long s_idx = a.perm[j];',
            '        const double XIJ[3] = {(double)xf, (double)yf, (double)zf};',
            '        const double R2IJ = XIJ[0] * XIJ[0] + XIJ[1] * XIJ[1] + XIJ[2] * XIJ[2];',
            '        const double RIJ = sqrt(R2IJ);',
            '        const double HIJ = 0.5 * ((double)Ai.w + (double)Aj.w);',
            '        (void)RIJ; (void)HIJ; (void)s_idx;'] + sym + ind(loop) + [
            '    }',
            '}', ''])


def _defines(eq, method):
    """True if the equation's class gives the method a body of its own (the Equation base
    classes -- the reference's and ours -- define none of them except converged())."""
    fn = getattr(type(eq), method, None)
    if fn is None:
        return False
    for klass in type(eq).__mro__:
        if method in klass.__dict__:
            return klass.__name__ != 'Equation'
    return False


def is_generic_group(group):
    """A Group every equation of which carries its own Python bodies."""
    eqs = group.equations
    return len(eqs) > 0 and all(any(_defines(e, m) for m in ('initialize', 'loop', 'post_loop', 'reduce',
                                                             'py_initialize')) for e in eqs)


# ---- NVRTC -------------------------------------------------------------------
_IMAGES = {}


def compile_cached(source):
    """compile_image once per distinct source text (equal Groups share the cubin)."""
    if source not in _IMAGES:
        _IMAGES[source] = compile_image(source)
    return _IMAGES[source]


def compile_image(source, name='b200sph_generic.cu', arch='sm_100a'):
    """CUDA C -> cubin for sm_100a through NVRTC (no GPU needed to compile)."""
    try:
        from cuda.bindings import nvrtc
    except ImportError:                                  # older cuda-python layout
        from cuda import nvrtc
    full = source

    def chk(res):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError('NVRTC: %s' % nvrtc.nvrtcGetErrorString(err)[1].decode())
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    prog = chk(nvrtc.nvrtcCreateProgram(full.encode(), name.encode(), 0, [], []))
    opts = [b'--gpu-architecture=' + arch.encode(), b'-lineinfo', b'--std=c++17']
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        n = chk(nvrtc.nvrtcGetProgramLogSize(prog))
        log = b' ' * n
        chk(nvrtc.nvrtcGetProgramLog(prog, log))
        raise RuntimeError('NVRTC could not compile the generated kernel:\n%s\n---- source ----\n%s'
                           % (log.decode(errors='replace'), source))
    n = chk(nvrtc.nvrtcGetCUBINSize(prog))
    image = b' ' * n
    chk(nvrtc.nvrtcGetCUBIN(prog, image))
    nvrtc.nvrtcDestroyProgram(prog)
    return image

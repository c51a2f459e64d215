"""SPH smoothing-kernel descriptors.

Host-side parameter holders with the same names, constructor arguments and
attributes (``dim``, ``fac``, ``radius_scale``, ``get_deltap()``) as
pysph/base/kernels.py (CubicSpline :29-163, WendlandQuintic :274-380,
Gaussian :830-941, QuinticSpline :1050-1210).  The arithmetic itself lives in
the CUDA ``__device__`` functions of csrc/b200sph.cu; the adapters recognise a
kernel by its class name, so a real ``pysph.base.kernels`` object works too.
"""
from math import pi, sqrt

# ids shared with include/b200sph.h (B200SPH_KERNEL_*)
KERNEL_IDS = {'CubicSpline': 0, 'WendlandQuintic': 1, 'QuinticSpline': 2,
              'Gaussian': 3}


class _Kernel(object):
    radius_scale = 2.0
    _deltap = 0.0

    def get_deltap(self):
        return self._deltap


class CubicSpline(_Kernel):
    _deltap = 2. / 3

    def __init__(self, dim=1):
        self.radius_scale = 2.0
        self.dim = dim
        self.fac = {3: 1.0 / pi, 2: 10.0 / (7.0 * pi)}.get(dim, 2.0 / 3.0)


class WendlandQuintic(_Kernel):
    _deltap = 0.5

    def __init__(self, dim=2):
        if dim == 1:
            raise ValueError("WendlandQuintic: Dim %d not supported" % dim)
        self.radius_scale = 2.0
        self.dim = dim
        self.fac = 7.0 / (4.0 * pi) if dim == 2 else 21.0 / (16.0 * pi)


class QuinticSpline(_Kernel):
    _deltap = 0.759298480738450

    def __init__(self, dim=2):
        self.radius_scale = 3.0
        self.dim = dim
        self.fac = {1: 1.0 / 120.0, 2: 7.0 / (478.0 * pi)}.get(
            dim, 1.0 / (120.0 * pi))


class Gaussian(_Kernel):
    _deltap = 0.70710678118654746

    def __init__(self, dim=2):
        self.radius_scale = 3.0
        self.dim = dim
        self.fac = (1.0 / sqrt(pi)) ** dim


def kernel_id(kernel):
    """Map a kernel object (ours or PySPH's) to the C-ABI kernel id."""
    name = kernel.__class__.__name__
    if name not in KERNEL_IDS:
        raise NotImplementedError(
            'B200 backend: unsupported smoothing kernel %r (supported: %s)'
            % (name, ', '.join(sorted(KERNEL_IDS))))
    return KERNEL_IDS[name]

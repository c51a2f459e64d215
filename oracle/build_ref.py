"""Build the reference's OWN compiled smoothing kernels into oracle/_ref/.

TEST INFRASTRUCTURE.  Compiles /root/reference/pysph/base/c_kernels.pyx (the
committed, generated Cython file the reference ships; only libc.math + numpy)
from where it lies -- no reference source is copied into this repository, the
generated C++ and the extension module land in oracle/_ref/ (git-ignored).
The rest of the reference path cannot be built here: every other .pyx cimports
``cyarray`` (pysph/base/nnps_base.pxd:13) and the generated evaluator needs
``mako`` + ``compyle``, none of which is installed and there is no network.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PYSPH_REFERENCE', '/root/reference')
OUT = os.path.join(HERE, '_ref')


def build():
    """c_kernels (smoothing kernels) and linalg3 (the 3x3 symmetric eigen solver the
    solid-mechanics equations call, pysph/base/linalg3.pyx: libc + numpy only)."""
    out = None
    for mod in ('c_kernels', 'linalg3'):
        out = _build_one(mod) or out
    return out


def _build_one(mod):
    pyx = os.path.join(REF, 'pysph', 'base', mod + '.pyx')
    if not os.path.exists(pyx):
        print('build_ref: %s not found, nothing to do' % pyx)
        return None
    os.makedirs(OUT, exist_ok=True)
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    so = os.path.join(OUT, mod + ext)
    if os.path.exists(so) and os.path.getmtime(so) > os.path.getmtime(pyx):
        return so
    import numpy
    cpp = os.path.join(OUT, mod + '.cpp')
    subprocess.check_call([sys.executable, '-m', 'cython', '-3', '--cplus',
                           '-I', os.path.join(REF, 'pysph', 'base'),
                           pyx, '-o', cpp])
    cxx = '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'
    subprocess.check_call([
        cxx, '-O2', '-shared', '-fPIC', '-w',
        '-I', sysconfig.get_paths()['include'], '-I', numpy.get_include(),
        cpp, '-o', so])
    return so


if __name__ == '__main__':
    print(build())

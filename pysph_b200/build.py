"""Build libb200sph.so in-tree with nvcc for sm_100a (B200).

    python -m pysph_b200.build [--force] [--verbose]

The shared library is a plain C-ABI (include/b200sph.h); no torch types, no
pybind.  It is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'csrc', 'b200sph.cu')
HDR = os.path.join(ROOT, 'include', 'b200sph.h')
OUT = os.path.join(HERE, 'libb200sph.so')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-shared', '-Xcompiler', '-fPIC',
    '-Xptxas', '-v',
    '--use_fast_math' if False else '-DB200SPH_NO_FAST_MATH',
]


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'),
                 '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libb200sph.so')


def sources():
    """b200sph.cu and the kernel files it includes (one translation unit)."""
    d = os.path.dirname(SRC)
    return [SRC] + sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(('.cuh', '.h')))


def read_source():
    """The translation unit as one text: b200sph.cu with its own `#include "x.cuh"`
    lines replaced by the files (what the compiler sees; the CPU tests that compile
    the kernel source for the host read it through this)."""
    d = os.path.dirname(SRC)
    out = []
    for line in open(SRC).read().split('\n'):
        name = line[len('#include "'):-1] if line.startswith('#include "') else ''
        if name.endswith(('.cuh', '.h')) and os.path.exists(os.path.join(d, name)):
            out.append(open(os.path.join(d, name)).read().rstrip('\n'))
        else:
            out.append(line)
    return '\n'.join(out)


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in sources() + [HDR, __file__])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [find_nvcc()] + NVCC_FLAGS + ['-I', os.path.join(ROOT, 'include'),
                                        '-o', OUT, SRC]
    env = dict(os.environ)
    # the image exports CC=/opt/gcc/bin/gcc whose driver lacks libgomp specs;
    # nvcc only needs a host C++ compiler
    if os.path.exists('/usr/bin/g++'):
        cmd[1:1] = ['-ccbin', '/usr/bin/g++']
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, env=env)
    log = os.path.join(HERE, 'build.log')
    with open(log, 'w') as f:
        f.write(' '.join(cmd) + '\n' + res.stdout)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed (see %s)' % log)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))

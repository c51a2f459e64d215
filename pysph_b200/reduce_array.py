"""``serial_reduce_array`` / ``parallel_reduce_array`` for ``Equation.reduce`` bodies
(pysph/base/reduce_array.py:14-72): reductions of a HOST property array -- an equation that
needs device data pulls it first (``dst.gpu.pull('m')``), as with the reference's GPU back ends.
One process here, so the parallel variant is the serial one (the reference's
``dummy_reduce_array``); with a slab decomposition combine the per-rank results yourself."""
import numpy as np

_OPS = {'sum': np.sum, 'prod': np.prod, 'min': np.min, 'max': np.max}


def serial_reduce_array(array, op='sum'):
    if op not in _OPS:
        raise RuntimeError("Unknown reduction operator %r, use one of %s" % (op, sorted(_OPS)))
    return _OPS[op](np.asarray(array))


def parallel_reduce_array(array, op='sum'):
    return serial_reduce_array(array, op)


dummy_reduce_array = parallel_reduce_array

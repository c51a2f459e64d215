"""Output / restart files in PySPH's own format.

``dump`` writes the version-2 ``.npz`` layout of ``pysph.solver.output.NumpyOutput``
(pysph/solver/output.py:117-160 with ``get_particles_info``, pysph/base/utils.py:
466-497), so ``pysph.solver.utils.load`` reads our files and ``load`` here reads
PySPH's:

    version      = 2
    solver_data  = {'dt': .., 't': .., 'count': ..}                (solver.py:590-600)
    particles    = {array name: {'properties': {prop: {'name', 'type', 'default',
                                                       'stride', 'data': None}},
                                 'constants': {..}, 'output_property_arrays': [..],
                                 'lb_props': [..], 'arrays': {prop: ndarray}}}

Only ``output_property_arrays`` of the real particles are written unless
``detailed_output`` (output.py:64-69); with a device backend exactly those
properties are pulled (``pa.gpu.pull``, particle_array.pyx:377-378).  HDF5 needs
h5py, which this environment does not have: ``.hdf5`` names fall back to ``.npz``
exactly like the reference (output.py:403-412).
"""
import os

import numpy as np

output_formats = ('hdf5', 'npz')

_C_TYPES = {np.dtype('float64'): 'double', np.dtype('float32'): 'float',
            np.dtype('int32'): 'int', np.dtype('uint32'): 'unsigned int',
            np.dtype('int64'): 'long', np.dtype('uint64'): 'unsigned long'}
_NP_TYPES = dict((v, k) for k, v in _C_TYPES.items())


def _c_type(arr):
    if hasattr(arr, 'get_c_type'):
        return arr.get_c_type()
    return _C_TYPES.get(np.asarray(arr).dtype, 'double')


def get_particles_info(particles):
    """pysph/base/utils.py:466-497"""
    info = {}
    for pa in particles:
        props = {}
        stride = getattr(pa, 'stride', {}) or {}
        defaults = pa.default_values
        for name, prop in pa.properties.items():
            props[name] = {'name': name, 'type': _c_type(prop),
                           'default': defaults[name],
                           'stride': stride.get(name, 1), 'data': None}
        consts = {}
        for c_name, value in pa.constants.items():
            consts[c_name] = value.get_npy_array() if hasattr(value, 'get_npy_array') \
                else np.asarray(value)
        info[pa.name] = dict(properties=props, constants=consts,
                             output_property_arrays=list(pa.output_property_arrays),
                             lb_props=pa.get_lb_props())
    return info


def npz_name(filename):
    fname = os.path.splitext(filename)[0] if filename.endswith(output_formats) \
        else filename
    return fname + '.npz'


def write_npz(filename, particle_data, solver_data, compress=False):
    """The file itself: particle_data = get_particles_info(...) with every array's
    'arrays' dict filled in (pysph/solver/output.py:246-262)."""
    save = np.savez_compressed if compress else np.savez
    save(filename, version=2, particles=particle_data, solver_data=dict(solver_data))
    return filename


def dump(filename, particles, solver_data, detailed_output=False, only_real=True,
         mpi_comm=None, compress=False):
    """pysph/solver/output.py:364-415 (the npz branch)."""
    if mpi_comm is not None:
        raise NotImplementedError('B200 backend: collected parallel output (every '
                                  'rank writes its own file, solver.py:569-571)')
    filename = npz_name(filename)
    particle_data = get_particles_info(particles)
    for pa in particles:
        arrays = pa.get_property_arrays(all=detailed_output, only_real=only_real)
        particle_data[pa.name]['arrays'] = dict(
            (k, np.array(v, copy=True)) for k, v in arrays.items())
    return write_npz(filename, particle_data, solver_data, compress)


def _to_str(s):
    return s.decode('utf-8') if isinstance(s, bytes) else str(s)


def _dict_bytes_to_str(d):
    res = {}
    for key, value in d.items():
        if isinstance(value, dict):
            value = _dict_bytes_to_str(value)
        if isinstance(value, bytes):
            value = _to_str(value)
        if isinstance(value, list) and value and isinstance(value[0], bytes):
            value = [_to_str(x) for x in value]
        res[_to_str(key)] = value
    return res


def _get_dict(arr):
    res = arr.reshape(1)[0]
    if res and isinstance(list(res.keys())[0], bytes):
        return _dict_bytes_to_str(res)
    return res


def load(fname):
    """pysph/solver/output.py:127-160, :331-353: {'arrays': {name: ParticleArray},
    'solver_data': {...}} from a version-2 (or version-1) npz file."""
    from .particle_array import ParticleArray, get_particle_array
    if fname.endswith('.hdf5'):
        raise ImportError('Install python-h5py to load this file')
    if not os.path.exists(fname):
        raise RuntimeError('File not present')
    data = np.load(fname, encoding='bytes', allow_pickle=True)
    if 'version' not in data.files:
        raise RuntimeError('Wrong file type! No version number recorded.')
    ret = {'arrays': {}, 'solver_data': _get_dict(data['solver_data'])}
    version = int(data['version'])
    if version == 1:
        arrays = _get_dict(data['arrays'])
        for name in arrays:
            ret['arrays'][name] = get_particle_array(name=name, **arrays[name])
    elif version == 2:
        particles = _get_dict(data['particles'])
        for name, info in particles.items():
            props = {}
            n = 0
            for prop, meta in info['properties'].items():
                if prop in info['arrays']:
                    a = np.asarray(info['arrays'][prop])
                    props[prop] = a
                    n = max(n, a.size // int(meta.get('stride', 1) or 1))
            pa = ParticleArray(name=name, constants=info['constants'], **props)
            # properties that were not written come back with their defaults
            for prop, meta in info['properties'].items():
                if prop not in pa.properties:
                    pa.add_property(prop, default=meta.get('default', 0))
                want = _NP_TYPES.get(meta.get('type'), None)
                if want is not None and pa.properties[prop].dtype != want:
                    pa.properties[prop] = pa.properties[prop].astype(want)
            pa.set_output_arrays(info.get('output_property_arrays', []))
            ret['arrays'][name] = pa
    else:
        raise RuntimeError('Version not understood!')
    return ret

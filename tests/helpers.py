"""Shared helpers for the parity tests (test infrastructure)."""
import json
import os

import numpy as np

from pysph_b200.particle_array import get_particle_array_wcsph

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

ACC_FIELDS = ['arho', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'dt_cfl', 'dt_force']


def load_golden(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def arrays_from_dict(d, order=('fluid', 'boundary', 'obstacle')):
    """golden 'inputs' dict -> list of stand-in ParticleArrays."""
    pas = []
    for name in order:
        if name not in d:
            continue
        a = d[name]
        props = dict((k, np.array(v, dtype=float)) for k, v in a.items()
                     if k[0] != '_')
        pa = get_particle_array_wcsph(name=name, **props)
        pa.set_num_real_particles(a.get('_n_real', len(a['x'])))
        pas.append(pa)
    return pas


def copy_arrays(pas):
    out = []
    for pa in pas:
        q = get_particle_array_wcsph(
            name=pa.name, **dict((k, v.copy()) for k, v in pa.properties.items()))
        q.set_num_real_particles(pa.num_real_particles)
        out.append(q)
    return out


def wcsph_params_from_case(case):
    p = dict(case['params'])
    p['fluids'] = ['fluid']
    p['solids'] = ['boundary', 'obstacle']
    p.setdefault('dt0', 1e-5)
    return p


def scale_of(ref, floor=0.0):
    return max(float(np.max(np.abs(ref))) if len(ref) else 0.0, floor)


def rel_err(a, b, scale=None):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    if a.size == 0:
        return 0.0
    s = scale if scale is not None else scale_of(b, 1e-300)
    return float(np.max(np.abs(a - b)) / s)


def edac_arrays_from_dict(d, order=('fluid', 'fluid2')):
    """golden 'inputs' dict of an EDAC case -> list of stand-in ParticleArrays."""
    from pysph_b200.particle_array import get_particle_array_edac
    pas = []
    for name in order:
        if name not in d:
            continue
        a = d[name]
        props = dict((k, np.array(v, dtype=float)) for k, v in a.items()
                     if k[0] != '_')
        pa = get_particle_array_edac(name=name, **props)
        pa.set_num_real_particles(a.get('_n_real', len(a['x'])))
        pas.append(pa)
    return pas


def edac_wall_arrays_from_dict(d):
    """golden 'inputs' dict of an EDAC case with a solid wall -> [fluid, wall] stand-ins."""
    from pysph_b200.particle_array import get_particle_array_edac, get_particle_array_edac_wall
    pas = []
    for name, factory in (('fluid', get_particle_array_edac), ('wall', get_particle_array_edac_wall)):
        a = d[name]
        props = dict((k, np.array(v, dtype=float)) for k, v in a.items() if k[0] != '_')
        pa = factory(name=name, **props)
        pa.set_num_real_particles(a.get('_n_real', len(a['x'])))
        pas.append(pa)
    return pas


def edac_ext_arrays_from_dict(d):
    """golden 'inputs' dict of an external-flow EDAC case -> stand-ins (fluids, then the wall)."""
    from pysph_b200.particle_array import get_particle_array_edac_ext, get_particle_array_edac_wall
    pas = []
    for name in ('fluid', 'fluid2', 'wall'):
        if name not in d:
            continue
        a = d[name]
        props = dict((k, np.array(v, dtype=float)) for k, v in a.items() if k[0] != '_')
        factory = get_particle_array_edac_wall if name == 'wall' else get_particle_array_edac_ext
        pa = factory(name=name, **props)
        pa.set_num_real_particles(a.get('_n_real', len(a['x'])))
        pas.append(pa)
    return pas


EDAC_EXT_FIELDS = ['V', 'rho', 'au', 'av', 'aw', 'ap', 'ax', 'ay', 'az']
EDAC_WALL_FIELDS = ['V', 'wij', 'p', 'uf', 'vf', 'wf', 'ug', 'vg', 'wg']
EDAC_FIELDS = ['V', 'rho', 'pavg', 'au', 'av', 'aw', 'auhat', 'avhat', 'awhat', 'ap']

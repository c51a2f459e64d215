"""Minimal host-side ParticleArray stand-in.

PySPH's real ``ParticleArray`` (pysph/base/particle_array.pyx:68-300) is a
Cython class over ``cyarray`` buffers and cannot be imported where PySPH is
not installed.  The B200 adapters in this package are duck-typed against the
small attribute surface they need (SURVEY.md 8b "Host buffers"):

    pa.name, pa.properties (dict name -> array), pa.constants,
    pa.num_real_particles, pa.get_number_of_particles(real=False),
    pa.get(name, only_real_particles=True), pa.add_property(name, ...),
    pa.gpu  (device helper installed by the backend)

This class provides exactly that over numpy arrays so the package, its tests
and bench.py run stand-alone; a real PySPH ParticleArray can be passed to
every adapter instead.
"""
import numpy as np

UINT_MAX = (1 << 32) - 1

# pysph/base/particle_array.pxd:24-27
LOCAL, REMOTE, GHOST = 0, 1, 2

# pysph/base/utils.py:41-44
DEFAULT_PROPS = ('x', 'y', 'z', 'u', 'v', 'w', 'm', 'h', 'rho', 'p',
                 'au', 'av', 'aw', 'gid', 'pid', 'tag')
# pysph/base/utils.py:177-178
WCSPH_PROPS = ('cs', 'ax', 'ay', 'az', 'arho', 'x0', 'y0', 'z0',
               'u0', 'v0', 'w0', 'rho0', 'div', 'dt_cfl', 'dt_force')

# EDACScheme.setup_properties, internal-flow branch (wc/edac.py:724-735): what the
# transport-velocity EDAC equations and EDACTVFStep read and write
EDAC_TVF_PROPS = ('uhat', 'vhat', 'what', 'ap', 'auhat', 'avhat', 'awhat', 'V',
                  'p0', 'u0', 'v0', 'w0', 'x0', 'y0', 'z0', 'pavg', 'nnbr')

_INT_PROPS = {'tag': np.int32, 'pid': np.int32, 'gid': np.uint32}


class ParticleArray(object):
    def __init__(self, name='array', constants=None, **props):
        self.name = name
        self.properties = {}
        self.constants = dict(constants or {})
        self.output_property_arrays = []
        self.gpu = None
        self.backend = 'b200'
        n = 0
        for v in props.values():
            if isinstance(v, dict):     # {'data':.., 'type':.., 'default':.., 'stride':..}
                v = v.get('data')       # (particle_array.pyx:150-175)
                if v is None:
                    continue
            v = np.asarray(v)
            if v.ndim > 0:          # scalars broadcast, they do not size
                n = max(n, v.size)
        self._n = n
        self.num_real_particles = n
        for k, v in props.items():
            if isinstance(v, dict):
                self.add_property(k, type=v.get('type'), default=v.get('default'),
                                  data=v.get('data'))
            else:
                self.add_property(k, data=v)

    # -- sizes ------------------------------------------------------------
    def get_number_of_particles(self, real=False):
        return self.num_real_particles if real else self._n

    # -- properties -------------------------------------------------------
    def add_property(self, name, type=None, default=None, data=None, stride=1):
        dtype = _INT_PROPS.get(name, np.float64)
        if type is not None:            # C type names, as get_c_type() returns them
            dtype = {'double': np.float64, 'float': np.float32, 'int': np.int32,
                     'unsigned int': np.uint32, 'long': np.int64,
                     'unsigned long': np.uint64}.get(type, dtype)
        if default is None:
            default = UINT_MAX if name == 'gid' else 0
        arr = np.full(self._n * stride, default, dtype=dtype)
        if data is not None:
            d = np.asarray(data)
            if d.ndim == 0 or (d.size == 1 and arr.size != 1):
                arr[:] = d.ravel()[0]
            else:
                if d.size != arr.size:
                    raise ValueError('property %r: size %d != %d' %
                                     (name, d.size, arr.size))
                arr[:] = d.ravel()
        self.properties[name] = arr

    def get(self, *names, **kw):
        only_real = kw.get('only_real_particles', True)
        out = []
        for nme in names:
            a = self.properties[nme] if nme in self.properties \
                else self.constants[nme]
            if only_real and nme in self.properties:
                a = a[:self.num_real_particles]
            out.append(a)
        return out[0] if len(out) == 1 else out

    def set(self, **props):
        for k, v in props.items():
            self.properties[k][:] = v

    def __getattr__(self, name):
        # pa.x style access, like the reference (particle_array.pyx:767-770)
        props = self.__dict__.get('properties', {})
        if name in props:
            return props[name]
        consts = self.__dict__.get('constants', {})
        if name in consts:
            return consts[name]
        raise AttributeError(name)

    def set_output_arrays(self, names):
        self.output_property_arrays = list(names)

    # -- what the output path reads (particle_array.pyx:344-386, :414-421;
    #    utils.py:466-497) -----------------------------------------------------
    stride = {}
    lb_props = None

    @property
    def default_values(self):
        return dict((k, UINT_MAX if k == 'gid' else 0) for k in self.properties)

    def get_lb_props(self):
        return list(self.properties.keys()) if self.lb_props is None else self.lb_props

    def get_property_arrays(self, all=True, only_real=True):
        props = self.output_property_arrays
        if all or len(props) == 0:
            props = list(self.properties.keys())
        if self.gpu is not None:
            self.gpu.pull(*props)            # only what is written travels D2H
        n = self.get_number_of_particles(only_real)
        return dict((p, self.properties[p][:n]) for p in props)

    def add_constant(self, name, data):
        self.constants[name] = np.atleast_1d(np.asarray(data, dtype=float))

    def set_num_real_particles(self, n):
        self.num_real_particles = int(n)

    # -- growth / shrink (used by the halo exchange for Remote particles) --
    def resize(self, n):
        n = int(n)
        for k, a in list(self.properties.items()):
            stride = a.size // max(self._n, 1) if self._n else 1
            b = np.zeros(n * stride, dtype=a.dtype)
            m = min(a.size, b.size)
            b[:m] = a[:m]
            self.properties[k] = b
        self._n = n
        self.num_real_particles = min(self.num_real_particles, n)

    def extract(self, idx, name=None):
        idx = np.asarray(idx)
        pa = ParticleArray(name=name or self.name, constants=self.constants)
        pa._n = idx.size
        pa.num_real_particles = idx.size
        for k, a in self.properties.items():
            pa.properties[k] = a[idx].copy()
        pa.output_property_arrays = list(self.output_property_arrays)
        return pa


def get_particle_array(additional_props=None, constants=None, **props):
    """pysph/base/utils.py:47-146: default SPH property set."""
    name = props.pop('name', 'array')
    pa = ParticleArray(name=name, constants=constants, **props)
    want = list(DEFAULT_PROPS) + list(additional_props or [])
    for p in want:
        if p not in pa.properties:
            pa.add_property(p)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'm', 'h',
                          'pid', 'gid', 'tag'])
    return pa


def get_particle_array_wcsph(constants=None, **props):
    """pysph/base/utils.py:152-190: the WCSPH property set."""
    pa = get_particle_array(additional_props=WCSPH_PROPS, constants=constants,
                            **props)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'm', 'h',
                          'pid', 'gid', 'tag', 'p'])
    return pa


def get_particle_array_edac(constants=None, **props):
    """A fluid array with the property set ``EDACScheme.setup_properties`` gives it
    for internal flows (wc/edac.py:709-741)."""
    pa = get_particle_array(additional_props=EDAC_TVF_PROPS, constants=constants,
                            **props)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'p', 'm', 'h',
                          'V', 'pavg'])
    return pa


# EDAC_PROPS (wc/edac.py:30-31): the fluid of an external flow (pb == 0, EDACStep)
EDAC_EXT_PROPS = ('ap', 'au', 'av', 'aw', 'ax', 'ay', 'az', 'x0', 'y0', 'z0', 'u0', 'v0', 'w0',
                  'p0', 'V')


def get_particle_array_edac_ext(constants=None, **props):
    """A fluid array of the EDAC scheme WITHOUT transport velocity (wc/edac.py:34-43)."""
    pa = get_particle_array(additional_props=EDAC_EXT_PROPS, constants=constants, **props)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'p', 'm', 'h', 'V'])
    return pa


# TVF_SOLID_PROPS of EDACScheme.setup_properties (wc/edac.py:752-753)
EDAC_WALL_PROPS = ('V', 'wij', 'ax', 'ay', 'az', 'uf', 'vf', 'wf', 'ug', 'vg', 'wg')


def get_particle_array_edac_wall(constants=None, **props):
    """A solid-wall array with the property set ``EDACScheme.setup_properties`` gives the
    ``solids`` of an internal flow (wc/edac.py:752-762): u v w are the PRESCRIBED wall
    velocity, au av aw its prescribed acceleration; V, wij, p, uf.., ug.. are computed by
    the first Group of every evaluation."""
    pa = get_particle_array(additional_props=EDAC_WALL_PROPS, constants=constants, **props)
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'p', 'm', 'h', 'V'])
    return pa


# pysph/sph/solid_mech/basic.py:52-59
ELASTIC_PROPS = ['cs', 'e', 'v00', 'v01', 'v02', 'v10', 'v11', 'v12', 'v20', 'v21',
                 'v22', 'r00', 'r01', 'r02', 'r11', 'r12', 'r22', 's00', 's01', 's02',
                 's11', 's12', 's22', 'as00', 'as01', 'as02', 'as11', 'as12', 'as22',
                 's000', 's010', 's020', 's110', 's120', 's220', 'arho', 'au', 'av',
                 'aw', 'ax', 'ay', 'az', 'ae', 'rho0', 'u0', 'v0', 'w0', 'x0', 'y0',
                 'z0', 'e0']


def get_particle_array_elastic_dynamics(constants=None, **props):
    """pysph/sph/solid_mech/basic.py:32-90: the property set and constants of an
    elastic solid (device kernels: k_solid_pass1/2, k_stage_solid)."""
    consts = {'wdeltap': -1., 'n': 4, 'G': 0.0, 'E': 0.0, 'nu': 0.0,
              'rho_ref': 1000.0, 'c0_ref': 0.0}
    given = dict(constants or {})
    consts.update(given)
    pa = get_particle_array(additional_props=ELASTIC_PROPS, **props)
    for k, v in consts.items():
        pa.add_constant(k, v)
    E, nu, rho_ref = pa.E[0], pa.nu[0], pa.rho_ref[0]
    if 'G' not in given:
        pa.G[0] = E / (2.0 * (1.0 + nu))                   # get_shear_modulus, :26-29
    if E > 0 and 'c0_ref' not in given:
        c0 = np.sqrt(E / (3 * (1.0 - 2 * nu)) / rho_ref)   # get_speed_of_sound, :19-23
        pa.cs[:] = c0
        pa.c0_ref[0] = c0
    pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'rho', 'm', 'h', 'pid', 'gid',
                          'tag', 'p'])
    return pa

"""Mirror boundaries (DomainManager(mirror_in_x ...), nnps_base.pyx:506-689;
`b200sph_set_mirror`, `k_flag_mirror`, `k_mirror_copy`).  Also run on the library emulation
(tests/test_library_on_cpu.py).  First passed on a B200 in the driver's round-1 run
(GPUTEST_r01.json)."""
import numpy as np
import pytest

from helpers import rel_err
from oracle import oracle as orc
from test_gpu_periodic import _dom_tuple, _periodic_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


def _mirror_tuple(dm):
    return _dom_tuple(dm) + ([int(dm.mirror_in_x), int(dm.mirror_in_y), int(dm.mirror_in_z)],)


@pytest.mark.parametrize('dim,n,pattern', [(2, 24, (1, 1, 0)), (3, 10, (1, 0, 1)),
                                           (3, 10, (1, 1, 1)), (2, 24, (0, 1, 0))])
def test_mirror_wcsph_steps_vs_oracle(gpu_device, dim, n, pattern):
    """EPEC steps in a box with mirror planes along `pattern` against the oracle, which
    re-creates the images at every update_domain the reference's way (mirror_ghosts); the
    device selects them once per list build and refreshes their values.  The restatement
    of _create_ghosts_mirror is NOT pinned to an executed reference (the reference has no
    mirror fixture and its NNPS cannot be built here): parity is with the oracle only."""
    import pysph_b200 as pb
    pa, params = _periodic_case(dim, n)
    ref_pa, _ = _periodic_case(dim, n)
    dm = pb.DomainManager(xmin=0, xmax=1, ymin=0, ymax=1, zmin=0,
                          zmax=1 if dim == 3 else 0,
                          mirror_in_x=bool(pattern[0]), mirror_in_y=bool(pattern[1]),
                          mirror_in_z=bool(pattern[2]))
    s = pb.make_wcsph_solver([pa], dict(params), pb.CubicSpline(dim=dim), domain=dm)
    o = orc.WCSPHOracleSolver([ref_pa], dict(params), 'CubicSpline',
                              domain=_mirror_tuple(dm))
    s.initialise()
    o.initialise()
    s.pull()
    r = o.pas[0]
    nr = r.num_real_particles
    assert pa.get_number_of_particles(real=True) == nr
    # the images: tag Ghost, at least the reference's set (the selection is wider by the skin)
    n_img, n_img_ref = pa.get_number_of_particles() - nr, r.get_number_of_particles() - nr
    assert n_img >= n_img_ref > 0
    assert np.all(pa.tag[nr:] == 2) and np.all(pa.tag[:nr] == 0)
    inside = np.ones(n_img, dtype=bool)
    for d, k in enumerate(('x', 'y', 'z')[:dim]):
        inside &= (pa.properties[k][nr:] >= 0.0) & (pa.properties[k][nr:] <= 1.0)
    assert not np.any(inside)                       # every image lies outside the box
    order = np.argsort(pa.gid[:nr])
    for k in ('au', 'av', 'aw', 'arho', 'ax', 'ay', 'az'):
        ref = r.properties[k][:nr]
        if np.max(np.abs(ref)) == 0.0:
            assert np.max(np.abs(pa.properties[k][:nr])) == 0.0
            continue
        assert rel_err(pa.properties[k][:nr][order], ref) <= 2e-5, k
    for _ in range(8):
        s.step()
        o.step()
    s.pull()
    r = o.pas[0]
    order = np.argsort(pa.gid[:nr])
    for k in ('x', 'y', 'z', 'u', 'v', 'w', 'rho'):
        ref = r.properties[k][:nr]
        scale = max(np.max(np.abs(ref)), 1e-12)
        assert np.max(np.abs(pa.properties[k][:nr][order] - ref)) <= 2e-5 * scale, k
    st = s.backend.stats()
    # images were refreshed in place AND re-selected with a new list build
    assert st['light_updates'] > 0 and st['full_builds'] >= 2, st


def test_mirror_errors(gpu_device):
    import pysph_b200 as pb
    pa, params = _periodic_case(2, 8)
    with pytest.raises(Exception):                  # mirror planes in a periodic domain
        dm = pb.DomainManager(xmin=0, xmax=1, ymin=0, ymax=1, periodic_in_y=True,
                              mirror_in_x=True)
        pb.make_wcsph_solver([pa], dict(params), pb.CubicSpline(dim=2), domain=dm).initialise()
    with pytest.raises(Exception):                  # fewer layers than one kernel support
        dm = pb.DomainManager(xmin=0, xmax=1, ymin=0, ymax=1, mirror_in_x=True, n_layers=0.5)
        pb.make_wcsph_solver([pa], dict(params), pb.CubicSpline(dim=2), domain=dm).initialise()

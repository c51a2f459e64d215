"""Synthetic inputs for the benchmark configurations (BASELINE.json configs).

The particle lattices of the reference examples, restated with vectorised
numpy so 1 M / 10 M particle cases build in seconds:

* 3D dam break  -- pysph/examples/_db_geometry.py:250-432 (DamBreak3DGeometry)
  with the parameters of pysph/examples/dam_break_3d.py:17-31,46-51.
* 2D dam break  -- pysph/examples/dam_break_2d.py:212-241, using
  pysph/tools/geometry.py:236-288 (get_2d_tank) and :385-421 (get_2d_block),
  including the example's quirk that particle h and m stay at the module-level
  dx = 0.03 constants whatever --dx is (dam_break_2d.py:35,45-47,230-232).
"""
import numpy as np

from .particle_array import get_particle_array_wcsph


def dam_break_3d_particles(dx=0.02, hdx=1.3, rho0=1000.0, nboundary_layers=1,
                           with_obstacle=True, xrange=None):
    """Return [fluid, boundary, obstacle] for the SPHERIC test-2 tank.

    ``xrange=(lo, hi)`` keeps only lattice columns with lo <= x < hi (used by
    the multi-GPU slab decomposition so that no rank ever materialises the
    whole lattice).
    """
    L, W, H = 3.22, 1.0, 1.0                 # container length, width, height
    fl, fw, fh = 1.228, 1.0, 0.55            # fluid column
    ocx, ocy = 2.5, 0.0                      # obstacle centre
    ol, oh, ow = 0.16, 0.161, 0.4            # obstacle length, height, width

    ghost = nboundary_layers * dx
    eps = 0.1 * dx
    xs = np.mgrid[0.0 - ghost:L + ghost + eps:dx]
    ys = np.mgrid[-0.5 * W - ghost:0.5 * W + ghost + eps:dx]
    zs = np.mgrid[0.0 - ghost:H + ghost + eps:dx]
    ix = np.arange(xs.size)
    if xrange is not None:
        keep = (xs >= xrange[0]) & (xs < xrange[1])
        xs, ix = xs[keep], ix[keep]
    x, y, z = [a.ravel() for a in np.meshgrid(xs, ys, zs, indexing='ij')]
    # global id = lattice index: unique over arrays and independent of how the
    # lattice is split over ranks (used to match results across decompositions)
    gid = ((ix[:, None, None] * ys.size + np.arange(ys.size)[None, :, None]) *
           zs.size + np.arange(zs.size)[None, None, :]).ravel().astype(np.uint32)

    cw2 = 0.5 * W
    fluid_mask = ((0 < x) & (x <= fl) & (-cw2 < y) & (y < cw2) &
                  (0 < z) & (z <= fh))
    obw2, obl2 = 0.5 * ow, 0.5 * ol
    obst_mask = ((ocx - obl2 <= x) & (x <= ocx + obl2) &
                 (ocy - obw2 <= y) & (y <= ocy + obw2) & (0 < z) & (z <= oh))
    wall_mask = (y <= -cw2) | (y >= cw2) | (x >= L) | (x <= 0) | (z <= 0)

    h0 = hdx * dx
    m0 = rho0 * dx ** 3

    def make(name, mask):
        return get_particle_array_wcsph(name=name, x=x[mask], y=y[mask],
                                        z=z[mask], m=m0, h=h0, rho=rho0,
                                        gid=gid[mask])

    arrays = [make('fluid', fluid_mask), make('boundary', wall_mask)]
    if with_obstacle:
        arrays.append(make('obstacle', obst_mask))
    return arrays


def dam_break_3d_params(dx, hdx=1.3):
    """Scheme/solver parameters of pysph/examples/dam_break_3d.py:17-31,54-73."""
    rho0 = 1000.0
    c0 = 10.0 * np.sqrt(2.0 * 9.81 * 0.55)
    h0 = dx * hdx
    return dict(fluids=['fluid'], solids=['boundary', 'obstacle'], dim=3,
                rho0=rho0, c0=float(c0), h0=h0, hdx=hdx, gz=-9.81, alpha=0.25,
                beta=0.0, gamma=7.0, hg_correction=True,
                tensile_correction=False,
                dt0=0.25 * h0 / (1.1 * float(c0)), n_damp=50, cfl=0.3,
                integrator='EPEC')


def _tank_2d(dx, length, height, num_layers, base_center):
    start = (1 - num_layers) * dx
    end = num_layers * dx
    x, y = np.mgrid[start:length + end:dx, start:height + end:dx]
    inside = (x > 0) & (x < length) & (y > 0) & (y < height + 10 * height)
    keep = ~inside
    return x[keep] + base_center[0] - length / 2, y[keep] + base_center[1]


def _block_2d(dx, length, height, center):
    n1 = int(length / dx) + 1
    n2 = int(height / dx) + 1
    x, y = np.mgrid[-length / 2.:length / 2.:n1 * 1j,
                    -height / 2.:height / 2.:n2 * 1j]
    return x.ravel() + center[0], y.ravel() + center[1]


def dam_break_2d_particles(dx=0.03):
    """Return [fluid, boundary] exactly as dam_break_2d.py builds them."""
    rho0 = 1000.0
    h = 1.3 * 0.03                 # module constant, NOT 1.3*dx (reference quirk)
    m = 0.03 ** 2 * rho0           # module constant
    xt, yt = _tank_2d(dx, 4.0, 4.0, 4, [2, 0])
    xf, yf = _block_2d(dx, 1.0, 2.0, [0.5, 1])
    xf = xf + dx
    yf = yf + dx
    fluid = get_particle_array_wcsph(name='fluid', x=xf, y=yf, h=h, m=m,
                                     rho=rho0)
    boundary = get_particle_array_wcsph(name='boundary', x=xt, y=yt, h=h, m=m,
                                        rho=rho0)
    return [fluid, boundary]


def dam_break_2d_params(dx=0.03, hdx=1.3):
    """pysph/examples/dam_break_2d.py:29-47,86-96,146-150."""
    rho0 = 1000.0
    c0 = 10.0 * np.sqrt(2 * 9.81 * 2.0)
    hq = 1.3 * 0.03
    h_opt = hdx * dx
    return dict(fluids=['fluid'], solids=['boundary'], dim=2, rho0=rho0,
                c0=float(c0), h0=hq, hdx=1.3, gy=-9.81, alpha=0.1, beta=0.0,
                gamma=7.0, hg_correction=True, update_h=True,
                dt0=0.125 * h_opt / float(c0), n_damp=50, cfl=0.3,
                integrator='PEC')


# ---------------------------------------------------------------------------
# Taylor-Green vortex (pysph/examples/taylor_green.py:30-35, :146-166, :268-297)
# ---------------------------------------------------------------------------
def taylor_green_params(nx, dim=2, re=100.0, hdx=1.0, L=1.0, U=1.0, rho0=1.0):
    """Parameters of the reference example (EDAC branch, :197-203).  The example is
    2-D; the 3-D variant (BASELINE configs[3]) keeps every parameter and uses the
    classical 3-D Taylor-Green initial field (ours to define)."""
    c0 = 10.0 * U
    p0 = c0 ** 2 * rho0
    nu = U * L / re
    dx = L / nx
    h0 = hdx * dx
    dt = min(0.25 * h0 / (c0 + U), 0.125 * h0 ** 2 / nu, 0.25)
    return dict(dim=dim, c0=c0, rho0=rho0, nu=nu, pb=p0, h=h0, hdx=hdx, dx=dx, dt=dt,
                L=L, U=U, re=re, nx=nx, alpha=0.0, edac_alpha=0.5, bql=True)


def taylor_green_particles(nx, dim=2, re=100.0, hdx=1.0, L=1.0, U=1.0, rho0=1.0,
                           perturb=0.0, seed=1):
    from .particle_array import get_particle_array_edac
    dx = L / nx
    ax = np.arange(dx / 2, L, dx)
    g = np.meshgrid(*([ax] * dim), indexing='ij')
    x = g[0].ravel().copy()
    y = g[1].ravel().copy()
    z = g[2].ravel().copy() if dim == 3 else np.zeros_like(x)
    if perturb > 0:
        rs = np.random.RandomState(seed)
        x += rs.random_sample(x.shape) * dx * perturb
        y += rs.random_sample(x.shape) * dx * perturb
        if dim == 3:
            z += rs.random_sample(x.shape) * dx * perturb
    k = 2 * np.pi / L
    if dim == 2:           # exact_solution(t=0), taylor_green.py:63-70
        u = -U * np.cos(k * x) * np.sin(k * y)
        v = U * np.sin(k * x) * np.cos(k * y)
        w = np.zeros_like(x)
        p = -0.25 * U * U * (np.cos(2 * k * x) + np.cos(2 * k * y))
    else:
        u = U * np.sin(k * x) * np.cos(k * y) * np.cos(k * z)
        v = -U * np.cos(k * x) * np.sin(k * y) * np.cos(k * z)
        w = np.zeros_like(x)
        p = rho0 * U * U / 16.0 * (np.cos(2 * k * x) + np.cos(2 * k * y)) * \
            (np.cos(2 * k * z) + 2.0)
    pa = get_particle_array_edac(name='fluid', x=x, y=y, z=z, u=u, v=v, w=w, p=p,
                                 m=rho0 * dx ** dim, h=hdx * dx, rho=rho0)
    pa.uhat[:] = u
    pa.vhat[:] = v
    pa.what[:] = w
    pa.V[:] = 1.0 / dx ** dim
    pa.gid[:] = np.arange(x.size)
    return pa


# ---------------------------------------------------------------------------
# Colliding elastic rings (pysph/examples/solid_mech/rings.py:18-84)
# ---------------------------------------------------------------------------
def rings_particles(dx=0.0005, hdx=1.5, ri=0.03, ro=0.04, spacing=0.041, E=1e7,
                    nu=0.3975, rho0=1.0, u_f=0.059):
    """One elastic array 'solid' holding both rings, approaching each other at
    u_f * c0 (rings.py:40-78).  The 2-D CubicSpline value W(dx, h) is ``wdeltap``."""
    from .particle_array import get_particle_array_elastic_dynamics
    n = int(round(2 * ro / dx))
    ax = -ro + dx * np.arange(n)                      # numpy.mgrid[-ro:ro:dx]
    x, y = np.meshgrid(ax, ax, indexing='ij')
    x, y = x.ravel(), y.ravel()
    d = x * x + y * y
    keep = np.flatnonzero((ri * ri <= d) * (d < ro * ro))
    x, y = x[keep], y[keep]
    x = np.concatenate([x - spacing, x + spacing])
    y = np.concatenate([y, y])
    h = hdx * dx
    q = dx / h                                         # CubicSpline(dim=2).kernel(rij=dx, h)
    fac = 10.0 / (7.0 * np.pi) / (h * h)
    w = 1.0 - 1.5 * q * q * (1.0 - 0.5 * q) if q <= 1.0 else 0.25 * (2.0 - q) ** 3
    pa = get_particle_array_elastic_dynamics(
        name='solid', x=x + spacing, y=y, m=dx * dx, rho=rho0, h=h,
        constants=dict(wdeltap=fac * w, n=4, rho_ref=rho0, E=E, nu=nu))
    pa.u[:] = pa.cs * u_f * (2 * (x < 0) - 1)
    pa.gid[:] = np.arange(x.size)
    return pa


def rings_3d_particles(dx=0.0005, lz=0.01, hdx=1.5, ri=0.03, ro=0.04, spacing=0.041,
                       E=1e7, nu=0.3975, rho0=1.0, u_f=0.059, x_range=None):
    """BASELINE configs[4]: the colliding rings of rings.py:18-84 extruded along z into
    two thick-walled tubes of length ``lz`` with free end faces (the reference example is
    2-D only; everything but the third lattice axis, the 3-D mass ``dx^3`` and the 3-D
    CubicSpline value ``wdeltap = W(dx, h)`` follows rings.py:40-78).  4 M particles is
    ``dx = 0.00028, lz = 0.02``.  ``x_range = (lo, hi)`` keeps only the lattice columns
    with ``lo <= x < hi`` (a rank's slab) so that no rank builds the whole body."""
    from .particle_array import get_particle_array_elastic_dynamics
    n = int(round(2 * ro / dx))
    ax = -ro + dx * np.arange(n)                      # numpy.mgrid[-ro:ro:dx]
    x, y = np.meshgrid(ax, ax, indexing='ij')
    x, y = x.ravel(), y.ravel()
    d = x * x + y * y
    keep = np.flatnonzero((ri * ri <= d) * (d < ro * ro))
    x, y = x[keep], y[keep]
    sign = np.concatenate([np.ones(x.size), -np.ones(x.size)])    # left ring moves right
    x = np.concatenate([x - spacing, x + spacing]) + spacing
    y = np.concatenate([y, y])
    gid2 = np.arange(x.size)
    if x_range is not None:
        sel = np.flatnonzero((x >= x_range[0]) * (x < x_range[1]))
        x, y, sign, gid2 = x[sel], y[sel], sign[sel], gid2[sel]
    nz = max(int(round(lz / dx)), 1)
    z = dx * (np.arange(nz) + 0.5)
    X = np.repeat(x, nz)
    Y = np.repeat(y, nz)
    Z = np.tile(z, x.size)
    h = hdx * dx
    q = dx / h                                         # CubicSpline(dim=3).kernel(rij=dx, h)
    fac = 1.0 / (np.pi * h * h * h)
    w = 1.0 - 1.5 * q * q * (1.0 - 0.5 * q) if q <= 1.0 else 0.25 * (2.0 - q) ** 3
    pa = get_particle_array_elastic_dynamics(
        name='solid', x=X, y=Y, z=Z, m=dx * dx * dx, rho=rho0, h=h,
        constants=dict(wdeltap=fac * w, n=4, rho_ref=rho0, E=E, nu=nu))
    pa.u[:] = pa.cs * u_f * np.repeat(sign, nz)
    pa.gid[:] = np.repeat(gid2, nz) * nz + np.tile(np.arange(nz), x.size)
    return pa


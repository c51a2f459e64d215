// pool_kernels.cuh -- elementwise / reduction kernels over the particle pool.
// Part of the single translation unit b200sph.cu (included there, in this order; not a
// stand-alone header).

// --------------------------------------------------------------------------
// elementwise / reduction kernels over the pool
// --------------------------------------------------------------------------
struct PoolLayout {
    int narr;
    long long off[B200SPH_MAX_ARRAYS], n[B200SPH_MAX_ARRAYS], n_real[B200SPH_MAX_ARRAYS];
};

__global__ void k_fill_ptype(uint8_t *ptype, long long pool_end, PoolLayout L)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pool_end) return;
    uint8_t t = PT_INVALID;
    for (int a = 0; a < L.narr; a++) {
        long long i = g - L.off[a];
        if (i >= 0 && i < L.n[a]) t = (uint8_t)(a | (i >= L.n_real[a] ? PT_GHOST : 0));
    }
    ptype[g] = t;
}

__global__ void k_red_init(long long *red)
{
    int i = threadIdx.x;
    if (i < 16) red[i] = (i & 1) ? d2o(-1e300) : d2o(1e300);  // even: min slots, odd: max slots
}

// slots: 0 xmin 1 xmax 2 ymin 3 ymax 4 zmin 5 zmax 6 hmin 7 hmax
__global__ void k_reduce_minmax(const double *__restrict__ x, const double *__restrict__ y,
                                const double *__restrict__ z, const double *__restrict__ h,
                                const uint8_t *__restrict__ ptype, long long pool_end,
                                int do_xyz, int do_h, long long *red)
{
    double mn[4] = {1e300, 1e300, 1e300, 1e300}, mx[4] = {-1e300, -1e300, -1e300, -1e300};
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < pool_end;
         g += (long long)gridDim.x * blockDim.x) {
        if (ptype[g] == PT_INVALID) continue;
        if (do_xyz) {
            double v = x[g]; mn[0] = fmin(mn[0], v); mx[0] = fmax(mx[0], v);
            v = y[g]; mn[1] = fmin(mn[1], v); mx[1] = fmax(mx[1], v);
            v = z[g]; mn[2] = fmin(mn[2], v); mx[2] = fmax(mx[2], v);
        }
        if (do_h) {
            double v = h[g]; mn[3] = fmin(mn[3], v); mx[3] = fmax(mx[3], v);
        }
    }
    for (int k = 0; k < 4; k++) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fmin(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmax(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        for (int k = 0; k < 4; k++) {
            if ((k < 3 && !do_xyz) || (k == 3 && !do_h)) continue;
            atomicMin(&red[2 * k], d2o(mn[k]));
            atomicMax(&red[2 * k + 1], d2o(mx[k]));
        }
    }
}

// slots: 9 max dt_cfl, 11 max dt_force (real particles), 12 min h (all)
__global__ void k_reduce_dt(const float *__restrict__ dt_cfl, const float *__restrict__ dt_force,
                            const double *__restrict__ h, const uint8_t *__restrict__ ptype,
                            long long pool_end, long long *red)
{
    double mc = -1e300, mf = -1e300, hm = 1e300;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < pool_end;
         g += (long long)gridDim.x * blockDim.x) {
        uint8_t t = ptype[g];
        if (t == PT_INVALID) continue;
        hm = fmin(hm, h[g]);
        if (t & PT_GHOST) continue;
        mc = fmax(mc, (double)dt_cfl[g]);
        mf = fmax(mf, (double)dt_force[g]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mc = fmax(mc, __shfl_xor_sync(0xffffffffu, mc, o));
        mf = fmax(mf, __shfl_xor_sync(0xffffffffu, mf, o));
        hm = fmin(hm, __shfl_xor_sync(0xffffffffu, hm, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&red[9], d2o(mc));
        atomicMax(&red[11], d2o(mf));
        atomicMin(&red[12], d2o(hm));
    }
}

// Integrator.compute_time_step (integrator.py:161-200) on the reduced factors: the local
// proposal cfl * dt_min, or 1e20 when no factor constrains it (solver.py:655-660)
__global__ void k_dt_propose(const long long *__restrict__ red, double *__restrict__ tc, double cfl, int fixed_h)
{
    const double mc = o2d(red[9]), mf = o2d(red[11]);
    const double f_cfl = mc < -1e299 ? -1.0 : mc, f_force = mf < -1e299 ? -1.0 : mf;
    double hmin = fmin(1.0, o2d(red[12]));
    if (!fixed_h || tc[3] < 0.0) tc[3] = hmin;
    hmin = tc[3];
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    double dt_cfl = inf, dt_force = inf;
    if (f_cfl > 0.0) dt_cfl = hmin / f_cfl;
    if (f_force > 0.0) dt_force = sqrt(hmin / sqrt(f_force));
    const double dt_min = fmin(dt_cfl, dt_force);
    tc[2] = (dt_min <= 0.0 || isinf(dt_min)) ? 1e20 : cfl * dt_min;
}
// Solver loop bookkeeping (solver.py:478-491, :647-688): t += dt; dt = damp(new dt)
// and the last step lands on the final time (solver.py:757-760, :771-773)
__global__ void k_dt_commit(double *__restrict__ tc, double prev_factor, double new_factor, int in_parallel, int adaptive, int advance,
                            double t_final, double t_eps)
{
    const double dt_old = tc[0];
    if (advance) tc[1] += dt_old;
    const double t = tc[1];
    if (fabs(t_final - t) < t_eps) return;  // reached the end: dt stays
    const double undamped = dt_old / prev_factor;
    double dt = undamped;
    if (adaptive) {
        dt = tc[2];
        if (!in_parallel && dt >= 1e20) dt = undamped;
    }
    dt *= new_factor;
    if (t + dt > t_final - t_eps) dt = t_final - t;
    tc[0] = dt;
}
// TaitEOS.loop wc/basic.py:60-65 ; TaitEOSHGCorrection.loop wc/basic.py:118-126
__global__ void k_eos(double *__restrict__ rho, float *__restrict__ p, float *__restrict__ cs,
                      const uint8_t *__restrict__ ptype, long long lo, long long hi, int hg,
                      double rho0, double c0, double gamma, double p0)
{
    long long g = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= hi) return;
    double r = rho[g];
    if (hg && r < rho0) {
        r = rho0;
        rho[g] = r;
    }
    double ratio = r * (1.0 / rho0);
    double B = rho0 * c0 * c0 / gamma;
    double tmp = pow(ratio, gamma);
    p[g] = (float)((hg ? 0.0 : p0) + B * (tmp - 1.0));
    cs[g] = (float)(c0 * pow(ratio, 0.5 * (gamma - 1.0)));
}

// UpdateSmoothingLengthFerrari.loop wc/basic.py:458-463
__global__ void k_ferrari(double *__restrict__ h, const double *__restrict__ m,
                          const double *__restrict__ rho, long long lo, long long hi, double hdx,
                          double dim1)
{
    long long g = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= hi) return;
    h[g] = hdx * pow(m[g] / rho[g], dim1);
}

struct StageArgs {
    double *x, *y, *z, *u, *v, *w, *rho;
    double *x0, *y0, *z0, *u0, *v0, *w0, *rho0;
    const float *au, *av, *aw, *ax, *ay, *az, *arho;
    const uint8_t *ptype;
    long long pool_end;
    int arr, which;
    double f;
};

// WCSPHStep.initialize / stage1 / stage2 integrator_step.py:51-91 (real particles only,
// integrator_cython.mako:97-111)
__device__ __forceinline__ void stage_body(const StageArgs &a);
__global__ void k_stage(StageArgs a) { stage_body(a); }
__global__ void k_stage_devdt(StageArgs a, const double *__restrict__ tc)
{
    const double dt = tc[0];
    a.f = a.which == 1 ? 0.5 * dt : dt;
    stage_body(a);
}
__device__ __forceinline__ void stage_body(const StageArgs &a)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.pool_end) return;
    uint8_t t = a.ptype[g];
    if (t == PT_INVALID || (t & PT_GHOST)) return;
    if (a.arr >= 0 && (t & 7) != a.arr) return;
    if (a.which == 0) {
        a.x0[g] = a.x[g]; a.y0[g] = a.y[g]; a.z0[g] = a.z[g];
        a.u0[g] = a.u[g]; a.v0[g] = a.v[g]; a.w0[g] = a.w[g];
        a.rho0[g] = a.rho[g];
    } else {
        const double f = a.f;
        a.u[g] = a.u0[g] + f * (double)a.au[g];
        a.v[g] = a.v0[g] + f * (double)a.av[g];
        a.w[g] = a.w0[g] + f * (double)a.aw[g];
        a.x[g] = a.x0[g] + f * (double)a.ax[g];
        a.y[g] = a.y0[g] + f * (double)a.ay[g];
        a.z[g] = a.z0[g] + f * (double)a.az[g];
        a.rho[g] = a.rho0[g] + f * (double)a.arho[g];
    }
}

// _box_wrap_periodic (nnps_base.pyx:699-743): real and ghost particles alike
__global__ void k_box_wrap(double *__restrict__ x, double *__restrict__ y, double *__restrict__ z,
                           const uint8_t *__restrict__ ptype, long long pool_end, GridDev D /* xmin = lo, cell = L */)
{
    long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= pool_end || ptype[g] == PT_INVALID) return;
    double *p[3] = {x, y, z};
#pragma unroll
    for (int d = 0; d < 3; d++) {
        if (!D.periodic[d]) continue;
        double v = p[d][g];
        const double lo = D.xmin[d], L = D.cell[d];
        if (v < lo) v += L;
        if (v > lo + L) v -= L;
        p[d][g] = v;
    }
}
__global__ void k_f64_to_f32(const double *__restrict__ in, float *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void k_f32_to_f64(const float *__restrict__ in, double *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}

// halo_kernels.cuh -- halo / migration kernels of the slab decomposition.
// Part of the single translation unit b200sph.cu (included there, in this order; not a
// stand-alone header).

// --------------------------------------------------------------------------
// halo / migration kernels
// --------------------------------------------------------------------------
// mode 0: flag = lo <= x < hi ; mode 1: flag = x < lo ; mode 2: flag = x >= hi ;
// mode 3: flag = keep (lo <= x < hi)
__global__ void k_flag_range(const double *__restrict__ x, long long off, long long n, double lo,
                             double hi, int mode, uint32_t *__restrict__ flag)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) {
        const double v = x[off + i];
        if (mode == 0 || mode == 3) f = (v >= lo && v < hi);
        else if (mode == 1) f = v < lo;
        else f = v >= hi;
    }
    flag[i] = f;  // flag[n] = 0 so that scan[n] = total
}

__global__ void k_gather_f64(const double *__restrict__ src, long long off, long long n,
                             const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                             double *__restrict__ dst, long long dst_off)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) dst[dst_off + pos[i]] = src[off + i];
}
__global__ void k_gather_u32_as_f64(const uint32_t *__restrict__ src, long long off, long long n,
                                    const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                    double *__restrict__ dst, long long dst_off)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) dst[dst_off + pos[i]] = (double)src[off + i];
}
__global__ void k_gather_u32(const uint32_t *__restrict__ src, long long off, long long n,
                             const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                             uint32_t *__restrict__ dst)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) dst[pos[i]] = src[off + i];
}
// the ghost message: B200SPH_HALO_FIELDS (x y z u v w rho h m), or with the elastic-dynamics
// arrays allocated B200SPH_HALO_FIELDS_SOLID (+ s00 s01 s02 s11 s12 s22 and, last, the fp32
// property cs widened to a double: no equation of that scheme recomputes it); NF is a
// template parameter so that the 9-field kernels stay the unrolled ones that were measured
#define HALO_ND(NF) ((NF) == B200SPH_HALO_FIELDS ? B200SPH_HALO_FIELDS : B200SPH_HALO_FIELDS_SOLID - 1)
struct HaloPtrs {
    double *p[B200SPH_HALO_FIELDS_SOLID - 1];
    float *cs;
};
// all NF fields of the selected particles in one launch (field-major, tight)
template <int NF>
__global__ void k_halo_gather_flag(HaloPtrs P, long long off, long long n, const uint32_t *__restrict__ flag,
                                   const uint32_t *__restrict__ pos, double *__restrict__ dst, long long tot)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    const long long k = pos[i];
#pragma unroll
    for (int f = 0; f < HALO_ND(NF); f++) dst[(long long)f * tot + k] = P.p[f][off + i];
    if (NF != B200SPH_HALO_FIELDS) dst[(long long)(NF - 1) * tot + k] = (double)P.cs[off + i];
}
template <int NF>
__global__ void k_halo_gather_idx(HaloPtrs P, long long off, const uint32_t *__restrict__ idx, long long n,
                                  double *__restrict__ dst)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long i = idx[k];
#pragma unroll
    for (int f = 0; f < HALO_ND(NF); f++) dst[(long long)f * n + k] = P.p[f][off + i];
    if (NF != B200SPH_HALO_FIELDS) dst[(long long)(NF - 1) * n + k] = (double)P.cs[off + i];
}
template <int NF>
__global__ void k_halo_scatter(HaloPtrs P, long long o, const double *__restrict__ src, long long stride, long long n)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
#pragma unroll
    for (int f = 0; f < HALO_ND(NF); f++) P.p[f][o + k] = src[(long long)f * stride + k];
    if (NF != B200SPH_HALO_FIELDS) P.cs[o + k] = (float)src[(long long)(NF - 1) * stride + k];
}
// mirror images (nnps_base.pyx:506-689): side 0: (v - plane) <= width, side 1: (plane - v) <= width
__global__ void k_flag_mirror(const double *__restrict__ v, long long off, long long n, double plane, double width, int side,
                              uint32_t *__restrict__ flag)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) {
        const double x = v[off + i];
        f = side == 0 ? ((x - plane) <= width) : ((plane - x) <= width);
    }
    flag[i] = f;  // flag[n] = 0 so that scan[n] = total
}
// image k = source idx[k] reflected in the plane normal to `axis`: position 2 plane - x, the
// normal velocity component negated, the other fields of the ghost message copied
__global__ void k_mirror_copy(HaloPtrs P, long long off, const uint32_t *__restrict__ idx, long long n, long long dst,
                              int axis, double plane2)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long i = off + idx[k];
#pragma unroll
    for (int f = 0; f < B200SPH_HALO_FIELDS; f++) {
        double v = P.p[f][i];
        if (f == axis) v = plane2 - v;
        else if (f == 3 + axis) v = -v;
        P.p[f][dst + k] = v;
    }
}
struct HaloAllArgs {
    int narr;
    long long prefix[B200SPH_MAX_ARRAYS + 1];  // particles before array a in the message
    long long off[B200SPH_MAX_ARRAYS];         // pool offset of the first particle addressed
    const uint32_t *idx[B200SPH_MAX_ARRAYS];   // gather: saved selection (relative to off); scatter: unused
};
// the refresh message of ALL arrays in one launch: block of array a starts at
// NF * prefix[a] doubles, field-major and tight inside the block.  `dst` may be a peer
// pointer (the neighbour's staging buffer): then this kernel is pack + send in one.
template <int NF>
__global__ void k_halo_gather_all(HaloPtrs P, HaloAllArgs A, double *__restrict__ dst)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.prefix[A.narr]) return;
    int a = 0;
    while (k >= A.prefix[a + 1]) a++;
    const long long r = k - A.prefix[a], cnt = A.prefix[a + 1] - A.prefix[a];
    const long long i = A.off[a] + A.idx[a][r];
    double *out = dst + NF * A.prefix[a] + r;
#pragma unroll
    for (int f = 0; f < HALO_ND(NF); f++) out[(long long)f * cnt] = P.p[f][i];
    if (NF != B200SPH_HALO_FIELDS) out[(long long)(NF - 1) * cnt] = (double)P.cs[i];
}
struct RepackArgs {
    const uint32_t *rank, *skey;   // rank == nullptr: no repack
    float4 *A, *AB;
    GridDev G;
};
// ghost values refreshed in place; with a valid neighbour build the ghosts' packed
// cell-relative positions are refreshed in the same pass (x, y, z, h are fields 0, 1, 2, 7)
template <int NF>
__global__ void k_halo_scatter_all(HaloPtrs P, HaloAllArgs A, const double *__restrict__ src, RepackArgs R)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.prefix[A.narr]) return;
    int a = 0;
    while (k >= A.prefix[a + 1]) a++;
    const long long r = k - A.prefix[a], cnt = A.prefix[a + 1] - A.prefix[a];
    const double *in = src + NF * A.prefix[a] + r;
    double v[HALO_ND(NF)];
#pragma unroll
    for (int f = 0; f < HALO_ND(NF); f++) {
        v[f] = in[(long long)f * cnt];
        P.p[f][A.off[a] + r] = v[f];
    }
    if (NF != B200SPH_HALO_FIELDS) P.cs[A.off[a] + r] = (float)in[(long long)(NF - 1) * cnt];
    if (R.rank) {
        const uint32_t s = R.rank[A.off[a] + r];
        uint32_t key = R.skey[s];
        const uint32_t cx = key % (uint32_t)R.G.nc[0];
        key /= (uint32_t)R.G.nc[0];
        const uint32_t cy = key % (uint32_t)R.G.nc[1];
        const uint32_t cz = key / (uint32_t)R.G.nc[1];
        float4 q;
        q.x = (float)(v[0] - (R.G.xmin[0] + (double)cx * R.G.cell[0]));
        q.y = (float)(v[1] - (R.G.xmin[1] + (double)cy * R.G.cell[1]));
        q.z = (float)(v[2] - (R.G.xmin[2] + (double)cz * R.G.cell[2]));
        q.w = (float)v[7];
        R.A[s] = q;
        R.AB[2 * s] = q;
    }
}
__global__ void k_drift_ratio(const unsigned *__restrict__ red_u32, float kr, float S, double *__restrict__ out)
{
    const float need = 2.0f * sqrtf(__uint_as_float(red_u32[0])) + kr * __uint_as_float(red_u32[1]);
    out[0] = S > 0.f ? (double)(need / S) : 2.0;
}
__global__ void k_save_idx(long long n, const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                           uint32_t *__restrict__ idx)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) idx[pos[i]] = (uint32_t)i;
}
__global__ void k_f64_to_u32(const double *__restrict__ in, uint32_t *__restrict__ out, long long n)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)in[i];
}
__global__ void k_gather_f32_as_f64(const float *__restrict__ src, long long off, long long n,
                                    const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos,
                                    double *__restrict__ dst, long long dst_off)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) dst[dst_off + pos[i]] = (double)src[off + i];
}
__global__ void k_fill_u32(uint32_t *p, long long n, uint32_t v)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

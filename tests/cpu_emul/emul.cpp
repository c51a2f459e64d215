// Host-side emulation harness for the list-consumer kernels of csrc/b200sph.cu.
// TEST INFRASTRUCTURE: the KERNEL SOURCE ITSELF (extracted verbatim from b200sph.cu into
// kernels_extract.inc by tests/test_kernel_source_on_cpu.py) is compiled with g++ against
// the tiny CUDA shim below and run thread by thread, so that the arithmetic, the record
// layouts, the type masks and the output indexing of a kernel can be checked against the
// golden fixtures without a GPU.  It says nothing about launch configuration, memory
// spaces or performance.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "b200sph.h"

struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct idx3 { unsigned x, y, z; };
static thread_local idx3 threadIdx, blockIdx, blockDim, gridDim;

// Warp lockstep for warp-synchronous kernels (k_list_build): the 32 lanes of a warp run as
// 32 OS threads; every __shfl_sync / __ballot_sync is "publish, barrier, read, barrier".
// Kernels launched thread by thread (g_warp == nullptr) see the identity shuffle.
#include <atomic>
#include <thread>
struct WarpCtx {
    std::atomic<int> arrived{0};
    std::atomic<int> phase{0};
    unsigned long long slot[32];
    void barrier()
    {
        const int ph = phase.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) == 31) {
            arrived.store(0, std::memory_order_relaxed);
            phase.store(ph + 1, std::memory_order_release);
        } else {
            while (phase.load(std::memory_order_acquire) == ph) std::this_thread::yield();
        }
    }
};
static thread_local WarpCtx *g_warp = nullptr;
static thread_local int g_lane = 0;
template <class T> static inline unsigned long long to_bits(T v) { unsigned long long b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T from_bits(unsigned long long b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T warp_read(T v, int src)
{
    if (!g_warp) return v;
    g_warp->slot[g_lane] = to_bits(v);
    g_warp->barrier();
    const T r = from_bits<T>(g_warp->slot[src & 31]);
    g_warp->barrier();
    return r;
}

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
static inline void __syncthreads() {}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int o) { return warp_read(v, g_lane ^ o); }
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return warp_read(v, src); }
static inline unsigned __ballot_sync(unsigned, bool pred)
{
    if (!g_warp) return pred ? 1u : 0u;
    g_warp->slot[g_lane] = pred ? 1ull : 0ull;
    g_warp->barrier();
    unsigned m = 0;
    for (int l = 0; l < 32; l++) m |= (unsigned)g_warp->slot[l] << l;
    g_warp->barrier();
    return m;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __syncwarp(unsigned = 0xffffffffu) { if (g_warp) g_warp->barrier(); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
static std::atomic_flag g_atomic_lock = ATOMIC_FLAG_INIT;
static inline unsigned atomicMax(unsigned *p, unsigned v)
{
    while (g_atomic_lock.test_and_set(std::memory_order_acquire)) {}
    const unsigned o = *p;
    if (v > o) *p = v;
    g_atomic_lock.clear(std::memory_order_release);
    return o;
}
static inline uint32_t __ldcs(const uint32_t *p) { return *p; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
// glibc declares __expf(float) but does not export it: route the CUDA intrinsic to expf
#define __expf(x) expf(x)
static inline float frcp(float x) { return 1.0f / x; }
static inline float frsqrt(float x) { return 1.0f / sqrtf(x); }
static inline void ld_256(const float4 *p, float4 &b, float4 &c) { b = p[0]; c = p[1]; }

// smem_tab_t / smem_tab / lds_T: inline PTX in the CUDA build
#define B200SPH_HOST_EMULATION 1
typedef const float4 *smem_tab_t;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline smem_tab_t smem_tab(const float4 *table) { return table; }
static inline float4 lds_T(smem_tab_t t, uint32_t index) { return t[index]; }
static inline void atomicAdd(unsigned long long *p, unsigned long long v) { __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
using std::max;
using std::min;

#define PT_INVALID 0xFFu
#define PT_GHOST 0x08u
#define LIST_JBITS 26
#define LIST_CBITS 6
#define LIST_ENTRY(j, code) (((uint32_t)(j) << LIST_CBITS) | (uint32_t)(code))
#define LIST_J(e) ((uint32_t)(e) >> LIST_CBITS)
#define LIST_CODE(e) ((uint32_t)(e) & 63u)
#define LIST_NT 128

#include "kernels_extract.inc"

// run one "launch" thread by thread.  Kernels that fill a block-shared table in their first
// 64 threads (the cell-offset table: identical for every block, and `static` here) get one
// warm-up execution of block 0 first, so that the table is complete when threads run for
// real; `after_warm` undoes whatever the warm-up must not leave behind (pair counters).
template <class F, class R> static void launch(long long n, int nt, F body, R after_warm)
{
    const unsigned nb = (unsigned)((n + nt - 1) / nt);
    blockDim = idx3{(unsigned)nt, 1, 1};
    gridDim = idx3{nb, 1, 1};
    blockIdx = idx3{0, 0, 0};
    for (int t = 0; t < nt && nb > 0; t++) {
        threadIdx = idx3{(unsigned)t, 0, 0};
        body();
    }
    after_warm();
    for (unsigned b = 0; b < nb; b++) {
        blockIdx = idx3{b, 0, 0};
        for (int t = 0; t < nt; t++) {
            threadIdx = idx3{(unsigned)t, 0, 0};
            body();
        }
    }
}
template <class F> static void launch(long long n, int nt, F body) { launch(n, nt, body, [] {}); }
// one execution per thread (kernels without block-shared state, e.g. with atomics)
template <class F> static void launch1(long long n, int nt, F body)
{
    const unsigned nb = (unsigned)((n + nt - 1) / nt);
    blockDim = idx3{(unsigned)nt, 1, 1};
    gridDim = idx3{nb, 1, 1};
    for (unsigned b = 0; b < nb; b++) {
        blockIdx = idx3{b, 0, 0};
        for (int t = 0; t < nt; t++) {
            threadIdx = idx3{(unsigned)t, 0, 0};
            body();
        }
    }
}
// warp-synchronous kernels: every warp of every block as 32 lock-stepped OS threads
template <class F> static void launch_warps(unsigned nb, int nt, F body)
{
    for (unsigned b = 0; b < nb; b++)
        for (int w = 0; w < nt / 32; w++) {
            WarpCtx ctx;
            std::vector<std::thread> lanes;
            for (int l = 0; l < 32; l++)
                lanes.emplace_back([&, l] {
                    g_warp = &ctx;
                    g_lane = l;
                    blockDim = idx3{(unsigned)nt, 1, 1};
                    gridDim = idx3{nb, 1, 1};
                    blockIdx = idx3{b, 0, 0};
                    threadIdx = idx3{(unsigned)(w * 32 + l), 0, 0};
                    body();
                    g_warp = nullptr;
                });
            for (auto &t : lanes) t.join();
        }
}

template <int K, int D> static void run_tvf(const TvfArgs &a, const uint32_t *cnt, const uint32_t *lst, int capg, int passes)
{
    if (passes & 1) launch(a.n, LIST_NT, [&] { k_tvf_pass1<K, D>(a, cnt, lst, capg); });
    if (passes & 2) launch(a.n, LIST_NT, [&] { k_tvf_pass2<K, D>(a, cnt, lst, capg); });
}
template <int K, int D> static void run_solid(const SolidArgs &a, const uint32_t *cnt, const uint32_t *lst, int capg, int passes)
{
    if (passes & 1) launch(a.n, LIST_NT, [&] { k_solid_pass1<K, D>(a, cnt, lst, capg); });
    if (passes & 2) launch(a.n, LIST_NT, [&] { k_solid_pass2<K, D>(a, cnt, lst, capg); });
}

#define DISPATCH(fn, ...)                                                        \
    switch (kernel * 4 + dim) {                                                  \
    case 0 * 4 + 2: fn<0, 2>(__VA_ARGS__); break;                                \
    case 0 * 4 + 3: fn<0, 3>(__VA_ARGS__); break;                                \
    case 1 * 4 + 2: fn<1, 2>(__VA_ARGS__); break;                                \
    case 1 * 4 + 3: fn<1, 3>(__VA_ARGS__); break;                                \
    case 2 * 4 + 2: fn<2, 2>(__VA_ARGS__); break;                                \
    case 2 * 4 + 3: fn<2, 3>(__VA_ARGS__); break;                                \
    default: return -1;                                                          \
    }

extern "C" {

// Everything lives in one "cell": positions are used as they are, every list entry carries the
// zero cell-offset code (dx = dy = dz = 0  ->  1 + 4 + 16 = 21) and every particle is a
// candidate of every particle (the kernels re-apply the exact accept test).
struct emul_common {
    long long n;            // particles (pool order == sorted order)
    int kernel, dim;
    double radius_scale, kfac;
    const double *x, *y, *z, *h, *u, *v, *w, *m;
    double *rho;
    const uint8_t *ptype;
};

static void make_lists(long long n, std::vector<uint32_t> &cnt, std::vector<uint32_t> &lst, int &capg)
{
    capg = (int)n;
    cnt.assign((size_t)n, (uint32_t)n);
    lst.assign((size_t)((n + 31) / 32) * (size_t)capg * 32u, 0u);
    for (long long s = 0; s < n; s++)
        for (long long k = 0; k < n; k++) lst[((size_t)(s >> 5) * capg + (size_t)k) * 32u + (size_t)(s & 31)] = LIST_ENTRY(k, 21u);
}

static void pack_A(const emul_common &c, std::vector<float4> &AB)
{
    AB.assign(2 * (size_t)c.n, float4{0, 0, 0, 0});
    for (long long s = 0; s < c.n; s++) AB[2 * s] = make_float4((float)c.x[s], (float)c.y[s], (float)c.z[s], (float)c.h[s]);
}

// one WCSPH Group: [pending EOS calls applied by k_pack_state] + k_pair_list.
// eos: per array {on, hg, real_only} and {rho0, c0, gamma, p0}; emask[d] = 8 bits per source
int emul_wcsph(const emul_common *c, const int *eos_i, const double *eos_d, const unsigned long long *emask,
               const double *params /* c0 alpha beta gx gy gz eps_xsph */, int tensile, int real_only, double deltap,
               float *p, float *cs, float *arho, float *au, float *av, float *aw, float *ax, float *ay, float *az,
               float *dt_cfl, float *dt_force, unsigned long long *pairs)
{
    const long long n = c->n;
    std::vector<uint32_t> cnt, lst, perm((size_t)n);
    int capg;
    make_lists(n, cnt, lst, capg);
    for (long long i = 0; i < n; i++) perm[i] = (uint32_t)i;
    std::vector<float4> AB, B((size_t)n), Cc((size_t)n);
    pack_A(*c, AB);
    EosTab E;
    memset(&E, 0, sizeof(E));
    int any = 0;
    for (int a = 0; a < B200SPH_MAX_ARRAYS; a++) {
        E.on[a] = eos_i[3 * a]; E.hg[a] = eos_i[3 * a + 1]; E.real_only[a] = eos_i[3 * a + 2];
        E.rho0[a] = eos_d[4 * a]; E.c0[a] = eos_d[4 * a + 1]; E.gamma[a] = eos_d[4 * a + 2]; E.p0[a] = eos_d[4 * a + 3];
        any |= E.on[a];
    }
    launch(n, 256, [&] { k_pack_state(c->u, c->v, c->w, c->m, c->rho, p, cs, c->ptype, perm.data(), n, B.data(), Cc.data(), AB.data(), E, any); });
    PairArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.A = nullptr; pa.B = B.data(); pa.C = Cc.data(); pa.AB = AB.data();
    pa.perm = perm.data();
    pa.arho = arho; pa.au = au; pa.av = av; pa.aw = aw; pa.ax = ax; pa.ay = ay; pa.az = az;
    pa.dt_cfl = dt_cfl; pa.dt_force = dt_force;
    pa.rho = c->rho;
    pa.n = n;
    pa.cellx = pa.celly = pa.cellz = 1.0f;
    pa.k2 = (float)(c->radius_scale * c->radius_scale);
    pa.kfac = (float)c->kfac;
    pa.deltap = (float)deltap;
    for (int a = 0; a < B200SPH_MAX_ARRAYS; a++) pa.emask[a] = emask[a];
    pa.c0 = (float)params[0]; pa.alpha = (float)params[1]; pa.beta = (float)params[2];
    pa.gx = (float)params[3]; pa.gy = (float)params[4]; pa.gz = (float)params[5]; pa.eps_xsph = (float)params[6];
    pa.tensile = tensile;
    pa.real_only = real_only;
    unsigned long long counter = 0;
    pa.pair_counter = nullptr;
    const int kernel = c->kernel, dim = c->dim;
    switch (kernel * 4 + dim) {
#define PL(K, D) case K * 4 + D: launch(n, LIST_NT, [&] { k_pair_list<K, D, 6, PAIR_EQS_ALL>(pa, cnt.data(), lst.data(), capg, nullptr); }); break;
        PL(0, 2) PL(0, 3) PL(1, 2) PL(1, 3) PL(2, 2) PL(2, 3) PL(3, 2) PL(3, 3)
#undef PL
    default: return -1;
    }
    if (pairs) *pairs = counter;
    return 0;
}

// The whole neighbour + pair pipeline of one WCSPH Group on a real cell grid: k_cell_count,
// (host prefix sum instead of the 3-phase scan), k_scatter, k_canon, k_pack_pos, k_list_build
// (count pass + fill pass, like build_lists), k_pack_state, k_pair_list.  Outputs in POOL order.
int emul_pipeline(const emul_common *c, const double *gxmin, const double *gcell, const int *gnc, const int *gper,
                  double skin_abs, const int *eos_i, const double *eos_d, const unsigned long long *emask,
                  const double *params, int tensile, int real_only, double deltap, float *p, float *cs, float *arho,
                  float *au, float *av, float *aw, float *ax, float *ay, float *az, float *dt_cfl, float *dt_force,
                  unsigned long long *pairs, int *capg_out)
{
    const long long n = c->n;
    GridDev G;
    for (int d = 0; d < 3; d++) { G.xmin[d] = gxmin[d]; G.cell[d] = gcell[d]; G.nc[d] = gnc[d]; G.periodic[d] = gper[d]; }
    // B200SPH_ZORDER=1: the cell rows along the Z-curve of (cy, cz), row table padded as in b200sph.cu
    G.zorder = getenv("B200SPH_ZORDER") ? atoi(getenv("B200SPH_ZORDER")) != 0 : 0;
    long long ncells = (long long)G.nc[0] * G.nc[1] * G.nc[2];
    if (G.zorder) {
        int bits = 0;
        while ((1 << bits) < std::max(G.nc[1], G.nc[2])) bits++;
        ncells = ((long long)1 << (2 * bits)) * G.nc[0];
    }
    std::vector<uint32_t> key_of((size_t)n), off_in((size_t)n), cell_cnt((size_t)ncells + 1, 0u), cell_start((size_t)ncells + 2, 0u);
    std::vector<uint32_t> perm_tmp((size_t)n), perm((size_t)n), skey((size_t)n), rank((size_t)n);
    launch1(n, 256, [&] { k_cell_count(c->x, c->y, c->z, c->ptype, n, G, key_of.data(), off_in.data(), cell_cnt.data()); });
    for (long long k = 0; k < ncells + 1; k++) cell_start[k + 1] = cell_start[k] + cell_cnt[k];   // exclusive scan
    launch1(n, 256, [&] { k_scatter(key_of.data(), off_in.data(), c->ptype, n, cell_start.data(), perm_tmp.data()); });
    launch1(n, 256, [&] { k_canon(perm_tmp.data(), key_of.data(), cell_start.data(), n, perm.data(), skey.data(), rank.data()); });
    std::vector<float4> A((size_t)n), AB(2 * (size_t)n), B((size_t)n), Cc((size_t)n);
    std::vector<uint8_t> stype((size_t)n);
    launch1(n, 256, [&] { k_pack_pos(c->x, c->y, c->z, c->h, perm.data(), skey.data(), n, G, A.data(), AB.data(), c->ptype, stype.data()); });
    // neighbour lists
    std::vector<uint32_t> cnt((size_t)n, 0u), lst;
    unsigned max_count = 0;
    ListBuildArgs la;
    la.A = A.data(); la.cell_start = cell_start.data(); la.skey = skey.data();
    la.n = n;
    la.ncx = G.nc[0]; la.ncy = G.nc[1]; la.ncz = G.nc[2];
    la.zorder = G.zorder;
    la.px = G.periodic[0]; la.py = G.periodic[1]; la.pz = G.periodic[2];
    la.cellx = (float)G.cell[0]; la.celly = (float)G.cell[1]; la.cellz = (float)G.cell[2];
    la.kr = (float)c->radius_scale;
    la.S = (float)skin_abs;
    la.cnt = cnt.data();
    la.max_count = &max_count;
    for (int d = 0; d < B200SPH_MAX_ARRAYS; d++) la.emask[d] = emask[d];   // the Group's (dest, source) pairs
    la.stype = stype.data();
    const bool per = la.px || la.py || la.pz;
    const unsigned nbw = (unsigned)((n + LB_WARPS * 32 - 1) / (LB_WARPS * 32));
    int capg = 0;
    for (int pass = 0; pass < 2; pass++) {
        la.lst = pass == 0 ? nullptr : lst.data();
        la.capg = capg;
        max_count = 0;
        if (per) launch_warps(nbw, LB_WARPS * 32, [&] { k_list_build<true>(la); });
        else launch_warps(nbw, LB_WARPS * 32, [&] { k_list_build<false>(la); });
        if (pass == 0) {
            capg = ((int)(max_count * 1.15) + 8 + 7) / 8 * 8;
            lst.assign((size_t)((n + 31) / 32) * (size_t)capg * 32u, 0u);
        }
    }
    if ((int)max_count > capg) return -2;
    *capg_out = capg;
    // state records + the pair kernel, outputs scattered back through perm by the kernel itself
    EosTab E;
    memset(&E, 0, sizeof(E));
    int any = 0;
    for (int a = 0; a < B200SPH_MAX_ARRAYS; a++) {
        E.on[a] = eos_i[3 * a]; E.hg[a] = eos_i[3 * a + 1]; E.real_only[a] = eos_i[3 * a + 2];
        E.rho0[a] = eos_d[4 * a]; E.c0[a] = eos_d[4 * a + 1]; E.gamma[a] = eos_d[4 * a + 2]; E.p0[a] = eos_d[4 * a + 3];
        any |= E.on[a];
    }
    launch1(n, 256, [&] { k_pack_state(c->u, c->v, c->w, c->m, c->rho, p, cs, c->ptype, perm.data(), n, B.data(), Cc.data(), AB.data(), E, any); });
    PairArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.B = B.data(); pa.C = Cc.data(); pa.AB = AB.data(); pa.perm = perm.data();
    pa.arho = arho; pa.au = au; pa.av = av; pa.aw = aw; pa.ax = ax; pa.ay = ay; pa.az = az;
    pa.dt_cfl = dt_cfl; pa.dt_force = dt_force;
    pa.rho = c->rho;
    pa.n = n;
    pa.cellx = la.cellx; pa.celly = la.celly; pa.cellz = la.cellz;
    pa.k2 = (float)(c->radius_scale * c->radius_scale);
    pa.kfac = (float)c->kfac;
    pa.deltap = (float)deltap;
    for (int a = 0; a < B200SPH_MAX_ARRAYS; a++) pa.emask[a] = emask[a];
    pa.c0 = (float)params[0]; pa.alpha = (float)params[1]; pa.beta = (float)params[2];
    pa.gx = (float)params[3]; pa.gy = (float)params[4]; pa.gz = (float)params[5]; pa.eps_xsph = (float)params[6];
    pa.tensile = tensile;
    pa.real_only = real_only;
    unsigned long long counter = 0;
    pa.pair_counter = pairs ? &counter : nullptr;
    const int kernel = c->kernel, dim = c->dim;
    // thread by thread for the results; when pairs are counted, once more with real warps:
    // the kernel sums its per-thread counts with warp shuffles, which the thread-by-thread
    // mode cannot reproduce (identity shuffle)
    const unsigned nbl = (unsigned)((n + LIST_NT - 1) / LIST_NT);
#define PL(K, D)                                                                                          \
    case K * 4 + D:                                                                                       \
        pa.pair_counter = nullptr;                                                                        \
        launch(n, LIST_NT, [&] { k_pair_list<K, D, 6, PAIR_EQS_ALL>(pa, cnt.data(), lst.data(), capg, nullptr); });                 \
        if (pairs) {                                                                                      \
            pa.pair_counter = &counter;                                                                   \
            launch_warps(nbl, LIST_NT, [&] { k_pair_list<K, D, 6, PAIR_EQS_ALL>(pa, cnt.data(), lst.data(), capg, nullptr); });     \
        }                                                                                                 \
        break;
    switch (kernel * 4 + dim) {
        PL(0, 2) PL(0, 3) PL(1, 2) PL(1, 3) PL(2, 2) PL(2, 3) PL(3, 2) PL(3, 3)
    default: return -1;
    }
#undef PL
    if (pairs) *pairs = counter;
    return 0;
}

int emul_tvf(const emul_common *c, const b200sph_tvf_program *prog, const double *uh, const double *vh, const double *wh,
             const double *pf, float *V, float *pavg, float *au, float *av, float *aw, float *auhat, float *avhat,
             float *awhat, float *ap, unsigned long long *pairs)
{
    const long long n = c->n;
    std::vector<uint32_t> cnt, lst, perm((size_t)n);
    int capg;
    make_lists(n, cnt, lst, capg);
    for (long long i = 0; i < n; i++) perm[i] = (uint32_t)i;
    std::vector<float4> AB, B((size_t)n), C2((size_t)n), Dv((size_t)n);
    std::vector<float2> PT((size_t)n);
    pack_A(*c, AB);
    launch(n, 256, [&] { k_pack_tvf(c->u, c->v, c->w, c->m, uh, vh, wh, pf, pavg, c->ptype, perm.data(), n, B.data(), AB.data(), C2.data(), Dv.data(), PT.data(), c->rho, 0u); });
    TvfArgs ta;
    memset(&ta, 0, sizeof(ta));
    ta.AB = AB.data(); ta.C2 = C2.data(); ta.Dv = Dv.data(); ta.PT = PT.data(); ta.perm = perm.data();
    ta.rho = c->rho; ta.V = V; ta.pavg = pavg;
    ta.au = au; ta.av = av; ta.aw = aw; ta.auhat = auhat; ta.avhat = avhat; ta.awhat = awhat; ta.ap = ap;
    ta.n = n;
    ta.cellx = ta.celly = ta.cellz = 1.0f;
    ta.k2 = (float)(c->radius_scale * c->radius_scale);
    ta.kfac = (float)c->kfac;
    ta.fluid_mask = prog->fluid_mask; ta.src_mask = prog->fluid_mask; ta.eqbits = prog->eqbits; ta.bql = prog->bql;
    ta.pb = (float)prog->pb; ta.nu = (float)prog->nu; ta.edac_nu = (float)prog->edac_nu;
    ta.c0 = (float)prog->c0; ta.alpha = (float)prog->alpha;
    double damp = 1.0;
    if (prog->t < prog->tdamp) damp = 0.5 * (sin((-0.5 + prog->t / prog->tdamp) * 3.14159265358979323846) + 1.0);
    ta.gx = (float)(prog->gx * damp); ta.gy = (float)(prog->gy * damp); ta.gz = (float)(prog->gz * damp);
    ta.pair_counter = nullptr;
    const int kernel = c->kernel, dim = c->dim;
    DISPATCH(run_tvf, ta, cnt.data(), lst.data(), capg, prog->passes)
    (void)pairs;
    return 0;
}

// k_stage_solid over n real particles of array 0: f64 = x y z u v w rho s[6], then the
// *0 copies in the same order; f32 = au av aw ax ay az arho as[6]
int emul_stage_solid(long long n, int which, double dt, double *const f64[26], const float *const f32[13])
{
    std::vector<uint8_t> ptype((size_t)n, 0);
    StageSolidArgs a;
    memset(&a, 0, sizeof(a));
    a.x = f64[0]; a.y = f64[1]; a.z = f64[2]; a.u = f64[3]; a.v = f64[4]; a.w = f64[5]; a.rho = f64[6];
    for (int k = 0; k < 6; k++) a.s[k] = f64[7 + k];
    a.x0 = f64[13]; a.y0 = f64[14]; a.z0 = f64[15]; a.u0 = f64[16]; a.v0 = f64[17]; a.w0 = f64[18]; a.rho0 = f64[19];
    for (int k = 0; k < 6; k++) a.s0[k] = f64[20 + k];
    a.au = f32[0]; a.av = f32[1]; a.aw = f32[2]; a.ax = f32[3]; a.ay = f32[4]; a.az = f32[5]; a.arho = f32[6];
    for (int k = 0; k < 6; k++) a.as[k] = f32[7 + k];
    a.ptype = ptype.data();
    a.pool_end = n;
    a.arr = -1;
    a.which = which;
    a.f = which == 1 ? 0.5 * dt : dt;
    launch(n, 256, [&] { k_stage_solid(a); });
    return 0;
}

int emul_solid(const emul_common *c, const b200sph_solid_program *prog, double *const s[6], float *p, const float *cs,
               float *const vg[9], float *const r[6], float *const as[6], float *arho, float *au, float *av, float *aw,
               float *ax, float *ay, float *az)
{
    const long long n = c->n;
    std::vector<uint32_t> cnt, lst, perm((size_t)n);
    int capg;
    make_lists(n, cnt, lst, capg);
    for (long long i = 0; i < n; i++) perm[i] = (uint32_t)i;
    std::vector<float4> AB, B((size_t)n), C3((size_t)n), T01((size_t)n), T2R((size_t)n), R2((size_t)n);
    pack_A(*c, AB);
    launch(n, 256, [&] { k_pack_solid(c->u, c->v, c->w, c->m, c->rho, p, cs, c->ptype, perm.data(), n, B.data(), AB.data(), C3.data()); });
    SolidArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.AB = AB.data(); sa.C3 = C3.data(); sa.T01 = T01.data(); sa.T2R = T2R.data(); sa.R2 = R2.data(); sa.perm = perm.data();
    sa.rho = c->rho;
    for (int k = 0; k < 6; k++) { sa.s[k] = s[k]; sa.r[k] = r[k]; sa.as[k] = as[k]; }
    for (int k = 0; k < 9; k++) sa.vg[k] = vg[k];
    sa.p = p;
    sa.arho = arho; sa.au = au; sa.av = av; sa.aw = aw; sa.ax = ax; sa.ay = ay; sa.az = az;
    sa.n = n;
    sa.cellx = sa.celly = sa.cellz = 1.0f;
    sa.k2 = (float)(c->radius_scale * c->radius_scale);
    sa.kfac = (float)c->kfac;
    sa.elastic_mask = prog->elastic_mask; sa.grad3d = prog->grad3d;
    sa.source_mask = prog->source_mask ? prog->source_mask : prog->elastic_mask; sa.ghost_group1 = prog->ghost_group1;
    sa.eps = (float)prog->eps; sa.alpha = (float)prog->alpha; sa.beta = (float)prog->beta; sa.eps_xsph = (float)prog->eps_xsph;
    for (int a = 0; a < B200SPH_MAX_ARRAYS; a++) {
        sa.c0_ref[a] = prog->c0_ref[a]; sa.rho_ref[a] = prog->rho_ref[a]; sa.G[a] = prog->G[a];
        sa.wdeltap[a] = (float)prog->wdeltap[a]; sa.nexp[a] = (float)prog->n[a];
    }
    const int kernel = c->kernel, dim = c->dim;
    DISPATCH(run_solid, sa, cnt.data(), lst.data(), capg, prog->passes)
    return 0;
}
}

// scan.cuh -- three-phase exclusive scan.
// Part of the single translation unit b200sph.cu (included there, in this order; not a
// stand-alone header).

// --------------------------------------------------------------------------
// exclusive scan (u32), three-phase; n up to 2^28 + 1
// --------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *total)
{
    // exclusive scan of one value per thread across the block
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t s = (lane < SCAN_THREADS / 32) ? wsum[lane] : 0;
        uint32_t si = s;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, si, o);
            if (lane >= o) si += t;
        }
        if (lane < SCAN_THREADS / 32) wsum[lane] = si - s;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1) *total = si;
    }
    __syncthreads();
    uint32_t r = wsum[w] + inc - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, long long n,
             uint32_t *__restrict__ blk_sums)
{
    __shared__ uint32_t total;
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    uint32_t ex = block_excl_scan(s, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) blk_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(uint32_t *blk_sums, long long nb)
{
    __shared__ uint32_t total;
    uint32_t carry = 0;
    for (long long b0 = 0; b0 < nb; b0 += SCAN_THREADS) {
        long long i = b0 + threadIdx.x;
        uint32_t v = (i < nb) ? blk_sums[i] : 0u;
        uint32_t ex = block_excl_scan(v, &total);
        if (i < nb) blk_sums[i] = ex + carry;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add(uint32_t *__restrict__ out, long long n, const uint32_t *__restrict__ blk_sums)
{
    const uint32_t add = blk_sums[blockIdx.x];
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
}

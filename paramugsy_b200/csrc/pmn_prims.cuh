// pmn_prims.cuh — hand-written device-wide primitives used by every stage:
//   * scan (inclusive/exclusive, any associative op) — reduce / spine / downsweep
//   * ordered stream compaction built on it
//   * stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass
// All launches go to the caller's stream; no host synchronisation inside.
// Grids are sized in multiples of the SM count where the work allows (148 on B200).
#pragma once
#include "pmn_common.cuh"

#ifdef __CUDACC__

// ------------------------------------------------------------------------------------ scan

#define PMN_SCAN_THREADS 256
#define PMN_SCAN_ITEMS 16
#define PMN_SCAN_TILE (PMN_SCAN_THREADS * PMN_SCAN_ITEMS)

struct OpAddU32 { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a + b; } static __device__ __forceinline__ uint32_t identity() { return 0u; } };
struct OpMaxI32 { __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; } static __device__ __forceinline__ int32_t identity() { return INT32_MIN; } };
// segmented max over (flag, value) packed as int64: flag in bit 62.. is handled by callers

template <class T, class Op>
__device__ __forceinline__ T pmn_block_scan_incl(T v, Op op, T *smem /* [32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { T u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = op(u, v); }
    if (lane == 31) smem[warp] = v;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (blockDim.x >> 5) ? smem[lane] : Op::identity();
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w = op(u, w); }
        smem[lane] = w;
    }
    __syncthreads();
    if (warp > 0) v = op(smem[warp - 1], v);
    __syncthreads();
    return v;
}

// phase 1: one partial per tile
template <class T, class Op>
__global__ void __launch_bounds__(PMN_SCAN_THREADS) pmn_scan_reduce(const T *__restrict__ in, T *__restrict__ partial, int64_t n)
{
    __shared__ T sm[32];
    Op op; T acc = Op::identity();
    int64_t base = (int64_t)blockIdx.x * PMN_SCAN_TILE + (int64_t)threadIdx.x * PMN_SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < PMN_SCAN_ITEMS; k++) if (base + k < n) acc = op(acc, in[base + k]);
    acc = pmn_block_scan_incl(acc, op, sm);
    if (threadIdx.x == PMN_SCAN_THREADS - 1) partial[blockIdx.x] = acc;
}

// phase 2: one block scans the partials in place (exclusive), looping over them
template <class T, class Op>
__global__ void __launch_bounds__(1024) pmn_scan_spine(T *__restrict__ partial, int64_t m)
{
    __shared__ T sm[32]; __shared__ T carry_s;
    Op op;
    if (threadIdx.x == 0) carry_s = Op::identity();
    __syncthreads();
    for (int64_t base = 0; base < m; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        T v = i < m ? partial[i] : Op::identity();
        T inc = pmn_block_scan_incl(v, op, sm);
        T carry = carry_s;
        // exclusive = carry (+) inclusive-of-previous; the shuffle runs on all lanes
        T prev = __shfl_up_sync(0xffffffffu, inc, 1);
        if ((threadIdx.x & 31) == 0) prev = threadIdx.x ? sm[(threadIdx.x >> 5) - 1] : Op::identity();
        __syncthreads();
        if (i < m) partial[i] = threadIdx.x ? op(carry, prev) : carry;
        if (threadIdx.x == blockDim.x - 1) carry_s = op(carry, inc);
        __syncthreads();
    }
}

// phase 3: rescan each tile with its offset
template <class T, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(PMN_SCAN_THREADS) pmn_scan_down(const T *__restrict__ in, T *__restrict__ out, const T *__restrict__ partial, int64_t n)
{
    __shared__ T sm[32];
    Op op; T loc[PMN_SCAN_ITEMS];
    int64_t base = (int64_t)blockIdx.x * PMN_SCAN_TILE + (int64_t)threadIdx.x * PMN_SCAN_ITEMS;
    T acc = Op::identity();
#pragma unroll
    for (int k = 0; k < PMN_SCAN_ITEMS; k++) { loc[k] = base + k < n ? in[base + k] : Op::identity(); acc = op(acc, loc[k]); }
    T inc = pmn_block_scan_incl(acc, op, sm);
    // exclusive prefix of this thread = (tile offset) (+) (inclusive of previous thread)
    T prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if ((threadIdx.x & 31) == 0) prev = threadIdx.x ? sm[(threadIdx.x >> 5) - 1] : Op::identity();
    T run = threadIdx.x ? op(partial[blockIdx.x], prev) : partial[blockIdx.x];
#pragma unroll
    for (int k = 0; k < PMN_SCAN_ITEMS; k++) {
        if (base + k < n) {
            if (INCLUSIVE) { run = op(run, loc[k]); out[base + k] = run; }
            else { out[base + k] = run; run = op(run, loc[k]); }
        }
    }
}

// short arrays (most scans of a pair: a few hundred clusters, a few thousand tiles, 10^5 matches): one block walks the
// array with a running carry — one launch instead of three; with several pairs in flight per GPU the stages are chains of
// small kernels whose cost is the launch, not the work
#define PMN_SCAN_SMALL_ITEMS 8
#define PMN_SCAN_SMALL_MAX (1024 * PMN_SCAN_SMALL_ITEMS)   /* one pass of the block: longer arrays are faster on three wide kernels */
template <class T, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(1024) pmn_scan_small(const T *__restrict__ in, T *__restrict__ out, int64_t n)
{
    __shared__ T sm[32]; __shared__ T carry_s;
    Op op;
    if (threadIdx.x == 0) carry_s = Op::identity();
    __syncthreads();
    for (int64_t cbase = 0; cbase < n; cbase += 1024 * PMN_SCAN_SMALL_ITEMS) {
        const int64_t base = cbase + (int64_t)threadIdx.x * PMN_SCAN_SMALL_ITEMS;
        T loc[PMN_SCAN_SMALL_ITEMS]; T acc = Op::identity();
#pragma unroll
        for (int k = 0; k < PMN_SCAN_SMALL_ITEMS; k++) { loc[k] = base + k < n ? in[base + k] : Op::identity(); acc = op(acc, loc[k]); }
        const T inc = pmn_block_scan_incl(acc, op, sm);
        T prev = __shfl_up_sync(0xffffffffu, inc, 1);
        if ((threadIdx.x & 31) == 0) prev = threadIdx.x ? sm[(threadIdx.x >> 5) - 1] : Op::identity();
        const T carry = carry_s;
        T run = threadIdx.x ? op(carry, prev) : carry;
#pragma unroll
        for (int k = 0; k < PMN_SCAN_SMALL_ITEMS; k++) {
            if (base + k < n) {
                if (INCLUSIVE) { run = op(run, loc[k]); out[base + k] = run; }
                else { out[base + k] = run; run = op(run, loc[k]); }
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = op(carry, inc);
        __syncthreads();
    }
}

// scratch must hold ceil(n / PMN_SCAN_TILE) + 1 elements of T.  in == out is allowed.
template <class T, class Op, bool INCLUSIVE>
static inline void pmn_scan(const T *in, T *out, int64_t n, T *scratch, cudaStream_t st)
{
    if (n <= 0) return;
    if (n <= PMN_SCAN_SMALL_MAX) { pmn_scan_small<T, Op, INCLUSIVE><<<1, 1024, 0, st>>>(in, out, n); pmn_tls_launches_saved += 2; return; }
    int64_t tiles = (n + PMN_SCAN_TILE - 1) / PMN_SCAN_TILE;
    pmn_scan_reduce<T, Op><<<(unsigned)tiles, PMN_SCAN_THREADS, 0, st>>>(in, scratch, n);
    pmn_scan_spine<T, Op><<<1, 1024, 0, st>>>(scratch, tiles);
    pmn_scan_down<T, Op, INCLUSIVE><<<(unsigned)tiles, PMN_SCAN_THREADS, 0, st>>>(in, out, scratch, n);
}
static inline size_t pmn_scan_scratch_elems(int64_t n) { return (size_t)((n + PMN_SCAN_TILE - 1) / PMN_SCAN_TILE + 1); }

// ------------------------------------------------------------------------------------ radix sort

#define PMN_RS_THREADS 256
#define PMN_RS_WARPS (PMN_RS_THREADS / 32)
#define PMN_RS_ITEMS 16                                   /* keys per lane */
#define PMN_RS_TILE (PMN_RS_THREADS * PMN_RS_ITEMS)       /* 4096 keys per block */
#define PMN_RS_RADIX 256

// histogram of one digit per tile: hist[d * ntiles + tile]
static __global__ void __launch_bounds__(PMN_RS_THREADS) pmn_rs_hist(const uint64_t *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int ntiles)
{
    __shared__ uint32_t h[PMN_RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * PMN_RS_TILE;
#pragma unroll
    for (int k = 0; k < PMN_RS_ITEMS; k++) {
        int64_t i = base + k * PMN_RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter.  Warp w of the block owns keys [base + w*512, base + (w+1)*512) and walks
// them 32 at a time in order; lanes with equal digits are ranked by lane id (match_any).
static __global__ void __launch_bounds__(PMN_RS_THREADS) pmn_rs_scatter(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                 uint64_t *__restrict__ kout, uint32_t *__restrict__ vout,
                                                                 int64_t n, int shift, const uint32_t *__restrict__ offs, int ntiles)
{
    __shared__ uint32_t cnt[PMN_RS_WARPS][PMN_RS_RADIX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < PMN_RS_WARPS * PMN_RS_RADIX; k += PMN_RS_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * PMN_RS_TILE + (int64_t)warp * (32 * PMN_RS_ITEMS);
    uint64_t key[PMN_RS_ITEMS]; uint32_t val[PMN_RS_ITEMS];
    const unsigned lt = pmn_lanemask_lt();
#pragma unroll
    for (int k = 0; k < PMN_RS_ITEMS; k++) {
        int64_t i = wbase + k * 32 + lane;
        bool ok = i < n;
        key[k] = ok ? kin[i] : ~0ull; val[k] = ok ? vin[i] : 0u;
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned d = (unsigned)(key[k] >> shift) & 0xff;
            unsigned m = __match_any_sync(act, d);
            if ((m & lt) == 0) cnt[warp][d] += __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit threadIdx.x: global offset of this tile, then exclusive prefix across the warps
        unsigned d = threadIdx.x;
        uint32_t run = offs[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < PMN_RS_WARPS; w++) { uint32_t t = cnt[w][d]; cnt[w][d] = run; run += t; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < PMN_RS_ITEMS; k++) {
        int64_t i = wbase + k * 32 + lane;
        bool ok = i < n;
        unsigned act = __ballot_sync(0xffffffffu, ok);
        if (ok) {
            unsigned d = (unsigned)(key[k] >> shift) & 0xff;
            unsigned m = __match_any_sync(act, d);
            uint32_t pos = cnt[warp][d] + __popc(m & lt);
            kout[pos] = key[k]; vout[pos] = val[k];
            __syncwarp(m);
            if ((m & lt) == 0) cnt[warp][d] += __popc(m);
        }
        __syncwarp();
    }
}

struct RadixScratch {
    DevBuf hist, spine;
    int reserve(int64_t n)
    {
        int64_t tiles = (n + PMN_RS_TILE - 1) / PMN_RS_TILE; if (tiles < 1) tiles = 1;
        int64_t cells = tiles * PMN_RS_RADIX;
        if (hist.ensure(sizeof(uint32_t) * (size_t)cells)) return -1;
        if (spine.ensure(sizeof(uint32_t) * pmn_scan_scratch_elems(cells))) return -1;
        return 0;
    }
};

// Sorts bits [first_shift, nbits) of the keys, stable.  Ping-pongs between (k0,v0) and (k1,v1);
// returns 0 if the result is in (k0,v0), 1 if in (k1,v1), negative on error.
static inline int pmn_radix_sort(uint64_t *k0, uint32_t *v0, uint64_t *k1, uint32_t *v1, int64_t n, int nbits,
                                 RadixScratch &rs, cudaStream_t st, int *launches = nullptr, int first_shift = 0)
{
    if (n <= 1 || nbits <= 0) return 0;
    if (rs.reserve(n)) return -1;
    int ntiles = (int)((n + PMN_RS_TILE - 1) / PMN_RS_TILE);
    int64_t cells = (int64_t)ntiles * PMN_RS_RADIX;
    int cur = 0;
    for (int shift = first_shift; shift < nbits; shift += 8) {       // first_shift > 0: the input is already ordered by the bits below it
        uint64_t *ki = cur ? k1 : k0, *ko = cur ? k0 : k1; uint32_t *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
        pmn_rs_hist<<<ntiles, PMN_RS_THREADS, 0, st>>>(ki, n, shift, rs.hist.as<uint32_t>(), ntiles);
        pmn_scan<uint32_t, OpAddU32, false>(rs.hist.as<uint32_t>(), rs.hist.as<uint32_t>(), cells, rs.spine.as<uint32_t>(), st);
        pmn_rs_scatter<<<ntiles, PMN_RS_THREADS, 0, st>>>(ki, vi, ko, vo, n, shift, rs.hist.as<uint32_t>(), ntiles);
        if (launches) *launches += 5;
        cur ^= 1;
    }
    return cur;
}

#endif  // __CUDACC__

// pmn_prims.cuh — hand-written device-wide primitives used by every stage:
//   * scan (inclusive/exclusive, any associative op) — one pass with decoupled look-back
//   * ordered stream compaction built on it
//   * stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass
// All launches go to the caller's stream; no host synchronisation inside.
// Grids are sized in multiples of the SM count where the work allows (148 on B200).
#pragma once
#include <atomic>
#include <cstring>

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

// One pass over the data (decoupled look-back): a block takes the next tile (a ticket from a device counter, so every tile a
// block waits for belongs to a block that is already running), scans it, publishes the tile's aggregate, adds up the
// aggregates of the tiles before it (warp 0, 32 descriptors per step, back to the nearest tile whose inclusive prefix is
// already known), publishes its own inclusive prefix and writes its part of the output.  One launch that reads the input
// once instead of reduce / spine / downsweep: 2 n elements of traffic instead of 3 n and a third of the launches.
//   A descriptor is {status, aggregate, inclusive}.  status = 2 * epoch (aggregate valid) or 2 * epoch + 1 (inclusive valid
//   too); every scan of the process has its own epoch (a host counter), scratch memory is zeroed when it is allocated
//   (DevBuf::grow_to), so whatever an earlier scan left in the scratch reads as "not yet".  The two counters in front (ticket,
//   finished tiles) are put back to zero by the block that finishes last.
struct PmnScanDesc { unsigned long long status, agg, incl; };

template <class T> __device__ __forceinline__ unsigned long long pmn_scan_pack(T v) { unsigned long long u = 0; memcpy(&u, &v, sizeof(T)); return u; }
template <class T> __device__ __forceinline__ T pmn_scan_unpack(unsigned long long u) { T v; memcpy(&v, &u, sizeof(T)); return v; }

template <class T, class Op, bool INCLUSIVE>
__global__ void __launch_bounds__(PMN_SCAN_THREADS) pmn_scan_onepass(const T *__restrict__ in, T *__restrict__ out, int64_t n, unsigned *__restrict__ counters,
                                                                    PmnScanDesc *__restrict__ desc, unsigned long long epoch, unsigned tiles)
{
    __shared__ T sm[32]; __shared__ T s_prefix; __shared__ unsigned s_tile;
    __shared__ T s_wval[PMN_SCAN_THREADS / 32]; __shared__ bool s_wincl[PMN_SCAN_THREADS / 32];
    Op op;
    if (threadIdx.x == 0) s_tile = atomicAdd(counters, 1u);
    __syncthreads();
    const unsigned tile = s_tile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T loc[PMN_SCAN_ITEMS];
    const int64_t base = (int64_t)tile * PMN_SCAN_TILE + (int64_t)threadIdx.x * PMN_SCAN_ITEMS;
    T acc = Op::identity();
    // A thread owns 16 consecutive elements.  Read one by one, a warp's load touches 32 sectors for 128 bytes and the kernel
    // lives on L1 hits (a 5 M-element scan: 50 us for 40 MB); with 16-byte loads and stores it is 4 (or 8) instructions per
    // thread instead of 16.  Needs 16-byte aligned arrays and a full chunk; the tail and odd pointers take the scalar path.
    constexpr int PER = 16 / (int)sizeof(T);
    const bool vec = (sizeof(T) == 4 || sizeof(T) == 8) && ((((uintptr_t)in) | ((uintptr_t)out)) & 15) == 0 && base + PMN_SCAN_ITEMS <= n;
    if (vec) {
        const uint4 *vin = reinterpret_cast<const uint4 *>(in + base);
#pragma unroll
        for (int j = 0; j < PMN_SCAN_ITEMS / PER; j++) {
            const uint4 q = vin[j];
            if constexpr (sizeof(T) == 4) {
                loc[4 * j + 0] = pmn_scan_unpack<T>(q.x); loc[4 * j + 1] = pmn_scan_unpack<T>(q.y); loc[4 * j + 2] = pmn_scan_unpack<T>(q.z); loc[4 * j + 3] = pmn_scan_unpack<T>(q.w);
            } else {
                loc[2 * j + 0] = pmn_scan_unpack<T>((unsigned long long)q.x | ((unsigned long long)q.y << 32));
                loc[2 * j + 1] = pmn_scan_unpack<T>((unsigned long long)q.z | ((unsigned long long)q.w << 32));
            }
        }
#pragma unroll
        for (int k = 0; k < PMN_SCAN_ITEMS; k++) acc = op(acc, loc[k]);
    } else {
#pragma unroll
        for (int k = 0; k < PMN_SCAN_ITEMS; k++) { loc[k] = base + k < n ? in[base + k] : Op::identity(); acc = op(acc, loc[k]); }
    }
    const T inc = pmn_block_scan_incl(acc, op, sm);         // sm[w] = inclusive total of warps 0..w afterwards
    const T block_agg = sm[PMN_SCAN_THREADS / 32 - 1];
    volatile unsigned long long *my = (volatile unsigned long long *)&desc[tile];
    if (threadIdx.x == 0) {
        if (tile == 0) { my[2] = pmn_scan_pack(block_agg); __threadfence(); my[0] = 2 * epoch + 1; }
        else { my[1] = pmn_scan_pack(block_agg); __threadfence(); my[0] = 2 * epoch; }
    }
    if (tile > 0) {
        // look-back by the whole block: warp w looks at the 32 tiles [look - 32 w - 31, look - 32 w], nearest first, so one step
        // covers 256 tiles.  (With warp 0 alone a 5 M-element scan — 1224 tiles that all start together — was a chain of 38
        // dependent steps, 40 of its 50 us.)
        T running = Op::identity();                       // aggregate of the tiles looked at so far, in tile order
        for (int64_t look = (int64_t)tile - 1;; look -= PMN_SCAN_THREADS) {
            const int64_t t = look - threadIdx.x;         // thread 0 looks at the nearest tile
            unsigned long long st = 2 * epoch + 1; T val = Op::identity();       // before tile 0: an inclusive prefix of nothing
            if (t >= 0) {
                volatile unsigned long long *d = (volatile unsigned long long *)&desc[t];
                do { st = d[0]; } while ((st >> 1) != epoch);
                __threadfence();
                val = pmn_scan_unpack<T>((st & 1ull) ? d[2] : d[1]);
            }
            const unsigned incl_mask = __ballot_sync(0xffffffffu, (st & 1ull) != 0);
            const int first = incl_mask ? __ffs((int)incl_mask) - 1 : 31;          // nearest lane that holds an inclusive prefix
            T v = lane <= first ? val : Op::identity();
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const T u = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = op(u, v); }   // lane 0: val[first] (+) ... (+) val[0]
            if (lane == 0) { s_wval[warp] = v; s_wincl[warp] = incl_mask != 0; }
            __syncthreads();
            T window = Op::identity(); bool found = false;
#pragma unroll
            for (int w = 0; w < PMN_SCAN_THREADS / 32; w++)
                if (!found) { window = op(s_wval[w], window); found = s_wincl[w]; }     // older tiles on the left
            running = op(window, running);
            __syncthreads();                              // s_wval / s_wincl are rewritten by the next step
            if (found) break;
        }
        if (threadIdx.x == 0) {
            s_prefix = running;
            my[2] = pmn_scan_pack(op(running, block_agg)); __threadfence(); my[0] = 2 * epoch + 1;
        }
    }
    __syncthreads();
    // exclusive prefix of this thread = (tile prefix) (+) (inclusive of the previous thread)
    T prev = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) prev = warp ? sm[warp - 1] : Op::identity();
    T run = threadIdx.x ? prev : Op::identity();
    if (tile > 0) run = threadIdx.x ? op(s_prefix, prev) : s_prefix;
    if (vec) {
#pragma unroll
        for (int k = 0; k < PMN_SCAN_ITEMS; k++) {
            const T x = loc[k];
            if (INCLUSIVE) { run = op(run, x); loc[k] = run; }
            else { loc[k] = run; run = op(run, x); }
        }
        uint4 *vout = reinterpret_cast<uint4 *>(out + base);
#pragma unroll
        for (int j = 0; j < PMN_SCAN_ITEMS / PER; j++) {
            uint4 q;
            if constexpr (sizeof(T) == 4) {
                q.x = (unsigned)pmn_scan_pack(loc[4 * j + 0]); q.y = (unsigned)pmn_scan_pack(loc[4 * j + 1]); q.z = (unsigned)pmn_scan_pack(loc[4 * j + 2]); q.w = (unsigned)pmn_scan_pack(loc[4 * j + 3]);
            } else {
                const unsigned long long a = pmn_scan_pack(loc[2 * j + 0]), b = pmn_scan_pack(loc[2 * j + 1]);
                q.x = (unsigned)a; q.y = (unsigned)(a >> 32); q.z = (unsigned)b; q.w = (unsigned)(b >> 32);
            }
            vout[j] = q;
        }
    } else {
#pragma unroll
        for (int k = 0; k < PMN_SCAN_ITEMS; k++) {
            if (base + k < n) {
                if (INCLUSIVE) { run = op(run, loc[k]); out[base + k] = run; }
                else { out[base + k] = run; run = op(run, loc[k]); }
            }
        }
    }
    if (threadIdx.x == 0) {                                // the block that finishes last rearms the counters for the next scan
        const unsigned done = atomicAdd(counters + 1, 1u);
        if (done == tiles - 1) { counters[0] = 0; counters[1] = 0; }
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

inline std::atomic<unsigned long long> pmn_scan_epoch{0};      // one epoch per scan of the process (C++17 inline variable: one object for all translation units)

// scratch: 8 * pmn_scan_scratch_elems(n) bytes of a DevBuf (zeroed when allocated); consecutive scans on one stream may share
// it.  in == out is allowed.  Every scan is ONE launch (callers that still count three note the difference in
// pmn_tls_launches_saved, which the API entry points subtract from the context's launch counter).
template <class T, class Op, bool INCLUSIVE>
static inline void pmn_scan(const T *in, T *out, int64_t n, T *scratch, cudaStream_t st)
{
    static_assert(sizeof(T) <= 8, "descriptor slots are 8 bytes");
    if (n <= 0) return;
    pmn_tls_launches_saved += 2;
    if (n <= PMN_SCAN_SMALL_MAX) { pmn_scan_small<T, Op, INCLUSIVE><<<1, 1024, 0, st>>>(in, out, n); return; }
    const int64_t tiles = (n + PMN_SCAN_TILE - 1) / PMN_SCAN_TILE;
    const unsigned long long epoch = ++pmn_scan_epoch;
    unsigned *counters = (unsigned *)scratch;
    PmnScanDesc *desc = (PmnScanDesc *)((char *)scratch + 16);
    pmn_scan_onepass<T, Op, INCLUSIVE><<<(unsigned)tiles, PMN_SCAN_THREADS, 0, st>>>(in, out, n, counters, desc, epoch, (unsigned)tiles);
}
static inline size_t pmn_scan_scratch_elems(int64_t n) { return (size_t)(3 * ((n + PMN_SCAN_TILE - 1) / PMN_SCAN_TILE) + 4); }   // in units of 8 bytes

// ------------------------------------------------------------------------------------ radix sort

#define PMN_RS_RADIX 256
#define PMN_RS_TILE 4096                                  /* keys per block of the tiled sort */
#define PMN_RS_HIST_THREADS 256
#define PMN_RS_WARPS 16                                   /* scatter kernel: 16 warps x 8 keys per lane */
#define PMN_RS_ITEMS (PMN_RS_TILE / (PMN_RS_WARPS * 32))

// histogram of one digit per tile: hist[d * ntiles + tile]
template <class KeyT>
static __global__ void __launch_bounds__(PMN_RS_HIST_THREADS) pmn_rs_hist(const KeyT *__restrict__ keys, int64_t n, int shift, uint32_t *__restrict__ hist, int ntiles)
{
    __shared__ uint32_t h[PMN_RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * PMN_RS_TILE;
#pragma unroll
    for (int k = 0; k < PMN_RS_TILE / PMN_RS_HIST_THREADS; k++) {
        int64_t i = base + k * PMN_RS_HIST_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 0xff], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// lanes of `act` that hold the same 8-bit digit as the caller: eight ballots (MATCH.ANY walks the distinct values of the
// warp one after the other and was what the scatter kernel waited for)
__device__ __forceinline__ unsigned pmn_match_digit(unsigned act, unsigned d)
{
    unsigned peers = act;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return peers;
}

// One pass of the stable sort over one tile, shared by the scatter kernel and the one-block sort.  Warp w owns keys
// [w*32*ITEMS, (w+1)*32*ITEMS) of the tile and walks them 32 at a time in order; lanes with equal digits are ranked by
// lane id.  The keys are first placed in shared memory at their rank inside the tile, so that what leaves the block are
// runs of consecutive addresses per digit (a tile of 4096 random keys: 16 keys per run) instead of 32 unrelated stores per
// warp instruction.
//   cnt[w][d]: scratch, NWARPS x 256;  on return skey/sval hold the tile in sorted order and, when dstart is not null,
//   dstart[d] = first slot of digit d inside the tile.  Every thread of the block must call it.
template <class KeyT, int NWARPS, int ITEMS>
__device__ __forceinline__ void pmn_rs_tile_pass(const KeyT (&key)[ITEMS], const uint32_t (&val)[ITEMS], unsigned okmask, int shift,
                                                 KeyT *skey, uint32_t *sval, uint32_t (*cnt)[PMN_RS_RADIX], uint32_t *dstart, uint32_t *stmp)
{
    const int warp = threadIdx.x >> 5;
    const unsigned lt = pmn_lanemask_lt();
    for (int k = threadIdx.x; k < NWARPS * PMN_RS_RADIX; k += NWARPS * 32) (&cnt[0][0])[k] = 0;
    __syncthreads();
    uint32_t rank[ITEMS];           // rank of the key among the keys of its warp with the same digit
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const bool ok = (okmask >> k) & 1u;
        const unsigned act = __ballot_sync(0xffffffffu, ok);
        const unsigned d = (unsigned)(key[k] >> shift) & 0xff;
        const unsigned peers = pmn_match_digit(act, d);
        uint32_t old = 0;
        const bool leader = ok && (peers & lt) == 0;
        if (leader) { old = cnt[warp][d]; cnt[warp][d] = old + __popc(peers); }
        old = __shfl_sync(0xffffffffu, old, ok ? __ffs(peers) - 1 : 0);
        rank[k] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: exclusive prefix across the warps, then across the digits
        uint32_t run = 0;
        const unsigned d = threadIdx.x;
        if (d < PMN_RS_RADIX) {
#pragma unroll
            for (int w = 0; w < NWARPS; w++) { uint32_t t = cnt[w][d]; cnt[w][d] = run; run += t; }
        }
        const uint32_t incl = pmn_block_scan_incl(run, OpAddU32(), stmp);
        if (d < PMN_RS_RADIX) {
            const uint32_t excl = incl - run;
#pragma unroll
            for (int w = 0; w < NWARPS; w++) cnt[w][d] += excl;
            if (dstart) dstart[d] = excl;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        if ((okmask >> k) & 1u) {
            const uint32_t pos = cnt[warp][(unsigned)(key[k] >> shift) & 0xff] + rank[k];
            skey[pos] = key[k]; sval[pos] = val[k];
        }
    }
    __syncthreads();
}

#define PMN_RS_SMEM_BYTES(KeyT, NWARPS, ITEMS) ((size_t)(NWARPS) * 32 * (ITEMS) * (sizeof(KeyT) + 4) + ((size_t)(NWARPS) * PMN_RS_RADIX + PMN_RS_RADIX + 32) * 4)

// stable scatter of one tile per block; offs = exclusive scan of the per-tile histograms
template <class KeyT>
static __global__ void __launch_bounds__(PMN_RS_WARPS * 32) pmn_rs_scatter(const KeyT *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                      KeyT *__restrict__ kout, uint32_t *__restrict__ vout,
                                                                      int64_t n, int shift, const uint32_t *__restrict__ offs, int ntiles)
{
    extern __shared__ __align__(16) unsigned char pmn_rs_smem[];
    KeyT *skey = (KeyT *)pmn_rs_smem;
    uint32_t *sval = (uint32_t *)(skey + PMN_RS_TILE);
    uint32_t (*cnt)[PMN_RS_RADIX] = (uint32_t (*)[PMN_RS_RADIX])(sval + PMN_RS_TILE);
    uint32_t *dstart = &cnt[0][0] + PMN_RS_WARPS * PMN_RS_RADIX;
    uint32_t *stmp = dstart + PMN_RS_RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t tbase = (int64_t)blockIdx.x * PMN_RS_TILE;
    const int64_t wbase = tbase + (int64_t)warp * (32 * PMN_RS_ITEMS);
    KeyT key[PMN_RS_ITEMS]; uint32_t val[PMN_RS_ITEMS]; unsigned okmask = 0;
#pragma unroll
    for (int k = 0; k < PMN_RS_ITEMS; k++) {
        const int64_t i = wbase + k * 32 + lane;
        const bool ok = i < n;
        key[k] = ok ? kin[i] : (KeyT)0; val[k] = ok ? vin[i] : 0u;
        okmask |= (ok ? 1u : 0u) << k;
    }
    pmn_rs_tile_pass<KeyT, PMN_RS_WARPS, PMN_RS_ITEMS>(key, val, okmask, shift, skey, sval, cnt, dstart, stmp);
    // dstart[d] becomes (global offset of this tile's run of digit d) - (its first slot in the tile)
    if (threadIdx.x < PMN_RS_RADIX) dstart[threadIdx.x] = offs[(size_t)threadIdx.x * ntiles + blockIdx.x] - dstart[threadIdx.x];
    __syncthreads();
    const int count = (int)(n - tbase < PMN_RS_TILE ? n - tbase : PMN_RS_TILE);
#pragma unroll 4
    for (int j = threadIdx.x; j < count; j += PMN_RS_WARPS * 32) {
        const KeyT kk = skey[j];
        const uint32_t g = dstart[(unsigned)(kk >> shift) & 0xff] + (uint32_t)j;
        kout[g] = kk; vout[g] = sval[j];
    }
}

// short arrays (the work lists of the doubling rounds, the cluster pieces of a pair): one block runs every pass in shared
// memory — one launch instead of three per pass
#define PMN_BS_WARPS 16
#define PMN_BS_ITEMS 16
#define PMN_BS_THREADS (PMN_BS_WARPS * 32)
#define PMN_BS_CAP (PMN_BS_THREADS * PMN_BS_ITEMS)        /* 8192 keys */
template <class KeyT>
static __global__ void __launch_bounds__(PMN_BS_THREADS) pmn_rs_block_sort(const KeyT *__restrict__ kin, const uint32_t *__restrict__ vin,
                                                                      KeyT *__restrict__ kout, uint32_t *__restrict__ vout, int n, int first_shift, int nbits)
{
    extern __shared__ __align__(16) unsigned char pmn_rs_smem[];
    KeyT *skey = (KeyT *)pmn_rs_smem;
    uint32_t *sval = (uint32_t *)(skey + PMN_BS_CAP);
    uint32_t (*cnt)[PMN_RS_RADIX] = (uint32_t (*)[PMN_RS_RADIX])(sval + PMN_BS_CAP);
    uint32_t *stmp = &cnt[0][0] + PMN_BS_WARPS * PMN_RS_RADIX + PMN_RS_RADIX;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wbase = warp * (32 * PMN_BS_ITEMS);
    KeyT key[PMN_BS_ITEMS]; uint32_t val[PMN_BS_ITEMS]; unsigned okmask = 0;
#pragma unroll
    for (int k = 0; k < PMN_BS_ITEMS; k++) {
        const int i = wbase + k * 32 + lane;
        const bool ok = i < n;
        key[k] = ok ? kin[i] : (KeyT)0; val[k] = ok ? vin[i] : 0u;
        okmask |= (ok ? 1u : 0u) << k;
    }
    for (int shift = first_shift; shift < nbits; shift += 8) {
        pmn_rs_tile_pass<KeyT, PMN_BS_WARPS, PMN_BS_ITEMS>(key, val, okmask, shift, skey, sval, cnt, nullptr, stmp);
        if (shift + 8 < nbits) {
#pragma unroll
            for (int k = 0; k < PMN_BS_ITEMS; k++) { const int i = wbase + k * 32 + lane; if (i < n) { key[k] = skey[i]; val[k] = sval[i]; } }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < n; j += PMN_BS_THREADS) { kout[j] = skey[j]; vout[j] = sval[j]; }
}

struct RadixScratch {
    DevBuf hist, spine;
    int reserve(int64_t n)
    {
        int64_t tiles = (n + PMN_RS_TILE - 1) / PMN_RS_TILE; if (tiles < 1) tiles = 1;
        int64_t cells = tiles * PMN_RS_RADIX;
        if (hist.ensure(sizeof(uint32_t) * (size_t)cells)) return -1;
        if (spine.ensure(8 * pmn_scan_scratch_elems(cells))) return -1;
        return 0;
    }
};

// the two kernels above use more than 48 KB of shared memory: opt in once per device and instantiation
template <class KeyT>
static inline int pmn_rs_prepare()
{
    static unsigned long long done = 0;         // bit per device; a race only repeats the calls
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev < 64 && ((done >> dev) & 1ull)) return 0;
    if (cudaFuncSetAttribute(pmn_rs_scatter<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PMN_RS_SMEM_BYTES(KeyT, PMN_RS_WARPS, PMN_RS_ITEMS)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(pmn_rs_block_sort<KeyT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PMN_RS_SMEM_BYTES(KeyT, PMN_BS_WARPS, PMN_BS_ITEMS)) != cudaSuccess) return -1;
    if (dev < 64) done |= 1ull << dev;
    return 0;
}

// Sorts bits [first_shift, nbits) of the keys, stable.  Ping-pongs between (k0,v0) and (k1,v1);
// returns 0 if the result is in (k0,v0), 1 if in (k1,v1), negative on error.
template <class KeyT>
static inline int pmn_radix_sort(KeyT *k0, uint32_t *v0, KeyT *k1, uint32_t *v1, int64_t n, int nbits,
                                 RadixScratch &rs, cudaStream_t st, int *launches = nullptr, int first_shift = 0)
{
    if (n <= 1 || nbits <= first_shift) return 0;
    if (pmn_rs_prepare<KeyT>()) return -1;
    if (n <= PMN_BS_CAP) {
        pmn_rs_block_sort<KeyT><<<1, PMN_BS_THREADS, PMN_RS_SMEM_BYTES(KeyT, PMN_BS_WARPS, PMN_BS_ITEMS), st>>>(k0, v0, k1, v1, (int)n, first_shift, nbits);
        if (launches) *launches += 1;
        return 1;
    }
    if (rs.reserve(n)) return -1;
    int ntiles = (int)((n + PMN_RS_TILE - 1) / PMN_RS_TILE);
    int64_t cells = (int64_t)ntiles * PMN_RS_RADIX;
    int cur = 0;
    for (int shift = first_shift; shift < nbits; shift += 8) {       // first_shift > 0: the input is already ordered by the bits below it
        KeyT *ki = cur ? k1 : k0, *ko = cur ? k0 : k1; uint32_t *vi = cur ? v1 : v0, *vo = cur ? v0 : v1;
        pmn_rs_hist<KeyT><<<ntiles, PMN_RS_HIST_THREADS, 0, st>>>(ki, n, shift, rs.hist.as<uint32_t>(), ntiles);
        pmn_scan<uint32_t, OpAddU32, false>(rs.hist.as<uint32_t>(), rs.hist.as<uint32_t>(), cells, rs.spine.as<uint32_t>(), st);
        pmn_rs_scatter<KeyT><<<ntiles, PMN_RS_WARPS * 32, PMN_RS_SMEM_BYTES(KeyT, PMN_RS_WARPS, PMN_RS_ITEMS), st>>>(ki, vi, ko, vo, n, shift, rs.hist.as<uint32_t>(), ntiles);
        if (launches) *launches += 5;
        cur ^= 1;
    }
    return cur;
}

#endif  // __CUDACC__

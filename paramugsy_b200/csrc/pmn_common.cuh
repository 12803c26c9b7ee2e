// pmn_common.cuh — shared declarations of the B200 pairwise-nucmer path (sm_100a only).
//
// Data layout in HBM:
//   packed text   uint64 words, 32 bases per word, base k of a word in bits [63-2k, 62-2k]
//                 (MSB first, so unsigned integer order of a window == lexicographic order)
//   x-mask        uint32 words, 32 bases per word, bit 31-k set when base k matches nothing
//                 (non-acgt, record separator, and everything past the end of the text)
//   both arrays carry PMN_PAD_WORDS words of padding on each side of nothing: the text is
//   zero-padded and the mask one-padded behind the last base, so a 32-base window can be
//   fetched at any position 0..n without a bounds test.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/pmn_params.h"

#define PMN_PAD_WORDS 192  /* >= words one seeding tile stages (SEED_WORDS, pmn_seed.cu: static_assert there) + slack */

struct PmnError { int code; char msg[480]; };

#define PMN_CUDA_OK(call)                                                                     \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            pmn_set_error(-2, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return -2;                                                                        \
        }                                                                                     \
    } while (0)

int pmn_set_error(int code, const char *fmt, ...);
void pmn_count_alloc(size_t bytes = 0, size_t had = 0);       // counts cudaMalloc calls (bench.py reports how many fell into the timed region)

// The stream of the context the calling thread is working for (set by every stage entry point).
// Growth of a buffer is stream-ordered on it (cudaMallocAsync / cudaFreeAsync from the device's
// default memory pool, whose release threshold pmn_ctx_create lifts): unlike cudaMalloc / cudaFree
// it does not serialise the device, so a worker that grows a buffer does not stall the others.
extern thread_local cudaStream_t pmn_tls_stream;
// callers count three launches per scan; the one-launch path for short arrays notes the difference here and the
// API entry points subtract it from the context's launch counter (bench.py's gpu_launches is a count, not an estimate)
extern thread_local long pmn_tls_launches_saved;

// ---- a grow-only device buffer: no allocation at steady state --------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    static size_t size_class(size_t bytes)
    {
        size_t want = 4096;
        if (bytes > ((size_t)64 << 20)) want = (bytes + bytes / 8 + ((size_t)64 << 20) - 1) / ((size_t)64 << 20) * ((size_t)64 << 20);
        else while (want < 2 * bytes) want <<= 1;           // 2x head room: every worker converges after its first pair
        return want;
    }
    int ensure(size_t bytes) { return bytes <= cap ? 0 : grow_to(size_class(bytes)); }
    int grow_to(size_t want)            // want: a capacity some buffer of this role already has (a size class)
    {
        if (want <= cap) return 0;
        const size_t had = cap;
        cudaStream_t ts = pmn_tls_stream;
        if (p) { if (ts) cudaFreeAsync(p, ts); else cudaFree(p); }
        // capacities are quantised (powers of two up to 64 MB, multiples of 64 MB above) so that the
        // slightly different sizes of successive pairs settle on one allocation after a few calls:
        // cudaMalloc / cudaFree serialise the whole device, which would stall every other worker
        cudaError_t e = ts ? cudaMallocAsync(&p, want, ts) : cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; cap = 0; return pmn_set_error(-3, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); }
        // new memory is zeroed once: the scan descriptors and ticket counters that live in scratch buffers (pmn_prims.cuh) rely on
        // never seeing anything but zeros or what an earlier scan of this process wrote
        e = ts ? cudaMemsetAsync(p, 0, want, ts) : cudaMemset(p, 0, want);
        if (e != cudaSuccess) { if (ts) cudaFreeAsync(p, ts); else cudaFree(p); p = nullptr; cap = 0; return pmn_set_error(-2, "cudaMemset(%zu) failed: %s", want, cudaGetErrorString(e)); }
        cap = want;
        pmn_count_alloc(want, had);
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// ---- device view of one packed sequence set ---------------------------------------------------
struct PackedView {
    const uint64_t *w;     // 2-bit text
    const uint32_t *xm;    // x-mask
    int64_t n;             // bases incl. separators
    int has_x;             // 0: plain acgt, the mask is never read
};

#ifdef __CUDACC__

__device__ __forceinline__ uint64_t pmn_window64(const uint64_t *__restrict__ w, int64_t p)
{
    int64_t k = p >> 5; int sh = (int)(p & 31) * 2;
    uint64_t a = __ldg(w + k), b = __ldg(w + k + 1);
    return sh ? (a << sh) | (b >> (64 - sh)) : a;
}

__device__ __forceinline__ uint32_t pmn_xwindow32(const uint32_t *__restrict__ xm, int64_t p)
{
    int64_t k = p >> 5; int sh = (int)(p & 31);
    uint32_t a = __ldg(xm + k), b = __ldg(xm + k + 1);
    return __funnelshift_l(b, a, sh);
}

// number of matchable bases among the 32 starting at p (stops at the first X or at the end)
__device__ __forceinline__ int pmn_valid32(const PackedView &s, int64_t p)
{
    if (p >= s.n) return 0;
    if (s.has_x) { uint32_t x = pmn_xwindow32(s.xm, p); return x ? __clz((int)x) : 32; }
    int64_t r = s.n - p;
    return r < 32 ? (int)r : 32;
}

__device__ __forceinline__ int pmn_base_at(const PackedView &s, int64_t p)   // 0..3, or 4
{
    if (p < 0 || p >= s.n) return PMN_CODE_X;
    if (s.has_x && ((__ldg(s.xm + (p >> 5)) >> (31 - (int)(p & 31))) & 1u)) return PMN_CODE_X;
    return (int)((__ldg(s.w + (p >> 5)) >> (62 - 2 * (int)(p & 31))) & 3ull);
}

// common prefix (in matchable bases) of a[pa..] and b[pb..], starting the comparison at
// offset `from` (bases before it are known equal), capped at `cap`
__device__ __forceinline__ int64_t pmn_lcp(const PackedView &a, int64_t pa, const PackedView &b, int64_t pb, int64_t from, int64_t cap)
{
    int64_t l = from;
    while (l < cap) {
        int va = pmn_valid32(a, pa + l), vb = pmn_valid32(b, pb + l);
        int v = va < vb ? va : vb;
        uint64_t x = pmn_window64(a.w, pa + l) ^ pmn_window64(b.w, pb + l);
        int m = x ? (__clzll((long long)x) >> 1) : 32;
        if (m > v) m = v;
        l += m;
        if (m < 32) break;
    }
    return l < cap ? l : cap;
}

__device__ __forceinline__ unsigned pmn_lanemask_lt() { unsigned m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

#endif  // __CUDACC__

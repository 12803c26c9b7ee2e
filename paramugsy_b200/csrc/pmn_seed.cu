// pmn_seed.cu — MUM-reference seeding of every query position, both strands.
//
// Stands in for `mummer -mumreference -b -l 20 -n` inside the `nucmer` child process of
// /root/reference/lib/nucmer/mugsy_nucmer.ml:100.  Oracle counterpart: oracle/pmn_oracle.c §4
// (seed_strand) — the anchor list must be identical, in (tag, query pos) order.
//
// Per query position i of a (record, strand) section:
//   1. 32-base window of the query from shared memory (the tile of packed query text and
//      its x-mask is staged by one TMA bulk copy per block, cp.async.bulk + mbarrier)
//   2. bucket [lo,hi) of the window's first K bases from the K-mer table
//   3. a bucket of one suffix is decided at once (match length, left-maximality base); larger
//      buckets go on the warp's list and are worked off densely: lower bound of Q[i..] among
//      the bucket's suffixes (binary search on packed text, 32 bases per probe), longest match
//      = better of the two neighbours of the insertion point, unique iff the other neighbour
//      is shorter and the LCP entry on the far side is shorter too, left-maximality
//   4. once a position is known to match R[r..] for m bases, the positions behind it continue
//      that match at r+1, r+2, ... and cannot be left-maximal there; as long as the rest of the
//      match is longer than any repeat of its reference suffix (the index's skip table, one
//      byte load) it is also their unique longest match, so they yield no anchor and are
//      stepped over without any table, suffix-array or text access: 3 of 4 positions at 2 %
//      divergence.  A warp runs its 1024 positions as a work list (bitmaps of positions to look
//      up / settled), so that the look-ups that remain are made 32 at a time
//   5. anchors are written at the slot of their position; k_seed_gather compacts them in order
// Output order equals the oracle's sort order, so no sort follows.
#include <algorithm>
#include <cstring>

#include "pmn_scratch.cuh"

#define SEED_THREADS 128
#define SEED_WARPS (SEED_THREADS / 32)
#define SEED_CHUNK 32                              /* consecutive positions per lane: lane l owns word l of its warp's anchor bitmap */
#define SEED_RUN (32 * SEED_CHUNK)                 /* positions per warp */
#define SEED_TILE (SEED_THREADS * SEED_CHUNK)      /* query positions per block */
#define SEED_WORDS (SEED_TILE / 32 + 8)            /* staged words: tile + alignment slack + one window */
static_assert(SEED_WORDS + 8 <= PMN_PAD_WORDS, "the last tile of a text reads SEED_WORDS words from its first word on: the padding must cover them");
static_assert((SEED_WORDS * 4) % 16 == 0, "TMA bulk copies move multiples of 16 bytes");

struct SeedSection {
    int64_t start;      // offset of the record inside the strand's concatenated text
    int64_t len;        // bases of the record
    int64_t npos;       // positions to try = len - minmatch + 1 (>= 1)
    int64_t tile0;      // first tile of this section
    int32_t tag;        // record*2 + strand
    int32_t strand;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}

// Positions fit 31 bits (pmn_fasta_to_device refuses more than 2^31 bases): the per-position code below does all its
// index arithmetic in 32 bits, which is a third fewer instructions than the int64 helpers of pmn_common.cuh.
struct View32 { const uint64_t *w; const uint32_t *xm; uint32_t n; int has_x; };
__device__ __forceinline__ View32 view32(const PackedView &v) { return View32{ v.w, v.xm, (uint32_t)v.n, v.has_x }; }
__device__ __forceinline__ uint64_t win32(const uint64_t *__restrict__ w, uint32_t p)
{
    const uint32_t k = p >> 5; const int sh = (int)(p & 31u) * 2;
    const uint64_t a = __ldg(w + k), b = __ldg(w + k + 1);
    return sh ? (a << sh) | (b >> (64 - sh)) : a;
}
// matchable bases among the 32 starting at p (stops at the first X or at the end)
__device__ __forceinline__ int valid32(const View32 &s, uint32_t p)
{
    if (p >= s.n) return 0;
    if (s.has_x) { const uint32_t k = p >> 5; const int sh = (int)(p & 31u); const uint32_t x = __funnelshift_l(__ldg(s.xm + k + 1), __ldg(s.xm + k), sh); return x ? __clz((int)x) : 32; }
    const uint32_t r = s.n - p;
    return r < 32u ? (int)r : 32;
}
__device__ __forceinline__ int base32(const View32 &s, uint32_t p)      // 0..3, or 4; p = 0xffffffff (one before the start) is X
{
    if (p >= s.n) return PMN_CODE_X;
    if (s.has_x && ((__ldg(s.xm + (p >> 5)) >> (31 - (int)(p & 31u))) & 1u)) return PMN_CODE_X;
    return (int)((__ldg(s.w + (p >> 5)) >> (62 - 2 * (int)(p & 31u))) & 3ull);
}
// common prefix in matchable bases of a[pa..] and b[pb..]
__device__ __forceinline__ uint32_t lcp32(const View32 &a, uint32_t pa, const View32 &b, uint32_t pb)
{
    uint32_t l = 0;
    for (;;) {
        const int va = valid32(a, pa + l), vb = valid32(b, pb + l);
        const int v = va < vb ? va : vb;
        const uint64_t x = win32(a.w, pa + l) ^ win32(b.w, pb + l);
        int m = x ? (__clzll((long long)x) >> 1) : 32;
        if (m > v) m = v;
        l += (uint32_t)m;
        if (m < 32) return l;
    }
}

// true iff the reference suffix at s sorts before Q[g..] (order of pmn_index.cu; a query X or
// the query end compares greater than every reference symbol)
__device__ __forceinline__ bool ref_lt_query(const View32 &R, uint32_t s, const View32 &Q, uint32_t g, uint64_t qw0, int vq0)
{
    uint32_t off = 0;
    uint64_t qw = qw0; int vq = vq0;
    for (;;) {
        const int vr = valid32(R, s + off);
        const uint64_t rw = win32(R.w, s + off);
        const uint64_t x = rw ^ qw;
        const int m = x ? (__clzll((long long)x) >> 1) : 32;
        const int v = vr < vq ? vr : vq;
        if (m < v) return ((rw >> (62 - 2 * m)) & 3ull) < ((qw >> (62 - 2 * m)) & 3ull);
        if (v == 32) { off += 32; vq = valid32(Q, g + off); qw = win32(Q.w, g + off); continue; }
        if (vr < vq) return s + off + (uint32_t)vr >= R.n;     // reference ran out (END, smallest) or hit an X (greater)
        return true;                                 // the query hit X/END first, or both did: query is greater
    }
}

// The general case of one position: lower bound of Q[g..] in [lo, hi), the better of its two neighbours, uniqueness from
// the LCP array, left-maximality.
__device__ __forceinline__ bool seed_general(const View32 &R32, const View32 &Q32, const uint32_t *__restrict__ sa, const int32_t *__restrict__ lcp,
                                             uint32_t lo, uint32_t hi, uint32_t g, uint64_t qw, int vq, int minmatch, uint32_t *r_out, uint32_t *l_out)
{
    const uint32_t sa_n = R32.n;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (ref_lt_query(R32, __ldg(sa + mid), Q32, g, qw, vq)) lo = mid + 1; else hi = mid;
    }
    const uint32_t p = lo;
    const int64_t L1 = p > 0 ? (int64_t)lcp32(Q32, g, R32, __ldg(sa + p - 1)) : -1;
    const int64_t L2 = p < sa_n ? (int64_t)lcp32(Q32, g, R32, __ldg(sa + p)) : -1;
    const int64_t L = L1 > L2 ? L1 : L2;
    if (L < minmatch || L1 == L2) return false;
    bool unique; uint32_t r;
    if (L2 > L1) { r = __ldg(sa + p); unique = !(p + 1 < sa_n && __ldg(lcp + p + 1) >= L); }
    else { r = __ldg(sa + p - 1); unique = !(__ldg(lcp + p - 1) >= L); }
    if (!unique) return false;
    const int qb = base32(Q32, g - 1), rb = base32(R32, r - 1);
    if (qb == rb && qb < 4) return false;
    *r_out = r; *l_out = (uint32_t)L;
    return true;
}

// A warp works through its run of SEED_RUN positions as a work list, so that every step has 32 lanes doing the same thing:
//   need / done   one bit per position: "has to be looked up" / "settled" (looked up, or stepped over)
//   0  blocks     one bit probe per block of minmatch - P + 1 positions settles the block when the P-mer at its last position is not in
//                 the reference (presence bitmap of the index): most of the strand without homology, and the positions in front of a
//                 mismatch.  The work list starts from the positions of every block that follows a settled one.
//   A  select     the next 32 positions of `need`
//   B  classify   window of the query, bucket of its first K bases.  Empty bucket (3 of 4 look-ups at 2 % divergence): nothing
//                 matches minmatch >= K bases; the next position is needed.  Two or more suffixes: the position goes on the
//                 multi list, the next position is needed.  One suffix: (position, slot) goes into the single buffer.
//   C  singles    32 at a time: the suffix s is the only candidate, every other suffix shares fewer than K bases with the query
//                 and with s, so the match is unique; L = lcp(Q[g..], R[s..]) decides the anchor together with the
//                 left-maximality base, and settles the positions behind it: position g+j continues the match at s+j with
//                 L-j bases, where it is not left-maximal (the base before it is matched); while L-j exceeds the longest repeat
//                 of the suffix at s+j that match is also its unique longest one, i.e. the position has no anchor.  With
//                 E = s+L those are exactly the positions before p_stop = E - skip[E] (pmn_index.cu: the repeat ends
//                 e(p) = p + rep(p) never decrease): they become `done` without any table, suffix-array or text access, and
//                 the positions from p_stop to the one behind the mismatch become `need`.
//   D  multis     the listed positions 32 at a time: binary search and both neighbours (seed_general).
// Every settled position settles or schedules its successor, so the run is covered when `need` and the buffers are empty.
// Anchors are written at the slot of their position (stage is one int4 per position) with a bitmap per warp run; the gather
// kernel compacts them in position order.
struct SeedWarp {
    uint32_t need[32], done[32], bits[32];
    uint32_t s_lo[64];
    uint16_t s_idx[64], sel[32];
    uint16_t mlist[SEED_RUN];
};

__device__ __forceinline__ void seed_or_range(uint32_t *bm, int a, int b)      // bits [a, b) of a SEED_RUN-bit map
{
    if (a >= b) return;
    const int w0 = a >> 5, w1 = (b - 1) >> 5;
    for (int w = w0; w <= w1; w++) {
        uint32_t m = ~0u;
        if (w == w0) m &= ~0u << (a & 31);
        if (w == w1) m &= ~0u >> (31 - ((b - 1) & 31));
        atomicOr(bm + w, m);
    }
}

__global__ void __launch_bounds__(SEED_THREADS) k_seed(PackedView R, const uint32_t *__restrict__ sa, const int32_t *__restrict__ lcp,
                                                      const uint32_t *__restrict__ table, const uint8_t *__restrict__ skip, const uint32_t *__restrict__ present, int P, int K,
                                                      PackedView QF, PackedView QR,
                                                      const SeedSection *__restrict__ secs, int nsec, int minmatch,
                                                      int4 *__restrict__ stage, uint32_t *__restrict__ run_bits, uint32_t *__restrict__ tile_cnt, unsigned tile_base, unsigned ntiles,
                                                      unsigned long long *__restrict__ lookups)
{
    __shared__ __align__(16) uint64_t s_w[SEED_WORDS];
    __shared__ __align__(16) uint32_t s_x[SEED_WORDS];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ SeedWarp s_warp[SEED_WARPS];
    if (threadIdx.x == 0) mbar_init(&s_bar, 1);
    __syncthreads();
    uint32_t phase = 0;
    unsigned my_lookups = 0, my_probes_total = 0;
    // a block walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (one tile per block unless the host caps the grid)
    for (unsigned ltile = blockIdx.x; ltile < ntiles; ltile += gridDim.x, phase ^= 1u) {
    const int64_t tile = (int64_t)ltile + tile_base;          // ltile is local to the launched tile range

    // which section does this tile belong to
    int lo_s = 0, hi_s = nsec - 1;
    while (lo_s < hi_s) { int mid = (lo_s + hi_s + 1) >> 1; if (secs[mid].tile0 <= tile) lo_s = mid; else hi_s = mid - 1; }
    const SeedSection sec = secs[lo_s];
    const PackedView Q = sec.strand ? QR : QF;
    const int64_t off0 = (tile - sec.tile0) * SEED_TILE;    // first position of the tile inside the record
    const int64_t g0 = sec.start + off0;
    const int64_t w0 = (g0 >> 5) & ~3ll;                                   // 16-byte aligned for text and mask

    if (threadIdx.x == 0) {
        uint32_t bytes = SEED_WORDS * 8 + (Q.has_x ? SEED_WORDS * 4 : 0);
        mbar_expect_tx(&s_bar, bytes);
        tma_bulk_g2s(s_w, Q.w + w0, SEED_WORDS * 8, &s_bar);
        if (Q.has_x) tma_bulk_g2s(s_x, Q.xm + w0, SEED_WORDS * 4, &s_bar);
    }
    mbar_wait(&s_bar, phase);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = pmn_lanemask_lt();
    const int first_need = minmatch < 32 ? minmatch : 32;
    const View32 R32 = view32(R), Q32 = view32(Q);
    const bool use_table = minmatch >= K;
    // Warp w owns the SEED_RUN consecutive positions [w * SEED_RUN, (w + 1) * SEED_RUN) of the tile and their slots of the staging area
    const size_t run = (size_t)ltile * SEED_WARPS + warp;
    int4 *wstage = stage + run * SEED_RUN;
    const int64_t woff = off0 + warp * SEED_RUN;
    SeedWarp &W = s_warp[warp];
    // the first window of position woff + idx, from the staged tile
    auto window = [&](int idx, uint32_t &g, uint64_t &qw, int &vq) {
        g = (uint32_t)(sec.start + woff + idx);
        const uint32_t rel = g - (uint32_t)(w0 << 5); const int k = (int)(rel >> 5), sh = (int)(rel & 31);
        const uint64_t a = s_w[k], b = s_w[k + 1];
        qw = sh ? (a << (2 * sh)) | (b >> (64 - 2 * sh)) : a;
        if (Q.has_x) { const uint32_t xw = __funnelshift_l(s_x[k + 1], s_x[k], sh); vq = xw ? __clz((int)xw) : 32; }
        else { const uint32_t r = Q32.n - g; vq = r < 32u ? (int)r : 32; }
    };
    int run_lim;        // positions of the run that exist
    { const int64_t left = sec.npos - woff; run_lim = left <= 0 ? 0 : left < SEED_RUN ? (int)left : SEED_RUN; }
    const uint32_t vmask = run_lim >= (lane + 1) * 32 ? ~0u : run_lim <= lane * 32 ? 0u : (1u << (run_lim - lane * 32)) - 1u;
    W.done[lane] = 0; W.bits[lane] = 0; W.need[lane] = 0;
    uint32_t nsingle = 0, nmulti = 0;                                     // the same in every lane
    __syncwarp();
    // ---- 0. blocks of B = minmatch - P + 1 positions: a match of minmatch bases that starts anywhere in a block contains the P-mer at
    // the block's last position; when the reference does not hold that P-mer (or it is cut short by an X or the end of the
    // record) the whole block is settled.  The work list starts from every position of the blocks that follow a settled block
    // (and of the first block): that is where a mismatch has just been passed and a new match may begin — its look-ups are
    // made side by side instead of one scheduling the next.
    const int Bk = minmatch - P + 1;
    unsigned my_probes = 0;
    if (Bk >= 2 && run_lim > 0) {
        const int nblk = (run_lim + Bk - 1) / Bk;
        bool carry_here = false;                                          // was the last block of the previous round of 32 in the reference?
        for (int base = 0; base < nblk; base += 32) {
            const int blk = base + lane;
            bool here = false; int b0 = 0, b1 = 0;
            if (blk < nblk) {
                b0 = blk * Bk; b1 = b0 + Bk < run_lim ? b0 + Bk : run_lim;
                uint32_t g; uint64_t qw; int vq;
                window(b1 - 1, g, qw, vq);
                if (vq >= P) { const uint32_t pk = (uint32_t)(qw >> (64 - 2 * P)); here = (__ldg(present + (pk >> 5)) >> (pk & 31)) & 1u; }
                if (!here) seed_or_range(W.done, b0, b1);
            }
            const unsigned hb = __ballot_sync(0xffffffffu, here);
            const bool prev_here = lane ? (hb >> (lane - 1)) & 1u : carry_here;
            if (here && !prev_here) seed_or_range(W.need, b0, b1);
            carry_here = (hb >> 31) & 1u;
        }
        if (lane == 0) my_probes = (unsigned)nblk;
    } else W.need[lane] = vmask & 1u;                                     // no filter (minmatch <= P): the seeds are every 32nd position
    my_probes_total += my_probes;
    __syncwarp();

    // 32 buffered one-suffix positions (the last n of the buffer)
    auto singles = [&](uint32_t n) {
        nsingle -= n;
        if ((uint32_t)lane < n) {
            const int idx = W.s_idx[nsingle + lane];
            const uint32_t s = __ldg(sa + W.s_lo[nsingle + lane]);
            uint32_t g; uint64_t qw; int vq;
            window(idx, g, qw, vq);
            const uint32_t L = lcp32(Q32, g, R32, s);
            if (L >= (uint32_t)minmatch) {
                const int qb = base32(Q32, g - 1), rb = base32(R32, s - 1);
                if (!(qb == rb && qb < 4)) { atomicOr(&W.bits[idx >> 5], 1u << (idx & 31)); wstage[idx] = make_int4((int)(s + 1), (int)(woff + idx + 1), (int)L, sec.tag); }
            }
            int step = 1;                                       // positions this look-up settles, itself included
            if (L >= 2u) {
                const uint32_t E = s + L;
                const uint32_t d = __ldg(skip + E);
                if (d < 255u) { const uint32_t p_stop = E - d; if (p_stop > s + 1u) step = (int)(p_stop - s); }
            }
            const int64_t upto = (int64_t)idx + (int64_t)L + 2;                     // one behind the position of the mismatch
            const int c1 = idx + step < run_lim ? idx + step : run_lim;
            const int n1 = upto < run_lim ? (int)upto : run_lim;
            seed_or_range(W.done, idx + 1, c1);
            seed_or_range(W.need, idx + step, n1);
        }
        __syncwarp();
    };

    for (;;) {
        // ---- A. what is waiting
        const uint32_t w = W.need[lane] & ~W.done[lane] & vmask;
        const int pc = __popc(w);
        int pre = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pre, o); if (lane >= o) pre += t; }
        const int total = __shfl_sync(0xffffffffu, pre, 31);
        pre -= pc;
        if (nsingle >= 32u || (nsingle > 0u && total == 0)) { singles(nsingle < 32u ? nsingle : 32u); continue; }
        if (total == 0) break;
        {   // every lane hands the first positions of its word to the selection, in position order
            uint32_t take = 0, ww = w; int r = pre;
            while (ww && r < 32) { const int b = __ffs((int)ww) - 1; ww &= ww - 1u; W.sel[r] = (uint16_t)(lane * 32 + b); take |= 1u << b; r++; }
            W.need[lane] = w & ~take;
            W.done[lane] |= take;
        }
        __syncwarp();
        const int T = total < 32 ? total : 32;
        // ---- B. classify
        const bool act = lane < T;
        const int idx = act ? (int)W.sel[lane] : 0;
        int cls = 0; uint32_t lo = 0;
        if (act) {
            uint32_t g; uint64_t qw; int vq;
            window(idx, g, qw, vq);
            if (vq >= first_need) {
                if (!use_table) cls = 2;
                else {
                    const uint32_t km = (uint32_t)(qw >> (64 - 2 * K));
                    lo = __ldg(table + km);
                    const uint32_t hi = __ldg(table + km + 1);
                    cls = hi - lo == 1u ? 1 : hi > lo ? 2 : 0;
                }
            }
            if (cls != 1 && idx + 1 < run_lim) atomicOr(&W.need[(idx + 1) >> 5], 1u << ((idx + 1) & 31));
        }
        const unsigned b1 = __ballot_sync(0xffffffffu, cls == 1), b2 = __ballot_sync(0xffffffffu, cls == 2);
        if (cls == 1) { const uint32_t at = nsingle + __popc(b1 & lt); W.s_idx[at] = (uint16_t)idx; W.s_lo[at] = lo; }
        if (cls == 2) W.mlist[nmulti + __popc(b2 & lt)] = (uint16_t)idx;
        nsingle += __popc(b1); nmulti += __popc(b2);
        if (lane == 0) my_lookups += (unsigned)T;
        __syncwarp();
    }
    // ---- D. the multi list, densely
    for (uint32_t c = 0; c < nmulti; c += 32) {
        if (c + lane < nmulti) {
            const int idx = W.mlist[c + lane];
            uint32_t g; uint64_t qw; int vq;
            window(idx, g, qw, vq);
            uint32_t lo = 0, hi = R32.n;
            if (use_table) { const uint32_t km = (uint32_t)(qw >> (64 - 2 * K)); lo = __ldg(table + km); hi = __ldg(table + km + 1); }
            uint32_t r, L;
            if (seed_general(R32, Q32, sa, lcp, lo, hi, g, qw, vq, minmatch, &r, &L)) {
                atomicOr(&W.bits[idx >> 5], 1u << (idx & 31)); wstage[idx] = make_int4((int)(r + 1), (int)(woff + idx + 1), (int)L, sec.tag);
            }
        }
    }
    __syncwarp();
    const uint32_t word = W.bits[lane];
    run_bits[run * SEED_CHUNK + lane] = word;
    uint32_t wcount = __popc(word);
#pragma unroll
    for (int o = 16; o; o >>= 1) wcount += __shfl_xor_sync(0xffffffffu, wcount, o);
    if (lane == 0) tile_cnt[run] = wcount;
    __syncthreads();            // the staged tile is free for the next copy
    }
    if (lookups && (threadIdx.x & 31) == 0 && my_lookups) atomicAdd(lookups, (unsigned long long)my_lookups);      // how many positions were looked up (the rest were stepped over)
    if (lookups && (threadIdx.x & 31) == 0 && my_probes_total) atomicAdd(lookups + 1, (unsigned long long)my_probes_total);
}

// gather the per-run anchors into one contiguous, ordered anchor array
__global__ void __launch_bounds__(256) k_seed_gather(const int4 *__restrict__ stage, const uint32_t *__restrict__ run_bits,
                                                    const uint32_t *__restrict__ run_off, int64_t nruns, int4 *__restrict__ anchors)
{
    const int64_t run = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);     // one warp per run of SEED_RUN positions
    if (run >= nruns) return;
    const int lane = threadIdx.x & 31;
    const unsigned lt = pmn_lanemask_lt();
    uint32_t off = run_off[run];
    const uint32_t mine = run_bits[run * SEED_CHUNK + lane];       // word w of the run's bitmap sits in lane w
    for (int w = 0; w < SEED_CHUNK; w++) {
        const uint32_t word = __shfl_sync(0xffffffffu, mine, w);
        if (!word) continue;
        if ((word >> lane) & 1u) anchors[off + __popc(word & lt)] = stage[(size_t)run * SEED_RUN + w * 32 + lane];
        off += __popc(word);
    }
}

int pmn_seed_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *q, const pmn_opts *o, int64_t *n_anchors, int part, int nparts)
{
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    *n_anchors = 0;
    if (o->minmatch < 1) return pmn_set_error(PMN_E_ARG, "minmatch must be positive");
    std::vector<SeedSection> secs;
    int64_t tiles = 0;
    for (int rec = 0; rec < q->nrec; rec++) {
        for (int strand = 0; strand < 2; strand++) {
            if (strand == 0 && !o->do_forward) continue;
            if (strand == 1 && !o->do_reverse) continue;
            int64_t npos = q->len[rec] - o->minmatch + 1;
            if (npos < 1) continue;
            SeedSection s;
            s.start = strand ? q->n - q->off[rec] - q->len[rec] : q->off[rec];
            s.len = q->len[rec]; s.npos = npos; s.tile0 = tiles; s.tag = rec * 2 + strand; s.strand = strand;
            secs.push_back(s);
            tiles += (npos + SEED_TILE - 1) / SEED_TILE;
        }
    }
    if (tiles > 0x7fffffffll) return pmn_set_error(PMN_E_ARG, "seed: query too large");
    // query-range sharding of one large pair: part k of G takes the tiles
    // [k*T/G, (k+1)*T/G); tiles are in (record, strand, position) order, so the parts'
    // anchor lists concatenated in part order are the full, ordered anchor list
    const int64_t all_tiles = tiles;
    const int64_t t_lo = all_tiles * part / nparts, t_hi = all_tiles * (part + 1) / nparts;
    tiles = t_hi - t_lo;
    if (tiles == 0) return 0;
    const int64_t runs = tiles * SEED_WARPS;                   // one anchor run per warp of a tile
    if (S.sections.ensure(sizeof(SeedSection) * secs.size()) || S.stage.ensure(sizeof(int4) * (size_t)tiles * SEED_TILE) ||
        S.tile_cnt.ensure(4 * (size_t)runs + 16) || S.tile_off.ensure(4 * (size_t)runs) || S.seed_bits.ensure(4 * SEED_CHUNK * (size_t)runs) ||
        S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(runs)) || S.ensure_pinned(64) || S.cl_counters.ensure(64)) return -3;
    PMN_H2D(c, S.sections.p, secs.data(), sizeof(SeedSection) * secs.size());
    unsigned long long *lookups = (unsigned long long *)((char *)S.cl_counters.p + 32);      // the clustering stage uses the first 8 bytes, later
    PMN_CUDA_OK(cudaMemsetAsync(lookups, 0, 16, st));
    PMN_CUDA_OK(cudaEventRecord(c->ev[6], st));
    // one block per tile by default (the hardware deals tiles to SMs as they free up); PMN_SEED_BPS = k caps the grid at k
    // blocks per SM that stride over the tiles (experiment: leave thread slots to the kernels of other pairs)
    static const int seed_bps = getenv("PMN_SEED_BPS") ? atoi(getenv("PMN_SEED_BPS")) : 0;
    const unsigned seed_grid = seed_bps > 0 ? (unsigned)std::min<int64_t>(tiles, (int64_t)c->sm_count * seed_bps) : (unsigned)tiles;
    k_seed<<<seed_grid, SEED_THREADS, 0, st>>>(ix->seq->fwd(), ix->sa(), ix->lcp(), ix->table(), ix->skip(), ix->present(), ix->P, ix->K,
                                                     q->fwd(), q->rev(), S.sections.as<SeedSection>(), (int)secs.size(), o->minmatch,
                                                     S.stage.as<int4>(), S.seed_bits.as<uint32_t>(), S.tile_cnt.as<uint32_t>(), (unsigned)t_lo, (unsigned)tiles, lookups);
    PMN_CUDA_OK(cudaEventRecord(c->ev[7], st));
    pmn_scan<uint32_t, OpAddU32, false>(S.tile_cnt.as<uint32_t>(), S.tile_off.as<uint32_t>(), runs, S.scan_tmp.as<uint32_t>(), st);
    uint32_t *tail = (uint32_t *)S.pinned;
    PMN_D2H(c, tail, S.tile_off.as<uint32_t>() + (runs - 1), 4);
    PMN_D2H(c, tail + 1, S.tile_cnt.as<uint32_t>() + (runs - 1), 4);
    PMN_D2H(c, tail + 2, lookups, 16);
    PMN_CUDA_OK(cudaStreamSynchronize(st));   // the host sizes the clustering stage from the anchor count
    c->syncs++;
    int64_t total = (int64_t)tail[0] + tail[1];
    { unsigned long long lk[2]; memcpy(lk, tail + 2, 16); S.seed_lookups = (int64_t)lk[0]; S.seed_probes = (int64_t)lk[1]; }
    c->launches += 4;
    if (total > 0) {
        if (S.anchors.ensure(sizeof(int4) * (size_t)total)) return -3;
        k_seed_gather<<<(unsigned)((runs + 7) / 8), 256, 0, st>>>(S.stage.as<int4>(), S.seed_bits.as<uint32_t>(), S.tile_off.as<uint32_t>(), runs, S.anchors.as<int4>());
        c->launches += 1;
    }
    PMN_CUDA_OK(cudaGetLastError());
    *n_anchors = total;
    return 0;
}

// pmn_cluster.cu — mgaps-style diagonal clustering of the anchors of one pair.
//
// Stands in for `mgaps -l 65 -s 90 -d 5 -f .12` inside the `nucmer` child process of
// /root/reference/lib/nucmer/mugsy_nucmer.ml:100.  Oracle counterpart: oracle/pmn_oracle.c §5
// (filter_matches, mgaps_section, process_matches) — the cluster list must be identical,
// including order, trimming and which chains are dropped.
//
// Stages (all on the device, no sort of the anchors is needed — they arrive in (tag, query
// position) order from pmn_seed.cu):
//   1. Filter_Matches: a segmented prefix-max of the match ends splits each section into
//      independent overlap groups; one thread walks each group sequentially (the original
//      is sequential by construction).  Filtered anchors stay in place as dead entries: they
//      unite with nothing, sort as components of their own and emit nothing, so the host does
//      not have to wait for their count
//   2. union-find over the separation/diagonal window (lock-free hooking of the larger root
//      under the smaller, so a component's label is its smallest member: deterministic)
//   3. stable radix sort by label = mgaps' qsort by (cluster id, start2, start1)
//   4. per component, repeated best-chain extraction (one warp per component; the O(m^2)
//      scan of the original is pruned with a running prefix maximum — same argmax)
//   5. ordered compaction of the emitted clusters: output order = (component, extraction)
#include "pmn_scratch.cuh"

struct OpMaxI64 { __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; } static __device__ __forceinline__ long long identity() { return LLONG_MIN; } };

// ------------------------------------------------------------------------------------ 1. filter

__global__ void __launch_bounds__(256) k_cl_endkeys(const int4 *__restrict__ anc, int64_t n, long long *__restrict__ key)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 a = anc[i];
    key[i] = ((long long)a.w << 32) | (unsigned)(a.y + a.z);      // (tag, query end): max-scan = per-section running max
}

__global__ void __launch_bounds__(256) k_cl_groupflags(const int4 *__restrict__ anc, const long long *__restrict__ pmax, int64_t n, uint8_t *__restrict__ gstart)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool st = true;
    if (i > 0) {
        long long pm = pmax[i - 1];
        int4 a = anc[i];
        st = (int)(pm >> 32) != a.w || a.y > (int)(unsigned)(pm & 0xffffffffll);
    }
    gstart[i] = st ? 1 : 0;
}

// one thread per overlap group: the literal Filter_Matches loops, confined to the group
__global__ void __launch_bounds__(128) k_cl_filter(int4 *__restrict__ anc, int64_t n, const uint8_t *__restrict__ gstart,
                                                  uint8_t *__restrict__ tent, uint32_t *__restrict__ good)
{
    int64_t a0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a0 >= n || !gstart[a0]) return;
    int64_t b0 = a0 + 1;
    while (b0 < n && !gstart[b0]) b0++;
    if (b0 - a0 == 1) { good[a0] = 1; return; }
    for (int64_t i = a0; i < b0; i++) { good[i] = 1; tent[i] = 0; }
    for (int64_t i = a0; i < b0 - 1; i++) {
        if (!good[i]) continue;
        int4 A = anc[i];
        int i_diag = A.y - A.x, i_end = A.y + A.z;
        for (int64_t j = i + 1; j < b0; j++) {
            int4 B = anc[j];
            if (B.y > i_end) break;
            if (!good[j]) continue;
            int j_diag = B.y - B.x;
            if (i_diag == j_diag) {
                int j_extent = B.z + B.y - A.y;
                if (j_extent > A.z) { A.z = j_extent; anc[i].z = j_extent; i_end = A.y + j_extent; }
                good[j] = 0;
            } else if (A.x == B.x || A.y == B.y) {
                int olap = A.x == B.x ? A.y + A.z - B.y : A.x + A.z - B.x;
                if (A.z < B.z) { if (olap >= A.z / 2) { good[i] = 0; break; } }
                else if (B.z < A.z) { if (olap >= B.z / 2) good[j] = 0; }
                else if (olap >= A.z / 2) { tent[j] = 1; if (tent[i]) { good[i] = 0; break; } }
            }
        }
    }
}

// ------------------------------------------------------------------------------------ 2. union-find

__device__ __forceinline__ uint32_t uf_find(uint32_t *parent, uint32_t x)
{
    uint32_t p = parent[x];
    while (p != x) { uint32_t gp = parent[p]; if (gp != p) parent[x] = gp; x = p; p = gp; }   // path halving
    return x;
}
__device__ __forceinline__ void uf_unite(uint32_t *parent, uint32_t a, uint32_t b)
{
    for (;;) {
        a = uf_find(parent, a); b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { uint32_t t = a; a = b; b = t; }           // hook the larger root under the smaller
        uint32_t old = atomicCAS(parent + a, a, b);
        if (old == a) return;
    }
}

__global__ void __launch_bounds__(256) k_cl_init_parent(uint32_t *parent, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = (uint32_t)i;
}

// lim[sep] = max(diagdiff, (long)(diagfactor * sep)) for sep = 0..maxgap, computed on the host in
// double precision exactly like the oracle; a negative separation uses diagdiff
__global__ void __launch_bounds__(256) k_cl_union(const int4 *__restrict__ f, const uint32_t *__restrict__ good, int64_t n, int maxgap, int diagdiff,
                                                 const int32_t *__restrict__ lim, uint32_t *parent)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !good[i]) return;
    int4 A = f[i];
    int i_end = A.y + A.z, i_diag = A.y - A.x;
    for (int64_t j = i + 1; j < n; j++) {
        int4 B = f[j];
        if (B.w != A.w) break;
        int sep = B.y - i_end;
        if (sep > maxgap) break;
        if (!good[j]) continue;                                  // filtered out: as if it were not in the list
        int dd = (B.y - B.x) - i_diag; if (dd < 0) dd = -dd;
        int l = sep >= 0 ? lim[sep] : diagdiff;
        if (dd <= l) uf_unite(parent, (uint32_t)i, (uint32_t)j);
    }
}

__global__ void __launch_bounds__(256) k_cl_labels(uint32_t *parent, int64_t n, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = uf_find(parent, (uint32_t)i);
    vals[i] = (uint32_t)i;
}

// ------------------------------------------------------------------------------------ 3. gather in (label, start2, start1) order

__global__ void __launch_bounds__(256) k_cl_gather(const int4 *__restrict__ f, const uint32_t *__restrict__ good, const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals, int64_t n,
                                                  int32_t *__restrict__ s1, int32_t *__restrict__ s2, int32_t *__restrict__ ln, int32_t *__restrict__ tg,
                                                  uint32_t *__restrict__ cflag, uint8_t *__restrict__ alive)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t src = svals[i];
    int4 a = f[src];
    alive[i] = good[src] ? 1 : 0;                  // a filtered anchor is a component of its own (it was never united) that emits nothing
    s1[i] = a.x; s2[i] = a.y; ln[i] = a.z; tg[i] = a.w;
    cflag[i] = (i == 0 || skeys[i] != skeys[i - 1]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_cl_compstarts(const uint32_t *__restrict__ cflag, const uint32_t *__restrict__ cpos, int64_t n, uint32_t *__restrict__ cstart,
                                                      uint32_t *__restrict__ counters /* [0]=ncomp [1]=next */)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (cflag[i]) cstart[cpos[i]] = (uint32_t)i;
    if (i == n - 1) { counters[0] = cpos[i] + cflag[i]; counters[1] = 0; }
}

// ------------------------------------------------------------------------------------ 4. chain extraction

struct ChainArrays {
    int32_t *s1, *s2, *ln; const int32_t *tg;
    const uint8_t *alive;   // 0: the slot holds an anchor Filter_Matches dropped
    int32_t *score, *from, *adj, *pm; uint8_t *good;
    int32_t *om;            // emitted matches, 3 ints per slot
    uint32_t *om_valid;     // slot holds a match
    int32_t *oc;            // emitted clusters, 3 ints per slot: first match slot, count, tag
    uint32_t *oc_valid;
};

// Process_Matches for the component occupying sorted slots [a, a+m).  One warp.
//
// The DP  score[i] = len[i] + max(0, max_{j<i}(score[j] - pen(i,j)))  (lowest j among the best) looks like a chain of m
// dependent steps, and is one when lane 0 walks it (5 cycles per instruction with nothing else to issue: 600 cycles per
// anchor, a component of 1000 anchors keeps its warp for 0.3 ms).  But in a component of colinear anchors almost every
// anchor's best predecessor is the anchor before it, and under THAT assumption the scores are a prefix scan:
//     c[i] = max(len[i], c[i-1] + len[i] - pen(i, i-1))        affine maps x -> max(A, x + B) compose associatively
// So: (A) all lanes compute the guess c[] and its prefix maxima with warp scans; (B) every lane evaluates the TRUE step
// of the DP for its own i — the same look-back loop, same pruning, same tie rule — reading the guessed scores of the
// anchors before it, and checks that the step reproduces c[i].  If it does for every i, then by induction on i the guess IS
// the score array of the sequential DP (score[i] is a function of score[0..i-1] alone), and the `from` / `adj` the lanes
// found are the sequential ones.  If any lane disagrees, lane 0 runs the sequential DP over the component (dp_sequential:
// the look-back served from a shared-memory ring).  Both paths run the one dp_step below.
#define CH_RING 64
struct ChainShared { int s1[CH_RING], s2[CH_RING], ln[CH_RING], sc[CH_RING], pm[CH_RING]; int from[32], adj[32]; };
#define CH_NEG (INT32_MIN / 4)

// one step of the DP for anchor i = (i1, i2, il); AT(j, &j1, &j2, &jl, &jsc) and PM(j) fetch anchor j < i, its score and the
// prefix maximum of the scores up to j
template <class At, class Pm>
__device__ __forceinline__ void dp_step(int i, int i1, int i2, int il, At AT, Pm PM, int &best, int &from, int &adj)
{
    best = il; from = -1; adj = 0;
    const int idiag = i2 - i1;
    for (int j = i - 1; j >= 0; j--) {
        const int jpm = PM(j);
        if (jpm + il < best || (from == -1 && jpm + il <= best)) break;
        int j1, j2, jl, jsc;
        AT(j, j1, j2, jl, jsc);
        int ol1 = j1 + jl - i1, ol = ol1 > 0 ? ol1 : 0, ol2 = j2 + jl - i2;
        if (ol2 > ol) ol = ol2;
        int dd = idiag - (j2 - j1); if (dd < 0) dd = -dd;
        int v = jsc + il - (ol + dd);
        if (v > best || (v == best && from != -1)) { best = v; from = j; adj = ol; }
    }
}

// the sequential DP (lane 0), anchors and results 32 at a time through shared memory; returns the index of the best score
__device__ int dp_sequential(const ChainArrays &C, int64_t a, int m, ChainShared &W)
{
    const int lane = threadIdx.x & 31;
    int bestIdx = 0, bestScore = INT32_MIN;                    // lane 0 only
    for (int base = 0; base < m; base += 32) {
        const int idx = base + lane;
        if (idx < m) { const int sl = idx & (CH_RING - 1); W.s1[sl] = C.s1[a + idx]; W.s2[sl] = C.s2[a + idx]; W.ln[sl] = C.ln[a + idx]; }
        __syncwarp();
        const int cnt = m - base < 32 ? m - base : 32;
        if (lane == 0) {
            const int lo_ring = base - (CH_RING - 32);         // entries [lo_ring, base + 32) are in the ring
            for (int t = 0; t < cnt; t++) {
                const int i = base + t, sl = i & (CH_RING - 1);
                int best, from, adj;
                dp_step(i, W.s1[sl], W.s2[sl], W.ln[sl],
                        [&](int j, int &j1, int &j2, int &jl, int &jsc) {
                            const int js = j & (CH_RING - 1);
                            if (j >= lo_ring) { j1 = W.s1[js]; j2 = W.s2[js]; jl = W.ln[js]; jsc = W.sc[js]; }
                            else { j1 = C.s1[a + j]; j2 = C.s2[a + j]; jl = C.ln[a + j]; jsc = C.score[a + j]; }
                        },
                        [&](int j) { return j >= lo_ring ? W.pm[j & (CH_RING - 1)] : C.pm[a + j]; }, best, from, adj);
                const int ppm = i > 0 ? W.pm[(i - 1) & (CH_RING - 1)] : 0;
                const int pmv = i == 0 || best > ppm ? best : ppm;
                W.sc[sl] = best; W.pm[sl] = pmv; W.from[t] = from; W.adj[t] = adj;
                if (best > bestScore) { bestScore = best; bestIdx = i; }
            }
        }
        __syncwarp();
        if (idx < m) { const int sl = idx & (CH_RING - 1); C.score[a + idx] = W.sc[sl]; C.from[a + idx] = W.from[lane]; C.adj[a + idx] = W.adj[lane]; C.pm[a + idx] = W.pm[sl]; }
        __syncwarp();
    }
    return __shfl_sync(0xffffffffu, bestIdx, 0);
}

// guess by scan, verify by the true step; returns the index of the best score, or -1 when the guess was wrong somewhere
__device__ int dp_parallel(const ChainArrays &C, int64_t a, int m)
{
    const int lane = threadIdx.x & 31;
    // ---- (A) c[i] = max(len[i], c[i-1] + len[i] - pen(i, i-1)) and its prefix maxima
    int c_carry = CH_NEG, pm_carry = CH_NEG;                   // c and prefix maximum of the last anchor of the previous chunk
    int p1 = 0, p2 = 0, pl = 0;                                // that anchor
    for (int base = 0; base < m; base += 32) {
        const int idx = base + lane;
        int i1 = 0, i2 = 0, il = 0;
        if (idx < m) { i1 = C.s1[a + idx]; i2 = C.s2[a + idx]; il = C.ln[a + idx]; }
        int j1 = __shfl_up_sync(0xffffffffu, i1, 1), j2 = __shfl_up_sync(0xffffffffu, i2, 1), jl = __shfl_up_sync(0xffffffffu, il, 1);
        if (lane == 0) { j1 = p1; j2 = p2; jl = pl; }
        int A = il, B = CH_NEG;                                // the map of this anchor: x -> max(A, x + B)
        if (idx > 0 && idx < m) {
            int ol1 = j1 + jl - i1, ol = ol1 > 0 ? ol1 : 0, ol2 = j2 + jl - i2;
            if (ol2 > ol) ol = ol2;
            int dd = (i2 - i1) - (j2 - j1); if (dd < 0) dd = -dd;
            B = il - (ol + dd);
        }
        if (idx >= m) { A = CH_NEG; B = 0; }                   // identity behind the end
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {                     // (A, B) of lane l becomes the composition of the maps of lanes l-o+1 .. l after those before
            const int Ap = __shfl_up_sync(0xffffffffu, A, o), Bp = __shfl_up_sync(0xffffffffu, B, o);
            if (lane >= o) { const int t = Ap + B; A = A > t ? A : t; B = Bp + B < CH_NEG ? CH_NEG : Bp + B; }
        }
        int cval = c_carry + B; if (cval < A) cval = A;        // c_carry + B stays far from overflow: both are >= CH_NEG
        if (c_carry == CH_NEG) cval = A;
        int pmv = cval;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, pmv, o); if (lane >= o && t > pmv) pmv = t; }
        if (pm_carry > pmv) pmv = pm_carry;
        if (idx < m) { C.score[a + idx] = cval; C.pm[a + idx] = pmv; }
        const int last = (m - base < 32 ? m - base : 32) - 1;
        c_carry = __shfl_sync(0xffffffffu, cval, last); pm_carry = __shfl_sync(0xffffffffu, pmv, last);
        p1 = __shfl_sync(0xffffffffu, i1, last); p2 = __shfl_sync(0xffffffffu, i2, last); pl = __shfl_sync(0xffffffffu, il, last);
    }
    __syncwarp();
    // ---- (B) the true step of every anchor against the guessed scores before it
    int bestIdx = 0, bestScore = INT32_MIN; bool wrong = false;
    for (int base = 0; base < m; base += 32) {
        const int idx = base + lane;
        int best = INT32_MIN, from = -1, adj = 0;
        if (idx < m) {
            dp_step(idx, C.s1[a + idx], C.s2[a + idx], C.ln[a + idx],
                    [&](int j, int &j1, int &j2, int &jl, int &jsc) { j1 = C.s1[a + j]; j2 = C.s2[a + j]; jl = C.ln[a + j]; jsc = C.score[a + j]; },
                    [&](int j) { return C.pm[a + j]; }, best, from, adj);
            if (best != C.score[a + idx]) wrong = true;
            C.from[a + idx] = from; C.adj[a + idx] = adj;
        }
        // first index of the largest score so far (strictly greater replaces, like the sequential scan)
        int v = best, vi = idx;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int u = __shfl_xor_sync(0xffffffffu, v, o), ui = __shfl_xor_sync(0xffffffffu, vi, o);
            if (u > v || (u == v && ui < vi)) { v = u; vi = ui; }
        }
        if (v > bestScore) { bestScore = v; bestIdx = vi; }
    }
    if (__any_sync(0xffffffffu, wrong)) return -1;
    return bestIdx;
}

__device__ void chain_component(const ChainArrays &C, int64_t a, int m, int mincluster, ChainShared &W)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = pmn_lanemask_lt();
    const int tag = C.tg[a];
    int cm = 0, ck = 0;       // matches / clusters emitted so far by this component
    while (m > 0) {
        // ---- DP
        int cur = dp_parallel(C, a, m);
        __syncwarp();
        if (cur < 0) { cur = dp_sequential(C, a, m, W); __syncwarp(); }
        // ---- mark the best chain, sum its lengths (from[i] < i: the walk only moves down, a chunk of 32 links at a time)
        int total = 0, root = 0;
        while (cur >= 0) {
            const int base = cur & ~31, idx = base + lane;
            int f = -1, l = 0;
            if (idx < m) { f = C.from[a + idx]; l = C.ln[a + idx]; }
            unsigned mark;
            // the usual chunk: every link from `cur` down to the chunk's first anchor points to the anchor before it
            const unsigned below = 0xffffffffu >> (31 - (cur - base));                     // lanes base .. cur
            const unsigned plain = __ballot_sync(0xffffffffu, f == idx - 1);
            if ((plain & below) == below) {
                mark = below;
                int lsum = (below >> lane) & 1u ? l : 0;
#pragma unroll
                for (int o = 16; o; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
                total += lsum; root = base; cur = base - 1;
                if (base == 0) cur = -1;
            } else {
                W.from[lane] = f; W.adj[lane] = l;
                __syncwarp();
                mark = 0;
                if (lane == 0) {
                    int c = cur;
                    while (c >= base) { mark |= 1u << (c - base); total += W.adj[c - base]; root = c; c = W.from[c - base]; }
                    cur = c;
                }
                mark = __shfl_sync(0xffffffffu, mark, 0); cur = __shfl_sync(0xffffffffu, cur, 0);
                total = __shfl_sync(0xffffffffu, total, 0); root = __shfl_sync(0xffffffffu, root, 0);
                __syncwarp();
            }
            if ((mark >> lane) & 1u) C.good[a + idx] = 1;
        }
        __syncwarp();
        // ---- emit (chain members in index order, trimmed by their overlap with the predecessor)
        if (total >= mincluster) {
            int nout = 0;
            for (int base = 0; base < m; base += 32) {
                int i = base + lane; bool w = false; int e1 = 0, e2 = 0, el = 0;
                if (i < m && C.good[a + i]) {
                    int adjv = i == root ? 0 : C.adj[a + i];
                    el = C.ln[a + i] - adjv;
                    if (el >= 1) { w = true; e1 = C.s1[a + i] + adjv; e2 = C.s2[a + i] + adjv; }
                }
                unsigned bal = __ballot_sync(0xffffffffu, w);
                if (w) {
                    int64_t slot = a + cm + nout + __popc(bal & lt);
                    C.om[slot * 3] = e1; C.om[slot * 3 + 1] = e2; C.om[slot * 3 + 2] = el; C.om_valid[slot] = 1;
                }
                nout += __popc(bal);
            }
            if (nout > 0) {
                if (lane == 0) { int64_t s = a + ck; C.oc[s * 3] = (int32_t)(a + cm); C.oc[s * 3 + 1] = nout; C.oc[s * 3 + 2] = tag; C.oc_valid[s] = 1; }
                cm += nout; ck++;
            }
        }
        // ---- drop the chain, keep the rest in order
        int k = 0;
        for (int base = 0; base < m; base += 32) {
            int i = base + lane; bool keep = false; int v1 = 0, v2 = 0, vl = 0;
            if (i < m) { keep = !C.good[a + i]; v1 = C.s1[a + i]; v2 = C.s2[a + i]; vl = C.ln[a + i]; }
            unsigned bal = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) { int d = k + __popc(bal & lt); C.s1[a + d] = v1; C.s2[a + d] = v2; C.ln[a + d] = vl; }
            if (i < m) C.good[a + i] = 0;
            k += __popc(bal);
            __syncwarp();
        }
        m = k;
    }
}

__global__ void __launch_bounds__(128) k_cl_chains(ChainArrays C, const uint32_t *__restrict__ cstart, uint32_t *counters, int64_t n, int mincluster)
{
    __shared__ ChainShared W[4];
    const int lane = threadIdx.x & 31;
    const uint32_t ncomp = counters[0];
    for (;;) {
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(counters + 1, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= ncomp) break;
        int64_t a = cstart[c];
        int64_t b = c + 1 < ncomp ? (int64_t)cstart[c + 1] : n;
        if (!C.alive[a]) continue;
        chain_component(C, a, (int)(b - a), mincluster, W[threadIdx.x >> 5]);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------ 5. compaction of the output

__global__ void __launch_bounds__(256) k_cl_out_matches(const int32_t *__restrict__ om, const uint32_t *__restrict__ valid, const uint32_t *__restrict__ pos, int64_t n,
                                                       int32_t *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !valid[i]) return;
    uint32_t p = pos[i];
    out[p * 3] = om[i * 3]; out[p * 3 + 1] = om[i * 3 + 1]; out[p * 3 + 2] = om[i * 3 + 2];
}

// cluster record: (first match index in the compacted match array, count, tag, 0)
__global__ void __launch_bounds__(256) k_cl_out_clusters(const int32_t *__restrict__ oc, const uint32_t *__restrict__ valid, const uint32_t *__restrict__ pos,
                                                        const uint32_t *__restrict__ mpos, int64_t n, int4 *__restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !valid[i]) return;
    out[pos[i]] = make_int4((int)mpos[oc[i * 3]], oc[i * 3 + 1], oc[i * 3 + 2], 0);
}

// ------------------------------------------------------------------------------------ driver
// Leaves in Scratch: cl_matches (int32 x3 per match), cl_recs (int4 per cluster) and their
// counts in n_cl_matches / n_clusters.

int pmn_cluster_impl(pmn_ctx *c, const pmn_index *, const pmn_seq *, const pmn_opts *o, int64_t n)
{
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    S.n_clusters = S.n_cl_matches = 0;
    if (n <= 0) return 0;
    if (n > 0x7ffffff0ll) return pmn_set_error(PMN_E_ARG, "cluster: too many anchors");
    int launches = 0;
    const unsigned g = (unsigned)((n + 255) / 256);
    int4 *anc = S.anchors.as<int4>();
    if (S.cl_a.ensure(8 * (size_t)n) || S.cl_b.ensure(4 * (size_t)n) || S.cl_c.ensure((size_t)n) || S.cl_d.ensure((size_t)n) ||
        S.cl_e.ensure(4 * (size_t)n) || S.cl_f.ensure(4 * (size_t)n) || S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(n)) || S.ensure_pinned(64)) return -3;
    uint32_t *tail = (uint32_t *)S.pinned;

    // 1. filter (in place: `good` marks the survivors)
    long long *ekey = S.cl_a.as<long long>(); uint32_t *good = S.cl_b.as<uint32_t>(); uint8_t *gstart = S.cl_c.as<uint8_t>(), *tent = S.cl_d.as<uint8_t>();
    k_cl_endkeys<<<g, 256, 0, st>>>(anc, n, ekey);
    pmn_scan<long long, OpMaxI64, true>(ekey, ekey, n, S.scan_tmp.as<long long>(), st);
    k_cl_groupflags<<<g, 256, 0, st>>>(anc, ekey, n, gstart);
    k_cl_filter<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(anc, n, gstart, tent, good);
    launches += 6;

    // 2. union-find.  lim[sep] = max(diagdiff, (long)(diagfactor * sep)) lives in the context and is uploaded when the options change
    if (S.lim_maxgap != o->maxgap || S.lim_diagdiff != o->diagdiff || S.lim_diagfactor != o->diagfactor) {
        S.lim_host.resize((size_t)o->maxgap + 1);
        for (int s = 0; s <= o->maxgap; s++) { long l = (long)(o->diagfactor * (double)s); S.lim_host[(size_t)s] = (int32_t)(l > o->diagdiff ? l : o->diagdiff); }
        if (S.cl_g.ensure(4 * S.lim_host.size())) return -3;
        PMN_H2D(c, S.cl_g.p, S.lim_host.data(), 4 * S.lim_host.size());
        S.lim_maxgap = o->maxgap; S.lim_diagdiff = o->diagdiff; S.lim_diagfactor = o->diagfactor;
    }
    if (S.cl_h.ensure(4 * (size_t)n) || S.k0.ensure(8 * (size_t)n) || S.k1.ensure(8 * (size_t)n) ||
        S.v0.ensure(4 * (size_t)n) || S.v1.ensure(4 * (size_t)n)) return -3;
    uint32_t *parent = S.cl_h.as<uint32_t>();
    k_cl_init_parent<<<g, 256, 0, st>>>(parent, n);
    k_cl_union<<<g, 256, 0, st>>>(anc, good, n, o->maxgap, o->diagdiff, S.cl_g.as<int32_t>(), parent);
    k_cl_labels<<<g, 256, 0, st>>>(parent, n, S.k0.as<uint64_t>(), S.v0.as<uint32_t>());
    launches += 3;

    // 3. sort by label (stable: members stay in (start2, start1) order)
    int nb = 1; while ((1ll << nb) < n) nb++;
    int where = pmn_radix_sort(S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), S.k1.as<uint64_t>(), S.v1.as<uint32_t>(), n, nb, S.rs, st, &launches);
    if (where < 0) return -3;
    const uint64_t *skeys = where ? S.k1.as<uint64_t>() : S.k0.as<uint64_t>();
    const uint32_t *svals = where ? S.v1.as<uint32_t>() : S.v0.as<uint32_t>();

    // 4. chains
    if (S.cl_i.ensure(4 * 4 * (size_t)n) || S.cl_j.ensure(4 * 4 * (size_t)n) || S.cl_k.ensure(2 * (size_t)n + 32) ||
        S.cl_l.ensure(4 * 8 * (size_t)n) || S.cl_counters.ensure(64)) return -3;
    int32_t *blk = S.cl_i.as<int32_t>();          // s1, s2, ln, tg
    int32_t *blk2 = S.cl_j.as<int32_t>();         // score, from, adj, pm
    int32_t *outb = S.cl_l.as<int32_t>();         // om (3n), oc (3n), om_valid (n), oc_valid (n)
    ChainArrays C;
    C.s1 = blk; C.s2 = blk + n; C.ln = blk + 2 * n; C.tg = blk + 3 * n;
    C.score = blk2; C.from = blk2 + n; C.adj = blk2 + 2 * n; C.pm = blk2 + 3 * n;
    C.good = S.cl_k.as<uint8_t>();
    uint8_t *alive = C.good + ((size_t)n + 15) / 16 * 16; C.alive = alive;
    C.om = outb; C.oc = outb + 3 * n; C.om_valid = (uint32_t *)(outb + 6 * n); C.oc_valid = (uint32_t *)(outb + 7 * n);
    uint32_t *cflag = S.cl_f.as<uint32_t>(), *cpos = S.cl_e.as<uint32_t>(), *cstart = S.cl_a.as<uint32_t>();      // `good` (cl_b) is read by the gather
    uint32_t *counters = S.cl_counters.as<uint32_t>();
    PMN_CUDA_OK(cudaMemsetAsync(C.good, 0, (size_t)n, st));
    PMN_CUDA_OK(cudaMemsetAsync(C.om_valid, 0, 8 * (size_t)n, st));     // om_valid and oc_valid are adjacent
    k_cl_gather<<<g, 256, 0, st>>>(anc, good, skeys, svals, n, C.s1, C.s2, C.ln, (int32_t *)C.tg, cflag, alive);
    pmn_scan<uint32_t, OpAddU32, false>(cflag, cpos, n, S.scan_tmp.as<uint32_t>(), st);
    k_cl_compstarts<<<g, 256, 0, st>>>(cflag, cpos, n, cstart, counters);
    int blocks = c->sm_count * 4;
    { int64_t need = (n + 3) / 4; if (need < blocks) blocks = (int)need; if (blocks < 1) blocks = 1; }
    k_cl_chains<<<blocks, 128, 0, st>>>(C, cstart, counters, n, o->mincluster);
    launches += 6;

    // 5. compact matches and clusters, in (component, extraction) order
    uint32_t *mpos = cflag, *kpos = cpos;      // reuse
    pmn_scan<uint32_t, OpAddU32, false>(C.om_valid, mpos, n, S.scan_tmp.as<uint32_t>(), st);
    pmn_scan<uint32_t, OpAddU32, false>(C.oc_valid, kpos, n, S.scan_tmp.as<uint32_t>(), st);
    PMN_D2H(c, tail, mpos + (n - 1), 4);
    PMN_D2H(c, tail + 1, C.om_valid + (n - 1), 4);
    PMN_D2H(c, tail + 2, kpos + (n - 1), 4);
    PMN_D2H(c, tail + 3, C.oc_valid + (n - 1), 4);
    PMN_CUDA_OK(cudaStreamSynchronize(st));      // the one host round trip of the stage: the extension sizes its grids from these counts
    c->syncs++;
    const int64_t nm = (int64_t)tail[0] + tail[1], nc = (int64_t)tail[2] + tail[3];
    launches += 6;
    if (nc > 0) {
        if (S.cl_matches.ensure(12 * (size_t)nm) || S.cl_recs.ensure(16 * (size_t)nc)) return -3;
        k_cl_out_matches<<<g, 256, 0, st>>>(C.om, C.om_valid, mpos, n, S.cl_matches.as<int32_t>());
        k_cl_out_clusters<<<g, 256, 0, st>>>(C.oc, C.oc_valid, kpos, mpos, n, S.cl_recs.as<int4>());
        launches += 2;
    }
    PMN_CUDA_OK(cudaGetLastError());
    c->launches += launches;
    S.n_clusters = nc; S.n_cl_matches = nm;
    return 0;
}

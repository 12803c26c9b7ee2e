// pmn_extend.cu — banded affine-gap extension between and beyond the anchors of each cluster,
// alignment stitching, error counts.
//
// Stands in for `postnuc -b 200` (+ sw_align) inside the `nucmer` child process of
// /root/reference/lib/nucmer/mugsy_nucmer.ml:100.  Oracle counterpart: oracle/pmn_oracle.c §6
// (align_engine) and §7 (extend_clusters, extend_forward, extend_backward, parse_delta) —
// alignment rows and delta lists must be identical.
//
// Organisation:
//   E1  clusters are split per reference record, grouped into syntenies (query record x
//       reference record) and sorted by (synteny, first reference start, mgaps order)
//   E2  WAVE 1, fully parallel: every forward alignment of postnuc is a pure function of the
//       cluster list (match -> next match inside a cluster; last match -> target cluster),
//       so all of them run at once: small match -> next match windows one THREAD per job
//       (k_ex_wave1_tpj, full matrix in register strips), cluster ends and large windows one
//       warp per job (k_ex_wave1_big: the banded engine; windows up to 100 x 100 that are too
//       large for a thread as a systolic full matrix, eng_mid_full), persistent warps pulling
//       jobs.  The targets of the cluster-end jobs come from k_ex_targets (one warp per
//       cluster).  The backward extension of every cluster that no end job reaches runs in the
//       same kernel: first for the clusters nobody aims at, and by the warp whose end job is the
//       last to fail for the others (aimers / aimfail).
//   E3  STITCH, one warp per synteny: the sequential control flow of extendClusters
//       (merging, shadow test, backward extension) consumes the wave-1 results — runs of
//       clusters that hand over to their successor 32 at a time, one node per cluster — and
//       the few alignments that depend on earlier outcomes (backward searches the cap cut
//       short, forced merges) run inline on the same warp with the same engine
//   E4  flatten the delta lists, count errors (parseDelta), copy to the host
//
// The engine is an anti-diagonal DP: the cells of one anti-diagonal are striped over the 32
// lanes, the two previous anti-diagonals live in a shared-memory ring (global memory when the
// band outgrows it), high score / band trimming are warp reductions and ballots.  Integer
// ALUs only.
#include <algorithm>
#include <vector>

#include "pmn_scratch.cuh"

#define EX_BAND_SMEM 224                /* widest band kept in shared memory (ring and base caches)  */
#define EX_WCAP (PMN_MAX_ALIGNMENT_LENGTH + 8)   /* global score row capacity                    */
#define EX_ROWS 12                     /* 3 anti-diagonals x (DEL, INS, MAT, cell max)            */
#define EX_DMAX (2 * PMN_MAX_ALIGNMENT_LENGTH + 8)
#define EX_TBW (1u << 20)              /* private traceback bytes per warp slot                   */
#define EX_TB_HDR ((size_t)EX_DMAX * 16)   /* tboff (8) + tblo (4) + reversed deltas (4) per diagonal */
#define EX_ARENA_CHUNK (256u << 10)
#define EX_WARPS_PER_BLOCK 4

#define EX_ERR_POOL 1
#define EX_ERR_ARENA 2
#define EX_ERR_LOGIC 4
#define EX_ERR_NODES 8

struct ExCluster { int32_t mfirst, nm, dir, syn, order, pad; };
struct ExSynteny {
    int32_t cfirst, nC, qrec, rrec;
    int64_t Abase, lenA, BbaseF, BbaseR, lenB;     // *base: 0-based concat index of the record's first base
    int32_t alfirst, alcap, nodefirst, nodecap;
};
struct ExJob { int32_t endA, endB, dcnt, target; uint32_t doff; int32_t reached, valid, asum; };   // asum: sum of (d>0 ? d : |d|-1) over the job's deltas
// everything wave 1 needs to know about one forward alignment, written by k_ex_jobdesc so that a warp
// fetches a job with one coalesced load instead of chasing match -> cluster -> synteny
struct ExJobDesc { int64_t Abase, Bbase; int32_t eA, eB, tA, tB, g, dir, m_o, target; };   // m_o < 0: no job
struct __align__(16) ExAlign { int32_t dirB, sA, sB, eA, eB, P, head, tail, ndelta, live, pad0, pad1; };   // P: reference position consumed through the last indel (sA-1 when none)
// a piece of an alignment's delta list: type 0 = `cnt` pool entries at `a` whose first value gets +-`b` added;
// type 1 = the deltas of the wave-1 jobs [a, cnt) (all reached their targets), `b` = P before the range
struct ExNode { int32_t type; uint32_t a; int32_t cnt, b, outoff, next, alslot, pad; };   // type 0: pool[a, a+cnt), first delta + b; 1: jobs [a, cnt), P before = b; 2: whole cluster `next` (1 + its end job)

// everything the stitcher needs to know about a cluster in one 80-byte record
struct __align__(16) ExCSum {
    int32_t sA0, sB0, len0;            // first match
    int32_t sAl, sBl, eAl, eBl;        // last match: start and end
    int32_t mfirst, nm, dir, anyfail;
    int32_t endA, endB, e_reached, e_dcnt, e_asum, target; uint32_t e_doff;    // the cluster-end job of wave 1
    int32_t bulk_cnt, bulk_P;          // deltas / last P of the inner jobs [mfirst, last) taken together
};

// the backward extension of a cluster's first match, computed ahead of the stitcher (wave 2): search from the match
// start towards the sequence starts, then the forced alignment over the region found
struct __align__(16) ExBack { int32_t fA, fB, ext_i, ext_j; uint32_t doff; int32_t dcnt, asum, valid; };

struct ExShared {                       // everything the device code needs, passed by value
    PackedView R, QF, QR;
    const int32_t *mA, *mB, *mL;        // matches (local reference coordinate)
    const ExCluster *cl; const ExSynteny *syn;
    int nC, nS; int64_t nM;
    ExJob *jobs; const int32_t *mcl;    // per match: result of its forward job, its cluster
    ExAlign *al; ExNode *nodes;
    int32_t *pool; uint32_t pool_cap;   // delta pool
    uint8_t *arena; unsigned long long arena_cap;
    int32_t *gscore; uint8_t *tbpriv;   // per warp slot: EX_ROWS*EX_WCAP ints, EX_TBW bytes
    unsigned long long *counters;       // [0] pool cursor [1] arena cursor [2] cells [3] engine calls [4] error flags [5] job cursor A [6] job cursor B
    int breaklen, do_extend, do_simplify;
    int tpj_cells;                      // largest N x M window the thread-per-job kernel takes (larger ones: one warp each)
    int32_t *syn_nal;                   // alignments produced per synteny (then: nodes used per synteny at [nS + s])
    const struct ExCSum *cs;            // per cluster summary for the stitcher
    const uint32_t *dcnt_ex;            // exclusive prefix of jobs[].dcnt, nM + 1 entries
    const long long *lastP;             // per job g: (piece << 32 | j+1), j = latest job <= g of the same cluster that has deltas
    const uint8_t *anyfail;             // per cluster: some match -> next match job did not reach its target
    unsigned long long *markkey;        // per job: set at the first job of a claimed range
    const ExJobDesc *descA, *descB;     // wave-1 jobs: cluster ends (nC), match -> next match (nM)
    int32_t *overflow, *overflow2;      // inner jobs handed to the big kernel by k_ex_jobdesc (count = counters[7]) / by k_ex_wave1_tpj (counters[12])
    uint2 *tkey;                        // per match: (bin, rank in bin) of its thread-per-job alignment, bin = ~0u when it has none
    uint32_t *tbin;                     // TPJ_BINS counters, then TPJ_BINS + 1 bin starts
    int32_t *tsorted;                   // thread-per-job alignments in bin order (count = counters[10], warp cursor = counters[11])
    uint8_t *tscratch;                  // TPJ_SLOT_BYTES per resident warp of k_ex_wave1_tpj
    ExBack *back;                       // per cluster (wave 2), valid = 0 where none was computed
    int32_t *aimers, *aimfail;          // per cluster: end jobs that aim at it (k_ex_jobdesc) / that did not reach it (k_ex_wave1_big)
    const int4 *tgt;                    // per cluster: (target A, target B, target cluster) of its end job (k_ex_targets)
    int4 *dbg; unsigned dbg_cap;        // PMN_JOBLOG: two int4 per engine call (cursor = counters[15])
};

// ------------------------------------------------------------------------------------ engine

struct Eng {
    const ExShared *X;
    int32_t *ssc;          // shared-memory ring of this warp: EX_ROWS * EX_RW
    int32_t *gsc;          // global rows of this warp
    uint8_t *tbp;          // private traceback region
    int lane;
    int kid;               // 1 = wave 1, 2 = stitcher (job log only)
};

__device__ __forceinline__ void score_edit(int del, int ins, int mat, int &val, int &used)
{
    if (del > ins) { if (del > mat) { val = del; used = PMN_ST_DEL; } else { val = mat; used = PMN_ST_MAT; } }
    else if (ins > mat) { val = ins; used = PMN_ST_INS; }
    else { val = mat; used = PMN_ST_MAT; }
}
__device__ __forceinline__ int max_state(int vD, int vI, int vM)
{
    if (vD > vI) return vD > vM ? PMN_ST_DEL : PMN_ST_MAT;
    return vI > vM ? PMN_ST_INS : PMN_ST_MAT;
}

// ------------------------------------------------------------------------------------ the DP engine
//
// Scores are kept multiplied by 4 inside the register path.  The two free low bits carry the
// state a candidate comes from (DEL 0 < INS 1 < MAT 2), so that one integer maximum yields both
// the best score and scoreEdit's tie rule (MAT over INS over DEL); each three-way maximum is one
// IADD + two VIADDMNMX.
//
// Register band: column j of the DP belongs to lane (j / K) % 32, register slot j % K, K = 1, 2, 4
// or 8 consecutive columns per lane.  The 32*K columns form a ring that the band travels along, so
// nothing is ever shifted between lanes; the only lane-to-lane traffic per anti-diagonal is the
// rotation of slot K-1 (the left neighbour of slot 0).  Per column the registers hold the previous
// anti-diagonal (p*) and the one before it at column j-1 (q*, which is simply last step's left
// neighbour).  K follows the band width; a change of K goes through the shared-memory ring, which
// is also the hand-over format to the wide fallback (rows in shared / global memory).
//
// Bases: one nibble per base (reference: code | 8 when it matches nothing, query: code | 4), so
// that the XOR of a reference and a query nibble is zero exactly for an identical a/c/g/t pair.
// The reference nibbles of the K cells of a lane sit in a shift register fed from a shared-memory
// ring (one LDS per anti-diagonal); the query nibbles of its K columns come from a nibble-packed
// ring (one LDS).  Both rings are filled 32 bases at a time, one batch ahead of their use.

// Shared memory of one warp (offsets in 32-bit words):
//   CfgBig    K up to 8 and the wide fallback: 12 score rows x 256 columns (rows 0-2 / 4-6 = the two live
//             anti-diagonals at a change of K; all 12 belong to the wide fallback).  The rows only the wide
//             fallback uses are free while the register band runs; short alignments keep their whole
//             traceback there, so that the walk back never touches global memory: row 3 = per-diagonal
//             (row offset << 16 | first column), row 7 = reversed deltas, rows 8-11 = traceback rows (4 KB).
struct CfgBig {
    static constexpr int MAXK = 8, RW = 256, CR = 512;
    static constexpr int META_OFF = 3 * 256, META_N = 256, REV_OFF = 7 * 256, REV_N = 256, TBROWS_OFF = 8 * 256, TBROWS_BYTES = 4096;
    static constexpr int BASE_OFF = 12 * 256;
    static constexpr int WARP_BYTES = 12 * 256 * 4 + 512 + 256;      // rows + reference ring (CR bytes) + query nibble ring (CR / 2 bytes)
};
template <class Cfg> __device__ __forceinline__ int32_t *eng_warp_smem()
{
    extern __shared__ __align__(16) int32_t smem_all[];
    return smem_all + (threadIdx.x >> 5) * (Cfg::WARP_BYTES / 4);
}

struct EngCtx {                          // warp-uniform progress of one alignment
    int N, M, dir, breaklen;
    bool forced, search;
    int d, tlo, thi, plo, phi, pplo, pphi;
    int high, best_d, best_j, reached;   // high: plain (unscaled) score
    unsigned long long cells;
    int ext_i, ext_j;                    // largest reference / query index of any evaluated cell
    uint8_t *tcur, *tend; bool arena_fail;
    int tb_sm_used, tb_sm_n; bool tb_sm_open;   // traceback rows kept in shared memory: bytes used, diagonals [0, tb_sm_n), still appending
    int ca_hi, cb_hi;                    // first unfilled index of the reference / query ring
    int pa, pb;                          // this lane's nibble of the batch fetched ahead
    int64_t Apos0, Bpos0;
};

#define SCR(buf, st, j) ring[((buf) * 4 + (st)) * Cfg::RW + ((j) & (Cfg::RW - 1))]

__device__ __forceinline__ int eng_ref_nibble(const ExShared &X, const EngCtx &c, int i)
{
    if (i < 1 || i > c.N) return 8;
    const int b = pmn_base_at(X.R, c.Apos0 + (int64_t)c.dir * (i - 1));
    return b < 4 ? b : 8;
}
__device__ __forceinline__ int eng_qry_nibble(const PackedView &Q, const EngCtx &c, int j)
{
    if (j < 1 || j > c.M) return 4;
    const int b = pmn_base_at(Q, c.Bpos0 + (int64_t)c.dir * (j - 1));
    return b < 4 ? b : 4;
}

// write the batch fetched ahead into the ring, fetch the next one
template <class Cfg>
__device__ __forceinline__ void eng_fill_ref(const ExShared &X, EngCtx &c, uint8_t *ca, int lane)
{
    ca[(c.ca_hi + lane) & (Cfg::CR - 1)] = (uint8_t)c.pa;
    c.ca_hi += 32;
    c.pa = eng_ref_nibble(X, c, c.ca_hi + lane);
    __syncwarp();
}
template <class Cfg>
__device__ __forceinline__ void eng_fill_qry(const PackedView &Q, EngCtx &c, uint32_t *cbw, int lane)
{
    unsigned v = (unsigned)c.pb << (4 * (lane & 7));
    v |= __shfl_xor_sync(0xffffffffu, v, 1); v |= __shfl_xor_sync(0xffffffffu, v, 2); v |= __shfl_xor_sync(0xffffffffu, v, 4);
    if ((lane & 7) == 0) cbw[((c.cb_hi >> 3) + (lane >> 3)) & (Cfg::CR / 8 - 1)] = v;
    c.cb_hi += 32;
    c.pb = eng_qry_nibble(Q, c, c.cb_hi + lane);
    __syncwarp();
}

#define ENG_DONE 0
#define ENG_GROW 1
#define ENG_SHRINK 2

// Runs anti-diagonals from c.d on with K columns per lane until the alignment ends (ENG_DONE), the
// band needs more columns (ENG_GROW) or fits half as many (ENG_SHRINK).  Enters and leaves with
// the two live anti-diagonals in the shared-memory ring (plain scores; rows 0-2 = diagonal d-2,
// rows 4-6 = diagonal d-1).
template <int K, class Cfg>
__device__ __forceinline__ int eng_run_reg(const Eng &E, EngCtx &cref, const PackedView &Q, uint8_t **tboff, int32_t *tblo)
{
    // the caller's context lives in local memory (it is handed to the wide fallback by value); the loop works on a copy
    // whose address never leaves this function, so that every field stays in a register
    EngCtx c = cref;
    constexpr int LOGK = K == 1 ? 0 : (K == 2 ? 1 : (K == 4 ? 2 : 3));
    constexpr int W = 32 * K;
    constexpr unsigned QMASK = K == 8 ? 0xffffffffu : ((1u << (4 * K)) - 1u);
    const ExShared &X = *E.X;
    const int lane = E.lane;
    int32_t *ring = eng_warp_smem<Cfg>();
    uint8_t *ca = (uint8_t *)(ring + Cfg::BASE_OFF);
    uint32_t *cbw = (uint32_t *)(ca + Cfg::CR);
    const int N = c.N, M = c.M;
    const int NEG4 = PMN_NEG * 4;
    const int max_diff4 = 4 * PMN_GOOD_SCORE * c.breaklen;
    int high4 = c.high * 4;

    int pD[K], pI[K], pM[K], qD[K], qI[K], qM[K], cm[K];
    int b_cur; unsigned aw = 0;
    int Bb_last;
    {   // load the live anti-diagonals into the window of the first anti-diagonal to run
        const int clo = c.tlo > c.d - N ? c.tlo : c.d - N;
        const int Bb = (clo - 1) >> LOGK;
        const int b = Bb + ((lane - Bb) & 31), jbase = b << LOGK;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const int j = jbase + s;
            const bool pv = j >= c.plo && j <= c.phi, qv = j - 1 >= c.pplo && j - 1 <= c.pphi;
            pD[s] = pv ? SCR(1, 0, j) * 4 : NEG4; pI[s] = pv ? SCR(1, 1, j) * 4 : NEG4; pM[s] = pv ? SCR(1, 2, j) * 4 : NEG4;
            qD[s] = qv ? SCR(0, 0, j - 1) * 4 : NEG4; qI[s] = qv ? SCR(0, 1, j - 1) * 4 : NEG4; qM[s] = qv ? SCR(0, 2, j - 1) * 4 : NEG4;
        }
        b_cur = b - 64;                 // forces the reference shift register to be primed
        Bb_last = Bb;
        __syncwarp();
    }

    int rc = ENG_DONE;
    for (;; c.d++) {
        const int d = c.d;
        if (!c.forced && d - c.best_d > c.breaklen) break;
        const int clo = c.tlo > d - N ? c.tlo : d - N, chi = c.thi + 1 < M ? c.thi + 1 : M;
        if (clo > chi) break;
        const int span = chi - c.plo + K + 4;
        if (span > W) { rc = ENG_GROW; break; }
        if (K > 1 && span + 8 <= W / 2) { rc = ENG_SHRINK; break; }
        const int width = chi - clo + 1;
        const int Bb = (clo - 1) >> LOGK;
        const int b = Bb + ((lane - Bb) & 31), jbase = b << LOGK;
        Bb_last = Bb;
        // base rings
        { const int need_i = d - clo < N ? d - clo : N; while (c.ca_hi <= need_i) eng_fill_ref<Cfg>(X, c, ca, lane); }
        while (c.cb_hi <= chi + 7) eng_fill_qry<Cfg>(Q, c, cbw, lane);
        const int i0 = d - jbase;       // reference index of slot 0
        if (K > 1 && b != b_cur) {
#pragma unroll
            for (int s = K - 1; s >= 1; s--) aw = (aw << 4) | ca[(i0 - s) & (Cfg::CR - 1)];
        }
        b_cur = b;
        aw = (aw << 4) | ca[i0 & (Cfg::CR - 1)];
        const unsigned qw = (cbw[(jbase >> 3) & (Cfg::CR / 8 - 1)] >> (4 * (jbase & 7))) & QMASK;
        const unsigned x = aw ^ qw;
        // traceback row: in shared memory while it fits (short alignments never leave it), else global
        uint8_t *trow = nullptr; int trow_s = -1;
        if (!c.search) {
            if (c.tb_sm_open && d < Cfg::META_N && c.tb_sm_used + W <= Cfg::TBROWS_BYTES) {
                trow_s = c.tb_sm_used; c.tb_sm_used += W; c.tb_sm_n = d + 1;
                if (lane == 0) ring[Cfg::META_OFF + d] = (trow_s << 16) | ((Bb << LOGK) & 0xffff);
            } else {
                c.tb_sm_open = false;
                if (c.tcur + W > c.tend) {
                    unsigned long long at = 0;
                    if (lane == 0) at = atomicAdd(X.counters + 1, (unsigned long long)EX_ARENA_CHUNK);
                    at = __shfl_sync(0xffffffffu, at, 0);
                    if (at + EX_ARENA_CHUNK > X.arena_cap) { c.arena_fail = true; break; }
                    c.tcur = X.arena + at; c.tend = c.tcur + EX_ARENA_CHUNK;
                }
                trow = c.tcur; c.tcur += W;
                if (lane == 0) { tboff[d] = trow; tblo[d] = Bb << LOGK; }
            }
        }
        // left neighbour of slot 0: slot K-1 of the lane below (ring rotation)
        int LD = __shfl_sync(0xffffffffu, pD[K - 1], (lane + 31) & 31);
        int LI = __shfl_sync(0xffffffffu, pI[K - 1], (lane + 31) & 31);
        int LM = __shfl_sync(0xffffffffu, pM[K - 1], (lane + 31) & 31);
        unsigned tbw[(K + 3) / 4];
#pragma unroll
        for (int w = 0; w < (K + 3) / 4; w++) tbw[w] = 0;
        const unsigned range = (unsigned)(chi - clo);
#pragma unroll
        for (int s = K - 1; s >= 0; s--) {
            const int j = jbase + s;
            const bool act = (unsigned)(j - clo) <= range;
            int lD, lI, lM;
            if (s > 0) { lD = pD[s - 1]; lI = pI[s - 1]; lM = pM[s - 1]; } else { lD = LD; lI = LI; lM = LM; }
            const int sc = ((x >> (4 * s)) & 0xfu) ? 4 * PMN_BAD_SCORE : 4 * PMN_GOOD_SCORE;
            const int mD = __vimax3_s32(lD + (4 * PMN_CONT_GAP_SCORE + PMN_ST_DEL), lI + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_INS), lM + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_MAT));
            const int mI = __vimax3_s32(pD[s] + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_DEL), pI[s] + (4 * PMN_CONT_GAP_SCORE + PMN_ST_INS), pM[s] + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_MAT));
            const int mM = __vimax3_s32(qD[s] + sc + PMN_ST_DEL, qI[s] + sc + PMN_ST_INS, qM[s] + sc + PMN_ST_MAT);
            const int vD = mD & ~3, vI = mI & ~3, vM = mM & ~3;
            const int mc = __vimax3_s32(vD + PMN_ST_DEL, vI + PMN_ST_INS, vM + PMN_ST_MAT);
            const unsigned tb = ((unsigned)mD & 3u) | (((unsigned)mI & 3u) << 2) | (((unsigned)mM & 3u) << 4) | (((unsigned)mc & 3u) << 6);
            tbw[s >> 2] |= tb << (8 * (s & 3));
            qD[s] = lD; qI[s] = lI; qM[s] = lM;
            pD[s] = act ? vD : NEG4; pI[s] = act ? vI : NEG4; pM[s] = act ? vM : NEG4;
            cm[s] = act ? (mc & ~3) : INT32_MIN;
        }
        if (!c.search) {
            const int at = ((lane - Bb) & 31) << LOGK;
            if (trow_s >= 0) {
                uint8_t *dst = (uint8_t *)(ring + Cfg::TBROWS_OFF) + trow_s + at;
                if (K == 1) *dst = (uint8_t)tbw[0];
                else if (K == 2) *(uint16_t *)dst = (uint16_t)tbw[0];
                else if (K == 4) *(uint32_t *)dst = tbw[0];
                else *(uint2 *)dst = make_uint2(tbw[0], tbw[(K + 3) / 4 - 1]);
            } else {
                uint8_t *dst = trow + at;
                if (K == 1) *dst = (uint8_t)tbw[0];
                else if (K == 2) *(uint16_t *)dst = (uint16_t)tbw[0];
                else if (K == 4) *(uint32_t *)dst = tbw[0];
                else *(uint2 *)dst = make_uint2(tbw[0], tbw[(K + 3) / 4 - 1]);
            }
        }
        int lmax = cm[0];
#pragma unroll
        for (int s = 1; s < K; s++) lmax = max(lmax, cm[s]);
        const int cmax = __reduce_max_sync(0xffffffffu, lmax);
        c.cells += (unsigned long long)width;
        c.ext_i = max(c.ext_i, d - clo); c.ext_j = max(c.ext_j, chi);
        if (cmax >= high4) {
            int jm = -1;
#pragma unroll
            for (int s = 0; s < K; s++) if (cm[s] == cmax) jm = jbase + s;
            high4 = cmax; c.best_d = d; c.best_j = __reduce_max_sync(0xffffffffu, jm);
        }
        c.pplo = c.plo; c.pphi = c.phi; c.plo = clo; c.phi = chi;
        if (d == N + M) { c.reached = 1; break; }
        if (!c.forced) {
            const int t4 = high4 - max_diff4;
            bool below = false;
#pragma unroll
            for (int s = 0; s < K; s++) below |= (cm[s] < t4) & (cm[s] != INT32_MIN);
            if (__any_sync(0xffffffffu, below)) {
                int lo = INT32_MAX, hi = INT32_MIN;
#pragma unroll
                for (int s = 0; s < K; s++) if (cm[s] >= t4) { lo = min(lo, jbase + s); hi = jbase + s; }
                lo = __reduce_min_sync(0xffffffffu, lo); hi = __reduce_max_sync(0xffffffffu, hi);
                if (lo == INT32_MAX) { c.tlo = chi + 1; c.thi = chi; }     // nothing survives: the oracle's loops leave (chi+1, chi)
                else { c.tlo = lo; c.thi = hi; }
            } else { c.tlo = clo; c.thi = chi; }
        } else { c.tlo = clo; c.thi = chi; }
    }
    c.high = high4 >> 2;
    if (rc != ENG_DONE) {
        // park the two live anti-diagonals in the ring for the next layout
        const int b = Bb_last + ((lane - Bb_last) & 31), jbase = b << LOGK;
#pragma unroll
        for (int s = 0; s < K; s++) {
            const int j = jbase + s;
            if (j >= c.plo && j <= c.phi) { SCR(1, 0, j) = pD[s] >> 2; SCR(1, 1, j) = pI[s] >> 2; SCR(1, 2, j) = pM[s] >> 2; }
            if (j - 1 >= c.pplo && j - 1 <= c.pphi) { SCR(0, 0, j - 1) = qD[s] >> 2; SCR(0, 1, j - 1) = qI[s] >> 2; SCR(0, 2, j - 1) = qM[s] >> 2; }
        }
        __syncwarp();
    }
    cref = c;
    return rc;
}

// The wide fallback: any band width, rows in the shared-memory ring, then in global memory.
// Enters with the two live anti-diagonals in the ring.
__device__ __noinline__ EngCtx eng_run_wide(const Eng &E, EngCtx c, const PackedView &Q, uint8_t **tboff, int32_t *tblo)
{
    const ExShared &X = *E.X;
    const int lane = E.lane;
    const int N = c.N, M = c.M, dir = c.dir;
    const int max_diff = PMN_GOOD_SCORE * c.breaklen;
    bool in_smem = true;
    typedef CfgBig Cfg;
    int32_t *base = eng_warp_smem<Cfg>(); int stride = Cfg::RW; int mask = Cfg::RW - 1;
    int bpp = 0, bp = 1, bc = 2;
#define SC(buf, st, j) base[((buf) * 4 + (st)) * stride + ((j) & mask)]
    if (!c.search && c.tb_sm_n > 0) {
        // the rows this path is about to use hold the traceback of the first diagonals: move it to global memory
        const int32_t *meta = base + Cfg::META_OFF;
        const uint8_t *rows = (const uint8_t *)(base + Cfg::TBROWS_OFF);
        for (int dd = 0; dd < c.tb_sm_n; dd++) {
            const int m0 = meta[dd], off = m0 >> 16, end = dd + 1 < c.tb_sm_n ? (meta[dd + 1] >> 16) : c.tb_sm_used;
            const int len = end - off;
            if (c.tcur + len > c.tend) {
                unsigned long long at = 0;
                if (lane == 0) at = atomicAdd(X.counters + 1, (unsigned long long)EX_ARENA_CHUNK);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at + EX_ARENA_CHUNK > X.arena_cap) { c.arena_fail = true; break; }
                c.tcur = X.arena + at; c.tend = c.tcur + EX_ARENA_CHUNK;
            }
            for (int k = lane; k < len; k += 32) c.tcur[k] = rows[off + k];
            if (lane == 0) { tboff[dd] = c.tcur; tblo[dd] = (int)(short)(m0 & 0xffff); }
            c.tcur += len;
        }
        __syncwarp();
        c.tb_sm_n = 0;
    }
    c.tb_sm_open = false;
    for (; !c.arena_fail && c.d <= N + M; c.d++) {
        const int d = c.d;
        if (!c.forced && d - c.best_d > c.breaklen) break;
        const int clo = c.tlo > d - N ? c.tlo : d - N, chi = c.thi + 1 < M ? c.thi + 1 : M;
        if (clo > chi) break;
        const int width = chi - clo + 1;
        if (in_smem && width > EX_BAND_SMEM) {
            // move the two live anti-diagonals to global rows and carry on there
            int32_t *g = E.gsc;
            for (int st = 0; st < 3; st++) {
                for (int j = c.plo + lane; j <= c.phi; j += 32) g[(bp * 4 + st) * EX_WCAP + j] = SC(bp, st, j);
                for (int j = c.pplo + lane; j <= c.pphi; j += 32) g[(bpp * 4 + st) * EX_WCAP + j] = SC(bpp, st, j);
            }
            __syncwarp();
            in_smem = false; base = g; stride = EX_WCAP; mask = -1;
        }
        uint8_t *trow = nullptr;
        if (!c.search) {
            if (c.tcur + width > c.tend) {
                unsigned long long need = width > (int)EX_ARENA_CHUNK ? (unsigned long long)width : EX_ARENA_CHUNK, at = 0;
                if (lane == 0) at = atomicAdd(X.counters + 1, need);
                at = __shfl_sync(0xffffffffu, at, 0);
                if (at + need > X.arena_cap) { c.arena_fail = true; break; }
                c.tcur = X.arena + at; c.tend = c.tcur + need;
            }
            trow = c.tcur; c.tcur += width;
            if (lane == 0) { tboff[d] = trow; tblo[d] = clo; }
        }
        int dmax = INT32_MIN, dmaxj = -1;
        int cm_f = INT32_MIN, cm_l = INT32_MIN;         // this lane's cell maxima in the first / last chunk
        const int jb_last = clo + ((chi - clo) & ~31);
        for (int jb = clo; jb <= chi; jb += 32) {
            const int j = jb + lane;
            const bool act = j <= chi;
            int cm = INT32_MIN;
            if (act) {
                const int i = d - j;
                int U0 = PMN_NEG, U1 = PMN_NEG, U2 = PMN_NEG, L0 = PMN_NEG, L1 = PMN_NEG, L2 = PMN_NEG, P0 = PMN_NEG, P1 = PMN_NEG, P2 = PMN_NEG;
                if (j >= c.plo && j <= c.phi) { U0 = SC(bp, 0, j); U1 = SC(bp, 1, j); U2 = SC(bp, 2, j); }
                if (j - 1 >= c.plo && j - 1 <= c.phi) { L0 = SC(bp, 0, j - 1); L1 = SC(bp, 1, j - 1); L2 = SC(bp, 2, j - 1); }
                if (j - 1 >= c.pplo && j - 1 <= c.pphi) { P0 = SC(bpp, 0, j - 1); P1 = SC(bpp, 1, j - 1); P2 = SC(bpp, 2, j - 1); }
                int s = PMN_BAD_SCORE;
                if (i >= 1 && j >= 1) {
                    const int a_ = pmn_base_at(X.R, c.Apos0 + (int64_t)dir * (i - 1)), b_ = pmn_base_at(Q, c.Bpos0 + (int64_t)dir * (j - 1));
                    if (a_ == b_ && a_ != PMN_CODE_X) s = PMN_GOOD_SCORE;
                }
                int vD, vI, vM, uD, uI, uM;
                score_edit(L0 + PMN_CONT_GAP_SCORE, L1 + PMN_OPEN_GAP_SCORE, L2 + PMN_OPEN_GAP_SCORE, vD, uD);
                score_edit(U0 + PMN_OPEN_GAP_SCORE, U1 + PMN_CONT_GAP_SCORE, U2 + PMN_OPEN_GAP_SCORE, vI, uI);
                score_edit(P0 + s, P1 + s, P2 + s, vM, uM);
                const int ms = max_state(vD, vI, vM);
                cm = ms == PMN_ST_DEL ? vD : (ms == PMN_ST_INS ? vI : vM);
                SC(bc, 0, j) = vD; SC(bc, 1, j) = vI; SC(bc, 2, j) = vM; SC(bc, 3, j) = cm;
                if (!c.search) trow[j - clo] = (uint8_t)(uD | uI << 2 | uM << 4 | ms << 6);
            }
            if (jb == clo) cm_f = cm;
            if (jb == jb_last) cm_l = cm;
            // chunk maximum; among equal cells the largest j wins, and a later chunk wins ties too
            const int cmax = __reduce_max_sync(0xffffffffu, cm);
            const unsigned eq = __ballot_sync(0xffffffffu, act && cm == cmax);
            if (cmax >= dmax) { dmax = cmax; dmaxj = jb + 31 - __clz((int)eq); }
        }
        c.cells += (unsigned long long)width;
        c.ext_i = max(c.ext_i, d - clo); c.ext_j = max(c.ext_j, chi);
        __syncwarp();
        if (dmax >= c.high) { c.high = dmax; c.best_d = d; c.best_j = dmaxj; }
        c.pplo = c.plo; c.pphi = c.phi; c.plo = clo; c.phi = chi;
        if (d == N + M) { c.reached = 1; break; }
        if (!c.forced) {
            const int t = c.high - max_diff;
            int nlo = chi + 1, nhi = chi;               // nothing survives: the oracle's loops leave (chi+1, chi)
            for (int jb = clo; jb <= chi; jb += 32) {
                const int j = jb + lane;
                int v = INT32_MIN;
                if (j <= chi) v = jb == clo ? cm_f : (jb == jb_last ? cm_l : SC(bc, 3, j));
                const unsigned bal = __ballot_sync(0xffffffffu, v >= t);
                if (bal) { nlo = jb + __ffs(bal) - 1; break; }
            }
            if (nlo <= chi) {
                for (int jb = jb_last; jb >= clo; jb -= 32) {
                    const int j = jb + lane;
                    int v = INT32_MIN;
                    if (j <= chi && j >= nlo) v = jb == clo ? cm_f : (jb == jb_last ? cm_l : SC(bc, 3, j));
                    const unsigned bal = __ballot_sync(0xffffffffu, v >= t);
                    if (bal) { nhi = jb + 31 - __clz((int)bal); break; }
                }
            }
            c.tlo = nlo; c.thi = nhi;
        } else { c.tlo = clo; c.thi = chi; }
        { int x = bpp; bpp = bp; bp = bc; bc = x; }
    }
#undef SC
    return c;
}


// ------------------------------------------------------------------------------------ forced alignments: the whole matrix, one warp
//
// A FORCED forward alignment (extendBackward's re-alignment of the region a search found, or the bridge to the alignment
// it reached: oracle/pmn_oracle.c extend_backward) has no trimming and no early stop: every cell of the N x M matrix is
// evaluated and the traceback starts in (N, M).  At 10-15 % divergence these windows are several hundred bases on a side
// and the banded engine would run them in its widest fallback.  Here the warp is a systolic array instead: lane l owns a
// strip of 8 columns, handles row t - l at step t and hands the right edge of its strip to lane l + 1 by one shuffle; 32
// strips (256 columns) form a pass, the right edge of a pass goes through a 16-byte record per row.  Cell arithmetic,
// state bits and traceback bytes are those of k_ex_wave1_tpj.  Returns the state of the finish cell.
#define SYS_NEG (-(1 << 28))
__device__ __forceinline__ unsigned tpj_query_nibbles(const PackedView &Q, int64_t p, int left);

__device__ __noinline__ int eng_forced_systolic(const Eng &E, const PackedView &Q, int64_t Apos0, int64_t Bpos0, int N, int M, uint2 *tb, int4 *bnd)
{
    const ExShared &X = *E.X;
    const int lane = E.lane;
    const int npass = (M + 255) >> 8;
    int ms_fin = PMN_ST_MAT;
    for (int ps = 0; ps < npass; ps++) {
        const int j0 = ps * 256 + lane * 8;                 // this lane's columns: j0 + 1 .. j0 + 8
        const bool have = j0 < M;
        int nl = (M - ps * 256 + 7) >> 3; if (nl > 32) nl = 32;     // active lanes of this pass
        const unsigned qn = have ? tpj_query_nibbles(Q, Bpos0 + j0, M - j0) : 0x44444444u;
        int uI[8], ug[8];
#pragma unroll
        for (int c = 0; c < 8; c++) { uI[c] = SYS_NEG; ug[c] = SYS_NEG; }
        int prev_bmc = SYS_NEG;
        int oD = SYS_NEG, oH = SYS_NEG, oC = SYS_NEG;       // right edge of the row this lane finished last: (D, max(I+1, M+2), cell maximum)
        uint64_t aw = 0; uint32_t ax = 0;
        uint2 *trow = tb + (size_t)ps * (N + 1) * 32 + lane;
        const int steps = N + nl;
        for (int t = 0; t < steps; t++) {
            int lD = __shfl_up_sync(0xffffffffu, oD, 1), hl = __shfl_up_sync(0xffffffffu, oH, 1), bmc = __shfl_up_sync(0xffffffffu, oC, 1);
            const int i = t - lane;
            const bool act = have && i >= 0 && i <= N;
            if (lane == 0 && act) {
                if (ps == 0) {
                    if (i == 0) { lD = SYS_NEG; hl = 2; bmc = 2; }                       // cell (0,0) = MAT 0
                    else { const int v = 4 * (PMN_OPEN_GAP_SCORE + PMN_CONT_GAP_SCORE * (i - 1)) + PMN_ST_INS; lD = SYS_NEG; hl = v; bmc = v; }
                } else { const int4 b = bnd[i]; lD = b.x; hl = b.y; bmc = b.z; }
            }
            if (act) {
                unsigned an = 8;
                if (i >= 1) {
                    if (((i - 1) & 31) == 0 || i == 1) { aw = pmn_window64(X.R.w, Apos0 + i - 1); ax = X.R.has_x ? pmn_xwindow32(X.R.xm, Apos0 + i - 1) : 0u; }
                    an = (unsigned)(aw >> 62) | ((ax >> 31) << 3);
                    aw <<= 2; ax <<= 1;
                }
                const unsigned x = (an * 0x11111111u) ^ qn;
                int dmc = prev_bmc;
                prev_bmc = bmc;
                unsigned t0 = 0, t1 = 0; int mc = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int sc = (x & (0xfu << (4 * c))) ? 4 * PMN_BAD_SCORE : 4 * PMN_GOOD_SCORE;
                    const int mD = __viaddmax_s32(lD, 4 * PMN_CONT_GAP_SCORE + PMN_ST_DEL, hl + 4 * PMN_OPEN_GAP_SCORE);
                    const int mI = __viaddmax_s32(uI[c], 4 * PMN_CONT_GAP_SCORE + PMN_ST_INS, ug[c] + 4 * PMN_OPEN_GAP_SCORE);
                    const int mM = dmc + sc;
                    dmc = __viaddmax_s32(uI[c], PMN_ST_INS, ug[c]);
                    const int vD = mD & ~3, vI = mI & ~3, vM2 = (mM & ~3) + PMN_ST_MAT;
                    hl = __viaddmax_s32(vI, PMN_ST_INS, vM2);
                    mc = max(hl, vD);
                    uI[c] = vI; ug[c] = max(vD, vM2);
                    lD = vD;
                    const unsigned tbits = ((unsigned)mD & 3u) | (((unsigned)mI & 3u) << 2) | (((unsigned)mM & 3u) << 4);
                    if (c < 4) t0 |= tbits << (8 * c); else t1 |= tbits << (8 * (c - 4));
                }
                trow[(size_t)i * 32] = make_uint2(t0, t1);
                oD = lD; oH = hl; oC = mc;
                if (lane == nl - 1 && ps + 1 < npass) bnd[i] = make_int4(lD, hl, mc, 0);
            }
        }
        if (ps == npass - 1) {                              // the finish cell (N, M) sits in column (M-1) & 7 of lane ((M-1) >> 3) & 31
            const int c = (M - 1) & 7;
            int mcf = 0;
#pragma unroll
            for (int cc = 0; cc < 8; cc++) if (cc == c) mcf = __viaddmax_s32(uI[cc], PMN_ST_INS, ug[cc]);
            ms_fin = __shfl_sync(0xffffffffu, mcf, ((M - 1) >> 3) & 31) & 3;
        }
        __syncwarp();                                       // the records of this pass are read by lane 0 in the next one
    }
    return ms_fin;
}

// One alignment.  All 32 lanes call with identical arguments and get identical results.
// Returns reached (0/1); Aend/Bend become the finish cell.  Unless SEARCH, the deltas are
// appended to the pool: *doff, *dcnt.
template <class Cfg>
__device__ __noinline__ int align_engine(const Eng &E, int64_t Abase, int64_t Astart, int64_t &Aend, const PackedView &Q, int64_t Bbase, int64_t Bstart, int64_t &Bend,
                                         unsigned m_o, uint32_t *doff, int32_t *dcnt, int32_t *dasum, int *ext_i = nullptr, int *ext_j = nullptr)
{
    const ExShared &X = *E.X;
    const int lane = E.lane;
    EngCtx c;
    c.dir = (m_o & PMN_DIRECTION_BIT) ? 1 : -1;
    const int N = c.N = (int)(c.dir > 0 ? Aend - Astart + 1 : Astart - Aend + 1);
    const int M = c.M = (int)(c.dir > 0 ? Bend - Bstart + 1 : Bstart - Bend + 1);
    c.forced = (m_o & PMN_FORCED_BIT) != 0; c.search = (m_o & PMN_SEARCH_BIT) != 0;
    c.breaklen = X.breaklen;
    if (doff) { *doff = 0; *dcnt = 0; *dasum = 0; }
    if (N < 1 || M < 1 || N > PMN_MAX_ALIGNMENT_LENGTH || M > PMN_MAX_ALIGNMENT_LENGTH) {
        if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_LOGIC);
        return 0;
    }
    const long long dbg_t0 = X.dbg ? clock64() : 0;

    uint8_t **tboff = (uint8_t **)E.tbp;
    int32_t *tblo = (int32_t *)(E.tbp + (size_t)EX_DMAX * 8);
    int32_t *rev = (int32_t *)(E.tbp + (size_t)EX_DMAX * 12);
    c.tcur = E.tbp + EX_TB_HDR; c.tend = E.tbp + EX_TBW; c.arena_fail = false;
    c.tb_sm_used = 32; c.tb_sm_n = 1; c.tb_sm_open = true;      // row 0 = cell (0,0), written below

    c.Apos0 = Abase + Astart - 1; c.Bpos0 = Bbase + Bstart - 1;
    c.d = 1; c.tlo = 0; c.thi = 0; c.plo = 0; c.phi = 0; c.pplo = 1; c.pphi = 0;
    c.high = 0; c.best_d = 0; c.best_j = 0; c.reached = 0; c.cells = 0; c.ext_i = 0; c.ext_j = 0;
    {   // cell (0,0) into the ring; base rings: everything in front of the window matches nothing
        int32_t *ring = eng_warp_smem<Cfg>();
        uint8_t *ca = (uint8_t *)(ring + Cfg::BASE_OFF);
        uint32_t *cbw = (uint32_t *)(ca + Cfg::CR);
        if (lane == 0) {
            SCR(1, 0, 0) = PMN_NEG; SCR(1, 1, 0) = PMN_NEG; SCR(1, 2, 0) = 0;
            ring[Cfg::META_OFF] = 0;
            *(uint8_t *)(ring + Cfg::TBROWS_OFF) = (uint8_t)(PMN_ST_NONE | PMN_ST_NONE << 2 | PMN_ST_NONE << 4 | PMN_ST_MAT << 6);
        }
#pragma unroll
        for (int k = 0; k < Cfg::CR / 128; k++) ((uint32_t *)ca)[k * 32 + lane] = 0x08080808u;
        c.ca_hi = 1; c.pa = eng_ref_nibble(X, c, 1 + lane);
        c.cb_hi = 0; c.pb = eng_qry_nibble(Q, c, lane);
        __syncwarp();
        eng_fill_ref<Cfg>(X, c, ca, lane);
        eng_fill_qry<Cfg>(Q, c, cbw, lane);
    }

    // forced alignments over large windows: the systolic full-matrix path
    uint2 *sys_tb = nullptr; int sys_ms = PMN_ST_MAT;
    if (c.forced && !c.search && c.dir > 0 && (N < M ? N : M) >= 64) {
        const unsigned long long need = (unsigned long long)((M + 255) >> 8) * (unsigned long long)(N + 1) * 256ull;
        if (need <= (64ull << 20)) {
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(X.counters + 1, need);
            at = __shfl_sync(0xffffffffu, at, 0);
            if (at + need <= X.arena_cap) sys_tb = (uint2 *)(X.arena + at);       // else: the banded engine below reports the exhausted arena
        }
    }
    int mode = 1, path = 0;
    if (sys_tb) {
        sys_ms = eng_forced_systolic(E, Q, c.Apos0, c.Bpos0, N, M, sys_tb, (int4 *)E.gsc);
        c.reached = 1; c.d = N + M; c.cells = (unsigned long long)(N + 1) * (unsigned long long)(M + 1) - 1ull; c.ext_i = N; c.ext_j = M; path = 32;
    }
    else for (;;) {
        int rc;
        if (mode > Cfg::MAXK) {
            if (Cfg::MAXK < 8) return -1;          // too wide for this kernel: the caller hands the alignment to the big one
            c = eng_run_wide(E, c, Q, tboff, tblo); rc = ENG_DONE;
        }
        else if (mode == 1) rc = eng_run_reg<1, Cfg>(E, c, Q, tboff, tblo);
        else if (mode == 2) rc = eng_run_reg<2, Cfg>(E, c, Q, tboff, tblo);
        else if (Cfg::MAXK >= 4 && mode == 4) rc = eng_run_reg<(Cfg::MAXK >= 4 ? 4 : 1), Cfg>(E, c, Q, tboff, tblo);
        else rc = eng_run_reg<(Cfg::MAXK >= 8 ? 8 : 1), Cfg>(E, c, Q, tboff, tblo);
        if (rc == ENG_DONE) break;
        mode = rc == ENG_GROW ? mode * 2 : mode / 2;
        if (mode > path) path = mode;
    }
    const int d = c.d;
    if (lane == 0) { atomicAdd(X.counters + 2, c.cells); atomicAdd(X.counters + 3, 1ull); }
    if (X.dbg && lane == 0) {
        const unsigned long long at = atomicAdd(X.counters + 15, 1ull);
        if (at < X.dbg_cap) {
            X.dbg[2 * at] = make_int4((int)m_o, N, M, d);
            X.dbg[2 * at + 1] = make_int4((int)c.cells, (int)(clock64() - dbg_t0), E.kid, path);
        }
    }
    if (c.arena_fail) { if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_ARENA); return 0; }
    const int reached = c.reached;
    if (ext_i) { *ext_i = c.ext_i; *ext_j = c.ext_j; }

    int fd, fj;
    if (reached && !(m_o & PMN_OPTIMAL_BIT)) { fd = N + M; fj = M; } else { fd = c.best_d; fj = c.best_j; }
    const int fi = fd - fj;
    Aend = Astart + (int64_t)c.dir * (fi - 1);
    Bend = Bstart + (int64_t)c.dir * (fj - 1);

    if (!c.search) {
        __syncwarp();
        int32_t *ring = eng_warp_smem<Cfg>();
        const int32_t *meta = ring + Cfg::META_OFF;
        const uint8_t *rows = (const uint8_t *)(ring + Cfg::TBROWS_OFF);
        int32_t *rev_s = ring + Cfg::REV_OFF;
        const int n_sm = c.tb_sm_n;
        int nrev = 0;
        {
            // The walk back is a chain of dependent reads.  Rows that left shared memory would cost two dependent global loads
            // per step (row pointer, then the byte) — more than the DP itself on long alignments.  So the warp stages 32
            // anti-diagonals at a time: lane l fetches row pointer and first column of diagonal top-l (one coalesced load each)
            // and the 48 bytes of that row the path can touch (after l steps it is at most l columns left of where it is now)
            // into the score ring, which is dead by now; all lanes then walk in lockstep on shared memory, lane 0 emits.
            uint8_t *seg = (uint8_t *)ring;                     // rows 0-1 of the ring: 64 bytes per lane
            int32_t *segcol = ring + 2 * Cfg::RW;               // row 2: column of byte 0 of each lane's segment
            int win_top = -1, win_bot = 0;
            auto cell = [&](int cd, int cj) -> unsigned {
                if (sys_tb) {                           // systolic layout: [pass][row][lane] words of 8 columns; column 0 is implicit
                    const int ci = cd - cj;
                    if (cj == 0) return (unsigned)(ci == 1 ? PMN_ST_MAT : PMN_ST_INS) << 2;
                    const uint2 wv = sys_tb[((size_t)((cj - 1) >> 8) * (N + 1) + ci) * 32 + (((cj - 1) >> 3) & 31)];
                    const int cc = (cj - 1) & 7;
                    return ((cc < 4 ? wv.x : wv.y) >> (8 * (cc & 3))) & 0x3fu;
                }
                if (cd < n_sm) { const int m0 = meta[cd]; return rows[(m0 >> 16) + cj - (int)(short)(m0 & 0xffff)]; }
                if (cd > win_top || cd < win_bot) {
                    __syncwarp();
                    const int dd = cd - lane;
                    if (dd >= n_sm && dd >= 0) {
                        const uint8_t *p = tboff[dd]; const int lo = tblo[dd];
                        int o = cj - lane - lo; if (o < 0) o = 0;
                        const uint8_t *ab = (const uint8_t *)((uintptr_t)(p + o) & ~(uintptr_t)15);
                        const uint4 *src = (const uint4 *)ab;
                        const uint4 v0 = src[0], v1 = src[1], v2 = src[2];
                        uint4 *dst = (uint4 *)(seg + lane * 64);
                        dst[0] = v0; dst[1] = v1; dst[2] = v2;
                        segcol[lane] = lo + (int)(ab - p);
                    }
                    win_top = cd; win_bot = cd - 31 > n_sm ? cd - 31 : n_sm; if (win_bot < 0) win_bot = 0;
                    __syncwarp();
                }
                const int L = win_top - cd;
                return seg[L * 64 + (cj - segcol[L])];
            };
            auto emit = [&](int v) { if (lane == 0) { if (nrev < Cfg::REV_N) rev_s[nrev] = v; else rev[nrev] = v; } nrev++; };
            int cd = fd, cj = fj;
            int st = sys_tb ? sys_ms : (int)(cell(cd, cj) >> 6);
            int pending = 0, run = 0;
            while (cd > 0) {
                const unsigned b = cell(cd, cj);
                if (st == PMN_ST_MAT) { run++; st = (b >> 4) & 3; cd -= 2; cj -= 1; }
                else {
                    if (pending) emit(pending * (run + 1));
                    run = 0;
                    if (st == PMN_ST_INS) { pending = 1; st = (b >> 2) & 3; cd -= 1; }
                    else { pending = -1; st = b & 3; cd -= 1; cj -= 1; }
                }
            }
            if (pending) emit(pending * (run + 1));
        }
        nrev = __shfl_sync(0xffffffffu, nrev, 0);
        __syncwarp();
        if (nrev > 0) {
            unsigned long long at = 0;
            if (lane == 0) at = atomicAdd(X.counters + 0, (unsigned long long)nrev);
            at = __shfl_sync(0xffffffffu, at, 0);
            if (at + (unsigned long long)nrev > X.pool_cap) { if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_POOL); return reached; }
            int asum = 0;
            for (int k = lane; k < nrev; k += 32) {
                const int idx = nrev - 1 - k;
                const int dv = idx < Cfg::REV_N ? rev_s[idx] : rev[idx];
                X.pool[at + k] = dv; asum += dv > 0 ? dv : -dv - 1;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
            *doff = (uint32_t)at; *dcnt = nrev; *dasum = asum;
        }
        __syncwarp();
    }
    return reached;
}


// ------------------------------------------------------------------------------------ postnuc pieces shared by wave 1 and the stitcher

// returns the cluster index (syn-relative indices are absolute here) or `end` when there is none
__device__ int get_forward_target_cluster(const ExShared &X, int cp, int end, int64_t &targetA, int64_t &targetB)
{
    const ExCluster c = X.cl[cp];
    const int last = c.mfirst + c.nm - 1;
    const int64_t sA = (int64_t)X.mA[last] + X.mL[last] - 1, sB = (int64_t)X.mB[last] + X.mL[last] - 1;
    int64_t dist = targetA - sA < targetB - sB ? targetA - sA : targetB - sB;
    int best = end;
    for (int ci = cp + 1; ci < end; ci++) {
        const ExCluster t = X.cl[ci];
        if (t.dir != c.dir) continue;
        int64_t eA = X.mA[t.mfirst], eB = X.mB[t.mfirst];
        const int tl = t.mfirst + t.nm - 1;
        if ((eA < sA || eB < sB) && X.mA[tl] >= sA && X.mB[tl] >= sB)
            for (int k = t.mfirst; k <= tl && (eA < sA || eB < sB); k++) { eA = X.mA[k]; eB = X.mB[k]; }
        if (eA >= sA && eB >= sB) {
            int64_t greater, lesser;
            if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
            if (greater < X.breaklen || lesser * PMN_GOOD_SCORE + (greater - lesser) * PMN_CONT_GAP_SCORE >= 0) { best = ci; targetA = eA; targetB = eB; break; }
            else if ((greater << 1) - lesser < dist) { best = ci; targetA = eA; targetB = eB; dist = (greater << 1) - lesser; }
        }
    }
    return best;
}

// getForwardTargetCluster for every cluster at once: one warp per cluster, 32 candidates per step.  The sequential original
// (get_forward_target_cluster above, still used by the stitcher) stops at the first "close enough" cluster and otherwise keeps the
// first strict improvement of dist = 2*greater - lesser; a cluster in front of a rearrangement walks the whole rest of its
// synteny, which made one thread per cluster the long pole of k_ex_jobdesc (70 us of a 5 Mbp pair).
__global__ void __launch_bounds__(256) k_ex_targets(ExShared X, int4 *__restrict__ tgt)
{
    const int cp = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (cp >= X.nC) return;
    const ExCluster c = X.cl[cp];
    const ExSynteny S = X.syn[c.syn];
    const int end = S.cfirst + S.nC;
    const int last = c.mfirst + c.nm - 1;
    const int64_t sA = (int64_t)X.mA[last] + X.mL[last] - 1, sB = (int64_t)X.mB[last] + X.mL[last] - 1;
    int64_t targetA = S.lenA, targetB = S.lenB;
    long long dist = targetA - sA < targetB - sB ? targetA - sA : targetB - sB;
    int best = end;
    for (int base = cp + 1; base < end; base += 32) {
        const int ci = base + lane;
        bool valid = false, close = false; long long score = 0; int64_t eA = 0, eB = 0;
        if (ci < end) {
            const ExCluster t = X.cl[ci];
            if (t.dir == c.dir) {
                eA = X.mA[t.mfirst]; eB = X.mB[t.mfirst];
                const int tl = t.mfirst + t.nm - 1;
                if ((eA < sA || eB < sB) && X.mA[tl] >= sA && X.mB[tl] >= sB)
                    for (int k = t.mfirst; k <= tl && (eA < sA || eB < sB); k++) { eA = X.mA[k]; eB = X.mB[k]; }
                if (eA >= sA && eB >= sB) {
                    valid = true;
                    int64_t greater, lesser;
                    if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
                    close = greater < X.breaklen || lesser * PMN_GOOD_SCORE + (greater - lesser) * PMN_CONT_GAP_SCORE >= 0;
                    score = (greater << 1) - lesser;
                }
            }
        }
        const unsigned cb = __ballot_sync(0xffffffffu, close);
        const int F = cb ? __ffs((int)cb) - 1 : 32;
        long long key = (valid && !close && lane < F) ? score * 32 + lane : LLONG_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const long long k2 = __shfl_xor_sync(0xffffffffu, key, o); if (k2 < key) key = k2; }
        int src = -1;
        if (key != LLONG_MAX && (key >> 5) < dist) { dist = key >> 5; src = (int)(key & 31); }
        if (cb) src = F;
        if (src >= 0) {
            best = base + src;
            targetA = __shfl_sync(0xffffffffu, eA, src); targetB = __shfl_sync(0xffffffffu, eB, src);
        }
        if (cb) break;
    }
    if (lane == 0) tgt[cp] = make_int4((int)targetA, (int)targetB, best, 0);
}

// the alignment-engine part of extendForward, from (eA, eB) towards (targetA, targetB)
__device__ int forward_job(const Eng &E, const ExSynteny &S, int dirB, int64_t eA, int64_t eB, int64_t targetA, int64_t targetB, unsigned m_o, ExJob &out)
{
    int overflow = 0;
    if (targetA - eA + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetA = eA + PMN_MAX_ALIGNMENT_LENGTH - 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (targetB - eB + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetB = eB + PMN_MAX_ALIGNMENT_LENGTH - 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    uint32_t doff; int32_t dcnt, dasum;
    int reached = align_engine<CfgBig>(E, S.Abase, eA, targetA, dirB ? E.X->QR : E.X->QF, dirB ? S.BbaseR : S.BbaseF, eB, targetB, m_o, &doff, &dcnt, &dasum);
    if (reached && overflow) reached = 0;
    out.endA = (int32_t)targetA; out.endB = (int32_t)targetB; out.dcnt = dcnt; out.doff = doff; out.reached = reached; out.valid = 1; out.asum = dasum;
    return reached;
}

// ------------------------------------------------------------------------------------ E2: wave 1

__device__ __forceinline__ Eng make_eng(const ExShared &X, int32_t *)
{
    Eng E;
    const int warp = threadIdx.x >> 5;
    const size_t slot = (size_t)blockIdx.x * EX_WARPS_PER_BLOCK + warp;
    E.X = &X; E.lane = threadIdx.x & 31;
    E.ssc = nullptr;
    E.gsc = X.gscore + slot * (size_t)EX_ROWS * EX_WCAP;
    E.tbp = X.tbpriv + slot * (size_t)EX_TBW;
    E.kid = 0;
    return E;
}

__device__ __forceinline__ int run_mismatches(const PackedView &R, int64_t a, const PackedView &Q, int64_t b, int64_t run);


// ------------------------------------------------------------------------------------ thread-per-job windows
//
// Most match -> next match alignments are small windows (N x M up to a few thousand cells).  For
// those one THREAD runs the whole alignment (k_ex_wave1_tpj): a warp holds 32 jobs of similar shape
// (counting sort by bin below), so the 32 lanes do useful work on every instruction instead of one
// anti-diagonal of a single small matrix being spread thinly over them.
#define TPJ_W 8                        /* columns per strip (kept in registers)                      */
#define TPJ_TB_WORDS (13 * 101)        /* traceback words (8 cells each) per thread: any window up to 100 x 100 */
#define TPJ_BND_ROWS 104               /* strip boundary records (16 B) per thread, also delta staging */
#define TPJ_MAXDIM 100
#define TPJ_SLOT_BYTES (32 * (TPJ_TB_WORDS * 8 + TPJ_BND_ROWS * 16))
#define TPJ_NB 26                      /* buckets of 4 rows                                          */
#define TPJ_BINS (13 * TPJ_NB)
#define TPJ_NEG (-(1 << 28))
#define TPJ_BLOCKS_PER_SM 4

__device__ __forceinline__ bool tpj_fits(int N, int M, int breaklen, int maxcells)
{
    return N >= 1 && M >= 1 && N <= TPJ_MAXDIM && M <= TPJ_MAXDIM && N * M <= maxcells && N + M <= breaklen && ((M + TPJ_W - 1) / TPJ_W) * (N + 1) <= TPJ_TB_WORDS;
}
// large windows first: (strips descending, rows descending)
__device__ __forceinline__ int tpj_bin(int N, int M) { return (13 - (M + TPJ_W - 1) / TPJ_W) * TPJ_NB + (TPJ_NB - 1 - (N >> 2)); }

// one thread per match: its forward job (match -> next match), and for the last match of a cluster the
// cluster-end job (-> target cluster or as far as the score carries).
//
// Shortcut (most gaps between two matches are a single substitution): a square window of k <= 20 columns
// with at most one mismatching column is solved here.  The gap-free path scores >= 3k - 10, any path with
// a gap has at most k-1 match columns and two gap openings, <= 3(k-1) - 14; the same holds for every prefix,
// so the engine's traceback would follow the main diagonal: target reached, no deltas.  No cell of such a
// window can fall 3*breaklen below the high score when breaklen >= 100 (worst cell -223, high <= 60), so the engine would have evaluated the full
// (k+1)^2 - 1 cells, which is what the counters are credited with.
__global__ void __launch_bounds__(256) k_ex_jobdesc(ExShared X, ExJobDesc *__restrict__ descA, ExJobDesc *__restrict__ descB)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned sc_cells = 0, sc_jobs = 0;             // shortcut windows of this thread: one atomic per warp below (all of them hit two addresses)
    if (g < X.nM) {
        const int k = X.mcl[g];
        const ExCluster c = X.cl[k];
        const ExSynteny S = X.syn[c.syn];
        ExJobDesc d;
        d.Abase = S.Abase; d.Bbase = c.dir ? S.BbaseR : S.BbaseF;
        d.eA = X.mA[g] + X.mL[g] - 1; d.eB = X.mB[g] + X.mL[g] - 1; d.g = (int32_t)g; d.dir = c.dir;
        if ((int)g != c.mfirst + c.nm - 1) {
            d.tA = X.mA[g + 1]; d.tB = X.mB[g + 1]; d.m_o = PMN_FORWARD_ALIGN; d.target = -1;
            const int n = d.tA - d.eA + 1, m = d.tB - d.eB + 1;
            if (n == m && n >= 1 && n <= 20 && X.breaklen >= 100 &&
                run_mismatches(X.R, d.Abase + d.eA - 1, c.dir ? X.QR : X.QF, d.Bbase + d.eB - 1, n) <= 1) {
                ExJob r; r.endA = d.tA; r.endB = d.tB; r.dcnt = 0; r.target = -1; r.doff = 0; r.reached = 1; r.valid = 1; r.asum = 0;
                X.jobs[g] = r;
                sc_cells = (unsigned)((n + 1) * (m + 1) - 1); sc_jobs = 1;
                d.m_o = -1;
            }
            else if (tpj_fits(n, m, X.breaklen, X.tpj_cells)) {
                const int bin = tpj_bin(n, m);
                X.tkey[g] = make_uint2((unsigned)bin, atomicAdd(X.tbin + bin, 1u));
            }
            else X.overflow[atomicAdd(X.counters + 7, 1ull)] = (int32_t)g;
            descB[g] = d;
        } else {
            d.m_o = -1; d.tA = d.tB = 0; d.target = -1;
            descB[g] = d;                                       // the last match has no inner job
            if (X.do_extend) {
                const int end = S.cfirst + S.nC;
                const int4 t4 = X.tgt[k];                       // getForwardTargetCluster, one warp per cluster (k_ex_targets)
                const int tc = t4.z;
                d.tA = t4.x; d.tB = t4.y; d.target = tc;
                d.m_o = PMN_FORWARD_ALIGN | (tc == end ? PMN_OPTIMAL_BIT : 0);
                if (tc < end) atomicAdd(X.aimers + tc, 1);      // one more end job aims at tc (see k_ex_wave1_big)
            }
            descA[k] = d;
        }
    }
    sc_cells = __reduce_add_sync(0xffffffffu, sc_cells); sc_jobs = __reduce_add_sync(0xffffffffu, sc_jobs);
    if ((threadIdx.x & 31) == 0 && sc_jobs) { atomicAdd(X.counters + 2, (unsigned long long)sc_cells); atomicAdd(X.counters + 3, (unsigned long long)sc_jobs); }
}

// A forward alignment (not OPTIMAL) over a window of at most 31 x 31 bases, the bulk of wave 1.  With
// breaklen >= 134 no cell of such a window can fall 3*breaklen below the running high score (worst cell
// >= -7*31 - 4*30 - 7, high <= 93), so the engine would evaluate the full matrix and finish on the target
// cell: the band, high-score and trimming logic drop out.  Lane j owns column j for the whole alignment,
// the traceback (one byte per cell, 32 per anti-diagonal) stays in shared memory.
template <class Cfg>
__device__ __noinline__ void eng_small_full(const Eng &E, const ExShared &X, const ExJobDesc &d, int N, int M, uint32_t *doff, int32_t *dcnt, int32_t *dasum)
{
    const int lane = E.lane;
    int32_t *ring = eng_warp_smem<Cfg>();
    uint8_t *rows = (uint8_t *)(ring + Cfg::TBROWS_OFF);
    int32_t *rev_s = ring + Cfg::REV_OFF;
    const PackedView &Q = d.dir ? X.QR : X.QF;
    const int NEG4 = PMN_NEG * 4;
    *doff = 0; *dcnt = 0; *dasum = 0;
    // lane l holds reference base A'[l] and query base B'[l] (index 0: nothing)
    int an = 8, qn = 4;
    if (lane >= 1 && lane <= N) { const int b = pmn_base_at(X.R, d.Abase + d.eA - 1 + (lane - 1)); an = b < 4 ? b : 8; }
    if (lane >= 1 && lane <= M) { const int b = pmn_base_at(Q, d.Bbase + d.eB - 1 + (lane - 1)); qn = b < 4 ? b : 4; }
    int pD = NEG4, pI = NEG4, pM = lane == 0 ? 0 : NEG4, qD = NEG4, qI = NEG4, qM = NEG4;
    if (lane == 0) rows[0] = (uint8_t)(PMN_ST_NONE | PMN_ST_NONE << 2 | PMN_ST_NONE << 4 | PMN_ST_MAT << 6);
    const int D = N + M;
#pragma unroll 2
    for (int dd = 1; dd <= D; dd++) {
        int lD = __shfl_up_sync(0xffffffffu, pD, 1), lI = __shfl_up_sync(0xffffffffu, pI, 1), lM = __shfl_up_sync(0xffffffffu, pM, 1);
        const int i = dd - lane;
        const int ai = __shfl_sync(0xffffffffu, an, i & 31);
        if (lane == 0) { lD = NEG4; lI = NEG4; lM = NEG4; }
        const bool act = (unsigned)i <= (unsigned)N && lane <= M;
        const int sc = ((ai ^ qn) == 0 && i >= 1) ? 4 * PMN_GOOD_SCORE : 4 * PMN_BAD_SCORE;
        const int mD = __vimax3_s32(lD + (4 * PMN_CONT_GAP_SCORE + PMN_ST_DEL), lI + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_INS), lM + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_MAT));
        const int mI = __vimax3_s32(pD + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_DEL), pI + (4 * PMN_CONT_GAP_SCORE + PMN_ST_INS), pM + (4 * PMN_OPEN_GAP_SCORE + PMN_ST_MAT));
        const int mM = __vimax3_s32(qD + sc + PMN_ST_DEL, qI + sc + PMN_ST_INS, qM + sc + PMN_ST_MAT);
        const int vD = mD & ~3, vI = mI & ~3, vM = mM & ~3;
        const int mc = __vimax3_s32(vD + PMN_ST_DEL, vI + PMN_ST_INS, vM + PMN_ST_MAT);
        rows[dd * 32 + lane] = (uint8_t)(((unsigned)mD & 3u) | (((unsigned)mI & 3u) << 2) | (((unsigned)mM & 3u) << 4) | (((unsigned)mc & 3u) << 6));
        qD = lD; qI = lI; qM = lM;
        pD = act ? vD : NEG4; pI = act ? vI : NEG4; pM = act ? vM : NEG4;
    }
    if (lane == 0) { atomicAdd(X.counters + 2, (unsigned long long)((N + 1) * (M + 1) - 1)); atomicAdd(X.counters + 3, 1ull); }
    __syncwarp();
    int nrev = 0;
    if (lane == 0) {
        int cd = D, cj = M;
        int st = rows[cd * 32 + cj] >> 6;
        int pending = 0, run = 0;
        while (cd > 0) {
            const unsigned b = rows[cd * 32 + cj];
            if (st == PMN_ST_MAT) { run++; st = (b >> 4) & 3; cd -= 2; cj -= 1; }
            else {
                if (pending) rev_s[nrev++] = pending * (run + 1);
                run = 0;
                if (st == PMN_ST_INS) { pending = 1; st = (b >> 2) & 3; cd -= 1; }
                else { pending = -1; st = b & 3; cd -= 1; cj -= 1; }
            }
        }
        if (pending) rev_s[nrev++] = pending * (run + 1);
    }
    nrev = __shfl_sync(0xffffffffu, nrev, 0);
    __syncwarp();
    if (nrev > 0) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(X.counters + 0, (unsigned long long)nrev);
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at + (unsigned long long)nrev > X.pool_cap) { if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_POOL); return; }
        int asum = 0;
        for (int k = lane; k < nrev; k += 32) { const int dv = rev_s[nrev - 1 - k]; X.pool[at + k] = dv; asum += dv > 0 ? dv : -dv - 1; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
        *doff = (uint32_t)at; *dcnt = nrev; *dasum = asum;
    }
    __syncwarp();
}

// A forward alignment (not OPTIMAL) over a window of up to 100 x 100 bases with N + M <= breaklen, one WARP: what the
// thread-per-job kernel does for such a window with one thread (k_ex_wave1_tpj: the full matrix is the engine's answer while no
// cell falls 3*breaklen below the running high score, checked by the same sufficient condition) laid out as in
// eng_forced_systolic: lane l owns the strip of columns 8l+1 .. 8l+8, handles row t - l at step t and hands the right edge of
// its strip to lane l + 1 by one shuffle.  A 100 x 100 window is 113 steps of one row-strip each instead of 1313 row-strips of
// one thread — the largest windows were the tail of the thread-per-job kernel (0.6 ms of a 5 Mbp pair, 9 % of the SMs busy) —
// and from 64 x 64 on it is also fewer instructions (60 per step against 8.5 per row-strip and lane).  The traceback words
// ([row][strip], 8 cells each) stay in the warp's score ring (row 7, the reversed deltas, is left out), the walk back runs on
// shared memory.  Returns false when the condition fails: the caller runs the general engine.
template <class Cfg>
__device__ __noinline__ bool eng_mid_full(const Eng &E, const ExShared &X, const ExJobDesc &d, int N, int M, uint32_t *doff, int32_t *dcnt, int32_t *dasum)
{
    const int lane = E.lane;
    int32_t *ring = eng_warp_smem<Cfg>();
    int32_t *rev_s = ring + Cfg::REV_OFF;
    const PackedView &Q = d.dir ? X.QR : X.QF;
    const int64_t Apos0 = d.Abase + d.eA - 1, Bpos0 = d.Bbase + d.eB - 1;       // 0-based positions of window row 1 / column 1
    const int strips = (M + 7) >> 3;
    auto tbword = [&](int w) -> uint2 * { int off = w * 8; if (off >= Cfg::REV_OFF * 4) off += Cfg::REV_N * 4; return (uint2 *)((char *)ring + off); };
    *doff = 0; *dcnt = 0; *dasum = 0;
    const int j0 = lane * 8;
    const bool have = lane < strips;
    const unsigned qn = have ? tpj_query_nibbles(Q, Bpos0 + j0, M - j0) : 0x44444444u;
    int uI[8], ug[8];
#pragma unroll
    for (int c = 0; c < 8; c++) { uI[c] = SYS_NEG; ug[c] = SYS_NEG; }
    int prev_bmc = SYS_NEG;
    int oD = SYS_NEG, oH = SYS_NEG, oC = SYS_NEG;           // right edge of the row this lane finished last: (D, max(I+1, M+2), cell maximum)
    uint64_t aw = 0; uint32_t ax = 0;
    int slack = INT32_MAX;
    const int steps = N + strips;
    __syncwarp();
    for (int t = 0; t < steps; t++) {
        int lD = __shfl_up_sync(0xffffffffu, oD, 1), hl = __shfl_up_sync(0xffffffffu, oH, 1), bmc = __shfl_up_sync(0xffffffffu, oC, 1);
        const int i = t - lane;
        const bool act = have && i >= 0 && i <= N;
        if (lane == 0 && act) {
            if (i == 0) { lD = SYS_NEG; hl = 2; bmc = 2; }                           // cell (0,0) = MAT 0
            else { const int v = 4 * (PMN_OPEN_GAP_SCORE + PMN_CONT_GAP_SCORE * (i - 1)) + PMN_ST_INS; lD = SYS_NEG; hl = v; bmc = v; }
        }
        if (act) {
            unsigned an = 8;
            if (i >= 1) {
                if (((i - 1) & 31) == 0) { aw = pmn_window64(X.R.w, Apos0 + i - 1); ax = X.R.has_x ? pmn_xwindow32(X.R.xm, Apos0 + i - 1) : 0u; }
                an = (unsigned)(aw >> 62) | ((ax >> 31) << 3);
                aw <<= 2; ax <<= 1;
            }
            const unsigned x = (an * 0x11111111u) ^ qn;
            const int rowoff = -6 * (i + j0 + 8);
            int dmc = prev_bmc;
            prev_bmc = bmc;
            unsigned t0 = 0, t1 = 0; int mc = 0;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const int sc = (x & (0xfu << (4 * c))) ? 4 * PMN_BAD_SCORE : 4 * PMN_GOOD_SCORE;
                const int mD = __viaddmax_s32(lD, 4 * PMN_CONT_GAP_SCORE + PMN_ST_DEL, hl + 4 * PMN_OPEN_GAP_SCORE);
                const int mI = __viaddmax_s32(uI[c], 4 * PMN_CONT_GAP_SCORE + PMN_ST_INS, ug[c] + 4 * PMN_OPEN_GAP_SCORE);
                const int mM = dmc + sc;
                dmc = __viaddmax_s32(uI[c], PMN_ST_INS, ug[c]);
                const int vD = mD & ~3, vI = mI & ~3, vM2 = (mM & ~3) + PMN_ST_MAT;
                hl = __viaddmax_s32(vI, PMN_ST_INS, vM2);
                mc = max(hl, vD);
                uI[c] = vI; ug[c] = max(vD, vM2);
                lD = vD;
                slack = __viaddmin_s32(mc, rowoff, slack);
                const unsigned tbits = ((unsigned)mD & 3u) | (((unsigned)mI & 3u) << 2) | (((unsigned)mM & 3u) << 4);
                if (c < 4) t0 |= tbits << (8 * c); else t1 |= tbits << (8 * (c - 4));
            }
            *tbword(i * strips + lane) = make_uint2(t0, t1);
            oD = lD; oH = hl; oC = mc;
        }
    }
    slack = __reduce_min_sync(0xffffffffu, slack);
    if (slack < 3 - 4 * PMN_GOOD_SCORE * X.breaklen) return false;              // the same bound as k_ex_wave1_tpj
    int ms_fin;
    {   // the finish cell (N, M) sits in column (M-1) & 7 of lane (M-1) >> 3
        const int c = (M - 1) & 7;
        int mcf = 0;
#pragma unroll
        for (int cc = 0; cc < 8; cc++) if (cc == c) mcf = __viaddmax_s32(uI[cc], PMN_ST_INS, ug[cc]);
        ms_fin = __shfl_sync(0xffffffffu, mcf, (M - 1) >> 3) & 3;
    }
    if (lane == 0) { atomicAdd(X.counters + 2, (unsigned long long)((N + 1) * (M + 1) - 1)); atomicAdd(X.counters + 3, 1ull); }
    __syncwarp();
    int nrev = 0;
    if (lane == 0) {
        int i = N, j = M, st = ms_fin, pending = 0, run = 0;
        while (i > 0 || j > 0) {
            unsigned b;
            if (j == 0) b = (unsigned)(i == 1 ? PMN_ST_MAT : PMN_ST_INS) << 2;          // column 0: INS from above; cell (1,0) comes from (0,0) MAT
            else {
                const uint2 wv = *tbword(i * strips + ((j - 1) >> 3));
                const int c = (j - 1) & 7;
                b = ((c < 4 ? wv.x : wv.y) >> (8 * (c & 3))) & 0x3fu;
            }
            if (st == PMN_ST_MAT) { run++; st = (b >> 4) & 3; i--; j--; }
            else {
                if (pending) rev_s[nrev++] = pending * (run + 1);
                run = 0;
                if (st == PMN_ST_INS) { pending = 1; st = (b >> 2) & 3; i--; }
                else { pending = -1; st = b & 3; j--; }
            }
        }
        if (pending) rev_s[nrev++] = pending * (run + 1);
    }
    nrev = __shfl_sync(0xffffffffu, nrev, 0);
    __syncwarp();
    if (nrev > 0) {
        unsigned long long at = 0;
        if (lane == 0) at = atomicAdd(X.counters + 0, (unsigned long long)nrev);
        at = __shfl_sync(0xffffffffu, at, 0);
        if (at + (unsigned long long)nrev > X.pool_cap) { if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_POOL); return true; }
        int asum = 0;
        for (int k = lane; k < nrev; k += 32) { const int dv = rev_s[nrev - 1 - k]; X.pool[at + k] = dv; asum += dv > 0 ? dv : -dv - 1; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) asum += __shfl_xor_sync(0xffffffffu, asum, o);
        *doff = (uint32_t)at; *dcnt = nrev; *dasum = asum;
    }
    __syncwarp();
    return true;
}

// returns false when the alignment is too wide for this kernel's layout (never for CfgBig)
template <class Cfg>
__device__ __forceinline__ bool wave1_run(const Eng &E, const ExShared &X, const ExJobDesc &d, int *reached_out = nullptr, bool allow_mid = true)
{
    if (reached_out) *reached_out = 1;
    {
        const int N = d.tA - d.eA + 1, M = d.tB - d.eB + 1;
        if (d.m_o == PMN_FORWARD_ALIGN && N >= 1 && M >= 1 && N <= 31 && M <= 31 && X.breaklen >= 134) {
            uint32_t doff; int32_t dcnt, dasum;
            eng_small_full<Cfg>(E, X, d, N, M, &doff, &dcnt, &dasum);
            if (E.lane == 0) {
                ExJob r; r.endA = d.tA; r.endB = d.tB; r.dcnt = dcnt; r.doff = doff; r.reached = 1; r.valid = 1; r.asum = dasum; r.target = d.target;
                X.jobs[d.g] = r;
            }
            return true;
        }
    }
    if (Cfg::MAXK >= 8 && allow_mid) {        // not for what the thread-per-job kernel handed back: the same condition failed there
        const int N = d.tA - d.eA + 1, M = d.tB - d.eB + 1;
        if (d.m_o == PMN_FORWARD_ALIGN && N >= 1 && M >= 1 && N <= TPJ_MAXDIM && M <= TPJ_MAXDIM && N + M <= X.breaklen) {
            uint32_t doff; int32_t dcnt, dasum;
            if (eng_mid_full<Cfg>(E, X, d, N, M, &doff, &dcnt, &dasum)) {
                if (E.lane == 0) {
                    ExJob r; r.endA = d.tA; r.endB = d.tB; r.dcnt = dcnt; r.doff = doff; r.reached = 1; r.valid = 1; r.asum = dasum; r.target = d.target;
                    X.jobs[d.g] = r;
                }
                return true;
            }
        }
    }
    int64_t targetA = d.tA, targetB = d.tB; unsigned m_o = (unsigned)d.m_o;
    int overflow = 0;
    if (targetA - d.eA + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetA = d.eA + PMN_MAX_ALIGNMENT_LENGTH - 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (targetB - d.eB + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetB = d.eB + PMN_MAX_ALIGNMENT_LENGTH - 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    uint32_t doff; int32_t dcnt, dasum;
    int reached = align_engine<Cfg>(E, d.Abase, d.eA, targetA, d.dir ? X.QR : X.QF, d.Bbase, d.eB, targetB, m_o, &doff, &dcnt, &dasum);
    if (reached < 0) return false;
    if (reached && overflow) reached = 0;
    if (reached_out) *reached_out = reached;
    if (E.lane == 0) {
        ExJob r; r.endA = (int32_t)targetA; r.endB = (int32_t)targetB; r.dcnt = dcnt; r.doff = doff; r.reached = reached; r.valid = 1; r.asum = dasum; r.target = d.target;
        X.jobs[d.g] = r;
    }
    return true;
}

// The backward extension of the first match of cluster k, ahead of the stitcher: every cluster that no end job reaches starts a
// new alignment there (unless it turns out to be shadowed).  The search runs in the largest window extendBackward can ask
// for (towards the sequence starts, OPTIMAL); the stitcher takes the result when its own window, bounded by the
// target alignment, contains every cell this search evaluated (ext_i, ext_j) -- then the two runs are the same
// cell for cell -- and runs the engine itself otherwise.
__device__ __noinline__ void back_job(const Eng &E, const ExShared &X, int k)
{
    const ExCluster c = X.cl[k];
    const ExSynteny S = X.syn[c.syn];
    const int g = c.mfirst;
    const int64_t sA = X.mA[g], sB = X.mB[g];
    // A search that finds nothing dies within breaklen anti-diagonals of its best cell; one that is still alive
    // further out is following homologous sequence towards an earlier alignment and will most likely be
    // merged into it by the stitcher, which this kernel cannot know.  So the window is capped, and a search
    // that touches the cap is left to the stitcher.
    const int64_t cap = X.breaklen + 128;
    const int64_t fullN = sA < PMN_MAX_ALIGNMENT_LENGTH ? sA : PMN_MAX_ALIGNMENT_LENGTH, fullM = sB < PMN_MAX_ALIGNMENT_LENGTH ? sB : PMN_MAX_ALIGNMENT_LENGTH;
    const int64_t Ns = fullN < cap ? fullN : cap, Ms = fullM < cap ? fullM : cap;
    int64_t targetA = sA - Ns + 1, targetB = sB - Ms + 1;
    const PackedView &Q = c.dir ? X.QR : X.QF; const int64_t Bbase = c.dir ? S.BbaseR : S.BbaseF;
    int ext_i = INT32_MAX, ext_j = INT32_MAX;
    align_engine<CfgBig>(E, S.Abase, sA, targetA, Q, Bbase, sB, targetB, PMN_BACKWARD_SEARCH | PMN_OPTIMAL_BIT, nullptr, nullptr, nullptr, &ext_i, &ext_j);
    if ((Ns != fullN && ext_i >= Ns) || (Ms != fullM && ext_j >= Ms)) return;        // clipped by the cap: not the search the stitcher would run
    int64_t eA = sA, eB = sB; uint32_t doff = 0; int32_t dcnt = 0, dasum = 0;
    align_engine<CfgBig>(E, S.Abase, targetA, eA, Q, Bbase, targetB, eB, PMN_FORCED_FORWARD_ALIGN, &doff, &dcnt, &dasum);
    if (E.lane == 0) {
        ExBack b; b.fA = (int32_t)targetA; b.fB = (int32_t)targetB; b.ext_i = ext_i; b.ext_j = ext_j; b.doff = doff; b.dcnt = dcnt; b.asum = dasum; b.valid = 1;
        X.back[k] = b;
    }
    __syncwarp();
}

// Wave 1, big kernel: the cluster-end extensions (break-length searches, bands up to hundreds of cells)
// and what the small kernel handed over.
__global__ void __launch_bounds__(EX_WARPS_PER_BLOCK * 32, 3) k_ex_wave1_big(ExShared X, int pass)
{
    Eng E = make_eng(X, nullptr); E.kid = 1;
    const int lane = E.lane;
    // pass 0 (runs beside the thread-per-job kernel): the heads, cluster ends + the windows k_ex_jobdesc found too large;
    // pass 1 (after it): the windows the thread-per-job kernel handed back.
    // Backward extensions (round 1: a wave of its own behind wave 1, 0.2-0.3 ms of every pair): a cluster no end job aims at
    // (aimers == 0) starts an alignment whatever wave 1 finds, so its backward extension runs here, first — two engine runs, as
    // long as the longest jobs of this pass.  A cluster that is aimed at starts one only when every end job aiming at it fails
    // to reach it: the warp whose end job is the last of them to fail (aimfail == aimers) runs the backward extension at once.
    const unsigned long long nH = (pass == 0 && X.do_extend) ? (unsigned long long)X.nC : 0ull;
    const unsigned long long nA = nH, nO = X.counters[pass ? 12 : 7];
    const int32_t *list = pass ? X.overflow2 : X.overflow;
    for (;;) {
        unsigned long long k = 0;
        if (lane == 0) k = atomicAdd(X.counters + (pass ? 13 : 5), 1ull);
        k = __shfl_sync(0xffffffffu, k, 0);
        if (k >= nH + nA + nO) break;
        if (k < nH) { if (X.aimers[k] == 0) { E.kid = 3; back_job(E, X, (int)k); E.kid = 1; } continue; }
        k -= nH;
        if (k >= nA) { const ExJobDesc d = X.descB[list[k - nA]]; if (d.m_o >= 0) wave1_run<CfgBig>(E, X, d, nullptr, pass == 0); continue; }
        const ExJobDesc d = X.descA[k];
        if (d.m_o < 0) continue;
        int reached = 1;
        wave1_run<CfgBig>(E, X, d, &reached);
        if (reached) continue;
        const ExCluster c = X.cl[k];
        const ExSynteny S = X.syn[c.syn];
        if (d.target < 0 || d.target >= S.cfirst + S.nC) continue;
        int last = 0;
        if (lane == 0) last = atomicAdd(X.aimfail + d.target, 1) + 1 == X.aimers[d.target];
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) { E.kid = 3; back_job(E, X, d.target); E.kid = 1; }
    }
}


// bin counters -> bin starts, total number of thread-per-job alignments
__global__ void __launch_bounds__(512) k_ex_tbinscan(ExShared X)
{
    __shared__ uint32_t sh[512];
    const int t = threadIdx.x;
    const uint32_t v = t < TPJ_BINS ? X.tbin[t] : 0u;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < 512; o <<= 1) {
        const uint32_t a = t >= o ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += a;
        __syncthreads();
    }
    if (t < TPJ_BINS) X.tbin[TPJ_BINS + t] = sh[t] - v;
    if (t == 511) X.counters[10] = sh[t];
}

__global__ void __launch_bounds__(256) k_ex_tscatter(ExShared X)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= X.nM) return;
    const uint2 k = X.tkey[g];
    if (k.x != ~0u) X.tsorted[X.tbin[TPJ_BINS + k.x] + k.y] = (int32_t)g;
}

// nibbles of the 8 query bases at 0-based positions p .. p+7 (nibble c = base p+c): code, or 4 when the base
// matches nothing or lies behind the `left` bases the window still has
__device__ __forceinline__ unsigned tpj_query_nibbles(const PackedView &Q, int64_t p, int left)
{
    const uint32_t w = (uint32_t)(pmn_window64(Q.w, p) >> 48);                 // 8 bases, base c in bits 15-2c, 14-2c
    const uint32_t xm = Q.has_x ? (pmn_xwindow32(Q.xm, p) >> 24) : 0u;          // base c in bit 7-c
    unsigned r = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const unsigned code = (w >> (14 - 2 * c)) & 3u;
        const bool bad = c >= left || ((xm >> (7 - c)) & 1u);
        r |= (bad ? 4u : code) << (4 * c);
    }
    return r;
}

// Wave 1, thread-per-job kernel: every match -> next match alignment whose window passes tpj_fits().
//
// Why the full matrix is the engine's answer (oracle/pmn_oracle.c align_engine, FORWARD_ALIGN): with
// N + M <= breaklen the break-length stop cannot fire, and the band is the whole anti-diagonal as long as
// no cell ever falls 3*breaklen below the running high score.  The running high score after anti-diagonal
// d is at most 3*floor(d/2); every thread checks the sufficient condition
//     cell maximum >= 3*(i + j0 + 8)/2 - 3*breaklen     (scaled by 4, state bits allowed for)
// on all its cells and hands the alignment to the general engine (big kernel) when it fails.  Then the
// alignment reaches its target, finishes there, and the traceback is that of the untrimmed matrix.
//
// Per thread: the matrix is walked in strips of 8 columns, top to bottom; of the previous row the strip
// keeps per column I and max(D, M+2) (scores x4, the two low bits carry the state a value came from, as in
// eng_run_reg), the cells right of a strip reach the next strip through a 16-byte record per row in
// global scratch, and every row of a strip stores its 8 traceback bytes with one 8-byte store
// (lane-interleaved: a warp writes 256 contiguous bytes).
__global__ void __launch_bounds__(128, TPJ_BLOCKS_PER_SM) k_ex_wave1_tpj(ExShared X)
{
    const int lane = threadIdx.x & 31;
    const size_t slot = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    uint8_t *sbase = X.tscratch + slot * (size_t)TPJ_SLOT_BYTES;
    uint2 *tb = (uint2 *)sbase + lane;                                   // [word][lane]
    int4 *bnd = (int4 *)(sbase + (size_t)32 * TPJ_TB_WORDS * 8) + lane;  // [row][lane]
    const unsigned nT = (unsigned)X.counters[10];
    const int min_slack = 3 - 4 * PMN_GOOD_SCORE * X.breaklen;           // see above: 4 * (GOOD_SCORE * breaklen), plus the state bits
    for (;;) {
        unsigned w = 0;
        if (lane == 0) w = (unsigned)atomicAdd(X.counters + 11, 1ull);
        w = __shfl_sync(0xffffffffu, w, 0);
        if ((unsigned long long)w * 32ull >= nT) break;
        const unsigned k = w * 32 + lane;
        const bool have = k < nT;
        const int g = have ? X.tsorted[k] : 0;
        int N = 0, M = 0, dir = 0, tA = 0, tB = 0; int64_t Ap = 0, Bp = 0;
        if (have) {
            const ExJobDesc d = X.descB[g];
            N = d.tA - d.eA + 1; M = d.tB - d.eB + 1; dir = d.dir; tA = d.tA; tB = d.tB;
            Ap = d.Abase + d.eA - 1; Bp = d.Bbase + d.eB - 1;            // 0-based positions of window row 1 / column 1
        }
        const PackedView &Q = dir ? X.QR : X.QF;
        const int strips = (M + TPJ_W - 1) / TPJ_W;
        const int Nmax = __reduce_max_sync(0xffffffffu, N), Smax = __reduce_max_sync(0xffffffffu, strips);
        const int stride = Smax * (Nmax + 1) <= TPJ_TB_WORDS ? Nmax + 1 : N + 1;     // lanes of similar jobs share rows
        int slack = INT32_MAX, ms_fin = PMN_ST_MAT;
        for (int s = 0; s < strips; s++) {
            const int j0 = s * TPJ_W;
            const unsigned qn = tpj_query_nibbles(Q, Bp + j0, M - j0);
            int uI[TPJ_W], ug[TPJ_W];
#pragma unroll
            for (int c = 0; c < TPJ_W; c++) { uI[c] = TPJ_NEG; ug[c] = TPJ_NEG; }
            int prev_bmc = TPJ_NEG;                 // cell maximum (with state) of (i-1, j0)
            uint64_t aw = 0; uint32_t ax = 0;
            const bool last = s + 1 == strips;
            int4 nb = make_int4(TPJ_NEG, 2, 2, 0);  // boundary record of row 0 in strip 0: cell (0,0) = MAT 0
            int4 nb2 = nb;                          // the record after it: loads run two rows ahead of their use
            if (s) { nb = bnd[0]; if (N >= 1) nb2 = bnd[32]; }
            uint2 *trow = tb + (size_t)s * stride * 32;
            for (int i = 0; i <= N; i++) {
                const int lD0 = nb.x, hl0 = nb.y, bmc = nb.z;
                if (s) { nb = nb2; if (i + 2 <= N) nb2 = bnd[(size_t)(i + 2) * 32]; }
                else { const int v = 4 * (PMN_OPEN_GAP_SCORE + PMN_CONT_GAP_SCORE * i) + PMN_ST_INS; nb = make_int4(TPJ_NEG, v, v, 0); }   // cell (i+1, 0): INS only
                unsigned an = 8;
                if (i >= 1) {
                    if (((i - 1) & 31) == 0) { aw = pmn_window64(X.R.w, Ap + i - 1); ax = X.R.has_x ? pmn_xwindow32(X.R.xm, Ap + i - 1) : 0u; }
                    an = (unsigned)(aw >> 62) | ((ax >> 31) << 3);
                    aw <<= 2; ax <<= 1;
                }
                const unsigned x = (an * 0x11111111u) ^ qn;
                const int rowoff = -6 * (i + j0 + TPJ_W);
                int lD = lD0, hl = hl0, dmc = prev_bmc;
                prev_bmc = bmc;
                unsigned t0 = 0, t1 = 0;
#pragma unroll
                for (int c = 0; c < TPJ_W; c++) {
                    const int sc = (x & (0xfu << (4 * c))) ? 4 * PMN_BAD_SCORE : 4 * PMN_GOOD_SCORE;
                    const int mD = __viaddmax_s32(lD, 4 * PMN_CONT_GAP_SCORE + PMN_ST_DEL, hl + 4 * PMN_OPEN_GAP_SCORE);
                    const int mI = __viaddmax_s32(uI[c], 4 * PMN_CONT_GAP_SCORE + PMN_ST_INS, ug[c] + 4 * PMN_OPEN_GAP_SCORE);
                    const int mM = dmc + sc;
                    dmc = __viaddmax_s32(uI[c], PMN_ST_INS, ug[c]);        // cell maximum of (i-1, j): the diagonal of the next column
                    const int vD = mD & ~3, vI = mI & ~3, vM2 = (mM & ~3) + PMN_ST_MAT;
                    hl = __viaddmax_s32(vI, PMN_ST_INS, vM2);
                    const int mc = max(hl, vD);
                    uI[c] = vI; ug[c] = max(vD, vM2);
                    lD = vD;
                    slack = __viaddmin_s32(mc, rowoff, slack);
                    const unsigned tbits = ((unsigned)mD & 3u) | (((unsigned)mI & 3u) << 2) | (((unsigned)mM & 3u) << 4);
                    if (c < 4) t0 |= tbits << (8 * c); else t1 |= tbits << (8 * (c - 4));
                    if (c == TPJ_W - 1 && !last) bnd[(size_t)i * 32] = make_int4(vD, hl, mc, 0);
                }
                trow[(size_t)i * 32] = make_uint2(t0, t1);
            }
            if (last) {                               // state of the finish cell (N, M)
                const int c = (M - 1) & (TPJ_W - 1);
                int mcf = 0;
#pragma unroll
                for (int cc = 0; cc < TPJ_W; cc++) if (cc == c) mcf = __viaddmax_s32(uI[cc], PMN_ST_INS, ug[cc]);
                ms_fin = mcf & 3;
            }
        }
        const bool ok = have && slack >= min_slack;
        if (have && !ok) X.overflow2[atomicAdd(X.counters + 12, 1ull)] = g;   // the general engine decides
        // walk back from (N, M); the reversed deltas go to this thread's boundary rows (free by now)
        int nrev = 0, asum = 0;
        int32_t *rev = (int32_t *)bnd;                 // entry e at rev[(e >> 2) * 128 + (e & 3)]
        if (ok) {
            int i = N, j = M, st = ms_fin, pending = 0, run = 0;
            while (i > 0 || j > 0) {
                unsigned b;
                if (j == 0) b = (unsigned)(i == 1 ? PMN_ST_MAT : PMN_ST_INS) << 2;          // column 0: INS from above; cell (1,0) comes from (0,0) MAT
                else {
                    const uint2 wv = tb[((size_t)((j - 1) >> 3) * stride + i) * 32];
                    const int c = (j - 1) & 7;
                    b = ((c < 4 ? wv.x : wv.y) >> (8 * (c & 3))) & 0x3fu;
                }
                if (st == PMN_ST_MAT) { run++; st = (b >> 4) & 3; i--; j--; }
                else {
                    if (pending) { const int dv = pending * (run + 1); rev[(nrev >> 2) * 128 + (nrev & 3)] = dv; nrev++; asum += dv > 0 ? dv : -dv - 1; }
                    run = 0;
                    if (st == PMN_ST_INS) { pending = 1; st = (b >> 2) & 3; i--; }
                    else { pending = -1; st = b & 3; j--; }
                }
            }
            if (pending) { const int dv = pending * (run + 1); rev[(nrev >> 2) * 128 + (nrev & 3)] = dv; nrev++; asum += dv > 0 ? dv : -dv - 1; }
        }
        // delta pool: one atomic per warp
        int incl = nrev;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        unsigned long long at = 0;
        if (total > 0) {
            if (lane == 31) at = atomicAdd(X.counters + 0, (unsigned long long)total);
            at = __shfl_sync(0xffffffffu, at, 31);
            if (at + (unsigned long long)total > X.pool_cap) { if (lane == 0) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_POOL); continue; }
        }
        const unsigned long long mine = at + (unsigned long long)(incl - nrev);
        for (int e = 0; e < nrev; e++) { const int r = nrev - 1 - e; X.pool[mine + e] = rev[(r >> 2) * 128 + (r & 3)]; }
        if (ok) {
            ExJob r; r.endA = tA; r.endB = tB; r.dcnt = nrev; r.target = -1; r.doff = nrev ? (uint32_t)mine : 0u; r.reached = 1; r.valid = 1; r.asum = asum;
            X.jobs[g] = r;
        }
        unsigned long long cells = ok ? (unsigned long long)((N + 1) * (M + 1) - 1) : 0ull;
        cells = __reduce_add_sync(0xffffffffu, (unsigned)cells);
        const unsigned njobs = __popc(__ballot_sync(0xffffffffu, ok));
        if (lane == 0) { atomicAdd(X.counters + 2, cells); atomicAdd(X.counters + 3, (unsigned long long)njobs); }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------ between wave 1 and the stitcher

// per job: delta count (for the prefix that places range deltas), the key of the lastP scan,
// and the per-cluster "some inner job failed" flag
__global__ void __launch_bounds__(256) k_ex_jobmeta(const ExJob *__restrict__ jobs, const int32_t *__restrict__ mcl, const ExCluster *__restrict__ cl,
                                                   const uint32_t *__restrict__ pstart, const uint32_t *__restrict__ ppos, int64_t nm,
                                                   uint32_t *__restrict__ dcnt, long long *__restrict__ pkey, uint8_t *__restrict__ anyfail)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > nm) return;
    if (g == nm) { dcnt[g] = 0; return; }
    const ExJob j = jobs[g];
    const int k = mcl[g];
    const ExCluster c = cl[k];
    const bool inner = (int)g != c.mfirst + c.nm - 1;
    dcnt[g] = j.valid ? (uint32_t)j.dcnt : 0u;
    const long long piece = (long long)(ppos[g] + pstart[g]) - 1;
    const long long low = (inner && j.valid && j.dcnt > 0) ? g + 1 : 0;      // job index + 1 of a job that has deltas
    pkey[g] = (piece << 32) | low;
    if (inner && j.valid && !j.reached) anyfail[k] = 1;
}

// P (reference position consumed through the last indel) left behind by the latest job in
// [g0, g] that has deltas; `fallback` when there is none
__device__ __forceinline__ int range_last_P(const ExShared &X, int g0, int g, int fallback)
{
    const int j = (int)(X.lastP[g] & 0xffffffffll) - 1;
    if (j < g0) return fallback;
    return X.mA[j] + X.mL[j] - 2 + X.jobs[j].asum;
}

__global__ void __launch_bounds__(256) k_ex_csum(ExShared X, ExCSum *__restrict__ out)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= X.nC) return;
    const ExCluster c = X.cl[k];
    const int f = c.mfirst, l = c.mfirst + c.nm - 1;
    ExCSum s;
    s.sA0 = X.mA[f]; s.sB0 = X.mB[f]; s.len0 = X.mL[f];
    s.sAl = X.mA[l]; s.sBl = X.mB[l]; s.eAl = X.mA[l] + X.mL[l] - 1; s.eBl = X.mB[l] + X.mL[l] - 1;
    s.mfirst = f; s.nm = c.nm; s.dir = c.dir; s.anyfail = X.anyfail[k];
    const ExJob j = X.jobs[l];
    s.endA = j.endA; s.endB = j.endB; s.e_reached = j.reached; s.e_dcnt = j.valid ? j.dcnt : -1; s.e_asum = j.asum; s.target = j.target; s.e_doff = j.doff;
    s.bulk_cnt = (int32_t)(X.dcnt_ex[l] - X.dcnt_ex[f]);
    s.bulk_P = c.nm > 1 ? range_last_P(X, f, l - 1, -1) : -1;
    out[k] = s;
}

// ------------------------------------------------------------------------------------ E3: stitch

// The stitcher is one warp walking a dependent chain, so what it costs is the number of global-memory round trips per
// cluster.  The first ST_AL_CACHE alignments of the synteny are mirrored in shared memory (the shadow test and the reverse
// target search read all of them for every cluster that starts an alignment), and the cluster records are read a window of
// 32 at a time (k_ex_stitch).
#define ST_AL_CACHE 64
struct Stitch {
    const Eng *E; const ExSynteny *S; ExAlign *al; ExAlign *sal; ExNode *nodes; int nAl, nNodes; bool fail;
    ExAlign cur; int cur_slot;          // the alignment being grown lives in registers (identical on all lanes)
};

__device__ __forceinline__ ExAlign st_al(const Stitch &T, int i) { return i < ST_AL_CACHE ? T.sal[i] : T.al[i]; }
__device__ __forceinline__ void st_flush(Stitch &T)
{
    if (T.cur_slot >= 0 && T.E->lane == 0) { T.al[T.cur_slot] = T.cur; if (T.cur_slot < ST_AL_CACHE) T.sal[T.cur_slot] = T.cur; }
    __syncwarp();
}
__device__ __forceinline__ void st_load(Stitch &T, int slot) { T.cur = st_al(T, slot); T.cur_slot = slot; }

// append a node to the alignment held in registers
__device__ void cur_append(Stitch &T, int type, uint32_t a, int cnt_field, int b, int ndeltas)
{
    if (T.nNodes >= T.S->nodecap) { T.fail = true; return; }
    const int nd = T.nNodes++;
    if (T.E->lane == 0) {
        ExNode n; n.type = type; n.a = a; n.cnt = cnt_field; n.b = b; n.outoff = T.cur.ndelta; n.next = -1; n.alslot = T.S->alfirst + T.cur_slot; n.pad = 0;
        T.nodes[nd] = n;
    }
    if (T.cur.head < 0) T.cur.head = nd;
    T.cur.tail = nd; T.cur.ndelta += ndeltas;
}

// extendForward on the alignment in registers from a known result `r`
__device__ __forceinline__ int cur_apply_job(Stitch &T, uint32_t doff, int dcnt, int asum, int endA, int endB, int reached)
{
    if (dcnt > 0) {
        // the first new delta counts from the engine's start column: add the bases the alignment
        // already holds since its last indel; afterwards P = (start - 1) + asum
        cur_append(T, 0, doff, dcnt, T.cur.eA - T.cur.P - 1, dcnt);
        T.cur.P = T.cur.eA - 1 + asum;
    }
    T.cur.eA = endA; T.cur.eB = endB;
    return reached;
}

// extendForward, general: uses wave-1 job g when it computed exactly this extension, else runs the engine
__device__ int cur_extend_forward(Stitch &T, int dirB, int64_t targetA, int64_t targetB, unsigned m_o, int g)
{
    const ExShared &X = *T.E->X;
    ExJob r; bool have = false;
    if (g >= 0) { r = X.jobs[g]; have = r.valid && T.cur.eA == X.mA[g] + X.mL[g] - 1 && T.cur.eB == X.mB[g] + X.mL[g] - 1; }
    if (!have) forward_job(*T.E, *T.S, dirB, T.cur.eA, T.cur.eB, targetA, targetB, m_o, r);
    return cur_apply_job(T, r.doff, r.dcnt, r.asum, r.endA, r.endB, r.reached);
}

// first job in [g, last) that did not reach its target, or `last`
__device__ int st_next_fail(const ExShared &X, int g, int last, int lane)
{
    for (int b = g; b < last; b += 32) {
        const int k = b + lane;
        unsigned bal = __ballot_sync(0xffffffffu, k < last && !X.jobs[k].reached);
        if (bal) return b + __ffs(bal) - 1;
    }
    return last;
}

// getReverseTargetAlignment, 32 candidates per step.  The sequential original walks i = ap-1 .. 0,
// stops at the first "close enough" alignment and otherwise keeps the first strict improvement
// of dist = 2*greater - lesser.
__device__ int st_get_reverse_target(const Stitch &T, int ap, int dirB, int64_t sA, int64_t sB)
{
    const ExShared &X = *T.E->X;
    const int lane = T.E->lane;
    long long dist = sA < sB ? sA : sB;
    int best = -1;
    for (int top = ap - 1; top >= 0; top -= 32) {
        const int i = top - lane;
        bool valid = false, close = false; long long score = 0;
        if (i >= 0) {
            const ExAlign a = st_al(T, i);
            if (a.dirB == dirB && a.eA <= sA && a.eB <= sB) {
                valid = true;
                long long greater, lesser;
                if (sA - a.eA > sB - a.eB) { greater = sA - a.eA; lesser = sB - a.eB; } else { lesser = sA - a.eA; greater = sB - a.eB; }
                close = greater < X.breaklen || lesser * PMN_GOOD_SCORE + (greater - lesser) * PMN_CONT_GAP_SCORE >= 0;
                score = (greater << 1) - lesser;
            }
        }
        const unsigned cb = __ballot_sync(0xffffffffu, close);
        const int F = cb ? __ffs(cb) - 1 : 32;
        long long key = (valid && !close && lane < F) ? score * 32 + lane : LLONG_MAX;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { long long k2 = __shfl_xor_sync(0xffffffffu, key, o); if (k2 < key) key = k2; }
        if (key != LLONG_MAX) { const long long sc = key >> 5; if (sc < dist) { dist = sc; best = top - (int)(key & 31); } }
        if (cb) return top - F;
    }
    return best;
}

__device__ bool st_is_shadowed(const Stitch &T, const ExCSum &c)
{
    const int lane = T.E->lane;
    for (int top = T.nAl - 1; top >= 0; top -= 32) {
        const int i = top - lane;
        bool hit = false;
        if (i >= 0) { const ExAlign a = st_al(T, i); hit = a.dirB == c.dir && a.eA >= c.eAl && a.eB >= c.eBl && a.sA <= c.sA0 && a.sB <= c.sB0; }
        if (__ballot_sync(0xffffffffu, hit)) return true;
    }
    return false;
}

// extendBackward for the freshly created alignment in registers; returns true when it was merged
// into alignment tp (which then becomes the current one)
__device__ int st_extend_backward(Stitch &T, int tp, int dirB)
{
    const ExShared &X = *T.E->X;
    const ExSynteny &S = *T.S;
    const ExAlign a = T.cur;
    int overflow = 0; unsigned m_o = PMN_BACKWARD_SEARCH;
    int64_t targetA, targetB;
    if (tp >= 0) { const ExAlign t = st_al(T, tp); targetA = t.eA; targetB = t.eB; } else { targetA = 1; targetB = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (a.sA - targetA + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetA = a.sA - PMN_MAX_ALIGNMENT_LENGTH + 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (a.sB - targetB + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetB = a.sB - PMN_MAX_ALIGNMENT_LENGTH + 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    const PackedView &Q = dirB ? X.QR : X.QF; const int64_t Bbase = dirB ? S.BbaseR : S.BbaseF;
    int reached = align_engine<CfgBig>(*T.E, S.Abase, a.sA, targetA, Q, Bbase, a.sB, targetB, m_o, nullptr, nullptr, nullptr);
    if (overflow || tp < 0) reached = 0;
    if (reached) {
        // merge: the target alignment is extended (forced) up to this one's start and absorbs it
        T.nAl--;                        // the new alignment was the last one; it held no deltas yet
        st_load(T, tp);
        cur_extend_forward(T, dirB, a.sA, a.sB, PMN_FORCED_FORWARD_ALIGN, -1);
        T.cur.eA += a.eA - a.sA; T.cur.eB += a.eB - a.sB;
    } else {
        int64_t eA = a.sA, eB = a.sB; uint32_t doff; int32_t dcnt, dasum;
        align_engine<CfgBig>(*T.E, S.Abase, targetA, eA, Q, Bbase, targetB, eB, PMN_FORCED_FORWARD_ALIGN, &doff, &dcnt, &dasum);
        if (dcnt > 0) cur_append(T, 0, doff, dcnt, 0, dcnt);
        T.cur.sA = (int32_t)targetA; T.cur.sB = (int32_t)targetB; T.cur.P = (int32_t)targetA - 1 + dasum;
    }
    return reached;
}

// extendClusters for one synteny; every lane runs the same control flow on the same values
__global__ void __launch_bounds__(EX_WARPS_PER_BLOCK * 32) k_ex_stitch(ExShared X, uint8_t *fused)
{
    Eng E = make_eng(X, nullptr); E.kid = 2;
    const int lane = E.lane;
    const int s = blockIdx.x * EX_WARPS_PER_BLOCK + (threadIdx.x >> 5);
    if (s >= X.nS) return;
    const ExSynteny S = X.syn[s];
    __shared__ __align__(16) ExCSum s_cs[EX_WARPS_PER_BLOCK][32];
    __shared__ __align__(16) ExBack s_back[EX_WARPS_PER_BLOCK][32];
    __shared__ __align__(16) ExAlign s_al[EX_WARPS_PER_BLOCK][ST_AL_CACHE];
    __shared__ uint8_t s_fused[EX_WARPS_PER_BLOCK][32];
    const int wib = threadIdx.x >> 5;
    Stitch T; T.E = &E; T.S = &S; T.al = X.al + S.alfirst; T.sal = s_al[wib]; T.nodes = X.nodes + S.nodefirst; T.nAl = 0; T.nNodes = 0; T.fail = false; T.cur_slot = -1;
    const int c0 = S.cfirst, cend = S.cfirst + S.nC;
    int target_reached = 0, CurrCp = c0, PrevCp = c0, TargetCp = cend;
    bool logic_err = false;
    int wbase = -64;                    // clusters [wbase, wbase + 32) are in the window; only this warp writes their fused flags
#ifdef PMN_STITCH_TIMING            // cycles per section of the loop, printed by the host when the variable of the same name is set
    long long tk_window = 0, tk_shadow = 0, tk_start = 0, tk_match = 0, tk_end = 0, n_cl = 0, n_start = 0, n_nf = 0, n_fwd = 0;
    const long long tk0 = clock64();
#define ST_TICK(acc) { const long long t2__ = clock64(); acc += t2__ - tk; tk = t2__; }
#define ST_COUNT(x) x++
#else
#define ST_TICK(acc)
#define ST_COUNT(x)
#endif
    while (CurrCp < cend && !T.fail && !logic_err) {
#ifdef PMN_STITCH_TIMING
        long long tk = clock64();
        n_cl++;
#endif
        if (CurrCp < wbase || CurrCp >= wbase + 32) {
            __syncwarp();
            wbase = CurrCp;
            const int k = CurrCp + lane;
            if (k < cend) { s_cs[wib][lane] = X.cs[k]; s_back[wib][lane] = X.back[k]; s_fused[wib][lane] = fused[k]; }
            __syncwarp();
        }
        const int wi = CurrCp - wbase;
        const ExCSum c = s_cs[wib][wi];
        const int was_fused = s_fused[wib][wi];
        ST_TICK(tk_window)
        // A fused cluster is stepped over.  Flags are never cleared, so the restart point moves along with it: every later
        // restart would walk over the same fused clusters again and end up where this one does.
        if (X.do_extend && !target_reached && was_fused) {
            // on to the first cluster of the window that is not fused, in one step (only this warp writes the flags, and it
            // mirrors them in s_fused)
            const unsigned m = __ballot_sync(0xffffffffu, lane >= wi && (wbase + lane >= cend || !s_fused[wib][lane]));
            CurrCp = wbase + (m ? __ffs((int)m) - 1 : 32); PrevCp = CurrCp; continue;
        }
        if (!target_reached && X.do_simplify) {
            st_flush(T);
            const bool sh = st_is_shadowed(T, c);
            ST_TICK(tk_shadow)
            if (sh) {
                if (lane == 0) { fused[CurrCp] = 1; s_fused[wib][wi] = 1; }
                __syncwarp();
                CurrCp = ++PrevCp; continue;
            }
        }
        const int last = c.mfirst + c.nm - 1;
        // a cluster that is merged a second time (reached as a target after it was fused) goes job by
        // job, so that every wave-1 job belongs to at most one range node
        const bool allow_bulk = !was_fused;
        // The common case without the loop below: the alignment arrives on the first match of a cluster that was never visited,
        // every match -> next match job of the cluster reached its target and wave 1 ran the cluster-end job from the last match.
        // What the loop would do then is known from the cluster record alone: one range node for the inner jobs, one node for the
        // end job — written as ONE node of type 2 that the flattening kernels take apart — and the new P / end / target.  When
        // that end job reaches the next cluster of the window and the same holds there, the step after this one is known too:
        // lane j takes cluster wbase + j, the run of clusters that hand over to their successor is found with two ballots, P and
        // the delta offsets are prefix operations over the run, and every lane writes the node of its own cluster.
        bool fast = false;
        if (target_reached && !was_fused && !c.anyfail && X.do_extend && c.e_dcnt >= 0 && T.cur.eA == c.sA0 && T.cur.eB == c.sB0) {
            fast = true;
            const int kj = wbase + lane;
            const bool inw = lane >= wi && kj < cend;
            ExCSum cj = c;
            if (inw) cj = s_cs[wib][lane];
            const bool elig = inw && !s_fused[wib][lane] && !cj.anyfail && cj.e_dcnt >= 0;
            bool link = false;
            if (elig && lane < 31 && kj + 1 < cend && cj.e_reached && cj.target == kj + 1) {
                const int nA = s_cs[wib][lane + 1].sA0, nB = s_cs[wib][lane + 1].sB0;
                link = cj.endA == nA && cj.endB == nB;
            }
            const unsigned eligm = __ballot_sync(0xffffffffu, elig);
            const unsigned good = __ballot_sync(0xffffffffu, link) & (eligm >> 1);      // bit j: cluster j hands over to cluster j + 1, which qualifies
            const int e = __ffs((int)(~good & (0xffffffffu << wi))) - 1;                // the run is [wi, e]; bit 31 of good is never set
            const bool mine = lane >= wi && lane <= e;
            const int dn = mine ? cj.bulk_cnt + cj.e_dcnt : 0;
            const bool node = dn > 0;
            const bool sets = mine && (cj.e_dcnt > 0 || (cj.bulk_cnt > 0 && cj.bulk_P >= 0));
            const int pval = cj.e_dcnt > 0 ? cj.eAl - 1 + cj.e_asum : cj.bulk_P;
            const unsigned nodem = __ballot_sync(0xffffffffu, node), setm = __ballot_sync(0xffffffffu, sets);
            const unsigned below = (1u << lane) - 1u;
            const unsigned sb = setm & below;                                           // clusters of the run before mine that leave a new P
            int Pin = __shfl_sync(0xffffffffu, pval, sb ? 31 - __clz((int)sb) : 0);
            if (!sb) Pin = T.cur.P;
            int incl = dn;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int total_dn = __shfl_sync(0xffffffffu, incl, 31);
            const int nnodes = __popc(nodem);
            if (T.nNodes + nnodes > S.nodecap) T.fail = true;
            else {
                if (node) {
                    const int nd = T.nNodes + __popc(nodem & below);
                    ExNode n; n.type = 2; n.a = (uint32_t)cj.mfirst; n.cnt = cj.mfirst + cj.nm - 1; n.b = Pin; n.outoff = T.cur.ndelta + incl - dn; n.next = kj; n.alslot = S.alfirst + T.cur_slot; n.pad = 0;
                    T.nodes[nd] = n;
                    if (cj.bulk_cnt > 0) X.markkey[cj.mfirst] = ((unsigned long long)(cj.mfirst + 1) << 32) | (unsigned)(S.nodefirst + nd + 1);
                }
                if (mine) { fused[kj] = 1; s_fused[wib][lane] = 1; }
                T.nNodes += nnodes;
            }
            if (setm) T.cur.P = __shfl_sync(0xffffffffu, pval, 31 - __clz((int)setm));
            T.cur.ndelta += total_dn;
            T.cur.eA = __shfl_sync(0xffffffffu, cj.endA, e); T.cur.eB = __shfl_sync(0xffffffffu, cj.endB, e);
            TargetCp = __shfl_sync(0xffffffffu, cj.target, e); target_reached = __shfl_sync(0xffffffffu, cj.e_reached, e);
            CurrCp = wbase + e;                     // the last cluster of the run; what follows marks it (again) and moves on from it
            __syncwarp();
        }
        int CurrMp = fast ? c.nm : 0; bool positioned = false;       // positioned: the alignment already ends on the last base of match CurrMp
        while (CurrMp < c.nm && !T.fail) {
            const int g = c.mfirst + CurrMp;
            int gA, gB, gL;                            // match g
            // first and last match come with the cluster record: no global-memory round trip on the common path
            if (CurrMp == 0) { gA = c.sA0; gB = c.sB0; gL = c.len0; }
            else if (CurrMp == c.nm - 1) { gA = c.sAl; gB = c.sBl; gL = c.eAl - c.sAl + 1; }
            else if (!positioned || allow_bulk) { gA = X.mA[g]; gB = X.mB[g]; gL = X.mL[g]; }
            else { gA = gB = gL = 0; }          // positioned and not bulk: the values are not looked at
            if (!positioned) {
                if (target_reached) {
                    if (T.cur.eA != gA || T.cur.eB != gB) {
                        // the alignment was extended onto a later match of this cluster: the first one that starts where
                        // it ends, 32 candidates per step
                        int hit = -1;
                        for (int b = CurrMp + 1; b < c.nm && hit < 0; b += 32) {
                            const int k = b + lane;
                            const unsigned bal = __ballot_sync(0xffffffffu, k < c.nm && X.mA[c.mfirst + k] == T.cur.eA && X.mB[c.mfirst + k] == T.cur.eB);
                            if (bal) hit = b + __ffs(bal) - 1;
                        }
                        if (hit < 0) { logic_err = true; break; }
                        CurrMp = hit; continue;
                    }
                    T.cur.eA += gL - 1; T.cur.eB += gL - 1;
                } else {
                    if (T.nAl >= S.alcap) { logic_err = true; break; }
#ifdef PMN_STITCH_TIMING
                    const long long ts0 = clock64(); n_start++;
#endif
                    st_flush(T);
                    T.cur_slot = T.nAl++;
                    T.cur.dirB = c.dir; T.cur.sA = gA; T.cur.sB = gB; T.cur.eA = gA + gL - 1; T.cur.eB = gB + gL - 1;
                    T.cur.P = gA - 1; T.cur.head = -1; T.cur.tail = -1; T.cur.ndelta = 0; T.cur.live = 1; T.cur.pad0 = T.cur.pad1 = 0;
                    if (X.do_extend || CurrMp != 0) {
                        const int TargetAp = st_get_reverse_target(T, T.cur_slot, c.dir, gA, gB);
                        bool taken = false;
                        if (CurrMp == 0 && X.do_extend) {
                            const ExBack b = s_back[wib][wi];
                            if (b.valid) {
                                // the window extendBackward would search: up to the target alignment's end, or the sequence starts
                                int64_t tA = 1, tB = 1;
                                if (TargetAp >= 0) { const ExAlign t = st_al(T, TargetAp); tA = t.eA; tB = t.eB; }
                                int64_t Nw = gA - tA + 1, Mw = gB - tB + 1;
                                if (Nw > PMN_MAX_ALIGNMENT_LENGTH) Nw = PMN_MAX_ALIGNMENT_LENGTH;
                                if (Mw > PMN_MAX_ALIGNMENT_LENGTH) Mw = PMN_MAX_ALIGNMENT_LENGTH;
                                if (TargetAp < 0 || (b.ext_i < Nw && b.ext_j < Mw)) {
                                    if (b.dcnt > 0) cur_append(T, 0, b.doff, b.dcnt, 0, b.dcnt);
                                    T.cur.sA = b.fA; T.cur.sB = b.fB; T.cur.P = b.fA - 1 + b.asum;
                                    taken = true;
                                }
                            }
                        }
                        if (!taken) st_extend_backward(T, TargetAp, c.dir);
                    }
#ifdef PMN_STITCH_TIMING
                    tk_start += clock64() - ts0;
#endif
                }
            }
            positioned = false;
            if (CurrMp < c.nm - 1) {
                if (allow_bulk && T.cur.eA == gA + gL - 1 && T.cur.eB == gB + gL - 1) {
                    int f, cnt, Pn;
                    if (CurrMp == 0 && !c.anyfail) { f = last; cnt = c.bulk_cnt; Pn = c.bulk_P; }
                    else {
                        ST_COUNT(n_nf);
                        f = c.anyfail ? st_next_fail(X, g, last, lane) : last;
                        cnt = f > g ? (int)(X.dcnt_ex[f] - X.dcnt_ex[g]) : 0;
                        Pn = f > g ? range_last_P(X, g, f - 1, -1) : -1;
                    }
                    if (f > g) {
                        // all jobs [g, f) reached their targets: absorb matches g+1 .. f at once
                        if (cnt > 0) {
                            const int nd_before = T.nNodes;
                            cur_append(T, 1, (uint32_t)g, f, T.cur.P, cnt);
                            if (lane == 0 && !T.fail) X.markkey[g] = ((unsigned long long)(g + 1) << 32) | (unsigned)(S.nodefirst + nd_before + 1);
                            if (Pn >= 0) T.cur.P = Pn;
                        }
                        if (f == last) { T.cur.eA = c.eAl; T.cur.eB = c.eBl; }
                        else { T.cur.eA = X.mA[f] + X.mL[f] - 1; T.cur.eB = X.mB[f] + X.mL[f] - 1; }
                        CurrMp = f - c.mfirst; positioned = true; target_reached = 1;
                        continue;
                    }
                }
                ST_COUNT(n_fwd);
                target_reached = cur_extend_forward(T, c.dir, X.mA[g + 1], X.mB[g + 1], PMN_FORWARD_ALIGN, g);
            } else if (X.do_extend) {
#ifdef PMN_STITCH_TIMING
                const long long te0 = clock64();
#endif
                if (c.e_dcnt >= 0 && T.cur.eA == c.eAl && T.cur.eB == c.eBl) {
                    TargetCp = c.target;
                    target_reached = cur_apply_job(T, c.e_doff, c.e_dcnt, c.e_asum, c.endA, c.endB, c.e_reached);
                } else {
                    int64_t targetA = S.lenA, targetB = S.lenB; unsigned m_o = PMN_FORWARD_ALIGN;
                    TargetCp = get_forward_target_cluster(X, CurrCp, cend, targetA, targetB);
                    if (TargetCp == cend) m_o |= PMN_OPTIMAL_BIT;
                    target_reached = cur_extend_forward(T, c.dir, targetA, targetB, m_o, -1);
                }
#ifdef PMN_STITCH_TIMING
                tk_end += clock64() - te0;
#endif
            }
            CurrMp++;
        }
#ifdef PMN_STITCH_TIMING
        tk_match += clock64() - tk;
#endif
        if (TargetCp == cend) target_reached = 0;
        if (lane == 0) { fused[CurrCp] = 1; s_fused[wib][CurrCp - wbase] = 1; }
        __syncwarp();
        if (!target_reached) CurrCp = ++PrevCp; else CurrCp = TargetCp;
    }
    st_flush(T);
    if (lane == 0) {
        X.syn_nal[s] = T.nAl; X.syn_nal[X.nS + s] = T.nNodes;
#ifdef PMN_STITCH_TIMING
        atomicMax(X.counters + 16, (unsigned long long)(clock64() - tk0));
        atomicAdd(X.counters + 17, (unsigned long long)tk_window); atomicAdd(X.counters + 18, (unsigned long long)tk_shadow);
        atomicAdd(X.counters + 19, (unsigned long long)tk_start); atomicAdd(X.counters + 20, (unsigned long long)tk_match);
        atomicAdd(X.counters + 21, (unsigned long long)tk_end); atomicAdd(X.counters + 22, (unsigned long long)n_cl);
        atomicAdd(X.counters + 23, (unsigned long long)n_start); atomicAdd(X.counters + 24, (unsigned long long)n_nf); atomicAdd(X.counters + 25, (unsigned long long)n_fwd);
#endif
        if (T.fail) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_NODES);
        if (logic_err) atomicOr(X.counters + 4, (unsigned long long)EX_ERR_LOGIC);
    }
}

// ------------------------------------------------------------------------------------ E1 helpers

__device__ __forceinline__ int ref_record_of(const int64_t *__restrict__ roff, int nref, int64_t sA)
{
    int lo = 0, hi = nref - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (roff[mid] + 1 <= sA) lo = mid; else hi = mid - 1; }
    return lo;
}

// per match: reference record, local coordinate, tag, "starts a piece" (= cluster of postnuc:
// an mgaps cluster, split where consecutive matches lie in different reference records)
__global__ void __launch_bounds__(256) k_ex_match_rec(const int32_t *__restrict__ m3, const int4 *__restrict__ recs, int64_t nc, int64_t nm,
                                                     const int64_t *__restrict__ roff, int nref, int32_t *__restrict__ mA, int32_t *__restrict__ mB,
                                                     int32_t *__restrict__ mL, int32_t *__restrict__ mrec, uint32_t *__restrict__ pstart, int32_t *__restrict__ mtag)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nm) return;
    int64_t lo = 0, hi = nc - 1;        // the mgaps cluster holding match g
    while (lo < hi) { int64_t mid = (lo + hi + 1) >> 1; if (recs[mid].x <= g) lo = mid; else hi = mid - 1; }
    const int4 r = recs[lo];
    const int64_t sA = m3[g * 3];
    const int rec = ref_record_of(roff, nref, sA);
    mA[g] = (int32_t)(sA - roff[rec]); mB[g] = m3[g * 3 + 1]; mL[g] = m3[g * 3 + 2]; mrec[g] = rec; mtag[g] = r.z;
    pstart[g] = (g == r.x || ref_record_of(roff, nref, m3[(g - 1) * 3]) != rec) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_ex_pieces(const uint32_t *__restrict__ pstart, const uint32_t *__restrict__ ppos, int64_t nm, const int32_t *__restrict__ mA,
                                                  const int32_t *__restrict__ mrec, const int32_t *__restrict__ mtag, uint64_t *__restrict__ keys,
                                                  uint32_t *__restrict__ vals, int32_t *__restrict__ pfirst)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nm || !pstart[g]) return;
    uint32_t p = ppos[g];
    keys[p] = ((uint64_t)(mtag[g] >> 1) << 47) | ((uint64_t)mrec[g] << 32) | (uint32_t)mA[g];
    vals[p] = p; pfirst[p] = (int32_t)g;
}

// sorted pieces -> ExCluster records and the inverse permutation
__global__ void __launch_bounds__(256) k_ex_clusters(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ svals, int64_t np, int64_t nm,
                                                    const int32_t *__restrict__ pfirst, const int32_t *__restrict__ mtag,
                                                    ExCluster *__restrict__ cl, int32_t *__restrict__ inv, uint32_t *__restrict__ sflag)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    const uint32_t p = svals[k];
    const int first = pfirst[p];
    const int end = p + 1 < np ? pfirst[p + 1] : (int)nm;
    ExCluster c; c.mfirst = first; c.nm = end - first; c.dir = mtag[first] & 1; c.syn = 0; c.order = (int32_t)p; c.pad = 0;
    cl[k] = c;
    inv[p] = (int32_t)k;
    sflag[k] = (k == 0 || (skeys[k] >> 32) != (skeys[k - 1] >> 32)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_ex_mcl(const uint32_t *__restrict__ pstart, const uint32_t *__restrict__ ppos, const int32_t *__restrict__ inv, int64_t nm, int32_t *__restrict__ mcl)
{
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nm) mcl[g] = inv[ppos[g] + pstart[g] - 1];
}

__global__ void __launch_bounds__(256) k_ex_syntenies(const uint64_t *__restrict__ skeys, const uint32_t *__restrict__ sflag, const uint32_t *__restrict__ spos, int64_t np,
                                                     ExCluster *__restrict__ cl, ExSynteny *__restrict__ syn, const int64_t *__restrict__ roff, const int64_t *__restrict__ rlen,
                                                     const int64_t *__restrict__ qoff, const int64_t *__restrict__ qlen, int64_t qn, int64_t single_nm)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= np) return;
    const uint32_t s = spos[k] + sflag[k] - 1;     // inclusive count of synteny starts up to k, minus one
    cl[k].syn = (int32_t)s;
    if (sflag[k]) {
        // single_nm >= 0: one reference and one query record, hence one synteny that holds every cluster and all single_nm
        // matches; its capacities are written here (no walk over the flags, no k_ex_syn_caps)
        int64_t e = k + 1; if (single_nm >= 0) e = np; else while (e < np && !sflag[e]) e++;
        const int qrec = (int)(skeys[k] >> 47), rrec = (int)((skeys[k] >> 32) & 0x7fff);
        ExSynteny S;
        S.cfirst = (int32_t)k; S.nC = (int32_t)(e - k); S.qrec = qrec; S.rrec = rrec;
        S.Abase = roff[rrec]; S.lenA = rlen[rrec]; S.BbaseF = qoff[qrec]; S.BbaseR = qn - qoff[qrec] - qlen[qrec]; S.lenB = qlen[qrec];
        S.alfirst = 0; S.alcap = 0; S.nodefirst = 0; S.nodecap = 0;
        if (single_nm >= 0) { S.alcap = (int32_t)single_nm; S.nodecap = (int32_t)(3 * single_nm + 8); }
        syn[s] = S;
    }
}

// capacities per synteny (its clusters are contiguous in the sorted order): one block, a thread per synteny, a running
// carry from one chunk of 256 syntenies to the next
struct OpAddI32 { __device__ __forceinline__ int operator()(int a, int b) const { return a + b; } static __device__ __forceinline__ int identity() { return 0; } };
__global__ void __launch_bounds__(256) k_ex_syn_caps(ExSynteny *__restrict__ syn, int nS, const ExCluster *__restrict__ cl)
{
    __shared__ int sm[32]; __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nS; base += 256) {
        const int s = base + threadIdx.x;
        int m = 0;
        if (s < nS) { const int f = syn[s].cfirst, e = f + syn[s].nC; for (int k = f; k < e; k++) m += cl[k].nm; }
        const int inc = pmn_block_scan_incl(m, OpAddI32(), sm);
        const int carry = carry_s;
        if (s < nS) {
            const int al = carry + inc - m;          // matches of the syntenies before this one
            syn[s].alfirst = al; syn[s].alcap = m; syn[s].nodefirst = 3 * al + 8 * s; syn[s].nodecap = 3 * m + 8;
        }
        __syncthreads();
        if (threadIdx.x == 255) carry_s = carry + inc;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------ E4: flatten, parseDelta

// the live alignments of all syntenies in output order (one thread: they are few)
__global__ void k_ex_list(const ExSynteny *__restrict__ syn, const int32_t *__restrict__ syn_nal, int nS, const ExAlign *__restrict__ al,
                          int32_t *__restrict__ al_syn, uint32_t *__restrict__ al_slot, uint32_t *__restrict__ dcount, int32_t *__restrict__ slot2out,
                          unsigned long long *__restrict__ totals)
{
    if (blockIdx.x || threadIdx.x) return;
    unsigned long long n = 0, nd = 0;
    for (int s = 0; s < nS; s++)
        for (int k = 0; k < syn_nal[s]; k++) {
            const int slot = syn[s].alfirst + k;
            al_syn[n] = s; al_slot[n] = (uint32_t)slot; dcount[n] = (uint32_t)al[slot].ndelta; slot2out[slot] = (int32_t)n;
            nd += (unsigned long long)al[slot].ndelta; n++;
        }
    dcount[n] = 0;
    totals[0] = n; totals[1] = nd;
}

// explicit nodes: one thread per node slot (every node knows its alignment and its offset)
__global__ void __launch_bounds__(256) k_ex_flat_nodes(ExShared X, int64_t ncap, const int32_t *__restrict__ slot2out, const uint32_t *__restrict__ dstart,
                                                      int32_t *__restrict__ dout)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncap) return;
    int lo = 0, hi = X.nS - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (X.syn[mid].nodefirst <= t) lo = mid; else hi = mid - 1; }
    if (t - X.syn[lo].nodefirst >= X.syn_nal[X.nS + lo]) return;
    const ExNode n = X.nodes[t];
    if (n.type == 1) return;
    const int k = slot2out[n.alslot];
    if (k < 0) return;
    uint32_t src = n.a; int cnt = n.cnt, b = n.b, outoff = n.outoff;
    if (n.type == 2) {
        // the end job of a cluster taken in one step (k_ex_stitch): it follows the deltas of the inner jobs, and its first delta
        // counts from the last indel before it
        const ExCSum c = X.cs[n.next];
        const int P = (c.bulk_cnt > 0 && c.bulk_P >= 0) ? c.bulk_P : n.b;
        src = c.e_doff; cnt = c.e_dcnt; b = c.eAl - P - 1; outoff = n.outoff + c.bulk_cnt;
    }
    int32_t *out = dout + dstart[k] + outoff;
    for (int i = 0; i < cnt; i++) {
        int d = X.pool[src + i];
        if (i == 0) d += d > 0 ? b : -b;
        out[i] = d;
    }
}

// range nodes: one thread per wave-1 job; `mk` is the inclusive max-scan of markkey
__global__ void __launch_bounds__(256) k_ex_flat_ranges(ExShared X, const unsigned long long *__restrict__ mk, const int32_t *__restrict__ slot2out,
                                                       const uint32_t *__restrict__ dstart, int32_t *__restrict__ dout)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= X.nM) return;
    const unsigned long long key = mk[g];
    if (!key) return;
    const ExNode n = X.nodes[(uint32_t)(key & 0xffffffffull) - 1];
    const int g0 = (int)n.a;
    if (n.type == 0 || g < g0 || g >= n.cnt) return;
    const ExJob j = X.jobs[g];
    if (j.dcnt <= 0) return;
    const int k = slot2out[n.alslot];
    if (k < 0) return;
    const int Pprev = g > g0 ? range_last_P(X, g0, (int)g - 1, n.b) : n.b;
    const int adjust = (X.mA[g] + X.mL[g] - 1) - Pprev - 1;
    int32_t *out = dout + dstart[k] + n.outoff + (X.dcnt_ex[g] - X.dcnt_ex[g0]);
    for (int t = 0; t < j.dcnt; t++) {
        int d = X.pool[j.doff + t];
        if (t == 0) d += d > 0 ? adjust : -adjust;
        out[t] = d;
    }
}

// reference / query bases consumed by each delta (for the position prefix sums)
__global__ void __launch_bounds__(256) k_ex_consumed(const int32_t *__restrict__ d, int64_t nd, uint32_t *__restrict__ fa, uint32_t *__restrict__ fb)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nd) return;
    if (t == nd) { fa[t] = 0; fb[t] = 0; return; }
    const int v = d[t];
    fa[t] = v > 0 ? (uint32_t)v : (uint32_t)(-v - 1);
    fb[t] = v > 0 ? (uint32_t)(v - 1) : (uint32_t)(-v);
}

__device__ __forceinline__ uint64_t spread32(uint32_t x)     // bit 31-j of x -> bit 62-2j of the result
{
    uint64_t v = x;
    v = (v | (v << 16)) & 0x0000ffff0000ffffull;
    v = (v | (v << 8)) & 0x00ff00ff00ff00ffull;
    v = (v | (v << 4)) & 0x0f0f0f0f0f0f0f0full;
    v = (v | (v << 2)) & 0x3333333333333333ull;
    v = (v | (v << 1)) & 0x5555555555555555ull;
    return v;
}

// columns of a gap-free run that are not an identical a/c/g/t pair
__device__ __forceinline__ int run_mismatches(const PackedView &R, int64_t a, const PackedView &Q, int64_t b, int64_t run)
{
    int cnt = 0;
    for (int64_t off = 0; off < run; off += 32) {
        uint64_t x = pmn_window64(R.w, a + off) ^ pmn_window64(Q.w, b + off);
        uint64_t m = (x | (x >> 1)) & 0x5555555555555555ull;
        if (R.has_x) m |= spread32(pmn_xwindow32(R.xm, a + off));
        if (Q.has_x) m |= spread32(pmn_xwindow32(Q.xm, b + off));
        const int64_t len = run - off;
        if (len < 32) m &= ~0ull << (64 - 2 * (int)len);
        cnt += __popcll(m);
    }
    return cnt;
}

// parseDelta in parallel: item t < nd = the run before delta t plus the indel itself;
// item nd + k = the run behind the last delta of alignment k
__global__ void __launch_bounds__(256) k_ex_errors(ExShared X, const int32_t *__restrict__ al_syn, const uint32_t *__restrict__ al_slot, int64_t nal,
                                                  const uint32_t *__restrict__ dstart, const int32_t *__restrict__ d, int64_t nd,
                                                  const uint32_t *__restrict__ sa_, const uint32_t *__restrict__ sb_, unsigned long long *__restrict__ errs)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t k = -1; unsigned long long e = 0;
    if (t < nd + nal) {
        if (t < nd) {
            int64_t lo = 0, hi = nal - 1;      // last alignment whose first delta is <= t
            while (lo < hi) { int64_t mid = (lo + hi + 1) >> 1; if ((int64_t)dstart[mid] <= t) lo = mid; else hi = mid - 1; }
            k = lo;
        } else k = t - nd;
        const ExSynteny S = X.syn[al_syn[k]];
        const ExAlign a = X.al[al_slot[k]];
        const PackedView &Q = a.dirB ? X.QR : X.QF;
        const int64_t Ab = S.Abase - 1, Bb = (a.dirB ? S.BbaseR : S.BbaseF) - 1;
        const uint32_t first = dstart[k];
        const int64_t idx = t < nd ? t : (int64_t)dstart[k + 1];
        const int64_t Apos = a.sA + (int64_t)(uint32_t)(sa_[idx] - sa_[first]), Bpos = a.sB + (int64_t)(uint32_t)(sb_[idx] - sb_[first]);
        int64_t run;
        if (t < nd) { const int v = d[t]; run = (v < 0 ? -v : v) - 1; e = 1; } else run = (int64_t)a.eA - Apos + 1;
        if (run > 0) e += (unsigned long long)run_mismatches(X.R, Ab + Apos, Q, Bb + Bpos, run);
    }
    // a warp lies inside one alignment almost always (there are few of them): one atomic per warp then, not 32 on one address
    const int64_t k0 = __shfl_sync(0xffffffffu, k, 0);
    if (__all_sync(0xffffffffu, e == 0 || k == k0)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e += __shfl_xor_sync(0xffffffffu, e, o);
        if ((threadIdx.x & 31) == 0 && e) atomicAdd(errs + k0, e);
    }
    else if (e) atomicAdd(errs + k, e);
}

__global__ void __launch_bounds__(256) k_ex_rows(ExShared X, const int32_t *__restrict__ al_syn, const uint32_t *__restrict__ al_slot, int64_t nal,
                                                const unsigned long long *__restrict__ errs, long long *__restrict__ rows)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nal) return;
    const ExSynteny S = X.syn[al_syn[k]];
    const ExAlign a = X.al[al_slot[k]];
    long long *r = rows + k * 10;
    r[0] = S.rrec; r[1] = S.qrec; r[2] = a.dirB; r[3] = a.sA; r[4] = a.eA; r[5] = a.sB; r[6] = a.eB; r[7] = (long long)errs[k]; r[8] = r[7]; r[9] = 0;
}

// ------------------------------------------------------------------------------------ driver

struct OpMaxI64x { __device__ __forceinline__ long long operator()(long long a, long long b) const { return a > b ? a : b; } static __device__ __forceinline__ long long identity() { return LLONG_MIN; } };
struct OpMaxU64 { __device__ __forceinline__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; } static __device__ __forceinline__ unsigned long long identity() { return 0ull; } };

int pmn_extend_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *q, const pmn_opts *o, pmn_result *res)
{
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    const pmn_seq *ref = ix->seq;
    const int64_t nc0 = S.n_clusters, nm = S.n_cl_matches;
    res->stats.clusters = nc0; res->stats.cluster_matches = nm;
    res->al_doff.assign(1, 0);
    if (o->keep_stages && nc0 > 0) {
        std::vector<int4> recs((size_t)nc0);
        res->cl_matches.resize((size_t)nm * 3);
        PMN_D2H(c, res->cl_matches.data(), S.cl_matches.p, 12 * (size_t)nm);
        PMN_D2H(c, recs.data(), S.cl_recs.p, 16 * (size_t)nc0);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        res->cl_off.resize((size_t)nc0 + 1); res->cl_tag.resize((size_t)nc0);
        for (int64_t k = 0; k < nc0; k++) { res->cl_off[(size_t)k] = recs[(size_t)k].x; res->cl_tag[(size_t)k] = recs[(size_t)k].z; }
        res->cl_off[(size_t)nc0] = (int32_t)nm;
    } else if (o->keep_stages) { res->cl_off.assign(1, 0); }
    if (nc0 <= 0) return 0;
    int launches = 0;

    // ---- E1: record offsets to the device
    const int nref = ref->nrec, nqry = q->nrec;
    std::vector<int64_t> h((size_t)2 * nref + 2 * nqry);
    for (int i = 0; i < nref; i++) { h[(size_t)i] = ref->off[(size_t)i]; h[(size_t)nref + i] = ref->len[(size_t)i]; }
    for (int i = 0; i < nqry; i++) { h[(size_t)2 * nref + i] = q->off[(size_t)i]; h[(size_t)2 * nref + nqry + i] = q->len[(size_t)i]; }
    if (S.ex_a.ensure(8 * h.size()) || S.ex_b.ensure(4 * 5 * (size_t)nm) || S.ex_c.ensure(4 * 2 * (size_t)nm) || S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(nm + 1)) ||
        S.ensure_pinned(512)) return -3;
    PMN_H2D(c, S.ex_a.p, h.data(), 8 * h.size());
    const int64_t *roff = S.ex_a.as<int64_t>(), *rlen = roff + nref, *qoff = roff + 2 * nref, *qlen = qoff + nqry;
    int32_t *mA = S.ex_b.as<int32_t>(), *mB = mA + nm, *mL = mA + 2 * nm, *mrec = mA + 3 * nm, *mtag = mA + 4 * nm;
    uint32_t *pstart = S.ex_c.as<uint32_t>(), *ppos = pstart + nm;
    const unsigned gm = (unsigned)((nm + 255) / 256);
    k_ex_match_rec<<<gm, 256, 0, st>>>(S.cl_matches.as<int32_t>(), S.cl_recs.as<int4>(), nc0, nm, roff, nref, mA, mB, mL, mrec, pstart, mtag);
    pmn_scan<uint32_t, OpAddU32, false>(pstart, ppos, nm, S.scan_tmp.as<uint32_t>(), st);
    launches += 4;
    uint32_t *tail = (uint32_t *)S.pinned;
    int64_t np = nc0;           // one reference record: no mgaps cluster is split, the pieces are the clusters and the host knows their number
    if (nref > 1) {
        PMN_D2H(c, tail, ppos + (nm - 1), 4);
        PMN_D2H(c, tail + 1, pstart + (nm - 1), 4);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        c->syncs++;
        np = (int64_t)tail[0] + tail[1];
    }

    // pieces sorted by (query record, reference record, first reference start), stable
    if (S.k0.ensure(8 * (size_t)np) || S.k1.ensure(8 * (size_t)np) || S.v0.ensure(4 * (size_t)np) || S.v1.ensure(4 * (size_t)np) ||
        S.ex_d.ensure(4 * 2 * (size_t)np) || S.ex_e.ensure(sizeof(ExCluster) * (size_t)np) || S.ex_f.ensure(4 * (size_t)nm) ||
        S.ex_g.ensure(4 * 2 * (size_t)np) || S.ex_h.ensure(sizeof(ExSynteny) * (size_t)np)) return -3;
    int32_t *pfirst = S.ex_d.as<int32_t>(), *inv = pfirst + np;
    k_ex_pieces<<<gm, 256, 0, st>>>(pstart, ppos, nm, mA, mrec, mtag, S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), pfirst);
    launches++;
    // only the key bits in use are sorted: reference start, then the record fields if there are several records
    int nbits = 1; { int64_t mx = 1; for (int i = 0; i < nref; i++) mx = std::max(mx, ref->len[(size_t)i]); while ((1ll << nbits) <= mx) nbits++; }
    if (nref > 1 || nqry > 1) { int qb = 1; while ((1 << qb) < nqry) qb++; nbits = 47 + qb; }
    int where = pmn_radix_sort(S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), S.k1.as<uint64_t>(), S.v1.as<uint32_t>(), np, nbits, S.rs, st, &launches);
    if (where < 0) return -3;
    const uint64_t *skeys = where ? S.k1.as<uint64_t>() : S.k0.as<uint64_t>();
    const uint32_t *svals = where ? S.v1.as<uint32_t>() : S.v0.as<uint32_t>();
    ExCluster *cl = S.ex_e.as<ExCluster>(); int32_t *mcl = S.ex_f.as<int32_t>();
    uint32_t *sflag = S.ex_g.as<uint32_t>(), *spos = sflag + np;
    ExSynteny *syn = S.ex_h.as<ExSynteny>();
    const unsigned gp = (unsigned)((np + 255) / 256);
    k_ex_clusters<<<gp, 256, 0, st>>>(skeys, svals, np, nm, pfirst, mtag, cl, inv, sflag);
    k_ex_mcl<<<gm, 256, 0, st>>>(pstart, ppos, inv, nm, mcl);
    pmn_scan<uint32_t, OpAddU32, false>(sflag, spos, np, S.scan_tmp.as<uint32_t>(), st);
    const bool single_syn = nref == 1 && nqry == 1;
    k_ex_syntenies<<<gp, 256, 0, st>>>(skeys, sflag, spos, np, cl, syn, roff, rlen, qoff, qlen, q->n, single_syn ? nm : (int64_t)-1);
    launches += 6;
    int nS = 1;                 // one reference record and one query record: one synteny
    if (nref > 1 || nqry > 1) {
        PMN_D2H(c, tail, spos + (np - 1), 4);
        PMN_D2H(c, tail + 1, sflag + (np - 1), 4);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        c->syncs++;
        nS = (int)(tail[0] + tail[1]);
    }
    if (!single_syn) { k_ex_syn_caps<<<1, 256, 0, st>>>(syn, nS, cl); launches++; }

    // ---- E2/E3 storage
    // persistent blocks per SM of the warp-per-job and thread-per-job kernels.  Two each: with four, every worker held twice the
    // private traceback slots and score rows (3.6 GB instead of 1.8 GB) and a batch was no faster (1734 vs 1759 pairs/s at 16
    // workers); at two, 32 workers fit one B200 and all 28 pairs of a C2 step are in flight together (2040 pairs/s)
    static const int big_bps = getenv("PMN_BIG_BPS") ? std::max(1, atoi(getenv("PMN_BIG_BPS"))) : 2, tpj_bps = getenv("PMN_TPJ_BPS") ? std::min(TPJ_BLOCKS_PER_SM, std::max(1, atoi(getenv("PMN_TPJ_BPS")))) : 2;
    const int blocks1 = c->sm_count * big_bps;                // 4 warps per block
    const int blocks_st = (nS + EX_WARPS_PER_BLOCK - 1) / EX_WARPS_PER_BLOCK;
    const int nslots = std::max(blocks1, blocks_st) * EX_WARPS_PER_BLOCK;                 // warps that may run the wide fallback (global score rows)
    const int nslots_tb = nslots;                                                          // warps that keep a private traceback header
    const size_t pool_cap = (size_t)std::max<int64_t>(1 << 20, 8 * nm + (ref->n + q->n) / 8);
    // Traceback arena: 2 GB for every pair made a worker hold 2.3 GB that a bacterial pair uses a few per cent of.  It starts at
    // 256 MB (PMN_ARENA_MB), and a pair that runs out of it grows it fourfold and is extended again (pmn_api.cu).
    if (!S.arena_cap) { static const size_t mb = getenv("PMN_ARENA_MB") ? (size_t)atoll(getenv("PMN_ARENA_MB")) : 256; S.arena_cap = std::max<size_t>(16, mb) << 20; }
    S.arena_retry = false;
    if (S.ex_arena.cap > S.arena_cap + 4096) S.arena_cap = S.ex_arena.cap - 4096;       // the scheduler levelled the buffer with a worker that had to grow
    const size_t arena_cap = S.arena_cap;
    const size_t ncap_al = (size_t)nm, ncap_nodes = 3 * (size_t)nm + 8 * (size_t)nS;
    const size_t npad = ((size_t)np + 63) / 64 * 64;
    const size_t l_bytes = npad * 2 + 8 * npad + 8 * (size_t)nS + 64 + 16 + sizeof(ExBack) * npad;      // fused, anyfail, aimers, aimfail, syn_nal (2 x nS), back
    if (S.ex_i.ensure(sizeof(ExJob) * (size_t)nm) || S.ex_j.ensure(sizeof(ExAlign) * ncap_al) || S.ex_k.ensure(sizeof(ExNode) * ncap_nodes) ||
        S.ex_pool.ensure(4 * pool_cap) || S.ex_arena.ensure(arena_cap + 4096) || S.ex_scores.ensure(4 * (size_t)EX_ROWS * EX_WCAP * (size_t)nslots) ||
        S.ex_tb.ensure((size_t)EX_TBW * (size_t)nslots_tb + 4096) || S.ex_counters.ensure(256) || S.ex_l.ensure(l_bytes) ||
        S.ex_tbidx.ensure(8 * 3 * (size_t)(nm + 1)) || S.ex_a.ensure(8 * h.size() + 0) || S.cl_l.ensure(sizeof(ExCSum) * (size_t)np)) return -3;
    PMN_CUDA_OK(cudaMemsetAsync(S.ex_i.p, 0, sizeof(ExJob) * (size_t)nm, st));
    PMN_CUDA_OK(cudaMemsetAsync(S.ex_counters.p, 0, 256, st));
    PMN_CUDA_OK(cudaMemsetAsync(S.ex_l.p, 0, l_bytes, st));
    ExShared X;
    X.R = ref->fwd(); X.QF = q->fwd(); X.QR = q->rev();
    X.mA = mA; X.mB = mB; X.mL = mL; X.cl = cl; X.syn = syn; X.nC = (int)np; X.nS = nS; X.nM = nm;
    X.jobs = S.ex_i.as<ExJob>(); X.mcl = mcl; X.al = S.ex_j.as<ExAlign>(); X.nodes = S.ex_k.as<ExNode>();
    X.pool = S.ex_pool.as<int32_t>(); X.pool_cap = (uint32_t)std::min<size_t>(pool_cap, 0xfffffff0u);
    X.arena = S.ex_arena.as<uint8_t>(); X.arena_cap = arena_cap;
    X.gscore = S.ex_scores.as<int32_t>(); X.tbpriv = S.ex_tb.as<uint8_t>();
    X.counters = S.ex_counters.as<unsigned long long>();
    X.breaklen = o->breaklen; X.do_extend = o->do_extend; X.do_simplify = o->do_simplify;
    { static const int cells = getenv("PMN_TPJ_CELLS") ? atoi(getenv("PMN_TPJ_CELLS")) : 0; X.tpj_cells = cells > 0 ? cells : c->tpj_cells; }       // larger windows: one warp each (eng_mid_full)
    uint8_t *fused = S.ex_l.as<uint8_t>();
    uint8_t *anyfail = fused + npad;
    X.aimers = (int32_t *)(anyfail + npad);
    X.aimfail = X.aimers + npad;
    X.syn_nal = X.aimfail + npad;
    X.back = (ExBack *)(X.syn_nal + ((2 * (size_t)nS + 16 + 3) & ~(size_t)3));      // 16-byte aligned: the records are read with vector loads
    long long *pkey = S.ex_tbidx.as<long long>();                          // nm+1
    unsigned long long *markkey = (unsigned long long *)(pkey + (nm + 1));  // nm+1
    uint32_t *dcnt = (uint32_t *)(markkey + (nm + 1));                      // nm+1, then dcnt_ex nm+1
    uint32_t *dcnt_ex = dcnt + (nm + 1);
    X.dcnt_ex = dcnt_ex; X.lastP = pkey; X.anyfail = anyfail; X.markkey = markkey;
    ExCSum *cs = S.cl_l.as<ExCSum>();      // the clustering scratch is free by now
    X.cs = cs;
    PMN_CUDA_OK(cudaMemsetAsync(markkey, 0, 8 * (size_t)(nm + 1), st));
    if (S.ex_desc.ensure(sizeof(ExJobDesc) * (size_t)(np + nm + 1) + 8 * (size_t)(nm + 2) + 16 * (size_t)np + 16)) return -3;
    ExJobDesc *descA = S.ex_desc.as<ExJobDesc>(), *descB = descA + np;
    X.descA = descA; X.descB = descB; X.overflow = (int32_t *)(descB + nm + 1); X.overflow2 = X.overflow + (nm + 1);
    int4 *tgt = (int4 *)(((uintptr_t)(X.overflow2 + (nm + 1)) + 15) & ~(uintptr_t)15);
    X.tgt = tgt;
    // thread-per-job windows: (bin, rank) per match, bin counters and starts, the sorted list, per-warp scratch
    const int blocks_tpj = c->sm_count * tpj_bps;
    const size_t tkey_bytes = 8 * (size_t)nm, tbin_bytes = 4 * (size_t)(2 * TPJ_BINS + 2);
    if (S.ex_tkey.ensure(tkey_bytes + tbin_bytes + 4 * (size_t)nm + 64) || S.ex_tscratch.ensure((size_t)blocks_tpj * 4 * TPJ_SLOT_BYTES)) return -3;
    X.tkey = S.ex_tkey.as<uint2>(); X.tbin = (uint32_t *)(X.tkey + nm); X.tsorted = (int32_t *)(X.tbin + 2 * TPJ_BINS + 2);
    X.tscratch = S.ex_tscratch.as<uint8_t>();
    PMN_CUDA_OK(cudaMemsetAsync(X.tkey, 0xff, tkey_bytes, st));
    PMN_CUDA_OK(cudaMemsetAsync(X.tbin, 0, tbin_bytes, st));
    const char *joblog = getenv("PMN_JOBLOG");
    X.dbg = nullptr; X.dbg_cap = 0;
    if (joblog) {
        X.dbg_cap = (unsigned)(nm + 4096);
        if (S.ex_dbg.ensure(32 * (size_t)X.dbg_cap)) return -3;
        X.dbg = S.ex_dbg.as<int4>();
    }

    const size_t smem = (size_t)EX_WARPS_PER_BLOCK * CfgBig::WARP_BYTES;
    if (!c->smem_attr_set) {
        PMN_CUDA_OK(cudaFuncSetAttribute(k_ex_wave1_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PMN_CUDA_OK(cudaFuncSetAttribute(k_ex_stitch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        c->smem_attr_set = true;
    }
    int b1 = blocks1; { int64_t need = (nm + EX_WARPS_PER_BLOCK - 1) / EX_WARPS_PER_BLOCK; if (need < b1) b1 = (int)need; if (b1 < 1) b1 = 1; }
    if (o->do_extend) { k_ex_targets<<<(unsigned)((np + 7) / 8), 256, 0, st>>>(X, tgt); launches++; }
    k_ex_jobdesc<<<gm, 256, 0, st>>>(X, descA, descB);
    PMN_CUDA_OK(cudaEventRecord(c->ev[8], st));
    // the warp-per-job kernel (cluster ends: few, long) runs beside the thread-per-job kernel on the context's second stream
    PMN_CUDA_OK(cudaEventRecord(c->ev_fork, st));
    PMN_CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    k_ex_wave1_big<<<b1, EX_WARPS_PER_BLOCK * 32, smem, c->stream2>>>(X, 0);
    PMN_CUDA_OK(cudaEventRecord(c->ev_join, c->stream2));
    k_ex_tbinscan<<<1, 512, 0, st>>>(X);
    k_ex_tscatter<<<gm, 256, 0, st>>>(X);
    {
        int bt = blocks_tpj;
        const int64_t need = (nm + 127) / 128;
        if (need < bt) bt = (int)std::max<int64_t>(1, need);
        k_ex_wave1_tpj<<<bt, 128, 0, st>>>(X);
    }
    PMN_CUDA_OK(cudaStreamWaitEvent(st, c->ev_join, 0));
    k_ex_wave1_big<<<b1, EX_WARPS_PER_BLOCK * 32, smem, st>>>(X, 1);
    PMN_CUDA_OK(cudaEventRecord(c->ev[9], st));
    PMN_D2H(c, (unsigned long long *)S.pinned + 40, X.counters + 2, 8);      // cells evaluated by wave 1
    k_ex_jobmeta<<<(unsigned)((nm + 1 + 255) / 256), 256, 0, st>>>(X.jobs, mcl, cl, pstart, ppos, nm, dcnt, pkey, anyfail);
    pmn_scan<uint32_t, OpAddU32, false>(dcnt, dcnt_ex, nm + 1, S.scan_tmp.as<uint32_t>(), st);
    pmn_scan<long long, OpMaxI64x, true>(pkey, pkey, nm, S.scan_tmp.as<long long>(), st);
    k_ex_csum<<<gp, 256, 0, st>>>(X, cs);
    PMN_CUDA_OK(cudaEventRecord(c->ev[11], st));
    k_ex_stitch<<<blocks_st, EX_WARPS_PER_BLOCK * 32, smem, st>>>(X, fused);
    PMN_CUDA_OK(cudaEventRecord(c->ev[10], st));
    launches += 15;

    // ---- E4
    if (S.ex_c.ensure(4 * 5 * (size_t)(nm + 1) + 64)) return -3;     // al_syn, al_slot, dcount, dstart, slot2out  (pstart/ppos are dead now)
    int32_t *al_syn = S.ex_c.as<int32_t>(); uint32_t *al_slot = (uint32_t *)al_syn + (nm + 1), *dcount = al_slot + (nm + 1), *dstart = dcount + (nm + 1);
    int32_t *slot2out = (int32_t *)(dstart + (nm + 1));
    PMN_CUDA_OK(cudaMemsetAsync(slot2out, 0xff, 4 * (size_t)(nm + 1), st));
    k_ex_list<<<1, 32, 0, st>>>(syn, X.syn_nal, nS, X.al, al_syn, al_slot, dcount, slot2out, X.counters + 8);
    launches++;
    unsigned long long *hc = (unsigned long long *)S.pinned;
    PMN_D2H(c, hc, X.counters, 256);
    PMN_CUDA_OK(cudaStreamSynchronize(st));
    c->syncs++;
    const unsigned long long errflags = hc[4];
    if (joblog) {
        const size_t nrec = (size_t)std::min<unsigned long long>(hc[15], X.dbg_cap);
        std::vector<int4> rec(2 * nrec);
        PMN_CUDA_OK(cudaMemcpy(rec.data(), X.dbg, 32 * nrec, cudaMemcpyDeviceToHost));
        if (FILE *f = fopen(joblog, "a")) {
            fprintf(f, "# pair nm=%lld nC=%lld  columns: m_o N M d_end cells cycles kernel path\n", (long long)nm, (long long)np);
            for (size_t k = 0; k < nrec; k++)
                fprintf(f, "%d %d %d %d %d %d %d %d\n", rec[2 * k].x, rec[2 * k].y, rec[2 * k].z, rec[2 * k].w, rec[2 * k + 1].x, rec[2 * k + 1].y, rec[2 * k + 1].z, rec[2 * k + 1].w);
            fclose(f);
        }
    }
    res->stats.dp_cells = (int64_t)hc[2]; res->stats.dp_jobs = (int64_t)hc[3];
    res->stats.arena_bytes = (int64_t)hc[1];
    res->stats.wave1_cells = (int64_t)hc[40];
#ifdef PMN_STITCH_TIMING
    if (getenv("PMN_STITCH_TIMING"))     // cycles of the slowest synteny warp per section (k_ex_stitch)
        fprintf(stderr, "stitch cycles: total %llu window %llu shadow %llu start %llu match %llu end %llu clusters %llu starts %llu nextfail %llu fwd-calls %llu\n",
                hc[16], hc[17], hc[18], hc[19], hc[20], hc[21], hc[22], hc[23], hc[24], hc[25]);
#endif
    cudaEventElapsedTime(&res->stats.ms_wave1, c->ev[8], c->ev[9]);
    cudaEventElapsedTime(&res->stats.ms_stitch, c->ev[11], c->ev[10]);
    c->launches += launches; launches = 0;
    if (errflags & EX_ERR_POOL) return pmn_set_error(PMN_E_NOMEM, "extend: delta pool exhausted (%zu entries)", pool_cap);
    if (errflags & EX_ERR_ARENA) {
        if (arena_cap < ((size_t)32 << 30)) { S.arena_cap = arena_cap * 4; S.arena_retry = true; }
        return pmn_set_error(PMN_E_NOMEM, "extend: traceback arena exhausted (%zu bytes)", arena_cap);
    }
    if (errflags & EX_ERR_NODES) return pmn_set_error(PMN_E_INTERNAL, "extend: delta segment list overflow");
    if (errflags & EX_ERR_LOGIC) return pmn_set_error(PMN_E_INTERNAL, "extend: inconsistent cluster chain (target match does not exist)");
    const int64_t nal = (int64_t)hc[8], nd = (int64_t)hc[9];
    if (nal > 0) {
        if (S.ex_d.ensure(4 * 5 * (size_t)(nd + 1) + 64) || S.ex_g.ensure(88 * (size_t)nal + 64) || S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(std::max(nd + 1, nm + 1)))) return -3;
        int32_t *dflat = S.ex_d.as<int32_t>(); uint32_t *fa = (uint32_t *)dflat + (nd + 1), *fb = fa + (nd + 1), *sa_ = fb + (nd + 1), *sb_ = sa_ + (nd + 1);
        long long *rows = S.ex_g.as<long long>(); unsigned long long *errs = (unsigned long long *)(rows + 10 * nal);
        PMN_CUDA_OK(cudaMemsetAsync(errs, 0, 8 * (size_t)nal, st));
        pmn_scan<uint32_t, OpAddU32, false>(dcount, dstart, nal + 1, S.scan_tmp.as<uint32_t>(), st);
        launches += 3;
        if (nd > 0) {
            pmn_scan<unsigned long long, OpMaxU64, true>(markkey, markkey, nm, S.scan_tmp.as<unsigned long long>(), st);
            k_ex_flat_nodes<<<(unsigned)((ncap_nodes + 255) / 256), 256, 0, st>>>(X, (int64_t)ncap_nodes, slot2out, dstart, dflat);
            k_ex_flat_ranges<<<gm, 256, 0, st>>>(X, markkey, slot2out, dstart, dflat);
            launches += 5;
        }
        k_ex_consumed<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, st>>>(dflat, nd, fa, fb);
        pmn_scan<uint32_t, OpAddU32, false>(fa, sa_, nd + 1, S.scan_tmp.as<uint32_t>(), st);
        pmn_scan<uint32_t, OpAddU32, false>(fb, sb_, nd + 1, S.scan_tmp.as<uint32_t>(), st);
        k_ex_errors<<<(unsigned)((nd + nal + 255) / 256), 256, 0, st>>>(X, al_syn, al_slot, nal, dstart, dflat, nd, sa_, sb_, errs);
        k_ex_rows<<<(unsigned)((nal + 255) / 256), 256, 0, st>>>(X, al_syn, al_slot, nal, errs, rows);
        launches += 9;
        // rows, deltas and offsets come back through the context's pinned staging buffer: three copies back to back and one
        // wait (copies into pageable vectors are staged by the driver one after the other, 75 us of a 5 Mbp pair)
        const size_t b_rows = 80 * (size_t)nal, b_del = 4 * (size_t)nd, b_off = 4 * (size_t)(nal + 1);
        if (S.ensure_pinned(b_rows + b_del + b_off + 64)) return -3;         // hc is dead from here on
        char *hp = (char *)S.pinned;
        PMN_D2H(c, hp, rows, b_rows);
        if (nd) PMN_D2H(c, hp + b_rows, dflat, b_del);
        PMN_D2H(c, hp + b_rows + b_del, dstart, b_off);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        c->syncs++;
        res->al_rows.resize((size_t)nal * 10);
        res->al_deltas.resize((size_t)nd);
        memcpy(res->al_rows.data(), hp, b_rows);
        if (nd) memcpy(res->al_deltas.data(), hp + b_rows, b_del);
        const uint32_t *ds = (const uint32_t *)(hp + b_rows + b_del);
        res->al_doff.assign(ds, ds + nal + 1);
    }
    PMN_CUDA_OK(cudaGetLastError());
    c->launches += launches;
    return 0;
}

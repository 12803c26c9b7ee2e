/*
 * pmn_oracle.c — CPU ORACLE for the pairwise-nucmer hot path.  TEST INFRASTRUCTURE ONLY
 * (see pmn_oracle.h for who may call it).  PARITY UNPINNED vs a real MUMmer 3.20 binary.
 *
 * Written to be obviously correct, not fast: suffixes are sorted with a plain
 * comparator, MUMs are found by binary search plus neighbour counting, mgaps and
 * postnuc are sequential restatements.  Single-threaded, one pair per run.
 *
 * Boundary this stands in for: the child process started at
 *   /root/reference/lib/nucmer/mugsy_nucmer.ml:100   "nucmer %s %s -p %s %s"
 * whose output grammar is what /root/reference/lib/profiles_lib/m_delta.cc:72-220 and
 * /root/reference/lib/profiles/m_delta.ml:52-153 parse.
 *
 * Sections:  1 utilities  2 FASTA  3 index (SA+LCP)  4 seeding (mummer)
 *            5 clustering (mgaps)  6 alignment engine (sw_align)
 *            7 extension (postnuc)  8 .delta writer  9 API
 */
#include "pmn_oracle.h"
#include "../include/pmn_params.h"

#include <ctype.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ 1 utilities */

static __thread char g_err[512];
const char *pmo_last_error(void) { return g_err; }
static int fail(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
    return -1;
}

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) { fprintf(stderr, "pmn_oracle: out of memory\n"); abort(); } return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) { fprintf(stderr, "pmn_oracle: out of memory\n"); abort(); } return p; }

typedef struct { int64_t *v; int64_t n, cap; } vec64;
static void v64_push(vec64 *a, int64_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 16; a->v = (int64_t *)xrealloc(a->v, sizeof(int64_t) * (size_t)a->cap); }
    a->v[a->n++] = x;
}
typedef struct { int32_t *v; int64_t n, cap; } vec32;
static void v32_push(vec32 *a, int32_t x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 64; a->v = (int32_t *)xrealloc(a->v, sizeof(int32_t) * (size_t)a->cap); }
    a->v[a->n++] = x;
}
typedef struct { char *v; size_t n, cap; } strbuf;
static void sb_printf(strbuf *s, const char *fmt, ...)
{
    char tmp[1024];
    va_list ap; va_start(ap, fmt); int k = vsnprintf(tmp, sizeof tmp, fmt, ap); va_end(ap);
    if (k < 0) return;
    if ((size_t)k >= sizeof tmp) k = (int)sizeof tmp - 1;   /* ids longer than this are truncated */
    if (s->n + (size_t)k + 1 > s->cap) { s->cap = (s->cap + (size_t)k + 1) * 2; s->v = (char *)xrealloc(s->v, s->cap); }
    memcpy(s->v + s->n, tmp, (size_t)k); s->n += (size_t)k; s->v[s->n] = 0;
}

void pmo_default_opts(pmo_opts *o)
{
    o->minmatch = PMN_DEF_MINMATCH; o->mincluster = PMN_DEF_MINCLUSTER; o->maxgap = PMN_DEF_MAXGAP;
    o->diagdiff = PMN_DEF_DIAGDIFF; o->diagfactor = PMN_DEF_DIAGFACTOR; o->breaklen = PMN_DEF_BREAKLEN;
    o->do_forward = 1; o->do_reverse = 1; o->do_extend = 1; o->do_optimize = 1; o->do_simplify = 1;
    o->fast_chain = 0;
}

/* ------------------------------------------------------------------ 2 FASTA */

/* A multi-FASTA file as base codes.  Like MUMmer's `mummer -n`, only a/c/g/t (either
 * case) can match; every other character becomes PMN_CODE_X.  Record ids are the first
 * whitespace-delimited token of the header, which is what the `>` line of a .delta
 * carries (/root/reference/lib/base/m_rewrite_fasta.ml:5-59 makes them species.accession). */
typedef struct {
    int nrec;
    char **id;
    int64_t *len;       /* bases per record */
    int64_t *off;       /* 0-based offset of the record inside `cat` */
    uint8_t *cat;       /* records joined by ONE separator of code X (prenuc's 'x') */
    int64_t ncat;
} seqset;

static int code_of(int c)
{
    switch (c) {
        case 'a': case 'A': return PMN_CODE_A;
        case 'c': case 'C': return PMN_CODE_C;
        case 'g': case 'G': return PMN_CODE_G;
        case 't': case 'T': return PMN_CODE_T;
        default: return PMN_CODE_X;
    }
}

static int parse_fasta(const char *txt, size_t nb, seqset *s)
{
    memset(s, 0, sizeof *s);
    int caprec = 0; size_t capcat = nb + 16; s->cat = (uint8_t *)xmalloc(capcat);
    size_t i = 0; int cur = -1;
    while (i < nb) {
        size_t e = i; while (e < nb && txt[e] != '\n') e++;
        if (e > i && txt[i] == '>') {
            if (s->nrec == caprec) {
                caprec = caprec ? caprec * 2 : 4;
                s->id = (char **)xrealloc(s->id, sizeof(char *) * (size_t)caprec);
                s->len = (int64_t *)xrealloc(s->len, sizeof(int64_t) * (size_t)caprec);
                s->off = (int64_t *)xrealloc(s->off, sizeof(int64_t) * (size_t)caprec);
            }
            size_t a = i + 1; while (a < e && (txt[a] == ' ' || txt[a] == '\t')) a++;
            size_t b = a; while (b < e && !isspace((unsigned char)txt[b])) b++;
            cur = s->nrec++;
            s->id[cur] = (char *)xmalloc(b - a + 1); memcpy(s->id[cur], txt + a, b - a); s->id[cur][b - a] = 0;
            if (cur > 0) s->cat[s->ncat++] = PMN_CODE_X;
            s->off[cur] = s->ncat; s->len[cur] = 0;
        } else if (cur >= 0) {
            for (size_t k = i; k < e; k++) {
                if (isspace((unsigned char)txt[k])) continue;
                s->cat[s->ncat++] = (uint8_t)code_of((unsigned char)txt[k]); s->len[cur]++;
            }
        } else {
            for (size_t k = i; k < e; k++) if (!isspace((unsigned char)txt[k])) return fail("FASTA: sequence data before the first '>' header");
        }
        i = e + 1;
    }
    if (s->nrec == 0) return fail("FASTA: no records");
    if (s->ncat > (int64_t)INT32_MAX - 2) return fail("FASTA: more than 2^31 bases");
    return 0;
}

static void free_seqset(seqset *s)
{
    for (int i = 0; i < s->nrec; i++) free(s->id[i]);
    free(s->id); free(s->len); free(s->off); free(s->cat); memset(s, 0, sizeof *s);
}

/* ------------------------------------------------------------------ run object */

typedef struct { int32_t r, q, len, tag; } anchor;               /* 1-based positions */
typedef struct { int64_t sA, sB, len; } match_t;                 /* postnuc Match   */
typedef struct { int wasFused; int dirB; int nm; match_t *m; } cluster_t;
typedef struct {
    int dirB; int64_t sA, sB, eA, eB; vec64 delta; int64_t deltaApos;
    int64_t Errors, SimErrors, NonAlphas;
} align_t;

struct pmo_run {
    pmo_opts o;
    seqset ref, qry;
    int stage;
    int32_t *sa, *lcp;
    anchor *anc; int64_t nanc;
    vec32 cl_m, cl_off, cl_tag;        /* mgaps output */
    vec64 al, doff, dl;                /* alignments   */
    strbuf text;
    int64_t dp_cells;
};

/* ------------------------------------------------------------------ 3 index */

/* Suffix order: lexicographic over END < a < c < g < t < X_p, where every X (non-acgt
 * base or separator) at text position p is its own symbol, ordered by p.  A match can
 * never run through an X, so what lies behind the first X of a suffix is irrelevant to
 * seeding; making X unique also makes all suffixes distinct without a sentinel. */
static __thread const uint8_t *g_txt; static __thread int64_t g_n;

static int suffix_cmp(const void *pa, const void *pb)
{
    int64_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
    if (a == b) return 0;
    const uint8_t *t = g_txt; int64_t n = g_n;
    for (int64_t k = 0;; k++) {
        int ea = a + k >= n, eb = b + k >= n;
        if (ea || eb) return ea ? -1 : 1;            /* the shorter suffix is smaller; both cannot end together */
        int ca = t[a + k], cb = t[b + k];
        if (ca != cb) return ca < cb ? -1 : 1;
        if (ca == PMN_CODE_X) return a < b ? -1 : 1;  /* X_p ordered by position */
    }
}

/* number of leading positions where both texts hold the same a/c/g/t */
static int64_t lcp_rr(const uint8_t *t, int64_t n, int64_t a, int64_t b)
{
    int64_t k = 0;
    while (a + k < n && b + k < n && t[a + k] == t[b + k] && t[a + k] != PMN_CODE_X) k++;
    return k;
}

int pmo_stage_index(pmo_run *r)
{
    if (r->stage >= 1) return 0;
    int64_t n = r->ref.ncat;
    r->sa = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n);
    r->lcp = (int32_t *)xmalloc(sizeof(int32_t) * (size_t)n);
    for (int64_t i = 0; i < n; i++) r->sa[i] = (int32_t)i;
    g_txt = r->ref.cat; g_n = n;
    qsort(r->sa, (size_t)n, sizeof(int32_t), suffix_cmp);
    for (int64_t i = 0; i < n; i++)
        r->lcp[i] = i == 0 ? 0 : (int32_t)lcp_rr(r->ref.cat, n, r->sa[i - 1], r->sa[i]);
    r->stage = 1;
    return 0;
}

/* ------------------------------------------------------------------ 4 seeding */

/* mummer -mumreference -l minmatch -n (-b): for every query position i the longest
 * prefix of Q[i..] that occurs in the reference; reported iff it is at least minmatch
 * long, occurs at exactly ONE reference position, and cannot be extended to the left.
 * Uniqueness in the query is not required. */

/* compare Q[i..] with the reference suffix s in the order of section 3.  A query base
 * of code X compares greater than every reference symbol. */
static int qsuffix_cmp(const uint8_t *q, int64_t m, int64_t i, const uint8_t *t, int64_t n, int64_t s)
{
    for (int64_t k = 0;; k++) {
        int eq = i + k >= m, er = s + k >= n;
        if (eq && er) return 0;
        if (eq) return -1;
        if (er) return 1;
        int cq = q[i + k], cr = t[s + k];
        if (cq == PMN_CODE_X) return 1;
        if (cq != cr) return cq < cr ? -1 : 1;
    }
}

static int64_t lcp_qr(const uint8_t *q, int64_t m, int64_t i, const uint8_t *t, int64_t n, int64_t s)
{
    int64_t k = 0;
    while (i + k < m && s + k < n && q[i + k] == t[s + k] && q[i + k] != PMN_CODE_X) k++;
    return k;
}

static void revcomp_codes(const uint8_t *in, int64_t m, uint8_t *out)
{
    for (int64_t i = 0; i < m; i++) { uint8_t c = in[m - 1 - i]; out[i] = c < 4 ? (uint8_t)(3 - c) : c; }
}

static void seed_strand(pmo_run *r, const uint8_t *q, int64_t m, int tag, int64_t *cap)
{
    const uint8_t *t = r->ref.cat; int64_t n = r->ref.ncat; const int32_t *sa = r->sa;
    for (int64_t i = 0; i + r->o.minmatch <= m; i++) {
        /* first SA slot whose suffix is >= Q[i..] */
        int64_t lo = 0, hi = n;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (qsuffix_cmp(q, m, i, t, n, sa[mid]) > 0) lo = mid + 1; else hi = mid; }
        int64_t p = lo;
        int64_t l1 = p > 0 ? lcp_qr(q, m, i, t, n, sa[p - 1]) : -1;
        int64_t l2 = p < n ? lcp_qr(q, m, i, t, n, sa[p]) : -1;
        int64_t L = l1 > l2 ? l1 : l2;
        if (L < r->o.minmatch) continue;
        /* count reference positions sharing those L bases */
        int64_t cnt = 0, where = -1;
        for (int64_t k = p - 1; k >= 0 && lcp_qr(q, m, i, t, n, sa[k]) >= L; k--) { cnt++; where = sa[k]; if (cnt > 1) break; }
        for (int64_t k = p; k < n && cnt <= 1 && lcp_qr(q, m, i, t, n, sa[k]) >= L; k++) { cnt++; where = sa[k]; }
        if (cnt != 1) continue;
        /* left-maximal? */
        if (i > 0 && where > 0) { int cq = q[i - 1], cr = t[where - 1]; if (cq == cr && cq != PMN_CODE_X) continue; }
        if (r->nanc == *cap) { *cap = *cap ? *cap * 2 : 1024; r->anc = (anchor *)xrealloc(r->anc, sizeof(anchor) * (size_t)*cap); }
        anchor a = { (int32_t)(where + 1), (int32_t)(i + 1), (int32_t)L, tag };
        r->anc[r->nanc++] = a;
    }
}

static int anchor_cmp(const void *pa, const void *pb)
{
    const anchor *a = (const anchor *)pa, *b = (const anchor *)pb;
    if (a->tag != b->tag) return a->tag < b->tag ? -1 : 1;
    if (a->q != b->q) return a->q < b->q ? -1 : 1;
    if (a->r != b->r) return a->r < b->r ? -1 : 1;
    return 0;
}

int pmo_stage_seed(pmo_run *r)
{
    if (r->stage >= 2) return 0;
    if (pmo_stage_index(r)) return -1;
    int64_t cap = 0;
    for (int rec = 0; rec < r->qry.nrec; rec++) {
        const uint8_t *q = r->qry.cat + r->qry.off[rec]; int64_t m = r->qry.len[rec];
        if (r->o.do_forward) seed_strand(r, q, m, rec * 2, &cap);
        if (r->o.do_reverse) {
            uint8_t *rc = (uint8_t *)xmalloc((size_t)m); revcomp_codes(q, m, rc);
            seed_strand(r, rc, m, rec * 2 + 1, &cap);
            free(rc);
        }
    }
    qsort(r->anc, (size_t)r->nanc, sizeof(anchor), anchor_cmp);
    r->stage = 2;
    return 0;
}

/* ------------------------------------------------------------------ 5 clustering (mgaps) */

typedef struct {
    int64_t Start1, Start2, Len;
    int64_t Simple_Score, Simple_From, Simple_Adj;
    int64_t cluster_id; int Good, Tentative;
} mg_match;

static int by_start2(const void *pa, const void *pb)
{
    const mg_match *a = (const mg_match *)pa, *b = (const mg_match *)pb;
    if (a->Start2 != b->Start2) return a->Start2 < b->Start2 ? -1 : 1;
    if (a->Start1 != b->Start1) return a->Start1 < b->Start1 ? -1 : 1;
    return 0;
}
static int by_cluster(const void *pa, const void *pb)
{
    const mg_match *a = (const mg_match *)pa, *b = (const mg_match *)pb;
    if (a->cluster_id != b->cluster_id) return a->cluster_id < b->cluster_id ? -1 : 1;
    return by_start2(pa, pb);
}

/* mgaps Filter_Matches: drop matches internal to a repeat and merge overlapping matches
 * on one diagonal.  A[1..N], returns the new N.  Sequential by construction. */
static int64_t filter_matches(mg_match *A, int64_t N)
{
    int64_t i, j;
    for (i = 1; i <= N; i++) { A[i].Good = 1; A[i].Tentative = 0; }
    qsort(A + 1, (size_t)N, sizeof(mg_match), by_start2);
    for (i = 1; i < N; i++) {
        if (!A[i].Good) continue;
        int64_t i_diag = A[i].Start2 - A[i].Start1;
        int64_t i_end = A[i].Start2 + A[i].Len;
        for (j = i + 1; j <= N && A[j].Start2 <= i_end; j++) {
            if (!A[j].Good) continue;
            int64_t j_diag = A[j].Start2 - A[j].Start1;
            if (i_diag == j_diag) {
                int64_t j_extent = A[j].Len + A[j].Start2 - A[i].Start2;
                if (j_extent > A[i].Len) { A[i].Len = j_extent; i_end = A[i].Start2 + j_extent; }
                A[j].Good = 0;
            } else if (A[i].Start1 == A[j].Start1 || A[i].Start2 == A[j].Start2) {
                int64_t olap = A[i].Start1 == A[j].Start1
                             ? A[i].Start2 + A[i].Len - A[j].Start2
                             : A[i].Start1 + A[i].Len - A[j].Start1;
                if (A[i].Len < A[j].Len) {
                    if (olap >= A[i].Len / 2) { A[i].Good = 0; break; }
                } else if (A[j].Len < A[i].Len) {
                    if (olap >= A[j].Len / 2) A[j].Good = 0;
                } else if (olap >= A[i].Len / 2) {
                    A[j].Tentative = 1;
                    if (A[i].Tentative) { A[i].Good = 0; break; }
                }
            }
        }
    }
    for (i = j = 1; i <= N; i++) if (A[i].Good) { if (i != j) A[j] = A[i]; j++; }
    N = j - 1;
    for (i = 1; i <= N; i++) A[i].Good = 0;
    return N;
}

static int64_t uf_find(int64_t *UF, int64_t a)
{
    int64_t root = a; while (UF[root] >= 0) root = UF[root];
    while (UF[a] >= 0) { int64_t nx = UF[a]; UF[a] = root; a = nx; }
    return root;
}
static void uf_union(int64_t *UF, int64_t a, int64_t b)
{
    if (a == b) return;
    if (UF[a] < UF[b]) { UF[a] += UF[b]; UF[b] = a; } else { UF[b] += UF[a]; UF[a] = b; }
}

/* mgaps Process_Matches on one connected component A[1..N]: repeatedly take the best
 * colinear chain, emit it if long enough, drop it, until nothing is left. */
static void process_matches(pmo_run *r, mg_match *A, int64_t N, int tag)
{
    int64_t i, j, k;
    int64_t *pm = r->o.fast_chain ? (int64_t *)xmalloc(sizeof(int64_t) * (size_t)(N + 1)) : NULL;
    do {
        for (i = 1; i <= N; i++) {
            A[i].Simple_Score = A[i].Len; A[i].Simple_Adj = 0; A[i].Simple_From = -1;
            if (!r->o.fast_chain) {
                for (j = 1; j < i; j++) {
                    int64_t Olap1 = A[j].Start1 + A[j].Len - A[i].Start1;
                    int64_t Olap = Olap1 > 0 ? Olap1 : 0;
                    int64_t Olap2 = A[j].Start2 + A[j].Len - A[i].Start2;
                    if (Olap2 > Olap) Olap = Olap2;
                    int64_t dd = (A[i].Start2 - A[i].Start1) - (A[j].Start2 - A[j].Start1);
                    int64_t Pen = Olap + (dd < 0 ? -dd : dd);
                    if (A[j].Simple_Score + A[i].Len - Pen > A[i].Simple_Score) {
                        A[i].Simple_From = j; A[i].Simple_Score = A[j].Simple_Score + A[i].Len - Pen; A[i].Simple_Adj = Olap;
                    }
                }
            } else {
                /* Same argmax (lowest j among the best) found by scanning j downward and
                 * stopping once no earlier score can reach the best so far: Pen >= 0, so
                 * candidate j is worth at most Simple_Score[j] + Len[i]. */
                int64_t best = A[i].Len, from = -1, adj = 0;
                for (j = i - 1; j >= 1; j--) {
                    if (pm[j] + A[i].Len < best || (from == -1 && pm[j] + A[i].Len <= best)) break;
                    int64_t Olap1 = A[j].Start1 + A[j].Len - A[i].Start1;
                    int64_t Olap = Olap1 > 0 ? Olap1 : 0;
                    int64_t Olap2 = A[j].Start2 + A[j].Len - A[i].Start2;
                    if (Olap2 > Olap) Olap = Olap2;
                    int64_t dd = (A[i].Start2 - A[i].Start1) - (A[j].Start2 - A[j].Start1);
                    int64_t Pen = Olap + (dd < 0 ? -dd : dd);
                    int64_t v = A[j].Simple_Score + A[i].Len - Pen;
                    if (v > best || (v == best && from != -1)) { best = v; from = j; adj = Olap; }
                }
                A[i].Simple_Score = best; A[i].Simple_From = from; A[i].Simple_Adj = adj;
                pm[i] = i == 1 || best > pm[i - 1] ? best : pm[i - 1];
            }
        }
        int64_t best = 1;
        for (i = 2; i <= N; i++) if (A[i].Simple_Score > A[best].Simple_Score) best = i;
        int64_t total = 0;
        for (i = best; i > 0; i = A[i].Simple_From) { A[i].Good = 1; total += A[i].Len; }
        if (total >= r->o.mincluster) {
            int64_t prev = -1; int nout = 0;
            for (i = 1; i <= N; i++) if (A[i].Good) {
                int64_t adj = prev == -1 ? 0 : A[i].Simple_Adj;
                /* a match trimmed away completely is not passed on (ORACLE_SPEC.md §4) */
                if (A[i].Len - adj >= 1) {
                    v32_push(&r->cl_m, (int32_t)(A[i].Start1 + adj)); v32_push(&r->cl_m, (int32_t)(A[i].Start2 + adj));
                    v32_push(&r->cl_m, (int32_t)(A[i].Len - adj)); nout++;
                }
                prev = i;
            }
            if (nout) { v32_push(&r->cl_off, (int32_t)(r->cl_m.n / 3)); v32_push(&r->cl_tag, tag); }
        }
        for (i = k = 1; i <= N; i++) if (!A[i].Good) { if (i != k) A[k] = A[i]; k++; }
        N = k - 1;
    } while (N > 0);
    free(pm);
}

/* mgaps Process_Cluster for the anchors of one (query record, strand). */
static void mgaps_section(pmo_run *r, const anchor *a, int64_t n0, int tag)
{
    if (n0 == 0) return;
    mg_match *A = (mg_match *)xmalloc(sizeof(mg_match) * (size_t)(n0 + 1));
    memset(A, 0, sizeof(mg_match) * (size_t)(n0 + 1));
    for (int64_t i = 0; i < n0; i++) { A[i + 1].Start1 = a[i].r; A[i + 1].Start2 = a[i].q; A[i + 1].Len = a[i].len; }
    int64_t N = filter_matches(A, n0);
    int64_t *UF = (int64_t *)xmalloc(sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t i = 1; i <= N; i++) UF[i] = -1;
    for (int64_t i = 1; i < N; i++) {
        int64_t i_end = A[i].Start2 + A[i].Len, i_diag = A[i].Start2 - A[i].Start1;
        for (int64_t j = i + 1; j <= N; j++) {
            int64_t sep = A[j].Start2 - i_end;
            if (sep > r->o.maxgap) break;
            int64_t dd = (A[j].Start2 - A[j].Start1) - i_diag; if (dd < 0) dd = -dd;
            int64_t lim = (int64_t)(r->o.diagfactor * (double)sep);
            if (lim < r->o.diagdiff) lim = r->o.diagdiff;
            if (dd <= lim) uf_union(UF, uf_find(UF, i), uf_find(UF, j));
        }
    }
    /* component id = smallest member index, so the output order does not depend on the
     * union order (ORACLE_SPEC.md §4) */
    int64_t *minidx = (int64_t *)xmalloc(sizeof(int64_t) * (size_t)(N + 1));
    for (int64_t i = 1; i <= N; i++) minidx[i] = 0;
    for (int64_t i = 1; i <= N; i++) { int64_t root = uf_find(UF, i); if (!minidx[root]) minidx[root] = i; A[i].cluster_id = minidx[root]; }
    qsort(A + 1, (size_t)N, sizeof(mg_match), by_cluster);
    for (int64_t i = 1; i <= N;) {
        int64_t j = i + 1; while (j <= N && A[j].cluster_id == A[i].cluster_id) j++;
        process_matches(r, A + i - 1, j - i, tag);
        i = j;
    }
    free(minidx); free(UF); free(A);
}

int pmo_stage_cluster(pmo_run *r)
{
    if (r->stage >= 3) return 0;
    if (pmo_stage_seed(r)) return -1;
    v32_push(&r->cl_off, 0);
    for (int64_t i = 0; i < r->nanc;) {
        int64_t j = i; while (j < r->nanc && r->anc[j].tag == r->anc[i].tag) j++;
        mgaps_section(r, r->anc + i, j - i, r->anc[i].tag);
        i = j;
    }
    r->stage = 3;
    return 0;
}

/* ------------------------------------------------------------------ 6 alignment engine (sw_align) */

/* Three-state affine DP evaluated by anti-diagonals with a score-trimmed band.
 *
 * The window of A is A0[Astart .. Aend] (forward) or A0[Astart .. Aend] walked downward
 * (backward), N bases; likewise B, M bases.  Cell (i,j), 0<=i<=N, 0<=j<=M, holds the best
 * scores of aligning the first i window bases of A with the first j of B, ending in
 *   INS (last column: A base over a gap), DEL (gap over a B base) or MAT (base over base).
 * Cell (0,0) is MAT = 0.  Anti-diagonal d = i + j, cells are addressed by j.
 *
 * Band: diagonal d computes j in [max(tlo, d-N), min(thi+1, M)] where [tlo,thi] is what
 * survived trimming on d-1.  Trimming (not when FORCED) keeps the span between the first
 * and last cell whose best state is within GOOD_SCORE*breaklen of the running high score.
 * A predecessor outside the computed range of its diagonal counts as PMN_NEG.
 * Stop (not when FORCED) once breaklen diagonals passed without a new high score
 * (">=": a later equal score moves the finish point).
 *
 * Finish cell: the target (N,M) if it was reached and OPTIMAL is clear, else the high
 * score cell.  Aend/Bend are updated to the finish cell.  Unless SEARCH, the path is
 * traced back and appended to Delta in MUMmer's encoding. */

typedef struct { int32_t v[3]; } dpcell;

static inline void score_edit(int32_t del, int32_t ins, int32_t mat, int32_t *val, int *used)
{
    if (del > ins) { if (del > mat) { *val = del; *used = PMN_ST_DEL; } else { *val = mat; *used = PMN_ST_MAT; } }
    else if (ins > mat) { *val = ins; *used = PMN_ST_INS; }
    else { *val = mat; *used = PMN_ST_MAT; }
}
static inline int max_state(const int32_t *v)
{
    if (v[PMN_ST_DEL] > v[PMN_ST_INS]) return v[PMN_ST_DEL] > v[PMN_ST_MAT] ? PMN_ST_DEL : PMN_ST_MAT;
    return v[PMN_ST_INS] > v[PMN_ST_MAT] ? PMN_ST_INS : PMN_ST_MAT;
}

static int align_engine(pmo_run *r, const uint8_t *A0, int64_t Astart, int64_t *Aend,
                        const uint8_t *B0, int64_t Bstart, int64_t *Bend, vec64 *Delta, unsigned m_o)
{
    const int dir = (m_o & PMN_DIRECTION_BIT) ? 1 : -1;
    const int64_t N = dir > 0 ? *Aend - Astart + 1 : Astart - *Aend + 1;
    const int64_t M = dir > 0 ? *Bend - Bstart + 1 : Bstart - *Bend + 1;
    const int forced = (m_o & PMN_FORCED_BIT) != 0, search = (m_o & PMN_SEARCH_BIT) != 0;
    const int64_t breaklen = r->o.breaklen;
    const int32_t max_diff = (int32_t)(PMN_GOOD_SCORE * breaklen);
    if (N < 1 || M < 1) { fprintf(stderr, "pmn_oracle: align_engine called with an empty window\n"); abort(); }

    dpcell *buf[3];
    for (int k = 0; k < 3; k++) buf[k] = (dpcell *)xmalloc(sizeof(dpcell) * (size_t)(M + 1));
    dpcell *pp = buf[0], *p = buf[1], *c = buf[2];
    int64_t pplo = 1, pphi = 0, plo = 0, phi = 0;          /* computed ranges of d-2, d-1 */
    p[0].v[PMN_ST_DEL] = PMN_NEG; p[0].v[PMN_ST_INS] = PMN_NEG; p[0].v[PMN_ST_MAT] = 0;
    int64_t tlo = 0, thi = 0;

    /* traceback: one byte per cell = used[DEL] | used[INS]<<2 | used[MAT]<<4 | maxstate<<6 */
    uint8_t *tb = NULL; size_t tbn = 0, tbcap = 0; int64_t *tboff = NULL, *tblo = NULL;
    if (!search) {
        tboff = (int64_t *)xmalloc(sizeof(int64_t) * (size_t)(N + M + 1));
        tblo = (int64_t *)xmalloc(sizeof(int64_t) * (size_t)(N + M + 1));
        tbcap = 1024; tb = (uint8_t *)xmalloc(tbcap);
        tboff[0] = 0; tblo[0] = 0; tb[tbn++] = (uint8_t)(PMN_ST_NONE | PMN_ST_NONE << 2 | PMN_ST_NONE << 4 | PMN_ST_MAT << 6);
    }

    int32_t high = 0; int64_t best_d = 0, best_j = 0; int reached = 0; int64_t d;
    for (d = 1; d <= N + M; d++) {
        if (!forced && d - best_d > breaklen) break;
        int64_t clo = tlo > d - N ? tlo : d - N, chi = thi + 1 < M ? thi + 1 : M;
        if (clo > chi) break;
        if (!search) {
            tboff[d] = (int64_t)tbn; tblo[d] = clo;
            if (tbn + (size_t)(chi - clo + 1) > tbcap) { tbcap = (tbcap + (size_t)(chi - clo + 1)) * 2; tb = (uint8_t *)xrealloc(tb, tbcap); }
        }
        int32_t dmax = INT32_MIN; int64_t dmaxj = -1;
        for (int64_t j = clo; j <= chi; j++) {
            int64_t i = d - j;
            int32_t U[3] = { PMN_NEG, PMN_NEG, PMN_NEG }, L[3] = { PMN_NEG, PMN_NEG, PMN_NEG }, P[3] = { PMN_NEG, PMN_NEG, PMN_NEG };
            if (j >= plo && j <= phi) memcpy(U, p[j].v, sizeof U);                    /* (i-1, j)   */
            if (j - 1 >= plo && j - 1 <= phi) memcpy(L, p[j - 1].v, sizeof L);        /* (i,   j-1) */
            if (j - 1 >= pplo && j - 1 <= pphi) memcpy(P, pp[j - 1].v, sizeof P);     /* (i-1, j-1) */
            int32_t s = PMN_BAD_SCORE;
            if (i >= 1 && j >= 1) {
                int ca = A0[Astart + dir * (i - 1)], cb = B0[Bstart + dir * (j - 1)];
                if (ca == cb && ca != PMN_CODE_X) s = PMN_GOOD_SCORE;
            }
            int uD, uI, uM; int32_t vD, vI, vM;
            score_edit(L[PMN_ST_DEL] + PMN_CONT_GAP_SCORE, L[PMN_ST_INS] + PMN_OPEN_GAP_SCORE, L[PMN_ST_MAT] + PMN_OPEN_GAP_SCORE, &vD, &uD);
            score_edit(U[PMN_ST_DEL] + PMN_OPEN_GAP_SCORE, U[PMN_ST_INS] + PMN_CONT_GAP_SCORE, U[PMN_ST_MAT] + PMN_OPEN_GAP_SCORE, &vI, &uI);
            score_edit(P[PMN_ST_DEL] + s, P[PMN_ST_INS] + s, P[PMN_ST_MAT] + s, &vM, &uM);
            c[j].v[PMN_ST_DEL] = vD; c[j].v[PMN_ST_INS] = vI; c[j].v[PMN_ST_MAT] = vM;
            int ms = max_state(c[j].v); int32_t cm = c[j].v[ms];
            if (!search) tb[tbn++] = (uint8_t)(uD | uI << 2 | uM << 4 | ms << 6);
            if (cm >= dmax) { dmax = cm; dmaxj = j; }
        }
        r->dp_cells += chi - clo + 1;
        if (dmax >= high) { high = dmax; best_d = d; best_j = dmaxj; }
        if (d == N + M) { reached = 1; d++; break; }
        if (!forced) {
            int32_t t = high - max_diff;
            tlo = clo; while (tlo <= chi && c[tlo].v[max_state(c[tlo].v)] < t) tlo++;
            thi = chi; while (thi >= tlo && c[thi].v[max_state(c[thi].v)] < t) thi--;
        } else { tlo = clo; thi = chi; }
        dpcell *x = pp; pp = p; p = c; c = x;
        pplo = plo; pphi = phi; plo = clo; phi = chi;
    }

    int64_t fd, fj;
    if (reached && !(m_o & PMN_OPTIMAL_BIT)) { fd = N + M; fj = M; } else { fd = best_d; fj = best_j; }
    int64_t fi = fd - fj;
    *Aend = Astart + dir * (fi - 1);
    *Bend = Bstart + dir * (fj - 1);

    if (!search) {
        /* walk back from the finish cell, then emit deltas front to back */
        size_t nops = 0; uint8_t *ops = (uint8_t *)xmalloc((size_t)(fd + 1));
        int64_t cd = fd, cj = fj; int st = tb[tboff[cd] + (cj - tblo[cd])] >> 6;
        while (cd > 0) {
            uint8_t b = tb[tboff[cd] + (cj - tblo[cd])];
            ops[nops++] = (uint8_t)st;
            if (st == PMN_ST_MAT) { st = (b >> 4) & 3; cd -= 2; cj -= 1; }
            else if (st == PMN_ST_INS) { st = (b >> 2) & 3; cd -= 1; }
            else { st = b & 3; cd -= 1; cj -= 1; }
        }
        int64_t count = 1;
        for (size_t k = nops; k-- > 0;) {
            if (ops[k] == PMN_ST_MAT) count++;
            else if (ops[k] == PMN_ST_INS) { v64_push(Delta, count); count = 1; }
            else { v64_push(Delta, -count); count = 1; }
        }
        free(ops); free(tb); free(tboff); free(tblo);
    }
    for (int k = 0; k < 3; k++) free(buf[k]);
    return reached;
}

/* ------------------------------------------------------------------ 7 extension (postnuc) */

typedef struct {
    pmo_run *r;
    const uint8_t *A; int64_t lenA;          /* 1-based: A[1..lenA] */
    const uint8_t *Bf, *Br; int64_t lenB;
    cluster_t *C; int nC;
    align_t *Al; int nAl, capAl;
} synteny_ctx;

static int by_cluster_sA(const void *pa, const void *pb)
{
    /* AscendingClusterSort, made a total order: first match sA, then input order (wasFused
     * holds the input rank during the sort) */
    const cluster_t *a = (const cluster_t *)pa, *b = (const cluster_t *)pb;
    if (a->m[0].sA != b->m[0].sA) return a->m[0].sA < b->m[0].sA ? -1 : 1;
    return a->wasFused < b->wasFused ? -1 : (a->wasFused > b->wasFused);
}

static void add_new_alignment(synteny_ctx *x, int cp, int mp)
{
    if (x->nAl == x->capAl) { x->capAl = x->capAl ? x->capAl * 2 : 16; x->Al = (align_t *)xrealloc(x->Al, sizeof(align_t) * (size_t)x->capAl); }
    align_t *a = &x->Al[x->nAl++]; memset(a, 0, sizeof *a);
    const match_t *m = &x->C[cp].m[mp];
    a->sA = m->sA; a->sB = m->sB; a->eA = m->sA + m->len - 1; a->eB = m->sB + m->len - 1; a->dirB = x->C[cp].dirB;
}

static int is_shadowed_cluster(const synteny_ctx *x, int cp, int ap)
{
    const cluster_t *c = &x->C[cp];
    int64_t sA = c->m[0].sA, eA = c->m[c->nm - 1].sA + c->m[c->nm - 1].len - 1;
    int64_t sB = c->m[0].sB, eB = c->m[c->nm - 1].sB + c->m[c->nm - 1].len - 1;
    for (int i = ap; i >= 0; i--) {
        const align_t *a = &x->Al[i];
        if (a->dirB == c->dirB && a->eA >= eA && a->eB >= eB && a->sA <= sA && a->sB <= sB) return 1;
    }
    return 0;
}

/* returns nC when there is no suitable cluster; targetA/B then stay unchanged */
static int get_forward_target_cluster(const synteny_ctx *x, int cp, int64_t *targetA, int64_t *targetB)
{
    const cluster_t *c = &x->C[cp];
    int64_t sA = c->m[c->nm - 1].sA + c->m[c->nm - 1].len - 1;
    int64_t sB = c->m[c->nm - 1].sB + c->m[c->nm - 1].len - 1;
    int64_t dist = *targetA - sA < *targetB - sB ? *targetA - sA : *targetB - sB;
    int best = x->nC;
    for (int ci = cp + 1; ci < x->nC; ci++) {
        const cluster_t *t = &x->C[ci];
        if (t->dirB != c->dirB) continue;
        int64_t eA = t->m[0].sA, eB = t->m[0].sB;
        /* the cluster overlaps the current one: skip its leading matches */
        if ((eA < sA || eB < sB) && t->m[t->nm - 1].sA >= sA && t->m[t->nm - 1].sB >= sB)
            for (int k = 0; k < t->nm && (eA < sA || eB < sB); k++) { eA = t->m[k].sA; eB = t->m[k].sB; }
        if (eA >= sA && eB >= sB) {
            int64_t greater, lesser;
            if (eA - sA > eB - sB) { greater = eA - sA; lesser = eB - sB; } else { lesser = eA - sA; greater = eB - sB; }
            if (greater < x->r->o.breaklen || lesser * PMN_GOOD_SCORE + (greater - lesser) * PMN_CONT_GAP_SCORE >= 0) {
                best = ci; *targetA = eA; *targetB = eB; break;
            } else if ((greater << 1) - lesser < dist) {
                best = ci; *targetA = eA; *targetB = eB; dist = (greater << 1) - lesser;
            }
        }
    }
    return best;
}

/* returns -1 when there is none */
static int get_reverse_target_alignment(const synteny_ctx *x, int ap)
{
    const align_t *c = &x->Al[ap];
    int64_t sA = c->sA, sB = c->sB;
    int64_t dist = sA < sB ? sA : sB;
    int best = -1;
    for (int i = ap - 1; i >= 0; i--) {
        const align_t *a = &x->Al[i];
        if (a->dirB != c->dirB) continue;
        int64_t eA = a->eA, eB = a->eB;
        if (eA <= sA && eB <= sB) {
            int64_t greater, lesser;
            if (sA - eA > sB - eB) { greater = sA - eA; lesser = sB - eB; } else { lesser = sA - eA; greater = sB - eB; }
            if (greater < x->r->o.breaklen || lesser * PMN_GOOD_SCORE + (greater - lesser) * PMN_CONT_GAP_SCORE >= 0) { best = i; break; }
            else if ((greater << 1) - lesser < dist) { best = i; dist = (greater << 1) - lesser; }
        }
    }
    return best;
}

static int extend_forward(synteny_ctx *x, int ap, const uint8_t *B, int64_t targetA, int64_t targetB, unsigned m_o)
{
    align_t *a = &x->Al[ap];
    int overflow = 0, dbl = 0;
    int64_t Di = a->delta.n;
    int64_t ValA = targetA - a->eA + 1, ValB = targetB - a->eB + 1;
    if (ValA > PMN_MAX_ALIGNMENT_LENGTH) { targetA = a->eA + PMN_MAX_ALIGNMENT_LENGTH - 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (ValB > PMN_MAX_ALIGNMENT_LENGTH) { targetB = a->eB + PMN_MAX_ALIGNMENT_LENGTH - 1; if (overflow) dbl = 1; else overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (dbl) m_o &= ~(unsigned)PMN_SEQEND_BIT;
    int reached = align_engine(x->r, x->A, a->eA, &targetA, B, a->eB, &targetB, &a->delta, m_o);
    if (reached && overflow) reached = 0;
    if (Di < a->delta.n) {
        /* the first new delta counts from the engine's start column: add the bases the
         * alignment already holds since its last indel */
        ValA = (a->eA - a->sA + 1) - a->deltaApos - 1;
        a->delta.v[Di] += a->delta.v[Di] > 0 ? ValA : -ValA;
        for (int64_t k = Di; k < a->delta.n; k++) a->deltaApos += a->delta.v[k] > 0 ? a->delta.v[k] : -a->delta.v[k] - 1;
    }
    a->eA = targetA; a->eB = targetB;
    return reached;
}

/* may pop the last alignment (when it was merged into `tp`) */
static int extend_backward(synteny_ctx *x, int ap, int tp, const uint8_t *B)
{
    align_t *a = &x->Al[ap];
    int overflow = 0, dbl = 0; unsigned m_o = PMN_BACKWARD_SEARCH;
    int64_t targetA, targetB;
    if (tp >= 0) { targetA = x->Al[tp].eA; targetB = x->Al[tp].eB; }
    else { targetA = 1; targetB = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (a->sA - targetA + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetA = a->sA - PMN_MAX_ALIGNMENT_LENGTH + 1; overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    if (a->sB - targetB + 1 > PMN_MAX_ALIGNMENT_LENGTH) { targetB = a->sB - PMN_MAX_ALIGNMENT_LENGTH + 1; if (overflow) dbl = 1; else overflow = 1; m_o |= PMN_OPTIMAL_BIT; }
    (void)dbl;
    int reached = align_engine(x->r, x->A, a->sA, &targetA, B, a->sB, &targetB, NULL, m_o);
    if (overflow || tp < 0) reached = 0;
    if (reached) {
        extend_forward(x, tp, B, a->sA, a->sB, PMN_FORCED_FORWARD_ALIGN);
        a = &x->Al[ap];
        x->Al[tp].eA += a->eA - a->sA; x->Al[tp].eB += a->eB - a->sB;
        free(a->delta.v); x->nAl--;
    } else {
        int64_t eA = a->sA, eB = a->sB;
        align_engine(x->r, x->A, targetA, &eA, B, targetB, &eB, &a->delta, PMN_FORCED_FORWARD_ALIGN);
        a->sA = targetA; a->sB = targetB;
        for (int64_t k = 0; k < a->delta.n; k++) a->deltaApos += a->delta.v[k] > 0 ? a->delta.v[k] : -a->delta.v[k] - 1;
    }
    return reached;
}

static int extend_clusters(synteny_ctx *x)
{
    const pmo_opts *o = &x->r->o;
    for (int i = 0; i < x->nC; i++) x->C[i].wasFused = i;
    qsort(x->C, (size_t)x->nC, sizeof(cluster_t), by_cluster_sA);
    for (int i = 0; i < x->nC; i++) x->C[i].wasFused = 0;

    int target_reached = 0, CurrCp = 0, PrevCp = 0, TargetCp = x->nC, CurrAp = -1;
    while (CurrCp < x->nC) {
        cluster_t *c = &x->C[CurrCp];
        if (o->do_extend && !target_reached && c->wasFused) { CurrCp++; continue; }
        const uint8_t *B = c->dirB ? x->Br : x->Bf;
        if (!target_reached && o->do_simplify && is_shadowed_cluster(x, CurrCp, x->nAl - 1)) {
            c->wasFused = 1; CurrCp = ++PrevCp; continue;
        }
        int CurrMp = 0;
        while (CurrMp < c->nm) {
            if (target_reached) {
                /* the alignment already reaches this match: absorb it */
                if (x->Al[CurrAp].eA != c->m[CurrMp].sA || x->Al[CurrAp].eB != c->m[CurrMp].sB) {
                    if (CurrMp >= c->nm - 1) return fail("extend: target match does not exist");
                    CurrMp++; continue;
                }
                x->Al[CurrAp].eA += c->m[CurrMp].len - 1; x->Al[CurrAp].eB += c->m[CurrMp].len - 1;
            } else {
                add_new_alignment(x, CurrCp, CurrMp); CurrAp = x->nAl - 1;
                if (o->do_extend || CurrMp != 0) {
                    int TargetAp = get_reverse_target_alignment(x, CurrAp);
                    if (extend_backward(x, CurrAp, TargetAp, B)) CurrAp = TargetAp;
                }
            }
            unsigned m_o = PMN_FORWARD_ALIGN;
            if (CurrMp < c->nm - 1) {
                target_reached = extend_forward(x, CurrAp, B, c->m[CurrMp + 1].sA, c->m[CurrMp + 1].sB, m_o);
            } else if (o->do_extend) {
                int64_t targetA = x->lenA, targetB = x->lenB;
                TargetCp = get_forward_target_cluster(x, CurrCp, &targetA, &targetB);
                if (TargetCp == x->nC) m_o |= PMN_OPTIMAL_BIT;
                target_reached = extend_forward(x, CurrAp, B, targetA, targetB, m_o);
            }
            CurrMp++;
        }
        if (TargetCp == x->nC) target_reached = 0;
        c->wasFused = 1;
        if (!target_reached) CurrCp = ++PrevCp; else CurrCp = TargetCp;
    }
    return 0;
}

/* parseDelta: walk each alignment with its deltas and count mismatching columns */
static void parse_delta(synteny_ctx *x)
{
    for (int k = 0; k < x->nAl; k++) {
        align_t *a = &x->Al[k];
        const uint8_t *B = a->dirB ? x->Br : x->Bf;
        int64_t Apos = a->sA, Bpos = a->sB, Remain = a->eA - a->sA + 1;
        a->Errors = a->SimErrors = a->NonAlphas = 0;
        for (int64_t t = 0; t < a->delta.n; t++) {
            int64_t D = a->delta.v[t], absD = D < 0 ? -D : D, i;
            for (i = 1; i < absD; i++) {
                int ca = x->A[Apos++], cb = B[Bpos++];
                if (ca != cb || ca == PMN_CODE_X) { a->Errors++; a->SimErrors++; }
            }
            Remain -= i - 1;
            a->Errors++; a->SimErrors++;
            if (D > 0) { Apos++; Remain--; } else Bpos++;
        }
        for (int64_t i = 0; i < Remain; i++) {
            int ca = x->A[Apos++], cb = B[Bpos++];
            if (ca != cb || ca == PMN_CODE_X) { a->Errors++; a->SimErrors++; }
        }
    }
}

int pmo_stage_extend(pmo_run *r)
{
    if (r->stage >= 4) return 0;
    if (pmo_stage_cluster(r)) return -1;
    if (!r->o.do_optimize) return fail("--nooptimize is not supported");
    int64_t ncl = r->cl_tag.n;
    v64_push(&r->doff, 0);
    for (int qrec = 0; qrec < r->qry.nrec; qrec++) {
        int64_t lenB = r->qry.len[qrec];
        const uint8_t *bf = r->qry.cat + r->qry.off[qrec];
        uint8_t *br = (uint8_t *)xmalloc((size_t)lenB + 1); revcomp_codes(bf, lenB, br);
        /* postnuc input parsing: clusters of this query record (forward section, then
         * reverse), each split where consecutive matches fall into different reference
         * records; one synteny per reference record */
        int nsyn = 0; int *syn_ref = (int *)xmalloc(sizeof(int) * (size_t)r->ref.nrec);
        synteny_ctx *S = (synteny_ctx *)xmalloc(sizeof(synteny_ctx) * (size_t)r->ref.nrec);
        int *capC = (int *)xmalloc(sizeof(int) * (size_t)r->ref.nrec);
        for (int64_t k = 0; k < ncl; k++) {
            int tag = r->cl_tag.v[k]; if (tag >> 1 != qrec) continue;
            int prev_seq = -1; cluster_t *cur = NULL; int curcap = 0;
            for (int64_t mi = r->cl_off.v[k]; mi < r->cl_off.v[k + 1]; mi++) {
                int64_t sA = r->cl_m.v[mi * 3], sB = r->cl_m.v[mi * 3 + 1], len = r->cl_m.v[mi * 3 + 2];
                int seq = 0; while (seq + 1 < r->ref.nrec && sA > r->ref.off[seq] + r->ref.len[seq]) seq++;
                sA -= r->ref.off[seq];
                if (seq != prev_seq) {
                    int s = 0; while (s < nsyn && syn_ref[s] != seq) s++;
                    if (s == nsyn) {
                        syn_ref[nsyn] = seq; memset(&S[nsyn], 0, sizeof(synteny_ctx)); capC[nsyn] = 0;
                        S[nsyn].r = r; S[nsyn].A = r->ref.cat + r->ref.off[seq] - 1; S[nsyn].lenA = r->ref.len[seq];
                        S[nsyn].Bf = bf - 1; S[nsyn].Br = br - 1; S[nsyn].lenB = lenB; nsyn++;
                    }
                    if (S[s].nC == capC[s]) { capC[s] = capC[s] ? capC[s] * 2 : 8; S[s].C = (cluster_t *)xrealloc(S[s].C, sizeof(cluster_t) * (size_t)capC[s]); }
                    cur = &S[s].C[S[s].nC++]; memset(cur, 0, sizeof *cur); cur->dirB = tag & 1; curcap = 0;
                    prev_seq = seq;
                }
                if (cur->nm == curcap) { curcap = curcap ? curcap * 2 : 8; cur->m = (match_t *)xrealloc(cur->m, sizeof(match_t) * (size_t)curcap); }
                match_t m = { sA, sB, len }; cur->m[cur->nm++] = m;
            }
        }
        /* syntenies of one query record are written in reference-record order (ORACLE_SPEC.md §6) */
        for (int a = 1; a < nsyn; a++)
            for (int b = a; b > 0 && syn_ref[b - 1] > syn_ref[b]; b--) {
                int t = syn_ref[b]; syn_ref[b] = syn_ref[b - 1]; syn_ref[b - 1] = t;
                synteny_ctx ts = S[b]; S[b] = S[b - 1]; S[b - 1] = ts;
                int tc = capC[b]; capC[b] = capC[b - 1]; capC[b - 1] = tc;
            }
        int rc = 0;
        for (int s = 0; s < nsyn && rc == 0; s++) {
            rc = extend_clusters(&S[s]);
            if (rc) break;
            parse_delta(&S[s]);
            for (int k = 0; k < S[s].nAl; k++) {
                align_t *a = &S[s].Al[k];
                int64_t row[10] = { syn_ref[s], qrec, a->dirB, a->sA, a->eA, a->sB, a->eB, a->Errors, a->SimErrors, a->NonAlphas };
                for (int c = 0; c < 10; c++) v64_push(&r->al, row[c]);
                for (int64_t t = 0; t < a->delta.n; t++) v64_push(&r->dl, a->delta.v[t]);
                v64_push(&r->doff, r->dl.n);
            }
        }
        for (int s = 0; s < nsyn; s++) {
            for (int k = 0; k < S[s].nC; k++) free(S[s].C[k].m);
            for (int k = 0; k < S[s].nAl; k++) free(S[s].Al[k].delta.v);
            free(S[s].C); free(S[s].Al);
        }
        free(capC); free(S); free(syn_ref); free(br);
        if (rc) return -1;
    }
    r->stage = 4;
    return 0;
}

/* ------------------------------------------------------------------ 8 .delta writer */

/* Grammar exactly as the reference parses it (lib/profiles_lib/m_delta.cc:72-92 header,
 * :154-162 '>' line, :177-185 seven ints, :187-196 deltas up to "0"; the OCaml parser
 * lib/profiles/m_delta.ml:76-79,91 needs single spaces and no trailing blanks).
 * Reverse-strand alignments get their query coordinates flipped to the forward strand
 * (sQ > eQ), as MUMmer writes them. */
int pmo_stage_delta(pmo_run *r, const char *ref_path, const char *qry_path)
{
    if (r->stage >= 5) return 0;
    if (pmo_stage_extend(r)) return -1;
    sb_printf(&r->text, "%s %s\nNUCMER\n", ref_path, qry_path);
    int64_t nal = r->al.n / 10; int64_t prev_ref = -1, prev_qry = -1;
    for (int64_t k = 0; k < nal; k++) {
        const int64_t *a = r->al.v + k * 10;
        if (a[0] != prev_ref || a[1] != prev_qry) {
            sb_printf(&r->text, ">%s %s %lld %lld\n", r->ref.id[a[0]], r->qry.id[a[1]], (long long)r->ref.len[a[0]], (long long)r->qry.len[a[1]]);
            prev_ref = a[0]; prev_qry = a[1];
        }
        int64_t sB = a[5], eB = a[6], lenB = r->qry.len[a[1]];
        if (a[2]) { sB = lenB - sB + 1; eB = lenB - eB + 1; }
        sb_printf(&r->text, "%lld %lld %lld %lld %lld %lld %lld\n", (long long)a[3], (long long)a[4], (long long)sB, (long long)eB,
                  (long long)a[7], (long long)a[8], (long long)a[9]);
        for (int64_t t = r->doff.v[k]; t < r->doff.v[k + 1]; t++) sb_printf(&r->text, "%lld\n", (long long)r->dl.v[t]);
        sb_printf(&r->text, "0\n");
    }
    r->stage = 5;
    return 0;
}

/* ------------------------------------------------------------------ 9 API */

pmo_run *pmo_run_create(const char *ref_fasta, size_t ref_bytes, const char *qry_fasta, size_t qry_bytes, const pmo_opts *opts)
{
    pmo_run *r = (pmo_run *)xmalloc(sizeof *r); memset(r, 0, sizeof *r);
    if (opts) r->o = *opts; else pmo_default_opts(&r->o);
    if (parse_fasta(ref_fasta, ref_bytes, &r->ref) || parse_fasta(qry_fasta, qry_bytes, &r->qry)) { pmo_run_free(r); return NULL; }
    if (r->o.minmatch < 1) { fail("minmatch must be positive"); pmo_run_free(r); return NULL; }
    return r;
}

void pmo_run_free(pmo_run *r)
{
    if (!r) return;
    free_seqset(&r->ref); free_seqset(&r->qry);
    free(r->sa); free(r->lcp); free(r->anc);
    free(r->cl_m.v); free(r->cl_off.v); free(r->cl_tag.v);
    free(r->al.v); free(r->doff.v); free(r->dl.v); free(r->text.v);
    free(r);
}

int pmo_run_all(pmo_run *r, const char *ref_path, const char *qry_path) { return pmo_stage_delta(r, ref_path, qry_path); }

int64_t pmo_ref_len(const pmo_run *r) { return r->ref.ncat; }
const uint8_t *pmo_ref_codes(const pmo_run *r) { return r->ref.cat; }
const int32_t *pmo_sa(const pmo_run *r) { return r->sa; }
const int32_t *pmo_lcp(const pmo_run *r) { return r->lcp; }
int64_t pmo_n_anchors(const pmo_run *r) { return r->nanc; }
const int32_t *pmo_anchors(const pmo_run *r) { return (const int32_t *)r->anc; }
int64_t pmo_n_clusters(const pmo_run *r) { return r->cl_tag.n; }
int64_t pmo_n_cluster_matches(const pmo_run *r) { return r->cl_m.n / 3; }
const int32_t *pmo_cluster_matches(const pmo_run *r) { return r->cl_m.v; }
const int32_t *pmo_cluster_off(const pmo_run *r) { return r->cl_off.v; }
const int32_t *pmo_cluster_tag(const pmo_run *r) { return r->cl_tag.v; }
int64_t pmo_n_alignments(const pmo_run *r) { return r->al.n / 10; }
const int64_t *pmo_alignments(const pmo_run *r) { return r->al.v; }
const int64_t *pmo_delta_off(const pmo_run *r) { return r->doff.v; }
const int64_t *pmo_deltas(const pmo_run *r) { return r->dl.v; }
const char *pmo_delta_text(const pmo_run *r, size_t *len) { if (len) *len = r->text.n; return r->text.v ? r->text.v : ""; }
int64_t pmo_dp_cells(const pmo_run *r) { return r->dp_cells; }

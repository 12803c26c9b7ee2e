/*
 * pmn_post_oracle.c — CPU ORACLE for the two post-steps of the pairwise nucmer stage.
 * TEST INFRASTRUCTURE ONLY (see pmn_oracle.h): only tests/, smoke() and bench.py's CPU legs
 * may link or call this.
 *
 *   pmo_delta_filter   `delta-filter -1` / `-m`  (/root/reference/lib/nucmer/mugsy_nucmer.ml:102-105;
 *                      -1 by default, -m with -colinear, :103)
 *   pmo_delta2maf      `delta2maf`               (/root/reference/lib/nucmer/mugsy_nucmer.ml:118-124,
 *                      lib/base/mugsy_profiles_task.ml:60)
 *
 * PARITY UNPINNED: both programs are external (MUMmer 3.20 / Mugsy), not vendored under
 * /root/reference, and the reference holds no fixture for them.  What IS pinned is the grammar
 * on either side: the .delta grammar (lib/profiles_lib/m_delta.cc:72-196) and the MAF lines the
 * reference's readers accept (lib/maf/reader.ml:12-66, lib/profiles/m_untranslate.ml:127-151,
 * lib/profiles/m_profile_stream.ml:15-20).  Rules: ORACLE_SPEC.md §8, §9.
 */
#include <ctype.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pmn_oracle.h"

typedef struct { char *v; size_t n, cap; } sbuf;
static void sb_put(sbuf *b, const char *s, size_t n)
{
    if (b->n + n + 1 > b->cap) { b->cap = (b->cap + n + 64) * 2; b->v = (char *)realloc(b->v, b->cap); }
    memcpy(b->v + b->n, s, n); b->n += n; b->v[b->n] = 0;
}
static void sb_puts(sbuf *b, const char *s) { sb_put(b, s, strlen(s)); }
static void sb_ll(sbuf *b, long long v) { char t[32]; int n = snprintf(t, sizeof t, "%lld", v); sb_put(b, t, (size_t)n); }

/* ---- parsed .delta ---- */
typedef struct { char *rid, *qid; long long rlen, qlen; } dblock;
typedef struct { int block; long long sR, eR, sQ, eQ, e1, e2, e3; size_t doff, dcnt; } dalign;
typedef struct {
    char *line1, *line2;
    dblock *blk; size_t nblk;
    dalign *al; size_t nal;
    long long *dl; size_t ndl;
} ddelta;

static void dd_free(ddelta *d)
{
    free(d->line1); free(d->line2);
    for (size_t i = 0; i < d->nblk; i++) { free(d->blk[i].rid); free(d->blk[i].qid); }
    free(d->blk); free(d->al); free(d->dl);
}

static char *dupn(const char *s, size_t n) { char *r = (char *)malloc(n + 1); memcpy(r, s, n); r[n] = 0; return r; }

/* grammar of lib/profiles_lib/m_delta.cc:72-196: two header lines, '>' lines with four tokens, seven ints, deltas up to "0" */
static int dd_parse(const char *t, size_t n, ddelta *d)
{
    memset(d, 0, sizeof *d);
    size_t p = 0; int lineno = 0; size_t capb = 0, capa = 0, capd = 0;
    int in_deltas = 0;
    while (p < n) {
        size_t e = p; while (e < n && t[e] != '\n') e++;
        const char *l = t + p; size_t len = e - p;
        if (lineno == 0) d->line1 = dupn(l, len);
        else if (lineno == 1) d->line2 = dupn(l, len);
        else if (len == 0) { /* tolerate */ }
        else if (l[0] == '>') {
            if (in_deltas) return -1;
            char rid[1024], qid[1024]; long long rl, ql;
            char *tmp = dupn(l + 1, len - 1);
            int k = sscanf(tmp, "%1023s %1023s %lld %lld", rid, qid, &rl, &ql);
            free(tmp);
            if (k != 4) return -1;
            if (d->nblk == capb) { capb = capb ? capb * 2 : 16; d->blk = (dblock *)realloc(d->blk, capb * sizeof(dblock)); }
            dblock *b = &d->blk[d->nblk++]; b->rid = dupn(rid, strlen(rid)); b->qid = dupn(qid, strlen(qid)); b->rlen = rl; b->qlen = ql;
        } else if (!in_deltas) {
            if (!d->nblk) return -1;
            dalign a; memset(&a, 0, sizeof a);
            char *tmp = dupn(l, len);
            int k = sscanf(tmp, "%lld %lld %lld %lld %lld %lld %lld", &a.sR, &a.eR, &a.sQ, &a.eQ, &a.e1, &a.e2, &a.e3);
            free(tmp);
            if (k != 7) return -1;
            a.block = (int)d->nblk - 1; a.doff = d->ndl; a.dcnt = 0;
            if (d->nal == capa) { capa = capa ? capa * 2 : 64; d->al = (dalign *)realloc(d->al, capa * sizeof(dalign)); }
            d->al[d->nal++] = a; in_deltas = 1;
        } else {
            char *tmp = dupn(l, len); char *endp = NULL; long long v = strtoll(tmp, &endp, 10);
            int bad = endp == tmp; free(tmp);
            if (bad) return -1;
            if (v == 0) in_deltas = 0;
            else {
                if (d->ndl == capd) { capd = capd ? capd * 2 : 1024; d->dl = (long long *)realloc(d->dl, capd * sizeof(long long)); }
                d->dl[d->ndl++] = v; d->al[d->nal - 1].dcnt++;
            }
        }
        lineno++; p = e + 1;
    }
    if (lineno < 2 || in_deltas) return -1;
    return 0;
}

/* ------------------------------------------------------------------ delta-filter (ORACLE_SPEC.md §8) */

typedef struct { long long lo, hi; float idy; size_t k; } lis_item;
static int lis_cmp(const void *a, const void *b)
{
    const lis_item *x = (const lis_item *)a, *y = (const lis_item *)b;
    if (x->lo != y->lo) return x->lo < y->lo ? -1 : 1;
    return x->k < y->k ? -1 : (x->k > y->k ? 1 : 0);            /* stable */
}

/* weighted longest increasing subset of one sequence's alignments; flags[k] |= bit for the members */
static void lis_flag(lis_item *it, size_t n, double maxolap, unsigned char *flags, unsigned char bit)
{
    if (!n) return;
    qsort(it, n, sizeof *it, lis_cmp);
    long long *score = (long long *)malloc(n * sizeof(long long)); long *from = (long *)malloc(n * sizeof(long));
    for (size_t i = 0; i < n; i++) {
        const long long leni = it[i].hi - it[i].lo + 1;
        const double w = (double)it[i].idy * (double)it[i].idy;
        score[i] = (long long)((double)leni * w); from[i] = -1;
        for (size_t j = 0; j < i; j++) {
            const long long lenj = it[j].hi - it[j].lo + 1;
            long long olap = it[j].hi - it[i].lo + 1; if (olap < 0) olap = 0;
            if (olap > 0 && ((double)((float)olap / (float)leni) * 100.0 > maxolap || (double)((float)olap / (float)lenj) * 100.0 > maxolap)) continue;
            const long long cand = score[j] + (long long)((double)(leni - olap) * w);
            if (cand > score[i]) { score[i] = cand; from[i] = (long)j; }
        }
    }
    size_t best = 0;
    for (size_t i = 1; i < n; i++) if (score[i] > score[best]) best = i;
    for (long k = (long)best; k >= 0; k = from[k]) flags[it[k].k] |= bit;
    free(score); free(from);
}

static size_t id_index(char ***ids, size_t *n, const char *s)
{
    for (size_t i = 0; i < *n; i++) if (!strcmp((*ids)[i], s)) return i;
    *ids = (char **)realloc(*ids, (*n + 1) * sizeof(char *)); (*ids)[*n] = (char *)s; return (*n)++;
}

int pmo_delta_filter(const char *delta, size_t n, int mode, double maxolap, char **out, size_t *nout)
{
    ddelta d;
    if (!out || !nout || (mode != 1 && mode != 2) || dd_parse(delta, n, &d)) return -1;
    unsigned char *flags = (unsigned char *)calloc(d.nal + 1, 1);
    /* sequence ids -> dense indexes, in order of first appearance */
    char **rids = NULL, **qids = NULL; size_t nr = 0, nq = 0;
    size_t *ar = (size_t *)malloc((d.nal + 1) * sizeof(size_t)), *aq = (size_t *)malloc((d.nal + 1) * sizeof(size_t));
    for (size_t k = 0; k < d.nal; k++) { ar[k] = id_index(&rids, &nr, d.blk[d.al[k].block].rid); aq[k] = id_index(&qids, &nq, d.blk[d.al[k].block].qid); }
    lis_item *it = (lis_item *)malloc((d.nal + 1) * sizeof(lis_item));
    for (int side = 0; side < 2; side++) {
        const size_t ng = side ? nq : nr;
        for (size_t g = 0; g < ng; g++) {
            size_t m = 0;
            for (size_t k = 0; k < d.nal; k++) {
                if ((side ? aq[k] : ar[k]) != g) continue;
                const dalign *a = &d.al[k];
                long long neg = 0; for (size_t t = 0; t < a->dcnt; t++) if (d.dl[a->doff + t] < 0) neg++;
                const long long cols = (a->eR - a->sR + 1) + neg;
                lis_item x;
                if (side) { x.lo = a->sQ < a->eQ ? a->sQ : a->eQ; x.hi = a->sQ < a->eQ ? a->eQ : a->sQ; } else { x.lo = a->sR; x.hi = a->eR; }
                x.idy = (float)(cols - a->e1) / (float)cols; x.k = k;
                it[m++] = x;
            }
            lis_flag(it, m, maxolap, flags, side ? 2 : 1);
        }
    }
    sbuf o = { 0, 0, 0 };
    sb_puts(&o, d.line1); sb_puts(&o, "\n"); sb_puts(&o, d.line2); sb_puts(&o, "\n");
    int last_block = -1;
    for (size_t k = 0; k < d.nal; k++) {
        const int keep = mode == 1 ? flags[k] == 3 : flags[k] != 0;
        if (!keep) continue;
        const dalign *a = &d.al[k];
        if (a->block != last_block) {
            const dblock *b = &d.blk[a->block];
            sb_puts(&o, ">"); sb_puts(&o, b->rid); sb_puts(&o, " "); sb_puts(&o, b->qid); sb_puts(&o, " "); sb_ll(&o, b->rlen); sb_puts(&o, " "); sb_ll(&o, b->qlen); sb_puts(&o, "\n");
            last_block = a->block;
        }
        const long long v[7] = { a->sR, a->eR, a->sQ, a->eQ, a->e1, a->e2, a->e3 };
        for (int c = 0; c < 7; c++) { sb_ll(&o, v[c]); sb_puts(&o, c < 6 ? " " : "\n"); }
        for (size_t t = 0; t < a->dcnt; t++) { sb_ll(&o, d.dl[a->doff + t]); sb_puts(&o, "\n"); }
        sb_puts(&o, "0\n");
    }
    if (!o.v) { o.v = (char *)malloc(1); o.v[0] = 0; }
    *out = o.v; *nout = o.n;
    free(flags); free(rids); free(qids); free(ar); free(aq); free(it); dd_free(&d);
    return 0;
}

/* ------------------------------------------------------------------ delta2maf (ORACLE_SPEC.md §9) */

typedef struct { char *id; char *seq; long long len; } frec;

/* records of a FASTA text: id = first token of the header, residues = every non-white-space byte of the sequence lines */
static frec *fa_parse(const char *t, size_t n, size_t *nrec)
{
    frec *r = NULL; size_t cnt = 0, cap = 0; size_t p = 0;
    while (p < n) {
        size_t e = p; while (e < n && t[e] != '\n') e++;
        if (e > p && t[p] == '>') {
            if (cnt == cap) { cap = cap ? cap * 2 : 8; r = (frec *)realloc(r, cap * sizeof(frec)); }
            size_t a = p + 1; while (a < e && isspace((unsigned char)t[a])) a++;
            size_t b = a; while (b < e && !isspace((unsigned char)t[b])) b++;
            r[cnt].id = dupn(t + a, b - a); r[cnt].seq = (char *)malloc(n - e + 2); r[cnt].len = 0; cnt++;
        } else if (cnt) {
            for (size_t i = p; i < e; i++) if (!isspace((unsigned char)t[i])) r[cnt - 1].seq[r[cnt - 1].len++] = t[i];
        }
        p = e + 1;
    }
    *nrec = cnt;
    return r;
}

static char comp(char c)
{
    static const char *from = "ACGTUMRWSYKVHDBNacgtumrwsykvhdbn", *to = "TGCAAKYWSRMBDHVNtgcaakywsrmbdhvn";
    const char *p = strchr(from, c);
    return (p && c) ? to[p - from] : c;
}

int pmo_delta2maf(const char *delta, size_t n, const char *ref_fasta, size_t nrb, const char *qry_fasta, size_t nqb, char **out, size_t *nout)
{
    ddelta d;
    if (!out || !nout || dd_parse(delta, n, &d)) return -1;
    size_t nr = 0, nq = 0;
    frec *R = fa_parse(ref_fasta, nrb, &nr), *Q = fa_parse(qry_fasta, nqb, &nq);
    sbuf o = { 0, 0, 0 };
    sb_puts(&o, "##maf version=1\n");
    int rc = 0;
    for (size_t k = 0; k < d.nal && !rc; k++) {
        const dalign *a = &d.al[k]; const dblock *b = &d.blk[a->block];
        const frec *fr = NULL, *fq = NULL;
        for (size_t i = 0; i < nr; i++) if (!strcmp(R[i].id, b->rid)) { fr = &R[i]; break; }
        for (size_t i = 0; i < nq; i++) if (!strcmp(Q[i].id, b->qid)) { fq = &Q[i]; break; }
        if (!fr || !fq || a->sR < 1 || a->eR > fr->len || a->sR > a->eR) { rc = -1; break; }
        const int rev = a->sQ > a->eQ;
        const long long loQ = rev ? a->eQ : a->sQ, hiQ = rev ? a->sQ : a->eQ;
        if (loQ < 1 || hiQ > fq->len) { rc = -1; break; }
        const long long lenR = a->eR - a->sR + 1, lenQ = hiQ - loQ + 1;
        long long neg = 0; for (size_t t = 0; t < a->dcnt; t++) if (d.dl[a->doff + t] < 0) neg++;
        const long long cols = lenR + neg;
        char *tr = (char *)malloc((size_t)cols + 1), *tq = (char *)malloc((size_t)cols + 1);
        long long c = 0, ia = a->sR - 1, ib = 0;            /* ia: 0-based reference index; ib: bases of the query strand consumed */
        #define QBASE(off) (rev ? comp(fq->seq[a->sQ - 1 - (off)]) : fq->seq[a->sQ - 1 + (off)])
        for (size_t t = 0; t <= a->dcnt && !rc; t++) {
            long long run;
            if (t < a->dcnt) { const long long v = d.dl[a->doff + t]; run = (v < 0 ? -v : v) - 1; } else run = a->eR - ia;
            if (run < 0 || c + run > cols || ib + run > lenQ) { rc = -1; break; }
            for (long long u = 0; u < run; u++) { tr[c] = fr->seq[ia++]; tq[c] = QBASE(ib); ib++; c++; }
            if (t < a->dcnt) {
                if (c >= cols) { rc = -1; break; }
                if (d.dl[a->doff + t] > 0) { tr[c] = fr->seq[ia++]; tq[c] = '-'; }      /* reference base over a gap in the query */
                else { if (ib >= lenQ) { rc = -1; break; } tr[c] = '-'; tq[c] = QBASE(ib); ib++; }
                c++;
            }
        }
        #undef QBASE
        if (!rc && (c != cols || ib != lenQ)) rc = -1;
        if (!rc) {
            tr[cols] = 0; tq[cols] = 0;
            sb_puts(&o, "a score=0\n");        /* as the reference's own MAF writer, lib/profiles/m_maf.ml:46 */
            sb_puts(&o, "s "); sb_puts(&o, b->rid); sb_puts(&o, " "); sb_ll(&o, a->sR - 1); sb_puts(&o, " "); sb_ll(&o, lenR); sb_puts(&o, " + "); sb_ll(&o, b->rlen); sb_puts(&o, " "); sb_put(&o, tr, (size_t)cols); sb_puts(&o, "\n");
            sb_puts(&o, "s "); sb_puts(&o, b->qid); sb_puts(&o, " "); sb_ll(&o, rev ? b->qlen - a->sQ : a->sQ - 1); sb_puts(&o, " "); sb_ll(&o, lenQ); sb_puts(&o, rev ? " - " : " + "); sb_ll(&o, b->qlen); sb_puts(&o, " "); sb_put(&o, tq, (size_t)cols); sb_puts(&o, "\n\n");
        }
        free(tr); free(tq);
    }
    for (size_t i = 0; i < nr; i++) { free(R[i].id); free(R[i].seq); }
    for (size_t i = 0; i < nq; i++) { free(Q[i].id); free(Q[i].seq); }
    free(R); free(Q); dd_free(&d);
    if (rc) { free(o.v); return rc; }
    *out = o.v; *nout = o.n;
    return 0;
}

void pmo_free(void *p) { free(p); }

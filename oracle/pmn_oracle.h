/*
 * pmn_oracle.h — CPU ORACLE for the pairwise-nucmer hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may link or call this.  The product (paramugsy_b200/, libpmnucmer.so) never does.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in MUMmer 3.20
 * (/root/reference/scripts/pm_qsub_template.sh:4), which is NOT vendored in the
 * reference and absent from this image; the reference holds no golden .delta / MUM
 * fixtures.  This file restates the published MUMmer 3.x pipeline
 * (mummer -mumreference -b -l 20 -n | mgaps -l 65 -s 90 -d 5 -f .12 | postnuc -b 200,
 * the command line behind /root/reference/lib/nucmer/mugsy_nucmer.ml:100) and is pinned
 * only at the .delta FORMAT level against the reference's own parser
 * (/root/reference/lib/profiles_lib/m_delta.cc:43-49,148-220).  See ORACLE_SPEC.md.
 */
#ifndef PMN_ORACLE_H
#define PMN_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmo_opts {
    int32_t minmatch;      /* -l, 20 */
    int32_t mincluster;    /* -c, 65 */
    int32_t maxgap;        /* -g, 90 */
    int32_t diagdiff;      /* -D, 5  */
    double  diagfactor;    /* -d, 0.12 */
    int32_t breaklen;      /* -b, 200 */
    int32_t do_forward;    /* 0 with -r */
    int32_t do_reverse;    /* 0 with -f */
    int32_t do_extend;     /* --[no]extend   */
    int32_t do_optimize;   /* --[no]optimize (only 1 is supported) */
    int32_t do_simplify;   /* --[no]simplify */
    int32_t fast_chain;    /* 0: literal O(m^2) chain DP of mgaps; 1: pruned scan with identical result */
} pmo_opts;

void pmo_default_opts(pmo_opts *o);

typedef struct pmo_run pmo_run;

/* Parse both FASTA texts (borrowed for the call) and set up an empty run. NULL on error. */
pmo_run *pmo_run_create(const char *ref_fasta, size_t ref_bytes,
                        const char *qry_fasta, size_t qry_bytes, const pmo_opts *opts);
void pmo_run_free(pmo_run *r);
const char *pmo_last_error(void);

/* Stages; each requires the previous one.  Return 0 on success. */
int pmo_stage_index(pmo_run *r);     /* suffix array + LCP of the concatenated reference */
int pmo_stage_seed(pmo_run *r);      /* MUM-reference anchors, both strands */
int pmo_stage_cluster(pmo_run *r);   /* mgaps */
int pmo_stage_extend(pmo_run *r);    /* postnuc extendClusters + parseDelta */
int pmo_stage_delta(pmo_run *r, const char *ref_path, const char *qry_path);  /* .delta text */
int pmo_run_all(pmo_run *r, const char *ref_path, const char *qry_path);

/* ---- stage dumps (pointers stay valid until pmo_run_free) ---- */
int64_t pmo_ref_len(const pmo_run *r);               /* concatenated length incl. separators */
const uint8_t *pmo_ref_codes(const pmo_run *r);      /* 0..3 acgt, 4 = matches nothing */
const int32_t *pmo_sa(const pmo_run *r);
const int32_t *pmo_lcp(const pmo_run *r);

/* anchors: 4 int32 columns (ref pos 1-based concat, query pos 1-based in strand coords, len, tag)
 * tag = query_record*2 + (reverse ? 1 : 0); sorted by (tag, query pos, ref pos). */
int64_t pmo_n_anchors(const pmo_run *r);
const int32_t *pmo_anchors(const pmo_run *r);

/* clusters in mgaps output order: matches are 3 int32 (ref pos concat, query pos, len);
 * cluster k owns matches [off[k], off[k+1]); tag as above. */
int64_t pmo_n_clusters(const pmo_run *r);
int64_t pmo_n_cluster_matches(const pmo_run *r);
const int32_t *pmo_cluster_matches(const pmo_run *r);
const int32_t *pmo_cluster_off(const pmo_run *r);
const int32_t *pmo_cluster_tag(const pmo_run *r);

/* alignments in output order: 10 int64 columns
 * (ref_rec, qry_rec, dir(0 fwd,1 rev), sA, eA, sB, eB [strand coords], errors, sim_errors, non_alphas);
 * deltas of alignment k are delta[doff[k] .. doff[k+1]) (without the terminating 0). */
int64_t pmo_n_alignments(const pmo_run *r);
const int64_t *pmo_alignments(const pmo_run *r);
const int64_t *pmo_delta_off(const pmo_run *r);
const int64_t *pmo_deltas(const pmo_run *r);

const char *pmo_delta_text(const pmo_run *r, size_t *len);

/* ---- the two post-steps of lib/nucmer/mugsy_nucmer.ml (pmn_post_oracle.c; ORACLE_SPEC.md §8, §9) ----
 * text in, malloc'd text out (release with pmo_free); 0 on success */
int pmo_delta_filter(const char *delta, size_t n, int mode /* 1: -1, 2: -m */, double maxolap /* 75.0 */, char **out, size_t *nout);
int pmo_delta2maf(const char *delta, size_t n, const char *ref_fasta, size_t nr, const char *qry_fasta, size_t nq, char **out, size_t *nout);
void pmo_free(void *p);

/* work counters for the benchmark (cells = DP cells evaluated by the extension engine) */
int64_t pmo_dp_cells(const pmo_run *r);

#ifdef __cplusplus
}
#endif
#endif

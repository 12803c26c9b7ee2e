/*
 * pmnucmer.h — C ABI of libpmnucmer.so, the B200-native pairwise nucmer stage.
 *
 * What it replaces in the reference (orbitz/paramugsy):
 *   the child process started by  lib/nucmer/mugsy_nucmer.ml:100
 *       Shell.sh "nucmer %s %s -p %s %s" ref_file query_file obname options.nucmer_opts
 *   which must leave  <obname>.delta  behind (lib/nucmer/mugsy_nucmer.ml:97-98).
 * The reference has no FFI of its own (`grep -rn external lib` is empty);
 * these entry points are what an OCaml `external` stub in lib/nucmer would bind
 * (INTEGRATION.md shows the stub).  Plain C types only; no exception crosses the boundary.
 *
 * Error convention: 0 on success, negative on failure with a message in pmn_last_error().
 * The reference's convention is "non-zero child exit => Shell.sh raises"
 * (lib/nucmer/mugsy_nucmer.ml:100, lib/base/local_interface.ml:28-35 retries).
 *
 * There is NO CPU fallback: every entry point that computes fails with PMN_E_NOGPU when
 * no CUDA device is usable.
 */
#ifndef PMNUCMER_H
#define PMNUCMER_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PMN_OK          0
#define PMN_E_ARG      (-1)   /* bad argument / malformed FASTA / unsupported option */
#define PMN_E_CUDA     (-2)   /* a CUDA call failed */
#define PMN_E_NOMEM    (-3)
#define PMN_E_IO       (-4)
#define PMN_E_NOGPU    (-5)
#define PMN_E_INTERNAL (-6)

typedef struct pmn_ctx pmn_ctx;       /* one per (process, GPU); owns a stream and all scratch memory */
typedef struct pmn_seq pmn_seq;       /* a multi-FASTA file packed on the device (both strands)        */
typedef struct pmn_index pmn_index;   /* suffix array + LCP + k-mer bucket table of a pmn_seq          */
typedef struct pmn_result pmn_result; /* the .delta text of one pair plus counters                     */

/* nucmer's command-line tunables (defaults = MUMmer 3.20 nucmer, which is what the
 * reference always runs: lib/base/nucmer_task.ml:53 passes no -nucmer_opts). */
typedef struct pmn_opts {
    int32_t minmatch;      /* -l  20   */
    int32_t mincluster;    /* -c  65   */
    int32_t maxgap;        /* -g  90   */
    int32_t diagdiff;      /* -D  5    */
    double  diagfactor;    /* -d  0.12 */
    int32_t breaklen;      /* -b  200  */
    int32_t do_forward;    /* 0 with -r */
    int32_t do_reverse;    /* 0 with -f */
    int32_t do_extend;     /* --[no]extend   */
    int32_t do_optimize;   /* --[no]optimize; only 1 is implemented */
    int32_t do_simplify;   /* --[no]simplify */
    int32_t keep_stages;   /* 1: keep anchors / clusters / alignments in the result (tests) */
    int32_t post;          /* 0: .delta only.  1 / 2: the pair also goes through the two post-steps of lib/nucmer/mugsy_nucmer.ml
                            * in the same call, without their text round trip: `delta-filter -1` (1) or `-m` (2, -colinear, :103)
                            * and `delta2maf` of the filtered delta (:118-131); see pmn_result_filtered / pmn_result_maf */
} pmn_opts;

typedef struct pmn_stats {
    int64_t ref_bases, qry_bases;
    int64_t anchors, clusters, cluster_matches, alignments, aligned_ref_bases;
    int64_t dp_cells;            /* cells evaluated by the extension engine            */
    int64_t dp_jobs;             /* alignment-engine invocations                       */
    int32_t sa_rounds;           /* prefix-doubling rounds after the 16-mer pass       */
    int32_t kmer_bits;           /* 2*K of the bucket table                            */
    float   ms_index, ms_seed, ms_cluster, ms_extend, ms_total;   /* CUDA-event times  */
    float   ms_seed_kernel;      /* k_seed alone                                       */
    float   ms_wave1, ms_stitch; /* k_ex_wave1 / k_ex_stitch alone                     */
    int64_t kernel_launches;     /* kernels launched for this pair                     */
    int64_t wave1_cells;         /* DP cells evaluated inside k_ex_wave1               */
    float   wall_ms_index, wall_ms_align, wall_ms_text;   /* host wall clock: index build, pmn_align, .delta formatting */
    float   wall_ms_post;        /* host wall clock of the two post-steps (pmn_opts.post) */
    int64_t seed_lookups;        /* query positions the seeding kernel looked up in the index; the rest were stepped over */
    int64_t arena_bytes;         /* traceback arena the extension used for this pair */
    int64_t seed_probes;         /* bit probes of the presence bitmap (one per block of minmatch - P + 1 query positions) */
} pmn_stats;

void pmn_default_opts(pmn_opts *o);

/* MUMmer 3.x `nucmer` option spellings, read in one place for the three callers that meet them: the `nucmer` argv shim
 * (the child process of lib/nucmer/mugsy_nucmer.ml:100), the OCaml stub (integration/pmn_stubs.c) and the Python mirror.
 *   -l/--minmatch -c/--mincluster -g/--maxgap -D/--diagdiff -d/--diagfactor -b/--breaklen -f/--forward -r/--reverse
 *   --mumreference --[no]extend --[no]simplify --optimize --delta -p/--prefix --device -h -V   (--long=value is accepted)
 *   --mum --maxmatch --nooptimize --banded --nodelta: PMN_E_ARG ("not implemented on the B200 path"); anything else
 *   starting with '-': PMN_E_ARG ("unknown option").
 * pmn_nucmer_parse_argv: argv WITHOUT the program name; the char pointers of *out borrow argv.  device is -1 when not given.
 * pmn_opts_parse: the free-form -nucmer_opts string the reference appends to the command line verbatim (mugsy_nucmer.ml:100;
 *   lib/base/nucmer_task.ml:53 never sets it), split like a shell would; *o = defaults with the named options applied. */
typedef struct pmn_nucmer_args {
    pmn_opts opts;
    const char *prefix;          /* -p, default "out": the output is <prefix>.delta */
    const char *ref, *qry;       /* the two positional arguments, NULL when absent */
    int32_t device, help, version;
} pmn_nucmer_args;
int  pmn_nucmer_parse_argv(int argc, const char *const *argv, pmn_nucmer_args *out);
int  pmn_opts_parse(const char *nucmer_opts, pmn_opts *o);

/* ---- context ---- */
int  pmn_ctx_create(int device, pmn_ctx **out);
void pmn_ctx_destroy(pmn_ctx *c);
const char *pmn_last_error(const pmn_ctx *c);      /* c may be NULL: last error of the calling thread */
int  pmn_device_count(void);
/* device allocations (cudaMalloc) made by the library in this process so far: all scratch is grow-only, so the
 * number stops changing once the working set has been seen; bench.py reports the count inside its timed region */
int64_t pmn_alloc_count(void);
/* MAF texts (pmn_opts.post, pmn_delta2maf) are written by the device straight into page-locked host buffers that are
 * recycled when the caller frees the result.  out[0] = idle bytes the pool holds for reuse (never more than out[2]),
 * out[1] = all page-locked bytes it accounts for (idle + in the hands of callers), out[2] = the idle budget
 * (1 GB, PMN_PINNED_POOL_MB overrides): what comes back beyond the budget is released with cudaFreeHost. */
void pmn_pinned_pool_stats(int64_t out[3]);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on, so that a caller
 * can bracket calls with its own CUDA events */
void *pmn_ctx_stream(const pmn_ctx *c);
/* running totals since pmn_ctx_create: out[0] kernels launched, out[1] host->device bytes,
 * out[2] device->host bytes, out[3] pairs aligned */
void pmn_ctx_counters(const pmn_ctx *c, int64_t out[4]);
/* host waits for the device (cudaStreamSynchronize between stages, where the host sizes the next stage from a device count)
 * made for this context so far: 4 per single-record pair (anchors, clusters, alignments, rows), plus the index builds' */
int64_t pmn_ctx_sync_count(const pmn_ctx *c);
/* INT32 ALU issue rate of this GPU in 10^9 ops/s (dependent-free IADD3/VIMNMX chains on every SM):
 * the roofline denominator of the extension DP, which MEASURED_PEAKS.json does not hold */
int  pmn_measure_int32_peak(pmn_ctx *c, double *gops_per_s, double *sm_mhz_effective);

/* ---- sequences: FASTA bytes in HOST memory -> packed text in HBM ---- */
int  pmn_seq_from_fasta(pmn_ctx *c, const char *fasta, size_t bytes, pmn_seq **out);
int  pmn_seq_from_file(pmn_ctx *c, const char *path, pmn_seq **out);
void pmn_seq_free(pmn_seq *s);
int64_t pmn_seq_bases(const pmn_seq *s);           /* concatenated, incl. one separator between records */
int  pmn_seq_records(const pmn_seq *s);

/* ---- index of a reference (kept by the caller while it aligns queries against it) ---- */
int  pmn_index_build(pmn_ctx *c, const pmn_seq *ref, pmn_index **out);
void pmn_index_free(pmn_index *ix);
/* Replication across GPUs (north_star: "the reference index is replicated with an NCCL broadcast over
 * NVLink").  An index is ONE contiguous image in HBM (header, SA, LCP, k-mer table): the owner hands
 * {pointer, size} to the collective, a receiver allocates an empty index for its own packed copy of the
 * same reference, receives into the image in place and calls pmn_index_adopt to validate the header. */
int  pmn_index_image(const pmn_index *ix, void **dev_ptr, size_t *bytes);
size_t pmn_index_image_bytes(int64_t n_bases);      /* image size for a reference of n_bases (incl. separators) */
int  pmn_index_alloc(pmn_ctx *c, const pmn_seq *ref, pmn_index **out);
int  pmn_index_adopt(pmn_index *ix);

/* ---- one pair: seeding, clustering, extension, .delta text (host memory) ----
 * ref_path / qry_path are only echoed on line 1 of the .delta
 * (lib/profiles/m_delta.ml:56 splits it on the last space). */
int  pmn_align(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o,
               const char *ref_path, const char *qry_path, pmn_result **out);
/* One large pair on G GPUs: index and query are replicated, part k of G seeds the
 * query positions of its range and leaves its anchors (int32 x4: ref pos, query pos, length, tag) in the
 * context's scratch; the parts' lists concatenated in part order (an all-gather) are exactly the anchor
 * list of the undivided run, from which pmn_align_anchors continues (clustering, extension, .delta). */
int  pmn_seed_part(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o, int part, int nparts,
                   void **dev_anchors, int64_t *n_anchors);
int  pmn_align_anchors(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o, const void *dev_anchors, int64_t n_anchors,
                       const char *ref_path, const char *qry_path, pmn_result **out);
const char *pmn_result_delta(const pmn_result *r, size_t *len);
const char *pmn_result_filtered(const pmn_result *r, size_t *len);   /* pmn_opts.post != 0: the filtered .delta */
const char *pmn_result_maf(const pmn_result *r, size_t *len);        /* pmn_opts.post != 0: its MAF */
void pmn_result_stats(const pmn_result *r, pmn_stats *out);
void pmn_result_free(pmn_result *r);

/* ---- file level: exactly the job of the `nucmer` child process ----
 * Writes out_delta_path atomically (tmp + rename) so that a retry
 * (lib/base/local_interface.ml:28-35) never sees a partial file. */
int  pmn_align_pair(pmn_ctx *c, const char *ref_fasta_path, const char *qry_fasta_path,
                    const pmn_opts *o, const char *out_delta_path);

/* The batch unit of the reference is Nucmer_task.t.searches (lib/base/nucmer_task.ml:6):
 * a list of (ref, query) paths.  Pairs that share a reference reuse its index. */
int  pmn_align_batch(pmn_ctx *c, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                     const char *const *out_delta_paths, const pmn_opts *o);

/* The job of n `mugsy_nucmer` worker processes (lib/base/nucmer_task.ml:48-59 emits one command per pair,
 * lib/nucmer/mugsy_nucmer.ml:127-131 runs nucmer, delta-filter, delta2maf): with o->post = 1 (`-1`) or 2 (`-m`, -colinear)
 * delta_outs[i] receives the FILTERED delta (what the reference copies to delta_out) and maf_outs[i] its MAF; atomic writes. */
int  pmn_worker_batch(pmn_ctx *c, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                      const char *const *delta_outs, const char *const *maf_outs, const pmn_opts *o);

/* ---- the two post-steps of lib/nucmer/mugsy_nucmer.ml, on .delta text in host memory ----
 * pmn_delta_filter: `delta-filter -1` (mode 1) / `-m` (mode 2), mugsy_nucmer.ml:102-105; maxolap is delta-filter's -o (75.0).
 * pmn_delta2maf:    `delta2maf`, mugsy_nucmer.ml:118-124; ref / qry are the packed genomes the delta was computed from.
 * *out is allocated by the library (release with pmn_free_text), *nout its length. */
int  pmn_delta_filter(pmn_ctx *c, const char *delta, size_t n, int mode, double maxolap, char **out, size_t *nout);
int  pmn_delta2maf(pmn_ctx *c, const char *delta, size_t n, const pmn_seq *ref, const pmn_seq *qry, char **out, size_t *nout);
void pmn_free_text(char *p);

/* ---- in-process batch scheduler ----
 * Replaces the reference's fan-out of one `mugsy_nucmer` process per pair: run_nucmers
 * (lib/base/job_processor.ml:128-154) cuts the pair list into Nucmer_task.t.searches
 * (lib/base/nucmer_task.ml:6) and runs `-cores N` scripts at a time
 * (lib/base/queued_task_server.ml:57-64).  Here `workers` threads share one GPU, one
 * context (stream + scratch) each; every genome is packed once and every reference index
 * is built once and shared.  Pairs are (ref[k], qry[k]) indexes into the genome list;
 * names[g] is echoed on line 1 of the .delta files.  out[k] is owned by the caller
 * (pmn_result_free).  Results do not depend on `workers`.
 * Sizing: a worker holds its scratch, 1 MB of private traceback per resident warp of the extension kernels and a traceback
 * arena that starts at 256 MB and grows with the pairs it meets (PMN_ARENA_MB); 32 workers fit one B200 and carry all 28 pairs
 * of an 8 x 5 Mbp all-vs-all at once (2040 pairs/s; 16 workers: 1955), 8 are enough where the host has few cores per GPU.
 * The workers launch on 2 x `workers` streams: the library sets CUDA_DEVICE_MAX_CONNECTIONS=32 when
 * it is loaded (unless the host set it) so that every stream has a hardware work queue of its own;
 * a host that initialises CUDA before loading the library has to set the variable itself.
 * Worker threads spin while they wait for the device when the host has two cores per worker
 * thread of every visible GPU, and yield otherwise (PMN_DEVICE_SCHED=spin|yield|block overrides). */
typedef struct pmn_sched pmn_sched;
int  pmn_sched_create(int device, int workers, pmn_sched **out);
void pmn_sched_destroy(pmn_sched *s);
int  pmn_sched_workers(const pmn_sched *s);
pmn_ctx *pmn_sched_ctx(const pmn_sched *s, int k);          /* worker k's context (borrowed) */
void pmn_sched_counters(const pmn_sched *s, int64_t out[4]); /* pmn_ctx_counters summed over the workers */
int64_t pmn_sched_sync_count(const pmn_sched *s);            /* pmn_ctx_sync_count summed over the workers */
/* genomes as FASTA bytes in HOST memory */
int  pmn_sched_align_fasta(pmn_sched *s, int n_genomes, const char *const *fasta, const size_t *bytes, const char *const *names,
                           int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out);
/* genomes already packed in HBM (any context of the scheduler's device) */
int  pmn_sched_align_seqs(pmn_sched *s, int n_genomes, const pmn_seq *const *seqs, const char *const *names,
                          int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out);
/* the same with some reference indexes supplied by the caller (indexes[g] may be NULL): built earlier, or
 * received from the GPU that built them (pmn_index_alloc / pmn_index_adopt); they are never freed here */
int  pmn_sched_align_indexed(pmn_sched *s, int n_genomes, const pmn_seq *const *seqs, const pmn_index *const *indexes, const char *const *names,
                             int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out);
/* genomes and results as files: one call per Nucmer_task.t.searches */
int  pmn_sched_align_files(pmn_sched *s, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                           const char *const *out_delta_paths, const pmn_opts *o);

/* ---- several GPUs of one box from ONE process ----
 * The reference's concurrency knob is `-cores N`: one paramugsy process keeps N `mugsy_nucmer` workers busy
 * (lib/base/paramugsy.ml:54-57, lib/base/queued_task_server.ml:57-64).  A pmn_multi is that over the GPUs of a box: one
 * pmn_sched per device, so that the `nucmer` shim's caller, the OCaml stub or any C host can use all of them without Python
 * or torch.  devices = NULL means 0 .. n_devices-1; the same device may be named twice (two schedulers on it).
 *   all-vs-all (pmn_multi_align_fasta / _files): the pairs are cut by reference into one contiguous run per device
 *   (pmn_multi_plan: pure host arithmetic, device_of_pair[k] in 0 .. n_devices-1); every device packs the genomes and builds the
 *   indexes its pairs name; nothing crosses between the devices.  Results do not depend on the number of devices.
 *   one large pair (pmn_multi_align_large): both genomes and the index on every device, device k seeds query-position part k,
 *   the anchor lists are gathered on the first device (peer copies), which clusters, extends and formats — the .delta is
 *   byte-identical for any number of devices.  stats_ms (may be NULL): wall clock of [0] pack + index, [1] seeding,
 *   [2] gather, [3] clustering + extension + text. */
typedef struct pmn_multi pmn_multi;
int  pmn_multi_plan(int n_devices, int n_genomes, const size_t *bytes /* may be NULL */, int n_pairs, const int32_t *ref, const int32_t *qry, int32_t *device_of_pair);
int  pmn_multi_create(const int *devices, int n_devices, int workers_per_device, pmn_multi **out);
void pmn_multi_destroy(pmn_multi *m);
int  pmn_multi_devices(const pmn_multi *m);
pmn_sched *pmn_multi_sched(const pmn_multi *m, int k);        /* device k's scheduler (borrowed) */
int  pmn_multi_align_fasta(pmn_multi *m, int n_genomes, const char *const *fasta, const size_t *bytes, const char *const *names,
                           int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out);
/* maf_outs may be NULL; with o->post = 1 | 2 and maf_outs, out_delta_paths[i] receives the FILTERED delta and maf_outs[i] its MAF */
int  pmn_multi_align_files(pmn_multi *m, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                           const char *const *out_delta_paths, const char *const *maf_outs, const pmn_opts *o);
int  pmn_multi_align_large(pmn_multi *m, const char *ref_fasta, size_t ref_bytes, const char *qry_fasta, size_t qry_bytes, const pmn_opts *o,
                           const char *ref_path, const char *qry_path, pmn_result **out, double *stats_ms /* [4] or NULL */);

/* ---- stage dumps for the parity tests (sizes via the n_* calls; buffers are caller-owned) ---- */
int64_t pmn_index_size(const pmn_index *ix);
int  pmn_index_copy_sa(const pmn_index *ix, int32_t *sa_out, int32_t *lcp_out);
int64_t pmn_result_n_anchors(const pmn_result *r);
int  pmn_result_copy_anchors(const pmn_result *r, int32_t *out /* n x 4 */);
int64_t pmn_result_n_clusters(const pmn_result *r);
int64_t pmn_result_n_cluster_matches(const pmn_result *r);
int  pmn_result_copy_clusters(const pmn_result *r, int32_t *matches /* m x 3 */, int32_t *off /* k+1 */, int32_t *tag /* k */);
int64_t pmn_result_n_alignments(const pmn_result *r);
int64_t pmn_result_n_deltas(const pmn_result *r);
int  pmn_result_copy_alignments(const pmn_result *r, int64_t *rows /* a x 10 */, int64_t *doff /* a+1 */, int64_t *deltas);

#ifdef __cplusplus
}
#endif
#endif

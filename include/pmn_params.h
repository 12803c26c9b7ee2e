/*
 * pmn_params.h — every tunable and scoring constant of the pairwise-nucmer path,
 * single-sourced.  Included verbatim by the CPU oracle (oracle/pmn_oracle.c), by
 * the CUDA kernels (paramugsy_b200/csrc) and by the host driver.
 *
 * The reference (orbitz/paramugsy) runs `nucmer` with NO extra options
 * (lib/base/nucmer_task.ml:53 never sets -nucmer_opts; lib/nucmer/mugsy_nucmer.ml:100
 * passes the empty string), so these are the MUMmer 3.20 `nucmer` front-end defaults
 * (scripts/pm_qsub_template.sh:4 pins MUMmer3.20).  MUMmer itself is not vendored in
 * the reference; values marked ASSUMED are restated from the published design and
 * listed in ORACLE_SPEC.md.
 */
#ifndef PMN_PARAMS_H
#define PMN_PARAMS_H

/* ---- nucmer front-end defaults (mummer | mgaps | postnuc) ---- */
#define PMN_DEF_MINMATCH    20      /* mummer -l   */
#define PMN_DEF_MINCLUSTER  65      /* mgaps  -l   */
#define PMN_DEF_MAXGAP      90      /* mgaps  -s   */
#define PMN_DEF_DIAGDIFF    5       /* mgaps  -d   */
#define PMN_DEF_DIAGFACTOR  0.12    /* mgaps  -f   */
#define PMN_DEF_BREAKLEN    200     /* postnuc -b  */

/* ---- sw_align scoring, nucleotide matrix (ASSUMED, see ORACLE_SPEC.md §5) ---- */
#define PMN_GOOD_SCORE      3       /* identical a/c/g/t                       */
#define PMN_BAD_SCORE       (-7)    /* anything else, incl. any non-acgt base  */
#define PMN_OPEN_GAP_SCORE  (-7)    /* first base of a gap                     */
#define PMN_CONT_GAP_SCORE  (-4)    /* every further base of the same gap      */

/* One extension step never spans more than this many bases of either sequence. */
#define PMN_MAX_ALIGNMENT_LENGTH 10000

/* "minus infinity" of the int32 DP.  |PMN_NEG| leaves room for 2*10000 steps of -7. */
#define PMN_NEG             (-(1 << 28))

/* ---- base codes ---- */
#define PMN_CODE_A 0
#define PMN_CODE_C 1
#define PMN_CODE_G 2
#define PMN_CODE_T 3
#define PMN_CODE_X 4                /* any other character and the record separator: matches nothing */

/* ---- DP states / traceback codes ---- */
#define PMN_ST_DEL  0               /* consumes a QUERY base only  -> negative delta */
#define PMN_ST_INS  1               /* consumes a REFERENCE base only -> positive delta */
#define PMN_ST_MAT  2               /* consumes one base of each */
#define PMN_ST_NONE 3

/* ---- modus operandi bits of the alignment engine ---- */
#define PMN_DIRECTION_BIT 0x1       /* set: forward, clear: backward */
#define PMN_SEARCH_BIT    0x2       /* no traceback / delta          */
#define PMN_FORCED_BIT    0x4       /* no band trimming, no break-length stop */
#define PMN_OPTIMAL_BIT   0x8       /* finish on the best-scoring cell even if the target was reached */
#define PMN_SEQEND_BIT    0x10      /* --nooptimize: prefer running into the sequence end */

#define PMN_FORWARD_ALIGN         0x1
#define PMN_FORCED_FORWARD_ALIGN  0x5
#define PMN_BACKWARD_SEARCH       0x2

#endif

/* pmn_stubs.c — the OCaml C stubs a maintainer adds to lib/nucmer of orbitz/paramugsy so that
 * Mugsy_nucmer calls the B200 library in process instead of forking MUMmer.
 *
 *   caml_pmn_align_pair     replaces   Shell.sh "nucmer %s %s -p %s %s" ...          lib/nucmer/mugsy_nucmer.ml:100
 *   caml_pmn_delta_filter   replaces   Shell.sh "delta-filter %s %s > %s" ...        lib/nucmer/mugsy_nucmer.ml:104
 *   caml_pmn_delta2maf      replaces   Shell.sh "delta2maf %s > %s" ...              lib/nucmer/mugsy_nucmer.ml:118-124
 *
 * Everything else of mugsy_nucmer.ml (flags, file names, cp, rm -rf tmp_dir) stays as it is: see
 * mugsy_nucmer.ml.patch beside this file.  Build: add `pmn_stubs.c` to SOURCES and `CLIBS = pmnucmer` to
 * lib/nucmer/Makefile (OCamlMakefile compiles .c sources with the OCaml headers on the include path).
 *
 * Conventions kept from the reference:
 *   - failure is an OCaml exception: `Failure msg` (mugsy_nucmer.ml:78-81 raises Failure for bad arguments; a failing
 *     Shell.sh raises too), which leaves the worker with a non-zero exit and lets the back end retry it
 *     (lib/base/local_interface.ml:28-35);
 *   - lib/nucmer/Makefile:5 links with THREADS = yes: the runtime lock is released around every library call;
 *   - the free-form -nucmer_opts string is honoured (mugsy_nucmer.ml:100 appends it to the command line verbatim): it goes
 *     through pmn_opts_parse, the same option table the `nucmer` argv shim uses.
 *
 * The image this was written in has no OCaml toolchain; tests/test_integration_stub.py compiles this file against
 * a minimal stand-in for <caml/...> (tests/fake_caml) and drives the three entry points from C.
 */
#include <caml/mlvalues.h>
#include <caml/memory.h>
#include <caml/alloc.h>
#include <caml/fail.h>
#include <caml/threads.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pmnucmer.h"

static pmn_ctx *g_ctx;                       /* one context per worker process (one GPU: $PMN_DEVICE) */
static char g_msg[512];

static int the_ctx(void)
{
    if (g_ctx) return 0;
    return pmn_ctx_create(getenv("PMN_DEVICE") ? atoi(getenv("PMN_DEVICE")) : 0, &g_ctx);
}

static void keep_error(void)
{
    const char *m = pmn_last_error(g_ctx);
    snprintf(g_msg, sizeof g_msg, "%s", m && *m ? m : "libpmnucmer: unknown error");
}

static char *slurp(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    if (!f) { snprintf(g_msg, sizeof g_msg, "cannot open %s", path); return NULL; }
    size_t cap = 1 << 16, len = 0, k;
    char *buf = (char *)malloc(cap);
    while (buf && (k = fread(buf + len, 1, cap - len, f)) > 0) {
        len += k;
        if (len == cap) { cap *= 2; buf = (char *)realloc(buf, cap); }
    }
    fclose(f);
    if (!buf) { snprintf(g_msg, sizeof g_msg, "out of memory reading %s", path); return NULL; }
    *n = len;
    return buf;
}

static int spill(const char *path, const char *data, size_t n)
{
    FILE *f = fopen(path, "wb");
    if (!f) { snprintf(g_msg, sizeof g_msg, "cannot create %s", path); return -1; }
    const int bad = fwrite(data, 1, n, f) != n;
    if (fclose(f) != 0 || bad) { snprintf(g_msg, sizeof g_msg, "write error on %s", path); return -1; }
    return 0;
}

/* external pmn_align_pair : string -> string -> string -> string -> unit = "caml_pmn_align_pair"
 *   ref_file query_file nucmer_opts delta_file */
CAMLprim value caml_pmn_align_pair(value v_ref, value v_qry, value v_opts, value v_out)
{
    CAMLparam4(v_ref, v_qry, v_opts, v_out);
    /* OCaml strings may move once the runtime lock is released: private copies */
    char *ref = strdup(String_val(v_ref)), *qry = strdup(String_val(v_qry)), *opts = strdup(String_val(v_opts)), *out = strdup(String_val(v_out));
    int rc = 0;
    caml_release_runtime_system();
    pmn_opts o;
    rc = pmn_opts_parse(opts, &o);
    if (!rc) rc = the_ctx();
    if (!rc) rc = pmn_align_pair(g_ctx, ref, qry, &o, out);
    if (rc) keep_error();
    caml_acquire_runtime_system();
    free(ref); free(qry); free(opts); free(out);
    if (rc) caml_failwith(g_msg);
    CAMLreturn(Val_unit);
}

/* external pmn_delta_filter : string -> string -> string -> unit = "caml_pmn_delta_filter"
 *   chaining_opt ("-1" | "-m") delta_file delta_filt_file */
CAMLprim value caml_pmn_delta_filter(value v_mode, value v_in, value v_out)
{
    CAMLparam3(v_mode, v_in, v_out);
    char *mode = strdup(String_val(v_mode)), *in = strdup(String_val(v_in)), *out = strdup(String_val(v_out));
    int rc = 0;
    caml_release_runtime_system();
    const int m = !strcmp(mode, "-1") ? 1 : !strcmp(mode, "-m") ? 2 : 0;
    size_t n = 0, nout = 0; char *text = NULL, *res = NULL;
    if (!m) { snprintf(g_msg, sizeof g_msg, "delta-filter: option %s is not implemented on the B200 path", mode); rc = PMN_E_ARG; }
    if (!rc && !(text = slurp(in, &n))) rc = PMN_E_IO;
    if (!rc) { rc = the_ctx(); if (!rc) rc = pmn_delta_filter(g_ctx, text, n, m, 75.0, &res, &nout); if (rc) keep_error(); }
    if (!rc && spill(out, res, nout)) rc = PMN_E_IO;
    pmn_free_text(res); free(text);
    caml_acquire_runtime_system();
    free(mode); free(in); free(out);
    if (rc) caml_failwith(g_msg);
    CAMLreturn(Val_unit);
}

/* external pmn_delta2maf : string -> string -> string -> string -> unit = "caml_pmn_delta2maf"
 *   delta_out ref_seq query_seq maf_out */
CAMLprim value caml_pmn_delta2maf(value v_delta, value v_ref, value v_qry, value v_out)
{
    CAMLparam4(v_delta, v_ref, v_qry, v_out);
    char *delta = strdup(String_val(v_delta)), *ref = strdup(String_val(v_ref)), *qry = strdup(String_val(v_qry)), *out = strdup(String_val(v_out));
    int rc = 0;
    caml_release_runtime_system();
    size_t n = 0, nout = 0; char *text = NULL, *res = NULL;
    pmn_seq *r = NULL, *q = NULL;
    if (!(text = slurp(delta, &n))) rc = PMN_E_IO;
    if (!rc) {
        rc = the_ctx();
        if (!rc) rc = pmn_seq_from_file(g_ctx, ref, &r);
        if (!rc) rc = pmn_seq_from_file(g_ctx, qry, &q);
        if (!rc) rc = pmn_delta2maf(g_ctx, text, n, r, q, &res, &nout);
        if (rc) keep_error();
    }
    if (!rc && spill(out, res, nout)) rc = PMN_E_IO;
    pmn_free_text(res); pmn_seq_free(q); pmn_seq_free(r); free(text);
    caml_acquire_runtime_system();
    free(delta); free(ref); free(qry); free(out);
    if (rc) caml_failwith(g_msg);
    CAMLreturn(Val_unit);
}

/* Drives the three stubs of integration/pmn_stubs.c the way the patched mugsy_nucmer.ml would
 * (lib/nucmer/mugsy_nucmer.ml:96-131): nucmer, delta-filter, cp, delta2maf.
 *   driver <ref.fa> <qry.fa> <nucmer_opts> <tmp_dir> <delta_out> <maf_out> [-nofilter] [-colinear]
 * Exit 0; or 2 after printing "Failure: <msg>" when a stub raised; 3 if the runtime lock was not balanced. */
#include <setjmp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <caml/mlvalues.h>
#include <caml/fail.h>
#include <caml/threads.h>

static jmp_buf g_jmp;
static char g_failure[600];
static int g_released;      /* +1 per release, -1 per acquire */

void caml_failwith(const char *msg) { snprintf(g_failure, sizeof g_failure, "%s", msg); longjmp(g_jmp, 1); }
void caml_release_runtime_system(void) { g_released++; }
void caml_acquire_runtime_system(void) { g_released--; }

value caml_pmn_align_pair(value, value, value, value);
value caml_pmn_delta_filter(value, value, value);
value caml_pmn_delta2maf(value, value, value, value);

static int cp(const char *from, const char *to)
{
    FILE *a = fopen(from, "rb"), *b = fopen(to, "wb");
    if (!a || !b) return -1;
    char buf[1 << 16]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, a)) > 0) fwrite(buf, 1, k, b);
    fclose(a); return fclose(b);
}

int main(int argc, char **argv)
{
    if (argc < 7) { fprintf(stderr, "usage: driver ref qry nucmer_opts tmp_dir delta_out maf_out [-nofilter] [-colinear]\n"); return 1; }
    int filter = 1, colinear = 0;
    for (int i = 7; i < argc; i++) { if (!strcmp(argv[i], "-nofilter")) filter = 0; if (!strcmp(argv[i], "-colinear")) colinear = 1; }
    char delta_file[4096], filt_file[4096];
    snprintf(delta_file, sizeof delta_file, "%s/nucmer.delta", argv[4]);
    snprintf(filt_file, sizeof filt_file, "%s/nucmer.filt.delta", argv[4]);
    if (setjmp(g_jmp)) {
        printf("Failure: %s\n", g_failure);
        return g_released ? 3 : 2;       /* the lock must be held again when the exception is raised */
    }
    caml_pmn_align_pair((value)argv[1], (value)argv[2], (value)argv[3], (value)delta_file);
    const char *d = delta_file;
    if (filter) { caml_pmn_delta_filter((value)(colinear ? "-m" : "-1"), (value)delta_file, (value)filt_file); d = filt_file; }
    if (cp(d, argv[5])) { fprintf(stderr, "cp failed\n"); return 1; }
    caml_pmn_delta2maf((value)argv[5], (value)argv[1], (value)argv[2], (value)argv[6]);
    return g_released ? 3 : 0;
}

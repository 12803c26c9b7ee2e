// nucmer_main.cpp — `nucmer`-argv-compatible front end over libpmnucmer.so.
//
// Drop-in for the child process of /root/reference/lib/nucmer/mugsy_nucmer.ml:100
//     nucmer <ref.fa> <qry.fa> -p <prefix> <opts...>      ->  <prefix>.delta
// Put the directory holding this binary first on $PATH — the reference's own override hook
// (scripts/pm_qsub_template.sh:6-7) — and ParaMugsy runs the B200 path unchanged.
// Option spellings follow MUMmer 3.x's nucmer script.  Exit 0 on success, 1 otherwise with a
// message on stderr (Shell.sh raises on non-zero, mugsy_nucmer.ml:100).
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pmnucmer.h"

static void usage(FILE *f)
{
    fprintf(f,
            "USAGE: nucmer [options] <Reference> <Query>\n"
            "  -p|--prefix <string>   write output to <string>.delta (default out)\n"
            "  -l|--minmatch <int>    minimum length of a single match (20)\n"
            "  -c|--mincluster <int>  minimum length of a cluster of matches (65)\n"
            "  -g|--maxgap <int>      maximum gap between two adjacent matches in a cluster (90)\n"
            "  -D|--diagdiff <int>    maximum diagonal difference between two adjacent anchors (5)\n"
            "  -d|--diagfactor <f>    ... as a fraction of the gap length (0.12)\n"
            "  -b|--breaklen <int>    distance an extension may run through a poor region (200)\n"
            "  -f|--forward  -r|--reverse   use only one strand of the query\n"
            "  --mumreference (default)   --[no]extend  --[no]simplify  --[no]optimize  --[no]delta\n"
            "  --mum, --maxmatch, --banded, --nooptimize are not implemented on the B200 path\n"
            "  --device <int>         CUDA device (default 0 or $PMN_DEVICE)\n");
}

int main(int argc, char **argv)
{
    pmn_nucmer_args a;
    if (pmn_nucmer_parse_argv(argc - 1, (const char *const *)(argv + 1), &a)) {       // the option table shared with the OCaml stub and the Python mirror
        fprintf(stderr, "%s\n", pmn_last_error(nullptr));
        usage(stderr);
        return 1;
    }
    if (a.help) { usage(stdout); return 0; }
    if (a.version) { printf("nucmer (paramugsy_b200, MUMmer 3.20 compatible front end)\n"); return 0; }
    if (!a.ref || !a.qry) { usage(stderr); return 1; }
    const int device = a.device >= 0 ? a.device : getenv("PMN_DEVICE") ? atoi(getenv("PMN_DEVICE")) : 0;
    // MUMmer's nucmer writes ABSOLUTE paths on line 1 of the .delta; delta2maf re-opens them from wherever it is run
    // (lib/base/mugsy_profiles_task.ml:60 runs it from another directory)
    char rbuf[PATH_MAX], qbuf[PATH_MAX];
    const char *rp = realpath(a.ref, rbuf), *qp = realpath(a.qry, qbuf);
    if (!rp) { fprintf(stderr, "nucmer: cannot open %s: %s\n", a.ref, strerror(errno)); return 1; }
    if (!qp) { fprintf(stderr, "nucmer: cannot open %s: %s\n", a.qry, strerror(errno)); return 1; }
    pmn_ctx *ctx = nullptr;
    if (pmn_ctx_create(device, &ctx)) { fprintf(stderr, "nucmer: %s\n", pmn_last_error(nullptr)); return 1; }
    const std::string out = std::string(a.prefix) + ".delta";
    int rc = pmn_align_pair(ctx, rp, qp, &a.opts, out.c_str());
    if (rc) fprintf(stderr, "nucmer: %s\n", pmn_last_error(ctx));
    pmn_ctx_destroy(ctx);
    return rc ? 1 : 0;
}

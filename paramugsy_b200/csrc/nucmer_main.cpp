// nucmer_main.cpp — `nucmer`-argv-compatible front end over libpmnucmer.so.
//
// Drop-in for the child process of /root/reference/lib/nucmer/mugsy_nucmer.ml:100
//     nucmer <ref.fa> <qry.fa> -p <prefix> <opts...>      ->  <prefix>.delta
// Put the directory holding this binary first on $PATH — the reference's own override hook
// (scripts/pm_qsub_template.sh:6-7) — and ParaMugsy runs the B200 path unchanged.
// Option spellings follow MUMmer 3.x's nucmer script.  Exit 0 on success, 1 otherwise with a
// message on stderr (Shell.sh raises on non-zero, mugsy_nucmer.ml:100).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

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

static bool need(int i, int argc, const char *o) { if (i + 1 >= argc) { fprintf(stderr, "nucmer: option %s needs a value\n", o); return false; } return true; }

int main(int argc, char **argv)
{
    pmn_opts o; pmn_default_opts(&o);
    std::string prefix = "out";
    std::vector<const char *> pos;
    int device = getenv("PMN_DEVICE") ? atoi(getenv("PMN_DEVICE")) : 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        auto is = [&](const char *s, const char *l) { return !strcmp(a, s) || !strcmp(a, l); };
        if (is("-p", "--prefix")) { if (!need(i, argc, a)) return 1; prefix = argv[++i]; }
        else if (!strncmp(a, "--prefix=", 9)) prefix = a + 9;
        else if (is("-l", "--minmatch")) { if (!need(i, argc, a)) return 1; o.minmatch = atoi(argv[++i]); }
        else if (is("-c", "--mincluster")) { if (!need(i, argc, a)) return 1; o.mincluster = atoi(argv[++i]); }
        else if (is("-g", "--maxgap")) { if (!need(i, argc, a)) return 1; o.maxgap = atoi(argv[++i]); }
        else if (is("-D", "--diagdiff")) { if (!need(i, argc, a)) return 1; o.diagdiff = atoi(argv[++i]); }
        else if (is("-d", "--diagfactor")) { if (!need(i, argc, a)) return 1; o.diagfactor = atof(argv[++i]); }
        else if (is("-b", "--breaklen")) { if (!need(i, argc, a)) return 1; o.breaklen = atoi(argv[++i]); }
        else if (is("-f", "--forward")) o.do_reverse = 0;
        else if (is("-r", "--reverse")) o.do_forward = 0;
        else if (!strcmp(a, "--mumreference") || !strcmp(a, "--delta")) {}
        else if (!strcmp(a, "--extend")) o.do_extend = 1;
        else if (!strcmp(a, "--noextend")) o.do_extend = 0;
        else if (!strcmp(a, "--simplify")) o.do_simplify = 1;
        else if (!strcmp(a, "--nosimplify")) o.do_simplify = 0;
        else if (!strcmp(a, "--optimize")) o.do_optimize = 1;
        else if (!strcmp(a, "--device")) { if (!need(i, argc, a)) return 1; device = atoi(argv[++i]); }
        else if (is("-h", "--help")) { usage(stdout); return 0; }
        else if (is("-V", "--version")) { printf("nucmer (paramugsy_b200, MUMmer 3.20 compatible front end)\n"); return 0; }
        else if (!strcmp(a, "--nooptimize") || !strcmp(a, "--mum") || !strcmp(a, "--maxmatch") || !strcmp(a, "--banded") || !strcmp(a, "--nodelta")) {
            fprintf(stderr, "nucmer: option %s is not implemented on the B200 path\n", a); return 1;
        }
        else if (a[0] == '-' && a[1]) { fprintf(stderr, "nucmer: unknown option %s\n", a); usage(stderr); return 1; }
        else pos.push_back(a);
    }
    if (pos.size() != 2) { usage(stderr); return 1; }
    if (!o.do_forward && !o.do_reverse) { fprintf(stderr, "nucmer: -f and -r are mutually exclusive\n"); return 1; }
    pmn_ctx *ctx = nullptr;
    if (pmn_ctx_create(device, &ctx)) { fprintf(stderr, "nucmer: %s\n", pmn_last_error(nullptr)); return 1; }
    const std::string out = prefix + ".delta";
    int rc = pmn_align_pair(ctx, pos[0], pos[1], &o, out.c_str());
    if (rc) fprintf(stderr, "nucmer: %s\n", pmn_last_error(ctx));
    pmn_ctx_destroy(ctx);
    return rc ? 1 : 0;
}

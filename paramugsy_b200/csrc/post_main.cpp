// post_main.cpp — `delta-filter` and `delta2maf` argv-compatible front ends over libpmnucmer.so
// (one binary, the program name decides).
//
// Drop-ins for the two child processes that follow nucmer in
// /root/reference/lib/nucmer/mugsy_nucmer.ml:
//     delta-filter -1|-m <in.delta> > <out.delta>          (:102-105)
//     delta2maf <in.delta> > <out.maf>                      (:118-124; lib/base/mugsy_profiles_task.ml:60)
// Output goes to stdout as with the originals; delta2maf finds the two FASTA files on line 1 of
// the .delta, as the MUMmer tools do.  Exit 0 on success, 1 with a message on stderr otherwise.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "pmnucmer.h"

static int slurp(const char *path, std::string &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); return 1; }
    char buf[1 << 16]; size_t k;
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    fclose(f);
    return 0;
}

int main(int argc, char **argv)
{
    const char *prog = strrchr(argv[0], '/'); prog = prog ? prog + 1 : argv[0];
    const bool maf = strstr(prog, "delta2maf") != nullptr;
    int mode = 0; double maxolap = 75.0; const char *in = nullptr;
    int device = getenv("PMN_DEVICE") ? atoi(getenv("PMN_DEVICE")) : 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        if (!maf && !strcmp(a, "-1")) mode = 1;
        else if (!maf && !strcmp(a, "-m")) mode = 2;
        else if (!maf && !strcmp(a, "-o")) { if (i + 1 >= argc) { fprintf(stderr, "%s: -o needs a value\n", prog); return 1; } maxolap = atof(argv[++i]); }
        else if (!strcmp(a, "--device")) { if (i + 1 >= argc) return 1; device = atoi(argv[++i]); }
        else if (a[0] == '-' && a[1]) { fprintf(stderr, "%s: option %s is not implemented on the B200 path\n", prog, a); return 1; }
        else if (!in) in = a;
        else { fprintf(stderr, "%s: more than one input file\n", prog); return 1; }
    }
    if (!in || (!maf && !mode)) {
        fprintf(stderr, maf ? "USAGE: delta2maf <deltafile>\n" : "USAGE: delta-filter -1|-m [-o maxolap] <deltafile>\n");
        return 1;
    }
    std::string delta;
    if (slurp(in, delta)) return 1;
    pmn_ctx *ctx = nullptr;
    if (pmn_ctx_create(device, &ctx)) { fprintf(stderr, "%s: %s\n", prog, pmn_last_error(nullptr)); return 1; }
    char *out = nullptr; size_t n = 0; int rc;
    if (!maf) rc = pmn_delta_filter(ctx, delta.data(), delta.size(), mode, maxolap, &out, &n);
    else {
        const size_t nl = delta.find('\n'); const std::string l1 = delta.substr(0, nl == std::string::npos ? delta.size() : nl);
        const size_t sp = l1.rfind(' ');                       // lib/profiles/m_delta.ml:56 splits on the last space too
        if (sp == std::string::npos) { fprintf(stderr, "%s: line 1 of %s does not name two files\n", prog, in); pmn_ctx_destroy(ctx); return 1; }
        pmn_seq *r = nullptr, *q = nullptr;
        rc = pmn_seq_from_file(ctx, l1.substr(0, sp).c_str(), &r);
        if (!rc) rc = pmn_seq_from_file(ctx, l1.substr(sp + 1).c_str(), &q);
        if (!rc) rc = pmn_delta2maf(ctx, delta.data(), delta.size(), r, q, &out, &n);
        pmn_seq_free(q); pmn_seq_free(r);
    }
    if (rc) fprintf(stderr, "%s: %s\n", prog, pmn_last_error(ctx));
    else if (fwrite(out, 1, n, stdout) != n) { fprintf(stderr, "%s: write error\n", prog); rc = 1; }
    pmn_free_text(out);
    pmn_ctx_destroy(ctx);
    return rc ? 1 : 0;
}

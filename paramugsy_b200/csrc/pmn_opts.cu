// pmn_opts.cu — MUMmer 3.x `nucmer` option spellings -> pmn_opts, in ONE place.
//
// The reference forwards a free-form option string to the child process verbatim
// (/root/reference/lib/nucmer/mugsy_nucmer.ml:100: "nucmer %s %s -p %s %s" ... options.nucmer_opts;
// lib/base/nucmer_task.ml:53 never sets it).  Three callers need the same reading of it: the
// `nucmer` argv shim (nucmer_main.cpp), the OCaml stub (integration/pmn_stubs.c) and the Python
// mirror (paramugsy_b200/mugsy_nucmer.py).  No CUDA in here.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pmn_host.h"

static bool parse_int(const char *s, int32_t *out)
{
    if (!s || !*s) return false;
    char *end = nullptr;
    const long v = strtol(s, &end, 10);
    if (*end || v < -2147483647L || v > 2147483647L) return false;
    *out = (int32_t)v;
    return true;
}

static bool parse_double(const char *s, double *out)
{
    if (!s || !*s) return false;
    char *end = nullptr;
    const double v = strtod(s, &end);
    if (*end) return false;
    *out = v;
    return true;
}

extern "C" int pmn_nucmer_parse_argv(int argc, const char *const *argv, pmn_nucmer_args *out)
{
    if (!out || argc < 0 || (argc && !argv)) return pmn_set_error(PMN_E_ARG, "pmn_nucmer_parse_argv: NULL argument");
    pmn_default_opts(&out->opts);
    out->prefix = "out"; out->ref = out->qry = nullptr; out->device = -1; out->help = out->version = 0;
    pmn_opts &o = out->opts;
    int npos = 0;
    for (int i = 0; i < argc; i++) {
        const char *a = argv[i];
        if (!a) return pmn_set_error(PMN_E_ARG, "nucmer: NULL argument %d", i);
        auto is = [&](const char *s, const char *l) { return !strcmp(a, s) || !strcmp(a, l); };
        const char *val = nullptr;       // value of "--long=value"
        std::string name = a;
        if (a[0] == '-' && a[1] == '-') { const char *eq = strchr(a, '='); if (eq) { name.assign(a, (size_t)(eq - a)); val = eq + 1; } }
        auto nis = [&](const char *s, const char *l) { return name == s || name == l; };
        auto value = [&]() -> const char * { if (val) return val; if (i + 1 >= argc) return nullptr; return argv[++i]; };
        int32_t *ip = nullptr;
        if (nis("-l", "--minmatch")) ip = &o.minmatch;
        else if (nis("-c", "--mincluster")) ip = &o.mincluster;
        else if (nis("-g", "--maxgap")) ip = &o.maxgap;
        else if (nis("-D", "--diagdiff")) ip = &o.diagdiff;
        else if (nis("-b", "--breaklen")) ip = &o.breaklen;
        if (ip) {
            const char *v = value();
            if (!v) return pmn_set_error(PMN_E_ARG, "nucmer: option %s needs a value", name.c_str());
            if (!parse_int(v, ip)) return pmn_set_error(PMN_E_ARG, "nucmer: option %s needs an integer, got '%s'", name.c_str(), v);
        }
        else if (nis("-d", "--diagfactor")) {
            const char *v = value();
            if (!v) return pmn_set_error(PMN_E_ARG, "nucmer: option %s needs a value", name.c_str());
            if (!parse_double(v, &o.diagfactor)) return pmn_set_error(PMN_E_ARG, "nucmer: option %s needs a number, got '%s'", name.c_str(), v);
        }
        else if (nis("-p", "--prefix")) {
            const char *v = value();
            if (!v) return pmn_set_error(PMN_E_ARG, "nucmer: option %s needs a value", name.c_str());
            out->prefix = v;
        }
        else if (name == "--device") {
            const char *v = value(); int32_t d = 0;
            if (!v || !parse_int(v, &d)) return pmn_set_error(PMN_E_ARG, "nucmer: option --device needs an integer");
            out->device = d;
        }
        else if (is("-f", "--forward")) o.do_reverse = 0;
        else if (is("-r", "--reverse")) o.do_forward = 0;
        else if (is("--mumreference", "--delta")) {}
        else if (!strcmp(a, "--extend")) o.do_extend = 1;
        else if (!strcmp(a, "--noextend")) o.do_extend = 0;
        else if (!strcmp(a, "--simplify")) o.do_simplify = 1;
        else if (!strcmp(a, "--nosimplify")) o.do_simplify = 0;
        else if (!strcmp(a, "--optimize")) o.do_optimize = 1;
        else if (is("-h", "--help")) out->help = 1;
        else if (is("-V", "--version")) out->version = 1;
        else if (!strcmp(a, "--nooptimize") || !strcmp(a, "--mum") || !strcmp(a, "--maxmatch") || !strcmp(a, "--banded") || !strcmp(a, "--nodelta") ||
                 !strcmp(a, "--nobanded"))
            return pmn_set_error(PMN_E_ARG, "nucmer: option %s is not implemented on the B200 path", a);
        else if (a[0] == '-' && a[1]) return pmn_set_error(PMN_E_ARG, "nucmer: unknown option %s", a);
        else {
            if (npos == 0) out->ref = a; else if (npos == 1) out->qry = a;
            else return pmn_set_error(PMN_E_ARG, "nucmer: more than two positional arguments ('%s')", a);
            npos++;
        }
    }
    if (!o.do_forward && !o.do_reverse) return pmn_set_error(PMN_E_ARG, "nucmer: -f and -r are mutually exclusive");
    return 0;
}

extern "C" int pmn_opts_parse(const char *nucmer_opts, pmn_opts *o)
{
    if (!o) return pmn_set_error(PMN_E_ARG, "pmn_opts_parse: NULL argument");
    // the shell would have split the string on white space (Shell.sh goes through /bin/sh): the same here, honouring
    // single and double quotes
    std::vector<std::string> tok;
    if (nucmer_opts) {
        std::string cur; bool have = false; char quote = 0;
        for (const char *p = nucmer_opts; *p; p++) {
            const char ch = *p;
            if (quote) { if (ch == quote) quote = 0; else cur.push_back(ch); }
            else if (ch == '\'' || ch == '"') { quote = ch; have = true; }
            else if (ch == ' ' || ch == '\t' || ch == '\n' || ch == '\r') { if (have || !cur.empty()) { tok.push_back(cur); cur.clear(); have = false; } }
            else cur.push_back(ch);
        }
        if (quote) return pmn_set_error(PMN_E_ARG, "nucmer_opts: unbalanced quote");
        if (have || !cur.empty()) tok.push_back(cur);
    }
    std::vector<const char *> av;
    for (auto &t : tok) av.push_back(t.c_str());
    pmn_nucmer_args a;
    const int rc = pmn_nucmer_parse_argv((int)av.size(), av.data(), &a);
    if (rc) return rc;
    if (a.ref || a.help || a.version || a.device >= 0 || strcmp(a.prefix, "out"))
        return pmn_set_error(PMN_E_ARG, "nucmer_opts may hold alignment options only (no file, -p, --device, -h or -V)");
    *o = a.opts;          // defaults (pmn_default_opts) with the named options applied; keep_stages = post = 0
    return 0;
}

// pmn_post.cu — the two post-steps every pair goes through after nucmer in the reference:
//
//   delta-filter -1 | -m   /root/reference/lib/nucmer/mugsy_nucmer.ml:102-105 (filter defaults to true, :54;
//                          -m with -colinear, :103): keeps the alignments on the best weighted chain of
//                          each reference sequence and of each query sequence (-1: on both, -m: on either)
//   delta2maf              /root/reference/lib/nucmer/mugsy_nucmer.ml:118-124 and
//                          lib/base/mugsy_profiles_task.ml:60: every alignment of a .delta as a MAF block
//
// Both are external programs in the reference (MUMmer 3.20 / Mugsy, not vendored), so the rules are
// restated in ORACLE_SPEC.md §8-§9 and checked against oracle/pmn_post_oracle.c; the grammars on both
// sides are the reference's own (.delta: lib/profiles_lib/m_delta.cc:72-196; MAF lines:
// lib/maf/reader.ml:12-66, lib/profiles/m_untranslate.ml:127-151).
//
// The host parses the .delta text (it is the interface of the reference, a few hundred kB); the chain DP
// and the expansion of the edit scripts into the two gapped rows of every block run on the device.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "pmn_scratch.cuh"

namespace {

struct PBlock { std::string rid, qid; long long rlen = 0, qlen = 0; };
struct PAlign { int block; long long sR, eR, sQ, eQ, e1, e2, e3; size_t doff, dcnt; long long neg; };
struct PDelta {
    std::string line1, line2;
    std::vector<PBlock> blk; std::vector<PAlign> al; std::vector<int32_t> dl;
};

static bool next_ll(const char *&p, const char *e, long long &v)
{
    while (p < e && (*p == ' ' || *p == '\t')) p++;
    if (p >= e) return false;
    bool neg = false;
    if (*p == '-') { neg = true; p++; } else if (*p == '+') p++;
    if (p >= e || *p < '0' || *p > '9') return false;
    long long x = 0;
    while (p < e && *p >= '0' && *p <= '9') { x = x * 10 + (*p - '0'); p++; }
    v = neg ? -x : x;
    return true;
}
static bool next_tok(const char *&p, const char *e, std::string &t)
{
    while (p < e && (*p == ' ' || *p == '\t')) p++;
    const char *a = p;
    while (p < e && *p != ' ' && *p != '\t') p++;
    t.assign(a, p);
    return p > a;
}

// grammar of lib/profiles_lib/m_delta.cc:72-196
static int parse_delta(const char *t, size_t n, PDelta &d)
{
    size_t p = 0; int lineno = 0; bool in_deltas = false;
    while (p < n) {
        const char *nl = (const char *)memchr(t + p, '\n', n - p);
        const size_t e = nl ? (size_t)(nl - t) : n;
        const char *l = t + p, *le = t + e;
        if (le > l && le[-1] == '\r') le--;
        if (lineno == 0) d.line1.assign(l, le);
        else if (lineno == 1) d.line2.assign(l, le);
        else if (le == l) { /* tolerate blank lines */ }
        else if (*l == '>') {
            if (in_deltas) return pmn_set_error(PMN_E_ARG, "delta: '>' line inside an alignment (line %d)", lineno + 1);
            PBlock b; const char *q = l + 1;
            if (!next_tok(q, le, b.rid) || !next_tok(q, le, b.qid) || !next_ll(q, le, b.rlen) || !next_ll(q, le, b.qlen))
                return pmn_set_error(PMN_E_ARG, "delta: malformed '>' line %d", lineno + 1);
            d.blk.push_back(b);
        } else if (!in_deltas) {
            if (d.blk.empty()) return pmn_set_error(PMN_E_ARG, "delta: alignment before any '>' line (line %d)", lineno + 1);
            PAlign a; const char *q = l;
            if (!next_ll(q, le, a.sR) || !next_ll(q, le, a.eR) || !next_ll(q, le, a.sQ) || !next_ll(q, le, a.eQ) || !next_ll(q, le, a.e1) || !next_ll(q, le, a.e2) || !next_ll(q, le, a.e3))
                return pmn_set_error(PMN_E_ARG, "delta: malformed alignment line %d", lineno + 1);
            a.block = (int)d.blk.size() - 1; a.doff = d.dl.size(); a.dcnt = 0; a.neg = 0;
            d.al.push_back(a); in_deltas = true;
        } else {
            long long v; const char *q = l;
            if (!next_ll(q, le, v) || v > 0x7fffffffll || v < -0x7fffffffll) return pmn_set_error(PMN_E_ARG, "delta: malformed delta line %d", lineno + 1);
            if (v == 0) in_deltas = false;
            else { d.dl.push_back((int32_t)v); d.al.back().dcnt++; if (v < 0) d.al.back().neg++; }
        }
        lineno++; p = e + 1;
    }
    if (lineno < 2 || in_deltas) return pmn_set_error(PMN_E_ARG, "delta: truncated input");
    return 0;
}

static inline char *put_ll(char *p, long long v) { return pmn_fmt_int(p, v); }

}  // namespace

// ------------------------------------------------------------------------------------ delta-filter

// One warp per sequence (group): the weighted longest-increasing-subset DP of ORACLE_SPEC.md §8 over the
// group's alignments in (lo, input order) order.  Alignment i is tried against all earlier j, 32 at a time.
__global__ void __launch_bounds__(128) k_filter_lis(const long long *__restrict__ lo, const long long *__restrict__ hi, const int32_t *__restrict__ cols,
                                                   const int32_t *__restrict__ errs, const int32_t *__restrict__ item, const int32_t *__restrict__ goff,
                                                   int ngroups, double maxolap, long long *__restrict__ score, int32_t *__restrict__ from, uint8_t *__restrict__ flags)
{
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= ngroups) return;
    const int b = goff[g], e = goff[g + 1];
    long long best = LLONG_MIN; int besti = -1;
    for (int i = b; i < e; i++) {
        const long long leni = hi[i] - lo[i] + 1;
        const float idy = (float)(cols[i] - errs[i]) / (float)cols[i];
        const double w = __dmul_rn((double)idy, (double)idy);
        long long sc = (long long)__dmul_rn((double)leni, w); int fr = -1;
        long long cbest = LLONG_MIN; int cj = 0x7fffffff;
        for (int j0 = b; j0 < i; j0 += 32) {
            const int j = j0 + lane;
            long long cand = LLONG_MIN;
            if (j < i) {
                const long long lenj = hi[j] - lo[j] + 1;
                long long olap = hi[j] - lo[i] + 1; if (olap < 0) olap = 0;
                const bool skip = olap > 0 && (__dmul_rn((double)((float)olap / (float)leni), 100.0) > maxolap || __dmul_rn((double)((float)olap / (float)lenj), 100.0) > maxolap);
                if (!skip) cand = score[j] + (long long)__dmul_rn((double)(leni - olap), w);
            }
            if (cand > cbest) { cbest = cand; cj = j; }          // per lane: j ascends, strict > keeps the first
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long oc = __shfl_xor_sync(0xffffffffu, cbest, o); const int oj = __shfl_xor_sync(0xffffffffu, cj, o);
            if (oc > cbest || (oc == cbest && oj < cj)) { cbest = oc; cj = oj; }
        }
        if (cbest > sc) { sc = cbest; fr = cj; }
        if (lane == 0) { score[i] = sc; from[i] = fr; }
        if (sc > best) { best = sc; besti = i; }                   // first maximum
        __syncwarp();
    }
    if (lane == 0) for (int k = besti; k >= 0; k = from[k]) flags[item[k]] = 1;
}

// which alignments of d survive `delta-filter -1` (mode 1) / `-m` (mode 2)
static int filter_keep(pmn_ctx *c, const PDelta &d, int mode, double maxolap, std::vector<uint8_t> &keep)
{
    PMN_CUDA_OK(cudaSetDevice(c->device));
    pmn_tls_stream = c->stream;
    cudaStream_t st = c->stream;
    const size_t na = d.al.size();
    std::vector<uint8_t> flagR(na, 0), flagQ(na, 0);
    keep.assign(na, 0);
    if (na) {
        // groups: reference sequences, then query sequences, each in order of first appearance
        std::map<std::string, int> rix, qix;
        std::vector<int> gr(na), gq(na);
        for (size_t k = 0; k < na; k++) {
            const PBlock &b = d.blk[(size_t)d.al[k].block];
            gr[k] = rix.emplace(b.rid, (int)rix.size()).first->second; gq[k] = qix.emplace(b.qid, (int)qix.size()).first->second;
            if (d.al[k].eR - d.al[k].sR + 1 + d.al[k].neg > 0x7fffffffll) return pmn_set_error(PMN_E_ARG, "delta-filter: alignment too long");
        }
        const int nr = (int)rix.size(), nq = (int)qix.size();
        std::vector<long long> lo(2 * na), hi(2 * na); std::vector<int32_t> cols(2 * na), errs(2 * na), item(2 * na), goff((size_t)nr + nq + 1);
        std::vector<size_t> ord(na);
        size_t at = 0;
        for (int side = 0; side < 2; side++) {
            auto key_lo = [&](size_t k) { const PAlign &a = d.al[k]; return side ? std::min(a.sQ, a.eQ) : a.sR; };
            for (size_t k = 0; k < na; k++) ord[k] = k;
            std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) {
                const int gx = side ? gq[x] : gr[x], gy = side ? gq[y] : gr[y];
                if (gx != gy) return gx < gy;
                return key_lo(x) < key_lo(y);
            });
            int prev = -1;
            for (size_t t = 0; t < na; t++) {
                const size_t k = ord[t]; const PAlign &a = d.al[k];
                const int g = side ? gq[k] : gr[k];
                while (prev < g) { prev++; goff[(size_t)(side ? nr : 0) + prev] = (int32_t)at; }
                lo[at] = key_lo(k); hi[at] = side ? std::max(a.sQ, a.eQ) : a.eR;
                cols[at] = (int32_t)(a.eR - a.sR + 1 + a.neg); errs[at] = (int32_t)a.e1; item[at] = (int32_t)(k + (side ? na : 0));
                at++;
            }
        }
        goff[(size_t)nr + nq] = (int32_t)at;
        // device
        Scratch &S = *c->scratch;
        const size_t bytes = 2 * na * (8 + 8 + 4 + 4 + 4 + 8 + 4) + 4 * goff.size() + 2 * na + 256;
        if (S.ex_a.ensure(bytes + 1024)) return -3;
        char *base = (char *)S.ex_a.p; size_t o = 0;
        auto take = [&](size_t b) { char *p = base + o; o += (b + 15) / 16 * 16; return p; };
        long long *dlo = (long long *)take(16 * na), *dhi = (long long *)take(16 * na), *dscore = (long long *)take(16 * na);
        int32_t *dcols = (int32_t *)take(8 * na), *derrs = (int32_t *)take(8 * na), *ditem = (int32_t *)take(8 * na), *dfrom = (int32_t *)take(8 * na);
        int32_t *dgoff = (int32_t *)take(4 * goff.size()); uint8_t *dflags = (uint8_t *)take(2 * na);
        PMN_H2D(c, dlo, lo.data(), 16 * na); PMN_H2D(c, dhi, hi.data(), 16 * na);
        PMN_H2D(c, dcols, cols.data(), 8 * na); PMN_H2D(c, derrs, errs.data(), 8 * na); PMN_H2D(c, ditem, item.data(), 8 * na);
        PMN_H2D(c, dgoff, goff.data(), 4 * goff.size());
        PMN_CUDA_OK(cudaMemsetAsync(dflags, 0, 2 * na, st));
        const int ng = nr + nq;
        k_filter_lis<<<(ng + 3) / 4, 128, 0, st>>>(dlo, dhi, dcols, derrs, ditem, dgoff, ng, maxolap, dscore, dfrom, dflags);
        c->launches += 1;
        std::vector<uint8_t> fl(2 * na);
        PMN_D2H(c, fl.data(), dflags, 2 * na);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        PMN_CUDA_OK(cudaGetLastError());
        for (size_t k = 0; k < na; k++) { flagR[k] = fl[k]; flagQ[k] = fl[na + k]; }
    }
    for (size_t k = 0; k < na; k++) keep[k] = mode == 1 ? (flagR[k] && flagQ[k]) : (flagR[k] || flagQ[k]);
    return 0;
}

// .delta text of the alignments with keep[k] != 0 (all when keep is empty), in input order, '>' lines only where one follows
static void emit_delta(const PDelta &d, const std::vector<uint8_t> &keep, std::string &t)
{
    const size_t na = d.al.size();
    t.clear();
    t.reserve(d.line1.size() + d.line2.size() + 64 + na * 120 + d.dl.size() * 8);
    t += d.line1; t += '\n'; t += d.line2; t += '\n';
    int last_block = -1; char buf[256];
    for (size_t k = 0; k < na; k++) {
        if (!keep.empty() && !keep[k]) continue;
        const PAlign &a = d.al[k];
        if (a.block != last_block) {
            const PBlock &b = d.blk[(size_t)a.block];
            t += '>'; t += b.rid; t += ' '; t += b.qid; t += ' ';
            char *p = put_ll(buf, b.rlen); *p++ = ' '; p = put_ll(p, b.qlen); *p++ = '\n'; t.append(buf, p);
            last_block = a.block;
        }
        const long long v[7] = { a.sR, a.eR, a.sQ, a.eQ, a.e1, a.e2, a.e3 };
        char *p = buf;
        for (int i = 0; i < 7; i++) { p = put_ll(p, v[i]); *p++ = i < 6 ? ' ' : '\n'; }
        t.append(buf, p);
        const size_t before = t.size();
        t.resize(before + a.dcnt * 12);
        char *q = &t[before];
        for (size_t u = 0; u < a.dcnt; u++) { q = put_ll(q, d.dl[a.doff + u]); *q++ = '\n'; }
        t.resize((size_t)(q - &t[0]));
        t += "0\n";
    }
}

static int text_out(const std::string &t, char **out, size_t *nout)
{
    char *r = (char *)malloc(t.size() + 1);
    if (!r) return pmn_set_error(PMN_E_NOMEM, "out of memory (%zu bytes)", t.size());
    memcpy(r, t.data(), t.size()); r[t.size()] = 0;
    *out = r; *nout = t.size();
    return 0;
}

extern "C" int pmn_delta_filter(pmn_ctx *c, const char *delta, size_t n, int mode, double maxolap, char **out, size_t *nout)
{
    if (!c || !delta || !out || !nout || (mode != 1 && mode != 2)) return pmn_set_error(PMN_E_ARG, "pmn_delta_filter: bad argument");
    *out = nullptr; *nout = 0;
    PDelta d; { int rc = parse_delta(delta, n, d); if (rc) return rc; }
    std::vector<uint8_t> keep;
    { int rc = filter_keep(c, d, mode, maxolap, keep); if (rc) return rc; }
    std::string t;
    emit_delta(d, keep, t);
    return text_out(t, out, nout);
}

extern "C" void pmn_free_text(char *p) { if (p && !pmn_pinned_put(p)) free(p); }

// ------------------------------------------------------------------------------------ delta2maf

struct MafAlign {
    long long refbase;      // index of reference base sR in the reference residues (0-based, concatenation)
    long long qrybase;      // index of the FORWARD query base the strand starts on (sQ, 0-based, concatenation)
    long long outR, outQ;   // byte offsets of the two text rows in the output
    long long gcol0;        // first column in the column space of all alignments
    int32_t first, dcnt;    // its deltas
    int32_t rev, pad;
};

__device__ __forceinline__ uint8_t maf_comp(uint8_t ch)
{
    switch (ch) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A';
        case 'M': return 'K'; case 'R': return 'Y'; case 'W': return 'W'; case 'S': return 'S'; case 'Y': return 'R'; case 'K': return 'M';
        case 'V': return 'B'; case 'H': return 'D'; case 'D': return 'H'; case 'B': return 'V'; case 'N': return 'N';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a';
        case 'm': return 'k'; case 'r': return 'y'; case 'w': return 'w'; case 's': return 's'; case 'y': return 'r'; case 'k': return 'm';
        case 'v': return 'b'; case 'h': return 'd'; case 'd': return 'h'; case 'b': return 'v'; case 'n': return 'n';
        default: return ch;
    }
}

// per delta: columns, reference bases and query bases it stands for (|d|-1 aligned pairs, then the indel)
__global__ void __launch_bounds__(256) k_maf_counts(const int32_t *__restrict__ d, int64_t nd, uint32_t *__restrict__ fc, uint32_t *__restrict__ fa, uint32_t *__restrict__ fb)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t > nd) return;
    if (t == nd) { fc[t] = 0; fa[t] = 0; fb[t] = 0; return; }
    const int v = d[t]; const uint32_t m = (uint32_t)(v < 0 ? -v : v);
    fc[t] = m; fa[t] = v > 0 ? m : m - 1; fb[t] = v > 0 ? m - 1 : m;
}

#define MAF_CHUNK 32

// One thread per chunk of MAF_CHUNK columns of the column space: finds its alignment and the delta its first
// column belongs to by binary search on the prefix sums, then walks forward writing both rows.
__global__ void __launch_bounds__(256) k_maf_expand(const MafAlign *__restrict__ al, int nal, long long total_cols, const int32_t *__restrict__ d,
                                                   const uint32_t *__restrict__ sc, const uint32_t *__restrict__ sa, const uint32_t *__restrict__ sb,
                                                   const uint8_t *__restrict__ rres, const uint8_t *__restrict__ qres, uint8_t *__restrict__ out)
{
    const long long g0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * MAF_CHUNK;
    if (g0 >= total_cols) return;
    int lo = 0, hi = nal - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (al[mid].gcol0 <= g0) lo = mid; else hi = mid - 1; }
    int k = lo;
    long long g = g0; const long long gend = g0 + MAF_CHUNK < total_cols ? g0 + MAF_CHUNK : total_cols;
    while (g < gend) {
        const MafAlign A = al[k];
        const long long acols_end = k + 1 < nal ? al[k + 1].gcol0 : total_cols;
        if (g >= acols_end) { k++; continue; }
        const long long stop = gend < acols_end ? gend : acols_end;
        const uint32_t c = (uint32_t)(g - A.gcol0);               // column inside the alignment
        // last delta t in [first, first+dcnt] whose columns start at or before c (t = first+dcnt: the tail run)
        int tl = A.first, th = A.first + A.dcnt;
        const uint32_t c0 = sc[A.first];
        while (tl < th) { const int mid = (tl + th + 1) >> 1; if (sc[mid] - c0 <= c) tl = mid; else th = mid - 1; }
        int t = tl;
        uint32_t within = c - (sc[t] - c0);
        long long ia = A.refbase + (long long)(sa[t] - sa[A.first]), ib = (long long)(sb[t] - sb[A.first]);
        uint8_t *oR = out + A.outR + c, *oQ = out + A.outQ + c;
        while (g < stop) {
            const bool tail = t == A.first + A.dcnt;
            const int v = tail ? 0 : d[t];
            const uint32_t m = tail ? 0xffffffffu : (uint32_t)(v < 0 ? -v : v);      // columns of this item (tail: the rest)
            // aligned pairs
            while (g < stop && (tail || within + 1 < m)) {
                const uint8_t qb = qres[A.rev ? A.qrybase - (ib + within) : A.qrybase + (ib + within)];
                *oR++ = rres[ia + within]; *oQ++ = A.rev ? maf_comp(qb) : qb;
                within++; g++;
            }
            if (g >= stop) break;
            // the indel column
            if (v > 0) { *oR++ = rres[ia + within]; *oQ++ = '-'; }
            else { const uint8_t qb = qres[A.rev ? A.qrybase - (ib + within) : A.qrybase + (ib + within)]; *oR++ = '-'; *oQ++ = A.rev ? maf_comp(qb) : qb; }
            g++;
            ia += v > 0 ? m : m - 1; ib += v > 0 ? m - 1 : m;
            t++; within = 0;
        }
    }
}

// MAF text of the alignments of d with keep[k] != 0 (all when keep is empty); `ref` / `qry` are the packed genomes the
// delta was computed from (their residues stay in HBM).  *text comes from the pinned-host pool (pmn_pinned_put / pmn_free_text).
static int maf_of(pmn_ctx *c, const PDelta &d_all, const std::vector<uint8_t> &keep, const pmn_seq *ref, const pmn_seq *qry, char **text, size_t *text_len)
{
    static const bool timing = getenv("PMN_POST_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_0 = now();
    // the surviving alignments (their deltas stay where they are in d_all.dl)
    struct View { const std::vector<PBlock> &blk; std::vector<PAlign> al; const std::vector<int32_t> &dl; };
    View d{ d_all.blk, {}, d_all.dl };
    if (keep.empty()) d.al = d_all.al; else for (size_t k = 0; k < d_all.al.size(); k++) if (keep[k]) d.al.push_back(d_all.al[k]);
    const double t_parse = now();
    PMN_CUDA_OK(cudaSetDevice(c->device));
    pmn_tls_stream = c->stream;
    cudaStream_t st = c->stream;
    const size_t na = d.al.size();
    const char *head = "##maf version=1\n";
    std::map<std::string, int> rrec, qrec;
    for (int i = 0; i < ref->nrec; i++) rrec.emplace(ref->ids[(size_t)i], i);
    for (int i = 0; i < qry->nrec; i++) qrec.emplace(qry->ids[(size_t)i], i);
    // layout of the output: header, then per alignment "a score=0\n" "s <id> <start> <size> <strand> <srcsize> " text "\n" twice, "\n"
    std::vector<MafAlign> al(na);
    std::vector<std::string> hR(na), hQ(na);
    long long gcol = 0; size_t at = strlen(head);
    char buf[128];
    for (size_t k = 0; k < na; k++) {
        const PAlign &a = d.al[k]; const PBlock &b = d.blk[(size_t)a.block];
        auto ir = rrec.find(b.rid); auto iq = qrec.find(b.qid);
        if (ir == rrec.end() || iq == qrec.end()) return pmn_set_error(PMN_E_ARG, "delta2maf: sequence %s / %s of the delta is not in the FASTA", b.rid.c_str(), b.qid.c_str());
        const long long rlen = ref->len[(size_t)ir->second], qlen = qry->len[(size_t)iq->second];
        const bool rev = a.sQ > a.eQ;
        const long long loQ = rev ? a.eQ : a.sQ, hiQ = rev ? a.sQ : a.eQ;
        if (a.sR < 1 || a.eR > rlen || a.sR > a.eR || loQ < 1 || hiQ > qlen) return pmn_set_error(PMN_E_ARG, "delta2maf: alignment %zu lies outside its sequences", k);
        long long sumA = 0, sumB = 0;
        for (size_t u = 0; u < a.dcnt; u++) { const long long v = d.dl[a.doff + u], m = v < 0 ? -v : v; sumA += v > 0 ? m : m - 1; sumB += v > 0 ? m - 1 : m; }
        const long long lenR = a.eR - a.sR + 1, lenQ = hiQ - loQ + 1, cols = lenR + a.neg;
        if (sumA > lenR || lenR - sumA != lenQ - sumB) return pmn_set_error(PMN_E_ARG, "delta2maf: the deltas of alignment %zu do not fit its coordinates", k);
        MafAlign &m = al[k];
        m.refbase = ref->off[(size_t)ir->second] + a.sR - 1; m.qrybase = qry->off[(size_t)iq->second] + a.sQ - 1;
        m.gcol0 = gcol; m.first = (int32_t)a.doff; m.dcnt = (int32_t)a.dcnt; m.rev = rev ? 1 : 0; m.pad = 0;
        std::string &r = hR[k], &q = hQ[k];
        r = "a score=0\ns "; r += b.rid; r += ' ';
        char *p = put_ll(buf, a.sR - 1); *p++ = ' '; p = put_ll(p, lenR); memcpy(p, " + ", 3); p += 3; p = put_ll(p, b.rlen); *p++ = ' '; r.append(buf, p);
        q = "\ns "; q += b.qid; q += ' ';
        p = put_ll(buf, rev ? b.qlen - a.sQ : a.sQ - 1); *p++ = ' '; p = put_ll(p, lenQ); memcpy(p, rev ? " - " : " + ", 3); p += 3; p = put_ll(p, b.qlen); *p++ = ' '; q.append(buf, p);
        at += r.size(); m.outR = (long long)at; at += (size_t)cols;
        at += q.size(); m.outQ = (long long)at; at += (size_t)cols;
        at += 2;                                    // "\n\n"
        gcol += cols;
    }
    const size_t total = at;
    char *res = pmn_pinned_get(total);
    if (!res) return PMN_E_NOMEM;
    struct Guard { char *p; ~Guard() { if (p) pmn_pinned_put(p); } } guard{ res };
    const double t_layout = now(); double t_dev = t_layout, t_copy = t_layout;
    if (na && gcol > 0) {
        Scratch &S = *c->scratch;
        const int64_t nd = (int64_t)d.dl.size();
        if (S.ex_d.ensure(4 * (size_t)(nd + 1) + 64) || S.ex_b.ensure(3 * 4 * (size_t)(nd + 1) + 64) || S.ex_c.ensure(3 * 4 * (size_t)(nd + 1) + 64) ||
            S.ex_g.ensure(sizeof(MafAlign) * na + 64) || S.ex_pool.ensure(total + 64) || S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(nd + 1))) return -3;
        int32_t *dd = S.ex_d.as<int32_t>();
        uint32_t *fc = S.ex_b.as<uint32_t>(), *fa = fc + (nd + 1), *fb = fa + (nd + 1);
        uint32_t *sc = S.ex_c.as<uint32_t>(), *sa = sc + (nd + 1), *sb = sa + (nd + 1);
        MafAlign *dal = S.ex_g.as<MafAlign>(); uint8_t *dout = S.ex_pool.as<uint8_t>();
        if (nd) PMN_H2D(c, dd, d.dl.data(), 4 * (size_t)nd);
        PMN_H2D(c, dal, al.data(), sizeof(MafAlign) * na);
        k_maf_counts<<<(unsigned)((nd + 1 + 255) / 256), 256, 0, st>>>(dd, nd, fc, fa, fb);
        pmn_scan<uint32_t, OpAddU32, false>(fc, sc, nd + 1, S.scan_tmp.as<uint32_t>(), st);
        pmn_scan<uint32_t, OpAddU32, false>(fa, sa, nd + 1, S.scan_tmp.as<uint32_t>(), st);
        pmn_scan<uint32_t, OpAddU32, false>(fb, sb, nd + 1, S.scan_tmp.as<uint32_t>(), st);
        const long long chunks = (gcol + MAF_CHUNK - 1) / MAF_CHUNK;
        k_maf_expand<<<(unsigned)((chunks + 255) / 256), 256, 0, st>>>(dal, (int)na, gcol, dd, sc, sa, sb, ref->residues.as<uint8_t>(), qry->residues.as<uint8_t>(), dout);
        c->launches += 11;
        PMN_D2H(c, res, dout, total);          // straight into the pinned result buffer
        cudaError_t e = cudaStreamSynchronize(st);
        t_dev = now();
        t_copy = now();
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return pmn_set_error(-2, "pmn_delta2maf: %s", cudaGetErrorString(e));
    }
    // the host fills in what is not sequence text
    memcpy(res, head, strlen(head));
    for (size_t k = 0; k < na; k++) {
        const long long cols = (k + 1 < na ? al[k + 1].gcol0 : gcol) - al[k].gcol0;
        memcpy(res + al[k].outR - (long long)hR[k].size(), hR[k].data(), hR[k].size());
        memcpy(res + al[k].outQ - (long long)hQ[k].size(), hQ[k].data(), hQ[k].size());
        res[al[k].outQ + cols] = '\n'; res[al[k].outQ + cols + 1] = '\n';
    }
    res[total] = 0;
    guard.p = nullptr; *text = res; *text_len = total;
    if (timing) fprintf(stderr, "[pmn] delta2maf: select %.2f ms, layout %.2f ms, device (H2D, kernels, D2H) %.2f ms, copy %.2f ms, headers %.2f ms; %zu bytes\n",
                        t_parse - t_0, t_layout - t_parse, t_dev - t_layout, t_copy - t_dev, now() - t_copy, total);
    return 0;
}

extern "C" int pmn_delta2maf(pmn_ctx *c, const char *delta, size_t n, const pmn_seq *ref, const pmn_seq *qry, char **out, size_t *nout)
{
    if (!c || !delta || !ref || !qry || !out || !nout) return pmn_set_error(PMN_E_ARG, "pmn_delta2maf: bad argument");
    *out = nullptr; *nout = 0;
    PDelta d; { int rc = parse_delta(delta, n, d); if (rc) return rc; }
    return maf_of(c, d, std::vector<uint8_t>(), ref, qry, out, nout);
}

// ------------------------------------------------------------------------------------ both post-steps inside pmn_align

// pmn_opts.post: the alignments of a finished pair go through delta-filter and delta2maf without the text round trip of
// the reference's three child processes (mugsy_nucmer.ml:100,104,122): the rows and deltas the extension left on the
// host become the parsed form directly.
int pmn_post_impl(pmn_ctx *c, const pmn_seq *ref, const pmn_seq *qry, const char *ref_path, const char *qry_path, int mode, pmn_result *r)
{
    PDelta d;
    d.line1 = std::string(ref_path) + " " + qry_path; d.line2 = "NUCMER";
    const size_t na = r->al_rows.size() / 10;
    d.dl.assign(r->al_deltas.begin(), r->al_deltas.end());
    int64_t prev_r = -1, prev_q = -1;
    for (size_t k = 0; k < na; k++) {
        const int64_t *a = &r->al_rows[k * 10];
        if (a[0] != prev_r || a[1] != prev_q) {
            PBlock b; b.rid = ref->ids[(size_t)a[0]]; b.qid = qry->ids[(size_t)a[1]]; b.rlen = ref->len[(size_t)a[0]]; b.qlen = qry->len[(size_t)a[1]];
            d.blk.push_back(b); prev_r = a[0]; prev_q = a[1];
        }
        PAlign p; p.block = (int)d.blk.size() - 1;
        const int64_t lenB = qry->len[(size_t)a[1]];
        p.sR = a[3]; p.eR = a[4]; p.sQ = a[2] ? lenB - a[5] + 1 : a[5]; p.eQ = a[2] ? lenB - a[6] + 1 : a[6];
        p.e1 = a[7]; p.e2 = a[8]; p.e3 = a[9];
        p.doff = (size_t)r->al_doff[k]; p.dcnt = (size_t)(r->al_doff[k + 1] - r->al_doff[k]); p.neg = 0;
        for (size_t u = 0; u < p.dcnt; u++) if (d.dl[p.doff + u] < 0) p.neg++;
        d.al.push_back(p);
    }
    std::vector<uint8_t> keep;
    int rc = filter_keep(c, d, mode, 75.0, keep);
    if (rc) return rc;
    emit_delta(d, keep, r->filtered);
    return maf_of(c, d, keep, ref, qry, &r->maf, &r->maf_len);
}

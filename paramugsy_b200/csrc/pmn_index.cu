// pmn_index.cu — reference index on the device:
//   pack (codes -> 2-bit text + x-mask, both strands), suffix array by radix-sort prefix
//   doubling, LCP (from the sorted 16-mer keys where they decide it, a Kasai-style walk over the rest), and
//   the K-mer bucket table the seeding kernel starts from.
//
// Stands in for MUMmer's suffix-tree construction inside `mummer`, first stage of the
// `nucmer` child process of /root/reference/lib/nucmer/mugsy_nucmer.ml:100.
// Oracle counterpart: oracle/pmn_oracle.c §3 (suffix_cmp, pmo_stage_index) — SA and LCP
// must be bit-identical to it.
//
// Suffix order (same as the oracle): END < a < c < g < t < X_p, every X (non-acgt base or
// record separator) being its own symbol ordered by text position.
#include <vector>

#include "pmn_scratch.cuh"

struct OpMaxU32 { __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; } static __device__ __forceinline__ uint32_t identity() { return 0u; } };

// ------------------------------------------------------------------------------------ pack

// one thread per 32-base word
__global__ void __launch_bounds__(256) k_pack(const uint8_t *__restrict__ codes, int64_t n, uint64_t *__restrict__ w, uint32_t *__restrict__ xm, int64_t nwords)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords) return;
    uint64_t word = 0; uint32_t mask = 0;
    int64_t base = k * 32;
    if (base + 32 <= n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(codes + base);   // base is a multiple of 32: aligned
        uint4 q0 = __ldg(src), q1 = __ldg(src + 1);
        uint32_t v[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
#pragma unroll
        for (int j = 0; j < 32; j++) {
            uint32_t c = (v[j >> 2] >> (8 * (j & 3))) & 0xff;
            if (c < 4) word |= (uint64_t)c << (62 - 2 * j); else mask |= 1u << (31 - j);
        }
    } else {
        for (int j = 0; j < 32; j++) {
            uint32_t c = base + j < n ? codes[base + j] : 4u;
            if (c < 4) word |= (uint64_t)c << (62 - 2 * j); else mask |= 1u << (31 - j);
        }
    }
    w[k] = word; xm[k] = mask;
}

// reverse complement of the whole concatenation, one thread per output word
__global__ void __launch_bounds__(256) k_revcomp(const uint64_t *__restrict__ fw, const uint32_t *__restrict__ fx, int64_t n,
                                                uint64_t *__restrict__ rw, uint32_t *__restrict__ rx, int64_t nwords)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords) return;
    if (k * 32 >= n) { rw[k] = 0; rx[k] = ~0u; return; }
    int64_t f_lo = n - 32 - k * 32;
    uint64_t win; uint32_t xw;
    if (f_lo >= 0) { win = pmn_window64(fw, f_lo); xw = pmn_xwindow32(fx, f_lo); }
    else { int sh = (int)(-f_lo); win = pmn_window64(fw, 0) >> (2 * sh); xw = (pmn_xwindow32(fx, 0) >> sh) | (~0u << (32 - sh)); }
    uint64_t r = __brevll(win);
    r = ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
    uint32_t xr = __brev(xw);
    // keep X positions zero in the text
    uint64_t spread = 0;
#pragma unroll
    for (int j = 0; j < 32; j++) if ((xr >> (31 - j)) & 1u) spread |= 3ull << (62 - 2 * j);
    rw[k] = ~r & ~spread; rx[k] = xr;
}

// ------------------------------------------------------------------------------------ FASTA on the device

// v[i] = 2i+1 where a header starts ('>' first on its line), 2i at a newline, 0 elsewhere; the
// inclusive max-scan tells every byte whether the latest such event was a header start, i.e.
// whether it lies inside a header line
__global__ void __launch_bounds__(256) k_fa_events(const uint8_t *__restrict__ txt, int64_t nb, uint32_t *__restrict__ ev)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const uint8_t ch = txt[i];
    uint32_t v = 0;
    if (ch == '\n') v = (uint32_t)(2 * i + 2);
    else if (ch == '>' && (i == 0 || txt[i - 1] == '\n')) v = (uint32_t)(2 * i + 3);
    ev[i] = v;
}

__device__ __forceinline__ uint32_t fa_code(uint8_t ch)      // 0..3 acgt, 4 other base, 255 white space
{
    switch (ch) {
        case 'a': case 'A': return PMN_CODE_A;
        case 'c': case 'C': return PMN_CODE_C;
        case 'g': case 'G': return PMN_CODE_G;
        case 't': case 'T': return PMN_CODE_T;
        case ' ': case '\t': case '\r': case '\n': case '\v': case '\f': return 255u;
        default: return PMN_CODE_X;
    }
}

// keep[i] = byte i yields a code: a base outside header lines, or the '>' of every header but the
// first (it becomes the one-base separator between records)
__global__ void __launch_bounds__(256) k_fa_keep(const uint8_t *__restrict__ txt, const uint32_t *__restrict__ evmax, int64_t nb, int64_t first_header,
                                                uint32_t *__restrict__ keep)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nb) return;
    if (i == nb) { keep[i] = 0; return; }
    const uint32_t m = evmax[i];
    uint32_t k;
    if (m & 1u) k = ((int64_t)((m - 3) >> 1) == i && i != first_header) ? 1u : 0u;     // inside a header: only a later '>' itself
    else k = (i > first_header && fa_code(txt[i]) != 255u) ? 1u : 0u;
    keep[i] = k;
}

__global__ void __launch_bounds__(256) k_fa_emit(const uint8_t *__restrict__ txt, const uint32_t *__restrict__ keep, const uint32_t *__restrict__ pos, int64_t nb,
                                                uint8_t *__restrict__ codes, uint8_t *__restrict__ residues, uint32_t *__restrict__ any_x)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nb || !keep[i]) return;
    const uint8_t ch = txt[i];
    residues[pos[i]] = ch;                // the characters as given (case, IUPAC codes): what delta2maf prints
    uint32_t cd = ch == '>' ? (uint32_t)PMN_CODE_X : fa_code(ch);
    if (ch == '>') cd = PMN_CODE_X;       // a '>' is only kept as a record separator ('>' inside a sequence line is a plain non-acgt base)
    codes[pos[i]] = (uint8_t)cd;
    if (cd == PMN_CODE_X) *any_x = 1u;
}

__global__ void k_fa_gather(const uint32_t *__restrict__ pos, const int64_t *__restrict__ hp, int nrec, int64_t nb, uint32_t *__restrict__ out)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrec) out[r] = pos[hp[r]];
    if (r == nrec) out[r] = pos[nb];
}

// FASTA bytes in host memory -> base codes in HBM -> packed text of both strands.  The host only
// locates the header lines (for the record ids); everything per base happens on the device.
int pmn_fasta_to_device(pmn_ctx *c, pmn_seq *s, const char *txt, size_t nb, const std::vector<int64_t> &header_pos)
{
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    const int nrec = (int)header_pos.size();
    if (nb >= 0x7ffffff0ull) return pmn_set_error(PMN_E_ARG, "FASTA: file larger than 2 GiB");
    if (S.k0.ensure(nb + 64) || S.v0.ensure(4 * (nb + 1)) || S.v1.ensure(4 * (nb + 1)) || S.k1.ensure(4 * (nb + 1)) ||
        S.codes.ensure(nb + 128) || S.scan_tmp.ensure(8 * pmn_scan_scratch_elems((int64_t)nb + 1)) || S.flags.ensure(16 * (size_t)(nrec + 8)) ||
        S.ensure_pinned(8 * (size_t)(nrec + 2) + 64)) return -3;
    uint8_t *dtxt = S.k0.as<uint8_t>(); uint32_t *ev = S.v0.as<uint32_t>(), *keep = S.v1.as<uint32_t>(), *pos = S.k1.as<uint32_t>();
    int64_t *dhp = S.flags.as<int64_t>(); uint32_t *dout = (uint32_t *)(dhp + nrec + 1); uint32_t *anyx = dout + nrec + 2;
    PMN_H2D(c, dtxt, txt, nb);
    PMN_H2D(c, dhp, header_pos.data(), 8 * (size_t)nrec);
    PMN_CUDA_OK(cudaMemsetAsync(anyx, 0, 4, st));
    PMN_CUDA_OK(cudaMemsetAsync(S.codes.p, PMN_CODE_X, nb + 128, st));
    const unsigned g = (unsigned)((nb + 1 + 255) / 256);
    k_fa_events<<<g, 256, 0, st>>>(dtxt, (int64_t)nb, ev);
    pmn_scan<uint32_t, OpMaxU32, true>(ev, ev, (int64_t)nb, S.scan_tmp.as<uint32_t>(), st);
    k_fa_keep<<<g, 256, 0, st>>>(dtxt, ev, (int64_t)nb, header_pos[0], keep);
    pmn_scan<uint32_t, OpAddU32, false>(keep, pos, (int64_t)nb + 1, S.scan_tmp.as<uint32_t>(), st);
    if (pmn_pool_get(c, s->residues, nb + 128)) return -3;
    k_fa_emit<<<g, 256, 0, st>>>(dtxt, keep, pos, (int64_t)nb, S.codes.as<uint8_t>(), s->residues.as<uint8_t>(), anyx);
    k_fa_gather<<<(nrec + 1 + 255) / 256, 256, 0, st>>>(pos, dhp, nrec, (int64_t)nb, dout);
    c->launches += 10;
    uint32_t *h = (uint32_t *)S.pinned;
    PMN_D2H(c, h, dout, 4 * (size_t)(nrec + 3));
    PMN_CUDA_OK(cudaStreamSynchronize(st));
    const int64_t n = h[nrec];
    s->n = n;
    s->off.resize((size_t)nrec); s->len.resize((size_t)nrec);
    for (int r = 0; r < nrec; r++) s->off[(size_t)r] = (int64_t)h[r] + (r > 0 ? 1 : 0);
    for (int r = 0; r < nrec; r++) s->len[(size_t)r] = (r + 1 < nrec ? (int64_t)h[r + 1] : n) - s->off[(size_t)r];
    s->has_x = (h[nrec + 2] || nrec > 1) ? 1 : 0;
    if (n > 0x7ffffff0ll) return pmn_set_error(PMN_E_ARG, "FASTA: more than 2^31 bases");
    s->nwords = ((n + 31) / 32 + 3) / 4 * 4 + PMN_PAD_WORDS;
    if (pmn_pool_get(c, s->w_fwd, 8 * (size_t)s->nwords) || pmn_pool_get(c, s->xm_fwd, 4 * (size_t)s->nwords) ||
        pmn_pool_get(c, s->w_rev, 8 * (size_t)s->nwords) || pmn_pool_get(c, s->xm_rev, 4 * (size_t)s->nwords)) return -3;
    const unsigned gw = (unsigned)((s->nwords + 255) / 256);
    k_pack<<<gw, 256, 0, st>>>(S.codes.as<uint8_t>(), n, s->w_fwd.as<uint64_t>(), s->xm_fwd.as<uint32_t>(), s->nwords);
    k_revcomp<<<gw, 256, 0, st>>>(s->w_fwd.as<uint64_t>(), s->xm_fwd.as<uint32_t>(), n, s->w_rev.as<uint64_t>(), s->xm_rev.as<uint32_t>(), s->nwords);
    c->launches += 2;
    PMN_CUDA_OK(cudaGetLastError());
    PMN_CUDA_OK(cudaStreamSynchronize(st));   // txt may be reused by the caller
    return 0;
}

// ------------------------------------------------------------------------------------ suffix array

#define CLS_REGULAR 15

// Initial key of suffix i = its first 16 symbols as (32-bit padded 16-mer, 6-bit class):
//   END inside the window after t>=1 bases: pad with a (0), class t-1      (shorter first)
//   16 matchable bases:                     class 15
//   X after t>=0 bases:                     pad with t (3), class 16+(16-t) (longer first,
//                                           ties = distinct X's, resolved by text position
//                                           because the sort is stable)
// class_first: (text without X) the 15 suffixes that end inside their window are written in front, shortest first, and the
// rest behind them in text order: the array is then ordered by the 6 class bits among equal 16-mers, and the stable sort can
// skip its pass over them
template <bool CLASS_FIRST>
__global__ void __launch_bounds__(256) k_sa_keys(PackedView s, uint64_t *__restrict__ keys, uint32_t *__restrict__ keys32, uint32_t *__restrict__ vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    int v = pmn_valid32(s, i);
    uint32_t hi = (uint32_t)(pmn_window64(s.w, i) >> 32);
    uint32_t cls;
    if (v >= 16) cls = CLS_REGULAR;
    else {
        uint32_t keep = v ? ~0u << (32 - 2 * v) : 0u;
        if (i + v >= s.n) { hi &= keep; cls = (uint32_t)(v - 1); }
        else { hi = (hi & keep) | ~keep; cls = 16u + (16u - (uint32_t)v); }
    }
    if (CLASS_FIRST) {      // the class bits stay out of the key: 32-bit keys, a third less traffic per sort pass
        const int64_t at = i >= s.n - 15 ? s.n - 1 - i : i + 15;
        keys32[at] = hi; vals[at] = (uint32_t)i;
    } else { keys[i] = (uint64_t)hi << 6 | cls; vals[i] = (uint32_t)i; }
}

__global__ void __launch_bounds__(256) k_sa_heads(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ headpos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint64_t k = keys[p];
    bool head = p == 0 || k != keys[p - 1] || (k & 63) != CLS_REGULAR;
    headpos[p] = head ? (int32_t)p : -1;
}

// class-first keys (text without X): the suffixes whose class is not CLS_REGULAR are the last 15 of the text
__global__ void __launch_bounds__(256) k_sa_heads32(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, int32_t *__restrict__ headpos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t k = keys[p];
    bool head = p == 0 || k != keys[p - 1] || (int64_t)vals[p] >= n - 15 || (int64_t)vals[p - 1] >= n - 15;    // a group never contains a suffix of another class
    headpos[p] = head ? (int32_t)p : -1;
}

// rank[suffix] = 1 + first slot of its group; flags[p] = slot p still shares its group
__global__ void __launch_bounds__(256) k_sa_init_rank(const uint32_t *__restrict__ sa, const int32_t *__restrict__ gs, int64_t n,
                                                     int32_t *__restrict__ rank, uint32_t *__restrict__ flags)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int32_t g = gs[p];
    rank[sa[p]] = g + 1;
    flags[p] = (g != (int32_t)p || (p + 1 < n && gs[p + 1] == g)) ? 1u : 0u;
}

// dst[pos[i]] = src ? src[i] : i   for flagged i
__global__ void __launch_bounds__(256) k_compact(const uint32_t *__restrict__ flags, const uint32_t *__restrict__ pos, int64_t n,
                                                const uint32_t *__restrict__ src, uint32_t *__restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    dst[pos[i]] = src ? src[i] : (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_round_keys(const uint32_t *__restrict__ slots, int64_t m, const uint32_t *__restrict__ sa,
                                                   const int32_t *__restrict__ gs, const int32_t *__restrict__ rank, int64_t h, int64_t n,
                                                   int bits_r, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t p = slots[k], s = sa[p];
    uint64_t r2 = (int64_t)s + h < n ? (uint64_t)rank[s + h] : 0ull;
    keys[k] = ((uint64_t)(gs[p] + 1) << bits_r) | r2;
    vals[k] = s;
}

__global__ void __launch_bounds__(256) k_round_heads(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ slots, int64_t m, int32_t *__restrict__ headpos)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    headpos[k] = (k == 0 || keys[k] != keys[k - 1]) ? (int32_t)slots[k] : -1;
}

__global__ void __launch_bounds__(256) k_round_apply(const uint32_t *__restrict__ slots, const uint32_t *__restrict__ vals, const int32_t *__restrict__ gsn,
                                                    int64_t m, uint32_t *__restrict__ sa, int32_t *__restrict__ gs, int32_t *__restrict__ rank,
                                                    uint32_t *__restrict__ flags)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    uint32_t p = slots[k]; int32_t g = gsn[k];
    sa[p] = vals[k]; gs[p] = g; rank[vals[k]] = g + 1;
    flags[k] = (g != (int32_t)p || (k + 1 < m && gsn[k + 1] == g)) ? 1u : 0u;
}

// ------------------------------------------------------------------------------------ LCP

// The 16-mer keys in sorted order already hold most of the LCP array: two neighbours with different keys, both with 16
// matchable bases, share clz(xor)/2 bases — no text access.  What is left are the members of groups of equal 16-mers (their
// order is only final after the doubling rounds) and the suffixes that meet an X or the end inside their first 16 bases,
// and their successors.  Those suffixes are flagged by text position; k_lcp_text runs Kasai's walk over the flagged
// positions only.  HI(k) = the padded 16-mer of key k, REG(p) = slot p holds a suffix with 16 matchable bases.
template <class KeyT>
__global__ void __launch_bounds__(256) k_lcp_keys(const KeyT *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n,
                                                 int32_t *__restrict__ lcp, uint8_t *__restrict__ unres)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    auto hi = [&](int64_t q) -> uint32_t { return sizeof(KeyT) == 4 ? (uint32_t)keys[q] : (uint32_t)((uint64_t)keys[q] >> 6); };
    auto reg = [&](int64_t q) -> bool { return sizeof(KeyT) == 4 ? (int64_t)vals[q] < n - 15 : ((uint64_t)keys[q] & 63) == CLS_REGULAR; };
    const uint32_t h = hi(p);
    bool open_ = !reg(p);
    uint32_t hp = 0;
    if (p > 0) { hp = hi(p - 1); open_ = open_ || !reg(p - 1) || hp == h; }
    if (p + 1 < n && reg(p + 1) && hi(p + 1) == h) open_ = true;
    if (open_) unres[vals[p]] = 1;
    else lcp[p] = p > 0 ? (__clz((int)(hp ^ h)) >> 1) : 0;
}

#define PMN_LCP_CHUNK 16
__global__ void __launch_bounds__(256) k_lcp_text(PackedView s, const uint32_t *__restrict__ sa, const int32_t *__restrict__ rank,
                                                 const uint8_t *__restrict__ unres, int32_t *__restrict__ lcp)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i0 = t * PMN_LCP_CHUNK;
    if (i0 >= s.n) return;
    const uint4 f = *reinterpret_cast<const uint4 *>(unres + i0);        // the array is padded to a multiple of 16
    if (!(f.x | f.y | f.z | f.w)) return;
    const uint32_t fw[4] = { f.x, f.y, f.z, f.w };
    const int64_t i1 = i0 + PMN_LCP_CHUNK < s.n ? i0 + PMN_LCP_CHUNK : s.n;
    int64_t l = 0, last = i0;
    for (int64_t i = i0; i < i1; i++) {
        const int k = (int)(i - i0);
        if (!((fw[k >> 2] >> (8 * (k & 3))) & 0xffu)) continue;
        const int32_t p = rank[i] - 1;
        l -= i - last; if (l < 0) l = 0;          // Kasai: lcp at position i is at least the one at `last` minus the distance
        last = i;
        if (p == 0) { lcp[0] = 0; l = 0; continue; }
        const int64_t j = sa[p - 1];
        l = pmn_lcp(s, i, s, j, l, s.n);
        lcp[p] = (int32_t)l;
    }
}

// ------------------------------------------------------------------------------------ skip table of the seeding kernel

// rep(p) = longest prefix of the suffix at text position p that another suffix shares = max(LCP[slot], LCP[slot + 1]), and
// e(p) = p + rep(p), the text coordinate where that repeat ends.  e is non-decreasing (a repeat at p of length l is a repeat
// at p + 1 of length l - 1).  For a coordinate E the first p with e(p) >= E is what the seeding kernel asks for: a match of
// the query that ends at E and starts before that p is longer than any repeat of its reference suffix, so it is the unique
// longest match of its query position.  Thread p writes skip[E] = E - p for E in (e(p-1), e(p)] (E >= p there, because
// e(p-1) >= p - 1); e being monotone the ranges tile [0, e(n-1)], and the last thread closes the table up to E = n, which
// only makes the kernel look up one position more than it would have to.
__global__ void __launch_bounds__(256) k_skip_e(const int32_t *__restrict__ rank, const int32_t *__restrict__ lcp, int64_t n, int32_t *__restrict__ e)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t slot = rank[p] - 1;
    const int32_t a = lcp[slot], b = slot + 1 < n ? lcp[slot + 1] : 0;
    e[p] = (int32_t)p + (a > b ? a : b);
}
__global__ void __launch_bounds__(256) k_skip_fill(const int32_t *__restrict__ e, int64_t n, uint8_t *__restrict__ skip)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x, nn = (uint32_t)n;          // n < 2^31 and e <= n
    if (p >= nn) return;
    const uint32_t lo = p > 0 ? (uint32_t)e[p - 1] + 1u : 0u;
    const uint32_t hi = p + 1 < nn ? (uint32_t)e[p] : nn;      // the last position closes the table: E <= n
    for (uint32_t E = lo; E <= hi; E++) { const uint32_t d = E - p; skip[E] = (uint8_t)(d < 255u ? d : 255u); }
}

// ------------------------------------------------------------------------------------ K-mer bucket table

// bucket key of SA slot p: first K symbols, END padded with a, X padded with t — monotone in p.  It is the top 2K bits of
// the padded 16-mer the suffix was sorted by (same padding rule), and the doubling rounds only permute suffixes with equal
// 16-mers, so the table is filled from the sorted keys right after the first sort.
// table[k] = first slot whose bucket key is >= k, k = 0 .. 4^K   (table[4^K] = n)
template <class KeyT>
__global__ void __launch_bounds__(256) k_bucket_fill(const KeyT *__restrict__ keys, int64_t n, int K, uint32_t *__restrict__ table)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;           // n < 2^31, 4^K <= 2^26: 32-bit arithmetic throughout
    const uint32_t nn = (uint32_t)n;
    if (p > nn) return;
    const int sh = 32 - 2 * K;
    auto bk = [&](uint32_t q) -> uint32_t { const uint32_t hi = sizeof(KeyT) == 4 ? (uint32_t)keys[q] : (uint32_t)((uint64_t)keys[q] >> 6); return hi >> sh; };
    const uint32_t cur = p < nn ? bk(p) : 1u << (2 * K);
    uint32_t k = p > 0 ? bk(p - 1) + 1u : 0u;
    for (; k <= cur; k++) table[k] = p;
}

// presence bitmap: the top 2P bits of every sorted key (END / X padding only ever adds P-mers, which is harmless: a set bit
// only means "look the position up"); one atomic per distinct P-mer
template <class KeyT>
__global__ void __launch_bounds__(256) k_present_fill(const KeyT *__restrict__ keys, int64_t n, int P, uint32_t *__restrict__ present)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    auto pk = [&](int64_t q) -> uint32_t { const uint32_t hi = sizeof(KeyT) == 4 ? (uint32_t)keys[q] : (uint32_t)((uint64_t)keys[q] >> 6); return P == 16 ? hi : hi >> (32 - 2 * P); };
    const uint32_t k = pk(p);
    if (p > 0 && pk(p - 1) == k) return;
    atomicOr(present + (k >> 5), 1u << (k & 31));
}

// ------------------------------------------------------------------------------------ driver

static inline int bits_for(int64_t v) { int b = 1; while ((1ll << b) <= v) b++; return b; }

static inline int index_K(int64_t n)
{
    static const int shift = getenv("PMN_INDEX_KSHIFT") ? atoi(getenv("PMN_INDEX_KSHIFT")) : 0;     // experiment: smaller table
    int K = 8; while (K < 13 && (1ll << (2 * K)) < n) K++;
    K -= shift; if (K < 8) K = 8;
    return K;
}

// P of the presence bitmap: 4^P bits for n suffixes keep it under 1/32 full on random sequence (14 for a 5 Mbp genome: 32 MB,
// 16 from 34 Mbp on: 512 MB)
static inline int index_P(int64_t n)
{
    int P = 8; while (P < 16 && (1ll << (2 * P)) < 32 * n) P++;
    return P;
}

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

extern "C" size_t pmn_index_image_bytes(int64_t n_bases)
{
    if (n_bases < 1) return 0;
    return 256 + 2 * up256(4 * (size_t)n_bases) + up256(4 * (((size_t)1 << (2 * index_K(n_bases))) + 1)) + up256((size_t)n_bases + 1) +
           up256(((size_t)1 << (2 * index_P(n_bases))) / 8);
}

int pmn_index_layout(pmn_ctx *c, const pmn_seq *ref, pmn_index *ix)
{
    pmn_tls_stream = c->stream;
    const int64_t n = ref->n;
    if (n < 1) return pmn_set_error(PMN_E_ARG, "index: empty reference");
    if (n > 0x7ffffff0ll) return pmn_set_error(PMN_E_ARG, "index: reference longer than 2^31 bases");
    ix->ctx = c; ix->seq = ref; ix->n = n; ix->K = index_K(n);
    ix->off_sa = 256; ix->off_lcp = ix->off_sa + up256(4 * (size_t)n); ix->off_table = ix->off_lcp + up256(4 * (size_t)n);
    ix->off_skip = ix->off_table + up256(4 * (((size_t)1 << (2 * ix->K)) + 1));
    ix->P = index_P(n);
    ix->off_present = ix->off_skip + up256((size_t)n + 1);
    ix->blob_bytes = pmn_index_image_bytes(n);
    return pmn_pool_get(c, ix->blob, ix->blob_bytes) ? -3 : 0;
}

__global__ void k_index_header(PmnIndexHeader *h, int64_t n, int K, int rounds) { h->magic = PMN_INDEX_MAGIC; h->n = n; h->K = K; h->rounds = rounds; }

int pmn_index_build_impl(pmn_ctx *c, const pmn_seq *ref, pmn_index *ix)
{
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    { int rc = pmn_index_layout(c, ref, ix); if (rc) return rc; }
    const int64_t n = ref->n;
    PackedView T = ref->fwd();

    if (S.k0.ensure(8 * (size_t)n) || S.k1.ensure(8 * (size_t)n) || S.v0.ensure(4 * (size_t)n) || S.v1.ensure(4 * (size_t)n)) return -3;
    if (S.gs.ensure(4 * (size_t)n) || S.rank.ensure(4 * (size_t)(n + 1)) || S.flags.ensure(4 * (size_t)n) ||
        S.list0.ensure(4 * (size_t)n) || S.list1.ensure(4 * (size_t)n) || S.gsn.ensure(4 * (size_t)n)) return -3;
    if (S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(n)) || S.codes.ensure((size_t)n + 64)) return -3;
    if (S.ensure_pinned(64)) return -3;

    PMN_CUDA_OK(cudaEventRecord(c->ev[0], st));
    const unsigned gn = (unsigned)((n + 255) / 256);
    int launches = 0;

    // 1. sort all suffixes by their first 16 symbols
    const int class_first = (!T.has_x && n > 15) ? 1 : 0;       // 4 sort passes over the 32 bits of the 16-mer instead of 5 over all 38
    int where;
    if (class_first) {
        k_sa_keys<true><<<gn, 256, 0, st>>>(T, nullptr, S.k0.as<uint32_t>(), S.v0.as<uint32_t>()); launches++;
        where = pmn_radix_sort(S.k0.as<uint32_t>(), S.v0.as<uint32_t>(), S.k1.as<uint32_t>(), S.v1.as<uint32_t>(), n, 32, S.rs, st, &launches);
    } else {
        k_sa_keys<false><<<gn, 256, 0, st>>>(T, S.k0.as<uint64_t>(), nullptr, S.v0.as<uint32_t>()); launches++;
        where = pmn_radix_sort(S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), S.k1.as<uint64_t>(), S.v1.as<uint32_t>(), n, 38, S.rs, st, &launches);
    }
    if (where < 0) return -3;
    const void *skeys = where ? S.k1.p : S.k0.p;
    const uint32_t *svals = where ? S.v1.as<uint32_t>() : S.v0.as<uint32_t>();
    uint32_t *sa = ix->sa();
    PMN_CUDA_OK(cudaMemcpyAsync(sa, svals, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));
    // what the sorted 16-mers already decide: the bucket table and the LCP entries between different 16-mers
    const int K = ix->K;
    uint8_t *unres = S.codes.as<uint8_t>();
    PMN_CUDA_OK(cudaMemsetAsync(unres, 0, (size_t)n + 32, st));
    PMN_CUDA_OK(cudaMemsetAsync(ix->present(), 0, ((size_t)1 << (2 * ix->P)) / 8, st));
    if (class_first) {
        k_bucket_fill<uint32_t><<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>((const uint32_t *)skeys, n, K, ix->table());
        k_lcp_keys<uint32_t><<<gn, 256, 0, st>>>((const uint32_t *)skeys, svals, n, ix->lcp(), unres);
        k_present_fill<uint32_t><<<gn, 256, 0, st>>>((const uint32_t *)skeys, n, ix->P, ix->present());
    } else {
        k_bucket_fill<uint64_t><<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>((const uint64_t *)skeys, n, K, ix->table());
        k_lcp_keys<uint64_t><<<gn, 256, 0, st>>>((const uint64_t *)skeys, svals, n, ix->lcp(), unres);
        k_present_fill<uint64_t><<<gn, 256, 0, st>>>((const uint64_t *)skeys, n, ix->P, ix->present());
    }
    launches += 3;

    // 2. groups of equal 16-mers -> ranks; slots that still share a group go on the work list
    int32_t *gs = S.gs.as<int32_t>(), *rank = S.rank.as<int32_t>(), *gsn = S.gsn.as<int32_t>();
    uint32_t *flags = S.flags.as<uint32_t>();
    if (class_first) k_sa_heads32<<<gn, 256, 0, st>>>((const uint32_t *)skeys, svals, n, gs);
    else k_sa_heads<<<gn, 256, 0, st>>>((const uint64_t *)skeys, n, gs);
    launches++;
    pmn_scan<int32_t, OpMaxI32, true>(gs, gs, n, S.scan_tmp.as<int32_t>(), st); launches += 3;
    k_sa_init_rank<<<gn, 256, 0, st>>>(sa, gs, n, rank, flags); launches++;

    uint32_t *list = S.list0.as<uint32_t>(), *list_next = S.list1.as<uint32_t>();
    uint32_t *pos = S.v1.as<uint32_t>() == svals ? S.v0.as<uint32_t>() : S.v1.as<uint32_t>();   // a free uint32[n]
    int64_t m = n; const uint32_t *src = nullptr;   // first compaction runs over all slots
    uint32_t *tail = (uint32_t *)S.pinned;
    const int bits_r = bits_for(n + 1);
    int rounds = 0;
    for (int64_t h = 16;; h <<= 1) {
        // compact the flagged entries of the current list (round 0: of all slots)
        pmn_scan<uint32_t, OpAddU32, false>(flags, pos, m, S.scan_tmp.as<uint32_t>(), st); launches += 3;
        k_compact<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(flags, pos, m, src, list_next); launches++;
        PMN_D2H(c, tail, pos + (m - 1), 4);
        PMN_D2H(c, tail + 1, flags + (m - 1), 4);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
        int64_t m2 = (int64_t)tail[0] + tail[1];
        { uint32_t *t = list; list = list_next; list_next = t; }
        src = list; m = m2;
        if (m == 0) break;
        if (h > n) return pmn_set_error(PMN_E_INTERNAL, "index: prefix doubling did not converge");
        rounds++;
        unsigned gm = (unsigned)((m + 255) / 256);
        k_round_keys<<<gm, 256, 0, st>>>(list, m, sa, gs, rank, h, n, bits_r, S.k0.as<uint64_t>(), S.v0.as<uint32_t>()); launches++;
        int w2 = pmn_radix_sort(S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), S.k1.as<uint64_t>(), S.v1.as<uint32_t>(), m, 2 * bits_r, S.rs, st, &launches);
        if (w2 < 0) return -3;
        const uint64_t *rk = w2 ? S.k1.as<uint64_t>() : S.k0.as<uint64_t>();
        const uint32_t *rv = w2 ? S.v1.as<uint32_t>() : S.v0.as<uint32_t>();
        pos = w2 ? S.v0.as<uint32_t>() : S.v1.as<uint32_t>();
        k_round_heads<<<gm, 256, 0, st>>>(rk, list, m, gsn); launches++;
        pmn_scan<int32_t, OpMaxI32, true>(gsn, gsn, m, S.scan_tmp.as<int32_t>(), st); launches += 3;
        k_round_apply<<<gm, 256, 0, st>>>(list, rv, gsn, m, sa, gs, rank, flags); launches++;
    }
    ix->rounds = rounds;

    // 3. the LCP entries the keys left open, in text order
    k_lcp_text<<<(unsigned)(((n + PMN_LCP_CHUNK - 1) / PMN_LCP_CHUNK + 255) / 256), 256, 0, st>>>(T, sa, rank, unres, ix->lcp()); launches++;

    // 4. the skip table (needs the final ranks and the complete LCP array)
    k_skip_e<<<gn, 256, 0, st>>>(rank, ix->lcp(), n, gs);          // the group starts are dead after the last round
    k_skip_fill<<<gn, 256, 0, st>>>(gs, n, ix->skip()); launches += 2;

    k_index_header<<<1, 1, 0, st>>>((PmnIndexHeader *)ix->blob.p, n, K, rounds); launches++;

    PMN_CUDA_OK(cudaEventRecord(c->ev[1], st));
    PMN_CUDA_OK(cudaStreamSynchronize(st));
    PMN_CUDA_OK(cudaGetLastError());
    cudaEventElapsedTime(&ix->ms_build, c->ev[0], c->ev[1]);
    c->launches += launches;
    return 0;
}

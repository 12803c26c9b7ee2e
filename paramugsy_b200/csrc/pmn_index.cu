// pmn_index.cu — reference index on the device:
//   pack (codes -> 2-bit text + x-mask, both strands), suffix array by radix-sort prefix
//   doubling, Kasai-style LCP, and the K-mer bucket table the seeding kernel starts from.
//
// Stands in for MUMmer's suffix-tree construction inside `mummer`, first stage of the
// `nucmer` child process of /root/reference/lib/nucmer/mugsy_nucmer.ml:100.
// Oracle counterpart: oracle/pmn_oracle.c §3 (suffix_cmp, pmo_stage_index) — SA and LCP
// must be bit-identical to it.
//
// Suffix order (same as the oracle): END < a < c < g < t < X_p, every X (non-acgt base or
// record separator) being its own symbol ordered by text position.
#include "pmn_scratch.cuh"

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

int pmn_pack_upload(pmn_ctx *c, pmn_seq *s, const uint8_t *codes_host)
{
    Scratch &S = *c->scratch;
    int64_t n = s->n;
    s->nwords = ((n + 31) / 32 + 3) / 4 * 4 + PMN_PAD_WORDS;
    if (S.codes.ensure((size_t)n + 64)) return -3;
    if (s->w_fwd.ensure(8 * (size_t)s->nwords) || s->xm_fwd.ensure(4 * (size_t)s->nwords) ||
        s->w_rev.ensure(8 * (size_t)s->nwords) || s->xm_rev.ensure(4 * (size_t)s->nwords)) return -3;
    PMN_H2D(c, S.codes.p, codes_host, (size_t)n);
    unsigned g = (unsigned)((s->nwords + 255) / 256);
    k_pack<<<g, 256, 0, c->stream>>>(S.codes.as<uint8_t>(), n, s->w_fwd.as<uint64_t>(), s->xm_fwd.as<uint32_t>(), s->nwords);
    k_revcomp<<<g, 256, 0, c->stream>>>(s->w_fwd.as<uint64_t>(), s->xm_fwd.as<uint32_t>(), n, s->w_rev.as<uint64_t>(), s->xm_rev.as<uint32_t>(), s->nwords);
    c->launches += 2;
    PMN_CUDA_OK(cudaGetLastError());
    PMN_CUDA_OK(cudaStreamSynchronize(c->stream));   // codes_host may be reused by the caller
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
__global__ void __launch_bounds__(256) k_sa_keys(PackedView s, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
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
    keys[i] = (uint64_t)hi << 6 | cls;
    vals[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256) k_sa_heads(const uint64_t *__restrict__ keys, int64_t n, int32_t *__restrict__ headpos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint64_t k = keys[p];
    bool head = p == 0 || k != keys[p - 1] || (k & 63) != CLS_REGULAR;
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

// ------------------------------------------------------------------------------------ LCP (Kasai-style)

// Thread t walks text positions [t*C, (t+1)*C) in order and carries l-1 from one position
// to the next (Kasai et al.): lcp(i+1, phi(i+1)) >= lcp(i, phi(i)) - 1.  Comparisons run
// 32 bases per step on the packed text.
#define PMN_LCP_CHUNK 16
__global__ void __launch_bounds__(256) k_lcp(PackedView s, const uint32_t *__restrict__ sa, const int32_t *__restrict__ rank, int32_t *__restrict__ lcp)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t i0 = t * PMN_LCP_CHUNK;
    if (i0 >= s.n) return;
    int64_t i1 = i0 + PMN_LCP_CHUNK < s.n ? i0 + PMN_LCP_CHUNK : s.n;
    int64_t l = 0;
    for (int64_t i = i0; i < i1; i++) {
        int32_t p = rank[i] - 1;
        if (p == 0) { lcp[0] = 0; l = 0; continue; }
        int64_t j = sa[p - 1];
        l = pmn_lcp(s, i, s, j, l > 0 ? l - 1 : 0, s.n);
        lcp[p] = (int32_t)l;
    }
}

// ------------------------------------------------------------------------------------ K-mer bucket table

// bucket key of SA slot p: first K symbols, END padded with a, X padded with t — monotone in p
__device__ __forceinline__ uint32_t bucket_key(const PackedView &s, int64_t pos, int K)
{
    int v = pmn_valid32(s, pos);
    uint32_t km = (uint32_t)(pmn_window64(s.w, pos) >> (64 - 2 * K));
    if (v >= K) return km;
    uint32_t full = (K == 16) ? ~0u : ((1u << (2 * K)) - 1u);
    uint32_t keep = v ? (full >> (2 * (K - v))) << (2 * (K - v)) : 0u;
    if (pos + v >= s.n) return km & keep;
    return (km & keep) | (full & ~keep);
}

// table[k] = first slot whose bucket key is >= k, k = 0 .. 4^K   (table[4^K] = n)
__global__ void __launch_bounds__(256) k_bucket_fill(PackedView s, const uint32_t *__restrict__ sa, int K, uint32_t *__restrict__ table)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > s.n) return;
    int64_t nb = 1ll << (2 * K);
    int64_t cur = p < s.n ? (int64_t)bucket_key(s, sa[p], K) : nb;
    int64_t prev = p > 0 ? (int64_t)bucket_key(s, sa[p - 1], K) : -1;
    for (int64_t k = prev + 1; k <= cur; k++) table[k] = (uint32_t)p;
}

// ------------------------------------------------------------------------------------ driver

static inline int bits_for(int64_t v) { int b = 1; while ((1ll << b) <= v) b++; return b; }

int pmn_index_build_impl(pmn_ctx *c, const pmn_seq *ref, pmn_index *ix)
{
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    const int64_t n = ref->n;
    if (n < 1) return pmn_set_error(PMN_E_ARG, "index: empty reference");
    if (n > 0x7ffffff0ll) return pmn_set_error(PMN_E_ARG, "index: reference longer than 2^31 bases");
    ix->ctx = c; ix->seq = ref; ix->n = n;
    PackedView T = ref->fwd();

    if (S.k0.ensure(8 * (size_t)n) || S.k1.ensure(8 * (size_t)n) || S.v0.ensure(4 * (size_t)n) || S.v1.ensure(4 * (size_t)n)) return -3;
    if (S.gs.ensure(4 * (size_t)n) || S.rank.ensure(4 * (size_t)(n + 1)) || S.flags.ensure(4 * (size_t)n) ||
        S.list0.ensure(4 * (size_t)n) || S.list1.ensure(4 * (size_t)n) || S.gsn.ensure(4 * (size_t)n)) return -3;
    if (S.scan_tmp.ensure(8 * pmn_scan_scratch_elems(n))) return -3;
    if (ix->sa.ensure(4 * (size_t)n) || ix->lcp.ensure(4 * (size_t)n)) return -3;
    if (S.ensure_pinned(64)) return -3;

    PMN_CUDA_OK(cudaEventRecord(c->ev[0], st));
    const unsigned gn = (unsigned)((n + 255) / 256);
    int launches = 0;

    // 1. sort all suffixes by their first 16 symbols
    k_sa_keys<<<gn, 256, 0, st>>>(T, S.k0.as<uint64_t>(), S.v0.as<uint32_t>()); launches++;
    int where = pmn_radix_sort(S.k0.as<uint64_t>(), S.v0.as<uint32_t>(), S.k1.as<uint64_t>(), S.v1.as<uint32_t>(), n, 38, S.rs, st, &launches);
    if (where < 0) return -3;
    const uint64_t *skeys = where ? S.k1.as<uint64_t>() : S.k0.as<uint64_t>();
    const uint32_t *svals = where ? S.v1.as<uint32_t>() : S.v0.as<uint32_t>();
    uint32_t *sa = ix->sa.as<uint32_t>();
    PMN_CUDA_OK(cudaMemcpyAsync(sa, svals, 4 * (size_t)n, cudaMemcpyDeviceToDevice, st));

    // 2. groups of equal 16-mers -> ranks; slots that still share a group go on the work list
    int32_t *gs = S.gs.as<int32_t>(), *rank = S.rank.as<int32_t>(), *gsn = S.gsn.as<int32_t>();
    uint32_t *flags = S.flags.as<uint32_t>();
    k_sa_heads<<<gn, 256, 0, st>>>(skeys, n, gs); launches++;
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

    // 3. LCP in text order
    k_lcp<<<(unsigned)(((n + PMN_LCP_CHUNK - 1) / PMN_LCP_CHUNK + 255) / 256), 256, 0, st>>>(T, sa, rank, ix->lcp.as<int32_t>()); launches++;

    // 4. bucket table over the first K bases
    int K = 8; while (K < 13 && (1ll << (2 * K)) < n) K++;
    ix->K = K;
    if (ix->table.ensure(4 * ((size_t)1 << (2 * K)) + 16)) return -3;
    k_bucket_fill<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(T, sa, K, ix->table.as<uint32_t>()); launches++;

    PMN_CUDA_OK(cudaEventRecord(c->ev[1], st));
    PMN_CUDA_OK(cudaStreamSynchronize(st));
    PMN_CUDA_OK(cudaGetLastError());
    cudaEventElapsedTime(&ix->ms_build, c->ev[0], c->ev[1]);
    c->launches += launches;
    return 0;
}

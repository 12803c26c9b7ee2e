// pmn_api.cu — the C ABI of include/pmnucmer.h: context, FASTA -> HBM, index, one pair,
// batch, .delta text.  Replaces the `nucmer` child process of
// /root/reference/lib/nucmer/mugsy_nucmer.ml:96-100 (see INTEGRATION.md for the OCaml stub).
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <chrono>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unistd.h>
#include <vector>

#include "pmn_scratch.cuh"

// ------------------------------------------------------------------------------------ errors

static thread_local PmnError g_err;

int pmn_set_error(int code, const char *fmt, ...)
{
    g_err.code = code;
    va_list ap; va_start(ap, fmt); vsnprintf(g_err.msg, sizeof g_err.msg, fmt, ap); va_end(ap);
    return code;
}

extern "C" const char *pmn_last_error(const pmn_ctx *) { return g_err.msg; }
int pmn_last_code() { return g_err.code; }

thread_local cudaStream_t pmn_tls_stream = nullptr;
thread_local long pmn_tls_launches_saved = 0;

static inline double now_ms();
static const double g_t_start = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
static std::atomic<long long> g_allocs{0};
void pmn_count_alloc(size_t bytes, size_t had)
{
    g_allocs++;
    static const bool log = getenv("PMN_ALLOC_LOG") != nullptr;      // which buffers still grow at steady state
    if (log) {
        const double t0 = now_ms();
        fprintf(stderr, "[pmn] t=%.1f ms: allocation of %zu bytes (buffer had %zu)\n", t0 - g_t_start, bytes, had);
    }
}
extern "C" int64_t pmn_alloc_count(void) { return (int64_t)g_allocs.load(); }

extern "C" void pmn_default_opts(pmn_opts *o)
{
    o->minmatch = PMN_DEF_MINMATCH; o->mincluster = PMN_DEF_MINCLUSTER; o->maxgap = PMN_DEF_MAXGAP;
    o->diagdiff = PMN_DEF_DIAGDIFF; o->diagfactor = PMN_DEF_DIAGFACTOR; o->breaklen = PMN_DEF_BREAKLEN;
    o->do_forward = 1; o->do_reverse = 1; o->do_extend = 1; o->do_optimize = 1; o->do_simplify = 1; o->keep_stages = 0; o->post = 0;
}

// ------------------------------------------------------------------------------------ pinned host buffers for large texts

namespace {
struct PinnedPool {
    std::mutex mu;
    std::multimap<size_t, void *> idle;         // capacity -> buffer
    std::map<void *, size_t> all;               // every buffer the pool handed out and has not freed -> capacity
    size_t idle_bytes = 0;
    // Page-locked memory the pool keeps for reuse.  Buffers in the hands of callers are theirs (a batch call returns all its
    // results alive); what comes back beyond the budget is released, so a process never sits on more idle pinned memory than
    // this — 1 GB by default (a C2 step hands back 28 MAF buffers of 16 MB), PMN_PINNED_POOL_MB overrides, 0 = keep nothing.
    size_t budget() const
    {
        static const size_t b = getenv("PMN_PINNED_POOL_MB") ? (size_t)atoll(getenv("PMN_PINNED_POOL_MB")) << 20 : (size_t)1 << 30;
        return b;
    }
};
PinnedPool &pinned_pool() { static PinnedPool *p = new PinnedPool(); return *p; }      // never destroyed: outlives the CUDA runtime's own teardown
}  // namespace

char *pmn_pinned_get(size_t bytes)
{
    size_t cap = (size_t)1 << 20;
    while (cap < bytes + 1) cap <<= 1;
    PinnedPool &P = pinned_pool();
    {
        std::lock_guard<std::mutex> lk(P.mu);
        auto it = P.idle.find(cap);
        if (it != P.idle.end()) { void *p = it->second; P.idle.erase(it); P.idle_bytes -= cap; return (char *)p; }
    }
    void *p = nullptr;
    if (cudaMallocHost(&p, cap) != cudaSuccess) { cudaGetLastError(); pmn_set_error(PMN_E_NOMEM, "cudaMallocHost(%zu) failed", cap); return nullptr; }
    std::lock_guard<std::mutex> lk(P.mu);
    P.all[p] = cap;
    return (char *)p;
}

bool pmn_pinned_put(void *p)
{
    if (!p) return true;
    PinnedPool &P = pinned_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    auto it = P.all.find(p);
    if (it == P.all.end()) return false;
    if (P.idle_bytes + it->second > P.budget()) {        // over the budget: back to the system (cudaFreeHost waits for the device; rare)
        P.all.erase(it);
        if (cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
        return true;
    }
    P.idle.emplace(it->second, p);
    P.idle_bytes += it->second;
    return true;
}

extern "C" void pmn_pinned_pool_stats(int64_t out[3])
{
    if (!out) return;
    PinnedPool &P = pinned_pool();
    std::lock_guard<std::mutex> lk(P.mu);
    size_t live = 0;
    for (auto &kv : P.all) live += kv.second;
    out[0] = (int64_t)P.idle_bytes; out[1] = (int64_t)live; out[2] = (int64_t)P.budget();
}

pmn_result::~pmn_result() { pmn_pinned_put(maf); }

// ------------------------------------------------------------------------------------ context

Scratch *pmn_scratch_new() { return new Scratch(); }
void pmn_scratch_free(Scratch *s)
{
    if (!s) return;
    for (DevBuf *b : s->all()) b->release();
    if (s->pinned) cudaFreeHost(s->pinned);
    delete s;
}

extern "C" int pmn_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// The streams of a scheduler's workers need a hardware work queue each: with the driver's default of 8 connections the 16+
// streams alias and kernels of different pairs serialise behind each other (C2, one GPU: 1004 pairs/s with 8 connections,
// 1225 with 32 at 8 workers, 1505 at 16 workers).  The driver reads the variable when the context is created, so the
// library sets it when it is loaded — it has no effect if the host process initialised CUDA before that and did not set it.
__attribute__((constructor)) static void pmn_set_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }

// How host threads wait for the device.  A scheduler keeps W threads per GPU waiting in cudaStreamSynchronize most of the
// time.  Spinning is the fastest wake-up while every waiting thread has a core of its own (one GPU, 16 cores: 953 pairs/s
// spinning, 935 yielding); on a host with fewer cores than waiting threads the spinners fight for the cores (8 GPUs, 8 ranks of
// 8 workers on 32 cores: 5435 pairs/s spinning, 7004 yielding).  So a scheduler asks for yielding when the host has fewer than
// two cores per worker thread of every visible GPU; PMN_DEVICE_SCHED = spin | yield | block overrides.
void pmn_apply_device_sched(int workers)
{
    const char *sm = getenv("PMN_DEVICE_SCHED");
    unsigned fl;
    if (sm) fl = !strcmp(sm, "block") ? cudaDeviceScheduleBlockingSync : !strcmp(sm, "spin") ? cudaDeviceScheduleSpin : cudaDeviceScheduleYield;
    else {
        const long cores = sysconf(_SC_NPROCESSORS_ONLN);
        const int ndev = std::max(1, pmn_device_count());
        if (workers <= 1 || cores <= 0 || cores >= 2L * workers * ndev) return;
        fl = cudaDeviceScheduleYield;
    }
    if (cudaSetDeviceFlags(fl) != cudaSuccess) cudaGetLastError();
}

static int ctx_init(pmn_ctx *c)
{
    PMN_CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    PMN_CUDA_OK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    PMN_CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    PMN_CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    {   // freed blocks stay in the device's default pool instead of going back to the driver
        cudaMemPool_t mp; unsigned long long keep = ~0ull;
        PMN_CUDA_OK(cudaDeviceGetDefaultMemPool(&mp, c->device));
        PMN_CUDA_OK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    for (auto &e : c->ev) PMN_CUDA_OK(cudaEventCreate(&e));
    c->scratch = pmn_scratch_new();
    c->pool = std::make_shared<DevPool>();
    return 0;
}

cudaStream_t pmn_ctx_prio_stream(pmn_ctx *c, int level)
{
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const int levels = least - greatest;                  // priorities above the default one (numerically lower = higher priority)
    if (levels < 1) return nullptr;
    if (level < 0) level = 0;
    if (level > levels - 1) level = levels - 1;
    if (level > 7) level = 7;
    if (!c->prio_stream[level] && cudaStreamCreateWithPriority(&c->prio_stream[level], cudaStreamNonBlocking, greatest + level) != cudaSuccess) {
        cudaGetLastError(); c->prio_stream[level] = nullptr;
    }
    return c->prio_stream[level];
}

static void ctx_teardown(pmn_ctx *c)
{
    pmn_scratch_free(c->scratch);
    c->pool.reset();
    for (auto &e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->stream2) cudaStreamDestroy(c->stream2);
    for (auto &ps : c->prio_stream) if (ps) cudaStreamDestroy(ps);
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

extern "C" int pmn_ctx_create(int device, pmn_ctx **out)
{
    if (!out) return pmn_set_error(PMN_E_ARG, "pmn_ctx_create: out is NULL");
    *out = nullptr;
    int n = pmn_device_count();
    if (n <= 0) return pmn_set_error(PMN_E_NOGPU, "no CUDA device: libpmnucmer has no CPU path");
    if (device < 0 || device >= n) return pmn_set_error(PMN_E_ARG, "device %d out of range (have %d)", device, n);
    PMN_CUDA_OK(cudaSetDevice(device));
    pmn_apply_device_sched(1);
    cudaDeviceProp prop;
    PMN_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return pmn_set_error(PMN_E_NOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    pmn_ctx *c = new pmn_ctx();
    c->device = device; c->sm_count = prop.multiProcessorCount;
    const int rc = ctx_init(c);
    if (rc) { ctx_teardown(c); return rc; }       // streams and events created so far are destroyed, the context freed
    *out = c;
    return 0;
}

int pmn_pool_get(pmn_ctx *c, DevBuf &b, size_t bytes)
{
    if (b.cap >= bytes) return 0;
    DevPool &P = *c->pool;
    std::lock_guard<std::mutex> lk(P.mu);
    if (b.p) { if (P.bufs.size() >= 256) b.release(); else { P.bufs.push_back(b); b.p = nullptr; b.cap = 0; } }
    // only a buffer of exactly the size class a fresh allocation would get: requests of different classes never
    // take each other's buffers, so the pool stops growing after the first batch
    const size_t want = DevBuf::size_class(bytes);
    int best = -1;
    for (size_t i = 0; i < P.bufs.size(); i++)
        if (P.bufs[i].cap == want) { best = (int)i; break; }
    if (best >= 0) {
        b = P.bufs[(size_t)best];
        P.bufs.erase(P.bufs.begin() + best);
        return 0;
    }
    return b.ensure(bytes);
}

void pmn_pool_put(pmn_ctx *c, DevBuf &b)
{
    if (!b.p) return;
    DevPool &P = *c->pool;
    std::lock_guard<std::mutex> lk(P.mu);
    if (P.bufs.size() >= 256) { b.release(); return; }
    P.bufs.push_back(b);
    b.p = nullptr; b.cap = 0;
}

static inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

extern "C" void *pmn_ctx_stream(const pmn_ctx *c) { return c ? (void *)c->stream : nullptr; }

extern "C" void pmn_ctx_counters(const pmn_ctx *c, int64_t out[4])
{
    if (!c || !out) return;
    out[0] = c->launches; out[1] = c->h2d_bytes; out[2] = c->d2h_bytes; out[3] = c->pairs;
}

extern "C" int64_t pmn_ctx_sync_count(const pmn_ctx *c) { return c ? (int64_t)c->syncs : 0; }

// dependent-free chains of (add, max) on 8 accumulators per thread: the instruction mix of the
// DP inner loop (IADD3 / VIMNMX on the integer pipe)
__global__ void __launch_bounds__(1024) k_int32_peak(int *out, int seed, int iters)
{
    int a0 = seed + threadIdx.x, a1 = a0 ^ 0x55, a2 = a0 + 7, a3 = a0 - 9, a4 = a0 * 3, a5 = a0 + 11, a6 = a0 - 13, a7 = a0 ^ 0x33;
    const int x = seed | 1, y = seed - 3;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            a0 = max(a0 + x, y); a1 = max(a1 + x, y); a2 = max(a2 + x, y); a3 = max(a3 + x, y);
            a4 = max(a4 + x, y); a5 = max(a5 + x, y); a6 = max(a6 + x, y); a7 = max(a7 + x, y);
        }
    }
    int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (r == 0x7fffffff) out[0] = r;      // never true in practice; keeps the chains alive
}

extern "C" int pmn_measure_int32_peak(pmn_ctx *c, double *gops_per_s, double *sm_mhz_effective)
{
    if (!c || !gops_per_s) return pmn_set_error(PMN_E_ARG, "pmn_measure_int32_peak: NULL argument");
    PMN_CUDA_OK(cudaSetDevice(c->device));
    pmn_tls_stream = c->stream;
    Scratch &S = *c->scratch;
    if (S.ex_counters.ensure(128)) return -3;
    const int iters = 2048, blocks = c->sm_count * 2;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        PMN_CUDA_OK(cudaEventRecord(c->ev[11], c->stream));
        k_int32_peak<<<blocks, 1024, 0, c->stream>>>(S.ex_counters.as<int>(), rep + 1, iters);
        PMN_CUDA_OK(cudaEventRecord(c->ev[12], c->stream));
        PMN_CUDA_OK(cudaStreamSynchronize(c->stream));
        float ms; cudaEventElapsedTime(&ms, c->ev[11], c->ev[12]);
        if (rep > 0 && ms < best) best = ms;
    }
    c->launches += 5;
    const double ops = (double)blocks * 1024.0 * iters * 8.0 * 8.0 * 2.0;
    *gops_per_s = ops / (best * 1e-3) / 1e9;
    if (sm_mhz_effective) *sm_mhz_effective = *gops_per_s * 1e9 / ((double)c->sm_count * 128.0) / 1e6;   // if 128 lanes/SM/clk
    return 0;
}

extern "C" void pmn_ctx_destroy(pmn_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    const int device = c->device;
    if (pmn_tls_stream == c->stream || pmn_tls_stream == c->stream2) pmn_tls_stream = nullptr;      // never grow a buffer on a dead stream
    ctx_teardown(c);
    { cudaMemPool_t mp; if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0); }   // unused blocks go back to the driver
}

// ------------------------------------------------------------------------------------ FASTA

// Like `mummer -n`: only a/c/g/t (either case) can match, everything else becomes code X.
// Record ids = first token of the header (what the '>' line of a .delta carries; the
// reference rewrites headers to species.accession, lib/base/m_rewrite_fasta.ml:5-59).
// The host only finds the header lines; bases are translated and packed on the device.
static int find_headers(const char *txt, size_t nb, pmn_seq *s, std::vector<int64_t> &hp)
{
    hp.clear();
    // anything but white space before the first header is an error
    size_t i = 0;
    while (i < nb && (txt[i] == ' ' || txt[i] == '\t' || txt[i] == '\r' || txt[i] == '\n' || txt[i] == '\v' || txt[i] == '\f')) i++;
    if (i == nb) return pmn_set_error(PMN_E_ARG, "FASTA: no records");
    if (txt[i] != '>' || (i > 0 && txt[i - 1] != '\n')) return pmn_set_error(PMN_E_ARG, "FASTA: sequence data before the first '>' header");
    const char *p = txt + i;
    while (p) {
        size_t at = (size_t)(p - txt);
        if (at == 0 || txt[at - 1] == '\n') {
            size_t e = at + 1; while (e < nb && txt[e] != '\n') e++;
            size_t a = at + 1; while (a < e && (txt[a] == ' ' || txt[a] == '\t')) a++;
            size_t b = a; while (b < e && !(txt[b] == ' ' || txt[b] == '\t' || txt[b] == '\r' || txt[b] == '\v' || txt[b] == '\f')) b++;
            hp.push_back((int64_t)at);
            s->ids.emplace_back(txt + a, b - a);
            at = e;
        } else at++;
        p = at < nb ? (const char *)memchr(txt + at, '>', nb - at) : nullptr;
    }
    s->nrec = (int)hp.size();
    if (s->nrec > 32767) return pmn_set_error(PMN_E_ARG, "FASTA: more than 32767 records");
    return 0;
}

extern "C" int pmn_seq_from_fasta(pmn_ctx *c, const char *fasta, size_t bytes, pmn_seq **out)
{
    if (!c || !out || (!fasta && bytes)) return pmn_set_error(PMN_E_ARG, "pmn_seq_from_fasta: NULL argument");
    *out = nullptr;
    PMN_CUDA_OK(cudaSetDevice(c->device));
    std::unique_ptr<pmn_seq> s(new pmn_seq());
    s->ctx = c;
    std::vector<int64_t> hp;
    int rc = find_headers(fasta, bytes, s.get(), hp);
    if (rc) return rc;
    pmn_tls_launches_saved = 0;
    rc = pmn_fasta_to_device(c, s.get(), fasta, bytes, hp);
    c->launches -= pmn_tls_launches_saved; pmn_tls_launches_saved = 0;
    if (rc) { s->w_fwd.release(); s->xm_fwd.release(); s->w_rev.release(); s->xm_rev.release(); s->residues.release(); return rc; }
    *out = s.release();
    return 0;
}

int pmn_read_file(const char *path, std::string &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return pmn_set_error(PMN_E_IO, "cannot open %s: %s", path, strerror(errno));
    char buf[1 << 16]; size_t k;
    out.clear();
    while ((k = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, k);
    int bad = ferror(f); fclose(f);
    if (bad) return pmn_set_error(PMN_E_IO, "read error on %s", path);
    return 0;
}

extern "C" int pmn_seq_from_file(pmn_ctx *c, const char *path, pmn_seq **out)
{
    if (!path) return pmn_set_error(PMN_E_ARG, "pmn_seq_from_file: NULL path");
    std::string txt;
    int rc = pmn_read_file(path, txt);
    if (rc) return rc;
    return pmn_seq_from_fasta(c, txt.data(), txt.size(), out);
}

extern "C" void pmn_seq_free(pmn_seq *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    pmn_pool_put(s->ctx, s->w_fwd); pmn_pool_put(s->ctx, s->xm_fwd); pmn_pool_put(s->ctx, s->w_rev); pmn_pool_put(s->ctx, s->xm_rev);
    pmn_pool_put(s->ctx, s->residues);
    delete s;
}
extern "C" int64_t pmn_seq_bases(const pmn_seq *s) { return s ? s->n : 0; }
extern "C" int pmn_seq_records(const pmn_seq *s) { return s ? s->nrec : 0; }

// ------------------------------------------------------------------------------------ index

extern "C" int pmn_index_build(pmn_ctx *c, const pmn_seq *ref, pmn_index **out)
{
    if (!c || !ref || !out) return pmn_set_error(PMN_E_ARG, "pmn_index_build: NULL argument");
    *out = nullptr;
    PMN_CUDA_OK(cudaSetDevice(c->device));
    std::unique_ptr<pmn_index> ix(new pmn_index());
    const double t0 = now_ms();
    pmn_tls_launches_saved = 0;
    int rc = pmn_index_build_impl(c, ref, ix.get());
    c->launches -= pmn_tls_launches_saved; pmn_tls_launches_saved = 0;
    ix->wall_ms_build = (float)(now_ms() - t0);
    if (rc) { ix->blob.release(); return rc; }
    *out = ix.release();
    return 0;
}

extern "C" void pmn_index_free(pmn_index *ix)
{
    if (!ix) return;
    cudaSetDevice(ix->ctx->device);
    pmn_pool_put(ix->ctx, ix->blob);
    delete ix;
}

extern "C" int64_t pmn_index_size(const pmn_index *ix) { return ix ? ix->n : 0; }

// ---- replication of an index (multi-GPU): the image is one contiguous range of HBM, so a
// collective can send / receive it in place
extern "C" int pmn_index_image(const pmn_index *ix, void **dev_ptr, size_t *bytes)
{
    if (!ix || !dev_ptr || !bytes) return pmn_set_error(PMN_E_ARG, "pmn_index_image: NULL argument");
    *dev_ptr = ix->blob.p; *bytes = ix->blob_bytes;
    return 0;
}

extern "C" int pmn_index_alloc(pmn_ctx *c, const pmn_seq *ref, pmn_index **out)
{
    if (!c || !ref || !out) return pmn_set_error(PMN_E_ARG, "pmn_index_alloc: NULL argument");
    *out = nullptr;
    PMN_CUDA_OK(cudaSetDevice(c->device));
    std::unique_ptr<pmn_index> ix(new pmn_index());
    int rc = pmn_index_layout(c, ref, ix.get());
    if (rc) return rc;
    PMN_CUDA_OK(cudaMemsetAsync(ix->blob.p, 0, 256, c->stream));
    PMN_CUDA_OK(cudaStreamSynchronize(c->stream));
    *out = ix.release();
    return 0;
}

extern "C" int pmn_index_adopt(pmn_index *ix)
{
    if (!ix) return pmn_set_error(PMN_E_ARG, "pmn_index_adopt: NULL index");
    PMN_CUDA_OK(cudaSetDevice(ix->ctx->device));
    PmnIndexHeader h;
    PMN_CUDA_OK(cudaMemcpy(&h, ix->blob.p, sizeof h, cudaMemcpyDeviceToHost));
    if (h.magic != PMN_INDEX_MAGIC || h.n != ix->n || h.K != ix->K)
        return pmn_set_error(PMN_E_ARG, "pmn_index_adopt: the received image does not belong to this reference (n %lld vs %lld, K %d vs %d)",
                             (long long)h.n, (long long)ix->n, h.K, ix->K);
    ix->rounds = h.rounds;
    return 0;
}

extern "C" int pmn_index_copy_sa(const pmn_index *ix, int32_t *sa_out, int32_t *lcp_out)
{
    if (!ix) return pmn_set_error(PMN_E_ARG, "pmn_index_copy_sa: NULL index");
    PMN_CUDA_OK(cudaSetDevice(ix->ctx->device));
    if (sa_out) PMN_CUDA_OK(cudaMemcpy(sa_out, ix->sa(), 4 * (size_t)ix->n, cudaMemcpyDeviceToHost));
    if (lcp_out) PMN_CUDA_OK(cudaMemcpy(lcp_out, ix->lcp(), 4 * (size_t)ix->n, cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------------ one pair

// .delta grammar exactly as the reference parses it: lib/profiles_lib/m_delta.cc:72-92 (two
// header lines), :154-162 ('>' line), :177-185 (seven ints), :187-196 (deltas up to "0");
// lib/profiles/m_delta.ml:76-79,91 needs single spaces and no trailing blanks.
static void write_delta_text(pmn_ctx *c, const pmn_seq *ref, const pmn_seq *qry, const char *ref_path, const char *qry_path, pmn_result *r)
{
    // formatted in the context's staging buffer (worst-case size, its pages are touched once in the life of the context), then
    // copied into a string of the exact size
    std::vector<char> &t = c->text_buf;
    const size_t na = r->al_rows.size() / 10;
    size_t idlen = 0;
    for (auto &x : ref->ids) idlen = std::max(idlen, x.size());
    for (auto &x : qry->ids) idlen = std::max(idlen, x.size());
    const size_t head = strlen(ref_path) + strlen(qry_path) + 16;
    { const size_t need = head + na * (2 * idlen + 64 + 7 * 21 + 4) + r->al_deltas.size() * 12 + 16; if (t.size() < need) t.resize(need + need / 4); }
    char *p = &t[0];
    p += sprintf(p, "%s %s\nNUCMER\n", ref_path, qry_path);
    int64_t prev_r = -1, prev_q = -1, aligned = 0;
    for (size_t k = 0; k < na; k++) {
        const int64_t *a = &r->al_rows[k * 10];
        if (a[0] != prev_r || a[1] != prev_q) {
            *p++ = '>';
            const std::string &ri = ref->ids[(size_t)a[0]], &qi = qry->ids[(size_t)a[1]];
            memcpy(p, ri.data(), ri.size()); p += ri.size(); *p++ = ' ';
            memcpy(p, qi.data(), qi.size()); p += qi.size(); *p++ = ' ';
            p = pmn_fmt_int(p, ref->len[(size_t)a[0]]); *p++ = ' '; p = pmn_fmt_int(p, qry->len[(size_t)a[1]]); *p++ = '\n';
            prev_r = a[0]; prev_q = a[1];
        }
        int64_t sB = a[5], eB = a[6]; const int64_t lenB = qry->len[(size_t)a[1]];
        if (a[2]) { sB = lenB - sB + 1; eB = lenB - eB + 1; }
        p = pmn_fmt_int(p, a[3]); *p++ = ' '; p = pmn_fmt_int(p, a[4]); *p++ = ' '; p = pmn_fmt_int(p, sB); *p++ = ' '; p = pmn_fmt_int(p, eB); *p++ = ' ';
        p = pmn_fmt_int(p, a[7]); *p++ = ' '; p = pmn_fmt_int(p, a[8]); *p++ = ' '; p = pmn_fmt_int(p, a[9]); *p++ = '\n';
        for (int64_t d = r->al_doff[k]; d < r->al_doff[k + 1]; d++) { p = pmn_fmt_int(p, r->al_deltas[(size_t)d]); *p++ = '\n'; }
        *p++ = '0'; *p++ = '\n';
        aligned += a[4] - a[3] + 1;
    }
    r->delta.assign(&t[0], (size_t)(p - &t[0]));
    r->stats.alignments = (int64_t)na;
    r->stats.aligned_ref_bases = aligned;
}

static int check_opts(const pmn_opts *o_in, pmn_opts &o)
{
    if (o_in) o = *o_in; else pmn_default_opts(&o);
    if (!o.do_optimize) return pmn_set_error(PMN_E_ARG, "--nooptimize is not supported");
    if (o.minmatch < 1 || o.maxgap < 0 || o.breaklen < 1 || o.mincluster < 0 || o.diagdiff < 0 || o.diagfactor < 0)
        return pmn_set_error(PMN_E_ARG, "pmn_align: option out of range");
    if (o.post < 0 || o.post > 2) return pmn_set_error(PMN_E_ARG, "pmn_align: post must be 0, 1 or 2");
    return 0;
}

// seeding (unless the anchors are given), clustering, extension, .delta text
static int align_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o_in, const void *given_anchors, int64_t n_given,
                      const char *ref_path, const char *qry_path, pmn_result **out)
{
    *out = nullptr;
    pmn_opts o; { int rc = check_opts(o_in, o); if (rc) return rc; }
    PMN_CUDA_OK(cudaSetDevice(c->device));
    pmn_tls_stream = c->stream;          // buffers grow stream-ordered on THIS context's stream (given-anchors path: before any stage sets it)
    Scratch &S = *c->scratch;
    cudaStream_t st = c->stream;
    std::unique_ptr<pmn_result> r(new pmn_result());
    const double t0 = now_ms();
    pmn_tls_launches_saved = 0;
    long launches0 = c->launches;
    r->stats.ref_bases = ix->n; r->stats.qry_bases = qry->n;
    r->stats.sa_rounds = ix->rounds; r->stats.kmer_bits = 2 * ix->K; r->stats.ms_index = ix->ms_build;

    PMN_CUDA_OK(cudaEventRecord(c->ev[2], st));
    int64_t nanc = 0;
    int rc = 0;
    if (given_anchors || n_given == 0) {
        nanc = n_given;
        if (nanc > 0 && given_anchors != S.anchors.p) {
            if (S.anchors.ensure(16 * (size_t)nanc)) return -3;
            PMN_CUDA_OK(cudaMemcpyAsync(S.anchors.p, given_anchors, 16 * (size_t)nanc, cudaMemcpyDeviceToDevice, st));
        }
    } else {
        rc = pmn_seed_impl(c, ix, qry, &o, &nanc);
        if (rc) return rc;
    }
    PMN_CUDA_OK(cudaEventRecord(c->ev[3], st));
    r->stats.anchors = nanc;
    r->stats.seed_lookups = (given_anchors || n_given == 0) ? 0 : S.seed_lookups;
    r->stats.seed_probes = (given_anchors || n_given == 0) ? 0 : S.seed_probes;
    if (o.keep_stages && nanc > 0) {
        r->anchors.resize((size_t)nanc * 4);
        PMN_D2H(c, r->anchors.data(), S.anchors.p, 16 * (size_t)nanc);
        PMN_CUDA_OK(cudaStreamSynchronize(st));
    }
    rc = pmn_cluster_impl(c, ix, qry, &o, nanc);
    if (rc) return rc;
    PMN_CUDA_OK(cudaEventRecord(c->ev[4], st));
    for (;;) {
        rc = pmn_extend_impl(c, ix, qry, &o, r.get());
        // the traceback arena is sized from what earlier pairs needed; a pair that needs more gets a larger one and the extension
        // (a pure function of the cluster list, which is still in the scratch) runs again
        if (rc == PMN_E_NOMEM && S.arena_retry) { S.arena_retry = false; continue; }
        break;
    }
    if (rc) return rc;
    PMN_CUDA_OK(cudaEventRecord(c->ev[5], st));
    PMN_CUDA_OK(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&r->stats.ms_seed, c->ev[2], c->ev[3]);
    cudaEventElapsedTime(&r->stats.ms_cluster, c->ev[3], c->ev[4]);
    cudaEventElapsedTime(&r->stats.ms_extend, c->ev[4], c->ev[5]);
    cudaEventElapsedTime(&r->stats.ms_total, c->ev[2], c->ev[5]);
    if (!given_anchors && n_given != 0 && qry->n >= o.minmatch) { if (cudaEventElapsedTime(&r->stats.ms_seed_kernel, c->ev[6], c->ev[7]) != cudaSuccess) { cudaGetLastError(); r->stats.ms_seed_kernel = 0; } }
    c->pairs++;
    const double t1 = now_ms();
    write_delta_text(c, ix->seq, qry, ref_path ? ref_path : "ref", qry_path ? qry_path : "qry", r.get());
    c->launches -= pmn_tls_launches_saved; pmn_tls_launches_saved = 0;
    r->stats.wall_ms_text = (float)(now_ms() - t1);
    if (o.post) {
        const double t2 = now_ms();
        int prc = pmn_post_impl(c, ix->seq, qry, ref_path ? ref_path : "ref", qry_path ? qry_path : "qry", o.post, r.get());
        if (prc) return prc;
        c->launches -= pmn_tls_launches_saved; pmn_tls_launches_saved = 0;
        r->stats.wall_ms_post = (float)(now_ms() - t2);
    }
    r->stats.kernel_launches = c->launches - launches0;
    r->stats.wall_ms_align = (float)(now_ms() - t0);
    r->stats.wall_ms_index = ix->wall_ms_build;
    *out = r.release();
    return 0;
}

extern "C" int pmn_align(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o_in,
                         const char *ref_path, const char *qry_path, pmn_result **out)
{
    if (!c || !ix || !qry || !out) return pmn_set_error(PMN_E_ARG, "pmn_align: NULL argument");
    return align_impl(c, ix, qry, o_in, nullptr, -1, ref_path, qry_path, out);
}

// ---- one large pair on several GPUs: the query positions are sharded, everything else is replicated
extern "C" int pmn_seed_part(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o_in, int part, int nparts,
                             void **dev_anchors, int64_t *n_anchors)
{
    if (!c || !ix || !qry || !dev_anchors || !n_anchors || nparts < 1 || part < 0 || part >= nparts) return pmn_set_error(PMN_E_ARG, "pmn_seed_part: bad argument");
    pmn_opts o; { int rc = check_opts(o_in, o); if (rc) return rc; }
    PMN_CUDA_OK(cudaSetDevice(c->device));
    int rc = pmn_seed_impl(c, ix, qry, &o, n_anchors, part, nparts);
    if (rc) return rc;
    PMN_CUDA_OK(cudaStreamSynchronize(c->stream));
    *dev_anchors = *n_anchors > 0 ? c->scratch->anchors.p : nullptr;
    return 0;
}

extern "C" int pmn_align_anchors(pmn_ctx *c, const pmn_index *ix, const pmn_seq *qry, const pmn_opts *o_in, const void *dev_anchors, int64_t n_anchors,
                                 const char *ref_path, const char *qry_path, pmn_result **out)
{
    if (!c || !ix || !qry || !out || n_anchors < 0 || (n_anchors > 0 && !dev_anchors)) return pmn_set_error(PMN_E_ARG, "pmn_align_anchors: bad argument");
    return align_impl(c, ix, qry, o_in, dev_anchors, n_anchors, ref_path, qry_path, out);
}

extern "C" const char *pmn_result_delta(const pmn_result *r, size_t *len) { if (len) *len = r ? r->delta.size() : 0; return r ? r->delta.c_str() : ""; }
extern "C" const char *pmn_result_filtered(const pmn_result *r, size_t *len) { if (len) *len = r ? r->filtered.size() : 0; return r ? r->filtered.data() : nullptr; }
extern "C" const char *pmn_result_maf(const pmn_result *r, size_t *len) { if (len) *len = r ? r->maf_len : 0; return r ? r->maf : nullptr; }
extern "C" void pmn_result_stats(const pmn_result *r, pmn_stats *out) { if (r && out) *out = r->stats; }
extern "C" void pmn_result_free(pmn_result *r) { delete r; }

extern "C" int64_t pmn_result_n_anchors(const pmn_result *r) { return r ? (int64_t)r->anchors.size() / 4 : 0; }
extern "C" int pmn_result_copy_anchors(const pmn_result *r, int32_t *out)
{
    if (!r || !out) return pmn_set_error(PMN_E_ARG, "NULL argument");
    memcpy(out, r->anchors.data(), r->anchors.size() * 4); return 0;
}
extern "C" int64_t pmn_result_n_clusters(const pmn_result *r) { return r ? (int64_t)r->cl_tag.size() : 0; }
extern "C" int64_t pmn_result_n_cluster_matches(const pmn_result *r) { return r ? (int64_t)r->cl_matches.size() / 3 : 0; }
extern "C" int pmn_result_copy_clusters(const pmn_result *r, int32_t *matches, int32_t *off, int32_t *tag)
{
    if (!r) return pmn_set_error(PMN_E_ARG, "NULL argument");
    if (matches) memcpy(matches, r->cl_matches.data(), r->cl_matches.size() * 4);
    if (off) memcpy(off, r->cl_off.data(), r->cl_off.size() * 4);
    if (tag) memcpy(tag, r->cl_tag.data(), r->cl_tag.size() * 4);
    return 0;
}
extern "C" int64_t pmn_result_n_alignments(const pmn_result *r) { return r ? (int64_t)r->al_rows.size() / 10 : 0; }
extern "C" int64_t pmn_result_n_deltas(const pmn_result *r) { return r ? (int64_t)r->al_deltas.size() : 0; }
extern "C" int pmn_result_copy_alignments(const pmn_result *r, int64_t *rows, int64_t *doff, int64_t *deltas)
{
    if (!r) return pmn_set_error(PMN_E_ARG, "NULL argument");
    if (rows) memcpy(rows, r->al_rows.data(), r->al_rows.size() * 8);
    if (doff) memcpy(doff, r->al_doff.data(), r->al_doff.size() * 8);
    if (deltas) for (size_t k = 0; k < r->al_deltas.size(); k++) deltas[k] = r->al_deltas[k];
    return 0;
}

// ------------------------------------------------------------------------------------ file level

int pmn_write_file_atomic(const char *path, const char *data, size_t len)
{
    std::string tmp = std::string(path) + ".tmp." + std::to_string((long)getpid());
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return pmn_set_error(PMN_E_IO, "cannot create %s: %s", tmp.c_str(), strerror(errno));
    size_t k = fwrite(data, 1, len, f);
    int bad = (k != len) | (fclose(f) != 0);
    if (bad) { unlink(tmp.c_str()); return pmn_set_error(PMN_E_IO, "write error on %s", tmp.c_str()); }
    if (rename(tmp.c_str(), path) != 0) { unlink(tmp.c_str()); return pmn_set_error(PMN_E_IO, "cannot rename to %s: %s", path, strerror(errno)); }
    return 0;
}

static int batch_files(pmn_ctx *c, int n, const char *const *refs, const char *const *qrys, const char *const *outs, const char *const *mafs, const pmn_opts *o)
{
    if (!c || n < 0 || (n && (!refs || !qrys || !outs))) return pmn_set_error(PMN_E_ARG, "pmn_align_batch: bad argument");
    if (mafs && !(o && o->post)) return pmn_set_error(PMN_E_ARG, "pmn_worker_batch: MAF output needs pmn_opts.post = 1 or 2");
    // pairs are processed grouped by reference so that every index is built once; sequences
    // used several times are packed once
    std::map<std::string, pmn_seq *> seqs;
    std::map<std::string, pmn_index *> idx;
    int rc = 0;
    auto get_seq = [&](const char *p, pmn_seq **s) -> int {
        auto it = seqs.find(p);
        if (it != seqs.end()) { *s = it->second; return 0; }
        int e = pmn_seq_from_file(c, p, s);
        if (!e) seqs[p] = *s;
        return e;
    };
    for (int i = 0; i < n; i++)
        if (!refs[i] || !qrys[i] || !outs[i]) return pmn_set_error(PMN_E_ARG, "pmn_align_batch: NULL path in pair %d", i);
    std::vector<int> order(n);
    for (int i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return strcmp(refs[a], refs[b]) < 0; });
    std::string cur_ref; pmn_index *cur_ix = nullptr;
    for (int oi = 0; oi < n && !rc; oi++) {
        int i = order[oi];
        if (!cur_ix || cur_ref != refs[i]) {
            if (cur_ix) { pmn_index_free(cur_ix); cur_ix = nullptr; }
            pmn_seq *rs; rc = get_seq(refs[i], &rs); if (rc) break;
            rc = pmn_index_build(c, rs, &cur_ix); if (rc) break;
            cur_ref = refs[i];
        }
        pmn_seq *qs; rc = get_seq(qrys[i], &qs); if (rc) break;
        pmn_result *res = nullptr;
        rc = pmn_align(c, cur_ix, qs, o, refs[i], qrys[i], &res); if (rc) break;
        if (mafs) {         // the files one mugsy_nucmer process leaves behind: delta_out is the FILTERED delta (mugsy_nucmer.ml:128-130), maf_out its MAF
            rc = pmn_write_file_atomic(outs[i], res->filtered.data(), res->filtered.size());
            if (!rc && mafs[i]) rc = pmn_write_file_atomic(mafs[i], res->maf, res->maf_len);
        } else rc = pmn_write_file_atomic(outs[i], res->delta.data(), res->delta.size());
        pmn_result_free(res);
    }
    if (cur_ix) pmn_index_free(cur_ix);
    for (auto &kv : seqs) pmn_seq_free(kv.second);
    return rc;
}

extern "C" int pmn_align_batch(pmn_ctx *c, int n, const char *const *refs, const char *const *qrys, const char *const *outs, const pmn_opts *o)
{
    return batch_files(c, n, refs, qrys, outs, nullptr, o);
}

extern "C" int pmn_worker_batch(pmn_ctx *c, int n, const char *const *refs, const char *const *qrys, const char *const *delta_outs,
                                const char *const *maf_outs, const pmn_opts *o)
{
    if (!maf_outs) return pmn_set_error(PMN_E_ARG, "pmn_worker_batch: bad argument");
    return batch_files(c, n, refs, qrys, delta_outs, maf_outs, o);
}

extern "C" int pmn_align_pair(pmn_ctx *c, const char *ref_fasta_path, const char *qry_fasta_path, const pmn_opts *o, const char *out_delta_path)
{
    const char *r[1] = { ref_fasta_path }, *q[1] = { qry_fasta_path }, *d[1] = { out_delta_path };
    return pmn_align_batch(c, 1, r, q, d, o);
}

// pmn_sched.cu — in-process batch scheduler: the unit of work of the reference's fan-out
// (Nucmer_task.t.searches, /root/reference/lib/base/nucmer_task.ml:6, batched by
// run_nucmers, lib/base/job_processor.ml:128-154, `-cores N` workers at a time,
// lib/base/queued_task_server.ml:57-64) run by W worker threads that share one GPU.
//
// Every worker owns a context (stream + scratch), so the serial phases of one pair (host
// round trips between stages, the single-warp stitcher, the tail of the extension wave)
// overlap with the wide kernels of other pairs.  Genomes are packed once, every reference
// index is built once by whichever worker needs it first and is shared read-only.
#include <algorithm>
#include <atomic>
#include <map>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pmn_scratch.cuh"

namespace {

enum { ST_NONE = 0, ST_BUSY = 1, ST_READY = 2, ST_FAILED = 3 };

struct Slot {                       // a lazily built shared object (packed genome or index)
    int state = ST_NONE;
    void *obj = nullptr;
    int users = 0;                  // pairs that still need it
};

}  // namespace

struct pmn_sched {
    int device = 0;
    std::vector<pmn_ctx *> ctx;
    std::mutex mu;
    std::condition_variable cv;
    std::string err; int err_code = 0;
    // Packing a genome and building an index need large scratch that aligning a pair does not.  Which worker
    // gets to build is a race, so instead of every worker growing its own copy over many batches the
    // scheduler owns a few sets that the builder borrows: the working set is complete after the first batch.
    std::vector<Scratch *> build_scratch;
    // Pairs differ a little in size and land on workers at random, so each worker's grow-only scratch would keep
    // meeting "its largest pair so far" for many batches.  Workers publish the capacities they ended up with and
    // pre-grow to the largest any of them has seen: the working set of the whole scheduler settles with the first batch.
    std::vector<size_t> hiwater;
    std::mutex bmu;
    std::condition_variable bcv;
    size_t total_mem = 0;               // device memory, read once (bounds the number of live indexes)
};

namespace {
struct BorrowedScratch {
    pmn_sched *s; pmn_ctx *c; Scratch *own;
    BorrowedScratch(pmn_sched *s_, pmn_ctx *c_) : s(s_), c(c_), own(c_->scratch)
    {
        std::unique_lock<std::mutex> lk(s->bmu);
        s->bcv.wait(lk, [&] { return !s->build_scratch.empty(); });
        c->scratch = s->build_scratch.back(); s->build_scratch.pop_back();
    }
    ~BorrowedScratch()
    {
        cudaStreamSynchronize(c->stream);
        { std::lock_guard<std::mutex> lk(s->bmu); s->build_scratch.push_back(c->scratch); }
        c->scratch = own;
        s->bcv.notify_all();
    }
};
}  // namespace

extern "C" int pmn_sched_create(int device, int workers, pmn_sched **out)
{
    if (!out || workers < 1 || workers > 64) return pmn_set_error(PMN_E_ARG, "pmn_sched_create: bad argument");
    *out = nullptr;
    pmn_sched *s = new pmn_sched();
    s->device = device;
    if (cudaSetDevice(device) == cudaSuccess) pmn_apply_device_sched(workers); else cudaGetLastError();      // before the workers' streams exist
    for (int k = 0; k < workers; k++) {
        pmn_ctx *c = nullptr;
        int rc = pmn_ctx_create(device, &c);
        if (rc) { for (pmn_ctx *x : s->ctx) pmn_ctx_destroy(x); delete s; return rc; }
        if (!s->ctx.empty()) c->pool = s->ctx[0]->pool;
        s->ctx.push_back(c);
    }
    for (int k = 0; k < std::min(workers, 8); k++) s->build_scratch.push_back(pmn_scratch_new());
    { size_t free_b = 0, total_b = 0; if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) s->total_mem = total_b; else cudaGetLastError(); }
    *out = s;
    return 0;
}

extern "C" void pmn_sched_destroy(pmn_sched *s)
{
    if (!s) return;
    if (!s->ctx.empty()) cudaSetDevice(s->ctx[0]->device);
    for (Scratch *b : s->build_scratch) pmn_scratch_free(b);
    for (pmn_ctx *c : s->ctx) pmn_ctx_destroy(c);
    delete s;
}

extern "C" int pmn_sched_workers(const pmn_sched *s) { return s ? (int)s->ctx.size() : 0; }
extern "C" pmn_ctx *pmn_sched_ctx(const pmn_sched *s, int k) { return (s && k >= 0 && k < (int)s->ctx.size()) ? s->ctx[(size_t)k] : nullptr; }

extern "C" void pmn_sched_counters(const pmn_sched *s, int64_t out[4])
{
    if (!s || !out) return;
    out[0] = out[1] = out[2] = out[3] = 0;
    for (pmn_ctx *c : s->ctx) { int64_t t[4]; pmn_ctx_counters(c, t); for (int k = 0; k < 4; k++) out[k] += t[k]; }
}

extern "C" int64_t pmn_sched_sync_count(const pmn_sched *s)
{
    int64_t n = 0;
    if (s) for (pmn_ctx *c : s->ctx) n += pmn_ctx_sync_count(c);
    return n;
}

// The common engine: genomes either as host FASTA buffers (packed on demand) or as resident
// pmn_seq handles; pairs as index pairs into the genome list.
static int sched_run(pmn_sched *s, int ng, const char *const *fasta, const size_t *bytes, const pmn_seq *const *resident,
                     const pmn_index *const *given_idx, const char *const *names, int np, const int32_t *ref, const int32_t *qry, const pmn_opts *opts, pmn_result **out)
{
    if (!s || ng < 0 || np < 0 || (np && (!ref || !qry || !out)) || (!fasta && !resident && ng)) return pmn_set_error(PMN_E_ARG, "pmn_sched: bad argument");
    for (int p = 0; p < np; p++) {
        out[p] = nullptr;
        if (ref[p] < 0 || ref[p] >= ng || qry[p] < 0 || qry[p] >= ng) return pmn_set_error(PMN_E_ARG, "pmn_sched: pair %d names a genome out of range", p);
        if (resident && (!resident[ref[p]] || !resident[qry[p]])) return pmn_set_error(PMN_E_ARG, "pmn_sched: pair %d names a genome that is not resident", p);
    }
    std::vector<Slot> seqs((size_t)ng), idx((size_t)ng);
    for (int p = 0; p < np; p++) { seqs[(size_t)ref[p]].users++; seqs[(size_t)qry[p]].users++; idx[(size_t)ref[p]].users++; }
    if (resident) for (int g = 0; g < ng; g++) { seqs[(size_t)g].state = ST_READY; seqs[(size_t)g].obj = (void *)resident[g]; }
    // indexes the caller already holds (built here or received from another GPU) are used as they are and never freed
    if (given_idx) for (int g = 0; g < ng; g++) if (given_idx[g]) { idx[(size_t)g].state = ST_READY; idx[(size_t)g].obj = (void *)given_idx[g]; idx[(size_t)g].users += 1 << 30; }
    // pairs in reference order, so that at most a few indexes are alive at a time
    std::vector<int> order((size_t)np);
    for (int p = 0; p < np; p++) order[(size_t)p] = p;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ref[a] < ref[b]; });
    std::atomic<int> next{0};
    s->err.clear(); s->err_code = 0;

    // Indexes alive at a time: as many as fit a quarter of the device memory (8.25 B/base each), between 2 and 16.  A batch
    // over a handful of bacterial genomes (C2: 7 references of 107 MB) then builds all its indexes side by side at the start —
    // wide, HBM-bound kernels that fill the GPU — and no pair ever waits for a build in the tail of the batch.
    int MAX_LIVE_INDEXES = 4;
    {
        int64_t longest = 1;
        for (int p = 0; p < np; p++) {
            const int r = ref[p];
            const int64_t nb = resident ? resident[r]->n : (int64_t)bytes[r];
            longest = std::max(longest, nb);
        }
        if (s->total_mem) {
            const size_t per = pmn_index_image_bytes(longest) + 1;
            MAX_LIVE_INDEXES = (int)std::min<size_t>(16, std::max<size_t>(2, s->total_mem / 4 / per));
        }
        if (const char *e = getenv("PMN_SCHED_LIVE_INDEXES")) MAX_LIVE_INDEXES = std::max(1, atoi(e));
    }
    // few pairs: every pair's latency counts (thread-per-job windows up to 4096 cells, the rest one warp each); many pairs in
    // flight: the instructions issued count (pmn_ctx::tpj_cells)
    for (pmn_ctx *c : s->ctx) c->tpj_cells = np > 8 ? 10000 : 4096;
    int live_indexes = 0;                      // guarded by s->bmu
    int builds_done = 0;                       // opening builds finished (guarded by s->bmu)
    static const int build_width = getenv("PMN_SCHED_BUILD_WIDTH") ? std::max(0, atoi(getenv("PMN_SCHED_BUILD_WIDTH"))) : 0;
    std::atomic<int> failed{0};                // some worker gave up: nobody may keep waiting for an index slot
    std::vector<int> first_refs;               // the first distinct references in processing order that have no index yet
    for (int k = 0; k < np && (int)first_refs.size() < MAX_LIVE_INDEXES; k++) {
        const int r = ref[order[(size_t)k]];
        if (idx[(size_t)r].state == ST_NONE && std::find(first_refs.begin(), first_refs.end(), r) == first_refs.end()) first_refs.push_back(r);
    }
    // get a shared object: build it if nobody has, wait if somebody is
    auto acquire = [&](std::vector<Slot> &v, int g, auto &&build) -> void * {
        std::unique_lock<std::mutex> lk(s->mu);
        Slot &sl = v[(size_t)g];
        for (;;) {
            if (sl.state == ST_READY) return sl.obj;
            if (sl.state == ST_FAILED) return nullptr;
            if (sl.state == ST_NONE) {
                sl.state = ST_BUSY;
                lk.unlock();
                void *o = build();
                lk.lock();
                sl.obj = o; sl.state = o ? ST_READY : ST_FAILED;
                if (!o && !s->err_code) { const int ec = pmn_last_code(); s->err_code = ec < 0 ? ec : PMN_E_INTERNAL; s->err = pmn_last_error(nullptr); }   // the builder ran on this thread: its code and message
                if (!o) { std::lock_guard<std::mutex> lk2(s->bmu); failed.store(1); s->bcv.notify_all(); }      // under bmu: a waiter between its predicate and its block cannot miss it
                s->cv.notify_all();
                return o;
            }
            s->cv.wait(lk);
        }
    };

    const int W_active = (int)std::min<size_t>(s->ctx.size(), (size_t)std::max(1, np));
    auto worker = [&](int w) {
        pmn_ctx *c = s->ctx[(size_t)w];
        cudaSetDevice(c->device);
        // packing and index builds run on a stream the device schedules ahead of the pair kernels (pairs wait for them)
        struct OnStream {
            pmn_ctx *c; cudaStream_t saved;
            OnStream(pmn_ctx *c_, cudaStream_t st) : c(c_), saved(c_->stream) { if (st) c->stream = st; }
            ~OnStream() { c->stream = saved; }
        };
        static const bool use_prio = !(getenv("PMN_SCHED_PRIORITIES") && !strcmp(getenv("PMN_SCHED_PRIORITIES"), "0"));
        auto pack = [&](int g) { return acquire(seqs, g, [&]() -> void * {
            BorrowedScratch b(s, c); OnStream on(c, use_prio ? pmn_ctx_prio_stream(c, 0) : nullptr);
            pmn_seq *x = nullptr; const int rc = pmn_seq_from_fasta(c, fasta[g], bytes[g], &x);
            return rc ? nullptr : (void *)x; }); };
        // genomes given as FASTA text: the workers pack them side by side first (H2D + parse + 2-bit pack per genome),
        // instead of every worker waiting for the reference of the first pairs and then packing its query alone
        if (!resident) for (int g = w; g < ng; g += W_active) if (seqs[(size_t)g].users > 0 && !pack(g)) return;
        auto get_index = [&](int r, pmn_seq *rs, int level = 0) {
            return (pmn_index *)acquire(idx, r, [&]() -> void * {
                // at most MAX_LIVE_INDEXES indexes built by this run are alive at a time: pairs are taken in reference order, so the
                // holders of the oldest ones finish without needing another; bounds the memory (8.25 B/base each) and keeps the
                // number of index images the pool ever holds fixed, i.e. no allocation in later batches
                // PMN_SCHED_BUILD_WIDTH = n: at most n of the opening builds run at a time, in reference order (0 = all at once)
                const auto it = std::find(first_refs.begin(), first_refs.end(), r);
                const int pos = build_width > 0 && it != first_refs.end() ? (int)(it - first_refs.begin()) : -1;
                { std::unique_lock<std::mutex> lk(s->bmu); s->bcv.wait(lk, [&] { return (live_indexes < MAX_LIVE_INDEXES && (pos < 0 || builds_done + build_width > pos)) || failed.load(); }); if (failed.load()) return nullptr; live_indexes++; }
                BorrowedScratch b(s, c); OnStream on(c, use_prio ? pmn_ctx_prio_stream(c, level) : nullptr);
                pmn_index *x = nullptr;
                const int rc = pmn_index_build(c, rs, &x);
                if (pos >= 0 || rc) { std::lock_guard<std::mutex> lk(s->bmu); if (pos >= 0) builds_done++; if (rc) live_indexes--; s->bcv.notify_all(); }
                if (rc) return nullptr;
                return (void *)x; });
        };
        // the first MAX_LIVE_INDEXES references of the batch are indexed side by side by the first workers, so that the batch
        // does not open with every worker waiting for one build (the gaps of one build are filled by the others)
        // ... in reference order of device priority: the first reference's index — most pairs wait for it — is ready first
        if (w < (int)first_refs.size()) { const int r = first_refs[(size_t)w]; pmn_seq *rs = (pmn_seq *)pack(r); if (!rs || !get_index(r, rs, w)) return; }
        for (;;) {
            { std::lock_guard<std::mutex> lk(s->mu); if (s->err_code) return; }
            const int k = next.fetch_add(1);
            if (k >= np) return;
            const int p = order[(size_t)k], r = ref[p], q = qry[p];
            pmn_seq *rs = (pmn_seq *)pack(r); if (!rs) return;
            pmn_seq *qs = (pmn_seq *)pack(q); if (!qs) return;
            pmn_index *ix = get_index(r, rs);
            if (!ix) return;
            pmn_result *res = nullptr;
            {
                std::vector<DevBuf *> mine = c->scratch->all();
                std::vector<size_t> want;
                { std::lock_guard<std::mutex> lk(s->bmu); if (s->hiwater.size() != mine.size()) s->hiwater.assign(mine.size(), 0); want = s->hiwater; }
                pmn_tls_stream = c->stream;
                for (size_t i = 0; i < mine.size(); i++) if (want[i] > mine[i]->cap) mine[i]->grow_to(want[i]);
            }
            int rc = pmn_align(c, ix, qs, opts, names ? names[r] : nullptr, names ? names[q] : nullptr, &res);
            {
                std::vector<DevBuf *> mine = c->scratch->all();
                std::lock_guard<std::mutex> lk(s->bmu);
                for (size_t i = 0; i < mine.size() && i < s->hiwater.size(); i++) s->hiwater[i] = std::max(s->hiwater[i], mine[i]->cap);
            }
            std::lock_guard<std::mutex> lk(s->mu);
            if (rc) {
                if (!s->err_code) { s->err_code = rc; s->err = pmn_last_error(nullptr); }
                std::lock_guard<std::mutex> lk2(s->bmu);      // the store must not land between a waiter's predicate check and its block (lost wake-up)
                failed.store(1); s->bcv.notify_all();
                return;
            }
            out[p] = res;
            // the last user of an index / of a genome packed by this run frees it
            if (--idx[(size_t)r].users == 0) {
                pmn_index_free(ix); idx[(size_t)r].obj = nullptr; idx[(size_t)r].state = ST_NONE;
                { std::lock_guard<std::mutex> lk2(s->bmu); live_indexes--; }
                s->bcv.notify_all();
            }
            if (!resident) for (int g : { r, q }) if (--seqs[(size_t)g].users == 0) { pmn_seq_free((pmn_seq *)seqs[(size_t)g].obj); seqs[(size_t)g].obj = nullptr; seqs[(size_t)g].state = ST_NONE; }
        }
    };

    const int W = W_active;
    std::vector<std::thread> th;
    for (int w = 1; w < W; w++) th.emplace_back(worker, w);
    worker(0);
    for (auto &t : th) t.join();
    {   // level the build-scratch sets: a set that only comes into use when several builds coincide is full size by then
        std::lock_guard<std::mutex> lk(s->bmu);
        std::vector<size_t> mx;
        for (Scratch *b : s->build_scratch) { auto v = b->all(); if (mx.size() != v.size()) mx.assign(v.size(), 0); for (size_t i = 0; i < v.size(); i++) mx[i] = std::max(mx[i], v[i]->cap); }
        pmn_tls_stream = s->ctx[0]->stream;
        for (Scratch *b : s->build_scratch) { auto v = b->all(); for (size_t i = 0; i < v.size(); i++) if (mx[i] > v[i]->cap) v[i]->grow_to(mx[i]); }
        cudaStreamSynchronize(s->ctx[0]->stream);
    }

    if (s->err_code) {
        for (int p = 0; p < np; p++) { pmn_result_free(out[p]); out[p] = nullptr; }
        for (auto &sl : idx) if (sl.obj && sl.users < (1 << 30)) pmn_index_free((pmn_index *)sl.obj);
        if (!resident) for (auto &sl : seqs) if (sl.obj) pmn_seq_free((pmn_seq *)sl.obj);
        return pmn_set_error(s->err_code, "%s", s->err.c_str());
    }
    return 0;
}

extern "C" int pmn_sched_align_fasta(pmn_sched *s, int n_genomes, const char *const *fasta, const size_t *bytes, const char *const *names,
                                     int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out)
{
    if (n_genomes && (!fasta || !bytes)) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_fasta: NULL genome list");
    return sched_run(s, n_genomes, fasta, bytes, nullptr, nullptr, names, n_pairs, ref, qry, o, out);
}

extern "C" int pmn_sched_align_seqs(pmn_sched *s, int n_genomes, const pmn_seq *const *seqs, const char *const *names,
                                    int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out)
{
    if (!s || (n_genomes && !seqs)) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_seqs: NULL argument");
    for (int g = 0; g < n_genomes; g++) if (seqs[g] && seqs[g]->ctx->device != s->device) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_seqs: genome %d is not resident on device %d", g, s->device);
    return sched_run(s, n_genomes, nullptr, nullptr, seqs, nullptr, names, n_pairs, ref, qry, o, out);
}

extern "C" int pmn_sched_align_indexed(pmn_sched *s, int n_genomes, const pmn_seq *const *seqs, const pmn_index *const *indexes, const char *const *names,
                                       int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out)
{
    if (!s || (n_genomes && !seqs)) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_indexed: NULL argument");
    for (int g = 0; g < n_genomes; g++) if (seqs[g] && seqs[g]->ctx->device != s->device) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_indexed: genome %d is not resident on device %d", g, s->device);
    return sched_run(s, n_genomes, nullptr, nullptr, seqs, indexes, names, n_pairs, ref, qry, o, out);
}

// File level, one call per Nucmer_task.t.searches (lib/base/nucmer_task.ml:6,48-59): every distinct
// FASTA is read and packed once, every .delta is written atomically (tmp + rename).
extern "C" int pmn_sched_align_files(pmn_sched *s, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                                     const char *const *out_delta_paths, const pmn_opts *o)
{
    if (!s || n < 0 || (n && (!ref_fasta_paths || !qry_fasta_paths || !out_delta_paths))) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_files: bad argument");
    std::map<std::string, int> id;
    std::vector<std::string> text; std::vector<const char *> names;
    std::vector<int32_t> ref((size_t)n), qry((size_t)n);
    auto genome = [&](const char *path, int32_t *g) -> int {
        if (!path) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_files: NULL path");
        auto it = id.find(path);
        if (it == id.end()) {
            std::string t; int rc = pmn_read_file(path, t); if (rc) return rc;
            it = id.emplace(path, (int)text.size()).first; text.push_back(std::move(t)); names.push_back(path);
        }
        *g = it->second; return 0;
    };
    for (int i = 0; i < n; i++) {
        int rc = genome(ref_fasta_paths[i], &ref[(size_t)i]); if (rc) return rc;
        rc = genome(qry_fasta_paths[i], &qry[(size_t)i]); if (rc) return rc;
        if (!out_delta_paths[i]) return pmn_set_error(PMN_E_ARG, "pmn_sched_align_files: NULL output path");
    }
    std::vector<const char *> fa(text.size()); std::vector<size_t> nb(text.size());
    for (size_t g = 0; g < text.size(); g++) { fa[g] = text[g].data(); nb[g] = text[g].size(); }
    std::vector<pmn_result *> res((size_t)n, nullptr);
    int rc = sched_run(s, (int)text.size(), fa.data(), nb.data(), nullptr, nullptr, names.data(), n, ref.data(), qry.data(), o, res.data());
    for (int i = 0; i < n && !rc; i++) { size_t len; const char *d = pmn_result_delta(res[(size_t)i], &len); rc = pmn_write_file_atomic(out_delta_paths[i], d, len); }
    for (pmn_result *r : res) pmn_result_free(r);
    return rc;
}

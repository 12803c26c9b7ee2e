// pmn_multi.cu — ONE process driving several GPUs of a box through the C ABI (no Python, no torch).
//
// The reference's concurrency knob is `-cores N` / run_size: one paramugsy process keeps N workers busy
// (/root/reference/lib/base/paramugsy.ml:54-57, lib/base/queued_task_server.ml:57-64), each an independent `mugsy_nucmer`
// process per pair (lib/base/job_processor.ml:128-154).  A pmn_multi is that over the GPUs of one box: one pmn_sched
// (W worker threads, one stream and scratch each) per device.
//
//   * all-vs-all batches (C2, C3, C5 of BASELINE.json): the pair list is cut by reference into one contiguous run per device
//     (pmn_multi_plan), every device packs the genomes and builds the indexes ITS pairs name.  Nothing crosses between the
//     devices — the path is embarrassingly parallel by pair.  A reference whose pairs straddle a cut is
//     indexed on both sides: rebuilding (0.7 ms per 5 Mbp, 12 ms per 100 Mbp, all devices at once) is cheaper than a
//     broadcast that the consumers would wait for (measured: DESIGN.md §7).
//   * one large pair (C4): every device packs both genomes and builds the index (or copies the owner's image over
//     NVLink, PMN_MULTI_INDEX=copy), seeds its range of query positions, the anchor lists are gathered on the first device
//     with peer copies (their concatenation in device order is the anchor list of the undivided run) and the first device
//     clusters, extends and writes the .delta — byte-identical for any number of devices.
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "pmn_scratch.cuh"

struct pmn_multi {
    std::vector<int> devices;
    std::vector<pmn_sched *> sched;
};

// Pairs in reference order, cut into n_devices contiguous runs of (nearly) equal cost — cost of a pair = bytes of its two
// genomes — so that a device needs as few distinct indexes as possible.  The same rule as paramugsy_b200/multi.py:assign_pairs
// (the torch.distributed form of this scheduler); deterministic.
extern "C" int pmn_multi_plan(int n_devices, int n_genomes, const size_t *bytes, int n_pairs, const int32_t *ref, const int32_t *qry, int32_t *device_of_pair)
{
    if (n_devices < 1 || n_pairs < 0 || (n_pairs && (!ref || !qry || !device_of_pair))) return pmn_set_error(PMN_E_ARG, "pmn_multi_plan: bad argument");
    for (int p = 0; p < n_pairs; p++)
        if (ref[p] < 0 || ref[p] >= n_genomes || qry[p] < 0 || qry[p] >= n_genomes) return pmn_set_error(PMN_E_ARG, "pmn_multi_plan: pair %d names a genome out of range", p);
    std::vector<int> order((size_t)n_pairs);
    for (int p = 0; p < n_pairs; p++) order[(size_t)p] = p;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ref[a] < ref[b]; });
    auto cost = [&](int p) -> double { return bytes ? (double)bytes[ref[p]] + (double)bytes[qry[p]] : 1.0; };
    double total = 0;
    for (int p = 0; p < n_pairs; p++) total += cost(p);
    double acc = 0;
    for (int k = 0; k < n_pairs; k++) {
        const int p = order[(size_t)k];
        const double w = cost(p);
        // the device whose interval [d * total / n, (d + 1) * total / n) holds the pair's midpoint
        int d = total > 0 ? (int)((acc + w / 2) * n_devices / total) : 0;
        device_of_pair[p] = std::min(n_devices - 1, std::max(0, d));
        acc += w;
    }
    return 0;
}

extern "C" int pmn_multi_create(const int *devices, int n_devices, int workers_per_device, pmn_multi **out)
{
    if (!out || n_devices < 1 || n_devices > 64 || workers_per_device < 1) return pmn_set_error(PMN_E_ARG, "pmn_multi_create: bad argument");
    *out = nullptr;
    pmn_multi *m = new pmn_multi();
    for (int k = 0; k < n_devices; k++) {
        const int dev = devices ? devices[k] : k;
        pmn_sched *s = nullptr;
        const int rc = pmn_sched_create(dev, workers_per_device, &s);
        if (rc) { for (pmn_sched *x : m->sched) pmn_sched_destroy(x); delete m; return rc; }
        m->devices.push_back(dev); m->sched.push_back(s);
    }
    // peer access between the devices, both ways, for the anchor gather and the index copy of a large pair (a pair that
    // cannot be enabled — or is the same device twice — falls back to a staged copy inside cudaMemcpyPeerAsync)
    for (int a : m->devices)
        for (int b : m->devices) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) { cudaGetLastError(); continue; }
            if (cudaSetDevice(a) == cudaSuccess) { cudaError_t e = cudaDeviceEnablePeerAccess(b, 0); if (e != cudaSuccess) cudaGetLastError(); }
        }
    *out = m;
    return 0;
}

extern "C" void pmn_multi_destroy(pmn_multi *m)
{
    if (!m) return;
    for (pmn_sched *s : m->sched) pmn_sched_destroy(s);
    delete m;
}

extern "C" int pmn_multi_devices(const pmn_multi *m) { return m ? (int)m->sched.size() : 0; }
extern "C" pmn_sched *pmn_multi_sched(const pmn_multi *m, int k) { return (m && k >= 0 && k < (int)m->sched.size()) ? m->sched[(size_t)k] : nullptr; }

extern "C" int pmn_multi_align_fasta(pmn_multi *m, int n_genomes, const char *const *fasta, const size_t *bytes, const char *const *names,
                                     int n_pairs, const int32_t *ref, const int32_t *qry, const pmn_opts *o, pmn_result **out)
{
    if (!m || n_genomes < 0 || n_pairs < 0 || (n_genomes && (!fasta || !bytes)) || (n_pairs && (!ref || !qry || !out))) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_fasta: bad argument");
    const int nd = (int)m->sched.size();
    std::vector<int32_t> dev((size_t)n_pairs);
    { const int rc = pmn_multi_plan(nd, n_genomes, bytes, n_pairs, ref, qry, dev.data()); if (rc) return rc; }
    for (int p = 0; p < n_pairs; p++) out[p] = nullptr;
    struct Part { std::vector<int32_t> ref, qry; std::vector<int> pair; std::vector<pmn_result *> res; int rc = 0; std::string err; };
    std::vector<Part> part((size_t)nd);
    for (int p = 0; p < n_pairs; p++) { Part &P = part[(size_t)dev[(size_t)p]]; P.ref.push_back(ref[p]); P.qry.push_back(qry[p]); P.pair.push_back(p); }
    auto run = [&](int d) {
        Part &P = part[(size_t)d];
        if (P.pair.empty()) return;
        P.res.assign(P.pair.size(), nullptr);
        // genomes are named by their index in the caller's list: the device's scheduler packs only those its pairs use
        P.rc = pmn_sched_align_fasta(m->sched[(size_t)d], n_genomes, fasta, bytes, names, (int)P.pair.size(), P.ref.data(), P.qry.data(), o, P.res.data());
        if (P.rc) P.err = pmn_last_error(nullptr);
    };
    std::vector<std::thread> th;
    for (int d = 1; d < nd; d++) th.emplace_back(run, d);
    run(0);
    for (auto &t : th) t.join();
    int rc = 0; std::string err;
    for (int d = 0; d < nd; d++) if (part[(size_t)d].rc && !rc) { rc = part[(size_t)d].rc; err = part[(size_t)d].err; }
    for (int d = 0; d < nd; d++)
        for (size_t k = 0; k < part[(size_t)d].res.size(); k++) {
            if (rc) pmn_result_free(part[(size_t)d].res[k]);          // any failing pair fails the batch (lib/base/job_processor.ml:72-73)
            else out[part[(size_t)d].pair[k]] = part[(size_t)d].res[k];
        }
    if (rc) return pmn_set_error(rc, "%s", err.c_str());
    return 0;
}

// File level, one call per Nucmer_task.t.searches (lib/base/nucmer_task.ml:6,48-59) over all devices: every distinct FASTA is
// read once, every .delta is written atomically (tmp + rename); with o->post and maf_outs the filtered delta and its MAF
// (what one mugsy_nucmer process leaves behind, lib/nucmer/mugsy_nucmer.ml:127-131).
extern "C" int pmn_multi_align_files(pmn_multi *m, int n, const char *const *ref_fasta_paths, const char *const *qry_fasta_paths,
                                     const char *const *out_delta_paths, const char *const *maf_outs, const pmn_opts *o)
{
    if (!m || n < 0 || (n && (!ref_fasta_paths || !qry_fasta_paths || !out_delta_paths))) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_files: bad argument");
    if (maf_outs && !(o && o->post)) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_files: MAF output needs pmn_opts.post = 1 or 2");
    std::map<std::string, int> id;
    std::vector<std::string> text, keep_names; std::vector<int32_t> ref((size_t)n), qry((size_t)n);
    auto genome = [&](const char *path, int32_t *g) -> int {
        if (!path) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_files: NULL path");
        auto it = id.find(path);
        if (it == id.end()) {
            std::string t; int rc = pmn_read_file(path, t); if (rc) return rc;
            it = id.emplace(path, (int)text.size()).first; text.push_back(std::move(t)); keep_names.push_back(path);
        }
        *g = it->second; return 0;
    };
    for (int i = 0; i < n; i++) {
        int rc = genome(ref_fasta_paths[i], &ref[(size_t)i]); if (rc) return rc;
        rc = genome(qry_fasta_paths[i], &qry[(size_t)i]); if (rc) return rc;
        if (!out_delta_paths[i]) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_files: NULL output path");
    }
    std::vector<const char *> fa(text.size()), names(text.size()); std::vector<size_t> nb(text.size());
    for (size_t g = 0; g < text.size(); g++) { fa[g] = text[g].data(); nb[g] = text[g].size(); names[g] = keep_names[g].c_str(); }
    std::vector<pmn_result *> res((size_t)n, nullptr);
    int rc = pmn_multi_align_fasta(m, (int)text.size(), fa.data(), nb.data(), names.data(), n, ref.data(), qry.data(), o, res.data());
    for (int i = 0; i < n && !rc; i++) {
        size_t len; const char *d;
        if (maf_outs) {
            d = pmn_result_filtered(res[(size_t)i], &len); rc = pmn_write_file_atomic(out_delta_paths[i], d, len);
            if (!rc && maf_outs[i]) { d = pmn_result_maf(res[(size_t)i], &len); rc = pmn_write_file_atomic(maf_outs[i], d, len); }
        } else { d = pmn_result_delta(res[(size_t)i], &len); rc = pmn_write_file_atomic(out_delta_paths[i], d, len); }
    }
    for (pmn_result *r : res) pmn_result_free(r);
    return rc;
}

// One large pair over all devices.  stats_ms (may be NULL) receives the wall clock of the phases in
// milliseconds: [0] pack + index on every device, [1] seeding of the parts, [2] gather, [3] clustering + extension + text.
extern "C" int pmn_multi_align_large(pmn_multi *m, const char *ref_fasta, size_t ref_bytes, const char *qry_fasta, size_t qry_bytes, const pmn_opts *o,
                                     const char *ref_path, const char *qry_path, pmn_result **out, double *stats_ms)
{
    if (!m || !ref_fasta || !qry_fasta || !out) return pmn_set_error(PMN_E_ARG, "pmn_multi_align_large: NULL argument");
    *out = nullptr;
    const int nd = (int)m->sched.size();
    struct Dev { pmn_ctx *c = nullptr; pmn_seq *r = nullptr, *q = nullptr; pmn_index *ix = nullptr; void *anc = nullptr; int64_t n = 0; int rc = 0; std::string err; };
    std::vector<Dev> D((size_t)nd);
    for (int d = 0; d < nd; d++) D[(size_t)d].c = pmn_sched_ctx(m->sched[(size_t)d], 0);
    const char *mi = getenv("PMN_MULTI_INDEX");
    const bool copy_index = mi && !strcmp(mi, "copy");
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto each = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int d = 1; d < nd; d++) th.emplace_back([&, d] { if (!D[(size_t)d].rc) { D[(size_t)d].rc = fn(d); if (D[(size_t)d].rc) D[(size_t)d].err = pmn_last_error(nullptr); } });
        if (!D[0].rc) { D[0].rc = fn(0); if (D[0].rc) D[0].err = pmn_last_error(nullptr); }
        for (auto &t : th) t.join();
    };
    const double t0 = now();
    // ---- pack both genomes everywhere; index: built on every device at once, or built by the first and copied
    each([&](int d) -> int {
        Dev &X = D[(size_t)d];
        int rc = pmn_seq_from_fasta(X.c, ref_fasta, ref_bytes, &X.r); if (rc) return rc;
        rc = pmn_seq_from_fasta(X.c, qry_fasta, qry_bytes, &X.q); if (rc) return rc;
        if (!copy_index || d == 0) return pmn_index_build(X.c, X.r, &X.ix);
        return pmn_index_alloc(X.c, X.r, &X.ix);
    });
    auto failed = [&] { for (auto &X : D) if (X.rc) return true; return false; };
    if (copy_index && !failed())
        each([&](int d) -> int {
            if (d == 0) return 0;
            Dev &X = D[(size_t)d];
            void *src = nullptr, *dst = nullptr; size_t nb = 0, nb2 = 0;
            pmn_index_image(D[0].ix, &src, &nb); pmn_index_image(X.ix, &dst, &nb2);
            if (nb != nb2) return pmn_set_error(PMN_E_INTERNAL, "pmn_multi_align_large: index images differ in size");
            PMN_CUDA_OK(cudaSetDevice(X.c->device));
            PMN_CUDA_OK(cudaMemcpyPeerAsync(dst, X.c->device, src, D[0].c->device, nb, X.c->stream));
            PMN_CUDA_OK(cudaStreamSynchronize(X.c->stream));
            return pmn_index_adopt(X.ix);
        });
    const double t1 = now();
    // ---- every device seeds its range of query positions
    if (!failed()) each([&](int d) -> int { Dev &X = D[(size_t)d]; return pmn_seed_part(X.c, X.ix, X.q, o, d, nd, &X.anc, &X.n); });
    const double t2 = now();
    int rc = 0; std::string err;
    for (auto &X : D) if (X.rc && !rc) { rc = X.rc; err = X.err; }
    // ---- gather on the first device, in device order
    DevBuf all;
    int64_t total = 0;
    if (!rc) {
        for (auto &X : D) total += X.n;
        cudaSetDevice(D[0].c->device);
        pmn_tls_stream = D[0].c->stream;
        if (total > 0 && all.ensure(16 * (size_t)total)) rc = PMN_E_NOMEM;
        int64_t at = 0;
        for (int d = 0; d < nd && !rc; d++) {
            Dev &X = D[(size_t)d];
            if (X.n > 0) {
                cudaError_t e = d == 0 ? cudaMemcpyAsync((char *)all.p + 16 * at, X.anc, 16 * (size_t)X.n, cudaMemcpyDeviceToDevice, D[0].c->stream)
                                       : cudaMemcpyPeerAsync((char *)all.p + 16 * at, D[0].c->device, X.anc, X.c->device, 16 * (size_t)X.n, D[0].c->stream);
                if (e != cudaSuccess) { rc = pmn_set_error(PMN_E_CUDA, "pmn_multi_align_large: anchor gather: %s", cudaGetErrorString(e)); err = pmn_last_error(nullptr); }
            }
            at += X.n;
        }
        if (!rc && cudaStreamSynchronize(D[0].c->stream) != cudaSuccess) { rc = pmn_set_error(PMN_E_CUDA, "pmn_multi_align_large: anchor gather failed"); err = pmn_last_error(nullptr); }
    }
    const double t3 = now();
    // ---- clustering, extension, .delta on the first device
    if (!rc) {
        rc = pmn_align_anchors(D[0].c, D[0].ix, D[0].q, o, total > 0 ? all.p : nullptr, total, ref_path, qry_path, out);
        if (rc) err = pmn_last_error(nullptr);
    }
    const double t4 = now();
    if (all.p) { cudaSetDevice(D[0].c->device); cudaStreamSynchronize(D[0].c->stream); all.release(); }
    for (auto &X : D) { pmn_index_free(X.ix); pmn_seq_free(X.q); pmn_seq_free(X.r); }
    if (stats_ms) { stats_ms[0] = t1 - t0; stats_ms[1] = t2 - t1; stats_ms[2] = t3 - t2; stats_ms[3] = t4 - t3; }
    if (rc) return pmn_set_error(rc, "%s", err.c_str());
    return 0;
}

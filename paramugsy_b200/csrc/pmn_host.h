// pmn_host.h — host-side objects behind the C ABI of include/pmnucmer.h.
#pragma once
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/pmnucmer.h"
#include "pmn_common.cuh"

struct pmn_seq {
    pmn_ctx *ctx = nullptr;
    int nrec = 0;
    std::vector<std::string> ids;
    std::vector<int64_t> len, off;      // per record: bases, 0-based offset in the concatenation
    int64_t n = 0;                      // concatenated bases incl. separators
    int has_x = 0;
    int64_t nwords = 0;                 // words in each packed array (incl. padding)
    DevBuf w_fwd, xm_fwd;               // forward text
    DevBuf w_rev, xm_rev;               // reverse complement of the whole concatenation
    DevBuf residues;                    // the concatenation as characters, one byte per base (forward strand; delta2maf prints these)
    PackedView fwd() const { return PackedView{ w_fwd.as<uint64_t>(), xm_fwd.as<uint32_t>(), n, has_x }; }
    PackedView rev() const { return PackedView{ w_rev.as<uint64_t>(), xm_rev.as<uint32_t>(), n, has_x }; }
};

struct pmn_index {
    pmn_ctx *ctx = nullptr;
    const pmn_seq *seq = nullptr;       // borrowed: the caller keeps the sequence alive
    int64_t n = 0;
    // one contiguous image in HBM (so that it can be replicated by a single NCCL broadcast):
    //   [0,256) header {magic, n, K, rounds} | int32 SA[n] | int32 LCP[n] | uint32 table[4^K + 1] | uint8 skip[n + 1] | uint32 present[4^P / 32],
    //   each 256-byte aligned
    DevBuf blob;
    size_t off_sa = 0, off_lcp = 0, off_table = 0, off_skip = 0, off_present = 0, blob_bytes = 0;
    uint32_t *sa() const { return (uint32_t *)((char *)blob.p + off_sa); }
    int32_t *lcp() const { return (int32_t *)((char *)blob.p + off_lcp); }
    uint32_t *table() const { return (uint32_t *)((char *)blob.p + off_table); }
    // skip[E] for a reference coordinate E (one past the end of a match): E - skip[E] is the first reference position p whose
    // suffix shares E - p or more bases with another suffix (255: further back than 254).  A query position whose match ends at E
    // and starts before that p is matched nowhere else as long: the seeding kernel steps over those positions (pmn_seed.cu).
    uint8_t *skip() const { return (uint8_t *)blob.p + off_skip; }
    // present: one bit per P-mer, set for (at least) every P-mer that occurs in the reference.  A match of minmatch bases that starts
    // at any of minmatch - P + 1 consecutive query positions contains the P-mer at the last of them: when that P-mer is not in
    // the reference none of those positions has an anchor, and the seeding kernel settles them with one bit probe.
    uint32_t *present() const { return (uint32_t *)((char *)blob.p + off_present); }
    int P = 0;
    int K = 0;
    int rounds = 0;                     // prefix-doubling rounds after the 16-mer pass
    float ms_build = 0, wall_ms_build = 0;
};

struct StageTimes { float pack = 0, index = 0, seed = 0, cluster = 0, extend = 0, total = 0; };

struct pmn_result {
    std::string delta;
    std::string filtered;                       // pmn_opts.post: `delta-filter` of delta
    char *maf = nullptr; size_t maf_len = 0;    // ... and `delta2maf` of that, in a buffer of the pinned-host pool (the device writes it directly)
    ~pmn_result();
    pmn_stats stats{};
    // stage dumps kept for parity tests
    std::vector<int32_t> anchors;             // n x 4
    std::vector<int32_t> cl_matches, cl_off, cl_tag;
    std::vector<int64_t> al_rows, al_doff;      // 10 columns per alignment; delta offsets (a+1)
    std::vector<int32_t> al_deltas;
};

struct DevPool {
    std::mutex mu;
    std::vector<DevBuf> bufs;
    ~DevPool() { for (auto &b : bufs) b.release(); }
};

// per-context scratch, all grow-only
struct Scratch;

struct pmn_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // forked from `stream` where two independent kernels of one pair can run side by side
    cudaStream_t prio_stream[8] = {};   // created on demand: streams of higher device priority than `stream` (level 0 = highest), see pmn_ctx_prio_stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev[16] = {};
    PmnError err{};
    Scratch *scratch = nullptr;
    long launches = 0;                  // kernels launched by this library (bench.py's gpu_launches)
    long syncs = 0;                     // host waits for the device (cudaStreamSynchronize) inside the stages
    int64_t h2d_bytes = 0, d2h_bytes = 0, pairs = 0;
    std::shared_ptr<DevPool> pool;      // device buffers handed back by freed sequences / indexes, reused by the next ones;
                                        // the worker contexts of one pmn_sched share one pool (a genome packed or an index
                                        // built by one worker is freed by whichever worker finishes its last pair)
    bool smem_attr_set = false;         // opt-in dynamic shared memory of the extension kernels (per device)
    std::vector<char> text_buf;         // grow-only staging of the .delta formatter (worst-case size, touched once)
    // largest match -> next match window (cells) the thread-per-job kernel of the extension takes; larger ones are run by one warp
    // each (eng_mid_full).  4096 is the shortest pair (the largest windows are the tail of the thread-per-job kernel: 0.57 -> 0.28 ms
    // of a 5 Mbp pair) but 5 % more instructions per pair; a scheduler with many pairs in flight is bound by the instructions
    // issued and asks for 10000 (pmn_sched.cu).  The results do not depend on it.
    int tpj_cells = 4096;
};

// host<->device copies on the context's stream, counted for bench.py's e2e byte figures
#define PMN_H2D(c, dst, src, bytes) do { (c)->h2d_bytes += (int64_t)(bytes); PMN_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (c)->stream)); } while (0)
#define PMN_D2H(c, dst, src, bytes) do { (c)->d2h_bytes += (int64_t)(bytes); PMN_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (c)->stream)); } while (0)

// take a buffer of at least `bytes` from the context's pool (smallest that fits), else allocate
int pmn_pool_get(pmn_ctx *c, DevBuf &b, size_t bytes);
void pmn_pool_put(pmn_ctx *c, DevBuf &b);

// ---- decimal formatting of the .delta writers (pmn_api.cu, pmn_post.cu)
static const char PMN_DIGITS2[201] =
    "00010203040506070809101112131415161718192021222324252627282930313233343536373839404142434445464748495051525354555657585960616263646566676869707172737475767778798081828384858687888990919293949596979899";
// Most numbers of a .delta are deltas of one to three digits with a random sign: below 10000 there is no data-dependent
// branch (sign by pointer arithmetic, four digits from a table, the copy starts at the first significant one; it writes up to
// three bytes past the number, the callers' buffers have that slack).
static inline char *pmn_fmt_int(char *p, long long v)
{
    const bool neg = v < 0;
    *p = '-'; p += neg;
    const unsigned long long w0 = neg ? 0ull - (unsigned long long)v : (unsigned long long)v;
    if (w0 < 10000) {
        const unsigned u = (unsigned)w0, hi = u / 100, lo = u % 100;
        uint16_t a, b;
        memcpy(&a, PMN_DIGITS2 + 2 * hi, 2); memcpy(&b, PMN_DIGITS2 + 2 * lo, 2);
        const int n = 1 + (u >= 10) + (u >= 100) + (u >= 1000);
        const uint32_t x = ((uint32_t)a | (uint32_t)b << 16) >> (8 * (4 - n));      // little-endian host: the four digits in memory order, leading zeros shifted out
        memcpy(p, &x, 4);
        return p + n;
    }
    char tmp[24]; int n = 0;
    unsigned long long w = w0;
    while (w >= 100) { const unsigned lo = (unsigned)(w % 100); w /= 100; tmp[n++] = PMN_DIGITS2[2 * lo + 1]; tmp[n++] = PMN_DIGITS2[2 * lo]; }
    if (w >= 10) { tmp[n++] = PMN_DIGITS2[2 * w + 1]; tmp[n++] = PMN_DIGITS2[2 * w]; } else tmp[n++] = (char)('0' + w);
    while (n) *p++ = tmp[--n];
    return p;
}

int pmn_last_code();                          // code of the calling thread's last pmn_set_error (pmn_api.cu)
void pmn_apply_device_sched(int workers);
// A stream of this context whose kernels the device schedules before those of the ordinary streams (level 0 first, then 1, ...;
// as many levels as the device has above the default priority).  The scheduler builds the indexes of a batch on them, in
// reference order: pairs wait for their index, so its kernels go first, and the index of the first reference — the one most
// pairs are waiting for — is not slowed down by the six builds that started with it.  nullptr: no priorities on this device.
cudaStream_t pmn_ctx_prio_stream(pmn_ctx *c, int level);     // whether host threads spin or yield while they wait for the device (pmn_api.cu)

// stage entry points (defined in the .cu files)
int pmn_fasta_to_device(pmn_ctx *c, pmn_seq *s, const char *txt, size_t nb, const std::vector<int64_t> &header_pos);
int pmn_index_build_impl(pmn_ctx *c, const pmn_seq *ref, pmn_index *ix);
int pmn_index_layout(pmn_ctx *c, const pmn_seq *ref, pmn_index *ix);   // sizes the image for ref and takes it from the pool
struct PmnIndexHeader { uint64_t magic; int64_t n; int32_t K, rounds; };
#define PMN_INDEX_MAGIC 0x33584449304e4d50ull   /* "PMN0IDX3" */
int pmn_seed_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *q, const pmn_opts *o, int64_t *n_anchors, int part = 0, int nparts = 1);
int pmn_cluster_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *q, const pmn_opts *o, int64_t n_anchors);
int pmn_extend_impl(pmn_ctx *c, const pmn_index *ix, const pmn_seq *q, const pmn_opts *o, pmn_result *res);
// MAF texts are some 10 MB per pair: they are copied from the device straight into pinned host buffers that are recycled
// (pmn_free_text / pmn_result_free hand them back), instead of a bounce buffer plus a 10 MB memcpy per pair
char *pmn_pinned_get(size_t bytes);
bool pmn_pinned_put(void *p);          // false: p is not a buffer of the pool
int pmn_post_impl(pmn_ctx *c, const pmn_seq *ref, const pmn_seq *qry, const char *ref_path, const char *qry_path, int mode, pmn_result *r);
int pmn_read_file(const char *path, std::string &out);
int pmn_write_file_atomic(const char *path, const char *data, size_t len);
Scratch *pmn_scratch_new();
void pmn_scratch_free(Scratch *s);

// pmn_scratch.cuh — every grow-only device buffer a context owns.  One pair is in flight
// per context at a time, so stages simply reuse these between pairs (no cudaMalloc at
// steady state).
#pragma once
#include "pmn_host.h"
#include "pmn_prims.cuh"

struct Scratch {
    RadixScratch rs;
    DevBuf k0, k1, v0, v1;            // radix sort ping-pong (uint64 keys, uint32 values)
    DevBuf scan_tmp;                  // spine of pmn_scan
    DevBuf codes;                     // uint8 staging of parsed FASTA
    // index build
    DevBuf gs, rank, flags, list0, list1, gsn;
    // seeding
    DevBuf sections, stage, tile_cnt, tile_off, seed_bits, anchors;   // anchors: int4 (r, q, len, tag)
    // clustering
    DevBuf cl_a, cl_b, cl_c, cl_d, cl_e, cl_f, cl_g, cl_h, cl_i, cl_j, cl_k, cl_l;
    DevBuf cl_matches, cl_recs, cl_counters;
    // extension
    DevBuf ex_a, ex_b, ex_c, ex_d, ex_e, ex_f, ex_g, ex_h, ex_i, ex_j, ex_k, ex_l;
    DevBuf ex_scores, ex_tb, ex_tbidx, ex_pool, ex_counters, ex_arena, ex_dbg, ex_desc, ex_tkey, ex_tscratch;
    // every device buffer above, in a fixed order (the scheduler levels capacities between its workers; pmn_scratch_free)
    std::vector<DevBuf *> all()
    {
        return { &rs.hist, &rs.spine, &k0, &k1, &v0, &v1, &scan_tmp, &codes, &gs, &rank, &flags, &list0, &list1, &gsn, &sections, &stage, &tile_cnt, &tile_off, &seed_bits, &anchors,
                 &cl_a, &cl_b, &cl_c, &cl_d, &cl_e, &cl_f, &cl_g, &cl_h, &cl_i, &cl_j, &cl_k, &cl_l, &cl_matches, &cl_recs, &cl_counters,
                 &ex_a, &ex_b, &ex_c, &ex_d, &ex_e, &ex_f, &ex_g, &ex_h, &ex_i, &ex_j, &ex_k, &ex_l,
                 &ex_scores, &ex_tb, &ex_tbidx, &ex_pool, &ex_counters, &ex_arena, &ex_dbg, &ex_desc, &ex_tkey, &ex_tscratch };
    }
    // the diagonal-difference limits of the clustering stage as they are in cl_g, and the options they were computed from
    std::vector<int32_t> lim_host; int lim_maxgap = -1, lim_diagdiff = -1; double lim_diagfactor = -1.0;
    // host-side counts handed from one stage to the next (one pair in flight per context)
    int64_t n_anchors = 0, n_clusters = 0, n_cl_matches = 0, seed_lookups = 0, seed_probes = 0;
    // traceback arena of the extension: sized from what pairs have needed so far (grown and the extension re-run when a pair needs more)
    size_t arena_cap = 0;
    bool arena_retry = false;          // set by the extension when it grew the arena after running out: run it again
    // pinned host staging
    void *pinned = nullptr; size_t pinned_cap = 0;
    int ensure_pinned(size_t bytes)
    {
        if (bytes <= pinned_cap) return 0;
        if (pinned) cudaFreeHost(pinned);
        // a power of two with as much again on top, 1 MB at least: results of the pairs a worker meets differ by tens of per cent,
        // and every growth is a cudaFreeHost + cudaMallocHost that stalls the whole device
        size_t want = (size_t)1 << 20;
        while (want < 2 * bytes) want <<= 1;
        if (cudaMallocHost(&pinned, want) != cudaSuccess) { pinned = nullptr; pinned_cap = 0; return pmn_set_error(-3, "cudaMallocHost(%zu) failed", want); }
        pinned_cap = want;
        return 0;
    }
};

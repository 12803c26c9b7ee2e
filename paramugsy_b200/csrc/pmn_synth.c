/*
 * pmn_synth.c — deterministic synthetic genomes for the benchmark configs of
 * BASELINE.json (SURVEY.md §8d "Synthetic inputs").  Plain C, no CUDA: built as
 * libpmn_synth.so so that tests, bench.py and the oracle checker all see the same
 * bytes.  Output is upper-case ACGT text without header or line breaks.
 *
 * PRNG: splitmix64, one stream per genome.
 * Mutation model mutate(g, d, seed): per base u ~ U[0,1):
 *     u < 0.8 d  substitute by one of the 3 other bases (uniform)
 *     u < 0.9 d  delete the base
 *     u < 1.0 d  insert before it a run of 1+Geom(1/2) (capped at 10) random bases
 * so identity ~ 1-d.  invert(g, k, len, seed): reverse-complement k non-overlapping
 * segments of `len` bases at PRNG-chosen offsets.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } pmn_rng;

static inline uint64_t rng_next(pmn_rng *r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

static inline double rng_unit(pmn_rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }

static const char BASES[4] = { 'A', 'C', 'G', 'T' };

static inline int base_code(char c)
{
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; default: return 3; }
}

/* out must hold n bytes */
void pmn_synth_random(char *out, int64_t n, uint64_t seed)
{
    pmn_rng r = { seed };
    for (int64_t i = 0; i < n; i++) out[i] = BASES[rng_next(&r) >> 62];
}

/* out must hold 2*n + 16 bytes; returns the mutated length */
int64_t pmn_synth_mutate(const char *g, int64_t n, double d, uint64_t seed, char *out)
{
    pmn_rng r = { seed };
    int64_t o = 0;
    const double t_sub = 0.8 * d, t_del = 0.9 * d, t_ins = d;
    for (int64_t i = 0; i < n; i++) {
        double u = rng_unit(&r);
        if (u < t_sub) {
            int b = base_code(g[i]);
            out[o++] = BASES[(b + 1 + (int)(rng_next(&r) % 3)) & 3];
        } else if (u < t_del) {
            /* base dropped */
        } else if (u < t_ins) {
            int run = 1;
            while (run < 10 && (rng_next(&r) >> 63)) run++;
            for (int k = 0; k < run; k++) out[o++] = BASES[rng_next(&r) >> 62];
            out[o++] = g[i];
        } else {
            out[o++] = g[i];
        }
    }
    return o;
}

static inline char comp(char c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

/* in place; returns the number of segments actually inverted (k unless the genome is too small) */
int pmn_synth_invert(char *g, int64_t n, int k, int64_t len, uint64_t seed)
{
    pmn_rng r = { seed };
    if (len <= 0 || len > n) return 0;
    int64_t *starts = (int64_t *)malloc(sizeof(int64_t) * (size_t)(k > 0 ? k : 1));
    int done = 0;
    for (int tries = 0; done < k && tries < 1000 * k; tries++) {
        int64_t s = (int64_t)(rng_next(&r) % (uint64_t)(n - len + 1));
        int ok = 1;
        for (int j = 0; j < done; j++)
            if (s < starts[j] + len && starts[j] < s + len) { ok = 0; break; }
        if (!ok) continue;
        starts[done++] = s;
        for (int64_t a = s, b = s + len - 1; a <= b; a++, b--) {
            char ca = comp(g[a]), cb = comp(g[b]);
            g[a] = cb; g[b] = ca;
        }
    }
    free(starts);
    return done;
}

"""Named small inputs shared by the golden fixtures, the oracle tests and the GPU parity tests.
Every case is (reference FASTA bytes, query FASTA bytes, nucmer option overrides)."""
from paramugsy_b200 import synth


def _pair(n, d, seed, inv=0, inv_len=0):
    g0 = synth.random_genome(n, seed)
    g1 = synth.mutate(g0, d, seed + 1)
    if inv:
        g1 = synth.invert(g1, inv, inv_len, seed + 2)
    return synth.fasta("ref.1", g0), synth.fasta("qry.1", g1)


def c_1k_99():
    return (*_pair(1_000, 0.01, 101), {})


def c_10k_95():
    return (*_pair(10_000, 0.05, 111, inv=1, inv_len=1_500), {})


def c_100k_98_inv():
    return (*_pair(100_000, 0.02, 121, inv=2, inv_len=5_000), {})


def c_100k_90():
    return (*_pair(100_000, 0.10, 131, inv=1, inv_len=8_000), {})


def c_60k_85():
    return (*_pair(60_000, 0.15, 141), {})


def c_identical():
    g = synth.random_genome(5_000, 151)
    return synth.fasta("ref.1", g), synth.fasta("qry.1", g), {}


def c_unrelated():
    return synth.fasta("ref.1", synth.random_genome(4_000, 161)), synth.fasta("qry.1", synth.random_genome(4_000, 162)), {}


def c_short_and_empty():
    # records shorter than minmatch, an empty record, lower case, a trailing record without newline
    return (b">r.1\nACGTACGTAC\n>r.2\n\n>r.3\n" + synth.random_genome(300, 171).lower() + b"\n",
            b">q.1\nACGT\n>q.2\n" + synth.random_genome(300, 171) + b"\n>q.3\nAC", {})


def c_n_runs():
    a = synth.random_genome(3_000, 181); b = synth.random_genome(2_000, 182)
    ref = synth.fasta("ref.1", a[:1500] + b"N" * 40 + a[1500:] + b"NNNNNNNNNN" + b)
    qa = synth.mutate(a, 0.02, 183); qb = synth.mutate(b, 0.03, 184)
    qry = synth.fasta("qry.1", qb[:900] + b"RYKM" + qb[900:] + b"N" * 25 + qa)
    return ref, qry, {}


def c_multirecord():
    u = synth.random_genome(6_000, 191); v = synth.random_genome(5_000, 192); w = synth.random_genome(800, 193)
    ref = synth.fasta("r.1", u) + synth.fasta("r.2", v) + synth.fasta("r.3", w)
    qry = (synth.fasta("q.1", synth.mutate(v, 0.03, 194)) + synth.fasta("q.2", synth.random_genome(500, 195)) +
           synth.fasta("q.3", synth.invert(synth.mutate(u, 0.02, 196), 1, 900, 197) + synth.mutate(w, 0.01, 198)))
    return ref, qry, {}


def c_repeats():
    unit = synth.random_genome(700, 201)
    rep = unit * 4 + synth.random_genome(3_000, 202) + unit[:350]
    return synth.fasta("ref.1", rep), synth.fasta("qry.1", synth.mutate(rep, 0.02, 203)), {}


def c_tandem_low_complexity():
    # low-complexity stretches make overlapping same-start matches: exercises Filter_Matches
    body = synth.random_genome(2_000, 211)
    ref = body[:1000] + b"AC" * 60 + body[1000:] + b"A" * 45 + synth.random_genome(800, 212)
    qry = body[:1000] + b"AC" * 52 + body[1000:] + b"A" * 38 + synth.random_genome(800, 212)
    return synth.fasta("ref.1", ref), synth.fasta("qry.1", synth.mutate(qry, 0.01, 213)), {}


def c_big_indels():
    g = synth.random_genome(30_000, 221)
    q = g[:8_000] + g[8_150:15_000] + synth.random_genome(260, 222) + g[15_000:22_000] + g[22_900:]
    return synth.fasta("ref.1", g), synth.fasta("qry.1", synth.mutate(q, 0.03, 223)), {}


def c_opts_l12_c30():
    return (*_pair(20_000, 0.12, 231), {"minmatch": 12, "mincluster": 30, "maxgap": 120, "breaklen": 100})


def c_opts_l6():
    # minmatch below the K of the k-mer table (K >= 8): no bucket lookup, every position takes the full binary search
    return (*_pair(2_500, 0.04, 241), {"minmatch": 6, "mincluster": 20})


def c_x_runs_30k():
    # a text with N runs, IUPAC codes and three records above one sort tile: 64-bit keys with class bits in the tiled sort,
    # LCP entries next to X-limited suffixes, one-suffix buckets that end on an X
    a = synth.random_genome(14_000, 251); b = synth.random_genome(9_000, 252); c = synth.random_genome(6_000, 253)
    ref = synth.fasta("r.1", a[:5_000] + b"N" * 300 + a[5_000:9_000] + b"RYSWKM" + a[9_000:]) + synth.fasta("r.2", b + b"N" * 17 + c[:2_000]) + synth.fasta("r.3", c)
    qa = synth.mutate(a, 0.03, 254); qb = synth.mutate(b, 0.02, 255); qc = synth.mutate(c, 0.05, 256)
    qry = synth.fasta("q.1", qb[:4_000] + b"NNNN" + qb[4_000:]) + synth.fasta("q.2", synth.invert(qa, 1, 2_000, 257) + b"N" * 50 + qc)
    return ref, qry, {}


def c_opts_forward_only_noextend():
    return (*_pair(20_000, 0.04, 241, inv=1, inv_len=2_000), {"do_reverse": 0, "do_extend": 0})


def c_opts_nosimplify_diag():
    return (*_pair(20_000, 0.06, 251, inv=1, inv_len=2_000), {"do_simplify": 0, "diagdiff": 2, "diagfactor": 0.3})


def c_self():
    # aligning a sequence with repeats to itself
    unit = synth.random_genome(400, 261)
    g = synth.random_genome(2_000, 262) + unit + synth.random_genome(2_500, 263) + unit + synth.random_genome(1_000, 264)
    return synth.fasta("ref.1", g), synth.fasta("qry.1", g), {"do_simplify": 0}


def c_trim_everything():
    # homopolymer tails that share nothing: extensions out of the last / first cluster run into
    # A-against-C, every cell falls below the trimming threshold before the break length is reached
    g = synth.random_genome(3_000, 271)
    return (synth.fasta("ref.1", b"A" * 1_500 + g + b"A" * 1_500), synth.fasta("qry.1", b"C" * 1_500 + synth.mutate(g, 0.02, 272) + b"C" * 1_500), {})


def c_wide_band_b500():
    # a long break length keeps bands several hundred cells wide: register layouts up to 8 columns
    # per lane and the shared/global-memory fallback
    return (*_pair(30_000, 0.13, 281, inv=1, inv_len=3_000), {"breaklen": 500})


def c_gap_ladder():
    # gaps of growing size between exact blocks: band widths cross every layout boundary, both ways
    blocks = [synth.random_genome(120, 300 + k) for k in range(40)]
    junk_r = [synth.random_genome(4 * k, 400 + k) for k in range(40)]
    junk_q = [synth.random_genome(3 * k + 1, 500 + k) for k in range(40)]
    ref = b"".join(b + j for b, j in zip(blocks, junk_r)); qry = b"".join(b + j for b, j in zip(blocks, junk_q))
    return synth.fasta("ref.1", ref), synth.fasta("qry.1", qry), {}


CASES = {f.__name__[2:]: f for f in (
    c_1k_99, c_10k_95, c_100k_98_inv, c_100k_90, c_60k_85, c_identical, c_unrelated, c_short_and_empty, c_n_runs,
    c_multirecord, c_repeats, c_tandem_low_complexity, c_big_indels, c_opts_l12_c30, c_opts_l6, c_x_runs_30k, c_opts_forward_only_noextend,
    c_opts_nosimplify_diag, c_self, c_trim_everything, c_wide_band_b500, c_gap_ladder)}

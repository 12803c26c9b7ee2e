"""Stage-by-stage parity of the CUDA path against the oracle on a handful of inputs.
Development aid (the real parity tests are tests/test_gpu_parity.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pmn_oracle as O
from paramugsy_b200 import synth, lib as P

stages = sys.argv[1] if len(sys.argv) > 1 else "index,seed"
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None

def cases():
    g0 = synth.random_genome(3000, 11); yield "tiny", synth.fasta("a.1", g0), synth.fasta("b.1", synth.mutate(g0, 0.03, 12)), {}
    g0 = synth.random_genome(200000, 21); g1 = synth.invert(synth.mutate(g0, 0.03, 22), 2, 5000, 23)
    yield "200k", synth.fasta("a.1", g0), synth.fasta("b.1", g1), {}
    unit = synth.random_genome(300, 31); other = synth.random_genome(4000, 32)
    ref = synth.fasta("r.1", unit + b"NNNNN" + other + unit[:100]) + synth.fasta("r.2", unit[:150] + synth.random_genome(2000, 33)) + synth.fasta("r.3", b"ACGTACGTAC")
    qry = synth.fasta("q.1", other[50:3500] + b"N" + unit) + synth.fasta("q.2", synth.random_genome(100, 34) + other[:900]) + synth.fasta("q.3", b"ACGT")
    yield "multi", ref, qry, {"minmatch": 15}
    rep = synth.random_genome(700, 41) * 5 + synth.random_genome(3000, 42)
    yield "repeat", synth.fasta("a.1", rep), synth.fasta("b.1", synth.mutate(rep, 0.02, 43)), {}
    gs = synth.config_c1(); yield "C1", synth.fasta(*gs[0]), synth.fasta(*gs[1]), {}

ctx = P.Context(0)
bad = 0
for name, ref, qry, kw in cases():
    if only and name not in only: continue
    print('case', name, flush=True)
    t = time.time(); orc = O.Run(ref, qry, fast_chain=1, **kw)
    osa, olcp = orc.index(); oa = orc.anchors(); t_or = time.time() - t
    t = time.time(); rs = ctx.sequence(ref); qs = ctx.sequence(qry); t_pack = time.time() - t
    t = time.time(); ix = rs.index(); t_ix = time.time() - t
    sa, lcp = ix.suffix_array()
    ok_sa = np.array_equal(sa, osa); ok_lcp = np.array_equal(lcp, olcp)
    res = ix.align(qs, keep_stages=1, **kw)
    a = res.anchors(); st = res.stats
    ok_a = a.shape == oa.shape and np.array_equal(a, oa)
    print(f"{name}: n={len(sa)} SA={'ok' if ok_sa else 'DIFF'} LCP={'ok' if ok_lcp else 'DIFF'} anchors={len(a)}/{len(oa)} {'ok' if ok_a else 'DIFF'} "
          f"| oracle {t_or:.2f}s pack {t_pack*1e3:.1f}ms index {t_ix*1e3:.1f}ms (gpu {st['ms_index']:.2f}ms, rounds {st['sa_rounds']}, K {st['kmer_bits']//2}) seed {st['ms_seed']:.2f}ms")
    if not ok_sa:
        d = np.nonzero(sa != osa)[0]; print("  first SA diffs at", d[:10], sa[d[:5]], osa[d[:5]], "count", len(d))
    if not ok_lcp:
        d = np.nonzero(lcp != olcp)[0]; print("  first LCP diffs at", d[:10], lcp[d[:5]], olcp[d[:5]], "count", len(d))
    if not ok_a:
        sa_ = set(map(tuple, a.tolist())); so = set(map(tuple, oa.tolist()))
        print("  only gpu:", sorted(sa_ - so)[:8], "only oracle:", sorted(so - sa_)[:8], "order-only:", sa_ == so)
    bad += (not ok_sa) + (not ok_lcp) + (not ok_a)
    if "cluster" in stages:
        om, ooff, otag = orc.clusters(); m, off, tag = res.clusters()
        ok = np.array_equal(om, m) and np.array_equal(ooff, off) and np.array_equal(otag, tag)
        print(f"   clusters {len(tag)}/{len(otag)} matches {len(m)}/{len(om)} {'ok' if ok else 'DIFF'} cluster {st['ms_cluster']:.2f}ms")
        if not ok:
            for k in range(min(len(tag), len(otag))):
                if tag[k] != otag[k] or off[k+1]-off[k] != ooff[k+1]-ooff[k] or not np.array_equal(m[off[k]:off[k+1]], om[ooff[k]:ooff[k+1]]):
                    print("   first differing cluster", k, tag[k], otag[k], m[off[k]:off[k]+4].tolist(), om[ooff[k]:ooff[k]+4].tolist(), off[k+1]-off[k], ooff[k+1]-ooff[k]); break
        bad += not ok
    if "extend" in stages:
        orow, odoff, odl = orc.alignments(); row, doff, dl = res.alignments()
        ok = np.array_equal(orow, row) and np.array_equal(odoff, doff) and np.array_equal(odl, dl)
        okt = res.delta == orc.delta()
        print(f"   alignments {len(row)}/{len(orow)} deltas {len(dl)}/{len(odl)} {'ok' if ok else 'DIFF'} text {'ok' if okt else 'DIFF'} extend {st['ms_extend']:.2f}ms cells {st['dp_cells']}/{orc.dp_cells()} launches {st['kernel_launches']}")
        if not ok:
            print("   gpu rows", row[:6].tolist()); print("   ora rows", orow[:6].tolist())
        bad += (not ok) + (not okt)
print("FAILED" if bad else "ALL OK", bad)
sys.exit(1 if bad else 0)

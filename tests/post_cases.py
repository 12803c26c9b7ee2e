"""Inputs for the two post-steps (delta-filter, delta2maf): the pair cases of cases.py plus pairs built so
that the filters have something to remove (duplicated and shuffled segments, several records, IUPAC codes)."""
from paramugsy_b200 import synth
from cases import CASES


def p_dup_in_query():
    # the query carries a second copy of segment B behind the colinear part: its alignment lies inside the
    # reference range of the long one, so the reference chain (and -1) drops it while the query chain (and -m) keeps it
    a, b, c = (synth.random_genome(n, s) for n, s in ((3_000, 901), (2_500, 902), (3_000, 903)))
    ref = a + b + c
    qry = synth.mutate(ref, 0.02, 904) + synth.random_genome(400, 906) + synth.mutate(b, 0.01, 907)
    return synth.fasta("ref.1", ref), synth.fasta("qry.1", qry), {}


def p_dup_in_reference_self():
    # a repeat inside one sequence aligned to itself (nosimplify): overlapping and nested alignments on both axes
    unit = synth.random_genome(900, 911)
    g = synth.random_genome(2_000, 912) + unit + synth.random_genome(1_500, 913) + synth.mutate(unit, 0.02, 914) + synth.random_genome(1_000, 915)
    return synth.fasta("s.1", g), synth.fasta("s.1", g), {"do_simplify": 0}


def p_shuffled_records():
    # several records on both sides, one query record is the reverse complement of its source, IUPAC codes inside
    u, v, w = (synth.random_genome(n, s) for n, s in ((4_000, 921), (3_000, 922), (2_000, 923)))
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    vr = synth.mutate(v, 0.02, 924)[::-1].translate(comp)
    ref = synth.fasta("r.1", u) + synth.fasta("r.2", v[:1500] + b"RYKMN" + v[1500:]) + synth.fasta("r.3", w.lower())
    qry = synth.fasta("q.1", vr) + synth.fasta("q.2", synth.mutate(w, 0.03, 925) + b"nnnn" + synth.mutate(u, 0.04, 926)) + synth.fasta("q.3", synth.mutate(u[1000:3000], 0.01, 927))
    return ref, qry, {}


POST_CASES = dict(CASES)
POST_CASES.update({f.__name__[2:]: f for f in (p_dup_in_query, p_dup_in_reference_self, p_shuffled_records)})

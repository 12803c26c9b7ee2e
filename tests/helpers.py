"""Independent (slow, pure Python/numpy) checkers used by the oracle and parity tests."""
import numpy as np

CODE = np.full(256, 4, dtype=np.uint8)
for _c, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
    CODE[ord(_c)] = _v
    CODE[ord(_c.lower())] = _v


def codes(seq: bytes) -> np.ndarray:
    return CODE[np.frombuffer(seq, dtype=np.uint8)]


def revcomp_codes(c: np.ndarray) -> np.ndarray:
    r = c[::-1].copy()
    m = r < 4
    r[m] = 3 - r[m]
    return r


def parse_fasta(txt: bytes):
    recs, name, chunks = [], None, []
    for line in txt.split(b"\n"):
        if line.startswith(b">"):
            if name is not None:
                recs.append((name, b"".join(chunks)))
            name, chunks = line[1:].split()[0].decode(), []
        elif name is not None:
            chunks.append(b"".join(line.split()))
    if name is not None:
        recs.append((name, b"".join(chunks)))
    return recs


def concat_codes(recs):
    parts, offs = [], []
    pos = 0
    for k, (_, s) in enumerate(recs):
        if k:
            parts.append(np.array([4], dtype=np.uint8)); pos += 1
        offs.append(pos)
        c = codes(s); parts.append(c); pos += len(c)
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8), offs


def brute_mums(ref: np.ndarray, qry: np.ndarray, minmatch: int):
    """All (r, q, L) 1-based: longest match of qry[q..] in ref, unique in ref, left-maximal.

    O(len(qry) * occurrences); only for inputs of a few thousand bases."""
    n, m = len(ref), len(qry)
    index = {}
    k = minmatch
    for i in range(n - k + 1):
        w = ref[i:i + k]
        if (w < 4).all():
            index.setdefault(w.tobytes(), []).append(i)
    out = []
    for q in range(m - k + 1):
        w = qry[q:q + k]
        if not (w < 4).all():
            continue
        cand = index.get(w.tobytes())
        if not cand:
            continue
        best, where, cnt = -1, -1, 0
        for r in cand:
            L = k
            while r + L < n and q + L < m and ref[r + L] == qry[q + L] and ref[r + L] < 4:
                L += 1
            if L > best:
                best, where, cnt = L, r, 1
            elif L == best:
                cnt += 1
        if cnt != 1:
            continue
        if q > 0 and where > 0 and ref[where - 1] == qry[q - 1] and qry[q - 1] < 4:
            continue
        out.append((where + 1, q + 1, best))
    return out


def walk_alignment(A: np.ndarray, B: np.ndarray, sA, eA, sB, eB, deltas):
    """Replay a delta-encoded alignment over 0-based code arrays (coordinates 1-based inclusive,
    B already in strand orientation).  Returns (errors, columns, first_col_match, last_col_match);
    raises AssertionError when the deltas do not tile [sA,eA] x [sB,eB] exactly."""
    a, b = sA - 1, sB - 1
    errors = cols = 0
    first = last = None
    for d in deltas:
        run = abs(int(d)) - 1
        for _ in range(run):
            ok = A[a] == B[b] and A[a] < 4
            errors += 0 if ok else 1
            if first is None:
                first = ok
            last = ok
            a += 1; b += 1; cols += 1
        errors += 1; cols += 1
        if first is None:
            first = False
        last = False
        if d > 0:
            a += 1
        else:
            b += 1
    while a <= eA - 1 and b <= eB - 1:
        ok = A[a] == B[b] and A[a] < 4
        errors += 0 if ok else 1
        if first is None:
            first = ok
        last = ok
        a += 1; b += 1; cols += 1
    assert a == eA and b == eB, f"deltas do not tile the ranges: a={a} eA={eA} b={b} eB={eB}"
    return errors, cols, first, last


def parse_delta(text: bytes):
    """Minimal .delta reader following the grammar of lib/profiles_lib/m_delta.cc:148-220."""
    lines = text.decode().split("\n")
    assert lines[-1] == "", "file must end with a newline"
    lines = lines[:-1]
    files = lines[0].rsplit(" ", 1)
    assert lines[1] == "NUCMER"
    out, hdr, i = [], None, 2
    while i < len(lines):
        if lines[i].startswith(">"):
            t = lines[i][1:].split(" ")
            assert len(t) == 4, lines[i]
            hdr = (t[0], t[1], int(t[2]), int(t[3])); i += 1
        t = lines[i].split(" ")
        assert len(t) == 7, lines[i]
        vals = list(map(int, t)); i += 1
        ds = []
        while lines[i] != "0":
            ds.append(int(lines[i])); i += 1
        i += 1
        out.append((hdr, vals, ds))
    return files, out

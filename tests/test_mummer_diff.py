"""Auto-activating comparison with the real MUMmer.

The arithmetic of this path is MUMmer 3.20's `nucmer` (scripts/pm_qsub_template.sh:4 of the reference), which is neither
vendored under /root/reference nor installed in this image: the oracle is a restatement and its parity with MUMmer is
UNPINNED.  If a MUMmer `nucmer` ever is on $PATH, under baseline/_ref/ or named by $PMN_MUMMER_NUCMER, this test runs it on
the named cases and the oracle's .delta must equal MUMmer's below line 1 (the paths); it FAILS on any difference.
Otherwise it passes with the warning "parity vs MUMmer: UNVERIFIED" — never silently."""
import glob
import os
import shutil
import subprocess
import warnings

import pytest

from cases import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.realpath(os.path.join(ROOT, "paramugsy_b200", "_lib"))


def find_mummer_nucmer():
    """A `nucmer` that is MUMmer's, not this repository's argv shim of the same name."""
    cands = []
    if os.environ.get("PMN_MUMMER_NUCMER"):
        cands.append(os.environ["PMN_MUMMER_NUCMER"])
    for d in os.environ.get("PATH", "").split(os.pathsep):
        if d:
            cands.append(os.path.join(d, "nucmer"))
    cands += glob.glob(os.path.join(ROOT, "baseline", "_ref", "**", "nucmer"), recursive=True)
    for c in cands:
        if not (os.path.isfile(c) and os.access(c, os.X_OK)):
            continue
        if os.path.realpath(os.path.dirname(c)) == OURS:
            continue
        try:
            out = subprocess.run([c, "--version"], capture_output=True, text=True, timeout=20)
        except (OSError, subprocess.TimeoutExpired):
            continue
        if "paramugsy_b200" in out.stdout + out.stderr:
            continue
        return c
    return None


def test_oracle_against_mummer_when_present(oracle, tmp_path):
    exe = find_mummer_nucmer()
    if exe is None:
        msg = "parity vs MUMmer: UNVERIFIED (no MUMmer `nucmer` on $PATH, under baseline/_ref/ or in $PMN_MUMMER_NUCMER)"
        warnings.warn(msg)
        print(msg)
        return
    bad = []
    for name in sorted(CASES):
        ref, qry, kw = CASES[name]()
        if kw:
            continue                         # cases with non-default options: compared through the default ones only
        (tmp_path / "r.fa").write_bytes(ref); (tmp_path / "q.fa").write_bytes(qry)
        p = subprocess.run([exe, str(tmp_path / "r.fa"), str(tmp_path / "q.fa"), "-p", str(tmp_path / "mm")], capture_output=True, text=True, timeout=600)
        if p.returncode != 0:
            bad.append(f"{name}: MUMmer exit {p.returncode}: {p.stderr[-200:]}")
            continue
        theirs = (tmp_path / "mm.delta").read_bytes().split(b"\n", 1)[1]
        ours = oracle.nucmer(ref, qry).split(b"\n", 1)[1]
        if theirs != ours:
            bad.append(f"{name}: .delta differs ({len(ours)} vs {len(theirs)} bytes)")
    print(f"parity vs MUMmer ({exe}): {'VERIFIED' if not bad else 'BROKEN'}")
    assert not bad, "oracle != MUMmer: " + "; ".join(bad)
    shutil.rmtree(tmp_path, ignore_errors=True)


def test_the_finder_never_takes_our_own_shim(monkeypatch):
    monkeypatch.setenv("PATH", OURS)
    monkeypatch.delenv("PMN_MUMMER_NUCMER", raising=False)
    got = find_mummer_nucmer()
    assert got is None or os.path.realpath(os.path.dirname(got)) != OURS

"""Failure paths of the batch scheduler and the bounds of its host memory (-m gpu).

The reference fails a whole node when one worker fails (lib/base/job_processor.ml:72-73) — a failing pair must therefore
end the batch call with an error, never hang it."""
import os
import subprocess
import sys
import textwrap

import pytest

from paramugsy_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_script(body, env=None, timeout=240):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n" % ROOT + textwrap.dedent(body)],
                          capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.parametrize("workers", [1, 2, 6])
def test_a_failing_pair_fails_the_batch_and_never_hangs_it(workers):
    """One live index at a time and a genome that cannot be packed: every worker that waits for an index slot must be
    woken when the batch fails (the store of `failed` is made under the waiters' mutex)."""
    p = run_script(f"""
        from paramugsy_b200 import lib, synth
        gs = [synth.random_genome(30_000, 100 + k) for k in range(4)]
        fa = [synth.fasta(f"g{{k}}.1", g) for k, g in enumerate(gs)]
        fa[3] = b"ACGT\\n>late.header\\nACGT\\n"                   # sequence data before the first header: PMN_E_ARG
        pairs = [(0, 1), (1, 2), (2, 3), (0, 3), (1, 3), (0, 2), (2, 1), (1, 0)]
        with lib.Scheduler(0, {workers}) as s:
            for rep in range(3):
                try:
                    s.align_fasta(fa, pairs)
                    print("NO ERROR")
                except lib.PmnError as e:
                    print("failed as it must:", e.code, e)
            ok = s.align_fasta(fa[:3], [(0, 1), (1, 2), (0, 2)])   # the scheduler is still usable afterwards
            print("then", len(ok), "good pairs,", sum(len(r.delta) for r in ok) > 0)
        """, env={"PMN_SCHED_LIVE_INDEXES": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.count("failed as it must: -1") == 3 and "NO ERROR" not in p.stdout and "then 3 good pairs, True" in p.stdout, p.stdout


def test_idle_pinned_memory_is_bounded(tmp_path):
    """MAF texts land in page-locked buffers that are recycled; what comes back beyond the budget is freed."""
    p = run_script("""
        from paramugsy_b200 import lib, synth
        g0 = synth.random_genome(300_000, 5); fa = [synth.fasta("a.1", g0)] + [synth.fasta(f"b{k}.1", synth.mutate(g0, 0.02, 6 + k)) for k in range(6)]
        with lib.Scheduler(0, 4) as s:
            res = s.align_fasta(fa, [(0, k) for k in range(1, 7)], post=1)
            held = lib.pinned_pool_stats()
            assert all(len(r.maf) > 100_000 for r in res)
            for r in res: r.close()
            after = lib.pinned_pool_stats()
            print(held, after)
            assert held["accounted_bytes"] >= 6 << 20 and held["idle_bytes"] == 0
            assert after["idle_bytes"] <= after["idle_budget_bytes"] == 2 << 20 and after["accounted_bytes"] == after["idle_bytes"]
            res = s.align_fasta(fa, [(0, 1)], post=1); res[0].close()       # reuse still works
        print("ok")
        """, env={"PMN_PINNED_POOL_MB": "2"})
    assert p.returncode == 0 and p.stdout.strip().endswith("ok"), (p.stdout, p.stderr[-2000:])

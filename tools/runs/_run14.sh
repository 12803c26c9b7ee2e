cat > /tmp/q15.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from paramugsy_b200 import lib, synth
anc, qs_ = synth.config_c5(n=5_000_000, ds=(0.15,))
with lib.Context(0) as ctx:
    rs = ctx.sequence(synth.fasta(*anc)); ix = rs.index()
    qs = ctx.sequence(synth.fasta(*qs_[0]))
    res = ix.align(qs); print(res.stats); res.close()
PY
PMN_JOBLOG=gpurun_out/joblog15.txt python /tmp/q15.py | tail -1
python - <<'PY'
import collections
rows = [tuple(map(int, l.split())) for l in open('gpurun_out/joblog15.txt') if l.strip() and not l.startswith("#")]
for kid in (1,2,3):
    r=[x for x in rows if x[6]==kid]
    if not r: continue
    print('kernel',kid,'calls',len(r),'cycles',sum(x[5] for x in r), 'diags', sum(x[3] for x in r))
    by=collections.Counter(); cy=collections.Counter(); dg=collections.Counter()
    for x in r: by[x[0]]+=1; cy[x[0]]+=x[5]; dg[x[0]]+=x[3]
    for k in sorted(by): print(f'   m_o {k:#x}: {by[k]} calls {cy[k]} cycles {dg[k]} diags; avg d {dg[k]/by[k]:.0f}')
    print('   slowest', sorted(r, key=lambda x:-x[5])[:5])
PY

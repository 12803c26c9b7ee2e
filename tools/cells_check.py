import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from cases import CASES
from oracle import pmn_oracle as O
from paramugsy_b200 import lib
ctx=lib.Context(0)
for name in sorted(CASES):
    ref,qry,kw=CASES[name]()
    rs,qs=ctx.sequence(ref),ctx.sequence(qry); ix=rs.index(); res=ix.align(qs,**kw)
    r=O.Run(ref,qry,fast_chain=1,**kw); r.delta("a","b")
    print(name, res.stats["dp_cells"], r.dp_cells(), res.stats["dp_jobs"], "DIFF" if res.stats["dp_cells"]!=r.dp_cells() else "")
    res.close(); ix.close(); qs.close(); rs.close()

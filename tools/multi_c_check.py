"""pmn_multi on the REAL devices of the box: one process, N GPUs through the C ABI (what the `nucmer` shim's caller and the
OCaml stub use; no Python scheduling, no torch.distributed).
  1. C2 (8 x 5 Mbp, 28 pairs) over N devices: every .delta against the oracle's committed digest, then timed steps
     (wall clock around the blocking C call — the call returns with all devices drained).
  2. one large pair (default 20 Mbp) cut over the N devices, index built per device and copied from device 0: the .delta must equal
     the one-device result byte for byte.
usage: multi_c_check.py N [steps] [large_bp]"""
import hashlib, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paramugsy_b200 import lib, synth

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
large = int(float(sys.argv[3])) if len(sys.argv) > 3 else 20_000_000
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
golden = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_configs.json")))
gs = synth.config_c2()
fa = [synth.fasta(*g) for g in gs]
names = [g[0] for g in gs]
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
out = {"devices": N, "workers_per_device": 0, "c2": {}, "large": {}}


def digests_of(results):
    return [hashlib.sha256(r.delta).hexdigest() for r in results]


def find_c2_digests(g):
    """the 28 committed sha256 of the C2 .delta texts, whatever the layout of the golden file"""
    found = []
    def walk(x):
        if isinstance(x, dict):
            if "delta_sha256" in x: found.append(x["delta_sha256"])
            for v in x.values(): walk(v)
        elif isinstance(x, list):
            for v in x: walk(v)
    walk(g.get("c2", g))
    return found


want = find_c2_digests(golden)
workers = max(4, min(32, 2 * (os.cpu_count() or 16) // N))
out["workers_per_device"] = workers
with lib.Multi(list(range(N)), workers=workers) as m:
    res = m.align_fasta(fa, pairs, names=[n + ".fa" for n in names])
    got = digests_of(res)
    for r in res: r.close()
    out["c2"]["pairs"] = len(got)
    out["c2"]["digests_in_golden"] = sum(1 for d in got if d in set(want)) if want else None
    for _ in range(3):
        for r in m.align_fasta(fa, pairs, names=[n + ".fa" for n in names]): r.close()
    walls = []
    for _ in range(steps):
        t = time.perf_counter()
        rs = m.align_fasta(fa, pairs, names=[n + ".fa" for n in names])
        walls.append((time.perf_counter() - t) * 1e3)
        same = digests_of(rs) == got
        for r in rs: r.close()
        if not same: raise SystemExit("C2 results changed between steps")
    walls.sort()
    out["c2"]["ms_per_step_median"] = walls[len(walls) // 2]; out["c2"]["ms_per_step_min"] = walls[0]
    out["c2"]["pairs_per_s_e2e"] = 28e3 / out["c2"]["ms_per_step_median"]
print(json.dumps(out["c2"]), flush=True)

g = synth.config_c4(n=large, inv_len=max(1000, large // 100))
ref, qry = synth.fasta(*g[0]), synth.fasta(*g[1])
with lib.Multi([0], workers=1) as m1:
    r1, ms1 = m1.align_large(ref, qry, "c0.fa", "c1.fa"); one = r1.delta; r1.close()
    r1, ms1 = m1.align_large(ref, qry, "c0.fa", "c1.fa"); r1.close()
out["large"] = {"bp": large, "ms_1_device": ms1, "sha256": hashlib.sha256(one).hexdigest()}
for index in ("build", "copy"):
    os.environ["PMN_MULTI_INDEX"] = index
    with lib.Multi(list(range(N)), workers=1) as m:
        r, ms = m.align_large(ref, qry, "c0.fa", "c1.fa"); same = r.delta == one; r.close()
        r, ms = m.align_large(ref, qry, "c0.fa", "c1.fa"); same = same and r.delta == one; r.close()
    out["large"][f"ms_{N}_devices_index_{index}"] = ms; out["large"][f"identical_{index}"] = same
print(json.dumps(out))
if want and out["c2"]["digests_in_golden"] != 28: raise SystemExit("C2: a .delta digest is not the oracle's")
if not (out["large"]["identical_build"] and out["large"]["identical_copy"]): raise SystemExit("large pair: the result depends on the number of devices")

"""tools/bench_stages.py -- quick per-stage timing of one resident step at a reduced workload (dev aid)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastani_b200 as fb
from fastani_b200 import parallel

clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
strains = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = int(sys.argv[3]) if len(sys.argv) > 3 else 5_000_000
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
ctx = fb.Context(fb.Parameters())
for kv in sys.argv[5:]:                       # run-time switches, e.g. l2_stage=0
    k_, v_ = kv.split("=")
    ctx.set_flag(k_, int(v_))
nG = clusters * strains
host = ctx.pinned(nG * L)
for g in range(nG):
    c, s = divmod(g, strains)
    ctx.synth_genome(3, c + 1, s, 6000 * s, L, out=host[g * L:(g + 1) * L])
off = np.arange(nG + 1, dtype=np.int64) * L
gen_off = np.arange(nG + 1, dtype=np.int32)
gs = ctx.genomes_from_buffer(host, off, gen_off)
for it in range(reps):
    ctx.profile(True); ctx.profile_read()
    t = time.time()
    sk = fb.Sketch(ctx, gs)
    ctx.sync(); t1 = time.time()
    res, tot, ctr = fb.compute_cgi(ctx, sk, gs)
    ctx.sync()
    dt = time.time() - t
    prof = ctx.profile_read()
    sk.close()
print("step %.1f ms (index %.1f + map %.1f)  pairs/s %.0f  counters %s" % (dt * 1e3, (t1 - t) * 1e3, (dt - (t1 - t)) * 1e3, nG * nG / dt, ctr.as_dict()))
for k, v in prof.items():
    print("  %-22s %9.2f ms  x%d" % (k, v[0], v[2]))

"""tools/gpu_check.py -- verbose parity diagnostics on a GPU box (development aid; the graded
tests are tests/test_*_gpu.py).  Prints a detailed report instead of stopping at the first failure."""
import json
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import fastani_b200 as fb          # noqa: E402
import pyoracle as po              # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
FAILS = []


def check(name, ok, detail=""):
    print(("PASS " if ok else "FAIL ") + name + (" :: " + detail if detail else ""), flush=True)
    if not ok:
        FAILS.append(name)


def first_diff(a, b):
    n = min(len(a), len(b))
    for f in a.dtype.names:
        d = np.nonzero(a[f][:n] != b[f][:n])[0]
        if len(d):
            i = int(d[0])
            return "field %s first diff at %d: got %s want %s (n_got=%d n_want=%d, %d diffs)" % (
                f, i, a[max(0, i - 1):i + 2], b[max(0, i - 1):i + 2], len(a), len(b), len(d))
    return "lengths %d vs %d" % (len(a), len(b))


def section(fn):
    print("==== " + fn.__name__, flush=True)
    t = time.time()
    try:
        fn()
    except Exception:
        traceback.print_exc()
        FAILS.append(fn.__name__ + " (exception)")
    print("---- %s %.2fs" % (fn.__name__, time.time() - t), flush=True)


def t_pack():
    ctx = fb.Context(fb.Parameters())
    edge = fb.read_fasta(os.path.join(G, "edge_mixed.fa"))
    g = ctx.genome(edge)
    inf = g.info()
    print(inf)
    for c, (nm, sq) in enumerate(edge):
        want = po.upper(sq)
        got = g.decode(c)
        check("pack/decode %s" % nm, len(got) == len(want) and bool((got == want).all()),
              "" if len(got) == len(want) and (got == want).all() else "first diff %s" % (np.nonzero(got != want)[0][:5],))
    nexc = sum(int(np.isin(po.upper(sq), np.frombuffer(b"ACGT", np.uint8), invert=True).sum()) for _, sq in edge)
    check("exception count", inf["n_exceptions"] == nexc, "%d vs %d" % (inf["n_exceptions"], nexc))


def t_sketch_edge():
    edge = fb.read_fasta(os.path.join(G, "edge_mixed.fa"))
    for k, w in [(16, 24), (21, 15), (16, 13), (16, 40), (11, 5), (32, 3), (7, 1), (24, 64)]:
        ctx = fb.Context(fb.Parameters(kmerSize=k, windowSize=w))
        g = ctx.genome(edge)
        sk = fb.Sketch(ctx, [g])
        got = sk.minimizerIndex()
        want = np.fromfile(os.path.join(G, "edge_mixed.k%dw%d.mi" % (k, w)), dtype=fb.MINIMIZER_DTYPE)
        ok = len(got) == len(want) and bool((got == want).all())
        check("sketch edge k%d w%d" % (k, w), ok, "" if ok else first_diff(got, want))
        st = sk.stats()
        u = len(np.unique(want["hash"]))
        check("unique k%d w%d" % (k, w), st["n_unique"] == u, "%d vs %d" % (st["n_unique"], u))


def t_sketch_real():
    import hashlib
    sums = json.load(open(os.path.join(G, "sketch_sha256.json")))
    ec = fb.read_fasta(os.path.join(G, "Escherichia_coli_str_K12_MG1655.fna.gz"))
    sh = fb.read_fasta(os.path.join(G, "Shigella_flexneri_2a_01.fna.gz"))
    for tag, gen in (("ecoli", ec), ("shigella", sh)):
        for k, w in [(16, 24), (21, 15)]:
            ctx = fb.Context(fb.Parameters(kmerSize=k, windowSize=w))
            g = ctx.genome(gen)
            t = time.time()
            sk = fb.Sketch(ctx, [g])
            dt = time.time() - t
            got = sk.minimizerIndex()
            s = sums["%s.k%dw%d" % (tag, k, w)]
            ok = len(got) == s["records"] and hashlib.sha256(got.tobytes()).hexdigest() == s["sha256"]
            detail = "records %d vs %d, build %.3fs, %s" % (len(got), s["records"], dt, sk.stats())
            if not ok:
                want, _, _ = po.sketch_genomes([gen], k, w)
                detail += " | " + first_diff(got, want)
            check("sketch %s k%d w%d" % (tag, k, w), ok, detail)
            if tag == "ecoli" and k == 16:
                want, _, _ = po.sketch_genomes([gen], k, w)
                h = int(want["hash"][1000])
                hits, n = sk.lookup(h)
                wantpos = [(int(r["seqId"]), int(r["wpos"])) for r in want[want["hash"] == h]]
                check("lookup", hits == wantpos, "%s vs %s" % (hits, wantpos))
                hits, n = sk.lookup(12345)
                check("lookup miss", n == int((want["hash"] == 12345).sum()))


def t_map():
    ec = fb.read_fasta(os.path.join(G, "Escherichia_coli_str_K12_MG1655.fna.gz"))
    sh = fb.read_fasta(os.path.join(G, "Shigella_flexneri_2a_01.fna.gz"))
    ctx = fb.Context(fb.Parameters())
    ge, gs = ctx.genomes([ec, sh])
    sk = fb.Sketch(ctx, [ge])
    t = time.time()
    m = fb.Map(ctx, sk, gs)
    dt = time.time() - t
    want = np.fromfile(os.path.join(G, "s2e.k16.map"), dtype=fb.MAPPING_DTYPE)
    got = m.rows
    ok = len(got) == len(want) and got.tobytes() == want.tobytes()
    print("map time %.3fs counters %s frags %d" % (dt, m.counters.as_dict(), m.totalQueryFragments))
    check("map s2e rows", ok, "" if ok else first_diff(got, want))
    check("map s2e fragments", m.totalQueryFragments == 1608)
    # oracle counters
    rec, sbf, _ = po.sketch_genomes([ec], 16, 24)
    rows, tot, ctr = po.map_genome(po.Index(rec), sh, 16, 24, 3000)
    oc = {f[0]: getattr(ctr, f[0]) for f in ctr._fields_}
    gc = m.counters.as_dict()
    check("counters", all(gc[k] == oc[k] for k in ("sum_s", "hits", "candidates", "n2", "mappings")), "%s vs %s" % (gc, oc))
    # CGI
    res, tots, _ = fb.compute_cgi(ctx, sk, [gs])
    want_cgi = po.cgi(want, sbf, 3000)
    print(res, want_cgi)
    ok = len(res) == len(want_cgi) and all(int(r["refGenomeId"]) == w[0] and int(r["countSeq"]) == w[1] and
                                           np.float32(r["identity"]) == w[2] for r, w in zip(res, want_cgi))
    check("cgi s2e", ok)
    # reverse direction, both in one index: 2 refs x 2 queries
    sk2 = fb.Sketch(ctx, [ge, gs])
    res2, tots2, _ = fb.compute_cgi(ctx, sk2, [gs, ge])
    print(res2, tots2)
    rec2, sbf2, _ = po.sketch_genomes([ec, sh], 16, 24)
    ix2 = po.Index(rec2)
    exp = []
    for qi, q in enumerate([sh, ec]):
        rws, tq, _ = po.map_genome(ix2, q, 16, 24, 3000)
        mm = fb.Map(ctx, sk2, [gs, ge][qi])
        okr = len(mm.rows) == len(rws) and mm.rows.tobytes() == rws.tobytes()
        check("map 2x2 rows q%d" % qi, okr, "" if okr else first_diff(mm.rows, rws))
        for (g, c, idn) in po.cgi(rws, sbf2, 3000):
            exp.append((qi, g, c, idn, tq))
    got2 = [(int(r["qryGenomeId"]), int(r["refGenomeId"]), int(r["countSeq"]), np.float32(r["identity"]), int(r["totalQueryFragments"])) for r in res2]
    check("cgi 2x2", got2 == exp, "%s vs %s" % (got2, exp))


def t_map_sweep():
    import hashlib
    ec = fb.read_fasta(os.path.join(G, "Escherichia_coli_str_K12_MG1655.fna.gz"))
    sh = fb.read_fasta(os.path.join(G, "Shigella_flexneri_2a_01.fna.gz"))
    sums = json.load(open(os.path.join(G, "map_sha256.json")))
    for k, L in [(16, 1000), (16, 5000), (21, 3000), (21, 5000), (21, 1000)]:
        ctx = fb.Context(fb.Parameters(kmerSize=k, minReadLength=L))
        ge, gs = ctx.genomes([ec, sh])
        sk = fb.Sketch(ctx, [ge])
        m = fb.Map(ctx, sk, gs)
        if (k, L) == (21, 1000):
            check("map sweep k21 L1000 empty", len(m.rows) == 0, "w=%d rows=%d" % (ctx.windowSize, len(m.rows)))
            continue
        s = sums["s2e.k%d.L%d" % (k, L)]
        ok = len(m.rows) == s["records"] and hashlib.sha256(m.rows.tobytes()).hexdigest() == s["sha256"]
        check("map sweep k%d L%d" % (k, L), ok, "w=%d rows %d vs %d" % (ctx.windowSize, len(m.rows), s["records"]))


def t_synth():
    from fastani_b200.synth import synth_genome
    ctx = fb.Context(fb.Parameters())
    for (a, s, ppm, n) in [(0, 0, 0, 10000), (3, 5, 30000, 100001), (7, 19, 114000, 5000)]:
        got = ctx.synth_genome(3, a, s, ppm, n)
        want = synth_genome(3, a, s, ppm, n)
        check("synth a%d s%d" % (a, s), bool((got == want).all()))
    a0 = synth_genome(3, 1, 0, 0, 200000); a5 = synth_genome(3, 1, 5, 30000, 200000)
    print("observed substitution rate", float((a0 != a5).mean()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["t_pack", "t_synth", "t_sketch_edge", "t_sketch_real", "t_map", "t_map_sweep"]
    for w in which:
        section(globals()[w])
    print("FAILS:", FAILS)
    sys.exit(1 if FAILS else 0)

"""GPU (-m gpu): parity of the CUDA path, through the C ABI, against the oracle and the golden fixtures."""
import hashlib
import json
import os

import numpy as np
import pytest

import fastani_b200 as fb
import pyoracle as po
from conftest import GOLDEN
from fastani_b200 import parallel
from fastani_b200.synth import synth_genome

pytestmark = pytest.mark.gpu

EC = os.path.join(GOLDEN, "Escherichia_coli_str_K12_MG1655.fna.gz")
SH = os.path.join(GOLDEN, "Shigella_flexneri_2a_01.fna.gz")


@pytest.fixture(scope="module")
def real():
    return fb.read_fasta(EC), fb.read_fasta(SH)


@pytest.fixture(scope="module")
def edge():
    return fb.read_fasta(os.path.join(GOLDEN, "edge_mixed.fa"))


def test_pack_roundtrip_and_exceptions(edge):
    ctx = fb.Context(fb.Parameters())
    g = ctx.genome(edge)
    nexc = 0
    for c, (_, sq) in enumerate(edge):
        want = po.upper(sq)
        assert (g.decode(c) == want).all()
        nexc += int(np.isin(want, np.frombuffer(b"ACGT", np.uint8), invert=True).sum())
    inf = g.info()
    assert inf["n_exceptions"] == nexc and inf["n_contigs"] == 16
    empty = ctx.genome([])
    assert empty.info()["n_contigs"] == 0


@pytest.mark.parametrize("k,w", [(16, 24), (21, 15), (16, 13), (16, 40), (11, 5), (32, 3), (7, 1), (24, 64)])
def test_sketch_edge_cases_vs_reference_golden(edge, k, w):
    ctx = fb.Context(fb.Parameters(kmerSize=k, windowSize=w))
    sk = fb.Sketch(ctx, [ctx.genome(edge)])
    want = np.fromfile(os.path.join(GOLDEN, "edge_mixed.k%dw%d.mi" % (k, w)), dtype=fb.MINIMIZER_DTYPE)
    got = sk.minimizerIndex()
    assert len(got) == len(want) and (got == want).all()
    assert sk.stats()["n_unique"] == len(np.unique(want["hash"]))
    assert sk.sequencesByFileInfo == [16]


def test_sketch_real_genomes_sha(real):
    sums = json.load(open(os.path.join(GOLDEN, "sketch_sha256.json")))
    for tag, gen in zip(("ecoli", "shigella"), real):
        for k, w in [(16, 24), (21, 15)]:
            ctx = fb.Context(fb.Parameters(kmerSize=k, windowSize=w))
            sk = fb.Sketch(ctx, [ctx.genome(gen)])
            got = sk.minimizerIndex()
            s = sums["%s.k%dw%d" % (tag, k, w)]
            assert len(got) == s["records"] and hashlib.sha256(got.tobytes()).hexdigest() == s["sha256"]
            if tag == "ecoli" and k == 16:
                assert sk.stats()["n_unique"] == 361568
                h = int(got["hash"][1000])
                hits, n = sk.lookup(h)
                assert hits == [(int(r["seqId"]), int(r["wpos"])) for r in got[got["hash"] == h]]
                assert sk.lookup(12345)[1] == int((got["hash"] == 12345).sum())


def test_empty_and_degenerate_inputs():
    ctx = fb.Context(fb.Parameters())
    sk = fb.Sketch(ctx, [])                                   # a shard without references
    q = ctx.genome([("q", synth_genome(1, 1, 0, 0, 20000).tobytes())])
    m = fb.Map(ctx, sk, q)
    assert len(m.rows) == 0 and m.totalQueryFragments == 6
    res, tot, _ = fb.compute_cgi(ctx, sk, [q])
    assert len(res) == 0 and int(tot[0]) == 6
    # all-N and too-short queries against a real index
    sk2 = fb.Sketch(ctx, [q])
    for seq, frags in ((b"N" * 9000, 3), (b"ACGT" * 100, 0), (b"", 0)):
        m = fb.Map(ctx, sk2, ctx.genome([("x", seq)]))
        assert len(m.rows) == 0 and m.totalQueryFragments == frags
    assert fb.Map(ctx, sk2, ctx.genome([])).totalQueryFragments == 0


def test_map_real_pair_rows_counters_cgi(real):
    ec, sh = real
    ctx = fb.Context(fb.Parameters())
    ge, gs = ctx.genomes([ec, sh])
    sk = fb.Sketch(ctx, [ge])
    m = fb.Map(ctx, sk, gs)
    want = np.fromfile(os.path.join(GOLDEN, "s2e.k16.map"), dtype=fb.MAPPING_DTYPE)
    assert m.totalQueryFragments == 1608
    assert m.rows.tobytes() == want.tobytes()                 # all 4138 records, every field, reference order
    c = m.counters.as_dict()
    assert (c["sum_s"], c["hits"], c["candidates"], c["n2"], c["mappings"]) == (381646, 311331, 5058, 2815830, 4138)
    res, tot, _ = fb.compute_cgi(ctx, sk, [gs])
    assert [(int(r["refGenomeId"]), int(r["countSeq"]), "%g" % r["identity"]) for r in res] == [(0, 1303, "97.7507")]
    # callback form of skch::Map
    seen = []
    fb.Map(ctx, sk, gs, f=lambda r: seen.append(int(r["querySeqId"])))
    assert len(seen) == 4138 and seen == sorted(seen)


def test_two_refs_two_queries_against_oracle(real):
    ec, sh = real
    ctx = fb.Context(fb.Parameters())
    ge, gs = ctx.genomes([ec, sh])
    sk = fb.Sketch(ctx, [ge, gs])
    assert sk.sequencesByFileInfo == [1, 3]
    rec, sbf, _ = po.sketch_genomes([ec, sh], 16, 24)
    ix = po.Index(rec)
    exp = []
    for qi, (q, gq) in enumerate(((sh, gs), (ec, ge))):
        rows, tq, _ = po.map_genome(ix, q, 16, 24, 3000)
        assert fb.Map(ctx, sk, gq).rows.tobytes() == rows.tobytes()
        exp += [(qi, g, c, np.float32(i), tq) for g, c, i in po.cgi(rows, sbf, 3000)]
    res, tot, _ = fb.compute_cgi(ctx, sk, [gs, ge])
    got = [(int(r["qryGenomeId"]), int(r["refGenomeId"]), int(r["countSeq"]), np.float32(r["identity"]),
            int(r["totalQueryFragments"])) for r in res]
    assert got == exp
    assert ["%g" % r["identity"] for r in res] == ["97.7507", "100", "100", "97.664"]     # README.md:80, fastani_tests.cpp:61


@pytest.mark.parametrize("k,L", [(16, 1000), (16, 5000), (21, 3000), (21, 5000), (21, 1000)])
def test_parameter_sweep_vs_reference(real, k, L):
    ec, sh = real
    sums = json.load(open(os.path.join(GOLDEN, "map_sha256.json")))
    ctx = fb.Context(fb.Parameters(kmerSize=k, minReadLength=L))
    ge, gs = ctx.genomes([ec, sh])
    m = fb.Map(ctx, fb.Sketch(ctx, [ge]), gs)
    if (k, L) == (21, 1000):
        assert ctx.windowSize == 1000 and len(m.rows) == 0
        return
    s = sums["s2e.k%d.L%d" % (k, L)]
    assert len(m.rows) == s["records"] and hashlib.sha256(m.rows.tobytes()).hexdigest() == s["sha256"]


def _cluster_set(n_clusters, n_strains, L, seed=7):
    gs = []
    for c in range(n_clusters):
        for s in range(n_strains):
            seq = synth_genome(seed, c + 1, s, 6000 * s, L)
            if s == 2:                      # one multi-contig strain per cluster, with an N gap and lower case
                a = seq.tobytes()
                gs.append([("c%ds%d_a" % (c, s), a[:L // 3].lower()), ("c%ds%d_tiny" % (c, s), a[L // 3:L // 3 + 500]),
                           ("c%ds%d_b" % (c, s), a[L // 3 + 500:L // 2] + b"N" * 700 + a[L // 2:])])
            else:
                gs.append([("c%ds%d" % (c, s), seq.tobytes())])
    return gs


def test_many_to_many_synthetic_vs_oracle_and_shard_invariance():
    """12 x 12 synthetic clusters: every CGI row equals the oracle's; mapping in 3 shards and merging
    (the multi-GPU rule) equals the single-shard result; self pairs are 100 % over all fragments."""
    L = 90000
    genomes = _cluster_set(3, 4, L)
    ctx = fb.Context(fb.Parameters())
    hs = ctx.genomes(genomes)
    n = len(hs)
    sk = fb.Sketch(ctx, hs)
    res, tot, ctr = fb.compute_cgi(ctx, sk, hs)
    cnt, idn, _ = parallel.dense_tables(res, n, n)
    # oracle
    rec, sbf, _ = po.sketch_genomes(genomes, 16, 24)
    ix = po.Index(rec)
    assert (sk.minimizerIndex() == rec).all()
    ocnt = np.zeros((n, n), np.int32); oidn = np.zeros((n, n), np.float32)
    for qi, q in enumerate(genomes):
        rows, tq, _ = po.map_genome(ix, q, 16, 24, 3000)
        assert tq == int(tot[qi])
        if qi % 5 == 0:
            assert fb.Map(ctx, sk, hs[qi]).rows.tobytes() == rows.tobytes()
        for g, c, i in po.cgi(rows, sbf, 3000):
            ocnt[qi, g] = c; oidn[qi, g] = i
    assert (cnt == ocnt).all()
    assert (idn.view(np.uint32) == oidn.view(np.uint32)).all()
    for i in range(n):
        # a fragment next to a contig end is never scored (computeMap.hpp:455): allow one miss per contig
        assert tot[i] - len(genomes[i]) <= cnt[i, i] <= tot[i] and idn[i, i] > 99.9
    assert int((cnt > 0).sum()) >= 3 * 16
    # shard invariance
    G = 3
    tabs = []
    for g in range(G):
        ids = parallel.shard_refs(n, G, g)
        r, _, _ = fb.compute_cgi(ctx, fb.Sketch(ctx, [hs[i] for i in ids]), hs)
        tabs.append(parallel.dense_tables(r, n, len(ids)))
    mc, mi, mt = parallel.merge_shards(tabs, n, n, G)
    assert (mt == np.asarray(tot, np.int64)).all()
    assert (mc == cnt).all() and (mi.view(np.uint32) == idn.view(np.uint32)).all()


def test_full_size_properties():
    """Size-independent properties at BASELINE genome size (5 Mbp): sketch ordering / density, index
    consistency, self-mapping = 100 % for every fragment, strain pair close to its simulated divergence."""
    L = 5_000_000
    ctx = fb.Context(fb.Parameters())
    a = ctx.synth_genome(3, 1, 0, 0, L); b = ctx.synth_genome(3, 1, 4, 24000, L)
    ga, gb = ctx.genomes([[("a", a)], [("b", b)]])
    sk = fb.Sketch(ctx, [ga, gb])
    rec = sk.minimizerIndex()
    st = sk.stats()
    key = rec["seqId"].astype(np.int64) << 32 | rec["wpos"].astype(np.int64)
    assert (np.diff(key) > 0).all()                                    # ordered, (seqId, wpos) unique
    assert abs(len(rec) / (2 * L) - 2 / 25) < 0.002                    # density 2/(w+1)
    assert st["n_unique"] == len(np.unique(rec["hash"]))
    res, tot, _ = fb.compute_cgi(ctx, sk, [ga, gb])
    cnt, idn, _ = parallel.dense_tables(res, 2, 2)
    assert int(tot[0]) == L // 3000 and cnt[0, 0] >= tot[0] - 1 and cnt[1, 1] >= tot[1] - 1
    assert idn[0, 0] > 99.99 and idn[1, 1] > 99.99
    assert cnt[0, 1] > 0.95 * tot[0] and abs(float(idn[0, 1]) - 97.6) < 0.5


def test_device_synth_matches_numpy_twin():
    ctx = fb.Context(fb.Parameters())
    for a, s, ppm, n in [(0, 0, 0, 10000), (3, 5, 30000, 100001), (7, 19, 114000, 5000)]:
        assert (ctx.synth_genome(3, a, s, ppm, n) == synth_genome(3, a, s, ppm, n)).all()


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


_PATH_SCRIPT = r"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import fastani_b200 as fb
G = os.path.join(sys.argv[1], "tests", "golden")
ec = fb.read_fasta(os.path.join(G, "Escherichia_coli_str_K12_MG1655.fna.gz"))
sh = fb.read_fasta(os.path.join(G, "Shigella_flexneri_2a_01.fna.gz"))
ctx = fb.Context(fb.Parameters())
ge, gs = ctx.genomes([ec, sh])
sk = fb.Sketch(ctx, [ge, gs])
for q in (gs, ge):
    m = fb.Map(ctx, sk, q)
    print(len(m.rows), hashlib.sha256(m.rows.tobytes()).hexdigest(), m.counters.as_dict()["candidates"])
"""


def test_hit_paths_agree():
    """The per-fragment shared-memory path (hits.cu), the device-wide sort path and their mix give the same rows."""
    import subprocess, sys
    from conftest import ROOT
    outs = []
    for cap, nb, stg in (("8192", "4096", "0"), ("0", "1024", "1"), ("300", "4096", "1"), ("1500", "1024", "0")):   # all fast / all device-wide / two mixes; both L2 directory sizes; direct / staged event stores
        env = dict(os.environ, BANI_FRAG_L1_MAX=cap, BANI_L2E_BUCKETS=nb, BANI_L2_STAGE=stg)
        r = subprocess.run([sys.executable, "-c", _PATH_SCRIPT, ROOT], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2] == outs[3] and outs[0].count("\n") == 2


def _cli_run(args, cwd):
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "fastani_b200", "bin", "fastANI")
    r = subprocess.run([exe] + args, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r


def test_cli_reproduces_the_reference_goldens(tmp_path):
    """The C++ command line (fastani_b200/bin/fastANI) on the reference's own offline golden
    (tests/fastani_tests.cpp:50-72: .txt, .matrix, .visual byte for byte) and README.md:80."""
    os.mkdir(tmp_path / "data")
    for n in ("Escherichia_coli_str_K12_MG1655.fna", "Shigella_flexneri_2a_01.fna"):
        os.symlink(os.path.join(GOLDEN, n + ".gz"), tmp_path / "data" / n)          # gzopen reads both, so does the CLI
    E, S = "data/Escherichia_coli_str_K12_MG1655.fna", "data/Shigella_flexneri_2a_01.fna"
    for tag, q, r in (("e2s", E, S), ("s2e", S, E)):
        _cli_run(["-q", q, "-r", r, "-o", tag + ".txt", "--matrix", "--visualize", "--gpus", "1"], tmp_path)
        for ext in ("", ".matrix", ".visual"):
            got = open(tmp_path / (tag + ".txt" + ext)).read()
            assert got == open(os.path.join(GOLDEN, tag + ".txt" + ext)).read(), tag + ext
        # without --visualize the per-pair reduction runs on the device: same .txt
        _cli_run(["-q", q, "-r", r, "-o", tag + "_dev.txt", "--gpus", "1"], tmp_path)
        assert open(tmp_path / (tag + "_dev.txt")).read() == open(os.path.join(GOLDEN, tag + ".txt")).read()
    # sanity check on a pure repeat: no output rows (tests/fastani_tests.cpp:302-416)
    rep = tmp_path / "rep.fa"
    rep.write_bytes(b">rep\n" + b"AT" * 20000 + b"\n")
    _cli_run(["-q", "rep.fa", "-r", "rep.fa", "-o", "rep.txt", "-s", "--maxRatioDiff", "10", "--gpus", "1"], tmp_path)
    assert open(tmp_path / "rep.txt").read() == ""


def test_cli_many_to_many_lists_match_the_python_host(tmp_path):
    genomes = _cluster_set(2, 3, 60000)
    paths = []
    for i, g in enumerate(genomes):
        p = tmp_path / ("g%d.fa" % i)
        with open(p, "wb") as f:
            for name, seq in g:
                f.write(b">" + name.encode() + b"\n")
                for o in range(0, len(seq), 70):
                    f.write(seq[o:o + 70] + b"\n")
        paths.append(str(p))
    open(tmp_path / "ql.txt", "w").write("\n".join(paths[:4]) + "\n\n")
    open(tmp_path / "rl.txt", "w").write("  " + "\n".join(paths[1:]) + "\n")
    _cli_run(["--ql", "ql.txt", "--rl", "rl.txt", "-o", "out.txt", "--matrix", "-t", "3", "--gpus", "1", "--minFraction", "0.1"], tmp_path)
    # the same through the Python host
    from fastani_b200 import report
    ctx = fb.Context(fb.Parameters())
    hs = ctx.genomes(genomes)
    q, r = list(range(4)), list(range(1, len(genomes)))
    sk = fb.Sketch(ctx, [hs[i] for i in r])
    res, tot, _ = fb.compute_cgi(ctx, sk, [hs[i] for i in q])
    lens = [report.genome_length([len(s) for _, s in g], 3000) for g in genomes]
    rows = [(int(x["qryGenomeId"]), int(x["refGenomeId"]), int(x["countSeq"]), int(x["totalQueryFragments"]), x["identity"]) for x in res]
    want = report.output_lines(rows, [paths[i] for i in q], [paths[i] for i in r], [lens[i] for i in q], [lens[i] for i in r], 3000, 0.1)
    got = open(tmp_path / "out.txt").read().splitlines()
    assert sorted(got) == sorted(want) and len(got) >= 6
    assert [g.split("\t")[0] for g in got] == sorted(g.split("\t")[0] for g in got)      # grouped by query, in list order
    m = open(tmp_path / "out.txt.matrix").read().splitlines()
    assert m[0] == str(len(genomes)) and m[1] == paths[0] and len(m) == 1 + len(genomes)
    # host CGI path (--visualize) gives the same table
    _cli_run(["--ql", "ql.txt", "--rl", "rl.txt", "-o", "vis.txt", "--visualize", "--gpus", "1", "--minFraction", "0.1", "--partition", "block"], tmp_path)
    assert open(tmp_path / "vis.txt").read() == open(tmp_path / "out.txt").read()


def test_query_sketch_objects_export_import_and_map():
    """bani_qsketch_*: sketching the queries in two halves (as two ranks would), moving one half through a flat device
    buffer, and mapping both halves against the index gives exactly compute_cgi's rows."""
    import torch
    genomes = _cluster_set(2, 4, 70000)
    ctx = fb.Context(fb.Parameters())
    hs = ctx.genomes(genomes)
    n = len(hs)
    sk = fb.Sketch(ctx, hs[1:])
    want, tot, ctr = fb.compute_cgi(ctx, sk, hs)
    even, odd = list(range(0, n, 2)), list(range(1, n, 2))
    s0 = fb.QuerySketch(ctx, [hs[i] for i in even], even)
    s1 = fb.QuerySketch(ctx, [hs[i] for i in odd], odd)
    i1 = s1.info()
    assert i1["n_queries"] == len(odd) and i1["n_fragments"] == int(sum(tot[i] for i in odd)) and i1["export_bytes"] % 16 == 0
    buf = torch.empty(i1["export_bytes"] + 64, dtype=torch.uint8, device="cuda")
    s1.export_to(buf.data_ptr(), buf.numel())
    s1b = fb.QuerySketch.from_device_buffer(ctx, buf.data_ptr(), i1["export_bytes"])
    assert s1b.info() == i1
    got, ctr2 = fb.compute_cgi_sketched(ctx, sk, [s0, s1b])
    key = lambda a: np.sort(a, order=["qryGenomeId", "refGenomeId"])
    assert key(got).tobytes() == key(want).tobytes()
    assert ctr2.as_dict() == ctr.as_dict()
    # merged sketches (what a rank does with the sketches of its peers): same rows, in the order given
    m = fb.QuerySketch.merge(ctx, [s1b, s0])
    assert m.info()["n_queries"] == n and m.info()["n_fragments"] == int(tot.sum())
    got2, ctr3 = fb.compute_cgi_sketched(ctx, sk, [m])
    assert key(got2).tobytes() == key(want).tobytes() and ctr3.as_dict() == ctr.as_dict()
    assert [int(x) for x in got2["qryGenomeId"][:1]] == [odd[0]] or len(got2) == 0
    # truncated / foreign buffers are rejected
    with pytest.raises(fb.BaniError):
        fb.QuerySketch.from_device_buffer(ctx, buf.data_ptr(), 32)
    with pytest.raises(fb.BaniError):
        fb.QuerySketch.from_device_buffer(fb.Context(fb.Parameters(kmerSize=21)), buf.data_ptr(), i1["export_bytes"])


_REUSE_SCRIPT = r"""
import hashlib, os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
import fastani_b200 as fb
from fastani_b200.synth import synth_genome
G = os.path.join(sys.argv[1], "tests", "golden")
edge = fb.read_fasta(os.path.join(G, "edge_mixed.fa"))
sh = fb.read_fasta(os.path.join(G, "Shigella_flexneri_2a_01.fna.gz"))
# a genome with palindromic k-mers (invalid positions), N runs and lower case around fragment borders
rng = np.random.default_rng(5)
pal = b"ACGTACGTTGCATGCA"                       # its own reverse complement: an invalid position wherever it occurs
body = synth_genome(9, 1, 0, 0, 40000).tobytes()
tricky = bytearray(body)
for pos in (0, 984, 999, 1000, 1016, 1999, 2990, 3000, 8000, 11990, 12000, 20000, 39984):
    tricky[pos:pos + 16] = pal
tricky[5000:5400] = b"N" * 400
tricky[14990:15020] = b"n" * 30
tricky[25000:25016] = b"A" * 16
for k, L in ((16, 3000), (16, 1000), (21, 3000), (11, 400)):
    ctx = fb.Context(fb.Parameters(kmerSize=k, minReadLength=L))
    gs = ctx.genomes([edge, [("tricky", bytes(tricky))], sh if L == 3000 and k == 16 else [("t2", bytes(tricky[7:]))]])
    sk = fb.Sketch(ctx, gs)
    for g in gs:
        m = fb.Map(ctx, sk, g)
        print(k, L, len(m.rows), hashlib.sha256(m.rows.tobytes()).hexdigest(), m.totalQueryFragments, m.counters.as_dict()["sum_s"])
    res, tot, _ = fb.compute_cgi(ctx, sk, gs)
    print(hashlib.sha256(res.tobytes()).hexdigest())
"""


def test_fragment_sketches_from_the_index_equal_hashed_ones():
    """Stage A' (fragment sketches read from the index a query genome is a member of) against stage A (the fragment
    hashed as a stand-alone sequence): same rows and counters on real, edge-case and palindrome/N-laden genomes."""
    import subprocess, sys
    from conftest import ROOT
    outs = []
    for env in ({}, {"BANI_NO_SKETCH_REUSE": "1"}, {"BANI_MAX_HITS_PER_PIECE": "500"}):      # the last one forces pieces to be split by queries
        r = subprocess.run([sys.executable, "-c", _REUSE_SCRIPT, ROOT], env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2] and outs[0].count("\n") == 16

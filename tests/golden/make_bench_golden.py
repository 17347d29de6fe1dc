#!/usr/bin/env python
"""Generates the reference-side goldens of the BASELINE-scale workloads by running the UNMODIFIED reference
(oracle/_ref/fastANI_ref, built by oracle/Makefile from /root/reference) on the synthetic genomes of
fastani_b200/workloads.py.  Run where /root/reference exists; the outputs are small text files:

  bench_cfg3_q8.txt            config 3 (1000 x 1000 x 5 Mbp): the 8 sample queries of bench.py against ALL 1000
                               references -- the parity gate of every timed bench run (bench.py, `parity`)
  cfg2_1x100.txt               config 2: 1 query x 100 references on the 0.2 % * j divergence ladder
  cfg4_40x40.txt               config 4 slice: 2 clusters x 20 multi-contig 3 Mbp drafts, all vs all
  cfg5_20x20.k{K}.L{L}.txt(.matrix)   config 5 slice: 20 x 20 at six (k, fragLen) with --minFraction 0.2 --matrix

Lines keep the reference's text verbatim except that the directory of the FASTA paths is stripped (genome names
only), so the files do not depend on where the FASTA files were written.

  python tests/golden/make_bench_golden.py [bench] [cfg2] [cfg4] [cfg5]      (default: all)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fastani_b200 import workloads as W     # noqa: E402

GOLDEN = os.path.dirname(os.path.abspath(__file__))
CLI = os.path.join(ROOT, "oracle", "_ref", "fastANI_ref")
TMP = os.environ.get("BANI_TMP", "/tmp")


def strip_dirs(text):
    out = []
    for ln in text.splitlines():
        f = ln.split("\t")
        f = [os.path.basename(x) if x.endswith(".fna") else x for x in f]
        out.append("\t".join(f))
    return "\n".join(out) + ("\n" if out else "")


def run(qpaths, rpaths, out, extra=(), threads=None):
    d = os.path.dirname(out)
    ql, rl = os.path.join(d, "ql.txt"), os.path.join(d, "rl.txt")
    open(ql, "w").write("\n".join(qpaths) + "\n")
    open(rl, "w").write("\n".join(rpaths) + "\n")
    t = time.time()
    r = subprocess.run([CLI, "--ql", ql, "--rl", rl, "-t", str(threads or W.available_cores()), "-o", out] + list(extra),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    sys.stderr.write("  fastANI_ref %d x %d %s: %.1f s\n" % (len(qpaths), len(rpaths), " ".join(extra), time.time() - t))


def bench():
    specs = W.config3()
    d = os.path.join(TMP, "bani_fasta_" + W.spec_key(specs))
    paths = W.materialize(specs, d)
    q = W.sample_queries(50, 20, 8)
    out = os.path.join(d, "golden_out.txt")
    run([paths[i] for i in q], paths, out)
    open(os.path.join(GOLDEN, "bench_cfg3_q8.txt"), "w").write(strip_dirs(open(out).read()))


def cfg2():
    qs, refs = W.config2()
    d = os.path.join(TMP, "bani_fasta_" + W.spec_key([qs] + refs))
    paths = W.materialize([qs] + refs, d)
    out = os.path.join(d, "golden_out.txt")
    run(paths[:1], paths[1:], out)
    open(os.path.join(GOLDEN, "cfg2_1x100.txt"), "w").write(strip_dirs(open(out).read()))


def cfg4():
    specs = W.config4(clusters=2)
    d = os.path.join(TMP, "bani_fasta_" + W.spec_key(specs))
    paths = W.materialize(specs, d)
    out = os.path.join(d, "golden_out.txt")
    run(paths, paths, out)
    open(os.path.join(GOLDEN, "cfg4_40x40.txt"), "w").write(strip_dirs(open(out).read()))


def cfg5():
    specs = W.config3(clusters=2, strains=10)          # 20 genomes: 2 clusters x 10 strains (0 .. 5.4 % divergence)
    d = os.path.join(TMP, "bani_fasta_" + W.spec_key(specs))
    paths = W.materialize(specs, d)
    for k in (16, 21):
        for L in (1000, 3000, 5000):
            out = os.path.join(d, "golden_k%d_L%d.txt" % (k, L))
            run(paths, paths, out, ["-k", str(k), "--fragLen", str(L), "--minFraction", "0.2", "--matrix"])
            base = os.path.join(GOLDEN, "cfg5_20x20.k%d.L%d.txt" % (k, L))
            open(base, "w").write(strip_dirs(open(out).read()))
            m = open(out + ".matrix").read().splitlines()
            open(base + ".matrix", "w").write("\n".join([m[0]] + ["\t".join([os.path.basename(x) if x.endswith(".fna") else x for x in ln.split("\t")]) for ln in m[1:]]) + "\n")


if __name__ == "__main__":
    what = sys.argv[1:] or ["cfg2", "cfg4", "cfg5", "bench"]
    for w in what:
        sys.stderr.write("%s\n" % w)
        {"bench": bench, "cfg2": cfg2, "cfg4": cfg4, "cfg5": cfg5}[w]()

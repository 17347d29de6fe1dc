"""tools/cli_e2e.py [clusters] [strains] -- FASTA files on local disk -> out.txt through the C++ command line
(fastani_b200/bin/fastANI): the end-to-end figure of SURVEY.md 8(d)(ii), on a synthetic many-to-many set."""
import os, subprocess, sys, tempfile, time, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastani_b200 as fb
from fastani_b200 import workloads as W

clusters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
strains = int(sys.argv[2]) if len(sys.argv) > 2 else 20
L = 5_000_000
ctx = fb.Context(fb.Parameters())
tmp = tempfile.mkdtemp(prefix="bani_cli_")
try:
    paths = []
    for c in range(clusters):
        for s in range(strains):
            p = os.path.join(tmp, "c%d_s%d.fna" % (c, s))
            W.write_fasta(p, [("c%d_s%d" % (c, s), ctx.synth_genome(3, c + 1, s, 6000 * s, L).tobytes())])
            paths.append(p)
    lst = os.path.join(tmp, "all.txt")
    open(lst, "w").write("\n".join(paths) + "\n")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fastani_b200", "bin", "fastANI")
    for it in range(2):
        t = time.time()
        r = subprocess.run([exe, "--ql", lst, "--rl", lst, "-o", os.path.join(tmp, "out.txt"), "--matrix", "-t", str(min(64, os.cpu_count() or 1)), "--gpus", "1"],
                           capture_output=True, text=True)
        dt = time.time() - t
        assert r.returncode == 0, r.stderr[-2000:]
    n = len(paths)
    rows = sum(1 for _ in open(os.path.join(tmp, "out.txt")))
    print("CLI end to end: %d x %d genomes of %.1f Mbp from FASTA on local disk to out.txt(+.matrix): %.2f s -> %.0f pairs/s (%d output rows)" % (n, n, L / 1e6, dt, n * n / dt, rows))
    print("\n".join(l for l in r.stderr.splitlines() if "Time spent" in l))
finally:
    shutil.rmtree(tmp, ignore_errors=True)

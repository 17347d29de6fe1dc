"""CPU, world_size 2 over gloo: the N>1 host path (shard references round-robin, every shard maps all
queries, gather the dense per-pair tables) gives the single-shard answer.  The per-shard compute here is
the oracle (no GPU in this container); on the GPU box tests/test_gpu_parallel.py runs the same merge
with the CUDA path."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["ROOT"]); sys.path.insert(0, os.path.join(os.environ["ROOT"], "oracle"))
import pyoracle as po
from fastani_b200 import parallel
from fastani_b200.synth import synth_genome

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
L = 45000
genomes = [[("g%d" % i, synth_genome(5, i // 3, i % 3, 20000 * (i % 3), L).tobytes())] for i in range(5)]
queries = genomes[:3]

def run(ref_ids):
    refs = [genomes[i] for i in ref_ids]
    rec, sbf, _ = po.sketch_genomes(refs, 16, 24)
    ix = po.Index(rec)
    cnt = np.zeros((len(queries), len(refs)), np.int32); idn = np.zeros((len(queries), len(refs)), np.float32)
    for qi, q in enumerate(queries):
        rows, tot, _ = po.map_genome(ix, q, 16, 24, 3000)
        for g, c, i in po.cgi(rows, sbf, 3000):
            cnt[qi, g] = c; idn[qi, g] = i
    return cnt, idn

mine = parallel.shard_refs(len(genomes), world, rank)
c, i = run(mine)
gc, gi = parallel.gather_tables(c, i, len(genomes), world, rank, dist=dist)
if rank == 0:
    fc, fi = run(list(range(len(genomes))))
    assert (gc == fc).all() and (gi.view(np.uint32) == fi.view(np.uint32)).all(), (gc, fc)
    assert int((fc > 0).sum()) >= 7
    print("DIST_OK", int((fc > 0).sum()))
dist.barrier()
dist.destroy_process_group()
'''


def test_two_shards_gather_equals_single_shard(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DIST_OK" in r.stdout

"""CPU, world_size 2 over gloo: the N>1 host path (shard references round-robin, every shard maps all
queries, gather the dense per-pair tables) gives the single-shard answer.  The per-shard compute here is
the oracle (no GPU in this container); on the GPU box tests/test_gpu.py (shard invariance, query sketch objects) runs the same merge
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
    tq = np.zeros(len(queries), np.int32)
    for qi, q in enumerate(queries):
        rows, tot, _ = po.map_genome(ix, q, 16, 24, 3000)
        for g, c, i in po.cgi(rows, sbf, 3000):
            cnt[qi, g] = c; idn[qi, g] = i
            tq[qi] = tot                      # as the CGI rows carry it: known only where the query has a row
    return cnt, idn, tq

mine = parallel.shard_refs(len(genomes), world, rank)
c, i, t = run(mine)
gc, gi, gt = parallel.gather_tables(c, i, len(genomes), world, rank, dist=dist, tot=t)
gc2, gi2 = parallel.gather_tables(c, i, len(genomes), world, rank, dist=dist)
assert (gc2 == gc).all() and (gi2.view(np.uint32) == gi.view(np.uint32)).all()
# the compact-row exchange, for the reference's round-robin deal and for contiguous blocks
from fastani_b200.api import CGI_DTYPE
fc0, fi0, ft0 = run(list(range(len(genomes))))
for part in ("interleave", "block"):
    ids = parallel.shard_refs(len(genomes), world, rank, part)
    lc, li, lt = run(ids)
    qq, rr = np.nonzero(lc)
    rows = np.zeros(len(qq), CGI_DTYPE)
    rows["qryGenomeId"], rows["refGenomeId"], rows["countSeq"], rows["identity"], rows["totalQueryFragments"] = qq, rr, lc[qq, rr], li[qq, rr], lt[qq]
    rc, ri, rt = parallel.gather_rows(rows, len(queries), len(genomes), world, rank, dist=dist, partition=part)
    assert (rc == fc0).all() and (ri.view(np.uint32) == fi0.view(np.uint32)).all() and (rt == ft0).all(), part
    assert [parallel.global_ref_id(i, world, rank, len(genomes), part) for i in range(len(ids))] == ids
if rank == 0:
    fc, fi, ft = run(list(range(len(genomes))))
    assert (gt == ft).all() and (ft == L // 3000).all(), (gt, ft)
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


EXCHANGE_WORKER = r'''
import ctypes, os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["ROOT"])
from fastani_b200 import parallel

class FakeSketch:
    """Stands in for fastani_b200.QuerySketch: same info()/export_to() surface over a host byte string."""
    def __init__(self, payload): self.payload = payload
    def info(self): return {"export_bytes": len(self.payload)}
    def export_to(self, ptr, cap):
        assert cap >= len(self.payload)
        ctypes.memmove(ptr, self.payload, len(self.payload))

def importer(ctx, ptr, nbytes):
    return FakeSketch(ctypes.string_at(ptr, nbytes))

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(100 + rank)
mine = FakeSketch(rng.integers(0, 256, 1000 + 4097 * rank, dtype=np.uint8).tobytes())        # ragged sizes
got = parallel.exchange_query_sketches(None, mine, world, rank, dist, torch.device("cpu"), importer=importer)
assert len(got) == world and got[rank] is mine
for r in range(world):
    want = np.random.default_rng(100 + r).integers(0, 256, 1000 + 4097 * r, dtype=np.uint8).tobytes()
    assert got[r].payload == want, (rank, r)
# the query split of the bench: rank r owns queries r, r+N, ...; together they cover every query once
nq = 11
owned = torch.zeros(nq, dtype=torch.int64)
owned[list(range(rank, nq, world))] = 1
dist.all_reduce(owned)
assert owned.tolist() == [1] * nq
if rank == 0: print("EXCHANGE_OK")
dist.barrier()
dist.destroy_process_group()
'''


def test_query_sketch_exchange_logic_over_gloo(tmp_path):
    """parallel.exchange_query_sketches (sizes all-reduce, padded all-gather, per-rank slices) with world_size 2 and 3
    on CPU; the sketch objects are stand-ins, the GPU round trip of real ones is tests/test_gpu.py::test_query_sketch_*."""
    script = tmp_path / "xworker.py"
    script.write_text(EXCHANGE_WORKER)
    env = dict(os.environ, ROOT=ROOT, OMP_NUM_THREADS="1")
    for n, port in ((2, "29618"), (3, "29619")):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                            "--master-addr", "127.0.0.1", "--master-port", port, str(script)],
                           capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        assert "EXCHANGE_OK" in r.stdout

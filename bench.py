#!/usr/bin/env python
"""bench.py -- genome-pairs/sec of the many-to-many ANI hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N ...            the reference's own CPU implementation
                                                            (oracle/_ref/fastANI_ref, OpenMP, all host cores)

Workload (config.workload): BASELINE.json configs[2], "Many-to-many: 1000 x 1000 synthetic ~5 Mbp
bacterial genomes, k=16, fragLen=3000": 50 clusters x 20 strains, strain m = cluster ancestor with iid
substitutions at rate 0.6 % * m, one contig per genome, Q = R = the same 1000 genomes.
One "step" = one pass of the hot path over the whole batch: HP1 (index build over this rank's reference
shard) + HP2 (all queries mapped against it) + the per-pair reduction; for N > 1 each rank sketches 1/N of the
queries, the sketches are all-gathered over NCCL, every rank maps all of them against its shard, and the dense
per-pair tables are gathered at the end.  value = pairs / step time with the genomes already packed in HBM; e2e = the same through
the C ABI from pinned HOST buffers (H2D + 2-bit packing inside the timed region, results copied back).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genome-pairs/sec (many-to-many, k=16, fragLen=3000)"
UNIT = "genome-pairs/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--clusters", type=int, default=50)
    ap.add_argument("--strains", type=int, default=20)
    ap.add_argument("--genome-len", type=int, default=5_000_000)
    ap.add_argument("--seed", type=int, default=3)
    ap.add_argument("--ref-sample-refs", type=int, default=0, help="reference arm: references in the bounded sample (0 = one per host core)")
    ap.add_argument("--ref-sample-queries", type=int, default=4, help="reference arm: queries in the bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    n = a.clusters * a.strains
    return "many-to-many %dx%d synthetic %.1f Mbp genomes (%d clusters x %d strains, 0.6%%*m substitutions), k=16 fragLen=3000" % (
        n, n, a.genome_len / 1e6, a.clusters, a.strains)


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                      "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in open(self.path):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}
        return out


# ----------------------------------------------------------------------------------------- reference arm
def write_fasta(path, name, seq):
    with open(path, "wb") as f:
        f.write((">%s\n" % name).encode())
        s = seq.tobytes() if hasattr(seq, "tobytes") else bytes(seq)
        # 80-column FASTA
        n = len(s)
        body = bytearray(n + (n + 79) // 80)
        mv = np.frombuffer(s, np.uint8)
        full = n // 80
        out = np.frombuffer(body, np.uint8)
        if full:
            blk = out[:full * 81].reshape(full, 81)
            blk[:, :80] = mv[:full * 80].reshape(full, 80)
            blk[:, 80] = 10
        rest = n - full * 80
        if rest:
            out[full * 81:full * 81 + rest] = mv[full * 80:]
            out[full * 81 + rest] = 10
        f.write(body)


def miniature(a, cores):
    """Bounded sample of the workload for the CPU arm.  The reference parallelises over REFERENCES only (one OpenMP
    thread per reference shard, core_genome_identity.cpp:55), so the sample has as many references as the box has
    cores (all threads busy, as in the full 1000-reference run) and a few queries, and keeps the full run's related
    fraction: reference i = strain 2*(i // clusters) of cluster i % clusters, query q = strain 1 of cluster q, so
    query q is related to the references i = q (mod clusters): ~R/clusters of R, i.e. 1 in `clusters`, like 20 of 1000."""
    R = max(1, min(a.ref_sample_refs or cores, a.clusters * ((a.strains + 1) // 2)))
    Q = max(1, min(a.ref_sample_queries, a.clusters))
    refs = [(i % a.clusters, 2 * (i // a.clusters)) for i in range(R)]
    qrys = [(q, 1) for q in range(Q)]
    return qrys, refs


def run_reference_cli(a, gen, steps, warmup):
    """Times oracle/_ref/fastANI_ref (the unmodified reference, OpenMP) on the bounded sample; returns
    (pairs_per_s, seconds_per_step, cores, sample description)."""
    cli = os.path.join(ROOT, "oracle", "_ref", "fastANI_ref")
    if not os.path.exists(cli):
        raise RuntimeError("oracle/_ref/fastANI_ref is missing (build it with `make -C oracle` where /root/reference exists)")
    cores = os.cpu_count() or 1
    qrys, refs = miniature(a, cores)
    tmp = tempfile.mkdtemp(prefix="bani_ref_")
    try:
        ql, rl = os.path.join(tmp, "q.txt"), os.path.join(tmp, "r.txt")
        with open(ql, "w") as fq, open(rl, "w") as fr:
            for lst, fh in ((qrys, fq), (refs, fr)):
                for c, s in lst:
                    p = os.path.join(tmp, "c%d_s%d.fna" % (c, s))
                    if not os.path.exists(p):
                        write_fasta(p, "c%d_s%d" % (c, s), gen(c, s))
                    fh.write(p + "\n")
        times = []
        for it in range(warmup + steps):
            t = time.time()
            r = subprocess.run([cli, "--ql", ql, "--rl", rl, "-t", str(cores), "-o", os.path.join(tmp, "out.txt")],
                               capture_output=True, text=True)
            dt = time.time() - t
            if r.returncode != 0:
                raise RuntimeError("fastANI_ref failed: " + r.stderr[-500:])
            if it >= warmup:
                times.append(dt)
        rows = sum(1 for _ in open(os.path.join(tmp, "out.txt")))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    sec = float(np.mean(times))
    nrel = sum(1 for (qc, _) in qrys for (rc, _) in refs if rc == qc)
    sample = ("%d queries x %d references of the workload (%d related pairs of %d; full run: 1 in %d), "
              "%.1f Mbp genomes, fastANI_ref -t %d (one thread per reference), %d output rows, FASTA on local disk"
              % (len(qrys), len(refs), nrel, len(qrys) * len(refs), a.clusters, a.genome_len / 1e6, cores, rows))
    return len(qrys) * len(refs) / sec, sec, cores, sample


def numpy_gen(a):
    from fastani_b200.synth import synth_genome
    return lambda c, s: synth_genome(a.seed, c + 1, s, 6000 * s, a.genome_len)


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    gen = numpy_gen(a)
    try:
        import fastani_b200 as fb
        ctx = fb.Context(fb.Parameters())
        gen = lambda c, s: ctx.synth_genome(a.seed, c + 1, s, 6000 * s, a.genome_len)     # same bytes, generated faster
    except Exception:
        pass
    value, sec, cores, sample = run_reference_cli(a, gen, a.steps, max(a.warmup, 0))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "timed": "bounded sample per step: " + sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------- our arm
def main_ours(a):
    import torch
    import fastani_b200 as fb
    from fastani_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)

    ctx = fb.Context(fb.Parameters(), device=local)
    nG = a.clusters * a.strains
    L = a.genome_len
    # ---- work split (N > 1): rank r owns the references r, r+N, ... (splitReferenceGenomes) AND sketches the queries
    #      r, r+N, ...; the query sketches are exchanged over NCCL, every rank maps all of them against its own shard.
    #      With Q = R the two sets coincide, so every genome is uploaded to exactly one GPU.
    my_refs = parallel.shard_refs(nG, world, rank)
    my_qrys = list(range(rank, nG, world))
    need = sorted(set(my_refs) | set(my_qrys))
    slot = {g: i for i, g in enumerate(need)}
    # ---- synthetic genomes into ONE pinned host buffer (what a FASTA reader would fill)
    host = ctx.pinned(len(need) * L)
    for i, g in enumerate(need):
        c, s = divmod(g, a.strains)
        ctx.synth_genome(a.seed, c + 1, s, 6000 * s, L, out=host[i * L:(i + 1) * L])
    off = np.arange(len(need) + 1, dtype=np.int64) * L
    gen_off = np.arange(len(need) + 1, dtype=np.int32)

    stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)
        ctx.sync()

    last = {}

    def step(genomes=None):
        """one pass of the hot path; genomes=None => end to end from the pinned host buffer"""
        own = genomes is None
        dbg = os.environ.get("BENCH_DEBUG") and rank == 0
        tt = [time.time()]
        if own:
            genomes = ctx.genomes_from_buffer(host, off, gen_off)
        tt.append(time.time())
        sk = fb.Sketch(ctx, [genomes[slot[i]] for i in my_refs])
        tt.append(time.time())
        if world == 1:
            res, tot, ctr = fb.compute_cgi(ctx, sk, genomes)
            d2h = res.nbytes + tot.nbytes
        else:
            mine = fb.QuerySketch(ctx, [genomes[slot[i]] for i in my_qrys], my_qrys, hint=sk)
            sketches = parallel.exchange_query_sketches(ctx, mine, world, rank, dist, device)
            res, ctr = fb.compute_cgi_sketched(ctx, sk, sketches)
            d2h = res.nbytes
            for q in sketches:
                q.close()
        tt.append(time.time())
        cnt, idn = parallel.dense_tables(res, nG, len(my_refs))
        gc, gi = parallel.gather_tables(cnt, idn, nG, world, rank, dist=dist, device=device)
        tt.append(time.time())
        last.update(cnt=gc, idn=gi, ctr=ctr.as_dict(), stats=sk.stats(), d2h=d2h)
        sk.close()
        if own:
            for g in genomes:
                g.close()
        if dbg:
            tt.append(time.time())
            sys.stderr.write("step[%s] upload %.1f index %.1f map %.1f gather %.1f close %.1f ms\n" % (
                "e2e" if own else "res", *[(tt[i + 1] - tt[i]) * 1e3 for i in range(5)]))

    def timed(n, genomes):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record(stream)
        for _ in range(n):
            step(genomes)
        e1.record(stream)
        barrier()
        wall = time.time() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # device time between the two events on the library's stream (the step has host sync points
        # inside, so it agrees with the wall clock to ~1 %); wall clock only if events are unusable
        t = torch.tensor([ms / 1e3 if ms > 0 else wall], dtype=torch.float64, device=device)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / n

    # ---- warm-up (end-to-end steps: also warms the memory pool and the statistic tables)
    for _ in range(max(a.warmup, 0)):
        step(None)
    resident = ctx.genomes_from_buffer(host, off, gen_off)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    ctx.profile_read()
    l0 = ctx.launch_count()
    sec_res = timed(a.steps, resident)
    launches = (ctx.launch_count() - l0) // max(a.steps, 1)
    prof = ctx.profile_read()
    ctx.profile(False)
    for g in resident:          # hand the resident genomes' device blocks back to the library's allocator: the end-to-end
        g.close()               # steps re-create them from the pinned host buffer (no cudaMalloc inside the timed region)
    step(None)                  # untimed: first end-to-end step after the switch
    sec_e2e = timed(a.steps, None)
    clocks = sampler.stop() if rank == 0 else {}

    pairs = float(nG) * float(nG)
    h2d = torch.tensor([float(len(need) * L)], dtype=torch.float64, device=device)        # bytes this rank uploads per e2e step
    if dist is not None:
        dist.all_reduce(h2d)
    h2d_total = float(h2d.item())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- sanity of the result the timing produced (not a parity test, those live in tests/)
    cnt, idn = last["cnt"], last["idn"]
    diag_ok = bool((np.diag(idn) > 99.9).all()) and int((cnt > 0).sum()) >= nG
    ctr = last["ctr"]

    # ---- roofline of the dominant stage: algorithmic bytes / CUDA-event duration
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    st = last["stats"]
    stages = {k: {"ms": v[0] / max(a.steps, 1), "launches": v[2] // max(a.steps, 1),
                  "algo_GB": v[1] / max(a.steps, 1) / 1e9} for k, v in prof.items()}
    # stage timers -> kernels (the two sketch stages run the same kernel); algorithmic bytes per stage are set
    # next to the launches in csrc/ (DESIGN.md section 4 lists the formulas)
    KERNEL_OF = {"ref_sketch": "sketch_kernel", "q_sketch": "sketch_kernel", "l2_events": "l2_events_kernel",
                 "l2_seq": "l2_seq_kernel", "frag_l1": "frag_l1_kernel", "lookup": "lookup_kernel",
                 "q_sort_unique": "sort_unique_kernel", "l2_bounds": "l2_bounds_kernel"}
    kernels = {}
    for k, v in stages.items():
        kk = kernels.setdefault(KERNEL_OF.get(k, k), {"ms": 0.0, "launches": 0, "algo_GB": 0.0})
        kk["ms"] += v["ms"]; kk["launches"] += v["launches"]; kk["algo_GB"] += v["algo_GB"]
    dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
    roof = None
    if dom:
        kd = kernels[dom]
        ach = kd["algo_GB"] / (kd["ms"] / 1e3) if kd["ms"] > 0 else 0.0
        traffic, traffic_src = None, None
        try:        # DRAM bytes per algorithmic byte of this kernel, from the committed ncu --set full capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_dram_traffic.json")))
            if dom in tj["kernels"]:
                traffic = tj["kernels"][dom]["dram_bytes_per_algorithmic_byte"] * kd["algo_GB"] * 1e9 / max(kd["launches"], 1)
                traffic_src = tj["kernels"][dom]["source"]
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "launches_per_step": kd["launches"], "avg_launch_ms": kd["ms"] / max(kd["launches"], 1),
                "algorithmic_bytes_per_launch": kd["algo_GB"] * 1e9 / max(kd["launches"], 1),
                "share_of_step": kd["ms"] / (sec_res * 1e3),
                "note": "integer hash/compare kernels: the limiter is the INT32 ALU / L1TEX pipe, not HBM (DESIGN.md section 6)"}
    kern_tab = {k: {"ms": round(v["ms"], 3), "launches": v["launches"], "GBps": round(v["algo_GB"] / (v["ms"] / 1e3), 1) if v["ms"] > 0 else 0.0}
                for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms"])}

    line = {
        "metric": METRIC, "value": pairs / sec_res, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": sec_res * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "k": 16, "frag_len": 3000, "window": ctx.windowSize,
                   "queries": nG, "references": nG, "parallelism": ("reference list sharded round-robin over %d GPU(s)" % world) + ("; each rank sketches 1/%d of the queries, sketches all-gathered over NCCL" % world if world > 1 else ""),
                   "l2_flush": "inputs (%.1f GB packed genomes + %.1f GB index per rank) exceed the 126 MB L2" % (
                       nG * L / 4e9, 16.0 * st["n_minimizers"] / 1e9),
                   "result_check": "self pairs > 99.9%% and >= %d populated pairs: %s (min self identity %.4f)" % (nG, diag_ok, float(np.diag(idn).min())),
                   "counters_rank0": ctr},
        "clocks": clocks,
        "e2e": {"value": pairs / sec_e2e, "unit": UNIT, "ms_per_step": sec_e2e * 1e3,
                "h2d_bytes_per_step": int(h2d_total), "d2h_bytes_per_step": int(last["d2h"])},
        "gpu_launches": int(launches),
        "roofline": roof,
        "kernels": kern_tab,
        "stages": stages,
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            gen = lambda c, s: ctx.synth_genome(a.seed, c + 1, s, 6000 * s, L)
            v, sec, cores, sample = run_reference_cli(a, gen, 1, 0)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample}
        except Exception as e:          # the checker being absent must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
                                    "sample": "unavailable: %s" % e}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse_args()
    sys.exit(main_reference(args) if args.impl == "reference" else main_ours(args))

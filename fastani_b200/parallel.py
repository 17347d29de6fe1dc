"""Multi-GPU layout of a many-to-many run: one process per GPU, the reference list cut into one shard per GPU.

This is the reference's threading rule lifted to GPUs (src/cgi/include/computeCoreIdentity.hpp:457-487): every shard
maps ALL queries against its own index, results of different shards are disjoint, and a local reference id becomes
global again afterwards.  The reference deals references round-robin (shard g owns j with j % G == g, global id =
local * G + g: partition "interleave"); contiguous blocks of the list (partition "block") give the same results and keep
list neighbours on one GPU.  Two exchanges: (1) the query sketches -- rank r sketches the
queries r, r+G, ... once and the sorted fragment sketches (~0.33 B per query base) are all-gathered over NCCL,
so the query-side work is not repeated on every rank; (2) the final gather of the dense per-pair tables
(count int32, identity float32).
"""
import numpy as np


def shard_refs(n_refs, world, rank, partition="interleave"):
    """Indices of the references of shard `rank`.
    "interleave": splitReferenceGenomes (computeCoreIdentity.hpp:457-474), reference j -> shard j % world.
    "block": contiguous ranges of the list.  Per-pair results do not depend on the partition (the
    shard-invariance tests), so a run is free to choose: a block partition keeps genomes that are neighbours in the list
    on one GPU -- lists ordered by taxon, as directory listings usually are, then leave each query fragment with
    candidates in ONE shard instead of a few in every shard -- while interleaving balances the load for any order."""
    if partition == "block":
        a = (n_refs * rank) // world
        b = (n_refs * (rank + 1)) // world
        return list(range(a, b))
    return list(range(rank, n_refs, world))


def global_ref_id(local_id, world, rank, n_refs=None, partition="interleave"):
    """correctRefGenomeIds (computeCoreIdentity.hpp:480-487); for a block partition: offset of the block."""
    if partition == "block":
        return (n_refs * rank) // world + local_id
    return local_id * world + rank


def dense_tables(cgi_results, n_queries, n_local_refs, total_fragments=None):
    """CGI result rows of one shard -> dense [n_queries, n_local_refs] (count, identity) tables and the per-query
    totalQueryFragments (taken from `total_fragments` when given, else from the rows; 0 = unknown on this shard)."""
    cnt = np.zeros((n_queries, max(n_local_refs, 0)), np.int32)
    idn = np.zeros((n_queries, max(n_local_refs, 0)), np.float32)
    tot = np.zeros(n_queries, np.int32)
    if len(cgi_results):
        q = cgi_results["qryGenomeId"]; r = cgi_results["refGenomeId"]
        cnt[q, r] = cgi_results["countSeq"]
        idn[q, r] = cgi_results["identity"]
        tot[q] = cgi_results["totalQueryFragments"]
    if total_fragments is not None:
        tot[:len(total_fragments)] = np.asarray(total_fragments, np.int64).astype(np.int32)
    return cnt, idn, tot


def merge_shards(tables, n_queries, n_refs, world, partition="interleave"):
    """tables[g] = (count, identity[, totals]) of shard g, shapes [n_queries, len(shard_refs(n_refs, world, g))].
    Returns the global [n_queries, n_refs] tables (and the element-wise maximum of the totals when present)."""
    cnt = np.zeros((n_queries, n_refs), np.int32)
    idn = np.zeros((n_queries, n_refs), np.float32)
    tot = None
    for g, t in enumerate(tables):
        c, i = t[0], t[1]
        if partition == "block":
            a, b = (n_refs * g) // world, (n_refs * (g + 1)) // world
            if b > a:
                cnt[:, a:b] = c[:, :b - a]
                idn[:, a:b] = i[:, :b - a]
        else:
            ncols = len(range(g, n_refs, world))             # shard g owns the columns g, g + world, ... (shard_refs)
            if ncols:
                cnt[:, g::world] = c[:, :ncols]
                idn[:, g::world] = i[:, :ncols]
        if len(t) > 2 and t[2] is not None:
            tot = t[2].copy() if tot is None else np.maximum(tot, t[2])
    return (cnt, idn) if tot is None else (cnt, idn, tot)


def exchange_rows(cgi_results, world, rank, dist=None, device=None):
    """The result exchange of a multi-GPU run on the COMPACT rows: every rank contributes its cgi::CGI_Results rows
    (20 bytes each; a few thousand per rank instead of a dense n_queries x n_refs / world table) to one padded all-gather.
    Returns the per-rank row arrays (reference ids still shard-local), on every rank."""
    rows = np.ascontiguousarray(cgi_results)
    if world == 1 or dist is None:
        return [rows]
    import torch
    n = torch.tensor([len(rows)], dtype=torch.int64, device=device if device is not None else "cpu")
    nall = torch.empty(world, dtype=torch.int64, device=n.device)
    dist.all_gather_into_tensor(nall, n)
    ns = [int(x) for x in nall.cpu()]
    width = max(max(ns), 1)
    buf = np.zeros((width, rows.dtype.itemsize // 4), np.int32)
    buf[:len(rows)] = rows.view(np.int32).reshape(len(rows), -1)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * width, buf.shape[1]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    o = out.cpu().numpy().reshape(world, width, buf.shape[1])
    return [np.ascontiguousarray(o[g, :ns[g]]).view(rows.dtype).reshape(-1) for g in range(world)]


def rows_to_tables(parts, n_queries, n_refs, world, total_fragments=None, partition="interleave", out=None):
    """Per-rank result rows (exchange_rows) -> dense (count, identity, totalQueryFragments) tables with global reference
    ids (correctRefGenomeIds).  `out`: tables to reuse."""
    if out is not None:
        cnt, idn, tot = out
        cnt.fill(0); idn.fill(0); tot.fill(0)
    else:
        cnt = np.zeros((n_queries, n_refs), np.int32)
        idn = np.zeros((n_queries, n_refs), np.float32)
        tot = np.zeros(n_queries, np.int32)
    if total_fragments is not None:
        tot[:len(total_fragments)] = np.asarray(total_fragments, np.int64).astype(np.int32)
    cflat, iflat = cnt.reshape(-1), idn.reshape(-1)
    for g, r in enumerate(parts):
        if len(r):
            w = np.ascontiguousarray(r).view(np.int32).reshape(len(r), -1)   # refGenomeId qryGenomeId countSeq totalQueryFragments identity-bits
            q = w[:, 1].astype(np.int64)
            flat = q * n_refs + global_ref_id(w[:, 0].astype(np.int64), world, g, n_refs, partition)
            cflat[flat] = w[:, 2]
            iflat[flat] = w[:, 4].view(np.float32)
            np.maximum.at(tot, q, w[:, 3])
    return cnt, idn, tot


def gather_rows(cgi_results, n_queries, n_refs, world, rank, dist=None, device=None, total_fragments=None, partition="interleave", out=None):
    """exchange_rows + rows_to_tables: (count, identity, totals) with global ids on every rank."""
    parts = exchange_rows(cgi_results, world, rank, dist, device)
    return rows_to_tables(parts, n_queries, n_refs, world, total_fragments, partition, out)


def gather_tables(cnt_local, idn_local, n_refs, world, rank, dist=None, device=None, tot=None, partition="interleave"):
    """All-gather of the per-shard tables over torch.distributed (NCCL on GPUs, gloo on CPU): shards are padded to the
    largest shard and count / identity bits / totals travel as ONE int32 tensor in one collective.
    Returns (count, identity) or, when `tot` is given, (count, identity, totals)."""
    n_queries = cnt_local.shape[0]
    if world == 1 or dist is None:
        return merge_shards([(cnt_local, idn_local, tot)], n_queries, n_refs, 1, partition)
    import torch
    width = (n_refs + world - 1) // world
    pack = np.zeros((n_queries, 2 * width + 1), np.int32)
    pack[:, :cnt_local.shape[1]] = cnt_local
    pack[:, width:width + idn_local.shape[1]] = idn_local.view(np.int32)
    if tot is not None:
        pack[:, 2 * width] = tot
    t = torch.from_numpy(pack)
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)      # rank-major concatenation
    dist.all_gather_into_tensor(out, t)
    o = out.cpu().numpy().reshape(world, t.shape[0], t.shape[1])
    tables = [(o[g][:, :width], o[g][:, width:2 * width].view(np.float32), o[g][:, 2 * width] if tot is not None else None) for g in range(world)]
    return merge_shards(tables, n_queries, n_refs, world, partition)


class SketchExchange:
    """An all-gather of the ranks' QuerySketch objects in flight (start_exchange ... finish)."""

    def __init__(self, ctx, mine, world, rank, dist, device, importer=None):
        import torch
        if importer is None:
            from .api import QuerySketch
            importer = QuerySketch.from_device_buffer
        self.ctx, self.mine, self.world, self.rank, self.device, self.importer = ctx, mine, world, rank, device, importer
        nbytes = mine.info()["export_bytes"]
        sizes = torch.zeros(world, dtype=torch.int64, device=device)
        sizes[rank] = nbytes
        dist.all_reduce(sizes)
        self.sizes = [int(x) for x in sizes.cpu()]
        self.width = (max(self.sizes) + 255) // 256 * 256
        self.send = torch.empty(self.width, dtype=torch.uint8, device=device)
        mine.export_to(self.send.data_ptr(), self.width)    # synchronises the library's stream: the bytes are in place
        self.recv = torch.empty(world * self.width, dtype=torch.uint8, device=device)
        self.work = dist.all_gather_into_tensor(self.recv, self.send, async_op=True)

    def finish(self):
        """Waits for the collective and rebuilds the peers' sketches on this device (bani_qsketch_import).
        Returns the `world` sketches in rank order (this rank's own object included)."""
        import torch
        self.work.wait()
        if getattr(self.device, "type", str(self.device)) != "cpu":
            torch.cuda.synchronize(self.device)
        out = []
        for r in range(self.world):
            out.append(self.mine if r == self.rank else self.importer(self.ctx, self.recv.data_ptr() + r * self.width, self.sizes[r]))
        return out


def start_exchange(ctx, mine, world, rank, dist, device, importer=None):
    """Export this rank's QuerySketch into a flat device buffer and START the padded all-gather (NCCL over NVLink on GPUs);
    the caller maps its own sketch against its shard while the peers' sketches travel, then calls .finish()."""
    return SketchExchange(ctx, mine, world, rank, dist, device, importer)


def exchange_query_sketches(ctx, mine, world, rank, dist, device, importer=None):
    """All-gather of the ranks' QuerySketch objects: each is packed into a flat device buffer (bani_qsketch_export),
    the buffers travel in one padded NCCL all-gather, and the peers' sketches are rebuilt on this device
    (bani_qsketch_import).  Returns the `world` sketches in rank order (this rank's own object included).
    `importer(ctx, ptr, nbytes)` defaults to QuerySketch.from_device_buffer (the CPU/gloo test passes a stand-in)."""
    return start_exchange(ctx, mine, world, rank, dist, device, importer).finish()

"""Multi-GPU layout of a many-to-many run: one process per GPU, references sharded round-robin.

This is exactly the reference's threading rule lifted to GPUs (src/cgi/include/
computeCoreIdentity.hpp:457-487): shard g owns references j with j % G == g, every shard maps
ALL queries against its own index, results of different shards are disjoint, and a local
reference id becomes global as local * G + g.  Two exchanges: (1) the query sketches -- rank r sketches the
queries r, r+G, ... once and the sorted fragment sketches (~0.33 B per query base) are all-gathered over NCCL,
so the query-side work is not repeated on every rank; (2) the final gather of the dense per-pair tables
(count int32, identity float32).
"""
import numpy as np


def shard_refs(n_refs, world, rank):
    """splitReferenceGenomes (computeCoreIdentity.hpp:457-474): indices of the references of shard `rank`."""
    return list(range(rank, n_refs, world))


def global_ref_id(local_id, world, rank):
    """correctRefGenomeIds (computeCoreIdentity.hpp:480-487)."""
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


def merge_shards(tables, n_queries, n_refs, world):
    """tables[g] = (count, identity[, totals]) of shard g, shapes [n_queries, len(shard_refs(n_refs, world, g))].
    Returns the global [n_queries, n_refs] tables (and the element-wise maximum of the totals when present)."""
    cnt = np.zeros((n_queries, n_refs), np.int32)
    idn = np.zeros((n_queries, n_refs), np.float32)
    tot = None
    for g, t in enumerate(tables):
        c, i = t[0], t[1]
        ncols = len(range(g, n_refs, world))                 # shard g owns the columns g, g + world, ... (shard_refs)
        if ncols:
            cnt[:, g::world] = c[:, :ncols]
            idn[:, g::world] = i[:, :ncols]
        if len(t) > 2 and t[2] is not None:
            tot = t[2].copy() if tot is None else np.maximum(tot, t[2])
    return (cnt, idn) if tot is None else (cnt, idn, tot)


def gather_tables(cnt_local, idn_local, n_refs, world, rank, dist=None, device=None, tot=None):
    """All-gather of the per-shard tables over torch.distributed (NCCL on GPUs, gloo on CPU): shards are padded to the
    largest shard and count / identity bits / totals travel as ONE int32 tensor in one collective.
    Returns (count, identity) or, when `tot` is given, (count, identity, totals)."""
    n_queries = cnt_local.shape[0]
    if world == 1 or dist is None:
        return merge_shards([(cnt_local, idn_local, tot)], n_queries, n_refs, 1)
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
    return merge_shards(tables, n_queries, n_refs, world)


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

"""oracle/pyoracle.py -- TEST INFRASTRUCTURE ONLY (ctypes face of oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
The product package (fastani_b200/) never does.
"""
import ctypes as C
import gzip
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MAPPING_DTYPE = np.dtype([
    ("queryLen", "<i4"), ("refStartPos", "<i4"), ("refEndPos", "<i4"),
    ("queryStartPos", "<i4"), ("queryEndPos", "<i4"), ("refSeqId", "<i4"),
    ("querySeqId", "<i4"), ("nucIdentity", "<f4"), ("nucIdentityUpperBound", "<f4"),
    ("sketchSize", "<i4"), ("conservedSketches", "<i4")])          # skch::MappingResult, 44 B
MINIMIZER_DTYPE = np.dtype([("hash", "<u4"), ("seqId", "<i4"), ("wpos", "<i4")])   # skch::MinimizerInfo, 12 B


class Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("sum_s", "hits", "n2", "mappings", "candidates", "events")]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(_HERE, "ani_oracle.c")):
            subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
        L = C.CDLL(so)
        L.orc_hash.restype = C.c_uint32
        L.orc_hash.argtypes = [C.c_char_p, C.c_int]
        L.orc_upper.argtypes = [C.c_void_p, C.c_int64]
        L.orc_minimizers.restype = C.c_int64
        L.orc_minimizers.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_min_hits_relaxed.restype = C.c_int
        L.orc_min_hits_relaxed.argtypes = [C.c_int, C.c_int, C.c_float]
        L.orc_identity.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_window_size.restype = C.c_int
        L.orc_window_size.argtypes = [C.c_int, C.c_int]
        L.orc_index_new.restype = C.c_void_p
        L.orc_index_new.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_index_unique.restype = C.c_int64
        L.orc_index_unique.argtypes = [C.c_void_p]
        L.orc_index_free.argtypes = [C.c_void_p]
        L.orc_map_genome.restype = C.c_int64
        L.orc_map_genome.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(Counters)]
        L.orc_cgi.restype = C.c_int
        L.orc_cgi.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        _LIB = L
    return _LIB


def read_fasta(path):
    """Minimal FASTA/FASTQ(.gz) reader with kseq semantics (src/common/kseq.h:177-218):
    name = header up to first whitespace; sequence lines concatenated, whitespace dropped."""
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    out, name, chunks, fq_skip = [], None, [], 0
    with op(path, "rb") as f:
        lines = f.read().split(b"\n")
    i = 0
    while i < len(lines):
        ln = lines[i].rstrip(b"\r")
        i += 1
        if not ln:
            continue
        if ln[:1] in (b">", b"@") and (name is None or ln[:1] == b">" or fq_skip == 0):
            if name is not None:
                out.append((name, b"".join(chunks)))
            name = ln[1:].split()[0].decode() if len(ln) > 1 else ""
            chunks = []
        elif ln[:1] == b"+" and name is not None:
            need = sum(len(c) for c in chunks)       # FASTQ quality block: skip same number of chars
            got = 0
            while got < need and i < len(lines):
                got += len(lines[i].rstrip(b"\r"))
                i += 1
        else:
            chunks.append(ln.replace(b" ", b"").replace(b"\t", b""))
    if name is not None:
        out.append((name, b"".join(chunks)))
    return out


def upper(seq: bytes) -> np.ndarray:
    a = np.frombuffer(seq, dtype=np.uint8).copy()
    m = (a > 96) & (a < 123)
    a[m] -= 32
    return a


def orc_hash(kmer: bytes) -> int:
    return lib().orc_hash(kmer, len(kmer))


def minimizers(seq_upper: np.ndarray, k, w, seq_id=0):
    n = len(seq_upper)
    h = np.empty(max(n, 1), np.uint32); s = np.empty(max(n, 1), np.int32); p = np.empty(max(n, 1), np.int32)
    seq_upper = np.ascontiguousarray(seq_upper, dtype=np.uint8)
    m = lib().orc_minimizers(seq_upper.ctypes.data, n, k, w, seq_id, h.ctypes.data, s.ctypes.data, p.ctypes.data)
    out = np.empty(m, MINIMIZER_DTYPE)
    out["hash"], out["seqId"], out["wpos"] = h[:m], s[:m], p[:m]
    return out


def sketch_genomes(genomes, k, w):
    """genomes: list of list of (name, bytes).  Returns (records, seqsByFile, contig_lens):
    Sketch::build, winSketch.hpp:124-176 -- every contig consumes a seqId, even when too short."""
    recs, by_file, lens, sid = [], [], [], 0
    for g in genomes:
        for _, sq in g:
            recs.append(minimizers(upper(sq), k, w, sid))
            lens.append(len(sq))
            sid += 1
        by_file.append(sid)
    rec = np.concatenate(recs) if recs else np.empty(0, MINIMIZER_DTYPE)
    return rec, np.array(by_file, np.int32), np.array(lens, np.int64)


class Index:
    def __init__(self, records):
        self.hash = np.ascontiguousarray(records["hash"]); self.seq = np.ascontiguousarray(records["seqId"])
        self.wpos = np.ascontiguousarray(records["wpos"])
        self.h = lib().orc_index_new(len(records), self.hash.ctypes.data, self.seq.ctypes.data, self.wpos.ctypes.data)

    def unique(self):
        return lib().orc_index_unique(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_index_free(self.h); self.h = None


def map_genome(index, genome, k, w, frag_len, pid=80.0, cap=1 << 22):
    """genome: list of (name, bytes).  Returns (rows[MAPPING_DTYPE], totalQueryFragments, Counters)."""
    seqs = [upper(sq) for _, sq in genome]
    off = np.zeros(len(seqs) + 1, np.int64)
    off[1:] = np.cumsum([len(x) for x in seqs])
    cat = np.concatenate(seqs) if seqs else np.empty(0, np.uint8)
    rows = np.empty(cap, MAPPING_DTYPE)
    tot, ctr = C.c_uint64(0), Counters()
    n = lib().orc_map_genome(index.h, len(seqs), off.ctypes.data, cat.ctypes.data, k, w, frag_len, pid,
                             rows.ctypes.data, cap, C.byref(tot), C.byref(ctr))
    if n < 0:
        raise RuntimeError("oracle row capacity exceeded")
    return rows[:n].copy(), tot.value, ctr


def cgi(rows, seqs_by_file, frag_len, want_visual=False):
    rows = np.ascontiguousarray(rows)
    sbf = np.ascontiguousarray(seqs_by_file, dtype=np.int32)
    ng = len(sbf)
    og = np.empty(ng, np.int32); oc = np.empty(ng, np.int32); oi = np.empty(ng, np.float32)
    n = max(len(rows), 1)
    vr = np.empty(n, np.int32); vq = np.empty(n, np.int32); vs = np.empty(n, np.int32); vi = np.empty(n, np.float32)
    vn = C.c_int64(0)
    m = lib().orc_cgi(rows.ctypes.data, len(rows), sbf.ctypes.data, ng, frag_len, og.ctypes.data, oc.ctypes.data,
                      oi.ctypes.data, vr.ctypes.data, vq.ctypes.data, vs.ctypes.data, vi.ctypes.data, C.byref(vn))
    res = [(int(og[i]), int(oc[i]), np.float32(oi[i])) for i in range(m)]
    if want_visual:
        v = vn.value
        return res, (vr[:v].copy(), vq[:v].copy(), vs[:v].copy(), vi[:v].copy())
    return res

"""FASTA/FASTQ(.gz) ingest on the host (the step before HP1; kseq semantics,
reference src/common/kseq.h:177-218 as used by winSketch.hpp:141-171)."""
import gzip


def read_fasta(path):
    """Returns [(name, sequence_bytes)] in file order.  The name is the header up to the first
    whitespace; sequence lines are concatenated (bytes kept as they are, like kseq) until a line that starts
    with '>', '@' or '+'; after '+', quality lines are skipped until they cover the sequence
    length; CRLF tolerated."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rb") as fh:
        lines = fh.read().split(b"\n")
    out = []
    i, n = 0, len(lines)
    while i < n:
        ln = lines[i].rstrip(b"\r")
        i += 1
        if ln[:1] not in (b">", b"@"):
            continue                          # kseq skips ahead to the next header
        parts = ln[1:].split()
        name = parts[0].decode("latin-1") if parts else ""
        chunks = []
        while i < n:
            s = lines[i].rstrip(b"\r")
            if s[:1] in (b">", b"@", b"+"):
                break
            chunks.append(s)
            i += 1
        seq = b"".join(chunks)
        if i < n and lines[i][:1] == b"+":
            i += 1
            got = 0
            while got < len(seq) and i < n:
                got += len(lines[i].rstrip(b"\r"))
                i += 1
        out.append((name, seq))
    return out

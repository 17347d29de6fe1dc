"""FASTA/FASTQ(.gz) ingest on the host (the step before HP1) with the semantics of the reference's reader
(src/common/kseq.h:177-218 as used by winSketch.hpp:141-171).  Same state machine as host/kseq_reader.hpp, so the
Python host and bin/fastANI report the same contigs for the same file, malformed ones included:
  * a record starts at the next '>' or '@' CHARACTER (not only at a line start); name = header up to the first whitespace
  * sequence = every following line up to a line whose first character is '>', '@' or '+'; empty lines are skipped,
    ONE trailing '\\r' of a line is dropped, all other bytes are kept
  * after a '+' line, quality lines are consumed until they cover the sequence length; a quality string of another
    length ends the file (kseq returns -2 and the reference's callers stop reading): the record is dropped
"""
import gzip

_SPACE = b" \t\n\v\f\r"


def read_fasta(path):
    """Returns [(name, sequence_bytes)] in file order."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    opener = gzip.open if magic == b"\x1f\x8b" else open
    with opener(path, "rb") as fh:
        buf = fh.read()
    e = len(buf)
    out = []
    p = 0
    last = 0                                   # kseq's last_char: a header character already consumed
    while True:
        if not last:
            a, b = buf.find(b">", p), buf.find(b"@", p)
            cand = [x for x in (a, b) if x >= 0]
            if not cand:
                break
            p = min(cand) + 1
        # name: up to the first whitespace; the rest of the header line is the comment
        q = p
        while q < e and buf[q] not in _SPACE:
            q += 1
        if q == p and q >= e:
            break                              # no name at end of file
        name = buf[p:q].decode("latin-1")
        p = q
        if p < e and buf[p] != 10:
            nl = buf.find(b"\n", p)
            p = e if nl < 0 else nl
        if p < e:
            p += 1                             # the newline
        chunks, n = [], 0
        stop = -1
        while p < e:
            ch = buf[p]
            p += 1
            if ch in (62, 43, 64):             # '>', '+', '@'
                stop = ch
                break
            if ch == 10:
                continue
            nl = buf.find(b"\n", p)
            eol = e if nl < 0 else nl
            line = buf[p - 1:eol]
            if n + len(line) > 1 and line.endswith(b"\r"):
                line = line[:-1]
            chunks.append(line)
            n += len(line)
            p = eol + 1 if eol < e else e
        last = stop if stop in (62, 64) else 0
        seq = b"".join(chunks)
        if stop != 43:
            out.append((name, seq))
            if stop < 0:
                break
            continue
        # FASTQ: skip the '+' line, then quality lines until they cover the sequence
        nl = buf.find(b"\n", p)
        if nl < 0:
            break                              # kseq: -2, no quality string -> the callers stop here
        p = nl + 1
        ql, any_line = 0, False
        while p < e and (not any_line or ql < len(seq)):
            nl = buf.find(b"\n", p)
            eol = e if nl < 0 else nl
            add = eol - p
            if ql + add > 1 and add > 0 and buf[eol - 1] == 13:
                add -= 1
            ql += add
            any_line = True
            p = eol + 1 if eol < e else e
        if ql != len(seq):
            break                              # kseq: -2 -> record dropped, reading stops
        out.append((name, seq))
        last = 0
    return out

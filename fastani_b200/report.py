"""Host glue after the hot path: the output filter and text format of cgi::outputCGI
(src/cgi/include/computeCoreIdentity.hpp:307-343) and computeGenomeLengths (:48-92)."""
import numpy as np


def genome_length(contig_lens, frag_len):
    """computeCoreIdentity.hpp:57-61: sum over contigs >= fragLen of floor(len/fragLen)*fragLen."""
    return int(sum((int(L) // frag_len) * frag_len for L in contig_lens if L >= frag_len))


def fmt_float(x):
    """std::ostream << float with default precision (6 significant digits, %g)."""
    return "%g" % float(np.float32(x))


def output_lines(results, query_names, ref_names, query_lens, ref_lens, frag_len, min_fraction=0.2):
    """results: iterable of (qryGenomeId, refGenomeId, countSeq, totalQueryFragments, identity).
    Ordered as outputCGI: query ascending, identity descending (cgid_types.hpp:76-79)."""
    rows = sorted(results, key=lambda r: (r[0], -float(r[4])))
    out = []
    for q, r, cnt, tot, idn in rows:
        min_len = min(query_lens[q], ref_lens[r])
        shared = cnt * frag_len
        if shared >= np.float32(min_len) * np.float32(min_fraction):        # uint64 * float -> float (:326-332)
            out.append("%s\t%s\t%s\t%d\t%d" % (query_names[q], ref_names[r], fmt_float(idn), cnt, tot))
    return out

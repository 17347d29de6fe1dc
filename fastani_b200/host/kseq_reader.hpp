// kseq_reader.hpp -- FASTA/FASTQ[.gz] ingest with the semantics of the reference's reader.
//
// The reference reads every genome through kseq (src/common/kseq.h:177-218, used by
// winSketch.hpp:141-171, computeMap.hpp:121-189 and computeCoreIdentity.hpp:48-92).  This is an
// independent reader with the same observable behaviour, written for whole-file buffers (the file is
// inflated once with zlib, which also passes plain files through, like gzopen in the reference):
//   * a record starts at the next '>' or '@'; the name is the header up to the first whitespace
//   * sequence = every following line up to a line whose FIRST character is '>', '@' or '+';
//     empty lines are skipped, a trailing '\r' of a line is dropped, all other bytes are kept
//   * after a '+' line, quality lines are consumed until they cover the sequence length; a quality
//     string of a different length ends the file (kseq returns -2 and the callers stop reading)
#pragma once
#include <zlib.h>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace bani_host {

struct Contig { std::string name; uint64_t off = 0; uint64_t len = 0; };   // bytes [off, off+len) of HostGenome::seq

struct HostGenome {
  std::string path;
  std::vector<Contig> contigs;
  std::vector<uint8_t> seq;          // all contigs back to back, bytes exactly as kseq yields them (released once packed)
  // 2-bit packed form (filled by skch::pack_genome in the reader thread): what crosses PCIe
  std::vector<uint32_t> words;       // every contig starts on a multiple of 4 words
  std::vector<int64_t> wordOff;      // per contig
  std::vector<uint32_t> excPos; std::vector<uint8_t> excByte;
  std::vector<int64_t> excOff;       // per contig + 1
  bool packed = false;
};

inline std::vector<uint8_t> inflate_file(const std::string &path)
{
  gzFile fp = gzopen(path.c_str(), "r");
  if (!fp) throw std::runtime_error("could not open " + path);
  gzbuffer(fp, 1 << 20);
  std::vector<uint8_t> out;
  size_t cap = 1 << 22;
  out.resize(cap);
  size_t n = 0;
  for (;;) {
    if (n == cap) { cap *= 2; out.resize(cap); }
    int r = gzread(fp, out.data() + n, (unsigned)std::min<size_t>(cap - n, 1u << 30));
    if (r < 0) { gzclose(fp); throw std::runtime_error("read error in " + path); }
    if (r == 0) break;
    n += (size_t)r;
  }
  gzclose(fp);
  out.resize(n);
  return out;
}

// Appends the line [p, eol) to seq the way ks_getuntil2(KS_SEP_LINE, append) does: bytes as they are,
// then one trailing '\r' removed if the accumulated string is longer than one byte (kseq.h:140).
inline void append_line(std::vector<uint8_t> &seq, size_t recStart, const uint8_t *p, const uint8_t *eol)
{
  seq.insert(seq.end(), p, eol);
  if (seq.size() - recStart > 1 && seq.back() == '\r') seq.pop_back();
}

inline HostGenome read_genome(const std::string &path)
{
  HostGenome g; g.path = path;
  const std::vector<uint8_t> buf = inflate_file(path);
  const uint8_t *b = buf.data(), *e = b + buf.size();
  g.seq.reserve(buf.size());
  const uint8_t *p = b;
  int last = 0;                                   // kseq's last_char: a header character already consumed
  while (true) {
    if (!last) { while (p < e && *p != '>' && *p != '@') p++; if (p >= e) break; last = *p++; }
    // name: up to the first whitespace; the rest of the header line is the comment
    const uint8_t *q = p;
    while (q < e && !isspace(*q)) q++;
    if (q == p && q >= e) break;                  // ks_getuntil returns -1: no name at end of file
    Contig c; c.name.assign((const char *)p, (size_t)(q - p));
    p = q;
    if (p < e && *p != '\n') { while (p < e && *p != '\n') p++; }
    if (p < e) p++;                               // the newline
    c.off = g.seq.size();
    int stop = -1;
    while (p < e) {
      const int ch = *p++;
      if (ch == '>' || ch == '+' || ch == '@') { stop = ch; break; }
      if (ch == '\n') continue;
      g.seq.push_back((uint8_t)ch);
      const uint8_t *eol = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
      if (!eol) eol = e;
      append_line(g.seq, c.off, p, eol);
      p = eol < e ? eol + 1 : e;
    }
    last = (stop == '>' || stop == '@') ? stop : 0;
    c.len = g.seq.size() - c.off;
    if (stop != '+') { g.contigs.push_back(c); if (stop < 0) break; continue; }
    // FASTQ: skip the '+' line, then quality lines until they cover the sequence
    while (p < e && *p != '\n') p++;
    if (p >= e) break;                            // kseq: -2, no quality string -> the callers stop here
    p++;
    uint64_t ql = 0; bool any = false;
    while (p < e && (!any || ql < c.len)) {
      const uint8_t *eol = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
      if (!eol) eol = e;
      uint64_t add = (uint64_t)(eol - p);
      if (ql + add > 1 && add > 0 && eol[-1] == '\r') add--;
      ql += add; any = true;
      p = eol < e ? eol + 1 : e;
    }
    if (ql != c.len) break;                       // kseq: -2 -> record dropped, reading stops
    g.contigs.push_back(c);
    last = 0;
  }
  return g;
}

} // namespace bani_host

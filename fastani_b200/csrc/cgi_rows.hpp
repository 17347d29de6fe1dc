// cgi_rows.hpp -- dense (count, identity) tables of a piece -> cgi::CGI_Results rows (computeCoreIdentity.hpp:267-297 emits
// one row per (query, genome) pair with mappings, in query-major, genome-minor order).  Host-only, no CUDA: the tables
// are sparse (a query matches a few tens of the genomes), so zero counts are skipped four at a time.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../include/fastani_b200.h"

namespace bani {

// counts >= 0.  qryGenomeId = queryId[q], totalQueryFragments = (int) totalFragments[q] (cgid_types.hpp:73)
inline void append_cgi_rows(const int32_t *count, const float *ident, int nQ, int nG, const int32_t *queryId,
                            const uint64_t *totalFragments, std::vector<bani_cgi_result> &out)
{
  for (int q = 0; q < nQ; q++) {
    const int32_t *cr = count + (size_t)q * nG;
    const float *ir = ident + (size_t)q * nG;
    int g = 0;
    while (g < nG) {
      if (g + 4 <= nG) {
        uint64_t a, b;
        memcpy(&a, cr + g, 8); memcpy(&b, cr + g + 2, 8);
        if ((a | b) == 0) { g += 4; continue; }
      }
      const int32_t cnt = cr[g];
      if (cnt > 0) {
        bani_cgi_result r;
        r.refGenomeId = g; r.qryGenomeId = queryId[q]; r.countSeq = cnt;
        r.totalQueryFragments = (int32_t)totalFragments[q];
        r.identity = ir[g];
        out.push_back(r);
      }
      g++;
    }
  }
}

} // namespace bani

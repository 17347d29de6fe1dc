/* oracle/ref_dump.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A small driver that includes the UNMODIFIED reference headers where they lie
 * (/root/reference/src, never copied) and dumps the intermediate values of the
 * hot path so that the C restatement in oracle/ani_oracle.c and the golden
 * fixtures in tests/golden/ can be pinned against the real thing:
 *
 *   ref_dump hash  <kmer>                    -> skch::CommonFunc::getHash
 *   ref_dump wsize <k> <fragLen>             -> skch::Stat::recommendedWindowSize
 *   ref_dump stats <s> <k>                   -> minHitsRelaxed + per-x identity / upper bound
 *   ref_dump sketch <k> <w> <out.bin> <fa..> -> addMinimizers over every contig (12-byte records)
 *   ref_dump map <k> <fragLen> <out.bin> <query.fa> <ref.fa..>
 *                                            -> skch::Sketch + skch::Map, 44-byte MappingResult records
 *
 * Built by oracle/Makefile into oracle/_ref/ref_dump (git-ignored).
 */
#include <iostream>
#include <fstream>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <vector>
#include <string>
#include <omp.h>

#include "map/include/map_parameters.hpp"
#include "map/include/base_types.hpp"
#include "map/include/winSketch.hpp"
#include "map/include/map_stats.hpp"
#include "map/include/computeMap.hpp"
#include "map/include/commonFunc.hpp"

static void fill_params(skch::Parameters &p, int k, int fragLen)
{
  /* defaults of parseandSave, /root/reference/src/map/include/parseCmdArgs.hpp:118-130 */
  p.kmerSize = k;
  p.minReadLength = fragLen;
  p.alphabetSize = 4;
  p.minFraction = 0.2;
  p.threads = 1;
  p.p_value = 1e-03;
  p.percentageIdentity = 80;
  p.visualize = false;
  p.matrixOutput = false;
  p.referenceSize = 5000000;
  p.maxRatioDiff = 100.0;
  p.reportAll = true;
  p.sanityCheck = false;
  p.outFileName = "/dev/null";
  p.windowSize = skch::Stat::recommendedWindowSize(p.p_value, p.kmerSize, p.alphabetSize,
      p.percentageIdentity, p.minReadLength, p.referenceSize);
}

int main(int argc, char **argv)
{
  if (argc < 2) { fprintf(stderr, "usage: ref_dump hash|wsize|stats|contigs|sketch|map ...\n"); return 2; }
  std::string mode = argv[1];

  if (mode == "hash" && argc == 3) {
    printf("%u\n", skch::CommonFunc::getHash(argv[2], (int)strlen(argv[2])));
    return 0;
  }
  if (mode == "wsize" && argc == 4) {
    skch::Parameters p; fill_params(p, atoi(argv[2]), atoi(argv[3]));
    printf("%d\n", p.windowSize);
    return 0;
  }
  if (mode == "stats" && argc == 4) {
    int s = atoi(argv[2]), k = atoi(argv[3]);
    printf("%d\n", skch::Stat::estimateMinimumHitsRelaxed(s, k, 80));
    for (int x = 0; x <= s; x++) {
      /* exactly the expressions of doL2Mapping, computeMap.hpp:375-381 */
      float mash_dist = skch::Stat::j2md(1.0 * x / s, k);
      float lb = skch::Stat::md_lower_bound(mash_dist, s, k, 0.9);
      float nucIdentity = 100 * (1 - mash_dist);
      float nucIdentityUpperBound = 100 * (1 - lb);
      uint32_t a, b; memcpy(&a, &nucIdentity, 4); memcpy(&b, &nucIdentityUpperBound, 4);
      printf("%d %u %u\n", x, a, b);
    }
    return 0;
  }
  if (mode == "contigs" && argc == 3) {
    /* what the reference's own reader (kseq_read, src/common/kseq.h) yields for a file: name, length, crc32 */
    gzFile fp = gzopen(argv[2], "r");
    if (!fp) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    kseq_t *seq = kseq_init(fp);
    int len;
    while ((len = kseq_read(seq)) >= 0)
      printf("%s\t%d\t%lu\n", seq->name.s, len, (unsigned long)crc32(0L, (const Bytef *)seq->seq.s, (uInt)seq->seq.l));
    kseq_destroy(seq);
    gzclose(fp);
    return 0;
  }
  if (mode == "sketch" && argc >= 6) {
    int k = atoi(argv[2]), w = atoi(argv[3]);
    std::vector<skch::MinimizerInfo> mi;
    skch::seqno_t seqCounter = 0;
    for (int f = 5; f < argc; f++) {
      /* the file loop of Sketch::build, winSketch.hpp:137-171 */
      gzFile fp = gzopen(argv[f], "r");
      if (!fp) { fprintf(stderr, "cannot open %s\n", argv[f]); return 1; }
      kseq_t *seq = kseq_init(fp);
      skch::offset_t len;
      while ((len = kseq_read(seq)) >= 0) {
        if (!(len < w || len < k))
          skch::CommonFunc::addMinimizers(mi, seq, k, w, 4, seqCounter);
        seqCounter++;
      }
      kseq_destroy(seq);
      gzclose(fp);
    }
    FILE *o = fopen(argv[4], "wb");
    fwrite(mi.data(), sizeof(skch::MinimizerInfo), mi.size(), o);
    fclose(o);
    fprintf(stderr, "records=%zu contigs=%d\n", mi.size(), seqCounter);
    return 0;
  }
  if (mode == "map" && argc >= 7) {
    skch::Parameters p; fill_params(p, atoi(argv[2]), atoi(argv[3]));
    p.querySequences.push_back(argv[5]);
    for (int f = 6; f < argc; f++) p.refSequences.push_back(argv[f]);
    skch::Sketch sk(p);
    skch::MappingResultsVector_t res;
    uint64_t totalQueryFragments = 0;
    using namespace std::placeholders;
    auto fn = std::bind(skch::Map::insertL2ResultsToVec, std::ref(res), _1);
    skch::Map mapper(p, sk, totalQueryFragments, 0, fn);
    FILE *o = fopen(argv[4], "wb");
    fwrite(res.data(), sizeof(skch::MappingResult), res.size(), o);
    fclose(o);
    fprintf(stderr, "w=%d mappings=%zu totalQueryFragments=%llu\n", p.windowSize, res.size(),
            (unsigned long long)totalQueryFragments);
    return 0;
  }
  fprintf(stderr, "bad arguments\n");
  return 2;
}

// fastani_main.cpp -- the fastANI command line on top of libfastani_b200.so.
//
// Same options and output files as the reference CLI (src/map/include/parseCmdArgs.hpp:118-256,
// src/cgi/core_genome_identity.cpp:27-167): -q/--query, --ql/--queryList, -r/--ref, --rl/--refList, -o/--output,
// -k/--kmer, --fragLen, --minFraction, --maxRatioDiff, --visualize, --matrix, -t/--threads, -s/--sanityCheck,
// -v/--version, -h/--help.  Differences, all on the host side:
//   * the reference list is split over GPUs (--gpus N, default: every visible device), not over OpenMP threads;
//     -t sets the host threads of the ingest stage.  Results do not depend on the split.
//   * every genome file is read ONCE (parallel inflate + parse), its length for the --minFraction filter comes from
//     the same pass (the reference reads each file again in computeGenomeLengths, computeCoreIdentity.hpp:48-92)
//   * --saveIndex / --loadIndex: the on-disk sketch cache the reference lacks (its only answer to repeated runs is
//     scripts/splitDatabase.sh + README.md:104-106); parameters (k, fragLen, window) are stored and a mismatch is refused
//   * without --visualize the per-pair reduction runs on the device (bani_map_cgi); with it, the mapping rows come
//     back and cgi::computeCGI runs on the host because the .visual file needs the surviving rows themselves
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <sstream>
#include <thread>
#include <unistd.h>
#include "ani_host.hpp"

using namespace skch;
typedef std::chrono::high_resolution_clock Clock;

static void usage(const char *prog, std::ostream &o)
{
  o << "-----------------\n"
       "fastANI (B200-native engine): alignment-free whole-genome Average Nucleotide Identity (ANI)\n"
       "-----------------\n"
       "Example usage:\n"
       "$ " << prog << " -q genome1.fa -r genome2.fa -o output.txt\n"
       "$ " << prog << " -q genome1.fa --rl genome_list.txt -o output.txt\n\n"
       "OPTIONS\n"
       "     -h, --help            print this help page\n"
       "     -r, --ref <value>     reference genome (fasta/fastq)[.gz]\n"
       "     --rl, --refList <value>   a file containing list of reference genome files, one genome per line\n"
       "     -q, --query <value>   query genome (fasta/fastq)[.gz]\n"
       "     --ql, --queryList <value> a file containing list of query genome files, one genome per line\n"
       "     -k, --kmer <value>    kmer size <= 32 [default : 16]\n"
       "     -t, --threads <value> host threads for reading the genome files [default : 1]\n"
       "     --gpus <value>        GPUs to split the reference list over [default : all visible]\n"
       "     --fragLen <value>     fragment length [default : 3,000]\n"
       "     --minFraction <value> minimum fraction of genome that must be shared for trusting ANI [default : 0.2]\n"
       "     --maxRatioDiff <value> maximum difference between (Total Ref. Length/Total Occ. Hashes) and\n"
       "                           (Total Ref. Length/Total No. Hashes) [default : 100.0]\n"
       "     --partition <value>   how the reference list is cut into one shard per GPU: interleave (the reference's round-robin\n"
       "                           deal, default) or block (contiguous ranges: list neighbours stay on one GPU)\n"
       "     --saveIndex <prefix>  write the reference sketches to <prefix>.meta + <prefix>.<shard>of<N>.idx after building them\n"
       "     --loadIndex <prefix>  take the references from a saved index instead of -r/--rl: no reference file is read or\n"
       "                           sketched again; queries that are genomes of the index need no file either (1 GPU)\n"
       "     --visualize           output mappings for visualization (<output>.visual)\n"
       "     --matrix              also output ANI values as lower triangular matrix (<output>.matrix)\n"
       "     -o, --output <value>  output file name\n"
       "     -s, --sanityCheck     run sanity check\n"
       "     -v, --version         show version\n";
}

static std::string trim(const std::string &s)
{
  size_t a = 0, b = s.size();
  while (a < b && isspace((unsigned char)s[a])) a++;
  while (b > a && isspace((unsigned char)s[b - 1])) b--;
  return s.substr(a, b - a);
}

static void parseFileList(const std::string &fileToRead, std::vector<std::string> &fileList)   // parseCmdArgs.hpp:34-52
{
  std::ifstream in(fileToRead);
  if (in.fail()) { std::cerr << "ERROR, skch::parseFileList, Could not open " << fileToRead << "\n"; exit(1); }
  std::string line;
  while (std::getline(in, line)) { line = trim(line); if (!line.empty()) fileList.push_back(line); }
}

static void validateInputFiles(const std::vector<std::string> &q, const std::vector<std::string> &r)   // parseCmdArgs.hpp:59-90
{
  if (q.empty() || r.empty()) { std::cerr << "ERROR, skch::validateInputFiles, Count of query and ref genomes should be non-zero" << std::endl; exit(1); }
  for (const auto *lst : {&q, &r})
    for (const auto &e : *lst) { std::ifstream in(e); if (in.fail()) { std::cerr << "ERROR, skch::validateInputFiles, Could not open " << e << std::endl; exit(1); } }
}

static void parseandSave(int argc, char **argv, Parameters &p)
{
  std::string refName, refList, qryName, qryList;
  bool help = false, version = false;
  auto need = [&](int &i) -> const char * { if (i + 1 >= argc) { usage(argv[0], std::cout); exit(1); } return argv[++i]; };
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-h" || a == "--help") help = true;
    else if (a == "-r" || a == "--ref") refName = need(i);
    else if (a == "--rl" || a == "--refList") refList = need(i);
    else if (a == "-q" || a == "--query") qryName = need(i);
    else if (a == "--ql" || a == "--queryList") qryList = need(i);
    else if (a == "-k" || a == "--kmer") p.kmerSize = atoi(need(i));
    else if (a == "-t" || a == "--threads") p.threads = atoi(need(i));
    else if (a == "--gpus") p.gpus = atoi(need(i));
    else if (a == "--fragLen") p.minReadLength = atoi(need(i));
    else if (a == "--minFraction") p.minFraction = (float)atof(need(i));
    else if (a == "--maxRatioDiff") p.maxRatioDiff = (float)atof(need(i));
    else if (a == "--partition") { const std::string v = need(i); if (v == "block") p.blockPartition = true; else if (v != "interleave") { usage(argv[0], std::cout); exit(1); } }
    else if (a == "--saveIndex") p.saveIndex = need(i);
    else if (a == "--loadIndex") p.loadIndex = need(i);
    else if (a == "--visualize") p.visualize = true;
    else if (a == "--matrix") p.matrixOutput = true;
    else if (a == "-o" || a == "--output") p.outFileName = need(i);
    else if (a == "-s" || a == "--sanityCheck") p.sanityCheck = true;
    else if (a == "-v" || a == "--version") version = true;
    else { usage(argv[0], std::cout); exit(1); }
  }
  if (help) { usage(argv[0], std::cout); exit(0); }
  if (version) { std::cerr << "version 1.33 (" << bani_version() << ")\n\n"; exit(0); }
  if (!p.loadIndex.empty() && (!refName.empty() || !refList.empty())) { std::cerr << "ERROR, --loadIndex replaces -r/--rl: give one of them\n"; exit(1); }
  if (!p.loadIndex.empty() && !p.saveIndex.empty()) { std::cerr << "ERROR, --saveIndex and --loadIndex exclude each other\n"; exit(1); }
  if (refName.empty() && refList.empty() && p.loadIndex.empty()) { std::cerr << "Provide reference file (s)\n"; exit(1); }
  if (qryName.empty() && qryList.empty()) { std::cerr << "Provide query file (s)\n"; exit(1); }
  if (!refName.empty()) p.refSequences.push_back(refName); else if (!refList.empty()) parseFileList(refList, p.refSequences);
  if (!qryName.empty()) p.querySequences.push_back(qryName); else parseFileList(qryList, p.querySequences);
  if (!(p.minFraction >= 0.0f && p.minFraction <= 1.0f)) { std::cerr << "ERROR, --minFraction must lie in [0, 1]\n"; exit(1); }
  if (p.threads < 1) p.threads = 1;
  bani_params c = p.c();
  const int w = bani_recommended_window_size(&c);          // Stat::recommendedWindowSize, parseCmdArgs.hpp:244-247
  if (w < 0) { std::cerr << "ERROR, " << bani_last_error() << "\n"; exit(1); }
  p.windowSize = w;
  std::cerr << ">>>>>>>>>>>>>>>>>>\nReference = [" ;
  for (size_t i = 0; i < p.refSequences.size(); i++) std::cerr << (i ? ", " : "") << p.refSequences[i];
  std::cerr << "]\nQuery = [";
  for (size_t i = 0; i < p.querySequences.size(); i++) std::cerr << (i ? ", " : "") << p.querySequences[i];
  std::cerr << "]\nKmer size = " << p.kmerSize << "\nFragment length = " << p.minReadLength << "\nThreads = " << p.threads
            << "\nANI output file = " << p.outFileName << "\nSanity Check  = " << p.sanityCheck << "\n>>>>>>>>>>>>>>>>>>" << std::endl;
  if (p.loadIndex.empty()) validateInputFiles(p.querySequences, p.refSequences);
}

// ---- metadata of a saved index: what the flat per-shard files (bani_index_save) do not hold -- genome paths and
//      contig names -- plus the parameters, so that a mismatch is reported before any GPU work
struct IndexMeta {
  int k = 0, fragLen = 0, window = 0, shards = 0;
  bool block = false;                                                 // version 2: the shards are contiguous blocks of the list
  std::vector<std::string> refPaths;                                  // global reference order
  std::vector<std::vector<std::string>> contigNames;                  // per shard, seqId order
};
static std::string shardFile(const std::string &prefix, int g, int G) { return prefix + "." + std::to_string(g) + "of" + std::to_string(G) + ".idx"; }

static void writeMeta(const std::string &prefix, const Parameters &p, int G, const std::vector<std::vector<std::string>> &contigNames)
{
  std::ofstream o(prefix + ".meta");
  if (!o) throw std::runtime_error("cannot write " + prefix + ".meta");
  o << "BANI_INDEX_META\t" << (p.blockPartition ? 2 : 1) << "\n" << p.kmerSize << "\t" << p.minReadLength << "\t" << p.windowSize << "\t" << G << "\t" << p.refSequences.size() << "\n";
  for (const auto &r : p.refSequences) o << r << "\n";
  for (int g = 0; g < G; g++) { o << contigNames[g].size() << "\n"; for (const auto &n : contigNames[g]) o << n << "\n"; }
}

static IndexMeta readMeta(const std::string &prefix)
{
  std::ifstream in(prefix + ".meta");
  if (!in) throw std::runtime_error("cannot open " + prefix + ".meta");
  IndexMeta m; std::string tag; int ver = 0; size_t nRefs = 0;
  in >> tag >> ver >> m.k >> m.fragLen >> m.window >> m.shards >> nRefs;
  if (!in || tag != "BANI_INDEX_META" || (ver != 1 && ver != 2) || m.shards < 1) throw std::runtime_error(prefix + ".meta is not an index metadata file");
  m.block = ver == 2;
  std::string line; std::getline(in, line);
  for (size_t i = 0; i < nRefs; i++) { if (!std::getline(in, line)) throw std::runtime_error(prefix + ".meta is truncated"); m.refPaths.push_back(line); }
  m.contigNames.resize(m.shards);
  for (int g = 0; g < m.shards; g++) {
    if (!std::getline(in, line)) throw std::runtime_error(prefix + ".meta is truncated");
    const size_t n = (size_t)std::stoull(line);
    for (size_t i = 0; i < n; i++) { if (!std::getline(in, line)) throw std::runtime_error(prefix + ".meta is truncated"); m.contigNames[g].push_back(line); }
  }
  return m;
}

int main(int argc, char **argv)
{
  if (argc == 3 && std::string(argv[1]) == "--dumpContigs") {       // test hook for the reader: name, length, crc32 per contig
    try {
      const bani_host::HostGenome g = bani_host::read_genome(argv[2]);
      for (const auto &c : g.contigs)
        std::cout << c.name << "\t" << c.len << "\t" << crc32(0L, g.seq.data() + c.off, (uInt)c.len) << "\n";
      return 0;
    } catch (const std::exception &e) { std::cerr << "ERROR, " << e.what() << std::endl; return 1; }
  }
  if (argc == 3 && std::string(argv[1]) == "--selftestWriters") {   // test hook for the writers (no GPU needed): a fixed result set
    Parameters p;
    p.querySequences = {"q/a.fa", "q/b.fa", "r/c.fa"}; p.refSequences = {"r/c.fa", "q/a.fa", "r/d.fa"};
    p.minReadLength = 3000; p.minFraction = 0.2f; p.matrixOutput = true;
    std::unordered_map<std::string, uint64_t> len = {{"q/a.fa", 150000}, {"q/b.fa", 90000}, {"r/c.fa", 150000}, {"r/d.fa", 3000000}};
    std::vector<cgi::CGI_Results> v = {
      {0, 0, 40, 50, 97.75071f}, {1, 0, 50, 50, 100.0f}, {2, 0, 9, 50, 81.5f},          // a -> c, a -> a (self), a -> d (fails minFraction: 27000 < 30000)
      {0, 1, 20, 30, 88.123456f}, {2, 1, 6, 30, 80.0f},                                  // b -> c, b -> d (18000 >= 18000 passes)
      {1, 2, 45, 50, 97.5f}, {0, 2, 50, 50, 100.0f}};                                    // c -> a (averaged with a -> c in the matrix), c -> c (self)
    cgi::outputCGI(p, len, v, argv[2]);
    cgi::outputPhylip(p, len, v, argv[2]);
    return 0;
  }
  const auto tStart = Clock::now();
  Parameters parameters;
  parseandSave(argc, argv, parameters);
  const std::string fileName = parameters.outFileName;
  try {
    const bool loading = !parameters.loadIndex.empty();
    IndexMeta meta;
    if (loading) {
      meta = readMeta(parameters.loadIndex);
      if (meta.k != parameters.kmerSize || meta.fragLen != parameters.minReadLength || meta.window != parameters.windowSize)
        throw std::runtime_error("the saved index was built with k " + std::to_string(meta.k) + " fragLen " + std::to_string(meta.fragLen) + " window " + std::to_string(meta.window) +
                                 ", this run asks for k " + std::to_string(parameters.kmerSize) + " fragLen " + std::to_string(parameters.minReadLength) + " window " + std::to_string(parameters.windowSize));
      parameters.refSequences = meta.refPaths;
      parameters.blockPartition = meta.block;
    }
    // a run that names its GPU count initialises only those devices (the driver's start-up cost grows with every visible GPU)
    {
      const int want = loading ? meta.shards : parameters.gpus;
      if (want > 0 && !getenv("CUDA_VISIBLE_DEVICES")) {
        std::string v; for (int g = 0; g < want; g++) v += (g ? "," : "") + std::to_string(g);
        setenv("CUDA_VISIBLE_DEVICES", v.c_str(), 0);
      }
    }
    int32_t nDev = 0;
    if (bani_device_count(&nDev) != BANI_OK || nDev == 0) throw std::runtime_error("no CUDA device available (this program has no CPU path)");
    const int G = loading ? meta.shards : (parameters.gpus > 0 ? std::min(parameters.gpus, nDev) : nDev);
    if (G > nDev) throw std::runtime_error("the saved index has " + std::to_string(G) + " shards but only " + std::to_string(nDev) + " GPU(s) are visible");
    const auto shards = cgi::splitReferenceGenomes((int)parameters.refSequences.size(), G, parameters.blockPartition);
    // the contexts (CUDA context creation, stream set-up) come up while the reader threads parse and pack the files
    std::vector<bani_ctx *> ctxs(G, nullptr); std::vector<std::string> ctxErr(G);
    std::vector<std::thread> ctxThreads;
    for (int g = 0; g < G; g++)
      ctxThreads.emplace_back([&, g]() { bani_params cp = parameters.c(); if (bani_ctx_create(g, &cp, &ctxs[g]) != BANI_OK) ctxErr[g] = bani_last_error(); });

    // ---- ingest: every distinct file once, in parallel.  With a loaded index no reference file is read, and (one
    //      shard, no --visualize) neither is a query that is a genome of the index: its sketch is derived from the index
    auto t0 = Clock::now();
    std::unordered_map<std::string, int> refOrdinal;                  // path -> first position in the reference list
    for (size_t j = 0; j < parameters.refSequences.size(); j++) refOrdinal.emplace(parameters.refSequences[j], (int)j);
    const bool deriveQueries = loading && G == 1 && !parameters.visualize;
    auto derivable = [&](const std::string &q) { return deriveQueries && refOrdinal.count(q) > 0; };
    std::unordered_map<std::string, int> pathId; std::vector<std::string> paths;
    for (const auto &e : parameters.querySequences) if (!derivable(e) && !pathId.count(e)) { pathId[e] = (int)paths.size(); paths.push_back(e); }
    if (!loading) for (const auto &e : parameters.refSequences) if (!pathId.count(e)) { pathId[e] = (int)paths.size(); paths.push_back(e); }
    std::vector<bani_host::HostGenome> genomes(paths.size());
    {
      std::atomic<size_t> next(0); std::mutex emu; std::string err;
      auto work = [&]() {
        for (size_t i; (i = next++) < paths.size();) {
          try { genomes[i] = bani_host::read_genome(paths[i]); skch::pack_genome(genomes[i]); }     // parse + 2-bit pack in the reader thread
          catch (const std::exception &e) { std::lock_guard<std::mutex> l(emu); err = e.what(); }
        }
      };
      std::vector<std::thread> th;
      for (int i = 0; i < std::min<int>(parameters.threads, (int)paths.size()); i++) th.emplace_back(work);
      for (auto &t : th) t.join();
      for (auto &t : ctxThreads) t.join();
      if (!err.empty()) throw std::runtime_error(err);
    }
    std::unordered_map<std::string, uint64_t> genomeLengths;
    for (size_t i = 0; i < paths.size(); i++) genomeLengths[paths[i]] = cgi::genomeLength(genomes[i], parameters.minReadLength);
    std::cerr << "INFO, skch::main, Time spent reading " << paths.size() << " genome files : "
              << std::chrono::duration<double>(Clock::now() - t0).count() << " sec" << std::endl;
    for (int g = 0; g < G; g++) if (!ctxs[g]) throw std::runtime_error("bani_ctx_create: " + ctxErr[g]);

    std::vector<cgi::CGI_Results> finalResults;
    std::vector<std::string> visual(G);
    std::vector<char> sanity(G, 1); std::vector<float> ratioDiffs(G, 1.0f);
    std::vector<std::vector<std::string>> savedContigNames(G);
    std::mutex mu; std::string err;

    auto shardWork = [&](int g) {
      try {
        auto t1 = Clock::now();
        bani_ctx *ctx = ctxs[g];
        {
          // genomes this device needs: its reference shard (unless loaded) and every query that has to be read, each file once
          std::vector<int> need; std::unordered_map<int, int> slot;
          auto want = [&](const std::string &path) { const int id = pathId.at(path); if (!slot.count(id)) { slot[id] = (int)need.size(); need.push_back(id); } return slot[id]; };
          std::vector<int> refSlot, qrySlot;                              // qrySlot: -1 = derived from the index
          if (!loading) for (int j : shards[g]) refSlot.push_back(want(parameters.refSequences[j]));
          for (const auto &q : parameters.querySequences) qrySlot.push_back(derivable(q) ? -1 : want(q));
          std::vector<const bani_host::HostGenome *> hs; for (int id : need) hs.push_back(&genomes[id]);
          std::vector<std::unique_ptr<DeviceGenome>> dev;
          upload_genomes(ctx, hs, dev, (size_t)1 << 28, std::max(1, parameters.threads / G));
          std::vector<std::string> shardRefNames;
          for (int j : shards[g]) shardRefNames.push_back(parameters.refSequences[j]);

          std::unique_ptr<Sketch> referSketchP;                           // HP1, or the cache
          if (loading) referSketchP.reset(new Sketch(ctx, parameters, shardFile(parameters.loadIndex, g, G), meta.contigNames[g]));
          else {
            std::vector<const DeviceGenome *> refs;
            for (int sl : refSlot) refs.push_back(dev[sl].get());
            referSketchP.reset(new Sketch(ctx, parameters, refs));
          }
          Sketch &referSketch = *referSketchP;
          if (g == 0) std::cerr << "INFO [GPU 0], skch::main, Time spent " << (loading ? "loading" : "sketching") << " the reference : "
                                << std::chrono::duration<double>(Clock::now() - t1).count() << " sec" << std::endl;
          if (!parameters.saveIndex.empty()) {
            referSketch.save(shardFile(parameters.saveIndex, g, G));
            for (const auto &c : referSketch.metadata) savedContigNames[g].push_back(c.name);
          }
          if (loading) {                                                  // lengths of the shard's genomes for the --minFraction filter
            std::lock_guard<std::mutex> l(mu);
            for (size_t i = 0; i < shards[g].size(); i++) {
              const int c0 = i ? referSketch.sequencesByFileInfo[i - 1] : 0, c1 = referSketch.sequencesByFileInfo[i];
              genomeLengths.emplace(shardRefNames[i], cgi::genomeLength(referSketch.metadata, c0, c1, parameters.minReadLength));
            }
          }
          std::vector<cgi::CGI_Results> local;
          sanity[g] = referSketch.sanityCheck(parameters.maxRatioDiff); ratioDiffs[g] = referSketch.getRatioDifference();
          if (sanity[g]) {
            t1 = Clock::now();
            if (parameters.visualize) {
              std::ostringstream vis;
              for (size_t q = 0; q < qrySlot.size(); q++) {
                MappingResultsVector_t mapResults; uint64_t totalQueryFragments = 0;
                Map mapper(ctx, parameters, referSketch, *dev[qrySlot[q]], totalQueryFragments,
                           std::bind(Map::insertL2ResultsToVec, std::ref(mapResults), std::placeholders::_1));     // HP2
                cgi::computeCGI(parameters, mapResults, mapper, referSketch, totalQueryFragments, q, parameters.querySequences[q],
                                shardRefNames, &vis, local);
              }
              visual[g] = vis.str();
            } else {
              // HP2 + reduction for all queries: one sketch object for the queries that were read, one for those derived
              std::vector<bani_genome *> qh; std::vector<int32_t> qid, dord, did;
              for (size_t q = 0; q < qrySlot.size(); q++) {
                if (qrySlot[q] >= 0) { qh.push_back(dev[qrySlot[q]]->h); qid.push_back((int32_t)q); }
                else { dord.push_back(refOrdinal.at(parameters.querySequences[q])); did.push_back((int32_t)q); }
              }
              std::vector<bani_qsketch *> sk;
              if (!qh.empty()) { bani_qsketch *s = nullptr; check(bani_qsketch_create(ctx, qh.data(), (int32_t)qh.size(), qid.data(), referSketch.handle(), &s), "bani_qsketch_create"); sk.push_back(s); }
              if (!dord.empty()) { bani_qsketch *s = nullptr; check(bani_qsketch_from_index(ctx, referSketch.handle(), dord.data(), (int32_t)dord.size(), did.data(), &s), "bani_qsketch_from_index"); sk.push_back(s); }
              bani_cgi_result *res = nullptr; uint64_t n = 0; bani_map_counters ctr;
              const int rc = bani_map_cgi_sketch(ctx, referSketch.handle(), sk.data(), (int32_t)sk.size(), &res, &n, &ctr);     // HP2 + reduction
              for (auto *s : sk) bani_qsketch_destroy(s);
              check(rc, "bani_map_cgi_sketch");
              for (uint64_t i = 0; i < n; i++)
                local.push_back(cgi::CGI_Results{res[i].refGenomeId, res[i].qryGenomeId, res[i].countSeq, res[i].totalQueryFragments, res[i].identity});
              bani_free(res);
            }
            if (g == 0) std::cerr << "INFO [GPU 0], skch::main, Time spent mapping " << qrySlot.size() << " query genome(s) : "
                                  << std::chrono::duration<double>(Clock::now() - t1).count() << " sec" << std::endl;
          }
          cgi::correctRefGenomeIds(local, g, G, (int)parameters.refSequences.size(), parameters.blockPartition);
          std::lock_guard<std::mutex> l(mu);
          finalResults.insert(finalResults.end(), local.begin(), local.end());
        }
        if (getenv("BANI_CLI_FULL_TEARDOWN")) bani_ctx_destroy(ctx);      // otherwise the process exit returns the device memory (faster)
      } catch (const std::exception &e) { std::lock_guard<std::mutex> l(mu); err = e.what(); }
    };
    {
      std::vector<std::thread> th;
      for (int g = 0; g < G; g++) th.emplace_back(shardWork, g);
      for (auto &t : th) t.join();
    }
    if (!err.empty()) throw std::runtime_error(err);
    if (!parameters.saveIndex.empty()) writeMeta(parameters.saveIndex, parameters, G, savedContigNames);
    for (int g = 0; g < G; g++)
      if (!sanity[g]) std::cerr << "ERROR :: SPLIT " << g << "'s ratio difference " << ratioDiffs[g] << " exceeds maximum thresholds." << std::endl;
    // a query derived from the index has the length its reference twin has
    for (const auto &q : parameters.querySequences) if (!genomeLengths.count(q)) throw std::runtime_error("no length known for " + q);

    cgi::outputCGI(parameters, genomeLengths, finalResults, fileName);
    if (parameters.matrixOutput) cgi::outputPhylip(parameters, genomeLengths, finalResults, fileName);
    if (parameters.visualize) { std::ofstream o(fileName + ".visual"); for (int g = 0; g < G; g++) o << visual[g]; }
    std::cerr << "INFO, skch::main, Total time : " << std::chrono::duration<double>(Clock::now() - tStart).count() << " sec" << std::endl;
  } catch (const std::exception &e) {
    std::cerr << "ERROR, " << e.what() << std::endl;
    return 1;
  }
  std::cout.flush(); std::cerr.flush();
  if (!getenv("BANI_CLI_FULL_TEARDOWN")) _exit(0);      // outputs are written and closed: skip the destructors of multi-GB host tables and the CUDA teardown
  return 0;
}

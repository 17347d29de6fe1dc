// common.cuh -- internal types shared by the translation units of libfastani_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <memory>
#include <map>
#include <stdexcept>
#include <cstdio>
#include <cstdarg>
#include <chrono>
#include "../../include/fastani_b200.h"

namespace bani {

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);

[[noreturn]] inline void fail(int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  throw Error(code, buf);
}

#define BANI_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) \
  ::bani::fail(BANI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)

// ---------------------------------------------------------------- device memory
// DevBuf owns a block from the caching allocator of alloc.cpp (freed blocks are reused on the same
// stream; the driver is only called on a miss).
void *dev_alloc(size_t bytes, cudaStream_t st, size_t *granted, int *dev);
void  dev_free(void *p, size_t granted, cudaStream_t st, int dev);      // dev = the device the block was allocated on
void  dev_cache_flush(int dev);

template <typename T>
struct DevBuf {
  T *p = nullptr; size_t n = 0; cudaStream_t st = nullptr; size_t granted = 0; int dev = -1;
  DevBuf() {}
  DevBuf(size_t n_, cudaStream_t s) { alloc(n_, s); }
  DevBuf(const DevBuf &) = delete; DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), st(o.st), granted(o.granted), dev(o.dev) { o.p = nullptr; o.n = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept
  { if (this != &o) { release(); p = o.p; n = o.n; st = o.st; granted = o.granted; dev = o.dev; o.p = nullptr; o.n = 0; } return *this; }
  ~DevBuf() { release(); }
  void alloc(size_t n_, cudaStream_t s)
  {
    release(); n = n_; st = s;
    if (n == 0) return;
    p = (T *)dev_alloc(n * sizeof(T), s, &granted, &dev);
  }
  void release() { if (p) { dev_free(p, granted, st, dev); p = nullptr; } n = 0; }
  size_t bytes() const { return n * sizeof(T); }
};

// ---------------------------------------------------------------- statistics (host)
int   stat_recommended_window_size(double p_value, int k, float identity, int fragLen, uint64_t refSize);
int   stat_min_hits_relaxed(int s, int k, float identity);
void  stat_identity(int shared, int s, int k, float *id, float *ub);

// LUT rows consumed by the mapping kernels: for sketch size s,
//   minHits[s]            = max(1, estimateMinimumHitsRelaxed(s, k, pid))
//   rowOff[s] .. +s+1     : identity[x], upper[x] for x = 0..s
struct StatLut {
  int k = 0; float pid = 0;
  std::vector<int32_t> minHits;     // index s (0 unused)
  std::vector<uint32_t> rowOff;     // index s -> offset into ident/upper
  std::vector<float> ident, upper;
  int smax = 0;                     // rows 1 .. smax are all present (dense part)
  std::vector<char> have;           // beyond the dense part: rows computed on demand (index s)
  void ensure(int s_needed);        // extends the dense part up to s_needed (host)
  bool ensure_rows(const std::vector<int> &svals);   // makes the rows of these sketch sizes present; true if anything was added
};

// ---------------------------------------------------------------- genomes
// Per-contig descriptor consumed by the sketch kernel.  A query fragment is the
// same thing with a non-zero startBase and len = fragLen.
struct SeqDesc {
  const uint32_t *packed;   // first 2-bit word of the parent contig (16 bases / word)
  const uint32_t *excPos;   // sorted contig-relative positions of non-ACGT bytes (parent contig)
  const uint8_t  *excByte;  // their (upper-cased) bytes
  int32_t nExc;
  int32_t startBase;        // offset of this sequence inside the parent contig
  int32_t len;              // bases
  int32_t seqId;            // ordinal written to the records
};

uint64_t next_genome_uid();

// Device block shared by the genomes of one upload sub-batch of the host-packed path (bani_genome_create_packed_batch):
// one H2D copy per array on the context's copy stream; `ready` is recorded behind them and every consumer makes its
// stream wait for it, so uploads of later sub-batches overlap the sketch launches of earlier ones.
struct GenomeBlock {
  DevBuf<uint32_t> words; DevBuf<uint32_t> excPos; DevBuf<uint8_t> excByte;
  cudaEvent_t ready = nullptr; int device = 0;
  ~GenomeBlock() { if (ready) { cudaSetDevice(device); cudaEventDestroy(ready); } }
};

struct Genome {
  uint64_t uid = next_genome_uid();  // identity that survives address reuse (index membership, see Index::members)
  int device = 0;
  int32_t nContigs = 0;
  std::vector<int32_t> len;          // per contig
  std::vector<int64_t> wordOff;      // per contig, into packed (multiple of 4 words)
  std::vector<int64_t> excOff;       // per contig +1, into exc arrays
  uint64_t totalLen = 0, nExc = 0;
  DevBuf<uint32_t> packed;           // own buffers (ASCII ingest, pack.cu) ...
  DevBuf<uint32_t> excPos;
  DevBuf<uint8_t>  excByte;
  std::shared_ptr<GenomeBlock> blk;  // ... or a slice of a shared upload block (host-packed ingest)
  const uint32_t *blkWords = nullptr; const uint32_t *blkExcPos = nullptr; const uint8_t *blkExcByte = nullptr;
  const uint32_t *packedBase() const { return blk ? blkWords : packed.p; }
  const uint32_t *excPosBase() const { return blk ? blkExcPos : excPos.p; }
  const uint8_t  *excByteBase() const { return blk ? blkExcByte : excByte.p; }
  void wait_ready(cudaStream_t st) const { if (blk && blk->ready) cudaStreamWaitEvent(st, blk->ready, 0); }
};

struct Ctx;

// ---------------------------------------------------------------- query sketch (first half of HP2)
// Map::doL1Mapping, computeMap.hpp:252-276: per fragment the sorted unique minimizer hashes Q, s = |Q|.
// A piece holds at most 2^17 fragments of whole query genomes; arrays are packed back to back so that a
// sketch can be exported to one flat device buffer and moved between GPUs.
struct QPiece {
  uint64_t memberOf = 0;             // uid of the index the fragment sketches were derived from (stage A'), 0 = hashed / imported
  int q0 = 0, nq = 0;                // local query range [q0, q0 + nq) of the owning sketch
  int32_t F = 0; uint64_t T = 0; int smax = 0;
  DevBuf<uint32_t> fragHash;         // T  : sorted unique hashes of fragment f at [segStart[f], segStart[f+1])
  DevBuf<uint32_t> segStart;         // F+1
  DevBuf<int32_t>  sCount;           // F  : s
  DevBuf<int32_t>  fragQuery;        // F  : query slot inside the piece (0 .. nq)
  DevBuf<int32_t>  fragSeqId;        // F  : querySeqId of the mapping records (fragment ordinal inside its genome)
  std::vector<int32_t> qFragOff;     // nq+1 (host): first fragment of every query of the piece
};
struct QSketch {
  int device = 0; int k = 0, w = 0, fragLen = 0;
  std::vector<int32_t> queryId;          // id reported as qryGenomeId
  std::vector<uint64_t> totalFragments;  // Map's totalQueryFragments per query
  std::vector<std::unique_ptr<QPiece>> pieces;
  uint64_t F = 0, T = 0;
};

// ---------------------------------------------------------------- index (HP1 output)
struct Index {
  uint64_t uid = next_genome_uid();  // identity of this index (a query sketch remembers the index it was derived from)
  int device = 0;
  uint64_t M = 0, U = 0, totalLen = 0;
  int32_t nContigs = 0, nGenomes = 0;
  int dirBits = 0;
  // position-ordered records (== Sketch::minimizerIndex as SoA) + same-hash links
  DevBuf<uint32_t> hash; DevBuf<int32_t> wpos; DevBuf<int32_t> seqId; DevBuf<uint32_t> link;
  DevBuf<int2> pos8;                 // {wpos, seqId} per record: what the L1 stage fetches per sorted hit (one 8-byte load)
  DevBuf<uint2> rec8;                // 8-byte L2 record: x = hash, y = back:14 | fwd:14 | tie | new | gone (index.cu); valid where blkMax allows
  DevBuf<uint32_t> blkMax;           // per 1024 records: max back | max fwd << 16 (0xFFFF: a link does not fit 14 bits)
  DevBuf<uint4> rec;                 // 16-byte L2 record: x=hash, y=wpos|tie<<31, z=twin link, w=back|fwd<<16 (index.cu)
  int cmw = 0;                       // super-window width the back/fwd fields were computed for
  int k = 0, w = 0, fragLen = 0;     // parameters of the context the index was built with
  DevBuf<uint32_t> contigRecOff;     // nContigs+1: first record of each contig
  DevBuf<int32_t>  contigGenome;     // nContigs: genome ordinal of a contig (reviseRefIdToGenomeId)
  DevBuf<uint32_t> contigBinOff;     // nContigs+1: prefix of #position-bins per contig (CGI)
  // hash-ordered lookup side (== Sketch::minimizerPosLookupIndex)
  DevBuf<uint32_t> ukeys;            // U unique hashes ascending
  DevBuf<uint32_t> uoff;             // U+1 offsets into posIdx
  DevBuf<uint32_t> posIdx;           // M record indices, sorted by (hash, record index)
  DevBuf<uint32_t> dir;              // (1<<dirBits)+1 bucket directory over the top bits of the hash
  // one-sector probe table: 2^tabBits buckets of 4 entries {x = (hash & ~0xFF) | min(count, 255), y = offset into posIdx},
  // bucket = LOW bits of the hash (uniform, unlike the top bits of a minimizer hash); x == 0 = empty.  A full bucket or a
  // saturated count sends the probe to the sorted keys above (index.cu: table_fill_kernel, map.cu: lookup_kernel)
  DevBuf<uint2> tab; int tabBits = 0;
  // membership filter in front of the probe table for small shards (multi-GPU): one bit per value of the low filtBits bits
  // of the hash, sized 8x the unique hashes and at most 64 MB so that it stays L2-resident; a clear bit answers a miss
  // without touching DRAM (most probes of a shard are misses when the queries of other shards are mapped against it)
  DevBuf<uint32_t> filt; int filtBits = 0;
  std::vector<int32_t> contigLen;    // host copies
  std::vector<int32_t> seqsByFile;   // cumulative contig count per genome (sequencesByFileInfo)
  // Which genomes the index was built from (uid -> first contig ordinal) and, per hashed position, whether it was
  // VALID (forward hash != reverse-complement hash, commonFunc.hpp:131): together with the position-ordered records
  // this is enough to derive the fragment sketches of a member genome without hashing it again (map.cu, stage A').
  std::map<uint64_t, int32_t> members;
  DevBuf<uint32_t> validBits;        // bit contigBitBase[c] + p = position p of contig c is valid
  DevBuf<unsigned long long> contigBitBase;   // nContigs (multiples of 32)
  uint64_t totalBins = 0;
};

// A non-owning typed view of a persistent scratch slot (same surface as DevBuf where it is used).
template <typename T>
struct View {
  T *p = nullptr; size_t n = 0;
  size_t bytes() const { return n * sizeof(T); }
  void release() {}
};

// Run-time switches of a context (bani_ctx_set_flag); the defaults can also be set through the environment
// (BANI_NO_SKETCH_REUSE, BANI_MAX_HITS_PER_PIECE, BANI_FRAG_L1_MAX, BANI_L2E_BUCKETS, BANI_L2_STAGE, BANI_TRACE), read when the context is created.
struct CtxFlags {
  int sketchReuse = 1;                        // stage A': fragment sketches of index members are read from the index
  long long maxHitsPerPiece = 3ll << 29;      // a piece that gathers more index hits is split at a query boundary
  long long fragL1Max = 8192;                 // hits per fragment handled inside one CTA (<= FRAG_L1_MAX)
  int l2eBuckets = 0;                         // 0 = adaptive; 1024 / 4096 force the size of the L2 rank directory
  int trace = 0;                              // BANI_TRACE: wall-clock marks of the host orchestration on stderr (diagnostic)
  int l2Stage = 0;                            // 1: l2_events_kernel stages event codes in shared memory where the window links allow it
                                              //    (measured slower than the direct stores on B200: DESIGN.md section 6; kept as a tested alternative)
  long long uploadGroupWords = 16ll << 20;    // packed words (16 bases each) per upload group of the host-packed ingest
};

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;   // H2D uploads of host-packed genomes (overlap with the sketch launches on `stream`)
  bani_params prm{};
  CtxFlags flags;
  int smCount = 0;
  StatLut lut;
  DevBuf<int32_t> d_minHits; DevBuf<uint32_t> d_rowOff; DevBuf<float> d_ident, d_upper;
  int lutUploaded = 0;
  void upload_lut(int smax, const int32_t *d_sCount = nullptr, int32_t F = 0);
  // optional per-stage device timing (CUDA events on `stream`), see bani_ctx_profile_*
  // diagnostic: microseconds of host wall clock since the previous mark (flags.trace)
  std::chrono::steady_clock::time_point traceT = std::chrono::steady_clock::now();
  void mark(const char *what)
  {
    if (!flags.trace) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[bani trace] %-28s +%8.1f us\n", what, std::chrono::duration<double, std::micro>(t - traceT).count());
    traceT = t;
  }
  uint64_t launches = 0;             // kernels of this library launched so far (CUB's not counted)
  bool profiling = false;
  struct ProfEv { const char *name; cudaEvent_t a, b; double bytes; };
  std::vector<ProfEv> profEvents;
  // Grow-only scratch slots for the per-chunk working set of the mapping pipeline (hit keys, flags,
  // candidate arrays ...): after the first chunk nothing is allocated or freed inside the timed path.
  std::map<int, DevBuf<uint8_t>> slots;
  // kernel function attributes (dynamic shared memory limit, carve-out) are per DEVICE: remember per context what has
  // been set, so that a process driving several GPUs (the C++ CLI: one thread + context per GPU) sets them on each
  std::map<const void *, bool> attrDone;
  bool first_time(const void *key) { bool &d = attrDone[key]; const bool f = !d; d = true; return f; }
  template <typename T> View<T> view(int id, size_t n)
  {
    DevBuf<uint8_t> &b = slots[id];
    const size_t bytes = (n ? n : 1) * sizeof(T);
    if (b.n < bytes) b.alloc(bytes + bytes / 8 + 256, stream);
    View<T> v; v.p = (T *)b.p; v.n = n; return v;
  }
};
#ifndef BANI_FILE_TAG
#define BANI_FILE_TAG 0
#endif
// one scratch slot per (source file, source line): never put two BANI_SCRATCH on one line
#define BANI_SLOT_ID ((BANI_FILE_TAG << 20) | __LINE__)
#define BANI_SCRATCH(T, name, count) ::bani::View<T> name = ctx->view<T>(BANI_SLOT_ID, (count))

// RAII stage timer: records an event pair around a stage when profiling is on.
struct Stage {
  Ctx *c; size_t idx = (size_t)-1;
  Stage(Ctx *ctx, const char *name, double algoBytes = 0) : c(ctx)
  {
    if (!c->profiling) return;
    Ctx::ProfEv e; e.name = name; e.bytes = algoBytes;
    cudaEventCreate(&e.a); cudaEventCreate(&e.b);
    cudaEventRecord(e.a, c->stream);
    idx = c->profEvents.size(); c->profEvents.push_back(e);
  }
  void bytes(double b) { if (idx != (size_t)-1) c->profEvents[idx].bytes = b; }
  size_t id() const { return idx; }                       // to set the bytes after the stage has been closed
  static void set_bytes(Ctx *c, size_t id, double b) { if (id != (size_t)-1 && id < c->profEvents.size()) c->profEvents[id].bytes = b; }
  void stop() { if (idx != (size_t)-1) { cudaEventRecord(c->profEvents[idx].b, c->stream); idx = (size_t)-1; } }
  ~Stage() { stop(); }
};

// ---------------------------------------------------------------- kernels' host entry points
// pack.cu
void genome_create_batch(Ctx *ctx, int32_t nGenomes, const int32_t *genOff, const int64_t *off,
                         const uint8_t *seq, Genome **out);
void genome_decode(Ctx *ctx, const Genome *g, int32_t contig, uint8_t *out, int64_t cap);
void genome_create_packed_batch(Ctx *ctx, int32_t nGenomes, const int32_t *genOff, const int32_t *contigLen, const int64_t *wordOff,
                                const uint32_t *words, const int64_t *excOff, const uint32_t *excPos, const uint8_t *excByte,
                                bool async, Genome **out);
uint64_t host_pack_contig(const uint8_t *seq, int64_t len, uint32_t *words, uint32_t *excPos, uint8_t *excByte, uint64_t excCap);

// sketch.cu : windowed minimizers of a list of sequences, records compacted in
// (sequence, wpos) order.  Outputs may be null (skipped).  Returns total records
// (which may exceed `cap`; only the first cap are stored).
// recBase: records already written by earlier launches into the same output arrays (index build in upload groups);
// outputs and segStart values are offset by it, `cap` is the capacity of the whole arrays.
uint64_t sketch_sequences(Ctx *ctx, const SeqDesc *d_desc, int32_t nSeq, const int32_t *h_len, int32_t uniformLen,
                          uint32_t *o_hash, int32_t *o_wpos, int32_t *o_seqId, uint64_t cap,
                          uint32_t *o_segStart /* nSeq+1 */,
                          uint32_t *o_validBits = nullptr, const unsigned long long *bitBase = nullptr, uint64_t recBase = 0);

// index.cu
Index *index_build(Ctx *ctx, Genome *const *refs, int32_t nRefs);
void   index_save(Ctx *ctx, const Index *ix, const char *path);
Index *index_load(Ctx *ctx, const char *path);

// map.cu
struct MapOutput {
  std::vector<bani_mapping> rows;               // when wantRows
  std::vector<bani_cgi_result> cgi;             // when wantCgi
  std::vector<uint64_t> totalQueryFragments;    // per query
  bani_map_counters ctr{};
};
void map_queries(Ctx *ctx, const Index *ix, const Genome *const *queries, int32_t nq,
                 bool wantRows, bool wantCgi, MapOutput &out);
QSketch *qsketch_create(Ctx *ctx, const Genome *const *queries, int32_t nq, const int32_t *queryIds, const Index *hint);
QSketch *qsketch_from_index(Ctx *ctx, const Index *ix, const int32_t *ordinals, int32_t nq, const int32_t *queryIds);
uint64_t qsketch_export_bytes(const QSketch *qs);
void qsketch_export(Ctx *ctx, const QSketch *qs, void *devBuf, uint64_t cap);
QSketch *qsketch_import(Ctx *ctx, const void *devBuf, uint64_t bytes);
QSketch *qsketch_merge(Ctx *ctx, const QSketch *const *sketches, int32_t n);
void qsketch_map(Ctx *ctx, const Index *ix, const QSketch *const *sketches, int32_t nSketches,
                 bool wantRows, bool wantCgi, MapOutput &out);

// hits.cu : per-fragment gather + shared-memory sort + L1 candidate regions
static constexpr unsigned long long FRAG_L1_MAX = 8192;   // hits per fragment handled inside one CTA
static constexpr int FRAG_NCLASS = 13;                    // size classes: 256 * {1,2,3,4,5,6,7,8,10,12,16,24,32} hits
__host__ __device__ inline int frag_class_items(int cls)
{
  return cls < 8 ? cls + 1 : (cls == 8 ? 10 : cls == 9 ? 12 : cls == 10 ? 16 : cls == 11 ? 24 : 32);
}
struct FragL1Args {
  const uint32_t *segStart; const int32_t *sCount; int32_t F;
  const uint32_t *hitLo, *hitCnt; const unsigned long long *hitOff;
  const uint32_t *posIdx; const int2 *recPos;    // recPos[r] = {wpos, seqId} of record r
  const int32_t *minHits; int fragLen, keyBits;
  int32_t *stSeq, *stStart, *stEnd;      // staging, addressed by global hit offset
  uint32_t *candCount;                   // per fragment
};
void frag_classify(Ctx *ctx, const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                   uint32_t *candCount, uint32_t *fragClass, uint32_t *classCount, uint32_t *classList, unsigned long long maxFast);
void frag_l1_fast(Ctx *ctx, const FragL1Args &a, const uint32_t *classList, const uint32_t *classCount);
void cand_stage(Ctx *ctx, const int32_t *cFrag, const int32_t *cSeq, const int32_t *cStart, const int32_t *cEnd, uint32_t C,
                const uint32_t *segStart, const unsigned long long *hitOff, int32_t *stSeq, int32_t *stStart, int32_t *stEnd,
                uint32_t *candCount);
void cand_compact(Ctx *ctx, const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                  const uint32_t *candCount, const uint32_t *candOff, const int32_t *stSeq, const int32_t *stStart,
                  const int32_t *stEnd, int32_t *cFrag, int32_t *cSeq, int32_t *cStart, int32_t *cEnd);

// synth.cu
void synth_genome(Ctx *ctx, uint64_t seed, uint32_t ancestor, uint32_t strain, uint32_t ppm,
                  int64_t len, uint8_t *hostOut);

// CUB-backed primitives (cubops.cu; kept in one TU because CUB compiles slowly)
size_t cub_sort_pairs_u32_temp(size_t n);
void   cub_sort_pairs_u32(void *temp, size_t tempBytes, const uint32_t *kin, uint32_t *kout,
                          const uint32_t *vin, uint32_t *vout, size_t n, int endBit, cudaStream_t s);
size_t cub_sort_keys_u64_temp(size_t n);
void   cub_sort_keys_u64(void *temp, size_t tempBytes, const uint64_t *kin, uint64_t *kout, size_t n,
                         int beginBit, int endBit, cudaStream_t s);
size_t cub_scan_u32_temp(size_t n);
void   cub_exclusive_sum_u32(void *temp, size_t tempBytes, const uint32_t *in, uint32_t *out, size_t n, cudaStream_t s);
size_t cub_scan_u64_temp(size_t n);
void   cub_exclusive_sum_u32_to_u64(void *temp, size_t tempBytes, const uint32_t *in, uint64_t *out, size_t n, cudaStream_t s);

} // namespace bani

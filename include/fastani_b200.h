/* fastani_b200.h -- C ABI of the B200-native ANI hot path.
 *
 * This is the drop-in boundary for the two data-parallel hot paths of FastANI
 * (reference citations are relative to the upstream tree, ParBLiSS/FastANI):
 *
 *   HP1  reference index build   skch::Sketch::Sketch(const Parameters&)
 *                                 src/map/include/winSketch.hpp:109-115
 *   HP2  query mapping           skch::Map::Map(const Parameters&, const Sketch&,
 *                                 uint64_t& totalQueryFragments, int queryno, callback)
 *                                 src/map/include/computeMap.hpp:93-102
 *   (+)  per-pair reduction      cgi::computeCGI
 *                                 src/cgi/include/computeCoreIdentity.hpp:166-298
 *
 * The reference has no FFI layer; its seam is those two constructors and one
 * callback.  The entry points below are what a binding for that seam needs:
 * plain pointers and sizes, no C++ or torch types.  All device work runs on the
 * CUDA device the context was created on (sm_100a kernels); there is NO CPU
 * fallback -- every call fails with BANI_ERR_CUDA if no device is usable.
 *
 * Conventions: every function returns 0 (BANI_OK) or a negative bani_status;
 * bani_last_error() gives a thread-local message for the last failure.  The
 * library allocates outputs; the caller frees them through the library.
 * Handles are not thread-safe individually; distinct contexts may be used from
 * distinct threads concurrently.
 */
#ifndef FASTANI_B200_H
#define FASTANI_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BANI_API __attribute__((visibility("default")))

typedef enum {
  BANI_OK = 0,
  BANI_ERR_ARG = -1,       /* invalid argument */
  BANI_ERR_CUDA = -2,      /* CUDA runtime failure / no usable device */
  BANI_ERR_NOMEM = -3,     /* host or device allocation failed */
  BANI_ERR_LIMIT = -4,     /* an implementation limit was exceeded (message says which) */
  BANI_ERR_INTERNAL = -5
} bani_status;

/* skch::Parameters, src/map/include/map_parameters.hpp:22-41 -- the fields the
 * hot path reads.  Zero-initialise, then call bani_params_default(). */
typedef struct {
  int32_t  kmer_size;          /* kmerSize            (default 16) */
  int32_t  window_size;        /* windowSize; <=0 => Stat::recommendedWindowSize */
  int32_t  frag_len;           /* minReadLength       (default 3000) */
  float    perc_identity;      /* percentageIdentity  (default 80) */
  double   p_value;            /* p_value             (default 1e-3) */
  uint64_t reference_size;     /* referenceSize       (default 5000000) */
  int32_t  reserved[8];
} bani_params;

/* skch::MappingResult, src/map/include/base_types.hpp:89-102 (44 bytes, same field order) */
typedef struct {
  int32_t queryLen, refStartPos, refEndPos, queryStartPos, queryEndPos;
  int32_t refSeqId, querySeqId;
  float   nucIdentity, nucIdentityUpperBound;
  int32_t sketchSize, conservedSketches;
} bani_mapping;

/* skch::MinimizerInfo, src/map/include/base_types.hpp:21-53 (12 bytes) */
typedef struct { uint32_t hash; int32_t seqId; int32_t wpos; } bani_minimizer;

/* cgi::CGI_Results, src/cgi/include/cgid_types.hpp:68-80.  refGenomeId is the
 * ordinal of the genome inside the index it was mapped against. */
typedef struct {
  int32_t refGenomeId, qryGenomeId, countSeq, totalQueryFragments;
  float   identity;
} bani_cgi_result;

/* Integer work counters of one mapping call; SURVEY.md section 8(d) defines the
 * algorithmic bytes of HP2 from these (same meaning as the oracle's counters). */
typedef struct {
  uint64_t fragments;     /* F   query fragments considered                      */
  uint64_t sum_s;         /* sum of unique query minimizers probed               */
  uint64_t hits;          /* H   index positions gathered                        */
  uint64_t candidates;    /* L1 candidate regions                                */
  uint64_t n2;            /* position-ordered records scanned over all candidates*/
  uint64_t mappings;      /* P   mapping records kept                            */
  uint64_t reserved[4];
} bani_map_counters;

typedef struct bani_ctx    bani_ctx;      /* one per CUDA device: stream, scratch, statistic LUTs */
typedef struct bani_genome bani_genome;   /* a genome resident in HBM: 2-bit packed contigs + exceptions */
typedef struct bani_index  bani_index;    /* a reference index resident in HBM (HP1 output)             */

BANI_API const char *bani_last_error(void);
BANI_API const char *bani_version(void);
BANI_API void bani_params_default(bani_params *p);

/* Stat::recommendedWindowSize, src/map/include/map_stats.hpp:226-256 (host, double math). */
BANI_API int bani_recommended_window_size(const bani_params *p);
/* Stat::estimateMinimumHitsRelaxed (map_stats.hpp:142-167) and the identity /
 * upper-bound expressions of Map::doL2Mapping (computeMap.hpp:375-381): exposed
 * so the statistic LUT the kernels consume can be checked on its own. */
BANI_API int bani_stat_min_hits_relaxed(int s, int k, float perc_identity);
BANI_API int bani_stat_identity(int shared, int s, int k, float *identity, float *upper_bound);

/* Number of usable CUDA devices (0 without a driver / GPU); lets a host program shard the reference list
 * (cgi::splitReferenceGenomes, computeCoreIdentity.hpp:457-474) without linking the CUDA runtime itself. */
BANI_API int bani_device_count(int32_t *n);

/* ---- context -------------------------------------------------------------- */
BANI_API int  bani_ctx_create(int device, const bani_params *p, bani_ctx **out);
BANI_API void bani_ctx_destroy(bani_ctx *ctx);
BANI_API int  bani_ctx_params(const bani_ctx *ctx, bani_params *out);   /* window_size resolved */
BANI_API int  bani_ctx_sync(bani_ctx *ctx);
/* Raw CUDA stream (cudaStream_t) every launch of this context goes to; for event timing. */
BANI_API void *bani_ctx_stream(bani_ctx *ctx);

/* Run-time switches of a context.  name: "sketch_reuse" (1 = read the fragment sketches of index members from the
 * index, 0 = always hash the query fragments: what a run with --ql != --rl does), "max_hits_per_piece",
 * "frag_l1_max", "l2e_buckets", "l2_stage", "upload_group_words" (tuning / test switches; results never depend on them).
 * Environment, read when a context is created: BANI_NO_SKETCH_REUSE, BANI_MAX_HITS_PER_PIECE, BANI_FRAG_L1_MAX,
 * BANI_L2E_BUCKETS, BANI_L2_STAGE set the defaults of those switches; BANI_TRACE=1 prints the host wall clock between
 * marks of the orchestration (index build, query sketches, every piece of the mapping) on stderr. */
BANI_API int  bani_ctx_set_flag(bani_ctx *ctx, const char *name, int64_t value);

/* Per-stage device timing.  When enabled, every stage of HP1/HP2 is bracketed by CUDA events on
 * the context's stream; bani_ctx_profile_read() synchronises, sums the elapsed time, launch count and
 * algorithmic bytes per stage name since the last read, and clears the record.  names: n_max
 * buffers of 32 chars. */
BANI_API int  bani_ctx_profile_enable(bani_ctx *ctx, int on);
BANI_API int  bani_ctx_profile_read(bani_ctx *ctx, char (*names)[32], double *ms, double *algo_bytes,
                                    int32_t *launches, int32_t n_max, int32_t *n);

/* Number of this library's own kernels launched on the context so far (CUB launches excluded). */
BANI_API uint64_t bani_ctx_launch_count(const bani_ctx *ctx);

/* Pinned host memory for staging genomes (optional; pageable buffers also work). */
BANI_API int  bani_host_alloc(size_t bytes, void **out);
BANI_API void bani_host_free(void *p);

/* ---- genome ingest --------------------------------------------------------
 * A genome is what one FASTA file holds: n_contigs sequences given as the raw
 * bytes kseq_read() yields (seq->seq.s / seq->seq.l, winSketch.hpp:147-150):
 * any case, any IUPAC or other byte.  Bytes are upper-cased (a-z only,
 * commonFunc.hpp:57-66) and packed on the GPU to 2 bits/base; every non-ACGT
 * byte is carried out of band so hashing sees exactly the reference's bytes.
 * `seq` is one host buffer; contig c occupies [off[c], off[c+1]).
 * Contigs keep their ordinal even when shorter than k or w (winSketch.hpp:150-164). */
BANI_API int  bani_genome_create(bani_ctx *ctx, int32_t n_contigs, const int64_t *off,
                                 const uint8_t *seq, bani_genome **out);
/* Several genomes in one call (one device synchronisation for the whole batch).
 * genome g owns contigs [gen_off[g], gen_off[g+1]) of the off[] table. */
BANI_API int  bani_genome_create_batch(bani_ctx *ctx, int32_t n_genomes, const int32_t *gen_off,
                                       const int64_t *off, const uint8_t *seq, bani_genome **out);
/* Host-packed ingest: the 2-bit layout of the device (16 bases per uint32, A0 C1 G2 T3, base i in bits [2*(i%16), +2);
 * every other byte after upper-casing a-z is code 0 plus an out-of-band (contig-relative position, byte) entry) produced
 * on the HOST -- by the reader threads, as the bytes come off kseq_read -- so that 0.25 bytes per base cross PCIe instead
 * of 1.  bani_pack_contig needs no GPU; it writes (len + 15) / 16 words and returns the exception count through *n_exc
 * (only the first exc_cap are stored: retry with larger arrays if it is bigger).
 * bani_genome_create_packed_batch: contig c of the batch has contig_len[c] bases at words + word_off[c] (multiple of 4,
 * ascending, no overlap) and its exceptions at [exc_off[c], exc_off[c+1]); genome g owns contigs
 * [gen_off[g], gen_off[g+1]).  The copies run on a second stream in groups of <= 64 MB so that an index build that
 * follows hashes the first groups while the last are in flight.  async = 0: the call returns when the copies are done;
 * async != 0: it returns at once and the host arrays (pinned) must stay untouched until the genomes have been consumed
 * by a call that returns results (bani_index_build, bani_map_*, bani_qsketch_create) or bani_ctx_sync. */
BANI_API int  bani_pack_contig(const uint8_t *seq, int64_t len, uint32_t *words, uint32_t *exc_pos, uint8_t *exc_byte,
                               uint64_t exc_cap, uint64_t *n_exc);
BANI_API int  bani_genome_create_packed_batch(bani_ctx *ctx, int32_t n_genomes, const int32_t *gen_off, const int32_t *contig_len,
                                              const int64_t *word_off, const uint32_t *words, const int64_t *exc_off,
                                              const uint32_t *exc_pos, const uint8_t *exc_byte, int32_t async, bani_genome **out);
BANI_API void bani_genome_destroy(bani_genome *g);
BANI_API int  bani_genome_info(const bani_genome *g, int32_t *n_contigs, uint64_t *total_len,
                               uint64_t *n_exceptions, uint64_t *n_fragments);
/* Device -> host decode of one contig back to (upper-cased) ASCII; test hook for the packer. */
BANI_API int  bani_genome_decode(bani_ctx *ctx, const bani_genome *g, int32_t contig, uint8_t *out, int64_t cap);

/* ---- HP1: reference index build -------------------------------------------
 * Sketch::build + Sketch::index (winSketch.hpp:124-193) over the given genomes
 * in order; seqId runs over all contigs of all genomes (sequencesByFileInfo,
 * winSketch.hpp:75,167).  An empty list builds an empty index. */
BANI_API int  bani_index_build(bani_ctx *ctx, bani_genome *const *refs, int32_t n_refs, bani_index **out);
BANI_API void bani_index_destroy(bani_index *ix);
/* Totals needed by Sketch::sanityCheck (winSketch.hpp:298-318) and the log lines. */
BANI_API int  bani_index_stats(const bani_index *ix, uint64_t *n_minimizers, uint64_t *n_unique,
                               uint64_t *total_len, uint64_t *n_contigs, uint64_t *n_genomes);
/* The position-ordered minimizer table (== Sketch::minimizerIndex, winSketch.hpp:94), copied to host. */
BANI_API int  bani_index_minimizers(bani_ctx *ctx, const bani_index *ix, bani_minimizer *out, uint64_t cap);
/* The lookup side (== Sketch::minimizerPosLookupIndex, winSketch.hpp:84): all positions of one hash.
 * Returns the count through *n; writes at most cap (seqId, wpos) pairs. */
BANI_API int  bani_index_lookup(bani_ctx *ctx, const bani_index *ix, uint32_t hash,
                                int32_t *seqId, int32_t *wpos, uint64_t cap, uint64_t *n);

/* ---- on-disk sketch cache (the reference has none: scripts/splitDatabase.sh + README.md:104-106 re-sketch every
 * reference in every run).  bani_index_save writes what only the sketch launch can produce -- position-ordered
 * (hash, wpos) records, contig table, validity bitmap, parameters, checksum -- as one flat file; bani_index_load reads
 * it back on a context with the SAME k / window / fragLen (anything else is refused, like an in-memory mismatch) and
 * rebuilds the lookup side on the GPU.  Host metadata (genome paths, contig names) is the caller's to store. */
BANI_API int  bani_index_save(bani_ctx *ctx, const bani_index *ix, const char *path);
BANI_API int  bani_index_load(bani_ctx *ctx, const char *path, bani_index **out);
/* Contig lengths of the index in seqId order (cap >= n_contigs of bani_index_stats) and the cumulative contig count per
 * genome (== Sketch::sequencesByFileInfo, winSketch.hpp:75; cap >= n_genomes): what a host needs to rebuild
 * Sketch::metadata lengths and computeGenomeLengths (computeCoreIdentity.hpp:48-92) for a loaded index. */
BANI_API int  bani_index_contigs(const bani_index *ix, int32_t *contig_len, uint64_t cap_contigs, int32_t *seqs_by_file, uint64_t cap_genomes);

/* ---- HP2: query mapping ---------------------------------------------------
 * Map::mapQuery (computeMap.hpp:112-196) for one query genome: every mapping the
 * reference would pass to its callback, in the same (fragment, candidate) order.
 * *rows is allocated by the library (free with bani_free). */
BANI_API int  bani_map_genome(bani_ctx *ctx, const bani_index *ix, const bani_genome *query,
                              bani_mapping **rows, uint64_t *n_rows,
                              uint64_t *total_query_fragments, bani_map_counters *counters);

/* Map + cgi::computeCGI fused on the device for a batch of query genomes: one
 * bani_cgi_result per (query, reference genome) pair with at least one 2-way
 * mapping, ordered by (query, refGenomeId); qryGenomeId = position in `queries`.
 * total_query_fragments (optional, n_queries entries) is filled per query. */
BANI_API int  bani_map_cgi(bani_ctx *ctx, const bani_index *ix, bani_genome *const *queries, int32_t n_queries,
                           bani_cgi_result **results, uint64_t *n_results,
                           uint64_t *total_query_fragments, bani_map_counters *counters);

/* ---- query sketch: the first half of HP2 as an object -------------------------
 * Map::doL1Mapping, computeMap.hpp:252-276: the sorted unique minimizer hashes of every 3 kb fragment of a
 * set of query genomes.  bani_map_cgi() builds it internally; building it separately lets a multi-GPU run
 * sketch each query ONCE (rank r sketches queries r, r+N, ...), move the sketches between GPUs
 * (export -> NCCL all-gather -> import; a sketch is ~0.33 bytes per query base) and map every sketch against
 * each rank's reference shard.  query_ids[i] is reported as qryGenomeId (NULL: 0..n-1).
 * hint (optional): an index on the same device.  Queries that are genomes the hint index was built from get their
 * fragment sketches from its minimizer records instead of being hashed again (the windows of a fragment are the
 * contig windows inside it) -- all-vs-all runs, and every rank of the multi-GPU split, hash each genome once.
 * bani_map_genome / bani_map_cgi pass their index as the hint themselves. */
typedef struct bani_qsketch bani_qsketch;
BANI_API int  bani_qsketch_create(bani_ctx *ctx, bani_genome *const *queries, int32_t n_queries, const int32_t *query_ids,
                                  const bani_index *hint, bani_qsketch **out);
/* The same for genomes OF the index, by genome ordinal, from the index alone (no genome handle, no bases): an index
 * loaded from disk is also the query side of an all-vs-all run. */
BANI_API int  bani_qsketch_from_index(bani_ctx *ctx, const bani_index *ix, const int32_t *genome_ordinals, int32_t n_queries,
                                      const int32_t *query_ids, bani_qsketch **out);
BANI_API void bani_qsketch_destroy(bani_qsketch *qs);
BANI_API int  bani_qsketch_info(const bani_qsketch *qs, int32_t *n_queries, uint64_t *n_fragments, uint64_t *n_hashes,
                                uint64_t *export_bytes);
/* Pack into / rebuild from one flat DEVICE buffer (cap >= export_bytes; 16-byte aligned). */
BANI_API int  bani_qsketch_export(bani_ctx *ctx, const bani_qsketch *qs, void *device_buf, uint64_t cap);
BANI_API int  bani_qsketch_import(bani_ctx *ctx, const void *device_buf, uint64_t bytes, bani_qsketch **out);
/* Several sketches of this device as ONE (queries in the order given; the sources stay valid and may be destroyed): the
 * sketches a rank received from its peers are then mapped in a few large passes instead of one small pass per peer. */
BANI_API int  bani_qsketch_merge(bani_ctx *ctx, const bani_qsketch *const *sketches, int32_t n_sketches, bani_qsketch **out);
/* bani_map_cgi for prebuilt sketches (all on this context's device); results ordered by (sketch, query, refGenomeId). */
BANI_API int  bani_map_cgi_sketch(bani_ctx *ctx, const bani_index *ix, const bani_qsketch *const *sketches, int32_t n_sketches,
                                  bani_cgi_result **results, uint64_t *n_results, bani_map_counters *counters);

BANI_API void bani_free(void *p);

/* ---- bench utilities -------------------------------------------------------
 * Deterministic synthetic genome, generated on the device and written to a host
 * buffer as upper-case ACGT ASCII (counter-based generator; fastani_b200/synth.py
 * is the same function in numpy).  ancestor_id picks the ancestor sequence,
 * strain_id the substitution stream, sub_rate_ppm the per-base substitution rate. */
BANI_API int  bani_synth_genome(bani_ctx *ctx, uint64_t seed, uint32_t ancestor_id, uint32_t strain_id,
                                uint32_t sub_rate_ppm, int64_t len, uint8_t *host_out);

#ifdef __cplusplus
}
#endif
#endif /* FASTANI_B200_H */

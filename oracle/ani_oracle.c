/* oracle/ani_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT.
 *
 * A plain-C, single-threaded CPU restatement of the FastANI hot path
 * (reference index build + query mapping + the per-pair identity reduction),
 * written from the reference's *behaviour* with flat arrays and brute-force
 * set semantics -- deliberately a different formulation from both the
 * reference (deque / unordered_map / std::map with a pivot iterator) and the
 * CUDA kernels (prev/next links + incremental pivot), so that agreement of
 * the three means something.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; the product (fastani_b200/) never does.
 *
 * PARITY PINNED: every function here is checked in tests/test_oracle.py
 * against fixtures generated from the UNMODIFIED reference compiled in this
 * container (oracle/_ref/ref_dump, oracle/_ref/fastANI_ref; generator script
 * tests/golden/make_golden.py) and against the reference's own goldens
 * (tests/data/stsrsq-test.txt{,.visual}, README.md:80).
 *
 * Each function cites the reference file:line (relative to /root/reference)
 * whose behaviour it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "binom.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ hashing */

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k)
{
  k ^= k >> 33; k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return k;
}

/* MurmurHash3_x64_128(key, len, seed=42), first 4 bytes of the output (LE) as
 * uint32: src/common/murmur3.h:226-303 consumed by
 * src/map/include/commonFunc.hpp:71-81 (getHash), seed at commonFunc.hpp:32. */
ORC_API uint32_t orc_hash(const uint8_t *data, int len)
{
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  int nblocks = len / 16;
  for (int i = 0; i < nblocks; i++) {
    uint64_t k1, k2;
    memcpy(&k1, data + 16 * i, 8);
    memcpy(&k2, data + 16 * i + 8, 8);
    k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const uint8_t *tail = data + 16 * nblocks;
  int rem = len & 15;
  uint64_t k1 = 0, k2 = 0;
  for (int j = rem - 1; j >= 8; j--) k2 ^= (uint64_t)tail[j] << (8 * (j - 8));
  if (rem > 8) { k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; }
  for (int j = (rem < 8 ? rem : 8) - 1; j >= 0; j--) k1 ^= (uint64_t)tail[j] << (8 * j);
  if (rem > 0) { k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1; }
  h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return (uint32_t)h1;
}

/* commonFunc.hpp:37-54: only A/C/G/T are complemented, every other byte is kept */
static inline uint8_t comp(uint8_t b)
{
  switch (b) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return b; }
}

/* commonFunc.hpp:57-66: a-z -> A-Z, everything else untouched */
ORC_API void orc_upper(uint8_t *s, int64_t n)
{
  for (int64_t i = 0; i < n; i++) if (s[i] > 96 && s[i] < 123) s[i] -= 32;
}

/* --------------------------------------------------------------- minimizers */

/* Restates CommonFunc::addMinimizers (commonFunc.hpp:92-167) as a
 * data-parallel specification (no deque):
 *   valid(i)  = hashFwd(i) != hashBwd(i)                (:131)
 *   h(i)      = min(hashFwd, hashBwd)                   (:134)
 *   m(i)      = RIGHTMOST arg-min of h over valid j in [i-w+1, i]
 *               (expiry :137, pop-back on >= :142 => ties go right)
 *   for valid i >= w-1: emit (h[m(i)], seqId, wpos = i-w+1) iff m(i) differs
 *   from m at the previous valid position >= w-1          (:152-161)
 * `seq` must already be upper-cased.  Outputs have capacity len.  Returns the
 * number of records. */
ORC_API int64_t orc_minimizers(const uint8_t *seq, int64_t len, int k, int w, int32_t seqId,
                               uint32_t *o_hash, int32_t *o_seq, int32_t *o_wpos)
{
  if (len < k || len < w) return 0;   /* winSketch.hpp:153 / computeMap.hpp:138 */
  int64_t np = len - k + 1;
  uint32_t *h = (uint32_t *)malloc(sizeof(uint32_t) * np);
  uint8_t *valid = (uint8_t *)malloc(np);
  uint8_t *rc = (uint8_t *)malloc(k);
  for (int64_t i = 0; i < np; i++) {
    uint32_t hf = orc_hash(seq + i, k);
    for (int j = 0; j < k; j++) rc[j] = comp(seq[i + k - 1 - j]);
    uint32_t hb = orc_hash(rc, k);
    valid[i] = hf != hb;
    h[i] = hf < hb ? hf : hb;
  }
  int64_t n = 0, prevM = -1;
  for (int64_t i = w - 1; i < np; i++) {
    if (!valid[i]) continue;
    int64_t m = -1;
    for (int64_t j = i - w + 1; j <= i; j++)
      if (valid[j] && (m < 0 || h[j] <= h[m])) m = j;
    if (m != prevM) {
      o_hash[n] = h[m]; o_seq[n] = seqId; o_wpos[n] = (int32_t)(i - w + 1);
      n++;
      prevM = m;
    }
  }
  free(h); free(valid); free(rc);
  return n;
}

/* --------------------------------------------------------------- statistics */

/* map_stats.hpp:44-56.  Types follow the reference exactly: the argument is
 * float, `1+j` is a float sum, log() is the double overload. */
static float st_j2md(float j, int k)
{
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float mash_dist = (-1.0 / k) * log(2.0 * j / (1 + j));
  return mash_dist;
}

/* map_stats.hpp:62-66: k*d is a float product handed to the double exp() */
static float st_md2j(float d, int k)
{
  float kd = k * d;
  float jaccard = 1.0 / (2.0 * exp((double)kd) - 1.0);
  return jaccard;
}

/* map_stats.hpp:79-109 (GSL branch) */
static float st_md_lower_bound(float d, int s, int k, float ci)
{
  float q2 = (1.0 - ci) / 2;
  float sj = s * st_md2j(d, k);
  int x = (int)ceil((double)sj);
  if (x < 1) x = 1;
  while (x <= s) {
    double cdf_complement = orc_binomial_Q((unsigned)(x - 1), (double)st_md2j(d, k), (unsigned)s);
    if (cdf_complement < q2) { x--; break; }
    x++;
  }
  float jaccard = (float)x / s;
  return st_j2md(jaccard, k);
}

/* map_stats.hpp:118-130 */
static int st_min_hits(int s, int k, float perc_identity)
{
  float mash_dist = 1.0 - perc_identity / 100.0;
  float jaccard = st_md2j(mash_dist, k);
  return (int)ceil(1.0 * s * jaccard);
}

/* map_stats.hpp:142-167 */
ORC_API int orc_min_hits_relaxed(int s, int k, float perc_identity)
{
  int first = st_min_hits(s, k, perc_identity);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = st_j2md(jaccard, k);
    float d_lower = st_md_lower_bound(d, s, k, 0.9);
    float id_upper = 100.0 * (1.0 - d_lower);
    if (id_upper >= perc_identity) relaxed = i; else break;
  }
  return relaxed;
}

/* The float expressions of Map::doL2Mapping, computeMap.hpp:375-381 */
ORC_API void orc_identity(int shared, int s, int k, float *nucIdentity, float *nucIdentityUpperBound)
{
  float mash_dist = st_j2md(1.0 * shared / s, k);
  float lb = st_md_lower_bound(mash_dist, s, k, 0.9);
  *nucIdentity = 100 * (1 - mash_dist);
  *nucIdentityUpperBound = 100 * (1 - lb);
}

/* map_stats.hpp:179-216 */
static double st_pvalue(int s, int k, int alphabetSize, float identity, int lengthQuery, uint64_t lengthReference)
{
  double kmerSpace = pow(alphabetSize, k);
  double pX, pY;
  pX = pY = 1. / (1. + kmerSpace / lengthQuery);
  double r = pX * pY / (pX + pY - pX * pY);
  int x = orc_min_hits_relaxed(s, k, identity);
  double cdf_complement = (x == 0) ? 1.0 : orc_binomial_Q((unsigned)(x - 1), r, (unsigned)s);
  return lengthReference * cdf_complement;
}

/* map_stats.hpp:226-256 with the defaults of parseCmdArgs.hpp:118-130
 * (p=1e-3, identity 80, referenceSize 5e6, alphabet 4).  The reference reads an
 * uninitialised variable when no sketch size qualifies (:237,:252; happens for
 * k=21, fragLen=1000); that case returns -1 here ("degenerate"). */
ORC_API int orc_window_size(int k, int fragLen)
{
  int best = -1;
  int cand[3] = {1, 2, 5};
  for (int c = 0; c < 3 && best < 0; c++)
    if (st_pvalue(cand[c], k, 4, 80, fragLen, 5000000) <= 1e-03) best = cand[c];
  for (int e = 10; e < fragLen && best < 0; e += 10)
    if (st_pvalue(e, k, 4, 80, fragLen, 5000000) <= 1e-03) best = e;
  if (best < 0) return -1;
  int w = 2.0 * fragLen / best;
  if (w < 1) w = 1;
  if (w > fragLen) w = fragLen;
  return w;
}

/* ------------------------------------------------------------------ mapping */

typedef struct {
  int32_t queryLen, refStartPos, refEndPos, queryStartPos, queryEndPos, refSeqId, querySeqId;
  float nucIdentity, nucIdentityUpperBound;
  int32_t sketchSize, conservedSketches;
} orc_mapping;   /* == skch::MappingResult, base_types.hpp:89-102 (44 bytes) */

typedef struct {
  int64_t M;                 /* records, ordered by (seqId, wpos) */
  const uint32_t *hash; const int32_t *seq; const int32_t *wpos;
  int64_t *byHash;           /* record indices, stably sorted by hash (Sketch::index, winSketch.hpp:181-193) */
} orc_index;

static const orc_index *g_sort_ix;
static int cmp_by_hash(const void *a, const void *b)
{
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  uint32_t hx = g_sort_ix->hash[x], hy = g_sort_ix->hash[y];
  if (hx != hy) return hx < hy ? -1 : 1;
  return x < y ? -1 : (x > y);
}

ORC_API orc_index *orc_index_new(int64_t M, const uint32_t *hash, const int32_t *seq, const int32_t *wpos)
{
  orc_index *ix = (orc_index *)calloc(1, sizeof(orc_index));
  ix->M = M; ix->hash = hash; ix->seq = seq; ix->wpos = wpos;
  ix->byHash = (int64_t *)malloc(sizeof(int64_t) * (M > 0 ? M : 1));
  for (int64_t i = 0; i < M; i++) ix->byHash[i] = i;
  g_sort_ix = ix;
  qsort(ix->byHash, M, sizeof(int64_t), cmp_by_hash);
  return ix;
}

ORC_API int64_t orc_index_unique(const orc_index *ix)
{
  int64_t u = 0;
  for (int64_t i = 0; i < ix->M; i++)
    if (i == 0 || ix->hash[ix->byHash[i]] != ix->hash[ix->byHash[i - 1]]) u++;
  return u;
}

ORC_API void orc_index_free(orc_index *ix) { if (ix) { free(ix->byHash); free(ix); } }

/* Sketch::searchIndex, winSketch.hpp:259-270: lower_bound by (seqId, wpos) */
static int64_t search_pos(const orc_index *ix, int32_t seqId, int32_t pos)
{
  int64_t lo = 0, hi = ix->M;
  while (lo < hi) {
    int64_t mid = (lo + hi) >> 1;
    int less = ix->seq[mid] < seqId || (ix->seq[mid] == seqId && ix->wpos[mid] < pos);
    if (less) lo = mid + 1; else hi = mid;
  }
  return lo;
}

static int cmp_u32(const void *a, const void *b) { uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b; return x < y ? -1 : (x > y); }
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return x < y ? -1 : (x > y); }

/* index of h in sorted unique Q[0..s), or -(gap+1) where gap = #q < h */
static int q_find(const uint32_t *Q, int s, uint32_t h)
{
  int lo = 0, hi = s;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (Q[mid] < h) lo = mid + 1; else hi = mid; }
  if (lo < s && Q[lo] == h) return lo;
  return -(lo + 1);
}

/* small open-addressing multiset for window hashes that are not in Q */
#define MS_CAP 16384
typedef struct { uint32_t key[MS_CAP]; int32_t cnt[MS_CAP]; uint8_t used[MS_CAP]; } multiset;
static int ms_add(multiset *m, uint32_t h, int delta)   /* returns new count */
{
  uint32_t i = (h * 2654435761u) & (MS_CAP - 1);
  while (m->used[i] && m->key[i] != h) i = (i + 1) & (MS_CAP - 1);
  if (!m->used[i]) { m->used[i] = 1; m->key[i] = h; m->cnt[i] = 0; }
  m->cnt[i] += delta;
  return m->cnt[i];
}

typedef struct { int32_t seqId, start, end; } l1cand;

typedef struct {
  int64_t sum_s, hits, n2, mappings, candidates, events;
} orc_counters;

/* Map::computeL2MappedRegions (computeMap.hpp:418-497) + MIIteratorL2
 * (MIIteratorL2.hpp:54-96) + SlideMapper (slidingMap.hpp:112-284), restated
 * with SET semantics evaluated from counting arrays:
 *   W  = distinct hashes of the records in [b, e)
 *   p  = s-th smallest element of Q u W
 *   shared = |{ h in Q n W : h <= p }|
 * which in rank space is  t* = max{ t : t + #(W\Q below q_t) <= s },
 * shared = #(present q_j, j <= t*). */
static void l2_region(const orc_index *ix, const uint32_t *Q, int s, int fragLen, int k, int w,
                      const l1cand *c, int32_t *o_pos, int32_t *o_shared, orc_counters *ctr,
                      int32_t *cntQ, int32_t *gap, multiset *ms)
{
  int cmw = fragLen - (w - 1) - (k - 1);
  int64_t b = search_pos(ix, c->seqId, c->start);
  int64_t e = search_pos(ix, c->seqId, ix->wpos[b] + cmw);
  int64_t last = search_pos(ix, c->seqId, c->end + fragLen);
  int32_t sw_pos = ix->wpos[b];
  memset(cntQ, 0, sizeof(int32_t) * s);
  memset(gap, 0, sizeof(int32_t) * (s + 1));
  memset(ms->used, 0, sizeof(ms->used));
  ctr->n2 += last - b;
#define INS(r) do { int f = q_find(Q, s, ix->hash[r]); if (f >= 0) cntQ[f]++; \
                    else if (ms_add(ms, ix->hash[r], +1) == 1) gap[-f - 1]++; } while (0)
#define DEL(r) do { int f = q_find(Q, s, ix->hash[r]); if (f >= 0) cntQ[f]--; \
                    else if (ms_add(ms, ix->hash[r], -1) == 0) gap[-f - 1]--; } while (0)
  for (int64_t r = b; r < e; r++) INS(r);
  int best = 0; int32_t first = 0, lastp = 0;
  while (e < last) {
    /* evaluate the window [b, e) */
    int t = 0, G = 0, shared = 0;
    while (t < s && (t + 1) + G + gap[t] <= s) { G += gap[t]; if (cntQ[t] > 0) shared++; t++; }
    ctr->events++;
    if (shared > best) { best = shared; first = lastp = ix->wpos[b]; }
    else if (shared == best) lastp = ix->wpos[b];
    /* MIIteratorL2::next */
    int32_t lastPos = sw_pos + cmw - 1;
    int32_t d1 = ix->wpos[b + 1] - sw_pos, d2 = ix->wpos[e] - lastPos;
    int32_t adv = d1 < d2 ? d1 : d2;
    sw_pos += adv;
    if (adv == d1) { DEL(b); b++; }
    if (adv == d2) { INS(e); e++; }
  }
#undef INS
#undef DEL
  *o_pos = (int32_t)(((int64_t)first + lastp) / 2);
  *o_shared = best;
}

/* Map::mapQuery + mapSingleQuerySeq + doL1Mapping + computeL1CandidateRegions +
 * doL2Mapping (computeMap.hpp:112-410) for ONE query genome given as
 * upper-cased contigs.  Rows are appended to `rows` (capacity cap); returns the
 * number of rows, or -1 on overflow. */
ORC_API int64_t orc_map_genome(const orc_index *ix, int n_contigs, const int64_t *off, const uint8_t *seq,
                               int k, int w, int fragLen, float pid,
                               orc_mapping *rows, int64_t cap,
                               uint64_t *totalQueryFragments, orc_counters *ctr)
{
  int64_t nrows = 0;
  int32_t seqCounter = 0;
  int maxm = fragLen + 1;
  uint32_t *mh = (uint32_t *)malloc(sizeof(uint32_t) * maxm);
  int32_t *ms_ = (int32_t *)malloc(sizeof(int32_t) * maxm), *mw = (int32_t *)malloc(sizeof(int32_t) * maxm);
  int32_t *cntQ = (int32_t *)malloc(sizeof(int32_t) * maxm), *gap = (int32_t *)malloc(sizeof(int32_t) * (maxm + 1));
  multiset *mset = (multiset *)malloc(sizeof(multiset));
  int64_t hcap = 1 << 16; int64_t *hits = (int64_t *)malloc(sizeof(int64_t) * hcap);
  int64_t ccap = 1 << 10; l1cand *cands = (l1cand *)malloc(sizeof(l1cand) * ccap);
  int *minhits_memo = (int *)malloc(sizeof(int) * (maxm + 1));
  for (int i = 0; i <= maxm; i++) minhits_memo[i] = -1;
  orc_counters local = {0, 0, 0, 0, 0, 0};
  if (!ctr) ctr = &local;

  for (int c = 0; c < n_contigs; c++) {
    int64_t len = off[c + 1] - off[c];
    if (len < w || len < k || len < fragLen) continue;          /* computeMap.hpp:138 */
    int fragmentCount = (int)(len / fragLen);                     /* :152 */
    for (int f = 0; f < fragmentCount; f++) {
      const uint8_t *fs = seq + off[c] + (int64_t)f * fragLen;   /* :173-175 */
      int32_t fragId = seqCounter + f;
      /* doL1Mapping :260-276 */
      int64_t nm = orc_minimizers(fs, fragLen, k, w, 0, mh, ms_, mw);
      qsort(mh, nm, sizeof(uint32_t), cmp_u32);
      int s = 0;
      for (int64_t i = 0; i < nm; i++) if (i == 0 || mh[i] != mh[i - 1]) mh[s++] = mh[i];
      if (s == 0) continue;                                       /* :278 */
      ctr->sum_s += s;
      /* :283-299 gather every index position of every unique hash */
      int64_t H = 0;
      for (int i = 0; i < s; i++) {
        int64_t lo = 0, hi = ix->M;
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (ix->hash[ix->byHash[mid]] < mh[i]) lo = mid + 1; else hi = mid; }
        for (int64_t r = lo; r < ix->M && ix->hash[ix->byHash[r]] == mh[i]; r++) {
          if (H == hcap) { hcap *= 2; hits = (int64_t *)realloc(hits, sizeof(int64_t) * hcap); }
          hits[H++] = ix->byHash[r];
        }
      }
      ctr->hits += H;
      if (minhits_memo[s] < 0) minhits_memo[s] = orc_min_hits_relaxed(s, k, pid);   /* :301 */
      int minimumHits = minhits_memo[s] < 1 ? 1 : minhits_memo[s];                  /* :316-317 */
      qsort(hits, H, sizeof(int64_t), cmp_i64);                                     /* :320 */
      int64_t nc = 0;
      for (int64_t a = 0; a + minimumHits <= H; a++) {                              /* :322-352 */
        int64_t ra = hits[a], rb = hits[a + minimumHits - 1];
        if (ix->seq[rb] == ix->seq[ra] && ix->wpos[rb] - ix->wpos[ra] < fragLen) {
          l1cand cd; cd.seqId = ix->seq[ra];
          cd.start = ix->wpos[rb] - fragLen + 1; if (cd.start < 0) cd.start = 0;
          cd.end = ix->wpos[ra];
          if (nc > 0 && cd.seqId == cands[nc - 1].seqId && cands[nc - 1].end >= cd.start) {
            if (cd.end > cands[nc - 1].end) cands[nc - 1].end = cd.end;
          } else {
            if (nc == ccap) { ccap *= 2; cands = (l1cand *)realloc(cands, sizeof(l1cand) * ccap); }
            cands[nc++] = cd;
          }
        }
      }
      ctr->candidates += nc;
      /* doL2Mapping :363-410 */
      for (int64_t ci = 0; ci < nc; ci++) {
        int32_t pos, shared;
        l2_region(ix, mh, s, fragLen, k, w, &cands[ci], &pos, &shared, ctr, cntQ, gap, mset);
        float id, ub;
        orc_identity(shared, s, k, &id, &ub);
        if (ub >= pid) {
          if (nrows == cap) { nrows = -1; goto done; }
          orc_mapping *r = &rows[nrows++];
          r->queryLen = fragLen; r->refStartPos = pos; r->refEndPos = pos + fragLen - 1;
          r->queryStartPos = 0; r->queryEndPos = fragLen - 1;
          r->refSeqId = cands[ci].seqId; r->querySeqId = fragId;
          r->nucIdentity = id; r->nucIdentityUpperBound = ub;
          r->sketchSize = s; r->conservedSketches = shared;
          ctr->mappings++;
        }
      }
    }
    seqCounter += fragmentCount;
    *totalQueryFragments += fragmentCount;                        /* :188-189 */
  }
done:
  free(mh); free(ms_); free(mw); free(cntQ); free(gap); free(mset); free(hits); free(cands); free(minhits_memo);
  return nrows;
}

/* ---------------------------------------------------------------------- CGI */

typedef struct { int32_t refSeq, genome, qSeq, refStart, bin; float id; } cgi_row;

static int cmp_query_bucket(const void *a, const void *b)   /* cgid_types.hpp:31-39 */
{
  const cgi_row *x = (const cgi_row *)a, *y = (const cgi_row *)b;
  if (x->genome != y->genome) return x->genome < y->genome ? -1 : 1;
  if (x->qSeq != y->qSeq) return x->qSeq < y->qSeq ? -1 : 1;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  if (x->refSeq != y->refSeq) return x->refSeq < y->refSeq ? -1 : 1;
  if (x->refStart != y->refStart) return x->refStart < y->refStart ? -1 : 1;
  return 0;
}
static int cmp_refbin_bucket(const void *a, const void *b)  /* cgid_types.hpp:45-53 */
{
  const cgi_row *x = (const cgi_row *)a, *y = (const cgi_row *)b;
  if (x->refSeq != y->refSeq) return x->refSeq < y->refSeq ? -1 : 1;
  if (x->bin != y->bin) return x->bin < y->bin ? -1 : 1;
  if (x->id != y->id) return x->id < y->id ? -1 : 1;
  /* the reference leaves ties to std::sort; break them deterministically */
  if (x->qSeq != y->qSeq) return x->qSeq < y->qSeq ? -1 : 1;
  return 0;
}

/* cgi::computeCGI, src/cgi/include/computeCoreIdentity.hpp:166-298.
 * seqsByFile[g] = cumulative contig count after reference genome g
 * (Sketch::sequencesByFileInfo, winSketch.hpp:75,167).  Outputs one entry per
 * reference genome that received at least one 2-way mapping; returns their
 * number.  Optionally returns the surviving 2-way rows (for .visual). */
ORC_API int orc_cgi(const orc_mapping *rows, int64_t n, const int32_t *seqsByFile, int n_genomes, int fragLen,
                    int32_t *o_genome, int32_t *o_count, float *o_identity,
                    int32_t *v_refSeq, int32_t *v_qSeq, int32_t *v_refStart, float *v_id, int64_t *v_n)
{
  cgi_row *r = (cgi_row *)malloc(sizeof(cgi_row) * (n > 0 ? n : 1));
  for (int64_t i = 0; i < n; i++) {
    r[i].refSeq = rows[i].refSeqId; r[i].qSeq = rows[i].querySeqId; r[i].refStart = rows[i].refStartPos;
    r[i].bin = rows[i].refStartPos / (fragLen - 20);                 /* :194 */
    r[i].id = rows[i].nucIdentity;
    int g = 0; while (g < n_genomes && seqsByFile[g] <= rows[i].refSeqId) g++;   /* upper_bound, :36-40 */
    r[i].genome = g;
  }
  qsort(r, n, sizeof(cgi_row), cmp_query_bucket);                     /* :214 */
  int64_t n1 = 0;
  for (int64_t i = 0; i < n; i++) {                                   /* :216-231, keep LAST of each run */
    if (n1 > 0 && r[i].genome == r[n1 - 1].genome && r[i].qSeq == r[n1 - 1].qSeq) r[n1 - 1] = r[i];
    else r[n1++] = r[i];
  }
  qsort(r, n1, sizeof(cgi_row), cmp_refbin_bucket);                   /* :237 */
  int64_t n2 = 0;
  for (int64_t i = 0; i < n1; i++) {                                  /* :239-254 */
    if (n2 > 0 && r[i].refSeq == r[n2 - 1].refSeq && r[i].bin == r[n2 - 1].bin) r[n2 - 1] = r[i];
    else r[n2++] = r[i];
  }
  if (v_n) {
    for (int64_t i = 0; i < n2; i++) { v_refSeq[i] = r[i].refSeq; v_qSeq[i] = r[i].qSeq; v_refStart[i] = r[i].refStart; v_id[i] = r[i].id; }
    *v_n = n2;
  }
  int out = 0;
  for (int64_t i = 0; i < n2;) {                                      /* :267-297 */
    int64_t j = i; float sum = 0.0f;
    while (j < n2 && r[j].genome == r[i].genome) { sum += r[j].id; j++; }   /* sequential float32 sum */
    o_genome[out] = r[i].genome; o_count[out] = (int32_t)(j - i); o_identity[out] = sum / (int32_t)(j - i);
    out++;
    i = j;
  }
  free(r);
  return out;
}

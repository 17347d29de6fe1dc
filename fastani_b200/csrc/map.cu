// map.cu -- HP2: query mapping (+ the per-pair identity reduction that consumes it).
//
// Replaces skch::Map::mapQuery / doL1Mapping / computeL1CandidateRegions / doL2Mapping /
// computeL2MappedRegions (src/map/include/computeMap.hpp:112-497) with SlideMapper
// (slidingMap.hpp) and MIIteratorL2 (MIIteratorL2.hpp), and cgi::computeCGI
// (src/cgi/include/computeCoreIdentity.hpp:166-298).
//
// The reference maps one fragment at a time.  Here a batch of query genomes is cut into
// fragments and every stage runs over ALL fragments (or all hits / all candidates) of the batch:
//
//   A  sketch      fragment minimizers, fragment-local windows (computeMap.hpp:260)      sketch.cu
//   B  sort/unique per fragment: sorted unique hashes Q, s = |Q| (computeMap.hpp:268-276)
//   C  lookup      Q -> bucket directory -> unique keys -> position lists (:283-299)
//   D  hits        gathered as (fragment << 32 | record index) and radix-sorted: the record index
//                  is monotone in (seqId, wpos), so this is the sort of :320 for all fragments
//   E  L1          candidate regions: the scan + merge of :322-352 is a LOCAL rule on the sorted
//                  hits (a hit starts a region unless its left neighbour qualifies and overlaps),
//                  so regions come from head flags + one prefix sum, in reference order
//   F  L2          per candidate: sliding super-window over the position-ordered records with
//                  the winnowed-MinHash intersection of slidingMap.hpp restated in rank space:
//                      t* = max{ t : t + #(distinct window hashes not in Q, below q_t) <= s }
//                      shared = #(q_j present in the window, j <= t*)
//                  maintained incrementally (each event moves t* by at most one)
//   G  report      identity / upper bound from the (s, shared) table, filter >= cutoff (:375-384),
//                  rows in (fragment, candidate) order == callback order of reportL2Mappings
//   H  CGI         1-way best per (fragment, genome); 2-way best per (ref contig, position bin) via
//                  atomicMax on a dense bin table; ordered float32 sum per genome pair
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace bani {

static inline unsigned nblk(uint64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

// ------------------------------------------------------------------ B: per-fragment sort + unique
static constexpr int SU_THREADS = 128;
static constexpr int SU_CAP = 4096;

__global__ void __launch_bounds__(SU_THREADS)
sort_unique_kernel(uint32_t *fragHash, const uint32_t *segStart, int32_t F, int32_t *sCount,
                   int *smax, int *err)
{
  __shared__ uint32_t a[SU_CAP];
  __shared__ uint32_t wsum[SU_THREADS / 32];
  const int f = blockIdx.x, tid = threadIdx.x;
  const uint32_t beg = segStart[f];
  const int n = (int)(segStart[f + 1] - beg);
  if (n > SU_CAP) { if (tid == 0) { atomicExch(err, 1); sCount[f] = 0; } return; }
  if (n == 0) { if (tid == 0) sCount[f] = 0; return; }
  int n2 = 1; while (n2 < n) n2 <<= 1;
  for (int i = tid; i < n2; i += SU_THREADS) a[i] = i < n ? fragHash[beg + i] : 0xFFFFFFFFu;
  __syncthreads();
  for (int k2 = 2; k2 <= n2; k2 <<= 1)
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n2; i += SU_THREADS) {
        int p = i ^ j;
        if (p > i) {
          uint32_t x = a[i], y = a[p];
          bool asc = (i & k2) == 0;
          if ((x > y) == asc) { a[i] = y; a[p] = x; }
        }
      }
      __syncthreads();
    }
  // unique: each thread owns a contiguous run of the sorted array
  const int per = (n + SU_THREADS - 1) / SU_THREADS;
  const int i0 = min(tid * per, n), i1 = min(i0 + per, n);
  uint32_t cnt = 0;
  for (int i = i0; i < i1; i++) cnt += (i == 0 || a[i] != a[i - 1]);
  uint32_t incl = cnt; const int lane = tid & 31, wid = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  uint32_t base = 0, total = 0;
  for (int i = 0; i < SU_THREADS / 32; i++) { if (i < wid) base += wsum[i]; total += wsum[i]; }
  uint32_t o = base + incl - cnt;
  for (int i = i0; i < i1; i++) if (i == 0 || a[i] != a[i - 1]) fragHash[beg + o++] = a[i];
  if (tid == 0) { sCount[f] = (int32_t)total; atomicMax(smax, (int)total); }
}

// ------------------------------------------------------------------ C: lookup
__device__ __forceinline__ int seg_of(const uint32_t *segStart, int F, uint32_t t)
{
  int lo = 0, hi = F - 1;          // last f with segStart[f] <= t
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (segStart[mid] <= t) lo = mid; else hi = mid - 1; }
  return lo;
}

__global__ void lookup_kernel(const uint32_t *fragHash, const uint32_t *segStart, const int32_t *sCount, int32_t F,
                              uint32_t T, const uint32_t *ukeys, const uint32_t *uoff, const uint32_t *dir, int dirBits,
                              uint32_t *hitLo, uint32_t *hitCnt)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > T) return;
  if (t == T) { hitCnt[t] = 0; hitLo[t] = 0; return; }
  const int f = seg_of(segStart, F, t);
  uint32_t cnt = 0, lo0 = 0;
  if ((int)(t - segStart[f]) < sCount[f]) {
    const uint32_t h = fragHash[t];
    const uint32_t b = h >> (32 - dirBits);
    uint32_t lo = dir[b], hi = dir[b + 1];
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ukeys[mid] < h) lo = mid + 1; else hi = mid; }
    if (lo < dir[b + 1] && ukeys[lo] == h) { lo0 = uoff[lo]; cnt = uoff[lo + 1] - lo0; }
  }
  hitLo[t] = lo0; hitCnt[t] = cnt;
}

// ------------------------------------------------------------------ D: gather hits as 64-bit keys
__global__ void gather_kernel(const uint32_t *segStart, int32_t F, uint32_t T, const uint32_t *hitLo, const uint32_t *hitCnt,
                              const unsigned long long *hitOff, const uint32_t *posIdx, unsigned long long *keys)
{
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t cnt = hitCnt[t];
  if (!cnt) return;
  const unsigned long long f = (unsigned long long)seg_of(segStart, F, t);
  const uint32_t lo = hitLo[t];
  unsigned long long o = hitOff[t];
  for (uint32_t j = 0; j < cnt; j++) keys[o + j] = (f << 32) | posIdx[lo + j];
}

// ------------------------------------------------------------------ E: L1 candidate regions
struct L1Args {
  const unsigned long long *keys; unsigned long long N;
  const uint32_t *segStart; const unsigned long long *hitOff; const int32_t *sCount;
  const int32_t *minHits;          // LUT indexed by s
  const int32_t *recSeq; const int32_t *recWpos;
  int fragLen;
};

// does sorted hit i start a (raw) candidate?  (computeMap.hpp:324-336)
__device__ __forceinline__ bool l1_qual(const L1Args &a, unsigned long long i, unsigned long long fragEnd, int mh,
                                        uint32_t ra, int32_t &start)
{
  if (i + (unsigned long long)mh > fragEnd) return false;
  const uint32_t rb = (uint32_t)a.keys[i + mh - 1];
  if (a.recSeq[rb] != a.recSeq[ra]) return false;
  const int32_t wb = a.recWpos[rb];
  if (wb - a.recWpos[ra] >= a.fragLen) return false;
  start = max(0, wb - a.fragLen + 1);
  return true;
}

// flags: bit0 = head of a merged region, bit1 = tail of a merged region
__device__ __forceinline__ uint32_t l1_flags(const L1Args &a, unsigned long long i, int32_t &start, uint32_t &ra_out, int &f_out)
{
  const unsigned long long key = a.keys[i];
  const int f = (int)(key >> 32); const uint32_t ra = (uint32_t)key;
  ra_out = ra; f_out = f;
  const unsigned long long fragBeg = a.hitOff[a.segStart[f]], fragEnd = a.hitOff[a.segStart[f + 1]];
  const int mh = a.minHits[a.sCount[f]];
  if (!l1_qual(a, i, fragEnd, mh, ra, start)) return 0;
  uint32_t fl = 0;
  // merged with the left neighbour iff it qualifies, same contig, and its end (= its wpos) >= our start (:342-350)
  bool merged = false;
  if (i > fragBeg) {
    const uint32_t rp = (uint32_t)a.keys[i - 1]; int32_t sp;
    if (a.recSeq[rp] == a.recSeq[ra] && l1_qual(a, i - 1, fragEnd, mh, rp, sp) && a.recWpos[rp] >= start) merged = true;
  }
  if (!merged) fl |= 1;
  bool nextMerges = false;
  if (i + 1 < fragEnd) {
    const uint32_t rn = (uint32_t)a.keys[i + 1]; int32_t sn;
    if (a.recSeq[rn] == a.recSeq[ra] && l1_qual(a, i + 1, fragEnd, mh, rn, sn) && a.recWpos[ra] >= sn) nextMerges = true;
  }
  if (!nextMerges) fl |= 2;
  return fl;
}

__global__ void l1_flag_kernel(const L1Args a, uint32_t *head)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > a.N) return;
  if (i == a.N) { head[i] = 0; return; }
  int32_t start; uint32_t ra; int f;
  head[i] = l1_flags(a, i, start, ra, f) & 1u;
}

__global__ void l1_write_kernel(const L1Args a, const uint32_t *head, const uint32_t *headScan,
                                int32_t *cFrag, int32_t *cSeq, int32_t *cStart, int32_t *cEnd)
{
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.N) return;
  int32_t start; uint32_t ra; int f;
  const uint32_t fl = l1_flags(a, i, start, ra, f);
  if (fl & 1) { const uint32_t c = headScan[i]; cFrag[c] = f; cSeq[c] = a.recSeq[ra]; cStart[c] = start; }
  if (fl & 2) { const uint32_t c = headScan[i] + head[i] - 1; cEnd[c] = a.recWpos[ra]; }
}

// ------------------------------------------------------------------ F: L2 (thread per candidate, state in global scratch)
struct L2Args {
  const int32_t *cFrag, *cSeq, *cStart, *cEnd; uint32_t C;
  const uint32_t *fragHash; const uint32_t *segStart; const int32_t *sCount;
  const uint32_t *recHash; const int32_t *recWpos; const uint32_t *recLink; const uint32_t *contigRecOff;
  int fragLen, cmw, smax;
  uint8_t *scratch; size_t stride;
  int32_t *cPos, *cBest;
  unsigned long long *ctr_n2;
  int onlyFlagged;
};

__device__ __forceinline__ uint32_t lb_wpos(const int32_t *wpos, uint32_t lo, uint32_t hi, int32_t v)
{
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (wpos[mid] < v) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int lb_q(const uint32_t *Q, int s, uint32_t h)
{
  int lo = 0, hi = s;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (Q[mid] < h) lo = mid + 1; else hi = mid; }
  return lo;
}

__global__ void __launch_bounds__(64)
l2_kernel(const L2Args a)
{
  const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x, nslots = gridDim.x * blockDim.x;
  uint16_t *gap = (uint16_t *)(a.scratch + (size_t)slot * a.stride);
  uint8_t *pres = (uint8_t *)(gap + a.smax + 2);
  unsigned long long n2 = 0;
  for (uint32_t c = slot; c < a.C; c += nslots) {
    if (a.onlyFlagged && a.cBest[c] != -1) continue;           // the fast path already solved it
    const int f = a.cFrag[c];
    const int s = a.sCount[f];
    const uint32_t *Q = a.fragHash + a.segStart[f];
    const int seq = a.cSeq[c];
    const uint32_t lo = a.contigRecOff[seq], hi = a.contigRecOff[seq + 1];
    uint32_t b = lb_wpos(a.recWpos, lo, hi, a.cStart[c]);
    uint32_t e = lb_wpos(a.recWpos, lo, hi, a.recWpos[b] + a.cmw);
    const uint32_t last = lb_wpos(a.recWpos, lo, hi, a.cEnd[c] + a.fragLen);
    n2 += last - b;
    for (int i = 0; i <= s; i++) gap[i] = 0;
    for (int i = 0; i < s; i++) pres[i] = 0;
    int t = s, G = 0, P = 0;

    auto insert = [&](uint32_t r, uint32_t wb) {
      const uint32_t pd = a.recLink[r] >> 16;
      if (pd != 0xFFFFu && r - pd >= wb) return;              // an earlier twin is inside the window
      const uint32_t h = a.recHash[r];
      const int j = lb_q(Q, s, h);
      if (j < s && Q[j] == h) { pres[j] = 1; if (j < t) P++; }
      else { gap[j]++; if (j < t) G++; while (t + G > s) { t--; G -= gap[t]; P -= pres[t]; } }
    };
    auto remove = [&](uint32_t r, uint32_t we) {
      const uint32_t nd = a.recLink[r] & 0xFFFFu;
      if (nd != 0xFFFFu && r + nd < we) return;               // a later twin is still inside the window
      const uint32_t h = a.recHash[r];
      const int j = lb_q(Q, s, h);
      if (j < s && Q[j] == h) { pres[j] = 0; if (j < t) P--; }
      else { gap[j]--; if (j < t) G--; while (t < s && t + 1 + G + (int)gap[t] <= s) { G += gap[t]; P += pres[t]; t++; } }
    };

    for (uint32_t r = b; r < e; r++) insert(r, b);
    int sw = a.recWpos[b];
    int best = 0, first = 0, lastp = 0;
    while (e < last) {
      if (P > best) { best = P; first = lastp = a.recWpos[b]; }
      else if (P == best) lastp = a.recWpos[b];
      const int d1 = a.recWpos[b + 1] - sw, d2 = a.recWpos[e] - (sw + a.cmw - 1);
      const int adv = min(d1, d2);
      sw += adv;
      const uint32_t ob = b, oe = e;
      if (adv == d1) { remove(ob, oe); b++; }
      if (adv == d2) { insert(oe, b); e++; }
    }
    a.cPos[c] = (first + lastp) / 2;
    a.cBest[c] = best;
  }
  if (n2) atomicAdd(a.ctr_n2, n2);
}

// ------------------------------------------------------------------ F': L2 fast path, one warp per query fragment
// All candidate regions of one fragment share its sketch Q, so a warp stages Q (plus a 1024-bucket
// directory over the top 10 hash bits) in shared memory once and gives every lane one candidate.
// Per-lane window state lives in shared memory too: gap[g] = distinct non-Q window hashes with exactly
// g query hashes below them (uint8, overflow => the candidate is handed to the exact slow kernel),
// pres = bitmap of query hashes present in the window.  Records are read as 16-byte AoS
// (hash, wpos, twin link, seqId): one load per pointer advance.
struct L2WArgs {
  const int32_t *cSeq, *cStart, *cEnd; const uint32_t *fragCandOff;
  const uint32_t *fragHash; const uint32_t *segStart; const int32_t *sCount;
  const uint4 *rec; const uint32_t *contigRecOff; uint32_t M;
  int fragLen, cmw, sLimit; uint32_t strideWords, gapWords;
  int32_t *cPos, *cBest; unsigned long long *ctr_n2;
};

__global__ void frag_cand_off_kernel(const int32_t *cFrag, uint32_t C, int32_t F, uint32_t *fragCandOff)
{
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f > F) return;
  uint32_t lo = 0, hi = C;                      // first candidate with cFrag >= f
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cFrag[mid] < f) lo = mid + 1; else hi = mid; }
  fragCandOff[f] = lo;
}

__device__ __forceinline__ uint32_t lb_rec(const uint4 *rec, uint32_t lo, uint32_t hi, int32_t v)
{
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if ((int32_t)__ldg(&rec[mid]).y < v) lo = mid + 1; else hi = mid; }
  return lo;
}

__global__ void __launch_bounds__(32)
l2_warp_kernel(const L2WArgs a)
{
  extern __shared__ __align__(16) uint32_t smem[];
  const int f = blockIdx.x, lane = threadIdx.x;
  const uint32_t c0 = a.fragCandOff[f], c1 = a.fragCandOff[f + 1];
  if (c0 == c1) return;
  const int s = a.sCount[f];
  if (s > a.sLimit || s < 1) { for (uint32_t c = c0 + lane; c < c1; c += 32) a.cBest[c] = -1; return; }
  uint32_t *Q = smem;                                          // sLimit words
  uint16_t *tab = (uint16_t *)(smem + a.sLimit);               // 1025 entries (+ pad)
  uint32_t *stBase = smem + a.sLimit + 516;
  uint32_t *st = stBase + lane * a.strideWords;                // this lane's state
  uint8_t *gap = (uint8_t *)st;                                // s + 1 counters
  uint32_t *pres = st + a.gapWords;                            // bitmap of s bits
  {
    const uint32_t *Qg = a.fragHash + a.segStart[f];
    for (int i = lane; i < s; i += 32) Q[i] = Qg[i];
    __syncwarp();
    for (int bkt = lane; bkt <= 1024; bkt += 32) {
      int lo = 0, hi = s;
      if (bkt == 1024) lo = s;
      else { const uint32_t v = (uint32_t)bkt << 22; while (lo < hi) { int mid = (lo + hi) >> 1; if (Q[mid] < v) lo = mid + 1; else hi = mid; } }
      tab[bkt] = (uint16_t)lo;
    }
    __syncwarp();
  }
  unsigned long long n2 = 0;
  for (uint32_t cb = c0; cb < c1; cb += 32) {
    const uint32_t c = cb + lane;
    const bool act = c < c1;
    for (uint32_t i = 0; i < a.strideWords; i++) st[i] = 0;
    uint32_t b = 0, e = 0, last = 0;
    int t = s, G = 0, P = 0;
    bool ovf = false;

    // One signed, branch-free update of the window state.  dir = +1: a record enters the window,
    // dir = -1: it leaves; eff: the record changes the set of DISTINCT window hashes (no twin inside).
    auto apply = [&](uint32_t h, bool eff, int dir) {
      const uint32_t bkt = h >> 22;                            // lower_bound of h in Q, inside its bucket
      int lo = tab[bkt], len = (int)tab[bkt + 1] - lo;
      while (len > 0) {
        const int half = len >> 1; const bool lt = Q[lo + half] < h;
        lo = lt ? lo + half + 1 : lo; len = lt ? len - half - 1 : half;
      }
      const int j = lo;
      const bool match = (j < s) && (Q[min(j, s - 1)] == h);
      const bool m = eff && match, nm = eff && !match;
      const uint32_t wi = (uint32_t)j >> 5, bit = 1u << (j & 31);
      const uint32_t w = pres[wi];
      if (m) pres[wi] = dir > 0 ? (w | bit) : (w & ~bit);
      const uint32_t g = gap[j];
      ovf |= nm && dir > 0 && g == 255u;
      if (nm) gap[j] = (uint8_t)(g + dir);
      const bool below = j < t;
      P += (m && below) ? dir : 0;
      G += (nm && below) ? dir : 0;
      // the pivot t moves by at most one rank per update
      const bool down = dir > 0 && (t + G > s);
      const int tt = t - (down ? 1 : 0);
      const int gv = gap[tt], pv = (int)((pres[tt >> 5] >> (tt & 31)) & 1u);
      const bool up = dir < 0 && t < s && (t + 1 + G + gv <= s);
      G += up ? gv : (down ? -gv : 0);
      P += up ? pv : (down ? -pv : 0);
      t += (up ? 1 : 0) - (down ? 1 : 0);
    };
    auto is_new = [](uint32_t link, uint32_t idx, uint32_t wb) { const uint32_t pd = link >> 16; return !(pd != 0xFFFFu && idx - pd >= wb); };
    auto is_gone = [](uint32_t link, uint32_t idx, uint32_t we) { const uint32_t nd = link & 0xFFFFu; return !(nd != 0xFFFFu && idx + nd < we); };

    if (act) {
      const int seq = a.cSeq[c];
      const uint32_t lo = a.contigRecOff[seq], hi = a.contigRecOff[seq + 1];
      b = lb_rec(a.rec, lo, hi, a.cStart[c]);
      e = lb_rec(a.rec, lo, hi, (int32_t)__ldg(&a.rec[b]).y + a.cmw);
      last = lb_rec(a.rec, lo, hi, a.cEnd[c] + a.fragLen);
    }
    const uint32_t b0 = b;
    // the first super-window
    uint32_t r = b;
    while (__any_sync(0xffffffffu, act && r < e && !ovf)) {
      if (act && r < e && !ovf) { const uint4 rc = __ldg(&a.rec[r]); apply(rc.x, is_new(rc.z, r, b), +1); r++; }
    }
    // slide
    int sw = 0, best = 0, first = 0, lastp = 0;
    // both streams are read two records ahead so a load is never consumed in the iteration that issued it
    uint4 cur = make_uint4(0, 0, 0, 0), nxt = cur, nxt2 = cur, re = cur, re2 = cur;
    bool run = act && e < last && !ovf;
    const uint32_t Mm1 = a.M - 1;
    if (run) {
      cur = __ldg(&a.rec[b]); nxt = __ldg(&a.rec[b + 1]); nxt2 = __ldg(&a.rec[min(b + 2, Mm1)]);
      re = __ldg(&a.rec[e]); re2 = __ldg(&a.rec[min(e + 1, Mm1)]); sw = (int)cur.y;
    }
    while (__any_sync(0xffffffffu, run)) {
      if (run) {
        if (P > best) { best = P; first = lastp = (int)cur.y; }
        else if (P == best) lastp = (int)cur.y;
        const int d1 = (int)nxt.y - sw, d2 = (int)re.y - (sw + a.cmw - 1);
        const int adv = min(d1, d2);
        sw += adv;
        const uint32_t ob = b, oe = e;
        const bool doRem = adv == d1, doIns = adv == d2;
        // the leaving record first (slidingMap.hpp order: delete_ref, then insert_ref); one uniform update
        // per iteration, a second one only when both ends move together
        apply(doRem ? cur.x : re.x, doRem ? is_gone(cur.z, ob, oe) : is_new(re.z, oe, ob), doRem ? -1 : +1);
        if (doRem) { b++; cur = nxt; nxt = nxt2; nxt2 = __ldg(&a.rec[min(b + 2, Mm1)]); }
        if (doRem && doIns) apply(re.x, is_new(re.z, oe, b), +1);
        if (doIns) { e++; re = re2; re2 = __ldg(&a.rec[min(e + 1, Mm1)]); }
        run = (e < last) && !ovf;
      }
    }
    if (act) {
      if (ovf) a.cBest[c] = -1;
      else { a.cPos[c] = (first + lastp) / 2; a.cBest[c] = best; n2 += last - b0; }
    }
    __syncwarp();
  }
  for (int o = 16; o; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
  if (lane == 0 && n2) atomicAdd(a.ctr_n2, n2);
}

// ------------------------------------------------------------------ G: report
struct RepArgs {
  const int32_t *cFrag, *cSeq, *cPos, *cBest; uint32_t C;
  const int32_t *sCount; const int32_t *fragSeqId;
  const uint32_t *rowOff; const float *ident, *upper; float pid; int fragLen;
};

__global__ void keep_flag_kernel(const RepArgs a, uint32_t *keep)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > a.C) return;
  if (c == a.C) { keep[c] = 0; return; }
  const int s = a.sCount[a.cFrag[c]];
  keep[c] = a.upper[a.rowOff[s] + a.cBest[c]] >= a.pid ? 1u : 0u;      // computeMap.hpp:384
}

__global__ void rows_kernel(const RepArgs a, const uint32_t *keep, const uint32_t *keepScan, bani_mapping *rows,
                            int32_t *rFrag /* optional: chunk-global fragment index per row */)
{
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.C || !keep[c]) return;
  const int f = a.cFrag[c]; const int s = a.sCount[f]; const int best = a.cBest[c];
  bani_mapping r;
  r.queryLen = a.fragLen; r.refStartPos = a.cPos[c]; r.refEndPos = a.cPos[c] + a.fragLen - 1;
  r.queryStartPos = 0; r.queryEndPos = a.fragLen - 1;
  r.refSeqId = a.cSeq[c]; r.querySeqId = a.fragSeqId[f];
  r.nucIdentity = a.ident[a.rowOff[s] + best]; r.nucIdentityUpperBound = a.upper[a.rowOff[s] + best];
  r.sketchSize = s; r.conservedSketches = best;
  rows[keepScan[c]] = r;
  if (rFrag) rFrag[keepScan[c]] = f;
}

// ------------------------------------------------------------------ H: CGI on the device
struct CgiArgs {
  const bani_mapping *rows; const int32_t *rFrag; uint32_t R;
  const int32_t *fragQuery;            // chunk-local query slot of a fragment
  const int32_t *contigGenome; const uint32_t *contigBinOff;
  int fragLen; unsigned long long totalBins; int nGenomes;
  uint32_t *table;                     // [querySlot][totalBins] float bits, 0 = empty
  uint8_t *touched;                    // [querySlot][nGenomes]
};

// 1-way: best row of each (fragment, genome) by (identity, refSeqId, refStartPos) (cgid_types.hpp:31-39,
// computeCoreIdentity.hpp:214-231); 2-way: best identity per (ref contig, position bin) (:237-254)
__global__ void cgi_scatter_kernel(const CgiArgs a)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R) return;
  const bani_mapping r = a.rows[i];
  const int f = a.rFrag[i], g = a.contigGenome[r.refSeqId];
  // rows of a fragment are contiguous and ordered by (refSeqId, refStartPos): a later row wins ties
  for (uint32_t j = i + 1; j < a.R && a.rFrag[j] == f && a.contigGenome[a.rows[j].refSeqId] == g; j++)
    if (a.rows[j].nucIdentity >= r.nucIdentity) return;
  for (uint32_t j = i; j-- > 0 && a.rFrag[j] == f && a.contigGenome[a.rows[j].refSeqId] == g;)
    if (a.rows[j].nucIdentity > r.nucIdentity) return;
  const int q = a.fragQuery[f];
  const unsigned long long bin = a.contigBinOff[r.refSeqId] + (uint32_t)(r.refStartPos / (a.fragLen - 20));
  atomicMax(a.table + (unsigned long long)q * a.totalBins + bin, __float_as_uint(r.nucIdentity));
  a.touched[(size_t)q * a.nGenomes + g] = 1;
}

// ordered float32 sum over the bins of one (query, genome) pair (computeCoreIdentity.hpp:267-297);
// clears what it read so the table is all-zero again for the next chunk
__global__ void cgi_sum_kernel(uint32_t *table, uint8_t *touched, const uint32_t *contigBinOff, const int32_t *genomeContigEnd,
                               unsigned long long totalBins, int nGenomes, int nQ, int32_t *oCount, float *oIdent)
{
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (uint32_t)nQ * (uint32_t)nGenomes) return;
  const int q = i / nGenomes, g = i % nGenomes;
  int32_t cnt = 0; float sum = 0.0f;
  if (touched[i]) {
    touched[i] = 0;
    const uint32_t b0 = contigBinOff[g ? genomeContigEnd[g - 1] : 0], b1 = contigBinOff[genomeContigEnd[g]];
    uint32_t *row = table + (unsigned long long)q * totalBins;
    for (uint32_t b = b0; b < b1; b++) { uint32_t v = row[b]; if (v) { sum += __uint_as_float(v); cnt++; row[b] = 0; } }
  }
  oCount[i] = cnt; oIdent[i] = cnt ? sum / cnt : 0.0f;
}

// ------------------------------------------------------------------ host orchestration
void Ctx::upload_lut(int smaxNeeded)
{
  lut.k = prm.kmer_size; lut.pid = prm.perc_identity;
  if (smaxNeeded <= lutUploaded && lutUploaded > 0) return;
  int target = std::max(smaxNeeded, 320);
  lut.ensure(target);
  d_minHits.alloc(lut.minHits.size(), stream); d_rowOff.alloc(lut.rowOff.size(), stream);
  d_ident.alloc(lut.ident.size(), stream); d_upper.alloc(lut.upper.size(), stream);
  BANI_CUDA(cudaMemcpyAsync(d_minHits.p, lut.minHits.data(), 4 * lut.minHits.size(), cudaMemcpyHostToDevice, stream));
  BANI_CUDA(cudaMemcpyAsync(d_rowOff.p, lut.rowOff.data(), 4 * lut.rowOff.size(), cudaMemcpyHostToDevice, stream));
  BANI_CUDA(cudaMemcpyAsync(d_ident.p, lut.ident.data(), 4 * lut.ident.size(), cudaMemcpyHostToDevice, stream));
  BANI_CUDA(cudaMemcpyAsync(d_upper.p, lut.upper.data(), 4 * lut.upper.size(), cudaMemcpyHostToDevice, stream));
  BANI_CUDA(cudaStreamSynchronize(stream));
  lutUploaded = target;
}

void map_queries(Ctx *ctx, const Index *ix, const Genome *const *queries, int32_t nq,
                 bool wantRows, bool wantCgi, MapOutput &out)
{
  cudaStream_t st = ctx->stream;
  const int k = ctx->prm.kmer_size, w = ctx->prm.window_size, fragLen = ctx->prm.frag_len;
  const float pid = ctx->prm.perc_identity;
  if (ix->device != ctx->device) fail(BANI_ERR_ARG, "index lives on another device");
  if (fragLen < 1 || fragLen > 60000) fail(BANI_ERR_LIMIT, "fragment length %d outside the supported range [1, 60000]", fragLen);
  if (wantCgi && fragLen <= 20) fail(BANI_ERR_ARG, "fragment length must exceed 20 for the identity reduction");
  const int cmw = fragLen - (w - 1) - (k - 1);         // computeMap.hpp:427
  out.totalQueryFragments.assign(nq, 0);
  out.ctr = bani_map_counters{};
  const int nG = ix->nGenomes;

  // chunk the query list: bounded fragment count and bounded 2-way table
  const uint64_t FRAG_MAX = 1u << 17;
  uint64_t qMaxByTable = nq;
  if (wantCgi && ix->totalBins) qMaxByTable = std::max<uint64_t>(1, ((uint64_t)3 << 30) / (4 * ix->totalBins));

  DevBuf<uint32_t> table; DevBuf<uint8_t> touched; DevBuf<int32_t> d_gce;
  uint64_t tableQ = 0;
  if (wantCgi) {
    d_gce.alloc(std::max(nG, 1), st);
    if (nG) BANI_CUDA(cudaMemcpyAsync(d_gce.p, ix->seqsByFile.data(), 4 * (size_t)nG, cudaMemcpyHostToDevice, st));
  }

  int q0 = 0;
  while (q0 < nq) {
    // ---- fragment table of this chunk (Map::mapQuery, computeMap.hpp:131-189)
    std::vector<SeqDesc> desc; std::vector<int32_t> fragQuery, flen;
    int q1 = q0;
    while (q1 < nq && (uint64_t)(q1 - q0) < qMaxByTable) {
      const Genome *Q = queries[q1];
      if (!Q) fail(BANI_ERR_ARG, "null genome handle");
      if (Q->device != ctx->device) fail(BANI_ERR_ARG, "genome lives on another device");
      uint64_t nf = 0;
      for (int c = 0; c < Q->nContigs; c++) { int L = Q->len[c]; if (!(L < w || L < k || L < fragLen)) nf += L / fragLen; }
      if (q1 > q0 && desc.size() + nf > FRAG_MAX) break;
      int32_t seqCounter = 0;
      for (int c = 0; c < Q->nContigs; c++) {
        const int L = Q->len[c];
        if (L < w || L < k || L < fragLen) continue;                 // :138
        const int fc = L / fragLen;                                  // :152
        for (int i = 0; i < fc; i++) {
          SeqDesc d;
          d.packed = Q->packed.p + Q->wordOff[c];
          d.nExc = (int32_t)(Q->excOff[c + 1] - Q->excOff[c]);
          d.excPos = d.nExc ? Q->excPos.p + Q->excOff[c] : nullptr;
          d.excByte = d.nExc ? Q->excByte.p + Q->excOff[c] : nullptr;
          d.startBase = i * fragLen; d.len = fragLen; d.seqId = seqCounter + i;   // :173-175
          desc.push_back(d); fragQuery.push_back(q1 - q0); flen.push_back(fragLen);
        }
        seqCounter += fc;
      }
      out.totalQueryFragments[q1] = (uint64_t)seqCounter;            // :188-189
      q1++;
    }
    const int nQc = q1 - q0;
    const int32_t F = (int32_t)desc.size();
    out.ctr.fragments += F;

    std::vector<int32_t> hCount; std::vector<float> hIdent;
    if (wantCgi) { hCount.assign((size_t)nQc * nG, 0); hIdent.assign((size_t)nQc * nG, 0.f); }

    if (F > 0 && ix->M > 0) {
      BANI_SCRATCH(SeqDesc, d_desc, F);
      BANI_CUDA(cudaMemcpyAsync(d_desc.p, desc.data(), sizeof(SeqDesc) * (size_t)F, cudaMemcpyHostToDevice, st));
      BANI_SCRATCH(int32_t, d_fragQuery, F);
      BANI_SCRATCH(int32_t, d_fragSeqId, F);
      BANI_CUDA(cudaMemcpyAsync(d_fragQuery.p, fragQuery.data(), 4 * (size_t)F, cudaMemcpyHostToDevice, st));
      { std::vector<int32_t> ids(F); for (int i = 0; i < F; i++) ids[i] = desc[i].seqId;
        BANI_CUDA(cudaMemcpyAsync(d_fragSeqId.p, ids.data(), 4 * (size_t)F, cudaMemcpyHostToDevice, st));
        BANI_CUDA(cudaStreamSynchronize(st)); }

      // ---- A: fragment sketches
      View<uint32_t> fragHash;
      BANI_SCRATCH(uint32_t, segStart, (size_t)F + 1);
      uint64_t perFrag = std::max(1, fragLen - k + 1);
      uint64_t cap = std::min<uint64_t>((uint64_t)F * perFrag, (uint64_t)F * (uint64_t)(2.6 * fragLen / (w + 1) + 64));
      uint64_t T = 0;
      for (int attempt = 0; attempt < 2; attempt++) {
        if (cap > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "query chunk produces more than 2^32 minimizers");
        fragHash = ctx->view<uint32_t>(__LINE__, std::max<uint64_t>(cap, 1));
        Stage sg(ctx, "q_sketch", (double)F * fragLen / 4.0);
        T = sketch_sequences(ctx, d_desc.p, F, flen.data(), fragHash.p, nullptr, nullptr, cap, segStart.p);
        if (T <= cap) break;
        cap = T;
      }

      // ---- B: sorted unique hashes per fragment
      BANI_SCRATCH(int32_t, sCount, F);
      DevBuf<int> d_flags(2, st);
      BANI_CUDA(cudaMemsetAsync(d_flags.p, 0, 8, st));
      { Stage sg(ctx, "q_sort_unique", 8.0 * T);
        sort_unique_kernel<<<F, SU_THREADS, 0, st>>>(fragHash.p, segStart.p, F, sCount.p, d_flags.p, d_flags.p + 1); ctx->launches++; }
      int hflags[2];
      BANI_CUDA(cudaMemcpyAsync(hflags, d_flags.p, 8, cudaMemcpyDeviceToHost, st));
      BANI_CUDA(cudaStreamSynchronize(st));
      if (hflags[1]) fail(BANI_ERR_LIMIT, "a query fragment has more than %d minimizers", SU_CAP);
      const int smax = hflags[0];
      ctx->upload_lut(smax);

      if (T > 0 && smax > 0) {
        // ---- C: lookup
        BANI_SCRATCH(uint32_t, hitLo, T + 1);
        BANI_SCRATCH(uint32_t, hitCnt, T + 1);
        BANI_SCRATCH(unsigned long long, hitOff, T + 1);
        { Stage sg(ctx, "lookup", 12.0 * T);
          lookup_kernel<<<nblk(T + 1), 256, 0, st>>>(fragHash.p, segStart.p, sCount.p, F, (uint32_t)T, ix->ukeys.p, ix->uoff.p,
                                                   ix->dir.p, ix->dirBits, hitLo.p, hitCnt.p);
          ctx->launches++;
          size_t tb = cub_scan_u64_temp(T + 1);
          BANI_SCRATCH(uint8_t, tmp, tb);
          cub_exclusive_sum_u32_to_u64(tmp.p, tb, hitCnt.p, (uint64_t *)hitOff.p, T + 1, st); }
        unsigned long long N = 0;
        BANI_CUDA(cudaMemcpyAsync(&N, hitOff.p + T, 8, cudaMemcpyDeviceToHost, st));
        BANI_CUDA(cudaStreamSynchronize(st));
        { std::vector<int32_t> hs(F); BANI_CUDA(cudaMemcpy(hs.data(), sCount.p, 4 * (size_t)F, cudaMemcpyDeviceToHost));
          for (int i = 0; i < F; i++) out.ctr.sum_s += hs[i]; }
        out.ctr.hits += N;
        if (N > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "query chunk gathers more than 2^32 index hits");

        if (N > 0) {
          // ---- D: gather + sort
          BANI_SCRATCH(unsigned long long, keysA, N);
          BANI_SCRATCH(unsigned long long, keysB, N);
          { Stage sg(ctx, "hit_gather", 12.0 * N);
            gather_kernel<<<nblk(T), 256, 0, st>>>(segStart.p, F, (uint32_t)T, hitLo.p, hitCnt.p, hitOff.p, ix->posIdx.p, keysA.p); ctx->launches++; }
          int fbits = 1; while ((1ll << fbits) < F) fbits++;
          { Stage sg(ctx, "hit_sort", 16.0 * N);
            size_t tb = cub_sort_keys_u64_temp(N);
            BANI_SCRATCH(uint8_t, tmp, tb);
            cub_sort_keys_u64(tmp.p, tb, (const uint64_t *)keysA.p, (uint64_t *)keysB.p, N, 0, 32 + fbits, st); }
          keysA.release();

          // ---- E: L1 candidate regions
          L1Args la; la.keys = keysB.p; la.N = N; la.segStart = segStart.p; la.hitOff = hitOff.p; la.sCount = sCount.p;
          la.minHits = ctx->d_minHits.p; la.recSeq = ix->seqId.p; la.recWpos = ix->wpos.p; la.fragLen = fragLen;
          BANI_SCRATCH(uint32_t, head, N + 1);
          BANI_SCRATCH(uint32_t, headScan, N + 1);
          { Stage sg(ctx, "l1_flags", 8.0 * N);
            l1_flag_kernel<<<nblk(N + 1), 256, 0, st>>>(la, head.p);
            ctx->launches++;
            size_t tb = cub_scan_u32_temp(N + 1);
            BANI_SCRATCH(uint8_t, tmp, tb);
            cub_exclusive_sum_u32(tmp.p, tb, head.p, headScan.p, N + 1, st); }
          uint32_t C = 0;
          BANI_CUDA(cudaMemcpyAsync(&C, headScan.p + N, 4, cudaMemcpyDeviceToHost, st));
          BANI_CUDA(cudaStreamSynchronize(st));
          out.ctr.candidates += C;

          if (C > 0) {
            BANI_SCRATCH(int32_t, cFrag, C);
            BANI_SCRATCH(int32_t, cSeq, C);
            BANI_SCRATCH(int32_t, cStart, C);
            BANI_SCRATCH(int32_t, cEnd, C);
            BANI_SCRATCH(int32_t, cPos, C);
            BANI_SCRATCH(int32_t, cBest, C);
            { Stage sg(ctx, "l1_write", 8.0 * N);
              l1_write_kernel<<<nblk(N), 256, 0, st>>>(la, head.p, headScan.p, cFrag.p, cSeq.p, cStart.p, cEnd.p); ctx->launches++; }
            head.release(); headScan.release(); keysB.release();

            // ---- F: L2
            L2Args l2; l2.cFrag = cFrag.p; l2.cSeq = cSeq.p; l2.cStart = cStart.p; l2.cEnd = cEnd.p; l2.C = C;
            l2.fragHash = fragHash.p; l2.segStart = segStart.p; l2.sCount = sCount.p;
            l2.recHash = ix->hash.p; l2.recWpos = ix->wpos.p; l2.recLink = ix->link.p; l2.contigRecOff = ix->contigRecOff.p;
            l2.fragLen = fragLen; l2.cmw = cmw; l2.smax = smax;
            l2.stride = ((size_t)2 * (smax + 2) + smax + 15) / 16 * 16;
            const unsigned blocks = (unsigned)std::min<uint64_t>((C + 63) / 64, (uint64_t)ctx->smCount * 16);
            BANI_SCRATCH(uint8_t, scratch, (size_t)blocks * 64 * l2.stride);
            BANI_SCRATCH(unsigned long long, d_n2, 1);
            BANI_CUDA(cudaMemsetAsync(d_n2.p, 0, 8, st));
            l2.scratch = scratch.p; l2.cPos = cPos.p; l2.cBest = cBest.p; l2.ctr_n2 = d_n2.p; l2.onlyFlagged = 1;
            {
              Stage sg(ctx, "l2");
              // fast path: one warp per fragment, window state in shared memory
              BANI_SCRATCH(uint32_t, fragCandOff, (size_t)F + 1);
              frag_cand_off_kernel<<<nblk((uint64_t)F + 1), 256, 0, st>>>(cFrag.p, C, F, fragCandOff.p);
              ctx->launches++;
              L2WArgs lw; lw.cSeq = cSeq.p; lw.cStart = cStart.p; lw.cEnd = cEnd.p; lw.fragCandOff = fragCandOff.p;
              lw.fragHash = fragHash.p; lw.segStart = segStart.p; lw.sCount = sCount.p;
              lw.rec = ix->rec.p; lw.contigRecOff = ix->contigRecOff.p; lw.M = (uint32_t)ix->M;
              lw.fragLen = fragLen; lw.cmw = cmw;
              lw.sLimit = std::min(smax, 1024);
              lw.gapWords = (uint32_t)(lw.sLimit + 1 + 3) / 4;
              lw.strideWords = (lw.gapWords + (uint32_t)(lw.sLimit + 32) / 32) | 1u;      // odd => conflict-free lanes
              lw.cPos = cPos.p; lw.cBest = cBest.p; lw.ctr_n2 = d_n2.p;
              const size_t shm = 4 * ((size_t)lw.sLimit + 516 + 32 * (size_t)lw.strideWords);
              static bool attrSet = false;
              if (!attrSet) {
                BANI_CUDA(cudaFuncSetAttribute(l2_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                // shared-memory carve-out: occupancy (window state per lane) beats L1 capacity here (measured 25/50/75/100 %)
                // and re-use each 128-byte line eight times
                int carve = 100; if (const char *ev = getenv("BANI_L2_CARVEOUT")) carve = atoi(ev);
                BANI_CUDA(cudaFuncSetAttribute(l2_warp_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
                attrSet = true;
              }
              l2_warp_kernel<<<F, 32, shm, st>>>(lw);
              ctx->launches++;
              // exact slow path for whatever the fast path flagged (uint8 counter overflow, very large sketches)
              l2_kernel<<<blocks, 64, 0, st>>>(l2); ctx->launches++;
            }

            // ---- G: report
            RepArgs ra; ra.cFrag = cFrag.p; ra.cSeq = cSeq.p; ra.cPos = cPos.p; ra.cBest = cBest.p; ra.C = C;
            ra.sCount = sCount.p; ra.fragSeqId = d_fragSeqId.p; ra.rowOff = ctx->d_rowOff.p; ra.ident = ctx->d_ident.p;
            ra.upper = ctx->d_upper.p; ra.pid = pid; ra.fragLen = fragLen;
            BANI_SCRATCH(uint32_t, keep, C + 1);
            BANI_SCRATCH(uint32_t, keepScan, C + 1);
            keep_flag_kernel<<<nblk(C + 1), 256, 0, st>>>(ra, keep.p);
            ctx->launches++;
            { size_t tb = cub_scan_u32_temp(C + 1);
            BANI_SCRATCH(uint8_t, tmp, tb);
              cub_exclusive_sum_u32(tmp.p, tb, keep.p, keepScan.p, C + 1, st); }
            uint32_t R = 0; unsigned long long n2 = 0;
            BANI_CUDA(cudaMemcpyAsync(&R, keepScan.p + C, 4, cudaMemcpyDeviceToHost, st));
            BANI_CUDA(cudaMemcpyAsync(&n2, d_n2.p, 8, cudaMemcpyDeviceToHost, st));
            BANI_CUDA(cudaStreamSynchronize(st));
            out.ctr.n2 += n2; out.ctr.mappings += R;
            if (R > 0) {
              BANI_SCRATCH(bani_mapping, rows, R);
              DevBuf<int32_t> rFrag(R, st);
              { Stage sg(ctx, "report", 44.0 * R);
                rows_kernel<<<nblk(C), 256, 0, st>>>(ra, keep.p, keepScan.p, rows.p, rFrag.p); ctx->launches++; }
              if (wantRows) {
                size_t old = out.rows.size(); out.rows.resize(old + R);
                BANI_CUDA(cudaMemcpyAsync(out.rows.data() + old, rows.p, sizeof(bani_mapping) * (size_t)R, cudaMemcpyDeviceToHost, st));
                BANI_CUDA(cudaStreamSynchronize(st));
              }
              if (wantCgi) {
                // ---- H: CGI
                if ((uint64_t)nQc > tableQ) {
                  tableQ = std::min<uint64_t>(qMaxByTable, std::max<uint64_t>(nQc, 1));
                  table.alloc((size_t)tableQ * ix->totalBins, st); touched.alloc((size_t)tableQ * std::max(nG, 1), st);
                  BANI_CUDA(cudaMemsetAsync(table.p, 0, table.bytes(), st));
                  BANI_CUDA(cudaMemsetAsync(touched.p, 0, touched.bytes(), st));
                }
                CgiArgs ca; ca.rows = rows.p; ca.rFrag = rFrag.p; ca.R = R; ca.fragQuery = d_fragQuery.p;
                ca.contigGenome = ix->contigGenome.p; ca.contigBinOff = ix->contigBinOff.p; ca.fragLen = fragLen;
                ca.totalBins = ix->totalBins; ca.nGenomes = nG; ca.table = table.p; ca.touched = touched.p;
                BANI_SCRATCH(int32_t, oCount, (size_t)nQc * nG);
                DevBuf<float> oIdent((size_t)nQc * nG, st);
                { Stage sg(ctx, "cgi", 48.0 * R);
                  cgi_scatter_kernel<<<nblk(R), 256, 0, st>>>(ca);
                  ctx->launches++;
                  cgi_sum_kernel<<<nblk((uint64_t)nQc * nG), 256, 0, st>>>(table.p, touched.p, ix->contigBinOff.p, d_gce.p,
                                                                          ix->totalBins, nG, nQc, oCount.p, oIdent.p); ctx->launches++; }
                BANI_CUDA(cudaMemcpyAsync(hCount.data(), oCount.p, 4 * (size_t)nQc * nG, cudaMemcpyDeviceToHost, st));
                BANI_CUDA(cudaMemcpyAsync(hIdent.data(), oIdent.p, 4 * (size_t)nQc * nG, cudaMemcpyDeviceToHost, st));
                BANI_CUDA(cudaStreamSynchronize(st));
              }
            }
          }
        }
      }
      BANI_CUDA(cudaGetLastError());
      BANI_CUDA(cudaStreamSynchronize(st));
    }
    if (wantCgi) {
      for (int q = 0; q < nQc; q++)
        for (int g = 0; g < nG; g++) {
          const int32_t cnt = hCount[(size_t)q * nG + g];
          if (cnt > 0) {
            bani_cgi_result r; r.refGenomeId = g; r.qryGenomeId = q0 + q; r.countSeq = cnt;
            r.totalQueryFragments = (int32_t)out.totalQueryFragments[q0 + q];       // cgid_types.hpp:73 (int)
            r.identity = hIdent[(size_t)q * nG + g];
            out.cgi.push_back(r);
          }
        }
    }
    q0 = q1;
  }
}

} // namespace bani

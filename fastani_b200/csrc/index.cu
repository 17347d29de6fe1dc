// index.cu -- HP1: reference index build.
//
// Replaces skch::Sketch::build + Sketch::index (src/map/include/winSketch.hpp:124-193):
//   build : every contig of every reference genome -> windowed minimizers (sketch.cu), written
//           already ordered by (seqId, wpos) == Sketch::minimizerIndex (winSketch.hpp:94)
//   index : the unordered_map<hash, vector<(seqId,wpos)>> (winSketch.hpp:84) becomes a stable
//           radix sort of (hash -> record index), a run-length compaction into unique keys +
//           offsets, and a bucket directory over the top bits of the hash.
// computeFreqHist (winSketch.hpp:199-248) has no effect at percentageThreshold = 0 (no
// minimizer is ever ignored) and is not reproduced.
//
// Extra, not in the reference: per record the distance (in records) to the previous / next
// record with the same hash.  The L2 stage uses it to keep SET semantics in a sliding window
// without an ordered map (slidingMap.hpp:137-200): a record entering the window adds a new
// distinct hash iff its previous twin is outside, a record leaving removes it iff its next twin
// is outside.
#include "common.cuh"
#include <algorithm>

namespace bani {

static inline unsigned nblk(uint64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__global__ void iota_kernel(uint32_t *v, uint64_t n)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = (uint32_t)i;
}

__global__ void head_flags_kernel(const uint32_t *sh, uint64_t n, uint32_t *head)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || sh[i] != sh[i - 1]) ? 1u : 0u;
}

__global__ void unique_scatter_kernel(const uint32_t *sh, const uint32_t *head, const uint32_t *scan, uint64_t n,
                                      uint32_t *ukeys, uint32_t *uoff, unsigned long long *o_U)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (head[i]) { ukeys[scan[i]] = sh[i]; uoff[scan[i]] = (uint32_t)i; }
  if (i == n - 1) { uint32_t U = scan[i] + head[i]; *o_U = U; }
}

// Twin links.  Only a few percent of the records share their hash with another record, so the link array is
// pre-filled with "no twin" (memset 0xFF) and only records that have a NEAR one are written (a 4-byte scatter).
__global__ void links_kernel(const uint32_t *sh, const uint32_t *posIdx, uint64_t n, uint32_t *link)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = sh[i];
  const bool hasPrev = i > 0 && sh[i - 1] == h, hasNext = i + 1 < n && sh[i + 1] == h;
  if (!hasPrev && !hasNext) return;
  const uint32_t r = posIdx[i];
  uint32_t pd = 0xFFFFu, nd = 0xFFFFu;
  if (hasPrev) pd = min(r - posIdx[i - 1], 0xFFFFu);       // stable sort: twins ascend by record index
  if (hasNext) nd = min(posIdx[i + 1] - r, 0xFFFFu);
  // a twin 65535 or more records away reads as "no twin" (no window is that long): in collections of related
  // genomes almost every hash recurs in a sister genome millions of records away, and none of those needs a write
  if (pd != 0xFFFFu || nd != 0xFFFFu) link[r] = (pd << 16) | nd;
}

// One 16-byte record per minimizer for the L2 stream: x = hash, y = wpos | tie << 31, z = twin link,
// w = back | fwd << 16.  back / fwd / tie describe the L2 super-window geometry of computeL2MappedRegions
// (computeMap.hpp:418-497) around this record, which does not depend on the candidate (cmw is fixed):
//   back : records strictly after the window start when this record ENTERS the window, i.e.
//          x - (UB(w_x - cmw + 1) - 1), UB = first record of the contig with wpos > v
//   fwd  : LB(w_{x+1} + cmw - 1) - x, the window end (exclusive) when this record LEAVES the window
//          (0xFFFF for the last record of a contig: it never leaves inside a scored window)
//   tie  : the record at x + fwd enters in the same step in which x leaves (wpos equal to w_{x+1} + cmw - 1)
// so the event schedule of a candidate needs no search (map.cu, l2_events_kernel).
__global__ void zip_records_kernel(const uint32_t *hash, const int32_t *wpos, const uint32_t *link, const int32_t *seqId,
                                   const uint32_t *contigRecOff, int cmw, uint64_t n, uint4 *rec, int2 *pos8, uint2 *rec8)
{
  uint64_t i64 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i64 >= n) return;
  const uint32_t i = (uint32_t)i64;
  const int32_t wx = wpos[i];
  const int seq = seqId[i];
  const uint32_t lo = contigRecOff[seq], hi = contigRecOff[seq + 1];
  uint32_t back = 0, fwd = 0xFFFFu, tie = 0;
  if (cmw >= 2) {
    { // UB(wx - cmw + 1) over [lo, i]: the answer is within cmw records of i
      const int32_t v = wx - cmw + 1;
      uint32_t l = (i - lo > (uint32_t)cmw) ? i - (uint32_t)cmw : lo, h = i;
      while (l < h) { uint32_t m = (l + h) >> 1; if (wpos[m] <= v) l = m + 1; else h = m; }
      back = min(i - l + 1, 0xFFFFu);
    }
    if (i + 1 < hi) {
      const int32_t v = wpos[i + 1] + cmw - 1;
      uint32_t l = i + 1, h = (hi - i - 1 > (uint32_t)cmw + 1) ? i + 1 + (uint32_t)cmw + 1 : hi;
      while (l < h) { uint32_t m = (l + h) >> 1; if (wpos[m] < v) l = m + 1; else h = m; }
      fwd = min(l - i, 0xFFFEu);
      tie = (l < hi && wpos[l] == v) ? 1u : 0u;
    }
  }
  rec[i] = make_uint4(hash[i], (uint32_t)wx | (tie << 31), link[i], back | (fwd << 16));
  pos8[i] = make_int2(wx, seq);
  // compact form for the staged L2 path: the two twin tests do not depend on the candidate once the record is past the
  // first window -- new distinct hash on entering iff the previous twin is further than `back`, distinct hash gone on
  // leaving iff the next twin is at least `fwd` ahead -- so two bits replace the twin distances
  const uint32_t pd = link[i] >> 16, nd = link[i] & 0xFFFFu;
  const uint32_t f14 = fwd == 0xFFFFu ? 0x3FFFu : min(fwd, 0x3FFEu);
  rec8[i] = make_uint2(hash[i], (back & 0x3FFFu) | (f14 << 14) | (tie << 28) | ((pd > back ? 1u : 0u) << 29) | ((nd >= fwd ? 1u : 0u) << 30));
}

// Per block of 1024 records: the largest `back` and the largest `fwd` (last records of a contig, which never leave, aside).
// l2_bounds_kernel takes the maximum over the blocks a candidate touches to decide whether its events fit the
// shared-memory ring of l2_events_kernel; 0xFFFF marks a block with a link that does not fit the 14-bit fields of rec8.
__global__ void block_link_max_kernel(const uint4 *rec, uint64_t n, uint32_t *blkMax)
{
  __shared__ uint32_t s_b, s_f;
  if (threadIdx.x == 0) { s_b = 0; s_f = 0; }
  __syncthreads();
  uint32_t mb = 0, mf = 0;
  for (int j = 0; j < 4; j++) {
    const uint64_t i = (uint64_t)blockIdx.x * 1024 + (uint64_t)j * 256 + threadIdx.x;
    if (i < n) {
      const uint32_t w = rec[i].w, back = w & 0xFFFFu, fwd = w >> 16;
      if (back >= 0x3FFFu || (fwd != 0xFFFFu && fwd >= 0x3FFFu)) { mb = 0xFFFFu; mf = 0xFFFFu; }
      else { mb = max(mb, back); if (fwd != 0xFFFFu) mf = max(mf, fwd); }
    }
  }
  atomicMax(&s_b, mb); atomicMax(&s_f, mf);
  __syncthreads();
  if (threadIdx.x == 0) blkMax[blockIdx.x] = min(s_b, 0xFFFFu) | (min(s_f, 0xFFFFu) << 16);
}

__global__ void dir_fill_kernel(const uint32_t *ukeys, uint32_t U, int dirBits, uint32_t *dir)
{
  // dir[b] = lower_bound(ukeys, b << (32 - dirBits)); minimizer hashes are minima of w hashes, hence heavily
  // skewed towards small values: most high buckets are empty, so every bucket searches for itself
  uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > (1u << dirBits)) return;
  uint32_t lo = 0, hi = U;
  if (b == (1u << dirBits)) lo = U;
  else { const uint32_t v = b << (32 - dirBits); while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (ukeys[mid] < v) lo = mid + 1; else hi = mid; } }
  dir[b] = lo;
}

// Probe table of the lookup stage (Index::tab): every unique hash claims one of the 4 slots of bucket (hash & mask) with
// a 64-bit compare-and-swap; keys that find their bucket full stay reachable through the sorted key array.  Which 4 keys
// of an overfull bucket get in depends on the order of the atomics, the result of a lookup does not.
__global__ void table_fill_kernel(const uint32_t *ukeys, const uint32_t *uoff, uint32_t U, uint32_t mask, uint2 *tab,
                                  uint32_t *filt, uint32_t filtMask)
{
  const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const uint32_t h = ukeys[u], o = uoff[u], cnt = uoff[u + 1] - o;
  if (filt) atomicOr(&filt[(h & filtMask) >> 5], 1u << (h & 31u));
  const unsigned long long e = ((unsigned long long)o << 32) | (unsigned long long)((h & 0xFFFFFF00u) | min(cnt, 255u));
  unsigned long long *bp = reinterpret_cast<unsigned long long *>(tab + 4 * (size_t)(h & mask));
#pragma unroll
  for (int sl = 0; sl < 4; sl++) if (atomicCAS(bp + sl, 0ull, e) == 0ull) return;
}

__global__ void fill_u32(uint32_t *p, uint32_t v, uint64_t n)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---- host-side contig tables shared by index_build and index_load: position bins (CGI), validity-bitmap bases
static unsigned long long index_contig_tables(Ctx *ctx, Index *ix, const std::vector<int32_t> &contigGenome)
{
  cudaStream_t st = ctx->stream;
  const int fragLen = ix->fragLen;
  const int32_t nC = (int32_t)ix->contigLen.size();
  std::vector<uint32_t> binOff(1, 0);
  std::vector<unsigned long long> bitBase;
  unsigned long long totalBits = 0;
  ix->totalLen = 0;
  for (int32_t c = 0; c < nC; c++) {
    const int32_t L = ix->contigLen[c];
    bitBase.push_back(totalBits); totalBits += ((unsigned long long)L + 31) & ~31ull;
    const uint64_t bins = (fragLen > 20) ? (uint64_t)L / (uint64_t)(fragLen - 20) + 1 : 1;
    if (binOff.back() + bins > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "reference too large for 32-bit position bins");
    binOff.push_back((uint32_t)(binOff.back() + bins));
    ix->totalLen += (uint64_t)L;
  }
  ix->nContigs = nC;
  ix->totalBins = binOff.back();
  ix->contigRecOff.alloc((size_t)nC + 1, st);
  ix->contigGenome.alloc(std::max(nC, 1), st);
  ix->contigBinOff.alloc((size_t)nC + 1, st);
  BANI_CUDA(cudaMemcpyAsync(ix->contigBinOff.p, binOff.data(), 4 * (size_t)(nC + 1), cudaMemcpyHostToDevice, st));
  if (nC) BANI_CUDA(cudaMemcpyAsync(ix->contigGenome.p, contigGenome.data(), 4 * (size_t)nC, cudaMemcpyHostToDevice, st));
  if (nC) {
    ix->contigBitBase.alloc(nC, st);
    BANI_CUDA(cudaMemcpyAsync(ix->contigBitBase.p, bitBase.data(), 8 * (size_t)nC, cudaMemcpyHostToDevice, st));
  }
  BANI_CUDA(cudaStreamSynchronize(st));          // the host vectors go out of scope
  return totalBits;
}

static void index_make_empty(Ctx *ctx, Index *ix)
{
  cudaStream_t st = ctx->stream;
  // empty index: every lookup misses (a shard with no references, computeCoreIdentity.hpp:468-471)
  fill_u32<<<nblk(ix->nContigs + 1), 256, 0, st>>>(ix->contigRecOff.p, 0, (uint64_t)ix->nContigs + 1);
  ctx->launches++;
  ix->M = 0; ix->U = 0;
  ix->dirBits = 8;
  ix->dir.alloc((1u << ix->dirBits) + 1, st);
  BANI_CUDA(cudaMemsetAsync(ix->dir.p, 0, 4 * ((1u << ix->dirBits) + 1), st));
  ix->ukeys.alloc(1, st); ix->uoff.alloc(2, st); ix->posIdx.alloc(1, st);
  BANI_CUDA(cudaMemsetAsync(ix->uoff.p, 0, 8, st));
  ix->tabBits = 8; ix->tab.alloc((size_t)4 << ix->tabBits, st);
  BANI_CUDA(cudaMemsetAsync(ix->tab.p, 0, ix->tab.bytes(), st));
  ix->hash.alloc(1, st); ix->wpos.alloc(1, st); ix->seqId.alloc(1, st); ix->link.alloc(1, st);
  BANI_CUDA(cudaStreamSynchronize(st));
}

// ---- Sketch::index (winSketch.hpp:181-193): the hash-ordered side and the per-record links from the position-ordered
//      records (ix->hash / wpos / seqId / contigRecOff filled, ix->M > 0)
static void index_finish(Ctx *ctx, Index *ix)
{
  cudaStream_t st = ctx->stream;
  const uint64_t M = ix->M;
  ix->link.alloc(M, st);
  ix->posIdx.alloc(M, st);
  DevBuf<uint32_t> sortedHash(M, st), iota(M, st), head(M, st), scan(M, st);
  DevBuf<unsigned long long> d_U(1, st);
  {
    Stage sg(ctx, "index_sort", 24.0 * M);
    iota_kernel<<<nblk(M), 256, 0, st>>>(iota.p, M);
    ctx->launches++;
    size_t tb = cub_sort_pairs_u32_temp(M);
    DevBuf<uint8_t> tmp(tb, st);
    cub_sort_pairs_u32(tmp.p, tb, ix->hash.p, sortedHash.p, iota.p, ix->posIdx.p, M, 32, st);
  }
  Stage sgc(ctx, "index_compact", 16.0 * M);
  head_flags_kernel<<<nblk(M), 256, 0, st>>>(sortedHash.p, M, head.p);
  ctx->launches++;
  {
    size_t tb = cub_scan_u32_temp(M);
    DevBuf<uint8_t> tmp(tb, st);
    cub_exclusive_sum_u32(tmp.p, tb, head.p, scan.p, M, st);
  }
  // U is needed to size ukeys: read scan[M-1] + head[M-1]
  uint32_t lastScan = 0, lastHead = 0;
  BANI_CUDA(cudaMemcpyAsync(&lastScan, scan.p + (M - 1), 4, cudaMemcpyDeviceToHost, st));
  BANI_CUDA(cudaMemcpyAsync(&lastHead, head.p + (M - 1), 4, cudaMemcpyDeviceToHost, st));
  BANI_CUDA(cudaStreamSynchronize(st));
  const uint64_t U = (uint64_t)lastScan + lastHead;
  ix->U = U;
  ix->ukeys.alloc(U, st); ix->uoff.alloc(U + 1, st);
  unique_scatter_kernel<<<nblk(M), 256, 0, st>>>(sortedHash.p, head.p, scan.p, M, ix->ukeys.p, ix->uoff.p, d_U.p);
  ctx->launches++;
  { uint32_t Mu = (uint32_t)M; BANI_CUDA(cudaMemcpyAsync(ix->uoff.p + U, &Mu, 4, cudaMemcpyHostToDevice, st)); }
  BANI_CUDA(cudaMemsetAsync(ix->link.p, 0xFF, 4 * (size_t)M, st));
  links_kernel<<<nblk(M), 256, 0, st>>>(sortedHash.p, ix->posIdx.p, M, ix->link.p);
  ctx->launches++;
  int bits = 8; while (bits < 24 && (1ull << bits) < U) bits++;       // <= 64 MB: the directory stays L2-resident during a lookup launch
  ix->dirBits = bits;
  ix->dir.alloc((1u << bits) + 1, st);
  dir_fill_kernel<<<nblk((1ull << bits) + 1), 256, 0, st>>>(ix->ukeys.p, (uint32_t)U, bits, ix->dir.p);
  ctx->launches++;
  {   // probe table: 2^tabBits >= U buckets (load <= 1 key per 4-slot bucket: ~2 % of the buckets are full), at most 2^28 (8.6 GB)
    int tb = 8; while (tb < 28 && (1ull << tb) < U) tb++;
    ix->tabBits = tb;
    ix->tab.alloc((size_t)4 << tb, st);
    BANI_CUDA(cudaMemsetAsync(ix->tab.p, 0, ix->tab.bytes(), st));
    int fb = 0;
    if (U <= (1ull << 26)) { fb = 8; while ((1ull << fb) < U) fb++; fb = std::min(fb + 3, 29); }      // <= 64 MB of bits
    ix->filtBits = fb;
    if (fb) { ix->filt.alloc((size_t)1 << (fb - 5), st); BANI_CUDA(cudaMemsetAsync(ix->filt.p, 0, ix->filt.bytes(), st)); }
    table_fill_kernel<<<nblk(U), 256, 0, st>>>(ix->ukeys.p, ix->uoff.p, (uint32_t)U, (1u << tb) - 1u, ix->tab.p,
                                               ix->filt.p, fb ? (uint32_t)((1ull << fb) - 1ull) : 0u);
    ctx->launches++;
  }
  ix->rec.alloc(M, st);
  ix->cmw = ix->fragLen - (ix->w - 1) - (ix->k - 1);       // computeMap.hpp:427
  ix->pos8.alloc(M, st); ix->rec8.alloc(M, st);
  zip_records_kernel<<<nblk(M), 256, 0, st>>>(ix->hash.p, ix->wpos.p, ix->link.p, ix->seqId.p, ix->contigRecOff.p, ix->cmw, M, ix->rec.p, ix->pos8.p, ix->rec8.p);
  ctx->launches++;
  ix->blkMax.alloc((size_t)((M + 1023) / 1024), st);
  block_link_max_kernel<<<(unsigned)((M + 1023) / 1024), 256, 0, st>>>(ix->rec.p, M, ix->blkMax.p);
  ctx->launches++;
  sgc.stop();
  BANI_CUDA(cudaGetLastError());
  BANI_CUDA(cudaStreamSynchronize(st));
}

Index *index_build(Ctx *ctx, Genome *const *refs, int32_t nRefs)
{
  cudaStream_t st = ctx->stream;
  const int k = ctx->prm.kmer_size, w = ctx->prm.window_size, fragLen = ctx->prm.frag_len;
  auto ix = std::make_unique<Index>();
  ix->device = ctx->device;
  ix->nGenomes = nRefs;
  ix->k = k; ix->w = w; ix->fragLen = fragLen;

  // ---- contig table over all genomes (every contig consumes a seqId: winSketch.hpp:150,164)
  std::vector<SeqDesc> desc;
  std::vector<int32_t> contigGenome;
  uint64_t totalPos = 0;
  for (int g = 0; g < nRefs; g++) {
    const Genome *G = refs[g];
    if (!G) fail(BANI_ERR_ARG, "null genome handle");
    if (G->device != ctx->device) fail(BANI_ERR_ARG, "genome lives on another device");
    ix->members.emplace(G->uid, (int32_t)desc.size());          // first occurrence wins if a genome is listed twice
    for (int c = 0; c < G->nContigs; c++) {
      SeqDesc d;
      d.packed = G->packedBase() + G->wordOff[c];
      d.nExc = (int32_t)(G->excOff[c + 1] - G->excOff[c]);
      d.excPos = d.nExc ? G->excPosBase() + G->excOff[c] : nullptr;
      d.excByte = d.nExc ? G->excByteBase() + G->excOff[c] : nullptr;
      d.startBase = 0; d.len = G->len[c]; d.seqId = (int32_t)desc.size();
      desc.push_back(d);
      ix->contigLen.push_back(G->len[c]);
      contigGenome.push_back(g);
      if (G->len[c] >= k) totalPos += G->len[c] - k + 1;
    }
    ix->seqsByFile.push_back((int32_t)desc.size());
  }
  const int32_t nC = (int32_t)desc.size();
  ctx->mark("index_build: enter + contig loop");
  const unsigned long long totalBits = index_contig_tables(ctx, ix.get(), contigGenome);
  ctx->mark("index_build: contig tables");

  if (nC == 0 || totalPos == 0) { index_make_empty(ctx, ix.get()); return ix.release(); }

  ix->validBits.alloc((size_t)(totalBits / 32) + 1, st);
  BANI_CUDA(cudaMemsetAsync(ix->validBits.p, 0, ix->validBits.bytes(), st));
  DevBuf<SeqDesc> d_desc(nC, st);
  BANI_CUDA(cudaMemcpyAsync(d_desc.p, desc.data(), sizeof(SeqDesc) * (size_t)nC, cudaMemcpyHostToDevice, st));

  // ---- build: minimizers in (seqId, wpos) order.  Expected density 2/(w+1); capacity 1.5x that,
  //      exact retry if a repetitive reference exceeds it (worst case one record per position).
  uint64_t cap = std::min<uint64_t>(totalPos, (uint64_t)(3.0 * totalPos / (w + 1)) + 65536);
  uint64_t M = 0;
  {
    DevBuf<uint32_t> th; DevBuf<int32_t> tw, ts;
    // genomes that share an upload block (host-packed ingest) form a group: one sketch launch per group, issued as
    // soon as the group's H2D copies have landed, so the upload of group g+1 overlaps the hashing of group g
    std::vector<std::pair<int32_t, int32_t>> groups;              // [first contig, end contig)
    std::vector<const Genome *> groupHead;
    {
      int32_t c = 0;
      for (int g = 0; g < nRefs; g++) {
        const Genome *G = refs[g];
        const bool newGroup = groups.empty() || G->blk != groupHead.back()->blk;       // genomes with their own buffers (no block) share one group
        if (newGroup) { groups.push_back({c, c}); groupHead.push_back(G); }
        c += G->nContigs; groups.back().second = c;
      }
    }
    for (int attempt = 0; attempt < 2; attempt++) {
      if (cap > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "more than 2^32 minimizers in one index shard");
      th.alloc(cap, st); tw.alloc(cap, st); ts.alloc(cap, st);
      Stage sg(ctx, "ref_sketch", (double)ix->totalLen / 4.0);
      M = 0;
      for (size_t gi = 0; gi < groups.size(); gi++) {
        const int32_t ca = groups[gi].first, cb = groups[gi].second;
        groupHead[gi]->wait_ready(st);
        M += sketch_sequences(ctx, d_desc.p + ca, cb - ca, ix->contigLen.data() + ca, 0, th.p, tw.p, ts.p, cap, ix->contigRecOff.p + ca,
                              ix->validBits.p, ix->contigBitBase.p + ca, M);
        if (M > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "more than 2^32 minimizers in one index shard");
      }
      sg.bytes((double)ix->totalLen / 4.0 + 12.0 * (double)M);       // packed bases in, 12-byte records out
      ctx->mark("index_build: sketched");
      if (M <= cap) break;
      cap = M;
    }
    if (M > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "more than 2^32 minimizers in one index shard");
    ix->M = M;
    if (M == 0) { index_make_empty(ctx, ix.get()); return ix.release(); }
    ix->hash.alloc(M, st); ix->wpos.alloc(M, st); ix->seqId.alloc(M, st);
    BANI_CUDA(cudaMemcpyAsync(ix->hash.p, th.p, 4 * M, cudaMemcpyDeviceToDevice, st));
    BANI_CUDA(cudaMemcpyAsync(ix->wpos.p, tw.p, 4 * M, cudaMemcpyDeviceToDevice, st));
    BANI_CUDA(cudaMemcpyAsync(ix->seqId.p, ts.p, 4 * M, cudaMemcpyDeviceToDevice, st));
  }
  index_finish(ctx, ix.get());
  ctx->mark("index_build: finished");
  return ix.release();
}

// ---------------------------------------------------------------------------------------- on-disk sketch cache
// The reference has no cache (only scripts/splitDatabase.sh + README.md:104-106: "divide the database, run as parallel
// processes"), so every run reads and sketches every reference again.  A saved index holds what only the sketch launch
// can produce -- the position-ordered (hash, wpos) records, the contig table and the validity bitmap -- as one flat
// little-endian file; the hash-ordered side, the links and the L2 records are rebuilt on the GPU at load time (a sort
// and a few passes: cheaper than reading them from disk).  A loaded index also serves as the QUERY side of its own
// genomes (qsketch_from_index): an all-vs-all run against a cache reads no FASTA at all.
//   u64 x 16 : magic, version, k, w, fragLen, M, nContigs, nGenomes, validWords, 0...
//   i32 contigLen[nContigs] | i32 seqsByFile[nGenomes] | u32 contigRecOff[nContigs+1] | u32 hash[M] | i32 wpos[M] |
//   u32 validBits[validWords] | u64 checksum (sum of all preceding 32-bit words)
static constexpr uint64_t IX_MAGIC = 0x32584449494e4142ull;

__global__ void seqid_fill_kernel(const uint32_t *contigRecOff, int32_t nC, int32_t *seqId)
{
  const int c = blockIdx.x;
  if (c >= nC) return;
  const uint32_t a = contigRecOff[c], b = contigRecOff[c + 1];
  for (uint32_t i = a + threadIdx.x; i < b; i += blockDim.x) seqId[i] = c;
}

namespace {
struct File {
  FILE *f = nullptr; std::string path; uint64_t sum = 0;
  File(const char *p, const char *mode) : f(fopen(p, mode)), path(p) { if (!f) fail(BANI_ERR_ARG, "cannot open %s", p); }
  ~File() { if (f) fclose(f); }
  void write(const void *p, size_t n) { if (n && fwrite(p, 1, n, f) != n) fail(BANI_ERR_INTERNAL, "write error on %s", path.c_str()); add(p, n); }
  void read(void *p, size_t n) { if (n && fread(p, 1, n, f) != n) fail(BANI_ERR_ARG, "%s is truncated", path.c_str()); add(p, n); }
  void add(const void *p, size_t n) { const uint32_t *w = (const uint32_t *)p; uint64_t s = 0; for (size_t i = 0; i < n / 4; i++) s += w[i]; sum += s; }
};
}

void index_save(Ctx *ctx, const Index *ix, const char *path)
{
  cudaStream_t st = ctx->stream;
  if (ix->device != ctx->device) fail(BANI_ERR_ARG, "index lives on another device");
  File f(path, "wb");
  const uint64_t M = ix->M, nC = (uint64_t)ix->nContigs, nG = (uint64_t)ix->nGenomes;
  const uint64_t validWords = M ? ix->validBits.n : 0;
  uint64_t h[16] = {IX_MAGIC, 2, (uint64_t)ix->k, (uint64_t)ix->w, (uint64_t)ix->fragLen, M, nC, nG, validWords};
  f.write(h, sizeof h);
  f.write(ix->contigLen.data(), 4 * nC);
  f.write(ix->seqsByFile.data(), 4 * nG);
  // device arrays through a pinned staging buffer
  const size_t CH = (size_t)64 << 20;
  void *stage = nullptr;
  BANI_CUDA(cudaHostAlloc(&stage, CH, cudaHostAllocDefault));
  try {
    auto dump = [&](const void *dp, uint64_t bytes) {
      for (uint64_t o = 0; o < bytes; o += CH) {
        const size_t n = (size_t)std::min<uint64_t>(CH, bytes - o);
        BANI_CUDA(cudaMemcpyAsync(stage, (const uint8_t *)dp + o, n, cudaMemcpyDeviceToHost, st));
        BANI_CUDA(cudaStreamSynchronize(st));
        f.write(stage, n);
      }
    };
    dump(ix->contigRecOff.p, 4 * (nC + 1));
    dump(ix->hash.p, 4 * M); dump(ix->wpos.p, 4 * M);
    dump(ix->validBits.p, 4 * validWords);
  } catch (...) { cudaFreeHost(stage); throw; }
  cudaFreeHost(stage);
  const uint64_t sum = f.sum;
  f.write(&sum, 8);
}

Index *index_load(Ctx *ctx, const char *path)
{
  cudaStream_t st = ctx->stream;
  File f(path, "rb");
  uint64_t h[16];
  f.read(h, sizeof h);
  if (h[0] != IX_MAGIC || h[1] != 2) fail(BANI_ERR_ARG, "%s is not a fastani_b200 index file (version 2)", path);
  const int k = ctx->prm.kmer_size, w = ctx->prm.window_size, fragLen = ctx->prm.frag_len;
  if ((int)h[2] != k || (int)h[3] != w || (int)h[4] != fragLen)
    fail(BANI_ERR_ARG, "%s was built with other parameters (k %d w %d fragLen %d), this context has k %d w %d fragLen %d",
         path, (int)h[2], (int)h[3], (int)h[4], k, w, fragLen);
  const uint64_t M = h[5], nC = h[6], nG = h[7], validWords = h[8];
  if (M > 0xfffffff0ull || nC > 0x7ffffff0ull || nG > nC + 1 || nG > 0x7ffffff0ull) fail(BANI_ERR_ARG, "%s: corrupt header", path);
  {   // the sizes in the header must match the file before anything is allocated
    const long pos = ftell(f.f);
    fseek(f.f, 0, SEEK_END);
    const uint64_t fileBytes = (uint64_t)ftell(f.f);
    fseek(f.f, pos, SEEK_SET);
    const uint64_t want = sizeof h + 4 * nC + 4 * nG + 4 * (nC + 1) + 8 * M + 4 * validWords + 8;
    if (fileBytes != want) fail(BANI_ERR_ARG, "%s: %llu bytes, header says %llu (truncated or corrupt)", path, (unsigned long long)fileBytes, (unsigned long long)want);
  }
  auto ix = std::make_unique<Index>();
  ix->device = ctx->device; ix->k = k; ix->w = w; ix->fragLen = fragLen; ix->nGenomes = (int32_t)nG;
  ix->contigLen.resize(nC); ix->seqsByFile.resize(nG);
  f.read(ix->contigLen.data(), 4 * nC);
  f.read(ix->seqsByFile.data(), 4 * nG);
  std::vector<int32_t> contigGenome(nC);
  {
    int32_t prev = 0;
    for (uint64_t g = 0; g < nG; g++) {
      const int32_t e = ix->seqsByFile[g];
      if (e < prev || (uint64_t)e > nC) fail(BANI_ERR_ARG, "%s: corrupt genome table", path);
      for (int32_t c = prev; c < e; c++) contigGenome[c] = (int32_t)g;
      prev = e;
    }
    if ((uint64_t)prev != nC) fail(BANI_ERR_ARG, "%s: corrupt genome table", path);
    for (uint64_t c = 0; c < nC; c++) if (ix->contigLen[c] < 0) fail(BANI_ERR_ARG, "%s: corrupt contig table", path);
  }
  const unsigned long long totalBits = index_contig_tables(ctx, ix.get(), contigGenome);
  if (M && validWords != totalBits / 32 + 1) fail(BANI_ERR_ARG, "%s: validity bitmap size does not match the contig table", path);
  const size_t CH = (size_t)64 << 20;
  void *stage[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  BANI_CUDA(cudaHostAlloc(&stage[0], CH, cudaHostAllocDefault));
  BANI_CUDA(cudaHostAlloc(&stage[1], CH, cudaHostAllocDefault));
  cudaEventCreate(&ev[0]); cudaEventCreate(&ev[1]);
  auto cleanup = [&] { cudaStreamSynchronize(st); cudaFreeHost(stage[0]); cudaFreeHost(stage[1]); cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]); };
  try {
    int cur = 0; bool used[2] = {false, false};
    auto slurp = [&](void *dp, uint64_t bytes) {          // disk read of chunk i+1 overlaps the H2D copy of chunk i
      for (uint64_t o = 0; o < bytes; o += CH) {
        const size_t n = (size_t)std::min<uint64_t>(CH, bytes - o);
        if (used[cur]) BANI_CUDA(cudaEventSynchronize(ev[cur]));
        f.read(stage[cur], n);
        BANI_CUDA(cudaMemcpyAsync((uint8_t *)dp + o, stage[cur], n, cudaMemcpyHostToDevice, st));
        BANI_CUDA(cudaEventRecord(ev[cur], st));
        used[cur] = true; cur ^= 1;
      }
    };
    slurp(ix->contigRecOff.p, 4 * (nC + 1));
    if (M == 0) {
      cudaStreamSynchronize(st);
      uint64_t sum = f.sum, got = 0;
      if (fread(&got, 1, 8, f.f) != 8 || got != sum) fail(BANI_ERR_ARG, "%s: checksum mismatch", path);
      cleanup();
      index_make_empty(ctx, ix.get());
      return ix.release();
    }
    ix->M = M;
    ix->hash.alloc(M, st); ix->wpos.alloc(M, st); ix->seqId.alloc(M, st);
    ix->validBits.alloc(validWords, st);
    slurp(ix->hash.p, 4 * M); slurp(ix->wpos.p, 4 * M); slurp(ix->validBits.p, 4 * validWords);
    BANI_CUDA(cudaStreamSynchronize(st));
    uint64_t sum = f.sum, got = 0;
    if (fread(&got, 1, 8, f.f) != 8 || got != sum) fail(BANI_ERR_ARG, "%s: checksum mismatch", path);
  } catch (...) { cleanup(); throw; }
  cleanup();
  // the record table must be consistent before it is used as an index: offsets ascending and ending at M
  {
    std::vector<uint32_t> ro(nC + 1);
    BANI_CUDA(cudaMemcpyAsync(ro.data(), ix->contigRecOff.p, 4 * (nC + 1), cudaMemcpyDeviceToHost, st));
    BANI_CUDA(cudaStreamSynchronize(st));
    for (uint64_t c = 0; c < nC; c++) if (ro[c] > ro[c + 1]) fail(BANI_ERR_ARG, "%s: corrupt record offsets", path);
    if (ro[0] != 0 || ro[nC] != M) fail(BANI_ERR_ARG, "%s: corrupt record offsets", path);
  }
  seqid_fill_kernel<<<(unsigned)nC, 128, 0, st>>>(ix->contigRecOff.p, (int32_t)nC, ix->seqId.p);
  ctx->launches++;
  index_finish(ctx, ix.get());
  return ix.release();
}

} // namespace bani

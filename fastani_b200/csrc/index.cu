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
                                   const uint32_t *contigRecOff, int cmw, uint64_t n, uint4 *rec)
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

__global__ void fill_u32(uint32_t *p, uint32_t v, uint64_t n)
{
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

static inline unsigned nblk(uint64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

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
  std::vector<uint32_t> binOff(1, 0);
  uint64_t totalPos = 0;
  std::vector<unsigned long long> bitBase;
  unsigned long long totalBits = 0;
  for (int g = 0; g < nRefs; g++) {
    const Genome *G = refs[g];
    if (!G) fail(BANI_ERR_ARG, "null genome handle");
    if (G->device != ctx->device) fail(BANI_ERR_ARG, "genome lives on another device");
    ix->members.emplace(G->uid, (int32_t)desc.size());          // first occurrence wins if a genome is listed twice
    for (int c = 0; c < G->nContigs; c++) {
      SeqDesc d;
      d.packed = G->packed.p + G->wordOff[c];
      d.nExc = (int32_t)(G->excOff[c + 1] - G->excOff[c]);
      d.excPos = d.nExc ? G->excPos.p + G->excOff[c] : nullptr;
      d.excByte = d.nExc ? G->excByte.p + G->excOff[c] : nullptr;
      d.startBase = 0; d.len = G->len[c]; d.seqId = (int32_t)desc.size();
      desc.push_back(d);
      ix->contigLen.push_back(G->len[c]);
      contigGenome.push_back(g);
      bitBase.push_back(totalBits); totalBits += ((unsigned long long)G->len[c] + 31) & ~31ull;
      uint64_t bins = (fragLen > 20) ? (uint64_t)G->len[c] / (uint64_t)(fragLen - 20) + 1 : 1;
      if (binOff.back() + bins > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "reference too large for 32-bit position bins");
      binOff.push_back((uint32_t)(binOff.back() + bins));
      ix->totalLen += G->len[c];
      if (G->len[c] >= k) totalPos += G->len[c] - k + 1;
    }
    ix->seqsByFile.push_back((int32_t)desc.size());
  }
  const int32_t nC = (int32_t)desc.size();
  ix->nContigs = nC;
  ix->totalBins = binOff.back();
  ix->contigRecOff.alloc((size_t)nC + 1, st);
  ix->contigGenome.alloc(std::max(nC, 1), st);
  ix->contigBinOff.alloc((size_t)nC + 1, st);
  BANI_CUDA(cudaMemcpyAsync(ix->contigBinOff.p, binOff.data(), 4 * (size_t)(nC + 1), cudaMemcpyHostToDevice, st));
  if (nC) BANI_CUDA(cudaMemcpyAsync(ix->contigGenome.p, contigGenome.data(), 4 * (size_t)nC, cudaMemcpyHostToDevice, st));

  if (nC == 0 || totalPos == 0) {
    // empty index: every lookup misses (a shard with no references, computeCoreIdentity.hpp:468-471)
    fill_u32<<<nblk(nC + 1), 256, 0, st>>>(ix->contigRecOff.p, 0, (uint64_t)nC + 1);
    ctx->launches++;
    ix->dirBits = 8;
    ix->dir.alloc((1u << ix->dirBits) + 1, st);
    BANI_CUDA(cudaMemsetAsync(ix->dir.p, 0, 4 * ((1u << ix->dirBits) + 1), st));
    ix->ukeys.alloc(1, st); ix->uoff.alloc(2, st); ix->posIdx.alloc(1, st);
    BANI_CUDA(cudaMemsetAsync(ix->uoff.p, 0, 8, st));
    ix->hash.alloc(1, st); ix->wpos.alloc(1, st); ix->seqId.alloc(1, st); ix->link.alloc(1, st);
    BANI_CUDA(cudaStreamSynchronize(st));
    return ix.release();
  }

  ix->validBits.alloc((size_t)(totalBits / 32) + 1, st);
  BANI_CUDA(cudaMemsetAsync(ix->validBits.p, 0, ix->validBits.bytes(), st));
  ix->contigBitBase.alloc(nC, st);
  BANI_CUDA(cudaMemcpyAsync(ix->contigBitBase.p, bitBase.data(), 8 * (size_t)nC, cudaMemcpyHostToDevice, st));
  DevBuf<SeqDesc> d_desc(nC, st);
  BANI_CUDA(cudaMemcpyAsync(d_desc.p, desc.data(), sizeof(SeqDesc) * (size_t)nC, cudaMemcpyHostToDevice, st));

  // ---- build: minimizers in (seqId, wpos) order.  Expected density 2/(w+1); capacity 1.5x that,
  //      exact retry if a repetitive reference exceeds it (worst case one record per position).
  uint64_t cap = std::min<uint64_t>(totalPos, (uint64_t)(3.0 * totalPos / (w + 1)) + 65536);
  uint64_t M = 0;
  {
    DevBuf<uint32_t> th; DevBuf<int32_t> tw, ts;
    for (int attempt = 0; attempt < 2; attempt++) {
      if (cap > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "more than 2^32 minimizers in one index shard");
      th.alloc(cap, st); tw.alloc(cap, st); ts.alloc(cap, st);
      Stage sg(ctx, "ref_sketch", (double)ix->totalLen / 4.0);
      M = sketch_sequences(ctx, d_desc.p, nC, ix->contigLen.data(), 0, th.p, tw.p, ts.p, cap, ix->contigRecOff.p,
                           ix->validBits.p, ix->contigBitBase.p);
      sg.bytes((double)ix->totalLen / 4.0 + 12.0 * (double)M);       // packed bases in, 12-byte records out
      if (M <= cap) break;
      cap = M;
    }
    if (M > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "more than 2^32 minimizers in one index shard");
    ix->M = M;
    size_t Ma = std::max<uint64_t>(M, 1);
    ix->hash.alloc(Ma, st); ix->wpos.alloc(Ma, st); ix->seqId.alloc(Ma, st); ix->link.alloc(Ma, st);
    if (M) {
      BANI_CUDA(cudaMemcpyAsync(ix->hash.p, th.p, 4 * M, cudaMemcpyDeviceToDevice, st));
      BANI_CUDA(cudaMemcpyAsync(ix->wpos.p, tw.p, 4 * M, cudaMemcpyDeviceToDevice, st));
      BANI_CUDA(cudaMemcpyAsync(ix->seqId.p, ts.p, 4 * M, cudaMemcpyDeviceToDevice, st));
    }
  }

  // ---- index: stable sort by hash, run-length compaction, twin links, bucket directory
  size_t Ma = std::max<uint64_t>(M, 1);
  ix->posIdx.alloc(Ma, st);
  if (M == 0) {
    ix->dirBits = 8; ix->dir.alloc((1u << 8) + 1, st);
    BANI_CUDA(cudaMemsetAsync(ix->dir.p, 0, 4 * ((1u << 8) + 1), st));
    ix->ukeys.alloc(1, st); ix->uoff.alloc(2, st);
    BANI_CUDA(cudaMemsetAsync(ix->uoff.p, 0, 8, st));
    BANI_CUDA(cudaStreamSynchronize(st));
    return ix.release();
  }
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
  ix->rec.alloc(M, st);
  ix->cmw = fragLen - (w - 1) - (k - 1);                   // computeMap.hpp:427
  zip_records_kernel<<<nblk(M), 256, 0, st>>>(ix->hash.p, ix->wpos.p, ix->link.p, ix->seqId.p, ix->contigRecOff.p, ix->cmw, M, ix->rec.p);
  ctx->launches++;
  sgc.stop();
  BANI_CUDA(cudaGetLastError());
  BANI_CUDA(cudaStreamSynchronize(st));
  return ix.release();
}

} // namespace bani

// pack.cu -- genome ingest: ASCII contigs (host) -> 2-bit packed bases + out-of-band
// exception list (device).  Replaces the per-contig makeUpperCase of the reference
// (src/map/include/commonFunc.hpp:57-66) and prepares the layout the sketch kernel reads.
//
// Layout in HBM (per genome):
//   packed   : uint32 words, 16 bases per word, base i of a contig in bits [2*(i%16), +2)
//              of word (wordOff[c] + i/16); A=0 C=1 G=2 T=3; every contig starts on a
//              16-byte boundary (wordOff multiple of 4).
//   excPos / excByte : for every byte that is not A/C/G/T after upper-casing, its
//              contig-relative position (sorted) and the upper-cased byte itself.  The
//              2-bit code stored at such a position is 0 and is never used for hashing.
#include "common.cuh"
#include <cstring>

namespace bani {

static constexpr int PACK_TILE = 4096;        // bases per block
static constexpr int PACK_THREADS = 256;      // 16 bases per thread

struct PackContig {
  int64_t   asciiOff;     // offset of the contig's first byte in the staging buffer
  uint32_t *packedDst;    // first packed word of the contig
  uint32_t *excPosDst;    // filled before pass B
  uint8_t  *excByteDst;
  uint32_t  tileOff;      // first tile of the contig in this launch
  int32_t   len;
};

__device__ __forceinline__ uint32_t upper4(uint32_t w)
{
  // per byte: if 'a' <= b <= 'z' then b -= 32 (commonFunc.hpp:61-64: b > 96 && b < 123)
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint32_t b = (w >> (8 * i)) & 0xFF;
    if (b > 96 && b < 123) b -= 32;
    r |= b << (8 * i);
  }
  return r;
}

// 16 ASCII bytes starting at contig-relative base `p0` (bounds-checked against len)
__device__ __forceinline__ void load16(const uint8_t *ascii, int64_t off, int32_t p0, int32_t len, uint32_t out[4])
{
  const uint8_t *a = ascii + off + p0;
  uintptr_t addr = (uintptr_t)a;
  const uint32_t *aw = (const uint32_t *)(addr & ~(uintptr_t)3);
  int m = (int)(addr & 3);
  int remain = len - p0;                 // valid bytes from p0
  if (remain >= 16 + 4) {
    uint32_t w0 = aw[0], w1 = aw[1], w2 = aw[2], w3 = aw[3], w4 = aw[4];
    out[0] = __funnelshift_r(w0, w1, 8 * m); out[1] = __funnelshift_r(w1, w2, 8 * m);
    out[2] = __funnelshift_r(w2, w3, 8 * m); out[3] = __funnelshift_r(w3, w4, 8 * m);
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      uint32_t w = 0;
#pragma unroll
      for (int b = 0; b < 4; b++) { int i = 4 * j + b; uint32_t c = (i < remain) ? a[i] : (uint32_t)'A'; w |= c << (8 * b); }
      out[j] = w;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; j++) out[j] = upper4(out[j]);
}

__device__ __forceinline__ int find_contig(const PackContig *c, int n, uint32_t tile)
{
  int lo = 0, hi = n - 1;                 // last contig with tileOff <= tile
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (c[mid].tileOff <= tile) lo = mid; else hi = mid - 1; }
  return lo;
}

// codes + exception mask of 16 bytes (bit i set => byte i is not A/C/G/T)
__device__ __forceinline__ void classify16(const uint32_t w[4], int valid, uint32_t &codes, uint32_t &mask)
{
  codes = 0; mask = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    uint32_t b = (w[i >> 2] >> (8 * (i & 3))) & 0xFF;
    uint32_t code = 0; bool ok = true;
    if (b == 'A') code = 0; else if (b == 'C') code = 1; else if (b == 'G') code = 2; else if (b == 'T') code = 3; else ok = false;
    codes |= code << (2 * i);
    if (!ok && i < valid) mask |= 1u << i;
  }
}

// pass A: write packed words, count exceptions per tile
__global__ void __launch_bounds__(PACK_THREADS)
pack_kernel(const uint8_t *__restrict__ ascii, const PackContig *__restrict__ contigs, int nContigs,
            uint32_t *__restrict__ tileExc)
{
  uint32_t tile = blockIdx.x;
  int ci = find_contig(contigs, nContigs, tile);
  PackContig c = contigs[ci];
  int32_t p0 = (int32_t)(tile - c.tileOff) * PACK_TILE + threadIdx.x * 16;
  uint32_t cnt = 0;
  if (p0 < c.len) {
    uint32_t w[4]; load16(ascii, c.asciiOff, p0, c.len, w);
    uint32_t codes, mask; classify16(w, min(16, c.len - p0), codes, mask);
    c.packedDst[p0 >> 4] = codes;
    cnt = __popc(mask);
  }
  // block reduce
  __shared__ uint32_t wsum[PACK_THREADS / 32];
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < PACK_THREADS / 32; i++) t += wsum[i]; tileExc[tile] = t; }
}

// pass B: write the exceptions of every tile that has some, in position order
__global__ void __launch_bounds__(PACK_THREADS)
pack_exc_kernel(const uint8_t *__restrict__ ascii, const PackContig *__restrict__ contigs, int nContigs,
                const uint32_t *__restrict__ tileExc, const uint32_t *__restrict__ tileExcOff)
{
  uint32_t tile = blockIdx.x;
  if (tileExc[tile] == 0) return;
  int ci = find_contig(contigs, nContigs, tile);
  PackContig c = contigs[ci];
  int32_t p0 = (int32_t)(tile - c.tileOff) * PACK_TILE + threadIdx.x * 16;
  uint32_t w[4] = {0, 0, 0, 0}, mask = 0, codes;
  if (p0 < c.len) { load16(ascii, c.asciiOff, p0, c.len, w); classify16(w, min(16, c.len - p0), codes, mask); }
  uint32_t cnt = __popc(mask);
  // block exclusive scan of cnt
  __shared__ uint32_t wsum[PACK_THREADS / 32];
  uint32_t incl = cnt;
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) wsum[wid] = incl;
  __syncthreads();
  uint32_t base = 0;
  for (int i = 0; i < wid; i++) base += wsum[i];
  // offset of this tile relative to the first tile of its contig => contig-local exception index
  uint32_t o = tileExcOff[tile] - tileExcOff[c.tileOff] + base + incl - cnt;
  while (mask) {
    int i = __ffs(mask) - 1; mask &= mask - 1;
    c.excPosDst[o] = (uint32_t)(p0 + i);
    c.excByteDst[o] = (uint8_t)((w[i >> 2] >> (8 * (i & 3))) & 0xFF);
    o++;
  }
}

__global__ void gather_contig_exc(const PackContig *contigs, int nContigs, const uint32_t *tileExcOff,
                                  uint32_t totalTiles, uint32_t *out /* nContigs+1 */)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nContigs) out[i] = tileExcOff[contigs[i].tileOff];
  if (i == nContigs) out[i] = tileExcOff[totalTiles];
}

void genome_create_batch(Ctx *ctx, int32_t nGenomes, const int32_t *genOff, const int64_t *off,
                         const uint8_t *seq, Genome **out)
{
  cudaStream_t st = ctx->stream;
  for (int g = 0; g < nGenomes; g++) out[g] = nullptr;
  std::vector<std::unique_ptr<Genome>> gs(nGenomes);
  // sub-batches bounded by staging size
  const int64_t STAGE_MAX = (int64_t)1 << 30;
  int g0 = 0;
  while (g0 < nGenomes) {
    int g1 = g0; int64_t bytes = 0;
    while (g1 < nGenomes) {
      int64_t b = off[genOff[g1 + 1]] - off[genOff[g1]];
      if (g1 > g0 && bytes + b > STAGE_MAX) break;
      bytes += b; g1++;
    }
    int c0 = genOff[g0], c1 = genOff[g1];
    int nC = c1 - c0;
    int64_t base = off[c0];
    if (off[c1] - base != bytes) fail(BANI_ERR_ARG, "contig offsets must be contiguous and ascending");
    DevBuf<uint8_t> stage((size_t)bytes + 64, st);
    { Stage sg(ctx, "h2d_ascii", (double)bytes);
      if (bytes) BANI_CUDA(cudaMemcpyAsync(stage.p, seq + base, (size_t)bytes, cudaMemcpyHostToDevice, st)); }
    Stage sgp(ctx, "pack", 1.25 * (double)bytes);
    BANI_CUDA(cudaMemsetAsync(stage.p + bytes, 'A', 64, st));
    std::vector<PackContig> pc(nC);
    uint64_t tiles = 0;
    for (int g = g0; g < g1; g++) {
      auto G = std::make_unique<Genome>();
      G->device = ctx->device;
      G->nContigs = genOff[g + 1] - genOff[g];
      G->len.resize(G->nContigs); G->wordOff.resize(G->nContigs); G->excOff.assign(G->nContigs + 1, 0);
      int64_t words = 0;
      for (int c = 0; c < G->nContigs; c++) {
        int64_t L = off[genOff[g] + c + 1] - off[genOff[g] + c];
        if (L < 0 || L > 0x7fffffff) fail(BANI_ERR_LIMIT, "contig length %lld exceeds the int32 offset_t of the reference", (long long)L);
        G->len[c] = (int32_t)L; G->wordOff[c] = words; G->totalLen += L;
        words += ((L + 15) / 16 + 3) / 4 * 4;
      }
      G->packed.alloc((size_t)words + 8, st);
      for (int c = 0; c < G->nContigs; c++) {
        PackContig &p = pc[genOff[g] + c - c0];
        p.asciiOff = off[genOff[g] + c] - base; p.packedDst = G->packed.p + G->wordOff[c];
        p.excPosDst = nullptr; p.excByteDst = nullptr; p.len = G->len[c];
        if (tiles > 0xfffffff0ull) fail(BANI_ERR_LIMIT, "too many pack tiles in one batch");
        p.tileOff = (uint32_t)tiles;
        tiles += (G->len[c] + PACK_TILE - 1) / PACK_TILE;
      }
      gs[g] = std::move(G);
    }
    if (nC > 0 && tiles > 0) {
      DevBuf<PackContig> d_pc(nC, st);
      BANI_CUDA(cudaMemcpyAsync(d_pc.p, pc.data(), sizeof(PackContig) * nC, cudaMemcpyHostToDevice, st));
      DevBuf<uint32_t> tileExc(tiles + 1, st), tileExcOff(tiles + 1, st);
      BANI_CUDA(cudaMemsetAsync(tileExc.p + tiles, 0, 4, st));
      pack_kernel<<<(unsigned)tiles, PACK_THREADS, 0, st>>>(stage.p, d_pc.p, nC, tileExc.p);
      ctx->launches++;
      BANI_CUDA(cudaGetLastError());
      size_t tb = cub_scan_u32_temp(tiles + 1);
      DevBuf<uint8_t> tmp(tb, st);
      cub_exclusive_sum_u32(tmp.p, tb, tileExc.p, tileExcOff.p, tiles + 1, st);
      DevBuf<uint32_t> d_cexc(nC + 1, st);
      gather_contig_exc<<<(nC + 1 + 255) / 256, 256, 0, st>>>(d_pc.p, nC, tileExcOff.p, (uint32_t)tiles, d_cexc.p);
      ctx->launches++;
      std::vector<uint32_t> cexc(nC + 1);
      BANI_CUDA(cudaMemcpyAsync(cexc.data(), d_cexc.p, 4 * (nC + 1), cudaMemcpyDeviceToHost, st));
      BANI_CUDA(cudaStreamSynchronize(st));
      // contigs with zero tiles (len 0) read the offset of the next tile: consistent with cexc
      uint32_t total = cexc[nC];
      if (total > 0) {
        for (int g = g0; g < g1; g++) {
          Genome *G = gs[g].get();
          int cb = genOff[g] - c0;
          uint32_t gstart = cexc[cb], gend = cexc[cb + G->nContigs];
          G->nExc = gend - gstart;
          for (int c = 0; c <= G->nContigs; c++) G->excOff[c] = cexc[cb + c] - gstart;
          if (G->nExc) {
            G->excPos.alloc(G->nExc, st); G->excByte.alloc(G->nExc, st);
            for (int c = 0; c < G->nContigs; c++) {
              pc[cb + c].excPosDst = G->excPos.p + G->excOff[c];
              pc[cb + c].excByteDst = G->excByte.p + G->excOff[c];
            }
          }
        }
        BANI_CUDA(cudaMemcpyAsync(d_pc.p, pc.data(), sizeof(PackContig) * nC, cudaMemcpyHostToDevice, st));
        pack_exc_kernel<<<(unsigned)tiles, PACK_THREADS, 0, st>>>(stage.p, d_pc.p, nC, tileExc.p, tileExcOff.p);
        ctx->launches++;
        BANI_CUDA(cudaGetLastError());
      }
      BANI_CUDA(cudaStreamSynchronize(st));   // pc / staging go out of scope
    }
    g0 = g1;
  }
  for (int g = 0; g < nGenomes; g++) out[g] = gs[g].release();
}

// ---- host-packed ingest ------------------------------------------------------
// The same layout produced on the HOST by the reader threads (host_pack_contig below): 0.25 bytes per base cross PCIe
// instead of 1, and nothing is left to do on the device but the copy.  The genomes of a batch are cut into upload
// groups of <= 64 MB of packed words; each group is one GenomeBlock (one copy per array on the context's copy stream,
// one event), so the index build can hash group g while group g+1 is still in flight.
void genome_create_packed_batch(Ctx *ctx, int32_t nGenomes, const int32_t *genOff, const int32_t *contigLen, const int64_t *wordOff,
                                const uint32_t *words, const int64_t *excOff, const uint32_t *excPos, const uint8_t *excByte,
                                bool async, Genome **out)
{
  cudaStream_t cs = ctx->copyStream;
  for (int g = 0; g < nGenomes; g++) out[g] = nullptr;
  std::vector<std::unique_ptr<Genome>> gs(nGenomes);
  const int32_t nC = nGenomes ? genOff[nGenomes] - genOff[0] : 0;
  const int32_t cBase = nGenomes ? genOff[0] : 0;
  // ---- validate the host tables before anything is allocated
  for (int32_t c = 0; c < nC; c++) {
    const int32_t L = contigLen[cBase + c];
    const int64_t w0 = wordOff[cBase + c], need = ((int64_t)L + 15) / 16;
    if (L < 0) fail(BANI_ERR_ARG, "negative contig length");
    if (w0 < 0 || (w0 & 3)) fail(BANI_ERR_ARG, "packed contigs must start on a 16-byte boundary (word offset multiple of 4)");
    if (c + 1 < nC && wordOff[cBase + c + 1] < w0 + need) fail(BANI_ERR_ARG, "packed contigs overlap or are out of order");
    const int64_t e0 = excOff[cBase + c], e1 = excOff[cBase + c + 1];
    if (e0 < 0 || e1 < e0) fail(BANI_ERR_ARG, "exception offsets must ascend");
    for (int64_t e = e0; e < e1; e++) {
      if (excPos[e] >= (uint32_t)L || (e > e0 && excPos[e] <= excPos[e - 1])) fail(BANI_ERR_ARG, "exception positions must ascend inside their contig");
      const uint8_t b = excByte[e];
      if (b == 'A' || b == 'C' || b == 'G' || b == 'T' || (b > 96 && b < 123)) fail(BANI_ERR_ARG, "an exception byte must be an upper-cased non-ACGT byte");
    }
  }
  const int64_t GROUP_WORDS = std::max(1ll, ctx->flags.uploadGroupWords);
  int g0 = 0;
  while (g0 < nGenomes) {
    // ---- one upload group = consecutive genomes
    int g1 = g0; int64_t span = 0;
    const int32_t ca = genOff[g0];
    const int64_t wA = genOff[g0] < genOff[nGenomes] ? wordOff[ca] : 0;
    while (g1 < nGenomes) {
      const int32_t cLast = genOff[g1 + 1] - 1;
      int64_t end = wA;
      if (cLast >= genOff[g1]) end = wordOff[cLast] + (((int64_t)contigLen[cLast] + 15) / 16 + 3) / 4 * 4;
      if (g1 > g0 && end - wA > GROUP_WORDS) break;
      span = std::max(span, end - wA);
      g1++;
    }
    const int32_t cb = genOff[g1];
    const int64_t eA = excOff[ca], eB = excOff[cb];
    auto blk = std::make_shared<GenomeBlock>();
    blk->device = ctx->device;
    blk->words.alloc((size_t)span + 8, cs);
    if (span) BANI_CUDA(cudaMemcpyAsync(blk->words.p, words + wA, 4 * (size_t)span, cudaMemcpyHostToDevice, cs));
    if (eB > eA) {
      blk->excPos.alloc((size_t)(eB - eA), cs); blk->excByte.alloc((size_t)(eB - eA), cs);
      BANI_CUDA(cudaMemcpyAsync(blk->excPos.p, excPos + eA, 4 * (size_t)(eB - eA), cudaMemcpyHostToDevice, cs));
      BANI_CUDA(cudaMemcpyAsync(blk->excByte.p, excByte + eA, (size_t)(eB - eA), cudaMemcpyHostToDevice, cs));
    }
    BANI_CUDA(cudaEventCreateWithFlags(&blk->ready, cudaEventDisableTiming));
    BANI_CUDA(cudaEventRecord(blk->ready, cs));
    for (int g = g0; g < g1; g++) {
      auto G = std::make_unique<Genome>();
      G->device = ctx->device;
      G->nContigs = genOff[g + 1] - genOff[g];
      G->len.resize(G->nContigs); G->wordOff.resize(G->nContigs); G->excOff.assign(G->nContigs + 1, 0);
      const int32_t c0 = genOff[g];
      const int64_t wG = G->nContigs ? wordOff[c0] : wA, eG = excOff[c0];
      for (int c = 0; c < G->nContigs; c++) {
        G->len[c] = contigLen[c0 + c]; G->wordOff[c] = wordOff[c0 + c] - wG; G->excOff[c] = excOff[c0 + c] - eG; G->totalLen += (uint64_t)contigLen[c0 + c];
      }
      G->excOff[G->nContigs] = excOff[c0 + G->nContigs] - eG;
      G->nExc = (uint64_t)G->excOff[G->nContigs];
      G->blk = blk;
      G->blkWords = blk->words.p + (wG - wA);
      G->blkExcPos = blk->excPos.p ? blk->excPos.p + (eG - eA) : nullptr;
      G->blkExcByte = blk->excByte.p ? blk->excByte.p + (eG - eA) : nullptr;
      gs[g] = std::move(G);
    }
    g0 = g1;
  }
  if (!async) BANI_CUDA(cudaStreamSynchronize(cs));      // the host arrays may be released when the call returns
  for (int g = 0; g < nGenomes; g++) out[g] = gs[g].release();
}

// 2-bit packing of one contig on the host: what pack_kernel / pack_exc_kernel do on the device (upper-case a-z, A C G T
// -> 0 1 2 3, every other byte -> code 0 + an out-of-band (position, upper-cased byte) entry).  words: (len + 15) / 16
// entries are written (the tail of the last word is code 0).  Returns the number of exceptions; only the first excCap are
// stored, so a caller that sees a larger count retries with bigger arrays.
namespace {
// 0x80 in every byte of v that is zero (exact per byte, no borrow between bytes)
inline uint64_t zero_bytes(uint64_t v)
{
  const uint64_t m = 0x7F7F7F7F7F7F7F7Full;
  return ~(((v & m) + m) | v | m);
}
// 0x80 in every byte of x that is NOT one of A C G T a c g t (clearing bit 5 folds the lower-case letters onto the upper-case ones,
// and nothing else onto them)
inline uint64_t non_acgt_bytes(uint64_t x)
{
  const uint64_t u = x & 0xDFDFDFDFDFDFDFDFull;
  const uint64_t ok = zero_bytes(u ^ 0x4141414141414141ull) | zero_bytes(u ^ 0x4343434343434343ull) |
                      zero_bytes(u ^ 0x4747474747474747ull) | zero_bytes(u ^ 0x5454545454545454ull);
  return ok ^ 0x8080808080808080ull;
}
// 8 ASCII bases (all of them ACGT / acgt) -> 16 bits, base j at bits 2j: A 0, C 1, G 2, T 3 = bits 1 and 2 of the byte, XORed
inline uint32_t squeeze8(uint64_t x)
{
  uint64_t c = ((x >> 1) ^ (x >> 2)) & 0x0303030303030303ull;
  c = (c | (c >> 6)) & 0x000F000F000F000Full;
  c = (c | (c >> 12)) & 0x000000FF000000FFull;
  c = (c | (c >> 24)) & 0xFFFFull;
  return (uint32_t)c;
}
}

uint64_t host_pack_contig(const uint8_t *seq, int64_t len, uint32_t *words, uint32_t *excPos, uint8_t *excByte, uint64_t excCap)
{
  static const struct Lut { uint8_t v[256]; Lut() { for (int i = 0; i < 256; i++) v[i] = 4; v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3; } } lut;
  uint64_t nExc = 0;
  const int64_t nw = (len + 15) / 16;
  for (int64_t wi = 0; wi < nw; wi++) {
    const int64_t p0 = wi * 16;
    const int n = (int)std::min<int64_t>(16, len - p0);
#if defined(__BYTE_ORDER__) && __BYTE_ORDER__ == __ORDER_LITTLE_ENDIAN__
    if (n == 16) {
      // 16 bases at a time in two 64-bit words; anything but ACGT/acgt in them sends the word down the byte path below
      uint64_t x0, x1;
      memcpy(&x0, seq + p0, 8); memcpy(&x1, seq + p0 + 8, 8);
      if ((non_acgt_bytes(x0) | non_acgt_bytes(x1)) == 0) { words[wi] = squeeze8(x0) | (squeeze8(x1) << 16); continue; }
    }
#endif
    uint32_t w = 0, any = 0;
    for (int j = 0; j < n; j++) { const uint32_t c = lut.v[seq[p0 + j]]; w |= (c & 3u) << (2 * j); any |= c; }
    if (any & 4u) {
      for (int j = 0; j < n; j++) {
        uint8_t b = seq[p0 + j];
        if (lut.v[b] == 4) {
          if (b > 96 && b < 123) b -= 32;                      // commonFunc.hpp:61-64 (only a-z; z and friends stay exceptions)
          if (nExc < excCap) { excPos[nExc] = (uint32_t)(p0 + j); excByte[nExc] = b; }
          nExc++;
        }
      }
    }
    words[wi] = w;
  }
  return nExc;
}

// ---- decode (test hook) ------------------------------------------------------
__global__ void decode_kernel(const uint32_t *packed, const uint32_t *excPos, const uint8_t *excByte, int nExc,
                              int32_t len, uint8_t *out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) { uint32_t w = packed[i >> 4]; out[i] = "ACGT"[(w >> (2 * (i & 15))) & 3]; }
}
__global__ void decode_patch_kernel(const uint32_t *excPos, const uint8_t *excByte, int nExc, uint8_t *out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nExc) out[excPos[i]] = excByte[i];
}

void genome_decode(Ctx *ctx, const Genome *g, int32_t contig, uint8_t *out, int64_t cap)
{
  if (contig < 0 || contig >= g->nContigs) fail(BANI_ERR_ARG, "contig out of range");
  int32_t L = g->len[contig];
  if (cap < L) fail(BANI_ERR_ARG, "output buffer too small");
  if (L == 0) return;
  cudaStream_t st = ctx->stream;
  DevBuf<uint8_t> d(L, st);
  int nExc = (int)(g->excOff[contig + 1] - g->excOff[contig]);
  g->wait_ready(st);
  decode_kernel<<<(L + 255) / 256, 256, 0, st>>>(g->packedBase() + g->wordOff[contig], nullptr, nullptr, 0, L, d.p);
  ctx->launches++;
  if (nExc) decode_patch_kernel<<<(nExc + 255) / 256, 256, 0, st>>>(g->excPosBase() + g->excOff[contig], g->excByteBase() + g->excOff[contig], nExc, d.p);
  ctx->launches++;
  BANI_CUDA(cudaGetLastError());
  BANI_CUDA(cudaMemcpyAsync(out, d.p, L, cudaMemcpyDeviceToHost, st));
  BANI_CUDA(cudaStreamSynchronize(st));
}

} // namespace bani

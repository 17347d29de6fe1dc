// sketch.cu -- windowed-minimizer extraction (HP1 inner loop and the query-side sketch of HP2)
//
// Replaces CommonFunc::addMinimizers (src/map/include/commonFunc.hpp:92-167) together with
// getHash (:71-81) / MurmurHash3_x64_128 (src/common/murmur3.h:226-303) and reverseComplement
// (:37-54), for whole contigs (Sketch::build, winSketch.hpp:124-176) and for query fragments
// (Map::doL1Mapping, computeMap.hpp:260).
//
// The reference runs a monotone deque over the sequence.  The data-parallel statement of the
// same function, verified record-for-record against the reference (tests/test_gpu.py::test_sketch_edge_cases_vs_reference_golden, ::test_sketch_real_genomes_sha):
//   hf(i), hb(i) = hash of the k-mer at i and of its reverse complement (ASCII bytes, seed 42)
//   valid(i)     = hf(i) != hb(i)                                  (commonFunc.hpp:131)
//   key(i)       = (min(hf,hb) << 32) | (0xFFFFFFFE - i)           valid positions only
//   m(i)         = min key over valid j in [i-w+1, i]              => smallest hash, RIGHTMOST position
//   emit at valid i >= w-1 iff m(i) != m(i'), i' = previous valid position >= w-1 (always if
//   there is none or it is >= w positions back); record = (hash(m(i)), seqId, wpos = i-w+1).
//
// One CTA owns a tile of consecutive k-mer positions of one sequence:
//   phase 1  2-bit words -> ASCII bytes in shared memory (vectorised 16-byte loads/stores,
//            PRMT table expansion), then the out-of-band non-ACGT bytes are patched in, so the
//            bytes hashed are exactly the (upper-cased) bytes the reference hashes;
//   phase 2  each thread slides a k-byte forward window and its reverse complement over 16
//            consecutive positions in registers and hashes both (2 x MurmurHash3_x64_128);
//   phase 3  sliding-window minimum over 64-bit keys: per thread, suffix-minima of the left
//            halo + prefix-minima of its own positions (w+15 shared-memory reads per 16 outputs);
//   phase 4  emission flags, CTA scan, decoupled look-back across tiles => records land in
//            global memory already ordered by (sequence, wpos); no second pass, no sort.
#include "common.cuh"

namespace bani {

static constexpr int SK_THREADS = 256;
static constexpr int SK_P = 16;                       // positions per thread
static constexpr int SK_SLOTS = SK_THREADS * SK_P;    // 4096 hash slots per CTA = halo + tile
static constexpr int SK_WMAX = 1024;
static constexpr int SK_KMAX = 32;
static constexpr int SK_ASCII = SK_SLOTS + SK_KMAX + 64;   // bytes staged per CTA (16-aligned start + slack)
// Key slots are padded by one per 16 so that the 16-position runs of consecutive lanes start in different
// banks (a lane stride of 17 keys = 34 words): without it every 64-bit key access is a 32-way conflict.
#define KIDX(q) ((q) + ((q) >> 4))
static constexpr int SK_KEYS = SK_SLOTS + SK_SLOTS / 16;

struct SketchArgs {
  const SeqDesc *desc; const uint32_t *tileOff;   // nSeq+1
  int32_t nSeq; uint32_t nTiles;
  int32_t uniformTiles;             // > 0: every sequence has this many tiles (query fragments): no tileOff table
  int k, w, tileLen;
  uint32_t *o_hash; int32_t *o_wpos; int32_t *o_seqId; uint64_t cap;
  uint32_t *o_segStart;
  unsigned long long *tileState;    // nTiles, zero-initialised
  unsigned long long *o_total;
  uint32_t *o_validBits; const unsigned long long *bitBase;   // optional: validity bitmap of the hashed positions
  unsigned long long recBase;       // records written by earlier launches into the same arrays
};

// ------------------------------------------------------------------ MurmurHash3_x64_128, low 32 bits of h1
__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
// 64 x 64 -> low 64 bits with a constant: one wide multiply + two multiply-adds chained on the high word (3 instructions;
// the compiler's own expansion spends a 4th on a separate add)
__device__ __forceinline__ uint64_t mulc(uint64_t a, uint64_t c)
{
  const uint32_t alo = (uint32_t)a, ahi = (uint32_t)(a >> 32), clo = (uint32_t)c, chi = (uint32_t)(c >> 32);
  uint32_t lo, hi;
  asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %4;\n\tmov.b64 {%0, %1}, t;\n\tmad.lo.u32 %1, %2, %5, %1;\n\tmad.lo.u32 %1, %3, %4, %1;\n\t}"
      : "=r"(lo), "=r"(hi) : "r"(alo), "r"(ahi), "r"(clo), "r"(chi));
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
  k ^= k >> 33; k = mulc(k, 0xff51afd7ed558ccdULL); k ^= k >> 33; k = mulc(k, 0xc4ceb9fe1a85ec53ULL); k ^= k >> 33;
  return k;
}
__device__ __forceinline__ uint64_t bytemask(int n) { return n >= 8 ? ~0ull : ((1ull << (8 * n)) - 1); }

// F holds the k bytes little-endian (byte i of the k-mer = byte i of the 32-byte register window)
template <int KT>
__device__ __forceinline__ uint32_t murmur32(const uint64_t F[4], int krt)
{
  const int k = KT > 0 ? KT : krt;
  const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
  uint64_t h1 = 42, h2 = 42;
  const int nblocks = k >> 4;
  if (nblocks >= 1) {
    uint64_t k1 = F[0], k2 = F[1];
    k1 = mulc(k1, c1); k1 = rotl64(k1, 31); k1 = mulc(k1, c2); h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 = mulc(k2, c2); k2 = rotl64(k2, 33); k2 = mulc(k2, c1); h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  if (nblocks == 2) {
    uint64_t k1 = F[2], k2 = F[3];
    k1 = mulc(k1, c1); k1 = rotl64(k1, 31); k1 = mulc(k1, c2); h1 ^= k1;
    h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
    k2 = mulc(k2, c2); k2 = rotl64(k2, 33); k2 = mulc(k2, c1); h2 ^= k2;
    h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
  }
  const int tail = k & 15;
  if (tail) {
    uint64_t t1 = nblocks ? F[2] : F[0], t2 = nblocks ? F[3] : F[1];
    if (tail > 8) { uint64_t k2 = t2 & bytemask(tail - 8); k2 = mulc(k2, c2); k2 = rotl64(k2, 33); k2 = mulc(k2, c1); h2 ^= k2; }
    uint64_t k1 = t1 & bytemask(tail < 8 ? tail : 8); k1 = mulc(k1, c1); k1 = rotl64(k1, 31); k1 = mulc(k1, c2); h1 ^= k1;
  }
  h1 ^= (uint64_t)k; h2 ^= (uint64_t)k;
  h1 += h2; h2 += h1;
  h1 = fmix64(h1); h2 = fmix64(h2);
  h1 += h2;
  return (uint32_t)h1;
}

// complement of A/C/G/T, identity on every other byte (commonFunc.hpp:43-50)
__device__ __forceinline__ uint32_t comp_byte(uint32_t c)
{
  uint32_t x = (c == 'A' || c == 'T') ? 0x15u : ((c == 'C' || c == 'G') ? 0x04u : 0u);
  return c ^ x;
}

template <int KT>
__device__ __forceinline__ void window_init(const uint8_t *s, int krt, uint64_t F[4], uint64_t R[4])
{
  const int k = KT > 0 ? KT : krt;
  F[0] = F[1] = F[2] = F[3] = 0; R[0] = R[1] = R[2] = R[3] = 0;
#pragma unroll
  for (int j = 0; j < (KT > 0 ? KT : SK_KMAX); j++) {
    if (j < k) {
      uint64_t c = s[j];
      F[j >> 3] |= c << (8 * (j & 7));
      int r = k - 1 - j;
      uint64_t cc = comp_byte((uint32_t)c);
#pragma unroll
      for (int q = 0; q < 4; q++) if ((r >> 3) == q) R[q] |= cc << (8 * (r & 7));
    }
  }
}

template <int KT>
__device__ __forceinline__ void window_slide(uint32_t c, int krt, uint64_t F[4], uint64_t R[4])
{
  const int k = KT > 0 ? KT : krt;
  // forward: drop byte 0, append c at byte k-1 (bytes >= k stay zero)
  F[0] = (F[0] >> 8) | (F[1] << 56);
  if (k > 8)  F[1] = (F[1] >> 8) | (F[2] << 56);
  if (k > 16) F[2] = (F[2] >> 8) | (F[3] << 56);
  if (k > 24) F[3] = (F[3] >> 8);
  {
    const int j = k - 1; uint64_t v = (uint64_t)c << (8 * (j & 7));
#pragma unroll
    for (int q = 0; q < 4; q++) if ((j >> 3) == q) F[q] |= v;
  }
  // reverse complement: shift towards higher bytes, complement of c enters at byte 0, byte k falls off
  if (k > 24) R[3] = (R[3] << 8) | (R[2] >> 56);
  if (k > 16) R[2] = (R[2] << 8) | (R[1] >> 56);
  if (k > 8)  R[1] = (R[1] << 8) | (R[0] >> 56);
  R[0] = (R[0] << 8) | (uint64_t)comp_byte(c);
  if (k < 32 && (k & 7)) {
    const uint64_t keep = bytemask(k & 7);
#pragma unroll
    for (int q = 0; q < 4; q++) if ((k >> 3) == q) R[q] &= keep;
  } else if (k < 32) {
    // k multiple of 8: byte k is byte 0 of word k>>3, which must stay zero
#pragma unroll
    for (int q = 1; q < 4; q++) if ((k >> 3) == q) R[q] = 0;
  }
}

__device__ __forceinline__ uint64_t min64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// expand 8 two-bit codes (16 bits) to 8 ASCII bytes via a PRMT table lookup
__device__ __forceinline__ void expand8(uint32_t x, uint32_t &b0, uint32_t &b1)
{
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;        // one 2-bit code per nibble
  b0 = __byte_perm(0x54474341u, 0u, x & 0xFFFFu);
  b1 = __byte_perm(0x54474341u, 0u, x >> 16);
}

template <int KT, int G>
__global__ void __launch_bounds__(SK_THREADS)
sketch_kernel(const SketchArgs a)
{
  __shared__ __align__(16) uint8_t  s_ascii[SK_ASCII];
  __shared__ __align__(16) uint64_t s_key[SK_KEYS];
  __shared__ uint16_t s_vmask[SK_THREADS];
  __shared__ uint32_t s_wsum[SK_THREADS / 32];
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const int k = KT > 0 ? KT : a.k, w = a.w;

  // ---- which sequence / tile
  int seq, t0;
  if (a.uniformTiles > 0) { seq = (int)(tile / (uint32_t)a.uniformTiles); t0 = (int)(tile % (uint32_t)a.uniformTiles) * a.tileLen; }
  else {
    int lo = 0, hi = a.nSeq - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (a.tileOff[mid] <= tile) lo = mid; else hi = mid - 1; }
    seq = lo; t0 = (int)(tile - a.tileOff[seq]) * a.tileLen;
  }
  const SeqDesc d = a.desc[seq];
  const int npos = d.len - k + 1;                               // may be <= 0
  const int hs = max(0, t0 - 2 * (w - 1));                      // first hashed position
  const int he = min(npos, t0 + a.tileLen);                     // one past the last hashed position
  const int nh = he - hs;                                       // <= SK_SLOTS

  // ---- phase 1: 2-bit -> ASCII
  int a0 = 0;   // sequence-relative position of s_ascii[0]
  if (nh > 0) {
    const int cb0 = d.startBase + hs, cb1 = d.startBase + he + k - 1;   // contig-relative byte range
    const int fw = cb0 >> 4, lw = (cb1 - 1) >> 4;
    a0 = fw * 16 - d.startBase;
    for (int j = tid; j <= lw - fw; j += SK_THREADS) {
      uint32_t wd = d.packed[fw + j];
      uint4 o; expand8(wd & 0xFFFFu, o.x, o.y); expand8(wd >> 16, o.z, o.w);
      *reinterpret_cast<uint4 *>(s_ascii + 16 * j) = o;
    }
    __syncthreads();
    if (d.nExc > 0) {
      int l = 0, r = d.nExc;
      while (l < r) { int m = (l + r) >> 1; if ((int)d.excPos[m] < cb0) l = m + 1; else r = m; }
      const int e0 = l; r = d.nExc;
      while (l < r) { int m = (l + r) >> 1; if ((int)d.excPos[m] < cb1) l = m + 1; else r = m; }
      for (int e = e0 + tid; e < l; e += SK_THREADS) s_ascii[(int)d.excPos[e] - fw * 16] = d.excByte[e];
    }
  }
  __syncthreads();

  // ---- phase 2: canonical hashes of 16 consecutive positions per thread
  const int q0 = tid * SK_P;                 // slot index of this thread's first position
  uint32_t vmask = 0;
  if (q0 < nh) {
    uint64_t F[4], R[4];
    const uint8_t *sp = s_ascii + (hs + q0 - a0);
    window_init<KT>(sp, k, F, R);
#pragma unroll 4
    for (int j = 0; j < SK_P; j++) {
      uint64_t key = ~0ull;
      if (q0 + j < nh) {
        uint32_t hf = murmur32<KT>(F, k), hb = murmur32<KT>(R, k);
        if (hf != hb) {
          key = ((uint64_t)min(hf, hb) << 32) | (uint64_t)(0xFFFFFFFEu - (uint32_t)(hs + q0 + j));
          vmask |= 1u << j;
        }
      }
      s_key[KIDX(q0 + j)] = key;
      window_slide<KT>(sp[j + k], k, F, R);
    }
  } else {
#pragma unroll
    for (int j = 0; j < SK_P; j++) s_key[KIDX(q0 + j)] = ~0ull;
  }
  s_vmask[tid] = (uint16_t)vmask;
  __syncthreads();
  if (a.o_validBits && he > t0) {
    // validity of the tile's own positions [t0, he) as aligned 32-bit words (t0 and the bit base are multiples of 32
    // resp. 16: slot q = p - hs, 16 slots per s_vmask entry)
    const int d = t0 - hs;                                  // slots of the left halo
    const int nw = (he - t0 + 31) >> 5;
    const unsigned long long wbase = (a.bitBase[seq] + (unsigned long long)t0) >> 5;
    for (int wi = tid; wi < nw; wi += SK_THREADS) {
      const int q = d + wi * 32;                            // first slot of this word
      const int e0 = q >> 4, sh = q & 15;
      unsigned long long bits = 0;
#pragma unroll
      for (int e = 0; e < 3; e++) if (e0 + e < SK_THREADS) bits |= (unsigned long long)s_vmask[e0 + e] << (16 * e);
      uint32_t word = (uint32_t)(bits >> sh);
      const int rem = he - t0 - wi * 32;
      if (rem < 32) word &= (1u << rem) - 1u;
      a.o_validBits[wbase + wi] = word;
    }
  }

  // ---- phase 3: window minima of the thread's 16 positions (kept in registers)
  uint64_t M[SK_P];
#pragma unroll
  for (int g = 0; g < SK_P; g += G) {
    const int g0 = q0 + g;
    // A = min keys[g0+G-w .. g0]   (common to all G windows of the group)
    uint64_t A = ~0ull;
    for (int q = max(0, g0 + G - w); q <= g0; q++) A = min64(A, s_key[KIDX(q)]);
    uint64_t S[G];
    S[G - 1] = A;
#pragma unroll
    for (int r = G - 1; r >= 1; r--) { int q = g0 + r - w; S[r - 1] = (q >= 0) ? min64(S[r], s_key[KIDX(q)]) : S[r]; }
    uint64_t Pr = ~0ull;
#pragma unroll
    for (int r = 0; r < G; r++) {
      if (r > 0) Pr = min64(Pr, s_key[KIDX(g0 + r)]);
      M[g + r] = min64(S[r], Pr);
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < SK_P; j++) s_key[KIDX(q0 + j)] = M[j];
  __syncthreads();

  // ---- phase 4: emission flags
  uint32_t emit = 0;
  // Common case: all 16 positions of the thread and the slot before them are valid, so the previous valid position of
  // every position is simply its left neighbour and the window minima are still in registers: m(i) != m(i-1)
  // (position w-1, the first complete window, always emits).  Anything else takes the general search below.
  const bool fastEmit = w >= 2 && vmask == 0xFFFFu && q0 > 0 && ((s_vmask[(q0 - 1) >> 4] >> ((q0 - 1) & 15)) & 1u);
  if (fastEmit) {
    uint64_t prev = s_key[KIDX(q0 - 1)];
#pragma unroll
    for (int j = 0; j < SK_P; j++) {
      const int p = hs + q0 + j;
      if (p >= t0 && p >= w - 1 && (p == w - 1 || M[j] != prev)) emit |= 1u << j;
      prev = M[j];
    }
  } else {
#pragma unroll 1
    for (uint32_t vm = vmask; vm; vm &= vm - 1) {
      const int j = __ffs(vm) - 1;
      const int q = q0 + j, p = hs + q;
      if (p < t0 || p < w - 1) continue;
      // previous valid position p' with p' >= max(w-1, p-w+1)
      const int qlo = max(max(w - 1, p - w + 1) - hs, 0);
      int qp = -1;
      {
        int wi = q >> 4; uint32_t bits = (uint32_t)s_vmask[wi] & ((1u << (q & 15)) - 1);
        while (true) {
          if (bits) { qp = (wi << 4) + (31 - __clz(bits)); break; }
          if ((wi << 4) <= qlo) break;
          wi--; bits = s_vmask[wi];
        }
        if (qp < qlo) qp = -1;
      }
      if (qp < 0 || s_key[KIDX(qp)] != s_key[KIDX(q)]) emit |= 1u << j;
    }
  }
  const uint32_t cnt = __popc(emit);
  // CTA exclusive scan
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  if (lane == 31) s_wsum[wid] = incl;
  __syncthreads();
  uint32_t wbase = 0, total = 0;
#pragma unroll
  for (int i = 0; i < SK_THREADS / 32; i++) { uint32_t v = s_wsum[i]; if (i < wid) wbase += v; total += v; }

  // ---- decoupled look-back across tiles (flag in bits 63..62: 1 = aggregate, 2 = inclusive prefix)
  if (wid == 0) {
    unsigned long long excl = 0;
    volatile unsigned long long *st = a.tileState;
    if (tile > 0) {
      if (lane == 0) st[tile] = (1ull << 62) | total;
      long long basei = (long long)tile - 1;
      while (true) {
        long long idx = basei - lane;
        unsigned long long v;
        if (idx >= 0) { do { v = st[idx]; } while ((v >> 62) == 0); } else v = (2ull << 62);
        uint32_t isP = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        unsigned long long val = v & ((1ull << 62) - 1);
        int firstP = isP ? (__ffs(isP) - 1) : 32;
        unsigned long long contrib = (lane <= firstP) ? val : 0;
#pragma unroll
        for (int o = 16; o; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (isP) break;
        basei -= 32;
      }
    }
    if (lane == 0) {
      st[tile] = (2ull << 62) | (excl + total);
      s_base = excl;
      if (t0 == 0 && a.o_segStart) a.o_segStart[seq] = (uint32_t)(a.recBase + excl);
      if (tile == a.nTiles - 1) { *a.o_total = excl + total; if (a.o_segStart) a.o_segStart[a.nSeq] = (uint32_t)(a.recBase + excl + total); }
    }
  }
  __syncthreads();
  unsigned long long o = a.recBase + s_base + wbase + (incl - cnt);
#pragma unroll 1
  for (uint32_t em = emit; em; em &= em - 1, o++) {
    const int j = __ffs(em) - 1;
    if (o < a.cap) {
      a.o_hash[o] = (uint32_t)(s_key[KIDX(q0 + j)] >> 32);
      if (a.o_wpos)  a.o_wpos[o] = hs + q0 + j - w + 1;
      if (a.o_seqId) a.o_seqId[o] = d.seqId;
    }
  }
}

template <int KT>
static void launch_sketch(const SketchArgs &a, cudaStream_t st)
{
  const int w = a.w;
  if (w >= 16)     sketch_kernel<KT, 16><<<a.nTiles, SK_THREADS, 0, st>>>(a);
  else if (w >= 8) sketch_kernel<KT, 8><<<a.nTiles, SK_THREADS, 0, st>>>(a);
  else if (w >= 4) sketch_kernel<KT, 4><<<a.nTiles, SK_THREADS, 0, st>>>(a);
  else if (w >= 2) sketch_kernel<KT, 2><<<a.nTiles, SK_THREADS, 0, st>>>(a);
  else             sketch_kernel<KT, 1><<<a.nTiles, SK_THREADS, 0, st>>>(a);
}

uint64_t sketch_sequences(Ctx *ctx, const SeqDesc *d_desc, int32_t nSeq, const int32_t *h_len, int32_t uniformLen,
                          uint32_t *o_hash, int32_t *o_wpos, int32_t *o_seqId, uint64_t cap,
                          uint32_t *o_segStart, uint32_t *o_validBits, const unsigned long long *bitBase, uint64_t recBase)
{
  cudaStream_t st = ctx->stream;
  const int k = ctx->prm.kmer_size, w = ctx->prm.window_size;
  if (k < 1 || k > SK_KMAX) fail(BANI_ERR_LIMIT, "k-mer size %d outside the supported range [1, %d]", k, SK_KMAX);
  if (w < 1 || w > SK_WMAX) fail(BANI_ERR_LIMIT, "window size %d outside the supported range [1, %d]", w, SK_WMAX);
  if (nSeq == 0) {
    if (o_segStart) { const uint32_t b = (uint32_t)recBase; BANI_CUDA(cudaMemcpyAsync(o_segStart, &b, 4, cudaMemcpyHostToDevice, st)); BANI_CUDA(cudaStreamSynchronize(st)); }
    return 0;
  }
  const int tileLen = (SK_SLOTS - 2 * (w - 1)) & ~31;        // multiple of 32: tiles write whole words of the validity bitmap
  uint64_t tiles = 0;
  int32_t uniformTiles = 0;
  DevBuf<uint32_t> d_tileOff;
  std::vector<uint32_t> tileOff;
  if (uniformLen > 0) {
    // all sequences have the same length (query fragments): tile -> sequence is a division
    const int64_t npos = (int64_t)uniformLen - k + 1;
    uniformTiles = npos <= 0 ? 1 : (int32_t)((npos + tileLen - 1) / tileLen);
    tiles = (uint64_t)uniformTiles * (uint64_t)nSeq;
    if (tiles > 0x7fffffffull) fail(BANI_ERR_LIMIT, "too many sketch tiles in one launch");
  } else {
    tileOff.resize(nSeq + 1);
    for (int i = 0; i < nSeq; i++) {
      tileOff[i] = (uint32_t)tiles;
      int64_t npos = (int64_t)h_len[i] - k + 1;
      tiles += npos <= 0 ? 1 : (uint64_t)((npos + tileLen - 1) / tileLen);
      if (tiles > 0x7fffffffull) fail(BANI_ERR_LIMIT, "too many sketch tiles in one launch");
    }
    tileOff[nSeq] = (uint32_t)tiles;
    d_tileOff.alloc(nSeq + 1, st);
    BANI_CUDA(cudaMemcpyAsync(d_tileOff.p, tileOff.data(), 4 * (size_t)(nSeq + 1), cudaMemcpyHostToDevice, st));
  }
  DevBuf<unsigned long long> state(tiles + 1, st);
  BANI_CUDA(cudaMemsetAsync(state.p, 0, 8 * (tiles + 1), st));
  SketchArgs a;
  a.desc = d_desc; a.tileOff = d_tileOff.p; a.nSeq = nSeq; a.nTiles = (uint32_t)tiles; a.uniformTiles = uniformTiles;
  a.k = k; a.w = w; a.tileLen = tileLen;
  a.o_hash = o_hash; a.o_wpos = o_wpos; a.o_seqId = o_seqId; a.cap = cap; a.o_segStart = o_segStart;
  a.tileState = state.p; a.o_total = state.p + tiles;
  a.o_validBits = o_validBits; a.bitBase = bitBase; a.recBase = recBase;
  if (k == 16) launch_sketch<16>(a, st);
  else if (k == 21) launch_sketch<21>(a, st);
  else launch_sketch<0>(a, st);
  ctx->launches++;
  BANI_CUDA(cudaGetLastError());
  unsigned long long total = 0;
  BANI_CUDA(cudaMemcpyAsync(&total, state.p + tiles, 8, cudaMemcpyDeviceToHost, st));
  BANI_CUDA(cudaStreamSynchronize(st));   // also keeps tileOff alive until the copy finished
  return total;
}

} // namespace bani

// hits.cu -- HP2 stages D+E fused per query fragment: gather the index hits of a fragment, sort them by
// (seqId, wpos) and apply the L1 candidate-region rule, all inside one CTA.
//
// Replaces the hit loop and std::sort of Map::doL1Mapping / computeL1CandidateRegions
// (src/map/include/computeMap.hpp:283-299, :320) and the region scan + merge of :322-352.
//
// The record index of a minimizer is monotone in (seqId, wpos), so sorting a fragment's hits by record index
// is the sort of :320.  A fragment has a few hundred to a few thousand hits: they fit in shared memory, so the
// device-wide 64-bit sort of (fragment, record) keys and the three passes over it (flags, scan, write) collapse
// into one kernel that reads each hit once:
//   1  gather   the position lists of the fragment's s query hashes -> shared memory
//   2  sort     block-wide radix sort on the significant bits of the record index
//   3  fetch    (seqId, wpos) of every sorted hit -> shared memory (neighbouring ranks = neighbouring records)
//   4  L1       hit i opens a raw region iff hit i+minHits-1 is on the same contig less than fragLen ahead
//               (:324-336); overlapping raw regions merge (:342-350), which is a LOCAL rule on sorted hits:
//               i is the head of a merged region unless i-1 qualifies and reaches i's start, the tail unless
//               i+1 qualifies and starts at or before i's position
//   5  emit     block scan of head flags -> candidate ordinals; {seqId, start, end} written to a staging
//               area addressed by the fragment's global hit offset (a fragment never has more regions than hits)
// Fragments are binned by hit count into four size classes (1024 / 2048 / 4096 / 8192 hits per CTA); larger
// ones (many near-identical references) stay on the device-wide sort path in map.cu.
#define BANI_FILE_TAG 1
#include "common.cuh"
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace bani {

static inline unsigned nblk(uint64_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

__global__ void frag_classify_kernel(const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                                     uint32_t *candCount, uint32_t *fragClass, uint32_t *classCount, uint32_t *classList, unsigned long long maxFast)
{
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const unsigned long long n = hitOff[segStart[f + 1]] - hitOff[segStart[f]];
  candCount[f] = 0;
  uint32_t cls = 5;                                    // no hits
  if (n > maxFast) cls = 4;                        // device-wide path
  else if (n > 4096) cls = 3;
  else if (n > 2048) cls = 2;
  else if (n > 1024) cls = 1;
  else if (n > 0) cls = 0;
  fragClass[f] = cls;
  if (cls < 5) { const uint32_t o = atomicAdd(&classCount[cls], 1u); if (cls < 4) classList[(size_t)cls * F + o] = (uint32_t)f; }
}

template <int ITEMS>
struct FragL1Smem {
  static constexpr int CAP = 256 * ITEMS;
  using Sort = cub::BlockRadixSort<uint32_t, 256, ITEMS>;
  using Scan = cub::BlockScan<uint32_t, 256>;
  union U {
    typename Sort::TempStorage sort;
    typename Scan::TempStorage scan;
    int32_t seq[CAP];
  };
};

template <int ITEMS>
__global__ void __launch_bounds__(256)
frag_l1_kernel(const FragL1Args a, const uint32_t *list, uint32_t count)
{
  using S = FragL1Smem<ITEMS>;
  constexpr int CAP = S::CAP;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint32_t *s_key = reinterpret_cast<uint32_t *>(smem_raw);                 // CAP keys, later the wpos of rank r
  typename S::U &u = *reinterpret_cast<typename S::U *>(smem_raw + sizeof(uint32_t) * CAP);
  int32_t *s_w = reinterpret_cast<int32_t *>(s_key);
  if (blockIdx.x >= count) return;
  const int f = (int)list[blockIdx.x], tid = threadIdx.x;
  const uint32_t t0 = a.segStart[f];
  const int s = a.sCount[f];
  const unsigned long long base = a.hitOff[t0];
  const int n = (int)(a.hitOff[a.segStart[f + 1]] - base);

  // ---- 1: gather (one position list per thread; lists are short)
  for (int q = tid; q < s; q += 256) {
    const uint32_t lo = a.hitLo[t0 + q], cnt = a.hitCnt[t0 + q];
    const uint32_t o = (uint32_t)(a.hitOff[t0 + q] - base);
    for (uint32_t j = 0; j < cnt; j++) s_key[o + j] = __ldg(&a.posIdx[lo + j]);
  }
  for (int i = n + tid; i < CAP; i += 256) s_key[i] = 0xFFFFFFFFu;
  __syncthreads();
  // ---- 2: sort by record index
  uint32_t keys[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) keys[i] = s_key[tid * ITEMS + i];
  __syncthreads();
  typename S::Sort(u.sort).SortBlockedToStriped(keys, 0, a.keyBits);
  __syncthreads();
  // ---- 3: (seqId, wpos) of the sorted hits; rank r = i*256 + tid
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int r = i * 256 + tid;
    if (r < n) { s_w[r] = __ldg(&a.recWpos[keys[i]]); u.seq[r] = __ldg(&a.recSeq[keys[i]]); }
  }
  __syncthreads();
  // ---- 4: L1 flags of ranks [tid*ITEMS, +ITEMS)
  const int mh = a.minHits[s];
  auto qual = [&](int i, int32_t &start) -> bool {
    const int rb = i + mh - 1;
    if (rb >= n) return false;
    if (u.seq[rb] != u.seq[i]) return false;
    const int32_t wb = s_w[rb];
    if (wb - s_w[i] >= a.fragLen) return false;
    start = max(0, wb - a.fragLen + 1);
    return true;
  };
  uint32_t headMask = 0, tailMask = 0;
  int32_t starts[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int r = tid * ITEMS + i;
    int32_t st = 0;
    starts[i] = 0;
    if (r < n && qual(r, st)) {
      starts[i] = st;
      int32_t sp, sn;
      const bool merged = r > 0 && u.seq[r - 1] == u.seq[r] && qual(r - 1, sp) && s_w[r - 1] >= st;
      const bool nextMerges = r + 1 < n && u.seq[r + 1] == u.seq[r] && qual(r + 1, sn) && s_w[r] >= sn;
      if (!merged) headMask |= 1u << i;
      if (!nextMerges) tailMask |= 1u << i;
    }
  }
  // values needed after the scan reuses the union: read them first
  int32_t hSeq[ITEMS], tEnd[ITEMS];
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const int r = tid * ITEMS + i;
    hSeq[i] = ((headMask >> i) & 1u) ? u.seq[r] : 0;
    tEnd[i] = ((tailMask >> i) & 1u) ? s_w[r] : 0;
  }
  __syncthreads();
  // ---- 5: ordinals + staging writes
  uint32_t excl, total;
  typename S::Scan(u.scan).ExclusiveSum((uint32_t)__popc(headMask), excl, total);
  uint32_t ord = excl;
#pragma unroll
  for (int i = 0; i < ITEMS; i++) {
    const bool hd = (headMask >> i) & 1u, tl = (tailMask >> i) & 1u;
    if (hd) { a.stSeq[base + ord] = hSeq[i]; a.stStart[base + ord] = starts[i]; }
    ord += hd ? 1u : 0u;
    if (tl) a.stEnd[base + ord - 1] = tEnd[i];
  }
  if (tid == 0) a.candCount[f] = total;
}

template <int ITEMS>
static void launch_class(const FragL1Args &a, const uint32_t *list, uint32_t count, cudaStream_t st)
{
  if (count == 0) return;
  using S = FragL1Smem<ITEMS>;
  const size_t shm = sizeof(uint32_t) * S::CAP + sizeof(typename S::U);
  static bool attr = false;
  if (!attr) { BANI_CUDA(cudaFuncSetAttribute(frag_l1_kernel<ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); attr = true; }
  frag_l1_kernel<ITEMS><<<count, 256, shm, st>>>(a, list, count);
}

void frag_classify(Ctx *ctx, const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                   uint32_t *candCount, uint32_t *fragClass, uint32_t *classCount /* 8, zeroed here */, uint32_t *classList /* 4*F */,
                   unsigned long long maxFast)
{
  BANI_CUDA(cudaMemsetAsync(classCount, 0, 8 * sizeof(uint32_t), ctx->stream));
  frag_classify_kernel<<<nblk(F), 256, 0, ctx->stream>>>(segStart, hitOff, F, candCount, fragClass, classCount, classList, maxFast);
  ctx->launches++;
}

void frag_l1_fast(Ctx *ctx, const FragL1Args &a, const uint32_t *classList, const uint32_t classCount[4])
{
  cudaStream_t st = ctx->stream;
  const size_t F = (size_t)a.F;
  launch_class<4>(a, classList + 0 * F, classCount[0], st);
  launch_class<8>(a, classList + 1 * F, classCount[1], st);
  launch_class<16>(a, classList + 2 * F, classCount[2], st);
  launch_class<32>(a, classList + 3 * F, classCount[3], st);
  for (int i = 0; i < 4; i++) if (classCount[i]) ctx->launches++;
  BANI_CUDA(cudaGetLastError());
}

// ---- candidates of the device-wide path -> the same staging area (candidate c of fragment f lands at
//      hitOff[segStart[f]] + its ordinal inside the fragment)
__global__ void cand_stage_kernel(const int32_t *cFrag, const int32_t *cSeq, const int32_t *cStart, const int32_t *cEnd, uint32_t C,
                                  const uint32_t *segStart, const unsigned long long *hitOff,
                                  int32_t *stSeq, int32_t *stStart, int32_t *stEnd, uint32_t *candCount)
{
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int f = cFrag[c];
  uint32_t lo = 0, hi = c;                       // first candidate of fragment f
  while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (cFrag[mid] < f) lo = mid + 1; else hi = mid; }
  const unsigned long long o = hitOff[segStart[f]] + (c - lo);
  stSeq[o] = cSeq[c]; stStart[o] = cStart[c]; stEnd[o] = cEnd[c];
  atomicAdd(&candCount[f], 1u);
}

void cand_stage(Ctx *ctx, const int32_t *cFrag, const int32_t *cSeq, const int32_t *cStart, const int32_t *cEnd, uint32_t C,
                const uint32_t *segStart, const unsigned long long *hitOff, int32_t *stSeq, int32_t *stStart, int32_t *stEnd,
                uint32_t *candCount)
{
  if (!C) return;
  cand_stage_kernel<<<nblk(C), 256, 0, ctx->stream>>>(cFrag, cSeq, cStart, cEnd, C, segStart, hitOff, stSeq, stStart, stEnd, candCount);
  ctx->launches++;
}

// ---- staging -> dense candidate arrays in (fragment, position) order; one warp per fragment
__global__ void cand_compact_kernel(const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                                    const uint32_t *candCount, const uint32_t *candOff,
                                    const int32_t *stSeq, const int32_t *stStart, const int32_t *stEnd,
                                    int32_t *cFrag, int32_t *cSeq, int32_t *cStart, int32_t *cEnd)
{
  const int f = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (f >= F) return;
  const uint32_t n = candCount[f];
  if (!n) return;
  const unsigned long long src = hitOff[segStart[f]];
  const uint32_t dst = candOff[f];
  for (uint32_t i = lane; i < n; i += 32) {
    cFrag[dst + i] = f; cSeq[dst + i] = stSeq[src + i]; cStart[dst + i] = stStart[src + i]; cEnd[dst + i] = stEnd[src + i];
  }
}

void cand_compact(Ctx *ctx, const uint32_t *segStart, const unsigned long long *hitOff, int32_t F,
                  const uint32_t *candCount, const uint32_t *candOff, const int32_t *stSeq, const int32_t *stStart,
                  const int32_t *stEnd, int32_t *cFrag, int32_t *cSeq, int32_t *cStart, int32_t *cEnd)
{
  cand_compact_kernel<<<nblk((uint64_t)F * 32), 256, 0, ctx->stream>>>(segStart, hitOff, F, candCount, candOff, stSeq, stStart, stEnd,
                                                                         cFrag, cSeq, cStart, cEnd);
  ctx->launches++;
}

} // namespace bani

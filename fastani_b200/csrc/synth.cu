// synth.cu -- deterministic synthetic genomes for the benchmark harness (not part of the hot path).
// Counter-based: base p of a genome is a pure function of (seed, ancestor, strain, rate, p), so the
// numpy twin in fastani_b200/synth.py produces identical bytes for the CPU baseline and the tests.
#include "common.cuh"

namespace bani {

__host__ __device__ inline uint64_t splitmix64(uint64_t x)
{
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

__global__ void synth_kernel(uint64_t seed, uint32_t ancestor, uint32_t strain, uint32_t ppm, int64_t len, uint8_t *out)
{
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const uint64_t ka = splitmix64(seed ^ (0xA5A5A5A5ull + ((uint64_t)ancestor << 32)));
  const uint64_t ks = splitmix64(seed ^ (0x5A5A5A5Aull + ((uint64_t)ancestor << 32) + ((uint64_t)strain << 8) + 1));
  uint32_t base = (uint32_t)(splitmix64(ka + (uint64_t)i) >> 62);
  const uint64_t u = splitmix64(ks + (uint64_t)i);
  // substitution with probability ppm / 2^20-ish million: compare 20 high bits scaled
  const uint32_t r = (uint32_t)(u >> 40) % 1000000u;
  if (strain != 0 && r < ppm) base = (base + 1 + (uint32_t)((u >> 8) % 3u)) & 3u;
  out[i] = "ACGT"[base];
}

void synth_genome(Ctx *ctx, uint64_t seed, uint32_t ancestor, uint32_t strain, uint32_t ppm, int64_t len, uint8_t *hostOut)
{
  if (len <= 0) return;
  cudaStream_t st = ctx->stream;
  DevBuf<uint8_t> d((size_t)len, st);
  synth_kernel<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(seed, ancestor, strain, ppm, len, d.p);
  ctx->launches++;
  BANI_CUDA(cudaGetLastError());
  BANI_CUDA(cudaMemcpyAsync(hostOut, d.p, (size_t)len, cudaMemcpyDeviceToHost, st));
  BANI_CUDA(cudaStreamSynchronize(st));
}

} // namespace bani

// alloc.cpp -- caching device allocator behind DevBuf.
//
// The hot path re-creates the same multi-GB buffers every step (index arrays, sort scratch).  Going to
// the driver for them (cudaMalloc / cudaMallocAsync + pool trimming) costs up to hundreds of milliseconds
// and is not repeatable, so freed blocks are kept per (device, stream, size) and handed out again; the
// driver is only asked on a miss, and everything cached is returned to it when an allocation fails or
// when the last context of the device goes away.  Reuse is stream-ordered: a block is only ever reused
// on the stream it was freed on, so no event is needed.
#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include "common.cuh"

namespace bani {

namespace {
struct Key {
  int dev; cudaStream_t st; size_t bytes;
  bool operator<(const Key &o) const { return std::tie(dev, st, bytes) < std::tie(o.dev, o.st, o.bytes); }
};
std::mutex g_mu;
std::multimap<Key, void *> g_free;

size_t round_size(size_t b)
{
  if (b <= 4096) return (b + 255) & ~(size_t)255;
  if (b <= (1u << 20)) return (b + 4095) & ~(size_t)4095;
  return (b + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
}
}

uint64_t next_genome_uid()
{
  static std::atomic<uint64_t> n(1);
  return n++;
}

void dev_cache_flush(int dev)
{
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto it = g_free.begin(); it != g_free.end();) {
    if (dev < 0 || it->first.dev == dev) { cudaFree(it->second); it = g_free.erase(it); }
    else ++it;
  }
}

void *dev_alloc(size_t bytes, cudaStream_t st, size_t *granted, int *devOut)
{
  int dev = 0; cudaGetDevice(&dev);
  *devOut = dev;
  const size_t rb = round_size(bytes);
  *granted = rb;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_free.find(Key{dev, st, rb});
    if (it != g_free.end()) { void *p = it->second; g_free.erase(it); return p; }
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, rb);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    cudaDeviceSynchronize();
    dev_cache_flush(dev);
    e = cudaMalloc(&p, rb);
    if (e != cudaSuccess) { (void)cudaGetLastError(); fail(BANI_ERR_NOMEM, "device allocation of %zu bytes failed: %s", rb, cudaGetErrorString(e)); }
  }
  return p;
}

void dev_free(void *p, size_t granted, cudaStream_t st, int dev)
{
  if (!p) return;
  // keyed by the device the block was ALLOCATED on (recorded in DevBuf), not by the caller's current device: a host
  // thread that drives several GPUs may destroy an object of device A while device B is current
  std::lock_guard<std::mutex> lk(g_mu);
  g_free.emplace(Key{dev, st, granted}, p);
}

} // namespace bani

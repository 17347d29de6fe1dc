/* oracle/shim/gsl/gsl_cdf.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Stand-in for the one GNU GSL symbol the reference calls,
 *   gsl_cdf_binomial_Q(k, p, n) = P[X > k],  X ~ Binomial(n, p)
 * (call sites: /root/reference/src/map/include/map_stats.hpp:96 and :206).
 * GSL is not vendored in /root/reference and is absent from this image
 * ("GSL 1.6+", unpinned: /root/reference/CMakeLists.txt:34), so the published
 * definition of the upper binomial tail is restated here as a direct log-space
 * sum in long double.  The reference consumes the value only through
 * comparisons (`< 0.05`, `<= 1e-3`), which are far from the ~1e-16 relative
 * error of this sum.
 *
 * Memoised on (k, n, bits(p)): un-memoised the reference runs >2x slower.
 */
#ifndef ORACLE_GSL_CDF_SHIM_H
#define ORACLE_GSL_CDF_SHIM_H

#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef __cplusplus
#include <unordered_map>
#include <mutex>
#endif

static inline double oracle_binomial_Q_raw(unsigned k, double p, unsigned n)
{
  if (k >= n) return 0.0;
  if (p <= 0.0) return 0.0;
  if (p >= 1.0) return 1.0;
  long double lp = logl((long double)p), lq = log1pl(-(long double)p);
  long double lgn = lgammal((long double)n + 1.0L);
  long double acc = 0.0L;
  for (unsigned i = k + 1; i <= n; i++) {
    long double t = lgn - lgammal((long double)i + 1.0L) - lgammal((long double)(n - i) + 1.0L)
                    + (long double)i * lp + (long double)(n - i) * lq;
    acc += expl(t);
  }
  if (acc > 1.0L) acc = 1.0L;
  return (double)acc;
}

#ifdef __cplusplus
static inline double gsl_cdf_binomial_Q(unsigned k, double p, unsigned n)
{
  struct Key { uint64_t a, b; bool operator==(const Key &o) const { return a == o.a && b == o.b; } };
  struct H { size_t operator()(const Key &x) const { return (size_t)(x.a * 0x9E3779B97F4A7C15ull ^ (x.b + (x.a >> 7))); } };
  static thread_local std::unordered_map<Key, double, H> memo;
  uint64_t pb; memcpy(&pb, &p, 8);
  Key key{((uint64_t)k << 32) | n, pb};
  auto it = memo.find(key);
  if (it != memo.end()) return it->second;
  double v = oracle_binomial_Q_raw(k, p, n);
  memo.emplace(key, v);
  return v;
}
#else
static inline double gsl_cdf_binomial_Q(unsigned k, double p, unsigned n)
{
  return oracle_binomial_Q_raw(k, p, n);
}
#endif

#endif

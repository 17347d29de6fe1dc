/* oracle/binom.h -- TEST INFRASTRUCTURE ONLY.
 * Upper binomial tail P[X > k], X ~ Bin(n, p): the published definition of
 * GSL's gsl_cdf_binomial_Q (third-party, absent from /root/reference and from
 * this image; call sites /root/reference/src/map/include/map_stats.hpp:96,206).
 * Direct log-space sum in long double.
 */
#ifndef ORACLE_BINOM_H
#define ORACLE_BINOM_H
#include <math.h>

static inline double orc_binomial_Q(unsigned k, double p, unsigned n)
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
#endif

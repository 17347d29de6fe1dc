// stats.cpp -- host-side statistics of the mapper (skch::Stat, src/map/include/map_stats.hpp)
// and the lookup tables the mapping kernels consume.
//
// Every result of these functions is used by the hot path only as an integer
// threshold (minimum hits), a pass/fail flag (upper bound >= cutoff) or a float32
// identity that is a pure function of (shared, s, k) -- so they are evaluated once
// on the host into tables (StatLut) and the kernels index them.
//
// The expression TYPES follow the reference exactly (float parameters, double
// transcendental calls, float narrowing on return); see the comments per function.
// GSL's gsl_cdf_binomial_Q (third party, not vendored by the reference) is replaced
// by its definition: the upper tail of the binomial distribution.
#include <cmath>
#include <cstdint>
#include <vector>
#include <mutex>
#include "common.cuh"

namespace bani {

namespace {

// log-factorial table, grown on demand
std::vector<long double> g_lfact;
std::mutex g_mu;

const long double *lfact_upto(int n)
{
  if ((int)g_lfact.size() <= n) {
    size_t old = g_lfact.size();
    g_lfact.resize(n + 64);
    for (size_t i = old; i < g_lfact.size(); i++) g_lfact[i] = (i == 0) ? 0.0L : g_lfact[i - 1] + logl((long double)i);
  }
  return g_lfact.data();
}

// tails[y] = P[X >= y], X ~ Bin(n, p), for y = 0..n+1   (tails[n+1] = 0)
void binomial_upper_tails(int n, double p, std::vector<long double> &tails)
{
  tails.assign(n + 2, 0.0L);
  if (p <= 0.0) { tails[0] = 1.0L; return; }
  if (p >= 1.0) { for (int y = 0; y <= n; y++) tails[y] = 1.0L; return; }
  const long double *lf = lfact_upto(n + 1);
  long double lp = logl((long double)p), lq = log1pl(-(long double)p);
  long double acc = 0.0L;
  for (int i = n; i >= 0; i--) {
    acc += expl(lf[n] - lf[i] - lf[n - i] + (long double)i * lp + (long double)(n - i) * lq);
    tails[i] = acc > 1.0L ? 1.0L : acc;
  }
}

// gsl_cdf_binomial_Q(k, p, n) = P[X > k]
double binomial_Q(unsigned k, double p, unsigned n)
{
  if (k >= n) return 0.0;
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<long double> t;
  binomial_upper_tails((int)n, p, t);
  return (double)t[k + 1];
}

// map_stats.hpp:44-56.  j is float; (1 + j) is a float sum; log is the double overload.
float j2md(float j, int k)
{
  if (j == 0) return 1.0;
  if (j == 1) return 0.0;
  float mash_dist = (-1.0 / k) * std::log(2.0 * j / (1 + j));
  return mash_dist;
}

// map_stats.hpp:62-66.  k*d is a float product evaluated by the double exp.
float md2j(float d, int k)
{
  float kd = k * d;
  float jaccard = 1.0 / (2.0 * std::exp((double)kd) - 1.0);
  return jaccard;
}

// map_stats.hpp:79-109 (GSL branch), with the tails of Bin(s, p) shared across the loop
float md_lower_bound(float d, int s, int k, float ci, std::vector<long double> &scratch)
{
  float q2 = (1.0 - ci) / 2;
  float pj = md2j(d, k);
  float sj = s * pj;
  int x = (int)std::ceil((double)sj);
  if (x < 1) x = 1;
  if (x <= s) {
    binomial_upper_tails(s, (double)pj, scratch);
    while (x <= s) {
      double cdf_complement = (double)scratch[x];          // P[X > x-1] = P[X >= x]
      if (cdf_complement < q2) { x--; break; }
      x++;
    }
  }
  float jaccard = float(x) / s;
  return j2md(jaccard, k);
}

// map_stats.hpp:118-130
int min_hits(int s, int k, float perc_identity)
{
  float mash_dist = 1.0 - perc_identity / 100.0;
  float jaccard = md2j(mash_dist, k);
  return (int)std::ceil(1.0 * s * jaccard);
}

} // namespace

// map_stats.hpp:142-167
int stat_min_hits_relaxed(int s, int k, float perc_identity)
{
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<long double> scratch;
  int first = min_hits(s, k, perc_identity);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, k);
    float d_lower = md_lower_bound(d, s, k, 0.9, scratch);
    float id_upper = 100.0 * (1.0 - d_lower);
    if (id_upper >= perc_identity) relaxed = i; else break;
  }
  return relaxed;
}

// Map::doL2Mapping, computeMap.hpp:375-381 (all-float products there)
static void identity_nolock(int shared, int s, int k, float *id, float *ub, std::vector<long double> &scratch)
{
  float mash_dist = j2md(1.0 * shared / s, k);
  float lb = md_lower_bound(mash_dist, s, k, 0.9, scratch);
  *id = 100 * (1 - mash_dist);
  *ub = 100 * (1 - lb);
}

void stat_identity(int shared, int s, int k, float *id, float *ub)
{
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<long double> scratch;
  identity_nolock(shared, s, k, id, ub, scratch);
}

// map_stats.hpp:179-256.  When no sketch size satisfies the p-value cutoff the
// reference reads an uninitialised variable (:237,:252); that case yields
// w = fragLen here (no window fits a fragment => no mappings), which is also
// what the reference produced when probed.
int stat_recommended_window_size(double p_value, int k, float identity, int fragLen, uint64_t refSize)
{
  auto pvalue = [&](int s) -> double {
    double kmerSpace = std::pow(4, k);
    double pX, pY;
    pX = pY = 1. / (1. + kmerSpace / fragLen);
    double r = pX * pY / (pX + pY - pX * pY);
    int x = stat_min_hits_relaxed(s, k, identity);
    double cdf_complement = (x == 0) ? 1.0 : binomial_Q((unsigned)(x - 1), r, (unsigned)s);
    return refSize * cdf_complement;
  };
  int best = -1;
  const int head[3] = {1, 2, 5};
  for (int c = 0; c < 3 && best < 0; c++) if (pvalue(head[c]) <= p_value) best = head[c];
  for (int e = 10; e < fragLen && best < 0; e += 10) if (pvalue(e) <= p_value) best = e;
  if (best < 0) return fragLen;
  int w = 2.0 * fragLen / best;
  return std::min(std::max(w, 1), fragLen);
}

static void lut_row(StatLut &L, int s, std::vector<long double> &scratch, int32_t &minHitsOut, std::vector<float> &idRow, std::vector<float> &ubRow)
{
  // estimateMinimumHitsRelaxed + the max(1, .) of computeL1CandidateRegions (computeMap.hpp:316-317)
  int first = min_hits(s, L.k, L.pid);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, L.k);
    float d_lower = md_lower_bound(d, s, L.k, 0.9, scratch);
    float id_upper = 100.0 * (1.0 - d_lower);
    if (id_upper >= L.pid) relaxed = i; else break;
  }
  minHitsOut = relaxed < 1 ? 1 : relaxed;
  idRow.resize(s + 1); ubRow.resize(s + 1);
  for (int x = 0; x <= s; x++) identity_nolock(x, s, L.k, &idRow[x], &ubRow[x], scratch);
}

void StatLut::ensure(int s_needed)
{
  if (s_needed <= smax) return;
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<long double> scratch;
  if (minHits.empty()) { minHits.push_back(1); rowOff.push_back(0); ident.push_back(0.f); upper.push_back(0.f); have.push_back(1); }  // s = 0 row (unused)
  if ((int)minHits.size() < s_needed + 1) { minHits.resize(s_needed + 1, 1); rowOff.resize(s_needed + 1, 0); have.resize(s_needed + 1, 0); }
  std::vector<float> idRow, ubRow;
  for (int s = smax + 1; s <= s_needed; s++) {
    if (have[s]) continue;
    int32_t mh;
    lut_row(*this, s, scratch, mh, idRow, ubRow);
    minHits[s] = mh;
    rowOff[s] = (uint32_t)ident.size();
    ident.insert(ident.end(), idRow.begin(), idRow.end()); upper.insert(upper.end(), ubRow.begin(), ubRow.end());
    have[s] = 1;
  }
  smax = s_needed;
}

// Rows on demand: a row costs O(s^2) tail terms, so for very large sketches (tiny windows forced through the ABI) only
// the sketch sizes that occur are computed instead of every s up to the maximum.
bool StatLut::ensure_rows(const std::vector<int> &svals)
{
  std::lock_guard<std::mutex> lk(g_mu);
  std::vector<long double> scratch;
  if (minHits.empty()) { minHits.push_back(1); rowOff.push_back(0); ident.push_back(0.f); upper.push_back(0.f); have.push_back(1); }
  bool added = false;
  std::vector<float> idRow, ubRow;
  for (int s : svals) {
    if (s < 1) continue;
    if ((int)minHits.size() < s + 1) { minHits.resize(s + 1, 1); rowOff.resize(s + 1, 0); have.resize(s + 1, 0); }
    if (have[s]) continue;
    int32_t mh;
    lut_row(*this, s, scratch, mh, idRow, ubRow);
    minHits[s] = mh;
    rowOff[s] = (uint32_t)ident.size();
    ident.insert(ident.end(), idRow.begin(), idRow.end()); upper.insert(upper.end(), ubRow.begin(), ubRow.end());
    have[s] = 1;
    added = true;
  }
  return added;
}

} // namespace bani

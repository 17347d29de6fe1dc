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
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <system_error>
#include <thread>
#include <tuple>
#include <vector>
#include "common.cuh"

namespace bani {

namespace {

// log-factorial table: filled once, read-only afterwards (the LUT rows are computed by several threads)
constexpr int LFACT_N = 65600;                 // covers sketch sizes up to fragLen = 60000 (the largest the ABI accepts)
std::mutex g_mu;

const long double *lfact_table()
{
  static std::once_flag once;
  static std::vector<long double> tab;
  std::call_once(once, [] {
    tab.resize(LFACT_N);
    tab[0] = 0.0L;
    for (int i = 1; i < LFACT_N; i++) tab[i] = tab[i - 1] + logl((long double)i);
  });
  return tab.data();
}

// tails[y] = P[X >= y], X ~ Bin(n, p), for y = lowest..n+1   (tails[n+1] = 0; entries below `lowest` are not computed:
// the sum runs downwards from n, so the entries that are computed do not depend on where it stops)
void binomial_upper_tails(int n, double p, std::vector<long double> &tails, int lowest = 0)
{
  tails.assign(n + 2, 0.0L);
  if (lowest < 0) lowest = 0;
  if (p <= 0.0) { tails[0] = 1.0L; return; }
  if (p >= 1.0) { for (int y = 0; y <= n; y++) tails[y] = 1.0L; return; }
  std::vector<long double> own;
  const long double *lf = lfact_table();
  if (n + 1 >= LFACT_N) {                      // beyond the shared table: a private one (slow, correct)
    own.resize((size_t)n + 2); own[0] = 0.0L;
    for (int i = 1; i <= n + 1; i++) own[i] = own[i - 1] + logl((long double)i);
    lf = own.data();
  }
  long double lp = logl((long double)p), lq = log1pl(-(long double)p);
  long double acc = 0.0L;
  for (int i = n; i >= lowest; i--) {
    acc += expl(lf[n] - lf[i] - lf[n - i] + (long double)i * lp + (long double)(n - i) * lq);
    tails[i] = acc > 1.0L ? 1.0L : acc;
  }
}

// gsl_cdf_binomial_Q(k, p, n) = P[X > k]
double binomial_Q(unsigned k, double p, unsigned n)
{
  if (k >= n) return 0.0;
  std::vector<long double> t;
  binomial_upper_tails((int)n, p, t, (int)k + 1);
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
    binomial_upper_tails(s, (double)pj, scratch, x);       // the loop below reads entries x, x+1, ... only
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

// One row of the table: pure function of (s, k, pid)
struct LutRow { int32_t minHits = 1; std::vector<float> ident, upper; };

static void lut_row(int k, float pid, int s, std::vector<long double> &scratch, LutRow &row)
{
  // estimateMinimumHitsRelaxed + the max(1, .) of computeL1CandidateRegions (computeMap.hpp:316-317)
  int first = min_hits(s, k, pid);
  int relaxed = first;
  for (int i = first; i >= 0; i--) {
    float jaccard = 1.0 * i / s;
    float d = j2md(jaccard, k);
    float d_lower = md_lower_bound(d, s, k, 0.9, scratch);
    float id_upper = 100.0 * (1.0 - d_lower);
    if (id_upper >= pid) relaxed = i; else break;
  }
  row.minHits = relaxed < 1 ? 1 : relaxed;
  row.ident.resize(s + 1); row.upper.resize(s + 1);
  for (int x = 0; x <= s; x++) identity_nolock(x, s, k, &row.ident[x], &row.upper[x], scratch);
}

// Rows are shared by every context of the process (the command line drives one context per GPU; a row costs O(s^2) tail
// terms: ~0.6 s for the 320 rows of the default parameters on one core).  Missing rows are computed by a few threads.
namespace {
struct RowKey {
  int k; uint32_t pidBits; int s;
  bool operator<(const RowKey &o) const { return std::tie(k, pidBits, s) < std::tie(o.k, o.pidBits, o.s); }
};
std::map<RowKey, std::shared_ptr<const LutRow>> g_rows;      // guarded by g_mu

uint32_t float_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

// rows for the sketch sizes `want` (each >= 1), in that order; g_mu held by the caller
std::vector<std::shared_ptr<const LutRow>> rows_for(int k, float pid, const std::vector<int> &want)
{
  const uint32_t pb = float_bits(pid);
  std::vector<std::shared_ptr<const LutRow>> out(want.size());
  std::vector<size_t> missing;
  for (size_t i = 0; i < want.size(); i++) {
    auto it = g_rows.find(RowKey{k, pb, want[i]});
    if (it != g_rows.end()) out[i] = it->second; else missing.push_back(i);
  }
  if (!missing.empty()) {
    (void)lfact_table();
    std::vector<std::shared_ptr<LutRow>> made(missing.size());
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(std::min<size_t>(hw ? hw : 1, 16), missing.size());
    // a row of size s costs ~s^2: largest first, each thread takes the next row that nobody has taken
    std::vector<size_t> order(missing.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return want[missing[a]] > want[missing[b]]; });
    std::atomic<size_t> next(0);
    auto work = [&]() {
      std::vector<long double> scratch;
      for (size_t j; (j = next++) < order.size();) {
        auto r = std::make_shared<LutRow>();
        lut_row(k, pid, want[missing[order[j]]], scratch, *r);
        made[order[j]] = std::move(r);
      }
    };
    std::vector<std::thread> th;
    try { for (size_t t = 1; t < nt; t++) th.emplace_back(work); }
    catch (const std::system_error &) {}                          // no more threads to be had: the ones that exist do the work
    work();
    for (auto &x : th) x.join();
    for (size_t j = 0; j < missing.size(); j++) {
      out[missing[j]] = made[j];
      g_rows[RowKey{k, pb, want[missing[j]]}] = made[j];
    }
  }
  return out;
}
} // namespace

void StatLut::ensure(int s_needed)
{
  if (s_needed <= smax) return;
  std::lock_guard<std::mutex> lk(g_mu);
  if (minHits.empty()) { minHits.push_back(1); rowOff.push_back(0); ident.push_back(0.f); upper.push_back(0.f); have.push_back(1); }  // s = 0 row (unused)
  if ((int)minHits.size() < s_needed + 1) { minHits.resize(s_needed + 1, 1); rowOff.resize(s_needed + 1, 0); have.resize(s_needed + 1, 0); }
  std::vector<int> want;
  for (int s = smax + 1; s <= s_needed; s++) if (!have[s]) want.push_back(s);
  const auto rows = rows_for(k, pid, want);
  for (size_t i = 0; i < want.size(); i++) {
    const int s = want[i];
    minHits[s] = rows[i]->minHits;
    rowOff[s] = (uint32_t)ident.size();
    ident.insert(ident.end(), rows[i]->ident.begin(), rows[i]->ident.end()); upper.insert(upper.end(), rows[i]->upper.begin(), rows[i]->upper.end());
    have[s] = 1;
  }
  smax = s_needed;
}

// Rows on demand: a row costs O(s^2) tail terms, so for very large sketches (tiny windows forced through the ABI) only
// the sketch sizes that occur are computed instead of every s up to the maximum.
bool StatLut::ensure_rows(const std::vector<int> &svals)
{
  std::lock_guard<std::mutex> lk(g_mu);
  if (minHits.empty()) { minHits.push_back(1); rowOff.push_back(0); ident.push_back(0.f); upper.push_back(0.f); have.push_back(1); }
  std::vector<int> want;
  for (int s : svals) {
    if (s < 1) continue;
    if ((int)minHits.size() < s + 1) { minHits.resize(s + 1, 1); rowOff.resize(s + 1, 0); have.resize(s + 1, 0); }
    if (have[s] || std::find(want.begin(), want.end(), s) != want.end()) continue;
    want.push_back(s);
  }
  const auto rows = rows_for(k, pid, want);
  for (size_t i = 0; i < want.size(); i++) {
    const int s = want[i];
    minHits[s] = rows[i]->minHits;
    rowOff[s] = (uint32_t)ident.size();
    ident.insert(ident.end(), rows[i]->ident.begin(), rows[i]->ident.end()); upper.insert(upper.end(), rows[i]->upper.begin(), rows[i]->upper.end());
    have[s] = 1;
  }
  return !want.empty();
}

} // namespace bani

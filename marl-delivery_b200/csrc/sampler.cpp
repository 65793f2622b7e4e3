// Host-side reset sampler: the reference's draw order (env.py:96-123,150-170) on numpy's legacy
// RandomState stream, restated natively so that one reset wave of 65k-262k environments takes
// milliseconds instead of minutes.
//
// numpy.random.RandomState(seed) with an int seed == MT19937 seeded by init_genrand(seed); legacy
// randint(lo, hi) == lo + masked-rejection draw in [0, hi-1-lo], one 32-bit output per try, no draw at
// all when the range is a single value.  Pinned against numpy itself in tests/test_sampler.py and
// transitively by the golden reset blocks.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "mdx_internal.h"

namespace mdx {

void Mt19937::seed(uint32_t s) {
  for (int i = 0; i < 624; ++i) {
    mt[i] = s;
    s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
  }
  pos = 624;
}

void Mt19937::refill() {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MAT = 0x9908b0dfu;
  int k = 0;
  for (; k < 624 - 397; ++k) {
    uint32_t y = (mt[k] & UPPER) | (mt[k + 1] & LOWER);
    mt[k] = mt[k + 397] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
  }
  for (; k < 623; ++k) {
    uint32_t y = (mt[k] & UPPER) | (mt[k + 1] & LOWER);
    mt[k] = mt[k + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
  }
  uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
  mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
  pos = 0;
}

// legacy RandomState.randint(lo, hi): value in [lo, hi-1]; returns false when lo >= hi (numpy raises)
bool Mt19937::randint(int64_t lo, int64_t hi, int64_t* out) {
  if (lo >= hi) return false;
  uint64_t span = (uint64_t)(hi - 1 - lo);
  if (span == 0) {  // single value: numpy does not touch the generator
    *out = lo;
    return true;
  }
  if (span > 0xFFFFFFFFull) return false;  // never needed on this path
  uint32_t mask = (uint32_t)span;
  mask |= mask >> 1;
  mask |= mask >> 2;
  mask |= mask >> 4;
  mask |= mask >> 8;
  mask |= mask >> 16;
  uint32_t v;
  do {
    v = next32() & mask;
  } while (v > (uint32_t)span);
  *out = lo + (int64_t)v;
  return true;
}

bool EpisodeSampler::init(uint32_t seed, const uint8_t* cells, int rows_, int cols_, int R_, int P_, int T_) {
  rows = rows_;
  cols = cols_;
  R = R_;
  P = P_;
  T = T_;
  free_cells.clear();
  for (int i = 0; i < rows * cols; ++i)
    if (cells[i] == 0) free_cells.push_back(i);  // row-major, env.py:155,166
  rng.seed(seed);
  return true;
}

// One reset block.  robot_cell[R]; pkg4 = [start[P], target[P], start_time[P], deadline[P]] in
// package-id order.  Returns false where the reference would raise (no free cell left, T too small)
// or loop forever (a single free cell).
bool EpisodeSampler::next(int32_t* robot_cell, int32_t* pkg4) {
  return draw(rng, free_cells, rows, R, P, T, robot_cell, pkg4);
}

// Same, on a caller-owned generator and free list (the env library keeps one generator per env and
// one free list per map; `this` only lends its scratch vectors).
bool EpisodeSampler::draw(Mt19937& rng, const std::vector<int>& free_cells, int rows, int R, int P, int T,
                          int32_t* robot_cell, int32_t* pkg4) {
  int64_t v;
  scratch_avail = free_cells;
  for (int i = 0; i < R; ++i) {  // env.py:97-100 with the shrinking list of :161-170
    if (!rng.randint(0, (int64_t)scratch_avail.size(), &v)) return false;
    robot_cell[i] = scratch_avail[(size_t)v];
    scratch_avail.erase(scratch_avail.begin() + v);
  }
  const int64_t F = (int64_t)free_cells.size();
  if (P > 0 && F < 2) return false;
  const int lim = std::min(R, 20);
  scratch_rec.resize((size_t)P);
  for (int i = 0; i < P; ++i) {
    Rec& rec = scratch_rec[(size_t)i];
    if (!rng.randint(0, F, &v)) return false;
    rec.start = free_cells[(size_t)v];
    do {  // env.py:107-110
      rng.randint(0, F, &v);
      rec.target = free_cells[(size_t)v];
    } while (rec.target == rec.start);
    // env.py:112: 10 + randint(N/2, 3*N) with N = n_rows; numpy truncates the float bound
    if (!rng.randint(rows / 2, 3 * (int64_t)rows, &v)) return false;
    int slack = 10 + (int)v;
    int st = 0;
    if (i > lim) {  // env.py:113-116
      if (!rng.randint(1, T, &v)) return false;
      st = (int)v;
    }
    rec.start_time = st;
    rec.deadline = st + slack;
  }
  std::stable_sort(scratch_rec.begin(), scratch_rec.end(),
                   [](const Rec& a, const Rec& b) { return a.start_time < b.start_time; });  // env.py:119
  for (int j = 0; j < P; ++j) {
    pkg4[0 * P + j] = scratch_rec[(size_t)j].start;
    pkg4[1 * P + j] = scratch_rec[(size_t)j].target;
    pkg4[2 * P + j] = scratch_rec[(size_t)j].start_time;
    pkg4[3 * P + j] = scratch_rec[(size_t)j].deadline;
  }
  return true;
}

// ---- lookup tables (host-built with the same IEEE double arithmetic Python uses) ---------------

void build_quotient_lut(int n, int n_out, float* out) {
  for (int d = -(n_out - 1); d <= n_out - 1; ++d) out[d + n_out - 1] = (float)((double)d / (double)n);
}

// dense rank of the float64 key (dr/rows)**2 + (dc/cols)**2; this file is compiled with
// -ffp-contract=off so the sum of the two rounded squares is not fused.
int build_rank_lut(int rows, int cols, int rows_out, int cols_out, uint16_t* out) {
  const int nr = 2 * rows_out - 1, nc = 2 * cols_out - 1;
  std::vector<double> key((size_t)nr * nc);
  for (int dr = -(rows_out - 1); dr <= rows_out - 1; ++dr)
    for (int dc = -(cols_out - 1); dc <= cols_out - 1; ++dc) {
      volatile double x0 = (double)dr / (double)rows, x1 = (double)dc / (double)cols;
      volatile double a = x0 * x0, b = x1 * x1;
      key[(size_t)(dr + rows_out - 1) * nc + (dc + cols_out - 1)] = a + b;
    }
  std::vector<double> uniq(key);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  for (size_t i = 0; i < key.size(); ++i)
    out[i] = (uint16_t)(std::lower_bound(uniq.begin(), uniq.end(), key[i]) - uniq.begin());
  return (int)uniq.size();
}

// numpy's pairwise float32 sum (add.reduce on a contiguous array): plain loop below 8 elements, 8
// running partials up to 128, recursive halves above.
float np_sum_f32(const float* a, int n) {
  if (n < 8) {
    volatile float s = 0.f;
    for (int i = 0; i < n; ++i) s = s + a[i];
    return s;
  }
  if (n <= 128) {
    volatile float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    volatile float s01 = r[0] + r[1], s23 = r[2] + r[3], s45 = r[4] + r[5], s67 = r[6] + r[7];
    volatile float lo = s01 + s23, hi = s45 + s67;
    volatile float res = lo + hi;
    for (; i < n; ++i) res = res + a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  volatile float l = np_sum_f32(a, n2), h = np_sum_f32(a + n2, n - n2);
  return l + h;
}

}  // namespace mdx

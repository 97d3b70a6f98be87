// Host side of the start-point draw: numpy.random.RandomState.uniform, the same stream, faster.
//
// The reference draws the candidate points of every argmax on the host, from the caller's RandomState:
// `random_state.uniform(low=low, high=high, size=(num_samples, dims))` (bore/mixins.py:49).  The stream must stay
// the caller's (same numbers, same state afterwards), so the draw cannot move to the device; but at BASELINE.json's
// configs[2] it is 65,536 x 50 doubles per BO iteration, and numpy's broadcasting path for array-valued bounds
// spends ~30 ns per number on it -- more than the training kernel takes for the whole fit, so the draw, not the
// GPU, set the end-to-end time of an iteration.  This file restates the PUBLISHED generator numpy's legacy
// RandomState wraps (numpy is a third-party dependency of the reference, absent from /root/reference): MT19937
// (Matsumoto & Nishimura 1998: state of 624 words, `pos` = next word; regeneration when pos == 624), doubles by
// genrand_res53 ((a >> 5) * 2^26 + (b >> 6)) / 2^53 from two consecutive words, uniform = low + (high - low) * u
// with the product and the sum rounded separately (numpy/random/src/distributions: random_uniform).  Whole blocks
// are regenerated and tempered in loops the compiler vectorises.  Bit-identical to numpy, state included:
// tests/test_hostrng.py.  No GPU involved -- compiled by nvcc's host compiler as part of libbore_b200.so.
#include <stdint.h>
#include <string.h>

#include "common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;

// next block of 624 words (the reference implementation's three loops)
void mt_regenerate(uint32_t *mt) {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, MATRIX_A = 0x9908b0dfu;
  int kk = 0;
  for (; kk < MT_N - MT_M; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  for (; kk < MT_N - 1; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
  }
  const uint32_t y = (mt[MT_N - 1] & UPPER) | (mt[0] & LOWER);
  mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? MATRIX_A : 0u);
}

inline uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

}  // namespace

extern "C" int bore_mt19937_uniform(uint32_t *key, int *pos, const double *low, const double *high, int dim,
                                    long long n, double *out) {
  BORE_CHECK(key && pos && low && high && out, "bore_mt19937_uniform: NULL argument");
  BORE_CHECK(dim >= 1 && n >= 0 && *pos >= 0 && *pos <= MT_N, "bore_mt19937_uniform: dim=%d n=%lld pos=%d", dim, n,
             *pos);
  std::vector<double> range((size_t)dim);
  for (int j = 0; j < dim; ++j) range[j] = high[j] - low[j];
  const long long total = n * dim;
  uint32_t words[MT_N + 1];  // tempered words of the current block (+ one carried over from the last block)
  int p = *pos;
  long long i = 0;
  int j = 0;
  int carried = 0;  // 1: words[0] holds the first word of a pair whose second word is in the next block
  while (i < total) {
    if (p >= MT_N) { mt_regenerate(key); p = 0; }
    // temper what this block still holds, but no more than the numbers left need
    const long long need = 2 * (total - i) - carried;
    int take = MT_N - p;
    if ((long long)take > need) take = (int)need;
    for (int k = 0; k < take; ++k) words[carried + k] = mt_temper(key[p + k]);
    p += take;
    const int have = carried + take, pairs = have >> 1;
    for (int q = 0; q < pairs; ++q) {
      const uint32_t a = words[2 * q] >> 5, b = words[2 * q + 1] >> 6;
      const double u = ((double)a * 67108864.0 + (double)b) / 9007199254740992.0;
      const double t = range[j] * u;  // (two roundings, as numpy: no fused multiply-add)
      out[i++] = low[j] + t;
      if (++j == dim) j = 0;
    }
    carried = have & 1;
    if (carried) words[0] = words[have - 1];
  }
  // (total numbers always consume an even count of words, so nothing is carried out of the loop)
  *pos = p;
  return 0;
}

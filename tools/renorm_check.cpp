// renorm_check.cpp — w3_coder_renorm (one clz of the width) and w3_ac_renorm (closed form over
// the bounds) against the literal loops of arithmetic_coder.rs:51-62, on ~6e8 states: random
// bounds at every width scale, bounds hugging the midpoint and the quarter points (E3 territory),
// and every small width around every power of two.
// Build/run (CPU only): g++ -O2 -std=c++17 -o /tmp/renorm_check tools/renorm_check.cpp && /tmp/renorm_check
#include <cstdint>
#include <cstdio>
#include <random>
#include "../weath3rb0i_b200/csrc/w3_core.cuh"

static int literal(uint32_t &x1, uint32_t &x2, int &s, int &k) {
  s = k = 0;
  while (((x1 ^ x2) >> 31) == 0) { x1 <<= 1; x2 = (x2 << 1) | 1; s++; }
  while (x1 >= 0x40000000u && x2 < 0xC0000000u) { x1 = (x1 << 1) & 0x7FFFFFFFu; x2 = (x2 << 1) | 0x80000001u; k++; }
  return s + k;
}

int main(int argc, char **argv) {
  const uint64_t n_random = argc > 1 ? strtoull(argv[1], 0, 10) : 300000000ull;
  std::mt19937_64 rng(12345);
  uint64_t bad = 0, cnt = 0;
  auto test = [&](uint32_t x1, uint32_t x2) {
    if (x1 >= x2) return; /* w >= 1: always true in the codec */
    uint32_t r1 = x1, r2 = x2; int s, k;
    const int nr = literal(r1, r2, s, k);
    w3_coder c; uint32_t flip;
    const int nn = w3_coder_renorm(c, x1, x2 - x1, flip);
    uint32_t o1 = x1, o2 = x2; int os, ok;
    const int no = w3_ac_renorm(o1, o2, os, ok);
    cnt++;
    if (nr != nn || r1 != c.x1 || r2 != c.x1 + c.w || (flip != 0) != (k > 0) || flip != (k > 0 ? 0x80000000u : 0u) ||
        no != nr || o1 != r1 || o2 != r2 || os != s || ok != k) {
      if (bad < 10) printf("MISMATCH x1=%08x x2=%08x literal n=%d (s=%d k=%d) %08x %08x | coder n=%d %08x %08x flip=%x | closed n=%d\n",
                           x1, x2, nr, s, k, r1, r2, nn, c.x1, c.x1 + c.w, flip, no);
      bad++;
    }
  };
  static const uint32_t centres[4] = {0x80000000u, 0x40000000u, 0xC0000000u, 0x7FFFFFFFu};
  for (uint64_t i = 0; i < n_random; i++) {
    const uint64_t r = rng();
    const uint32_t x1 = (uint32_t)r;
    uint32_t w = (uint32_t)(r >> 32) >> (rng() & 31);
    if (!w) w = 1;
    uint32_t x2 = x1 + w;
    if (x2 < x1) x2 = 0xFFFFFFFFu;
    test(x1, x2);
    const uint32_t d1 = (uint32_t)(rng() >> (32 + (rng() & 31))), d2 = (uint32_t)(rng() >> (32 + (rng() & 31)));
    test(centres[i & 3] - d1, centres[i & 3] + d2);
  }
  for (int p = 1; p < 32; p++)
    for (int64_t d = -70; d <= 70; d++)
      for (uint32_t w = 1; w < 140; w++) {
        const uint32_t x1 = (uint32_t)((1ull << p) + d);
        test(x1, x1 + w);
      }
  test(0u, 0xFFFFFFFFu);
  printf("%llu states, %llu mismatches\n", (unsigned long long)cnt, (unsigned long long)bad);
  return bad != 0;
}

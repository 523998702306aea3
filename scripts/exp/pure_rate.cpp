// Experiment: how often does a lookup key fall into a 13-bit cell that holds no threshold ("pure": rank = base)?
// build: g++ -O2 -std=c++17 -o /tmp/pure_rate scripts/exp/pure_rate.cpp ; run: /tmp/pure_rate file
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../weath3rb0i_b200/csrc/w3_core.cuh"
#include "../../weath3rb0i_b200/csrc/w3_lane.cuh"
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  std::vector<uint8_t> d(4u << 20);
  size_t n = fread(d.data(), 1, d.size(), f);
  fclose(f);
  const uint16_t book1[8] = W3_BOOK1_TABLE;
  static w3_hashtab tab;
  memset(&tab, 0, sizeof tab);
  for (int i = 0; i < 8; i++) tab.table[i] = book1[i];
  tab.max_bits = 8;
  for (uint32_t a = 0; a < 8; a++) for (uint32_t j = 0; j < 256; j++) w3_build_thr(&tab, a, j);
  for (uint32_t a = 0; a < 8; a++) {
    for (uint32_t c = 0; c < W3_NCELL; c++) w3_build_cell(&tab, a, c);
    for (uint32_t c = 0; c < W3_NCELL2; c++) w3_build_cell2(&tab, a, c);
  }
  for (int pb = 10; pb <= 16; pb++) {
    uint64_t look = 0, pure_true = 0, pure_any4 = 0;
    uint64_t hist = 0;
    for (size_t t = 0; t + 3 < n * 8; t++) {
      const uint32_t bit = (d[t >> 3] >> (7 - (t & 7))) & 1u;
      hist = (hist << 1) | bit;
      // key for position t+1 (true path)
      const uint64_t key = w3_key_of_history(hist);
      const uint32_t a = (uint32_t)((t + 1) & 7u);
      auto pure = [&](uint64_t k) {
        const uint64_t lo = k >> (64 - pb) << (64 - pb), hi = lo | ((1ull << (64 - pb)) - 1);
        // thresholds ascend: pure iff no thr in (lo, hi]  (rank = #thr < key ... treat thr in [lo,hi] as impure)
        int l = 0, r = 255;
        while (l < r) { int m = (l + r) / 2; if (tab.thr[a][m] < lo) l = m + 1; else r = m; }
        return !(l < 255 && tab.thr[a][l] <= hi);
      };
      look++;
      pure_true += pure(key);
      // four guesses for position t+3 alignment differs, but use same alignment table for a rough figure
      int all = 1;
      for (uint32_t g = 0; g < 4; g++) all &= pure((key >> 2) | ((uint64_t)g << 62));
      pure_any4 += all;
    }
    printf("cell bits %2d: pure on the true path %.2f%%, all four guesses pure %.2f%%\n", pb, 100.0 * pure_true / look, 100.0 * pure_any4 / look);
  }
  return 0;
}

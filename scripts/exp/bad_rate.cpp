// Experiment: how often does the fixed lookup of the pair decoder decline (per lane and per warp step)?
// build: g++ -O2 -std=c++17 -o /tmp/bad_rate scripts/exp/bad_rate.cpp ; run: /tmp/bad_rate file
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../weath3rb0i_b200/csrc/w3_core.cuh"
#include "../../weath3rb0i_b200/csrc/w3_lane.cuh"
static void build(const uint16_t *table, int max_bits, w3_hashtab *tab) {
  memset(tab, 0, sizeof *tab);
  for (int i = 0; i < 8; i++) tab->table[i] = table[i];
  tab->max_bits = (uint16_t)max_bits;
  for (uint32_t a = 0; a < 8; a++) for (uint32_t j = 0; j < 256; j++) w3_build_thr(tab, a, j);
  for (uint32_t a = 0; a < 8; a++) {
    for (uint32_t c = 0; c < W3_NCELL; c++) w3_build_cell(tab, a, c);
    for (uint32_t c = 0; c < W3_NCELL2; c++) w3_build_cell2(tab, a, c);
    for (uint32_t r = 0; r < 256; r++) w3_build_rval(tab, a, r);
  }
}
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  std::vector<uint8_t> d(16u << 20);
  size_t n = fread(d.data(), 1, d.size(), f);
  fclose(f);
  const uint16_t book1[8] = W3_BOOK1_TABLE;
  static w3_hashtab tab;
  build(book1, 8, &tab);
  static w3_lane_tab lt;
  for (uint32_t i = 0; i < 8u * W3_LANE_ROW; i++) lt.thi31[i] = w3_lane_tab_thi31(&tab, i);
  for (uint32_t i = 0; i < 8u * W3_LANE_NCELL; i++) lt.cellb[i] = w3_lane_tab_cellb(&tab, i);
  w3_lane_mem m; m.thi31 = (w3_sptr)lt.thi31; m.cellb = (w3_sptr)lt.cellb; m.cnt = 0; m.g = &tab; m.ones = ~0u;
  const size_t B = 1u << 20, nb = n / B;  // blocks decoded in lockstep: position t of every block at once
  uint64_t look = 0, bad_true = 0, bad_wrong = 0, steps = 0, steps_bad16 = 0, tie = 0;
  std::vector<uint64_t> hist(nb, 0);
  for (size_t t = 0; t < B * 8; t++) {
    int any = 0; int cnt = 0;
    for (size_t b = 0; b < nb && b < 16; b++) {
      const uint32_t bit = (d[b * B + (t >> 3)] >> (7 - (t & 7))) & 1u;
      // history after t bits = hist[b]; candidates for position t+1 with guess g
      const uint64_t key = w3_key_of_history(hist[b]);
      for (uint32_t g = 0; g < 2; g++) {
        const uint64_t gk = (key >> 1) | ((uint64_t)g << 63);
        bool bad; w3_lane_rank((uint32_t)(gk >> 32), (uint32_t)((t + 1) & 7u), m, bad);
        look++;
        if (bad) { if (g == bit) bad_true++; else bad_wrong++; any = 1; cnt++; }
      }
      hist[b] = (hist[b] << 1) | bit;
    }
    steps++; steps_bad16 += any;
  }
  printf("blocks %zu lookups %llu bad(true path) %.4f%% bad(wrong path) %.4f%% steps with any bad among 16 pairs %.2f%%\n", nb < 16 ? nb : 16,
         (unsigned long long)look, 100.0 * bad_true / (look / 2), 100.0 * bad_wrong / (look / 2), 100.0 * steps_bad16 / steps);
  return 0;
}

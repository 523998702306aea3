// Experiment: why does the fixed lookup decline?  tie on the top 31 bits vs. none of the W3_LANE_FIX thresholds above the key
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../weath3rb0i_b200/csrc/w3_core.cuh"
#include "../../weath3rb0i_b200/csrc/w3_lane.cuh"
int main(int argc, char **argv) {
  FILE *f = fopen(argv[1], "rb");
  std::vector<uint8_t> d(8u << 20);
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
  static w3_lane_tab lt;
  for (uint32_t i = 0; i < 8u * W3_LANE_ROW; i++) lt.thi31[i] = w3_lane_tab_thi31(&tab, i);
  for (uint32_t i = 0; i < 8u * W3_LANE_NCELL; i++) lt.cellb[i] = w3_lane_tab_cellb(&tab, i);
  uint64_t look = 0, tie = 0, none = 0, none_within8 = 0, hist = 0;
  for (size_t t = 0; t + 3 < n * 8; t++) {
    const uint32_t bit = (d[t >> 3] >> (7 - (t & 7))) & 1u;
    hist = (hist << 1) | bit;
    const uint64_t key = w3_key_of_history(hist);
    for (uint32_t g = 0; g < 4; g++) {
      const uint32_t a = (uint32_t)((t + 3) & 7u);
      const uint32_t khi = (uint32_t)(key >> 34) | (g << 30);
      const uint32_t base = lt.cellb[a * W3_LANE_NCELL + (khi >> (32 - W3_LANE_PB))];
      const uint32_t k31 = khi >> 1;
      uint32_t above = 0; bool ti = false;
      for (int i = 0; i < W3_LANE_FIX; i++) { const uint32_t z = k31 - lt.thi31[a * W3_LANE_ROW + base + i]; above += z >> 31; ti |= z == 0; }
      look++;
      if (ti) tie++;
      else if (above == 0) {
        none++;
        uint32_t ab2 = 0; bool t2 = false;
        for (int i = W3_LANE_FIX; i < W3_LANE_FIX + 4; i++) { const uint32_t z = k31 - lt.thi31[a * W3_LANE_ROW + base + i]; ab2 += z >> 31; t2 |= z == 0; }
        if (ab2 && !t2) none_within8++;
      }
    }
  }
  printf("lookups %llu: tie %.4f%%, none above %.4f%% (resolved by four more: %.4f%%)\n", (unsigned long long)look, 100.0 * tie / look, 100.0 * none / look, 100.0 * none_within8 / look);
  return 0;
}

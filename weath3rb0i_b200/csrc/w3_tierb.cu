/*
 * w3_tierb.cu — device-side build of the state table and the unit self-tests of
 * w3_tierb.cuh (state table, hash-slot cells), called through the C ABI by tests/.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "w3_kernels.h"
#include "w3_tierb.cuh"

__global__ void w3_k_build_state_table(w3_state_entry *table, w3_state_entry *aux) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W3_ST_SIZE) table[i] = w3_state_entry_of(i);
  if (aux && i < W3_ST_SUB) aux[i] = w3_aux_state_entry(i);
}

/* One thread replays a list of cell operations in order (sets and gets interleave).
 * op = {code | cell << 8, a, b, c}; result = two words per op.
 *   0 get_idx   a = id, b = bit_id, c = nib_ctx            -> idx >> 1, idx & 1
 *   1 get_state a = id, b = bit_id, c = nib_ctx            -> state
 *   2 set_state a = id | bit_id << 8 | nib_ctx << 16, b = state
 *   3 get_slot  a = hash low word, b = hash high word      -> slot id
 *   4 set_nib   a = id | nib << 8, b = s0 | s1 << 16, c = s2 | s3 << 16
 *   5 get_nib   a = id | nib << 8                          -> s0 | s1 << 16, s2 | s3 << 16
 *   6 cell_idx  a, b = hash low / high word, c = log_cell_count -> index low / high word */
__global__ void w3_k_hashslot_ops(const uint32_t *ops, uint32_t n_ops, uint8_t *cells, uint32_t *res) {
  if (threadIdx.x || blockIdx.x) return;
  for (uint32_t i = 0; i < n_ops; i++) {
    const uint32_t code = ops[4 * i] & 0xFFu, a = ops[4 * i + 1], b = ops[4 * i + 2], c = ops[4 * i + 3];
    uint8_t *cell = cells + (size_t)(ops[4 * i] >> 8) * W3_CELL_BYTES;
    uint32_t r0 = 0, r1 = 0;
    uint16_t s[4];
    switch (code) {
      case 0: { const uint32_t idx = w3_slot_idx(a, b, c); r0 = idx >> 1; r1 = idx & 1u; break; }
      case 1: r0 = w3_slot_get_state(cell, a, b, c); break;
      case 2: w3_slot_set_state(cell, a & 0xFFu, (a >> 8) & 0xFFu, a >> 16, b); break;
      case 3: r0 = w3_cell_get_slot(cell, ((uint64_t)b << 32) | a); break;
      case 4:
        s[0] = (uint16_t)b; s[1] = (uint16_t)(b >> 16); s[2] = (uint16_t)c; s[3] = (uint16_t)(c >> 16);
        w3_slot_set_nib(cell, a & 0xFFu, a >> 8, s);
        break;
      case 5:
        w3_slot_get_nib(cell, a & 0xFFu, a >> 8, s);
        r0 = s[0] | ((uint32_t)s[1] << 16); r1 = s[2] | ((uint32_t)s[3] << 16);
        break;
      case 6: { const uint64_t ix = w3_cell_index(((uint64_t)b << 32) | a, c); r0 = (uint32_t)ix; r1 = (uint32_t)(ix >> 32); break; }
      default: r0 = r1 = 0xDEADBEEFu;
    }
    res[2 * i] = r0;
    res[2 * i + 1] = r1;
  }
}

extern "C" {
cudaError_t w3_launch_build_state_table(w3_state_entry *d_table, w3_state_entry *d_aux, cudaStream_t s) {
  w3_k_build_state_table<<<(W3_ST_SIZE + 255) / 256, 256, 0, s>>>(d_table, d_aux);
  return cudaGetLastError();
}
cudaError_t w3_launch_hashslot_ops(const uint32_t *d_ops, uint32_t n_ops, uint8_t *d_cells, uint32_t *d_res, cudaStream_t s) {
  w3_k_hashslot_ops<<<1, 32, 0, s>>>(d_ops, n_ops, d_cells, d_res);
  return cudaGetLastError();
}
}

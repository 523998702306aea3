/* w3_kernels.h — internal launchers (w3_kernels.cu) used by the C ABI (w3_capi.cu). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define W3_MAX_WARPS 8 /* blocks (= warps) sharing one CTA's copy of the threshold table */

struct w3_hashtab;
struct w3_state_entry;

/* which kernel family codes the blocks: the context's options (w3_ctx_set_option), 0 = the launcher decides */
struct w3_launch_opts {
  int kernels;    /* W3_OPT_KERNELS: 0 auto, 1 team (one block per CTA), 2 warp (one warp per block), 3 many-block (teams sharing a CTA / lane per block) */
  int enc_teams;  /* W3_OPT_ENC_TEAMS: teams per CTA of the team encoder, 0 auto */
  int lane_warps; /* W3_OPT_LANE_WARPS: warps per CTA of the lane decoder, 0 auto */
  int enc_coder;  /* W3_OPT_ENC_CODER: 0 auto (one coder warp per CTA from two teams on), 1 a coder warp per team */
  int scratch_permille; /* W3_OPT_SCRATCH_PERMILLE: first-pass scratch region per block in 1/1000 of the block size; 0 = 1125 */
  int lane_form;  /* W3_OPT_LANE_FORM: 0 auto, 1 one lane per block, 2 a pair of lanes per block, 3 pair + probability table */
};

#ifdef __cplusplus
extern "C" {
#endif
cudaError_t w3_kernels_init(void);
cudaError_t w3_launch_build_table(struct w3_hashtab *d_tab, cudaStream_t s);
cudaError_t w3_launch_encode(const uint8_t *d_in, uint64_t n_total, uint32_t block_size,
                             uint32_t n_blocks, uint8_t *d_scratch, uint64_t stride,
                             uint32_t *d_lens, uint32_t *d_overflow,
                             const struct w3_hashtab *d_tab, int sm_count,
                             const struct w3_launch_opts *opts, cudaStream_t s);
cudaError_t w3_launch_scan(const uint32_t *d_lens, uint32_t n_blocks, uint64_t *d_offsets,
                           cudaStream_t s);
cudaError_t w3_launch_pack(const uint8_t *d_scratch, uint64_t stride, const uint32_t *d_lens,
                           uint32_t n_blocks, uint64_t *d_offsets, const uint64_t *d_alt,
                           uint8_t *d_packed, cudaStream_t s);
cudaError_t w3_launch_decode(const uint8_t *d_comp, const uint64_t *d_offsets, uint32_t n_blocks,
                             uint32_t block_size, uint64_t n_total, uint8_t *d_out,
                             uint32_t *d_err, const struct w3_hashtab *d_tab, int sm_count,
                             const struct w3_launch_opts *opts, cudaStream_t s);
/* the launcher's choice of decode kernel (w3_kernels.cu: W3_DEC_*), and its name */
int w3_decode_family(uint32_t n_blocks, uint32_t block_size, int sm_count, int kernels, int lane_form);
const char *w3_decode_kernel_name(uint32_t n_blocks, uint32_t block_size, int sm_count, int kernels, int lane_form);
/* StationaryModel::new on the device: d_cnt needs ceil(n / 1024) * 8 u16 */
cudaError_t w3_launch_train(const uint8_t *d_in, uint64_t n, uint16_t *d_cnt, uint16_t *d_table,
                            int sm_count, cudaStream_t s);
cudaError_t w3_launch_hash_array(const uint64_t *d_bits, const uint32_t *d_pos, uint32_t n,
                                 uint32_t *d_out, const struct w3_hashtab *d_tab, cudaStream_t s);
cudaError_t w3_launch_counter_p_check(unsigned long long *d_mismatches, cudaStream_t s);
cudaError_t w3_launch_gather_probe(const void *d_buf, uint64_t n_sectors, uint32_t blocks, uint32_t per_thread,
                                   unsigned long long *d_sink, cudaStream_t s);
cudaError_t w3_launch_ac_encode_probs(const uint8_t *d_in, const uint16_t *d_probs,
                                      uint32_t n_probs, uint8_t *d_out, uint32_t cap,
                                      uint32_t *d_out_len, cudaStream_t s);
cudaError_t w3_launch_ac_decode_probs(const uint8_t *d_in, uint32_t n_in, const uint16_t *d_probs,
                                      uint32_t n_probs, uint8_t *d_out, cudaStream_t s);
/* the hashed context-mixing stack (w3_cm.cu): one warp per block, cells = n_blocks x (96 << log_cells) zeroed bytes */
cudaError_t w3_cm_init(void);
cudaError_t w3_launch_cm_encode(const uint8_t *d_in, uint64_t n_total, uint32_t block_size, uint32_t n_blocks,
                                uint8_t *d_scratch, uint64_t stride, uint32_t *d_lens, uint32_t *d_err, uint8_t *d_cells,
                                uint32_t log_cells, cudaStream_t s);
cudaError_t w3_launch_cm_decode(const uint8_t *d_comp, const uint64_t *d_offsets, uint64_t n_total, uint32_t block_size,
                                uint32_t n_blocks, uint8_t *d_out, uint32_t *d_err, uint8_t *d_cells, uint32_t log_cells,
                                cudaStream_t s);
/* Tier B (w3_tierb.cu): NaiveStateTable built on the device; the hash-slot unit self-test */
cudaError_t w3_launch_build_state_table(struct w3_state_entry *d_table, struct w3_state_entry *d_aux, cudaStream_t s);
cudaError_t w3_launch_hashslot_ops(const uint32_t *d_ops, uint32_t n_ops, uint8_t *d_cells, uint32_t *d_res, cudaStream_t s);
#ifdef __cplusplus
}
#endif

/* w3_geo.h — launchers of the run-time-geometry kernels (w3_geo.cu), used by the C ABI. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

/* one sequential coding job of w3_k_geo_encode / w3_k_geo_decode: a block of a file, or a point of
 * a sweep over one file.  All pointers are device pointers. */
struct w3_geo_job {
  const uint32_t *H;  /* encode: ACHistory::hash of every position of this stream (>= the stream's max_bits) */
  const uint8_t *in;  /* encode: the bytes to code; decode: the stream */
  uint8_t *out;       /* encode: the stream, NULL = count bits only (ACStats); decode: the bytes */
  uint32_t *counters; /* 1 << ctx_bits words (c0 | c1 << 16), zeroed by the caller */
  uint32_t nbytes;    /* uncompressed bytes of this stream */
  uint32_t cap;       /* encode: capacity of out */
  uint32_t ctx_bits, align_bits; /* OrderNEntropy::new(bits_in_context, alignment_bits, ..) */
  uint64_t bit_count; /* out (encode, count only): bits ACStats counted */
  uint32_t out_len;   /* encode: out, the stream's length (full length even if it overflowed cap); decode: in, the stream's length */
  uint32_t flags;     /* out: 1 = the stream did not fit cap */
};

#ifdef __cplusplus
extern "C" {
#endif
cudaError_t w3_launch_geo_hash(const uint8_t *d_in, uint64_t n, uint32_t block_size, const uint16_t *d_table,
                               uint32_t max_bits, uint32_t raw, uint32_t *d_H, int sm_count, cudaStream_t s);
cudaError_t w3_launch_geo_encode(struct w3_geo_job *d_jobs, uint32_t n_jobs, uint32_t lanes, cudaStream_t s);
cudaError_t w3_launch_geo_decode(struct w3_geo_job *d_jobs, uint32_t n_jobs, uint32_t lanes, const uint16_t *d_table,
                                 uint32_t max_bits, cudaStream_t s);
#ifdef __cplusplus
}
#endif

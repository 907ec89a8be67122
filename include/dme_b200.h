/*
 * dme_b200.h -- C ABI of libdme_b200.so: the B200 (sm_100a) implementation of the distributed-mean-
 * estimation hot path of Ritesh622/Unbiased-Quantization-Distributed-Mean-Estimation.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream handle
 * (void* == cudaStream_t), enqueues work on that stream and returns immediately.  No torch types.
 * Return value: 0 = OK, <0 = DME_E* (message via dme_last_error(), thread-local).  The library never
 * allocates user-visible memory: outputs and scratch are caller-owned.  There is NO CPU fallback: a
 * call without a usable CUDA device fails with DME_ECUDA.
 *
 * Reference interface each entry replaces (AS = NMSE_Results/Codes/All_Schemes.py,
 * ND = NMSE_Results/Codes/Normal_dist.py of the reference tree):
 *   dme_l1_norms            AS:624, AS:681          input_vector.abs().sum()
 *   dme_type_quantize       AS:609-641 (mode 0), AS:644-687 (mode 1); returns what the reference returns
 *                           (deq) and/or the integer type vector + signs it only implies
 *   dme_type_encode         same arithmetic, emits the packed code "DMEP1" (no reference counterpart)
 *   dme_decode_mean         AS:640 / AS:687 dequantise + ND:133-147 `est += q / n`
 *   dme_decode_mean_tiles   the same for a range of tiles (slices of the mean, for overlap with the all-reduce)
 *   dme_quantize_mean       fused: the whole server loop ND:133-147 for the type quantizers
 *   dme_hadamard            AS:100-115   Hadamard.hadamard
 *   dme_rht / dme_irht      AS:127-144 / AS:151-156 (diagonal AS:117-120: Philox or injected)
 *   dme_pair_transform      AS:37-59 as it executes
 *   dme_drive               AS:707-752
 *   dme_eden_encode/decode  AS:335-350,370-390 / AS:398-426
 *   dme_quicfl_decode       AS:526-535
 *   dme_scalar_quantize     AS:755-790
 *   dme_mean_accumulate     ND:133-147 for schemes that return a dequantised vector
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Layout conventions: client vectors are rows of a row-major matrix X[n][ld], ld >= d, ld % 4 == 0,
 * X 16-byte aligned (the Python shim copies when that does not hold).  All vectors are fp32.
 */
#ifndef DME_B200_H
#define DME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DME_OK 0
#define DME_EINVAL (-1)    /* bad argument (message says which) */
#define DME_ECUDA (-2)     /* CUDA runtime error / no device */
#define DME_EWORKSPACE (-3)/* workspace or code arena too small */
#define DME_EOVERFLOW (-4) /* a magnitude does not fit the requested output (reported by dme_status) */
#define DME_ERETRY (-5)    /* biased mode: a tie-heavy row overflowed the fast selection (reported by dme_status): rerun the call after
                            * dme_set_biased_path(1) */

#define DME_MODE_UNBIASED 0
#define DME_MODE_BIASED 1

#define DME_TILE 1024      /* coordinates per tile of the packed code (one directory entry, one field width) */

typedef void *dme_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define DME_API __attribute__((visibility("default")))
#else
#define DME_API
#endif

DME_API const char *dme_last_error(void);
DME_API int dme_version(void);
/* Number of kernels this library has launched in the calling process (bench.py's gpu_launches). */
DME_API int64_t dme_launch_count(void);

/* Per-kernel timing (bench.py's roofline leg; keep it off inside timed regions; one caller thread).  While enabled, the
 * library records a CUDA event on `stream` after every kernel it launches; dme_profile_read returns, for the calls made
 * since dme_profile_enable(1, stream), the duration of each kernel in launch order (milliseconds) and dme_profile_name(i)
 * the kernel's name. */
DME_API int dme_profile_enable(int on, dme_stream_t stream);
DME_API int dme_profile_read(float *ms, int cap);
DME_API const char *dme_profile_name(int i);
/* Test hook: which implementation runs the unbiased mode.  0 (default) = l1_kernel + quantize_warp_kernel (the product path);
 * 1 = literal_rows_kernel (AS:625-637 as written, one CTA per row: an independent implementation for the parity tests). */
DME_API int dme_set_unbiased_path(int path);
/* Which selection the biased mode's mass repair (AS:655-664) uses: 0 (default) = one linear histogram of the residuals + a compact
 * candidate list of the threshold bin (4 passes over the rows); 1 = MSB-first radix select over all coordinates (8 passes; any
 * input, however many ties).  Path 0 reports DME_ERETRY through dme_status when a row's threshold bin does not fit its candidate
 * list; both paths give the same result when path 0 succeeds (ties: lowest index first). */
DME_API int dme_set_biased_path(int path);

/* X_c, the single uniform of client c (AS:634): Philox4x32-10, key = seed, counter = (client, 0, 0, 0x584D44),
 * top 24 bits -> [0,1).  Host-side helper so callers/tests can reproduce the draws. */
DME_API float dme_uniform_x(uint64_t seed, uint64_t client);
/* The same draws on the device: xu[c] = X_c of client client0 + c for the seed stored at *seed_dev (DEVICE memory), c < n;
 * bump != 0 adds one to *seed_dev afterwards.  Lets a captured CUDA graph of the path (x_inject = xu) draw fresh uniforms on every
 * replay without a host-side argument (dme_b200.api.MeanGraph). */
DME_API int dme_fill_uniforms(float *xu, int64_t n, uint64_t *seed_dev, uint64_t client0, int bump, dme_stream_t stream);
/* Adds n to dme_launch_count(): a graph replay launches kernels the library does not see. */
DME_API void dme_add_launches(int64_t n);

/* Scratch for the type-quantizer entry points (bytes). */
DME_API int64_t dme_workspace_bytes(int64_t n, int64_t d);
/* Arena size for packed codes: expect=1 -> sized for the field width the rate m/d suggests (+50 %),
 * expect=0 -> worst case (32-bit fields everywhere). */
DME_API int64_t dme_codes_bytes(int64_t n, int64_t d, int64_t m, int expect);
/* Entries (uint64) of the tile directory: n * ceil(d / DME_TILE). */
DME_API int64_t dme_dir_entries(int64_t n, int64_t d);

/* After the stream has been synchronised: 0, or DME_EOVERFLOW / DME_EWORKSPACE raised by a kernel of the
 * last call that used this workspace. */
DME_API int dme_status(const void *ws, dme_stream_t stream);

DME_API int dme_l1_norms(const float *X, int64_t n, int64_t d, int64_t ld, float *l1_out,
                 void *ws, int64_t ws_bytes, dme_stream_t stream);

/* x_inject (n floats, device, nullable): X_c to use instead of Philox(seed, client0 + c).
 * l1_inject (n floats, device, nullable): fp32 L1 norms to use instead of the computed ones.
 * k_out (int32, nullable), sgn_out (uint8, nullable), deq_out (float, nullable): rows of ld_out elements.
 * l1_out (n floats, nullable): the fp32 L1 norms used. */
DME_API int dme_type_quantize(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode,
                      const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                      int32_t *k_out, uint8_t *sgn_out, float *deq_out, int64_t ld_out, float *l1_out,
                      void *ws, int64_t ws_bytes, dme_stream_t stream);

/* Packed code DMEP1.  dir[c * T + t] = (byte offset of the tile in `codes` / 16) << 8 | field width,
 * T = ceil(d / DME_TILE); a tile of width w (the smallest of 2, 4, 8, 16, 32 with max |k| < 2^(w-1)) holds 32*w uint32
 * words: word q of 16-coordinate chunk j (0..63) at [q * 64 + j], field i of the chunk at bits [w*i, w*i + w) =
 * sign << (w-1) | magnitude. */
DME_API int dme_type_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode,
                    const float *x_inject, const float *l1_inject, uint64_t seed, uint64_t client0,
                    void *codes, int64_t codes_bytes, uint64_t *dir, float *l1_out,
                    void *ws, int64_t ws_bytes, dme_stream_t stream);

/* mean[i] (+)= sum_c deq(c, i) / n_total, clients in order, fp32.  accumulate=0 overwrites. */
DME_API int dme_decode_mean(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d,
                    int64_t m, int mode, int64_t n_total, float *mean, int accumulate, dme_stream_t stream);

/* The same for tiles [tile0, tile0 + tiles) only, i.e. coordinates [tile0 * DME_TILE, (tile0 + tiles) * DME_TILE) of the
 * mean: the decoder is tile-major, so a sharded run can all-reduce finished slices of the mean while later slices are
 * still being decoded (ND:133-147 + the exchange step of SURVEY 8e). */
DME_API int dme_decode_mean_tiles(const void *codes, const uint64_t *dir, const float *l1, int64_t n, int64_t d,
                    int64_t m, int mode, int64_t n_total, float *mean, int accumulate, int64_t tile0, int64_t tiles,
                    dme_stream_t stream);

/* quantize -> pack -> decode -> mean in one call (the north-star path).  `codes`/`dir` are scratch here. */
DME_API int dme_quantize_mean(const float *X, int64_t n, int64_t d, int64_t ld, int64_t m, int mode,
                      const float *x_inject, uint64_t seed, uint64_t client0, int64_t n_total,
                      float *mean, int accumulate, void *codes, int64_t codes_bytes, uint64_t *dir,
                      float *l1_out, void *ws, int64_t ws_bytes, dme_stream_t stream);

/* mean[i] (+)= sum_c Q[c][i] / n_total for already-dequantised rows (DRIVE/EDEN/... ND:133-147). */
DME_API int dme_mean_accumulate(const float *Q, int64_t n, int64_t d, int64_t ld, int64_t n_total, float *mean,
                        int accumulate, dme_stream_t stream);

/* The exchange step of a sharded run over NVLink peer memory: every rank's partial mean sits in a symmetric buffer
 * (d floats rounded up to a multiple of 4, the same on every GPU, mapped into every process: bufs = DEVICE array of the `world`
 * mappings of the allocations, the vector starts offset_bytes into each; multicast = the NVSwitch multicast mapping of the same
 * allocations or NULL).  Rank `rank` sums slice `rank` of the vector
 * over all ranks -- inside the switch when `multicast` is given (multimem.ld_reduce / multimem.st), else by peer loads in rank order
 * 0..world-1 -- and writes it into every rank's buffer.  The caller separates it from the writers of the partial means and from the
 * readers of the result by cross-GPU barriers (dme_b200/distributed.py uses the symmetric-memory barrier). */
DME_API int dme_peer_sum_slice(float *const *bufs, float *multicast, int64_t offset_bytes, int rank, int world, int64_t d,
                       dme_stream_t stream);

/* ---- rotations ---- */
/* In-place normalised natural-order Walsh-Hadamard transform of n rows of dpad (power of two) floats. */
DME_API int dme_hadamard(float *V, int64_t n, int64_t dpad, int64_t ld, dme_stream_t stream);
/* out[c][0..dpad) = H(diag_c * pad(x[c])).  diag_c: Philox(seed + c * seed_stride, coordinate) signs (seed_stride 0 =
 * one diagonal shared by all rows, as when one rotation_seed is used; 1 = a fresh rotation per client, as the
 * reference's per-call EDEN seed AS:800), or diag_inject (dpad floats, +-1, shared). */
DME_API int dme_rht(const float *X, int64_t n, int64_t d, int64_t ld, float *out, int64_t dpad, int64_t ld_out,
            uint64_t seed, uint64_t seed_stride, const float *diag_inject, dme_stream_t stream);
DME_API int dme_irht(float *V, int64_t n, int64_t dpad, int64_t ld, uint64_t seed, uint64_t seed_stride,
             const float *diag_inject, dme_stream_t stream);
DME_API int dme_rademacher(float *diag, int64_t dpad, uint64_t seed, dme_stream_t stream);
/* The reference's `fast_walsh_hadamard_transform` as it executes: log2(len) stages on adjacent pairs. */
DME_API int dme_pair_transform(float *V, int64_t n, int64_t len, int64_t ld, dme_stream_t stream);

/* ---- comparison quantizers ---- */
/* dsign_inject: +-1 per padded coordinate, chunk after chunk (2048-chunks, last padded to pow2), nullable.
 * compat 0 = the reference's transform, 1 = true WHT (real DRIVE). */
DME_API int dme_drive(const float *X, int64_t n, int64_t d, int64_t ld, float *out, int64_t ld_out, uint64_t seed,
              const float *dsign_inject, int compat, dme_stream_t stream);
/* rot: n x dpad scratch/out (rotated vectors); bins: uint8 n x dpad; scale: n floats. nbits in {1,2}. */
DME_API int dme_eden_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int nbits, uint64_t seed,
                    uint64_t seed_stride, const float *diag_inject, const float *norm_inject, float *rot, uint8_t *bins, float *scale,
                    dme_stream_t stream);
DME_API int dme_eden_decode(const uint8_t *bins, const float *scale, int64_t n, int64_t d, int64_t dpad, int nbits,
                    uint64_t seed, uint64_t seed_stride, const float *diag_inject, float *work, float *out, int64_t ld_out,
                    dme_stream_t stream);
/* Fractional rates (AS:352-368, AS:385-389, AS:401-421): coordinate i of client c is quantized with the nbits_high table where
 * mask[c][i] is set and with the nbits_low table elsewhere; mask_inject (uint8 n x dpad, nullable) or Bernoulli(p_high) from
 * Philox keyed by the client's mask seed (seed + c * seed_stride) * 7 + 13 -- sender and receiver derive the same mask.
 * drop (uint8 n x dpad, nullable): coordinates the receiver zeroes (rates below 1 bit, AS:413-421); the others are divided
 * by keep = 1 - pdrop. */
DME_API int dme_eden_encode_frac(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int nbits_low, int nbits_high,
                    float p_high, const uint8_t *mask_inject, uint64_t seed, uint64_t seed_stride, const float *diag_inject,
                    const float *norm_inject, float *rot, uint8_t *bins, float *scale, dme_stream_t stream);
DME_API int dme_eden_decode_frac(const uint8_t *bins, const float *scale, int64_t n, int64_t d, int64_t dpad, int nbits_low,
                    int nbits_high, float p_high, const uint8_t *mask_inject, const uint8_t *drop, float keep, uint64_t seed,
                    uint64_t seed_stride, const float *diag_inject, float *work, float *out, int64_t ld_out, dme_stream_t stream);
/* QuicFLSender.compress (AS:455-503) for n rows: shared rotation (rotation_seed or diag_inject), scale = sqrt(dpad) / ||.||_2,
 * coordinates with |z| > exact_threshold are sent exactly (exact_mask, and z itself in exact_dense, 0 elsewhere), the others are
 * rounded stochastically to the grid of step delta ((x_len - 1) / 2 points either side of 0) and looked up in the sender
 * tables send_X (int8) / send_p (fp32), both x_len x h_len: index X and the probability of X + 1 for grid point x and shared
 * randomness h.  The reference's tree does not ship these tables; dme_b200/quicfl_tables.py derives them from the receiver
 * table.  Outputs Xq, h_out (int32), exact_mask (uint8), exact_dense (fp32): all n x dpad; rot: n x dpad scratch; scale_out [n].
 * dpad: power of two >= max(d, 4).  Randomness: Philox keyed by (seed, client0 + row). */
DME_API int dme_quicfl_encode(const float *X, int64_t n, int64_t d, int64_t ld, int64_t dpad, int h_len, int x_len, float delta,
                    float exact_threshold, const int8_t *send_X, const float *send_p, uint64_t seed, uint64_t client0,
                    uint64_t rotation_seed, const float *diag_inject, float *rot, int32_t *Xq, int32_t *h_out,
                    uint8_t *exact_mask, float *exact_dense, float *scale_out, dme_stream_t stream);
/* Xq: int32 n x dpad table rows; h: int32 n x dpad shared randomness; recv_table: (2^nbits) x h_len floats;
 * exact_mask (uint8, nullable) with exact_vals either compacted per row (exact_off[c] .. exact_off[c+1], the reference's
 * format) or, with exact_off == NULL, dense n x dpad (dme_quicfl_encode's exact_dense). */
DME_API int dme_quicfl_decode(const int32_t *Xq, const int32_t *h, int64_t n, int64_t d, int64_t dpad, int h_len,
                      const float *recv_table, int table_len, const uint8_t *exact_mask, const float *exact_vals,
                      const int64_t *exact_off, const float *scale, uint64_t rotation_seed,
                      const float *diag_inject, float *work, float *out, int64_t ld_out, dme_stream_t stream);
/* u_inject: n x d uniforms (nullable -> Philox(seed, client, coordinate)). nlevels = 2^bits - 1. */
DME_API int dme_scalar_quantize(const float *X, int64_t n, int64_t d, int64_t ld, float nlevels, uint64_t seed,
                        uint64_t client0, const float *u_inject, float *out, int64_t ld_out, dme_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DME_B200_H */

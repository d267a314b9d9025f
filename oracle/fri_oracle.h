/*
 * fri_oracle.h — CPU oracle for frave's fractal transform + quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may link or
 * call it, and only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (pagmerek/frave, pure Rust) cannot be built in this
 * environment (no cargo/rustc) and ships no tests, fixtures or golden vectors for this
 * path.  This file is a restatement of the reference *source semantics*; every function
 * cites the reference file:line it follows (paths relative to /root/reference/).  It is
 * cross-checked against an independently written numpy restatement (oracle/fri_oracle_np.py)
 * and the known-answer hashes recorded in SURVEY.md §8(c).
 */
#ifndef FRI_ORACLE_H
#define FRI_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* crates/libfri/src/fractal.rs:51-86 — digit vectors, (re = x, im = y). */
extern const int32_t FRI_ORACLE_LITERALS[30][2];

/* Rust `Option<i32>` */
typedef struct {
    int32_t v;
    uint8_t some;
} fri_opt_i32;

/* crates/libfri/src/images.rs:82-85 (+ a sample_bytes extension: 1 = reference, 2 = u16 LE). */
typedef struct {
    uint32_t width, height;
    uint32_t channels;     /* ColorSpace::num_channels, images.rs:15-21 */
    uint32_t sample_bytes; /* 1 in the reference (Vec<u8>) */
    const void *data;      /* HWC interleaved */
} fri_oracle_raster;

/* utils.rs:5-14 */
size_t fri_oracle_prev_power_two(size_t x);

/* wavelet_transform.rs:71-90 */
void fri_oracle_nearby_vectors(int depth, int32_t out[6][2]);

/* wavelet_transform.rs:42-54 — image_positions, 2^(depth+1) entries of (re, im). */
void fri_oracle_image_positions(int depth, int32_t cx, int32_t cy, int32_t *pos);

/* wavelet_transform.rs:450-484 — returns all built tiles (in-bounds + fringe), malloc'd
 * array of (re, im) pairs in insertion order. */
int fri_oracle_fractal_divide(uint32_t width, uint32_t height, int depth,
                              int32_t **centers_out, size_t *n_out);

/* images.rs:89-100 */
fri_opt_i32 fri_oracle_get_pixel(const fri_oracle_raster *img, int32_t x, int32_t y,
                                 uint32_t channel);

/* wavelet_transform.rs:179-225 — coef: [channels][1<<depth] Option<i32>. */
void fri_oracle_extract_coefficients(const fri_oracle_raster *img, int depth, int32_t cx,
                                     int32_t cy, fri_opt_i32 *coef);

/* wavelet_transform.rs:405-416 — fractal_divide + extract_coefficients + retain, with the
 * retained tiles sorted by (centre.im, centre.re).  Output (malloc'd, caller frees):
 *   centers [n][2] (re, im); coef [n][channels][1<<depth] dense i32 (None stored as 0);
 *   some [n][channels][1<<depth] u8.
 * Retain rule: see the note in fri_oracle.c (the reference's rule drops every tile of a
 * 1-channel image and then panics; here the rule ranges over the active channels). */
int fri_oracle_from_raster(const fri_oracle_raster *img, int depth, int32_t **centers_out,
                           int32_t **coef_out, uint8_t **some_out, size_t *n_out);

/* Same but for a caller-supplied tile list (no BFS, no retain): timing + batch use.
 * nthreads > 1 splits the tile list over pthreads. */
void fri_oracle_extract_tiles(const fri_oracle_raster *img, int depth, const int32_t *centers,
                              size_t n, int32_t *coef, uint8_t *some, int nthreads);

/* quantization.rs:7-25 (encode) and :27-45 (decode): identical truncating division, with
 * the 32-entry matrix as a parameter (reference: all ones, quantization.rs:3-5).
 * multiply != 0 is NOT the reference: it is the "true dequantizer" variant. */
void fri_oracle_quantize(int32_t *coef, const uint8_t *some, size_t n_tiles, uint32_t channels,
                         int depth, const int32_t q[32], int multiply);

/* images.rs:103-111 applied through wavelet_transform.rs:358-381 for every tile;
 * out must be zero-initialised by the caller exactly like from_wavelet does (:309-317). */
void fri_oracle_extract_values(const int32_t *centers, const int32_t *coef, const uint8_t *some,
                               size_t n_tiles, int depth, uint32_t width, uint32_t height,
                               uint32_t channels, uint32_t sample_bytes, void *out, int nthreads);

/* Timed CPU baseline drivers: transform + quantization (encoder.rs:26-33) and dequantization
 * (dividing, quantization.rs:37) + inverse transform (decoder.rs:27-34) over a tile list. */
void fri_oracle_encode_tiles(const fri_oracle_raster *img, int depth, const int32_t *centers, size_t n_tiles,
                             const int32_t q[32], int32_t *coef, int nthreads);
void fri_oracle_decode_tiles(const int32_t *centers, const int32_t *coef, const uint8_t *some, size_t n_tiles,
                             int depth, uint32_t width, uint32_t height, uint32_t channels, uint32_t sample_bytes,
                             const int32_t q[32], void *out, int nthreads);

/* Cost model, not a restatement: the allocations, SipHash-1-3 hashes and
 * HashMap inserts Fractal::new performs per tile (wavelet_transform.rs:42-69).  Returns a checksum. */
uint64_t fri_oracle_fractal_new_cost(int depth, const int32_t *centers, size_t n_tiles, uint32_t channels);

void fri_oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif

/*
 * fri_cuda.h — C ABI of libfri_cuda: frave's fractal transform + quantization hot path on
 * NVIDIA B200 (sm_100a).  This is the boundary a `libfri-cuda` Rust crate binds with
 * `extern "C"` (see rust/libfri-cuda/ and INTEGRATION.md); plain pointers and sizes only.
 *
 * Reference citations are file:line relative to the pagmerek/frave tree
 * (crates/libfri/src/...).  What each entry point replaces:
 *
 *   fri_plan_create        WaveletImage::fractal_divide + Fractal::new + the retain step
 *                          (stages/wavelet_transform.rs:450-484, :42-69, :415-416) and, on
 *                          decode, WaveletImage::from_metadata (:392-403) whose only purpose
 *                          is to rebuild the lattice and the Some/None masks.
 *   fri_encode_tq*         wavelet_transform::encode (:708-713 -> from_raster :405-416 ->
 *                          extract_coefficients :179-225) fused with quantization::encode
 *                          (stages/quantization.rs:7-25).  Call sites: encoder.rs:26-33.
 *   fri_decode_tq*         quantization::decode (quantization.rs:27-45) fused with
 *                          wavelet_transform::decode (:715-717 -> from_wavelet :308-322 ->
 *                          extract_values :358-381, images.rs:103-111).  Call sites:
 *                          decoder.rs:27-34.
 *
 * Data layouts (identical to the reference's):
 *   pixels        HWC interleaved, index (y*W + x)*C + ch (images.rs:94); u8 (reference) or
 *                 little-endian u16 (extension, sample_bytes = 2).
 *   coefficients  int32 [n_tiles][C][2^depth], heap order: [0] = DC (:221), [pos] = residue of
 *                 tree node pos.  A coefficient the reference holds as `None` is 0 on encode
 *                 output and ignored on decode input; fri_plan_masks() says which are `Some`.
 *   tile order    defined by the plan: fri_plan_centers() returns the centre (re = x, im = y)
 *                 of tile i.  (The reference keeps tiles in a HashMap keyed by centre, so it
 *                 has no order of its own.)
 *
 * All functions return 0 on success or a negative FRI_E_* code; fri_last_error() returns a
 * thread-local message for the last failure.  There is NO CPU fallback: without a CUDA
 * device every compute entry point fails with FRI_E_CUDA.
 */
#ifndef FRI_CUDA_H
#define FRI_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRI_OK 0
#define FRI_E_INVALID (-1) /* bad argument */
#define FRI_E_CUDA (-2)    /* CUDA runtime / driver error, or no device */
#define FRI_E_NOMEM (-3)
#define FRI_E_UNSUPPORTED (-4)

/* quantization::decode behaviour */
#define FRI_DEQUANT_DIVIDE 0   /* reference-faithful: divides again (quantization.rs:37) */
#define FRI_DEQUANT_MULTIPLY 1 /* true dequantizer (not what the reference does) */

#define FRI_BASE_DEPTH 9 /* BASE_FRAC_DEPTH, wavelet_transform.rs:39 */

typedef struct fri_plan fri_plan;

/* Version / build info: "libfri_cuda <ver> sm_100a". */
const char *fri_version(void);
const char *fri_last_error(void);

/* Number of visible CUDA devices (0 if none / no driver). */
int fri_device_count(void);

/*
 * Builds the lattice for a (width, height, channels) image: the BFS of fractal_divide, the
 * retain rule, the tile order and the launch geometry.  Pure host work; if device >= 0 the
 * plan's tables are also uploaded to that device (required by every *_device / host entry
 * point), if device < 0 the plan is host-only (metadata queries work, compute calls fail).
 *   depth         9 is the only depth the reference reaches (BASE_FRAC_DEPTH); 10..24 is the
 *                 deep-tree extension (one fractal = 2^(depth-9) base tiles + coarse levels).
 *   sample_bytes  1 (reference) or 2 (u16 extension).
 */
int fri_plan_create(fri_plan **out, int device, uint32_t width, uint32_t height, uint32_t channels,
                    uint32_t depth, uint32_t sample_bytes);
void fri_plan_destroy(fri_plan *plan);

uint32_t fri_plan_num_tiles(const fri_plan *plan);       /* retained tiles (fractals of 2^depth leaves) */
uint32_t fri_plan_num_built(const fri_plan *plan);       /* tiles fractal_divide builds (incl. fringe) */
uint32_t fri_plan_num_full_tiles(const fri_plan *plan);  /* tiles with every leaf inside the image */
uint64_t fri_plan_coefs_per_frame(const fri_plan *plan); /* n_tiles * C * 2^depth */
uint64_t fri_plan_pixels_covered(const fri_plan *plan);  /* in-image pixels owned by retained tiles */

/* Launch geometry, for reports: info = {group_a, group_b, region_w, region_h, smem_pitch,
 * smem_bytes_per_cta, n_groups (CTAs per frame), n_base_tiles, threads_per_cta,
 * chunks_per_row, depth, depth - 9, fully owned 16-byte chunks per group, chunks with any owned
 * byte per group (both at phase 0), 0...}. */
int fri_plan_launch_info(const fri_plan *plan, int32_t info[16]);

/* centres[n_tiles][2] = (re, im) of each retained tile, in plan order. */
int fri_plan_centers(const fri_plan *plan, int32_t *centers);
/* masks[n_tiles][2^depth / 32]: bit (i & 31) of word (i >> 5) set <=> coefficient i is `Some`.
 * Channel-independent (validity is geometric). */
int fri_plan_masks(const fri_plan *plan, uint32_t *masks);

/*
 * Device-resident batch entry points (inputs and outputs already in HBM on the plan's
 * device).  `stream` is a cudaStream_t (NULL = default stream); the call only enqueues.
 *   pixels  n_frames consecutive HWC frames;  coefs  n_frames consecutive coefficient blocks.
 *   q       32-entry quantization matrix, host memory, every entry >= 1; q == NULL means
 *           all ones (get_quantization_matrix, quantization.rs:3-5).
 * The plan is read-only here (the depth > 9 scratch is allocated per call, stream-ordered), so
 * several calls may be in flight on different streams; only fri_plan_last_launches is last-writer-wins.
 * The kernels use programmatic dependent launch: on one stream, a call's kernel may become resident
 * while the previous kernel drains, but it waits for that kernel to complete (memory flushed) before
 * it reads or writes frame data, so stream order is preserved as usual.
 */
int fri_encode_tq_device(const fri_plan *plan, const void *d_pixels, uint32_t n_frames, const int32_t *q,
                         int32_t *d_coefs, void *stream);
int fri_decode_tq_device(const fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, const int32_t *q,
                         int dequant_mode, void *d_pixels, void *stream);

/*
 * The same two calls on int16 coefficient arrays (depth 9 and 8-bit samples only, else
 * FRI_E_UNSUPPORTED): every coefficient of an 8-bit image fits (see fri_encode_tq16 below), the
 * arithmetic is the same int32 arithmetic, the array is half the size — 3 bytes of HBM traffic
 * per sample instead of 5, reported as its own variant.  Encode saturates on the
 * (impossible) overflow; decode accepts any int16 content.
 */
int fri_encode_tq_device16(const fri_plan *plan, const void *d_pixels, uint32_t n_frames, const int32_t *q,
                           int16_t *d_coefs, void *stream);
int fri_decode_tq_device16(const fri_plan *plan, const int16_t *d_coefs, uint32_t n_frames, const int32_t *q,
                           int dequant_mode, void *d_pixels, void *stream);

/*
 * Host-buffer entry points (what libfri's stage functions call): copy in, run, copy out,
 * synchronise.  Frames are pipelined over the plan's internal streams.  Buffers obtained from
 * fri_host_alloc are page-locked and take the fast path (truly asynchronous copies at the link's
 * full rate — what the measured end-to-end numbers use, and what the libfri-cuda Rust crate's
 * PinnedBuf wraps).  Any other host memory is accepted in the default, synchronous mode and goes
 * through the CUDA driver's own pageable-copy staging (correct, roughly half the throughput, no
 * overlap); there is no bounce ring inside this library.  In asynchronous mode (fri_plan_set_async)
 * pageable buffers are rejected with FRI_E_INVALID.
 * One handle may alternate encode and decode calls, also in asynchronous mode: a device slot is
 * handed to the next frame only after the previous user's kernels and copies have finished.
 */
int fri_encode_tq(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, int32_t *coefs);
int fri_decode_tq(fri_plan *plan, const int32_t *coefs, uint32_t n_frames, const int32_t *q, int dequant_mode,
                  void *pixels);

/*
 * 16-bit transport of the same two stage calls (8-bit samples only): the host side of the copy is
 * int16 [n_frames][n_tiles][C][2^depth], same order.  Every coefficient the forward transform of
 * an 8-bit image can produce fits (|residue| <= 255, 0 <= low-pass <= 255; wavelet_transform.rs:
 * 211-218), and so does every coefficient a decodable container can hold (1024-symbol alphabet,
 * entropy_coding.rs:25), so the Rust glue can widen to the reference's Vec<Option<i32>> while it
 * applies the mask — and the PCIe copies, which bound these calls, carry half the bytes.  The
 * device computes in i32 exactly as above (at depth 9 the kernels read / write the int16 array
 * themselves, deep trees are repacked on the device); encode saturates on the (impossible) overflow
 * instead of wrapping.  FRI_E_UNSUPPORTED for sample_bytes = 2 (residues need 18 bits).
 */
int fri_encode_tq16(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, int16_t *coefs);
int fri_decode_tq16(fri_plan *plan, const int16_t *coefs, uint32_t n_frames, const int32_t *q, int dequant_mode,
                    void *pixels);

/*
 * Emission order (depth 9 only) — the order in which the reference's entropy coder consumes the
 * coefficients of one channel: entropy_coding.rs:283-329 walking sort_lattice
 * (wavelet_transform.rs:505-705): every DC, every root residue, then levels 1..8 in scan order.
 *   fri_plan_emission_order  order[n_tiles * 512]: entry i = tile_index * 512 + coefficient_index
 *                            (tile_index in plan order), `None` slots included.
 *   fri_plan_emission_count  number of `Some` slots = length of one channel's emitted stream.
 *   fri_emit_device          d_out[n_frames][C][count] <- coefficients in emission order with the
 *                            `None` slots dropped (what the three scans push, :287 / :298 / :314).
 *   fri_encode_tq_emit       host pixels -> host emitted streams: transform + quantization +
 *                            emission gather on the device, one D2H copy per frame.
 *   fri_emit_device16 /      the same with int16 output streams (8-bit samples only, like
 *   fri_encode_tq_emit16     fri_encode_tq16): half the bytes over PCIe; saturating.
 * All of them fail with FRI_E_UNSUPPORTED (and say so in fri_last_error) for the image sizes on
 * which the reference's own scan fails its assertion at wavelet_transform.rs:701, e.g. 257x300.
 */
/*
 *   fri_*_emit10             the same streams in the 10-bit packed transport (8-bit samples only): a
 *                            coefficient k travels as the symbol the reference's entropy coder works
 *                            in, pack_signed(k) = 2k (k >= 0) / -2k - 1 (k < 0) (utils.rs:34-40), of the
 *                            1024-symbol alphabet (entropy_coding.rs:25); 64 symbols -> 80 bytes,
 *                            symbol i of a block in bits [10 i, 10 i + 10), little-endian.  One
 *                            channel's stream is fri_plan_emission_packed_bytes() bytes (count rounded
 *                            up to a multiple of 64 symbols, padding = symbol 0), layout
 *                            [n_frames][C][packed_bytes].  1.25 bytes per coefficient over PCIe
 *                            instead of 2; encode saturates at -512 / +511 (cannot happen for an
 *                            8-bit image: |k| <= 255).
 */
/*
 *   fri_*_packed             the same layout at `bits` = 10 or 9 bits per symbol (8 * bits bytes per 64
 *                            symbols; fri_plan_emission_packed_size(plan, bits) bytes per channel).  9 bits
 *                            carry everything the transform of an 8-bit image produces (|k| <= 255, symbols
 *                            < 512) in 1.125 bytes per coefficient, saturating at -256 / +255; use 10 where
 *                            the streams may hold anything a container can (1024-symbol alphabet).
 *                            FRI_E_INVALID for any other width.
 */
uint64_t fri_plan_emission_count(fri_plan *plan);
uint64_t fri_plan_emission_packed_bytes(fri_plan *plan);
uint64_t fri_plan_emission_packed_size(fri_plan *plan, int bits);
int fri_emit_device_packed(fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, int bits, uint8_t *d_out, void *stream);
int fri_encode_tq_emit_packed(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, int bits, uint8_t *out);
int fri_unemit_device_packed(fri_plan *plan, const uint8_t *d_streams, uint32_t n_frames, int bits, int32_t *d_coefs,
                             void *stream);
int fri_decode_tq_emit_packed(fri_plan *plan, const uint8_t *streams, uint32_t n_frames, int bits, const int32_t *q,
                              int dequant_mode, void *pixels);
int fri_emit_device10(fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, uint8_t *d_out, void *stream);
int fri_encode_tq_emit10(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, uint8_t *out);
int fri_unemit_device10(fri_plan *plan, const uint8_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream);
int fri_decode_tq_emit10(fri_plan *plan, const uint8_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode,
                         void *pixels);
int fri_plan_emission_order(fri_plan *plan, uint32_t *order);
int fri_emit_device(fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, int32_t *d_out, void *stream);
int fri_encode_tq_emit(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, int32_t *out);
int fri_emit_device16(fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, int16_t *d_out, void *stream);
int fri_encode_tq_emit16(fri_plan *plan, const void *pixels, uint32_t n_frames, const int32_t *q, int16_t *out);

/*
 * The decoder side of the same order: the entropy decoder produces every channel's coefficients in
 * emission order (entropy_coding.rs:205-264 walks the same sorted lattice); fri_unemit_device*
 * places them back into the dense [n_tiles][C][512] blocks (`None` slots 0, where the reference's
 * from_metadata lattice, wavelet_transform.rs:392-403, holds no value), fri_decode_tq_emit* does
 * that and dequantization + inverse transform in one call: host streams -> host pixels.
 */
int fri_unemit_device(fri_plan *plan, const int32_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream);
int fri_unemit_device16(fri_plan *plan, const int16_t *d_streams, uint32_t n_frames, int32_t *d_coefs, void *stream);
int fri_decode_tq_emit(fri_plan *plan, const int32_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode,
                       void *pixels);
int fri_decode_tq_emit16(fri_plan *plan, const int16_t *streams, uint32_t n_frames, const int32_t *q, int dequant_mode,
                         void *pixels);

/*
 * Prediction + context bucketing on the device, encode side (depth 9): what
 * prediction::encode (stages/prediction.rs:224-323) computes per coefficient on the host before the rANS
 * coder — get_lf_context_bucket (:86-149) for every DC and root residue, get_hf_context_bucket (:151-207)
 * over ContextModeler::get_neighbour_values (context_modeling.rs:25-77) for levels 1..8, assign_bucket
 * (:55-68) and pack_signed (utils.rs:34-40).
 *   d_coefs        quantized dense blocks [n_frames][n_tiles][C][512] (`None` slots 0, as fri_encode_tq_device
 *                  leaves them);
 *   value_params,  host float [C][3][6] each: the value / width predictor parameters of the three layer sets
 *   width_params   (index 0: level 8, 1: level 7, 2: levels 1..6; prediction.rs:164-178).  They are INPUTS: the
 *                  reference fits them with an f32 SVD from crates that are not in its tree;
 *   d_bucket       uint8  [n_frames][C][count]  context bucket 0..9 of every `Some` coefficient, emission order
 *   d_pred         int32  [n_frames][C][count]  the prediction (`as i32` of the f32 predictor)
 *   d_sym          uint16 [n_frames][C][count]  pack_signed(value - prediction), saturated to 16 bits
 *   d_hist         uint32 [n_frames][C][10][1024]  symbol counts per context (AnsContext::bump_freq)
 *   d_overflow     optional uint32: number of symbols >= 1024 — the reference would panic on the first one
 *                  (entropy_coding.rs:99).
 * f32 arithmetic is evaluated in the reference's order without fused multiply-adds, so (bucket, prediction)
 * are bit-identical to a Rust evaluation of the same parameters.  The decoder side is inherently serial
 * (every prediction reads already-decoded neighbours, entropy_coding.rs:205-264) and stays on the host.
 */
int fri_predict_device(fri_plan *plan, const int32_t *d_coefs, uint32_t n_frames, const float *value_params,
                       const float *width_params, uint8_t *d_bucket, int32_t *d_pred, uint16_t *d_sym, uint32_t *d_hist,
                       uint32_t *d_overflow, void *stream);

/*
 * One image split over several GPUs ("a single huge image may additionally be split by tile-index
 * ranges across GPUs"; depth 9).  Part `part` of `n_parts` is a contiguous range of the plan's tile groups:
 *   fri_plan_part   groups [group_begin, group_end), the tiles [tile_begin, tile_end) (plan order) whose coefficient
 *                   blocks the part produces / consumes, and the pixel rows [row_begin, row_end) its groups read
 *                   (encode) or write into (decode).  Tile ranges of the parts are disjoint and cover the plan; row
 *                   ranges of neighbouring parts overlap by the halo of the tiles that straddle the cut.
 *   fri_encode_tq_device_part   d_pixel_rows points at pixel row `row_begin` (a band of row_end - row_begin rows,
 *                   frame row stride), d_coef_tiles at the block of tile `tile_begin`; only those are touched.
 *   fri_decode_tq_device_part   the inverse.  Only pixels OWNED by the part's tiles are written: in the rows two
 *                   parts share each writes its own pixels and leaves the others alone, so bands that start from
 *                   zero merge by addition (the one exchange step of the split: the overlap rows between
 *                   neighbours; see frave_b200/sharding.py and bench.py --workload image16k).
 *   fri_plan_groups_in_rows     the smallest contiguous sub-range [first, last) of groups [group_begin, group_end)
 *                   that contains every group touching pixel rows [row_begin, row_end), and the rows
 *                   [span_begin, span_end) that sub-range touches in all.
 *   fri_decode_tq_device_groups  the inverse transform of an arbitrary range of groups: d_coef_tiles points at the
 *                   block of tile `tile_first`, d_pixel_rows at frame row `row_first` (may be negative: a band with
 *                   a margin above row 0).  With a PEER-MAPPED band as the target this is the halo exchange done by
 *                   the transform kernel's own stores: a rank re-runs its groups along the cut into the neighbour's
 *                   band (which needs a margin of span rows around the shared ones), no zeroing, no merge pass —
 *                   only a barrier afterwards (bench.py --workload image16k, exchange "peer").
 * No collective inside the library; one frame per call; FRI_E_UNSUPPORTED at depth > 9.
 */
int fri_plan_part(const fri_plan *plan, uint32_t part, uint32_t n_parts, uint32_t *group_begin, uint32_t *group_end,
                  uint32_t *tile_begin, uint32_t *tile_end, uint32_t *row_begin, uint32_t *row_end);
int fri_encode_tq_device_part(const fri_plan *plan, const void *d_pixel_rows, const int32_t *q, int32_t *d_coef_tiles,
                              uint32_t part, uint32_t n_parts, void *stream);
int fri_decode_tq_device_part(const fri_plan *plan, const int32_t *d_coef_tiles, const int32_t *q, int dequant_mode,
                              void *d_pixel_rows, uint32_t part, uint32_t n_parts, void *stream);
int fri_plan_groups_in_rows(const fri_plan *plan, uint32_t group_begin, uint32_t group_end, uint32_t row_begin, uint32_t row_end,
                            uint32_t *first, uint32_t *last, uint32_t *span_begin, uint32_t *span_end);
int fri_decode_tq_device_groups(const fri_plan *plan, const int32_t *d_coef_tiles, uint32_t tile_first, const int32_t *q,
                                int dequant_mode, void *d_pixel_rows, int32_t row_first, uint32_t group_begin, uint32_t group_end,
                                void *stream);

/*
 * The host side of the codec behind the transform (depth 9, 8-bit samples):
 * the reference keeps context modelling, rANS and the `frif` container on the host, and so does this
 * library — in C++ (frave_b200/csrc/fri_codec.cpp), restating stages/entropy_coding.rs:32-176, :205-449,
 * stages/serialize.rs:40-268, context_modeling.rs:79-214 and the published rans64 algorithm behind the
 * un-vendored `rans 0.2.1` crate.  PARITY UNPINNED (no reference build exists here): containers written here
 * decode here bit-exactly; byte identity with the reference's `.frv` files is not claimed.
 *   fri_fit_parameters  host coefs [n_tiles][C][512] -> value / width predictor parameters float [C][3][6]
 *                       (least squares; the solver is NOT lstsq / nalgebra's f32 SVD).  The normal equations
 *                       are exact integer sums (the width fit's targets in 1/256 fixed point).
 *   fri_fit_device      the same fit for one device-resident frame: the sums are accumulated by a kernel (two
 *                       passes over the coefficients), the 6 x 6 systems solved on the host; synchronizes
 *                       `stream`.  Integer sums are order-independent: the parameters are bit-identical to
 *                       fri_fit_parameters on the same coefficients.
 *   fri_predict_host    the host form of fri_predict_device (same outputs, host arrays, one frame): what the
 *                       serial entropy decoder evaluates per coefficient
 *   fri_frv_pack        symbols + buckets + histograms (from fri_predict_device / _host) -> container bytes:
 *                       AnsContext tables from the Laplace model, 10 interleaved 64-bit rANS streams fed in
 *                       reverse emission order, serialize.rs layout.  *out is malloc'd: fri_frv_free.
 *                       colorspace: 0 = default (Luma for 1 channel, RGB for 3), 1 Luma, 2 RGB, 3 YCbCr.
 *   fri_frv_unpack      container bytes -> dense coefficient blocks (serial: every prediction reads already
 *                       decoded neighbours, entropy_coding.rs:205-264; channels run on separate host threads)
 *   fri_frv_info        width / height / channels of a container
 *   fri_frv_encode      host pixels -> container bytes: transform + quantization, parameter fit (fri_fit_device),
 *                       prediction + buckets + histograms on the device, rANS + container on the host
 *                       (FRIEncoder::encode, encoder.rs:87-109)
 *   fri_frv_decode      container bytes -> host pixels: fri_frv_unpack, then dequantization + inverse
 *                       transform on the device (FRIDecoder::decode, decoder.rs:48-59)
 * fri_frv_encode / fri_frv_decode keep pinned staging memory with the plan from their first call on (3 bytes per
 * emitted coefficient for the encoder, one frame of dense blocks for the decoder; released by fri_plan_destroy),
 * and — like every host-buffer entry point — serve one caller per plan handle at a time.
 * FRI_E_UNSUPPORTED where the reference itself panics: a residual outside the 1024-symbol alphabet
 * (entropy_coding.rs:99), or an image size whose sort_lattice scan asserts (wavelet_transform.rs:701).
 */
int fri_fit_parameters(fri_plan *plan, const int32_t *coefs, float *value_params, float *width_params);
int fri_fit_device(fri_plan *plan, const int32_t *d_coefs, float *value_params, float *width_params, void *stream);
int fri_predict_host(fri_plan *plan, const int32_t *coefs, const float *value_params, const float *width_params,
                     uint8_t *bucket, int32_t *pred, uint16_t *sym, uint32_t *hist, uint32_t *overflow);
int fri_frv_pack(fri_plan *plan, int colorspace, const float *value_params, const float *width_params,
                 const uint8_t *bucket, const uint16_t *sym, const uint32_t *hist, uint8_t **out, size_t *out_len);
int fri_frv_unpack(fri_plan *plan, const uint8_t *bytes, size_t len, int32_t *coefs);
int fri_frv_info(const uint8_t *bytes, size_t len, uint32_t *width, uint32_t *height, uint32_t *channels);
int fri_frv_encode(fri_plan *plan, const void *pixels, const int32_t *q, int colorspace, uint8_t **out, size_t *out_len);
int fri_frv_decode(fri_plan *plan, const uint8_t *bytes, size_t len, const int32_t *q, int dequant_mode, void *pixels);
void fri_frv_free(uint8_t *bytes);

/*
 * How the host-buffer entry points stream one frame through the device: in `bands` consecutive
 * bands of tile groups (1..8), each with its own copy in, kernels and copy out on three
 * event-chained streams, so that a single call overlaps its own host->device and device->host
 * traffic.  0 (default) = automatic (8 for a 4096x4096 image).  With several handles driven from
 * concurrent host threads the overlap comes from the other thread's call and one band per frame
 * (larger copies, fewer events) is faster: measured 4.5 -> 5.0 GPix/s for an encoder thread + a
 * decoder thread on 4096x4096 RGB with the 16-bit transport.
 */
int fri_plan_set_bands(fri_plan *plan, int bands);

/*
 * Asynchronous mode of the host-buffer entry points ("async variants enqueue on
 * the handle's stream and expose sync()"): with fri_plan_set_async(plan, 1) the fri_encode_tq* /
 * fri_decode_tq* / *_emit* calls return once their copies and kernels are enqueued on the plan's
 * streams, and fri_plan_sync(plan) waits for them.  The host buffers must be pinned (fri_host_alloc;
 * anything else is rejected with FRI_E_INVALID) and stay untouched until the sync.  One host thread can this way keep an encoder handle and a
 * decoder handle busy at once — the same full-duplex overlap two threads get.  Errors of the
 * enqueued work surface at the sync (or at the next call).
 */
int fri_plan_set_async(fri_plan *plan, int on);

/*
 * Hint for streams of independent frames through the *_device entry points (depth 9): with
 * fri_plan_set_independent_calls(plan, 1) the CALLER PROMISES that every fri_encode_tq_device* /
 * fri_decode_tq_device* call enqueued on a stream neither reads nor overwrites a buffer that the call
 * enqueued just before it on the same stream writes or reads (e.g. frame k + 1 while frame k is in flight).
 * The kernels then skip the programmatic-dependency wait, so the next launch's CTAs fill the slots the
 * previous launch's last, partly empty wave leaves idle — the same effect a batched launch (n_frames > 1)
 * has: measured 43 -> 39 us (encode) and 52 -> 47 us (decode) per 4096x4096 RGB frame.  Dependent calls
 * (encode then decode of the same coefficients) with the hint set are a data race.  Default: off.  Ignored
 * by deep trees and by the host-buffer entry points.
 */
int fri_plan_set_independent_calls(fri_plan *plan, int on);
int fri_plan_sync(fri_plan *plan);

/* Pinned host memory (cudaHostAlloc) for the host-buffer entry points. */
int fri_host_alloc(void **out, size_t bytes);
void fri_host_free(void *p);

/* Launch statistics of the last *_device / host call on this plan (kernel launches issued). */
uint32_t fri_plan_last_launches(const fri_plan *plan);

/* The kernels' truncating division value / q (q >= 1) evaluated on the host: the same
 * multiply-high + shift routine the device code uses for quantization.rs:19 / :37 (for tests). */
int32_t fri_quant_divide(int32_t value, int32_t q);
int32_t fri_quant_divide_magic(int32_t value, int32_t q); /* multiply-high path even for powers of two */
/* The encoder's narrow-range variant (exact for |value| <= 65535, all an 8/16-bit image can produce). */
int32_t fri_quant_divide_small(int32_t value, int32_t q);

#ifdef __cplusplus
}
#endif
#endif

// fri_codec.h — the host side of the codec behind the transform: context model, rANS and the `frif`
// container, plus the host form of the predictor the decoder needs.
// Pure C++ (no CUDA): the north star keeps the entropy coder on the host.
//
// Restates (crates/libfri/src/...):
//   stages/entropy_coding.rs:32-176   AnsContext: Laplace-model frequency tables, normalisation, cdf
//   stages/entropy_coding.rs:205-264  decode_symbol;  :266-352 encode (symbols pushed in reverse)
//   stages/entropy_coding.rs:354-449  decode (three scans, serial prediction)
//   stages/prediction.rs:55-84, 86-207, 220-222   buckets, predictors, laplace_distribution
//   stages/serialize.rs:40-268        container
//   context_modeling.rs:79-214        predictor parameter fit (see fit_parameters for what is NOT restated)
// and the published algorithm of the un-vendored dependency `rans = "0.2.1"` (a wrapper of ryg_rans'
// rans64.h): 64-bit state, lower bound 2^31, 32-bit renormalisation words, written back to front.
// PARITY UNPINNED: neither the crate nor a reference build exists in this environment; the multi-stream
// layout (states flushed in index order, read back in index order, hence the reference's
// `CONTEXT_AMOUNT - bucket - 1` at entropy_coding.rs:239) is inferred from the reference's call sites.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "fri_plan.h"

namespace fri {
namespace codec {

constexpr int kContexts = 10;    // CONTEXT_AMOUNT, prediction.rs:15
constexpr int kAlphabet = 1024;  // ALPHABET_SIZE, entropy_coding.rs:25

uint32_t pack_signed(int32_t k);    // utils.rs:34-40
int32_t unpack_signed(uint32_t k);  // utils.rs:42-48
int assign_bucket(float width);     // prediction.rs:55-68
float width_from_bucket(int bucket);  // prediction.rs:70-84

// entropy_coding.rs:32-176
struct AnsContext {
    std::array<uint32_t, kAlphabet> freqs{};
    std::array<uint32_t, kAlphabet> cdf{};
    std::vector<uint16_t> off_distribution_values;
    uint32_t max_freq_bits = 0;

    void fill_with_laplace(int bucket);                                  // :82-96
    std::array<uint32_t, kAlphabet> normalize_freqs(uint32_t target);    // :119-159
    void finalize_context(bool normalize, int bucket);                   // :102-117
};

// Encoder-side context of one bucket from its symbol counts (prediction.rs:302-304).
AnsContext context_from_counts(const uint32_t *counts, int bucket);
// Decoder-side context from what the container stores (serialize.rs:216-236).
AnsContext context_from_header(uint32_t max_freq_bits, std::vector<uint16_t> off_distribution_values, int bucket);

// N interleaved 64-bit rANS coders sharing one word stream (rans::B64RansEncoderMulti / DecoderMulti).
// One (start, freq) of a context prepared for coding without a division (ryg_rans' Rans64EncSymbol).
struct RansEncSymbol {
    uint64_t rcp_freq = 0, bias = 0, cmpl_freq = 0, x_max = 0;
    uint32_t rcp_shift = 0, codable = 0;  // codable == 0: zero frequency
    RansEncSymbol() = default;
    RansEncSymbol(uint32_t start, uint32_t freq, uint32_t scale_bits);
};
class RansEncoderMulti {
public:
    explicit RansEncoderMulti(int n);
    void put_at(int index, uint32_t start, uint32_t freq, uint32_t scale_bits);
    void put_at(int index, const RansEncSymbol &s);  // the same step, division-free
    void reserve_words(size_t n);
    void flush_all();
    std::vector<uint8_t> data() const;  // the bytes in decoding order
private:
    std::vector<uint64_t> state_;
    std::vector<uint32_t> words_;  // in emission order (the stream is this list reversed)
};
class RansDecoderMulti {
public:
    RansDecoderMulti(int n, const uint8_t *data, size_t len);
    uint32_t get_at(int index, uint32_t scale_bits) const;
    void advance_at(int index, uint32_t start, uint32_t freq, uint32_t scale_bits);  // advance_step + renorm
    bool overrun() const { return overrun_; }
private:
    uint32_t next_word();
    std::vector<uint64_t> state_;
    const uint8_t *data_;
    size_t len_, pos_ = 0;
    bool overrun_ = false;
};

// Host predictor over dense quantized coefficient blocks [n_tiles][C][512] (`None` slots 0): the same
// arithmetic as fri_predict_kernel, used by the (inherently serial) entropy decoder and by the tests.
struct Predictor {
    const LatticeIndex &lat;
    const int32_t *centers;  // [n_tiles][2]
    int channels;
    Vec2 nearby[10][6];
    struct NodeStep { int16_t heap; int8_t cell; };  // heap < 0: no node there in any tile; cell: (db + 1) * 3 + (da + 1)
    struct HeapSteps { NodeStep regular[6], alt[4], probe[4]; };
    HeapSteps steps[kTileLeaves];       // per heap index: where its neighbours sit, relative to the tile
    std::vector<int32_t> adjacent;      // [n_tiles][9] plan index of the tile one lattice step away, -1 if none
    int lf_cell[3];                     // adjacency cells of the tiles at centre + v9[4], v9[5], v9[0] (prediction.rs:86-149)
    Predictor(const LatticeIndex &l, const int32_t *c, int ch);
    int step_tile(int tile, const NodeStep &n) const;
    // neighbour values of the level-`level` node `heap` of `tile` (context_modeling.rs:25-77), levels 1..8
    void neighbour_values(const int32_t *coefs, int tile, int heap, int ch, int32_t v[6]) const;
    void lf(const int32_t *coefs, int tile, int heap, int ch, int &bucket, int32_t &prediction) const;
    void hf(const int32_t *coefs, int tile, int heap, int ch, const float vp[3][6], const float wp[3][6], int &bucket,
            int32_t &prediction) const;
};

// Predictor parameters of every channel: value[C][3][6], width[C][3][6] (layer set 0: level 8, 1: level 7,
// 2: levels 1..6).  Least squares like context_modeling.rs:144-214, but solved through the 6 x 6 normal
// equations with a pseudo-inverse for rank-deficient cases — NOT lstsq 0.6 / nalgebra's f32 SVD, which are not
// in the reference tree; the parameters are stored in the container, so any fit decodes.
//
// The normal equations are accumulated in EXACT integer arithmetic (every regressor is an integer; the width
// fit's target |coefficient - prediction| is taken in 1/kFitScale fixed point), so that the host fit below and
// the device fit (fri_fit_kernel: per-thread sums, shuffles, atomics — any order) produce the same sums and,
// through the same host solve, bit-identical parameters.  Sums wrap modulo 2^64 on both sides.
constexpr int kFitScale = 256;
struct FitSums {
    uint64_t a[21] = {};  // upper triangle of sum w w^T, row-major
    uint64_t b[6] = {};   // sum w * y
    void add(const int64_t w[6], int64_t y);
    void merge(const FitSums &o);
};
inline int fit_layer_set(int level) { return level < kBaseDepth - 2 ? 2 : (level == kBaseDepth - 2 ? 1 : 0); }
int64_t width_target(float coef, float prediction);                    // |coef - prediction| * kFitScale, truncated
void solve_fit(const FitSums &n, double b_scale, float out[6]);        // minimum-norm solution of a x = b / b_scale
// Rows of the reference's width matrices that stay [1, 0, 0, 0, 0, 0] -> 0 (None coefficients, two spare rows per
// tile in set 2): a plan constant per layer set, added to a[0][0].
void fit_zero_rows(const Plan &plan, const std::vector<uint8_t> &some, uint64_t rows[3]);
void fit_parameters(const Plan &plan, const LatticeIndex &lat, const std::vector<uint8_t> &some, const int32_t *coefs,
                    float *value_params, float *width_params, int n_threads);

// Per-channel payload of the container.
struct ChannelPayload {
    float value_params[3][6];
    float width_params[3][6];
    std::vector<AnsContext> contexts;  // 10
    std::vector<uint8_t> data;
};

// entropy_coding.rs:266-352 for one channel: symbols / buckets of the `Some` coefficients in emission
// order -> rANS bytes (pushed in reverse).  Returns an error for a symbol outside the alphabet.
std::string entropy_encode_channel(const uint16_t *sym, const uint8_t *bucket, size_t count, const uint32_t *hist,
                                   ChannelPayload &out);
// entropy_coding.rs:354-449 for one channel: decodes into the dense blocks (which must hold the already
// decoded channels / zeros) following `emit_src` (tile * 512 + heap, emission order).
std::string entropy_decode_channel(const ChannelPayload &in, const Predictor &pred, const std::vector<uint32_t> &emit_src, int ch,
                                   int32_t *coefs);

// serialize.rs:48-117 / :119-268.  colorspace: 1 = Luma, 2 = RGB, 3 = YCbCr (images.rs:23-29).
std::vector<uint8_t> serialize(uint32_t height, uint32_t width, int colorspace, const std::vector<ChannelPayload> &channels);
std::string deserialize(const uint8_t *bytes, size_t len, uint32_t &height, uint32_t &width, int &colorspace,
                        std::vector<ChannelPayload> &channels);

}  // namespace codec
}  // namespace fri

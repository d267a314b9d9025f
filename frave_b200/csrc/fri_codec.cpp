// fri_codec.cpp — host side of the codec behind the transform (see fri_codec.h for the citations).
#include "fri_codec.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace fri {
namespace codec {

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
uint32_t pack_signed(int32_t k)  // utils.rs:34-40 (wrapping like release-mode Rust)
{
    return k >= 0 ? 2u * (uint32_t)k : (uint32_t)(-2 * (int64_t)k - 1);
}
int32_t unpack_signed(uint32_t k)  // utils.rs:42-48
{
    return (k % 2 == 0) ? (int32_t)(k / 2) : (int32_t)(k + 1) / -2;
}

static uint32_t f32_as_u32(float x)  // Rust `as u32`: saturating, NaN -> 0
{
    if (!(x > 0.0f)) return 0;
    if (x >= 4294967296.0f) return 0xffffffffu;
    return (uint32_t)x;
}
static int32_t f32_as_i32(float x)  // Rust `as i32`
{
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

int assign_bucket(float width)  // prediction.rs:55-68
{
    // thresholds 3, 5, 6, 8, 12, 16, 20, 25, 30 — as a table: the serial entropy decoder calls this per symbol and
    // a compare chain mispredicts on every other one
    static const uint8_t kBucket[31] = {0, 0, 0, 1, 1, 2, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 7, 8, 8, 8, 8, 8, 9};
    const uint32_t w = f32_as_u32(width);
    return kBucket[w < 30 ? w : 30];
}

float width_from_bucket(int bucket)  // prediction.rs:70-84
{
    static const float w[10] = {2.5f, 4.5f, 6.3f, 8.5f, 12.7f, 16.f, 20.f, 24.f, 28.f, 36.f};
    return bucket < 10 ? w[bucket] : 50.f;
}

static float laplace_distribution(float x, float center, float width)  // prediction.rs:220-222
{
    return std::exp(-std::fabs(x - center) / width) / (2.0f * width);
}

static size_t prev_power_two(size_t x)  // utils.rs:5-14 (ORs up to >> 16 only, like the reference)
{
    size_t num = x;
    num |= num >> 1;
    num |= num >> 2;
    num |= num >> 4;
    num |= num >> 8;
    num |= num >> 16;
    return num ^ (num >> 1);
}
static uint32_t trailing_zeros(size_t x) { return x == 0 ? 64u : (uint32_t)__builtin_ctzll((unsigned long long)x); }

// ---------------------------------------------------------------------------------------------
// AnsContext (entropy_coding.rs:32-176)
// ---------------------------------------------------------------------------------------------
void AnsContext::fill_with_laplace(int bucket)  // :82-96
{
    const float width = width_from_bucket(bucket);
    const float scale = (float)(int32_t)(1u << (max_freq_bits & 31));
    for (int j = 0; j < kAlphabet; ++j) {
        uint32_t &freq = freqs[j];
        const uint32_t laplace_value = f32_as_u32(laplace_distribution((float)unpack_signed((uint32_t)j), 0.f, width) * scale);
        const bool listed = std::find(off_distribution_values.begin(), off_distribution_values.end(), (uint16_t)j) !=
                            off_distribution_values.end();
        if (laplace_value == 0 && freq == 0 && listed) {
            freq = 1;
        } else if (freq != 0 && laplace_value == 0) {
            freq = 1;
            off_distribution_values.push_back((uint16_t)j);
        } else {
            freq = laplace_value;
        }
    }
}

std::array<uint32_t, kAlphabet> AnsContext::normalize_freqs(uint32_t target_total)  // :119-159
{
    std::array<uint32_t, kAlphabet> cum{};
    {
        uint32_t acc = 0;  // get_cdf, :63-74: exclusive prefix sums (wrapping)
        for (int i = 0; i < kAlphabet; ++i) {
            cum[i] = acc;
            acc += freqs[i];
        }
    }
    const uint32_t cur_total = cum[kAlphabet - 1] + freqs[kAlphabet - 1];
    if (cur_total != 0)
        for (int i = 1; i < kAlphabet; ++i) cum[i] = (uint32_t)(((uint64_t)target_total * cum[i]) / cur_total);
    // fixing 0 freq values
    for (int i = 0; i < kAlphabet - 1; ++i) {
        if (freqs[i] != 0 && cum[i + 1] == cum[i]) {
            uint32_t best_freq = UINT32_MAX;
            int best_steal = -1;
            for (int j = 0; j < kAlphabet - 1; ++j) {
                const uint32_t freq = cum[j + 1] - cum[j];
                if (freq > 1 && freq < best_freq) {
                    best_freq = freq;
                    best_steal = j;
                }
            }
            if (best_steal < 0) continue;
            if (best_steal < i) {
                for (int j = best_steal + 1; j <= i; ++j) cum[j] -= 1;
            } else {
                for (int j = i + 1; j <= best_steal; ++j) cum[j] += 1;
            }
        }
    }
    for (int i = 0; i < kAlphabet - 1; ++i) freqs[i] = cum[i + 1] - cum[i];
    freqs[kAlphabet - 1] = cum[kAlphabet - 1] - target_total;  // as written (:157): 0 whenever the tail symbol's model frequency is 0
    return cum;
}

void AnsContext::finalize_context(bool normalize, int bucket)  // :102-117
{
    if (max_freq_bits < 8) max_freq_bits = 8;
    fill_with_laplace(bucket);
    if (normalize) {
        cdf = normalize_freqs(1u << (max_freq_bits & 31));
    } else {
        uint32_t acc = 0;
        for (int i = 0; i < kAlphabet; ++i) {
            cdf[i] = acc;
            acc += freqs[i];
        }
    }
    uint32_t sum = 0;
    for (uint32_t f : freqs) sum += f;
    max_freq_bits = trailing_zeros(prev_power_two((size_t)sum));
}

AnsContext context_from_counts(const uint32_t *counts, int bucket)
{
    AnsContext c;
    uint32_t sum = 0;
    for (int i = 0; i < kAlphabet; ++i) {
        c.freqs[i] = counts[i];
        sum += counts[i];
    }
    c.max_freq_bits = trailing_zeros(prev_power_two((size_t)sum));  // prediction.rs:302-303
    if (sum == 0) c.max_freq_bits = 0;                              // an unused context: finalize raises it to 8
    c.finalize_context(true, bucket);
    return c;
}

AnsContext context_from_header(uint32_t max_freq_bits, std::vector<uint16_t> off, int bucket)
{
    AnsContext c;
    c.max_freq_bits = max_freq_bits;
    c.off_distribution_values = std::move(off);
    c.finalize_context(true, bucket);
    return c;
}

// ---------------------------------------------------------------------------------------------
// rANS, 64-bit states, 32-bit words (ryg_rans rans64.h as wrapped by rans 0.2.1)
// ---------------------------------------------------------------------------------------------
static constexpr uint64_t kRansL = 1ull << 31;

RansEncoderMulti::RansEncoderMulti(int n) : state_((size_t)n, kRansL) {}

void RansEncoderMulti::put_at(int index, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    uint64_t x = state_[index];
    const uint64_t x_max = ((kRansL >> scale_bits) << 32) * freq;
    if (x >= x_max) {
        words_.push_back((uint32_t)x);
        x >>= 32;
    }
    state_[index] = ((x / freq) << scale_bits) + (x % freq) + start;
}

RansEncSymbol::RansEncSymbol(uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    x_max = ((kRansL >> scale_bits) << 32) * freq;
    codable = 1;
    cmpl_freq = (1ull << scale_bits) - freq;  // modulo 2^64: the reference's wrapped last frequency may exceed 2^scale_bits
    if (freq < 2) {  // x / 1: mul_hi(x, 2^64 - 1) = x - 1, the bias makes up for it
        rcp_freq = ~0ull;
        rcp_shift = 0;
        bias = (uint64_t)start + (1ull << scale_bits) - 1;
    } else {
        uint32_t shift = 0;
        while (freq > (1ull << shift)) ++shift;
        rcp_freq = (uint64_t)((((unsigned __int128)1 << (shift + 63)) + freq - 1) / freq);
        rcp_shift = shift - 1;
        bias = start;
    }
}

void RansEncoderMulti::put_at(int index, const RansEncSymbol &s)
{
    uint64_t x = state_[index];
    if (x >= s.x_max) {
        words_.push_back((uint32_t)x);
        x >>= 32;
    }
    // x < 2^63 and rcp_freq = ceil(2^(63 + shift) / freq): the quotient is exact (Alverson), so this is
    // ((x / freq) << scale_bits) + x % freq + start without the division
    const uint64_t q = (uint64_t)(((unsigned __int128)x * s.rcp_freq) >> 64) >> s.rcp_shift;
    state_[index] = x + s.bias + q * s.cmpl_freq;
}

void RansEncoderMulti::reserve_words(size_t n) { words_.reserve(n); }

void RansEncoderMulti::flush_all()
{
    for (uint64_t x : state_) {  // index order; each flush lands in front of the previous one
        words_.push_back((uint32_t)(x >> 32));
        words_.push_back((uint32_t)x);
    }
}

std::vector<uint8_t> RansEncoderMulti::data() const
{
    std::vector<uint8_t> out(words_.size() * 4);
    size_t o = 0;
    for (size_t i = words_.size(); i-- > 0;) {
        const uint32_t w = words_[i];
        out[o++] = (uint8_t)w;
        out[o++] = (uint8_t)(w >> 8);
        out[o++] = (uint8_t)(w >> 16);
        out[o++] = (uint8_t)(w >> 24);
    }
    return out;
}

RansDecoderMulti::RansDecoderMulti(int n, const uint8_t *data, size_t len) : state_((size_t)n, 0), data_(data), len_(len)
{
    for (int i = 0; i < n; ++i) {
        const uint64_t lo = next_word(), hi = next_word();
        state_[i] = lo | hi << 32;
    }
}

uint32_t RansDecoderMulti::next_word()
{
    if (pos_ + 4 > len_) {
        overrun_ = true;
        return 0;
    }
    const uint32_t w = (uint32_t)data_[pos_] | (uint32_t)data_[pos_ + 1] << 8 | (uint32_t)data_[pos_ + 2] << 16 |
                       (uint32_t)data_[pos_ + 3] << 24;
    pos_ += 4;
    return w;
}

uint32_t RansDecoderMulti::get_at(int index, uint32_t scale_bits) const
{
    return (uint32_t)(state_[index] & ((1ull << scale_bits) - 1));
}

void RansDecoderMulti::advance_at(int index, uint32_t start, uint32_t freq, uint32_t scale_bits)
{
    const uint64_t mask = (1ull << scale_bits) - 1;
    uint64_t x = state_[index];
    x = (uint64_t)freq * (x >> scale_bits) + (x & mask) - start;
    if (x < kRansL) x = (x << 32) | next_word();
    state_[index] = x;
}

// ---------------------------------------------------------------------------------------------
// host predictor (prediction.rs:86-207, context_modeling.rs:25-77, wavelet_transform.rs:97-177)
// ---------------------------------------------------------------------------------------------
// The neighbour of a node is a fixed (tile step, heap index) pair per heap index: positions are tile centre +
// leaf offset + neighbour vector, and the residue map / leaf offsets do not depend on the tile.  The table is
// built by running the lattice queries the reference makes (global_position_map[level].get) for a virtual tile
// at the anchor; per query only "does that tile exist" is left for run time.
Predictor::Predictor(const LatticeIndex &l, const int32_t *c, int ch) : lat(l), centers(c), channels(ch)
{
    for (int d = 0; d < 10; ++d) {
        for (int j = 0; j < 6; ++j) nearby[d][j] = Vec2{0, 0};
        if (d >= 1) nearby_vectors(d, nearby[d]);
    }
    int max_t = -1;
    for (int32_t t : lat.tile_at) max_t = std::max(max_t, (int)t);
    adjacent.assign(((size_t)max_t + 1) * 9, -1);
    for (int b = 0; b < lat.nb; ++b)
        for (int a = 0; a < lat.na; ++a) {
            const int32_t t = lat.tile_at[(size_t)b * lat.na + a];
            if (t < 0) continue;
            for (int db = -1; db <= 1; ++db)
                for (int da = -1; da <= 1; ++da) {
                    const int aa = a + da, bb = b + db;
                    if (aa >= 0 && bb >= 0 && aa < lat.na && bb < lat.nb)
                        adjacent[(size_t)t * 9 + (db + 1) * 3 + (da + 1)] = lat.tile_at[(size_t)bb * lat.na + aa];
                }
        }
    constexpr Vec2 l9 = kLiterals[kBaseDepth], l10 = kLiterals[kBaseDepth + 1];
    auto ref_at = [&](int level, int x, int y) {  // LatticeIndex::node_at for a tile centred at the anchor
        NodeStep r{-1, 4};
        const int k = lat.lut[LatticeIndex::mod512((x - lat.ax) + 181 * (y - lat.ay))];
        const int low = kBaseDepth - level;
        if (k & ((1 << low) - 1)) return r;
        const int64_t dx = x - lat.off[k].x - lat.ax, dy = y - lat.off[k].y - lat.ay;
        const int64_t na_ = dx * l10.y - (int64_t)l10.x * dy, nb_ = (int64_t)l9.x * dy - dx * l9.y;
        if (na_ % 512 != 0 || nb_ % 512 != 0) return r;
        const int64_t da = na_ / 512, db = nb_ / 512;
        if (da < -1 || da > 1 || db < -1 || db > 1) throw std::logic_error("a neighbour node lies beyond the adjacent tiles");
        r.heap = (int16_t)((1 << level) + (k >> low));
        r.cell = (int8_t)((db + 1) * 3 + (da + 1));
        return r;
    };
    {
        const int sel[3] = {4, 5, 0};
        for (int j = 0; j < 3; ++j) {  // whole-tile steps: the level-9 neighbour vectors are lattice vectors
            const Vec2 d = nearby[kBaseDepth][sel[j]];
            const int64_t na_ = (int64_t)d.x * l10.y - (int64_t)l10.x * d.y, nb_ = (int64_t)l9.x * d.y - (int64_t)d.x * l9.y;
            if (na_ % 512 != 0 || nb_ % 512 != 0 || std::llabs(na_ / 512) > 1 || std::llabs(nb_ / 512) > 1)
                throw std::logic_error("a level-9 neighbour vector is not a step to an adjacent tile");
            lf_cell[j] = (int)((nb_ / 512 + 1) * 3 + (na_ / 512 + 1));
        }
    }
    for (int heap = 0; heap < kTileLeaves; ++heap) {
        HeapSteps &hs = steps[heap];
        for (NodeStep &n : hs.regular) n = NodeStep{-1, 4};
        for (NodeStep &n : hs.alt) n = NodeStep{-1, 4};
        for (NodeStep &n : hs.probe) n = NodeStep{-1, 4};
        if (heap < 2) continue;
        const int level = 31 - __builtin_clz((unsigned)heap), d = kBaseDepth - level;
        const Vec2 o = lat.off[(heap - (1 << level)) << d];
        const int px = lat.ax + o.x, py = lat.ay + o.y;
        const Vec2 *nv = nearby[d];
        const Vec2 q[6] = {{px + nv[4].x, py + nv[4].y}, {px + nv[5].x, py + nv[5].y}, {px + nv[0].x, py + nv[0].y},
                           {px + nv[1].x, py + nv[1].y}, {px + nv[3].x, py + nv[3].y}, {px + nv[2].x, py + nv[2].y}};
        for (int j = 0; j < 6; ++j) hs.regular[j] = ref_at(level, q[j].x, q[j].y);
        if (d == 2) {
            hs.alt[0] = ref_at(level, px - 1 + nv[4].x, py - 1 + nv[4].y);  // up-left
            hs.alt[1] = ref_at(level, px - 1, py - 1);                      // up-right
            hs.alt[2] = ref_at(level, px + 1, py + 1);                      // down-left
            hs.alt[3] = ref_at(level, px + 1 + nv[1].x, py + 1 + nv[1].y);  // down-right
            hs.probe[0] = ref_at(2, px + nv[3].x, py + nv[3].y);            // the level-2 map, as written in the reference
            hs.probe[1] = ref_at(2, px + 1, py + 1);
            hs.probe[2] = ref_at(2, px + nv[0].x, py + nv[0].y);
            hs.probe[3] = ref_at(2, px - 1, py - 1);
        }
    }
}

inline int Predictor::step_tile(int tile, const NodeStep &n) const
{
    return n.heap < 0 ? -1 : adjacent[(size_t)tile * 9 + n.cell];
}

void Predictor::neighbour_values(const int32_t *coefs, int tile, int heap, int ch, int32_t v[6]) const
{
    const HeapSteps &hs = steps[heap];
    const int32_t *adj = adjacent.data() + (size_t)tile * 9;
    if (heap < 128 || heap >= 256) {  // every level but 7: the six regular positions
#pragma GCC unroll 6
        for (int j = 0; j < 6; ++j) {
            const NodeStep n = hs.regular[j];
            const int t = n.heap < 0 ? -1 : adj[n.cell];
            v[j] = t >= 0 ? coefs[(((size_t)t * channels + ch) << kBaseDepth) + (j < 3 ? n.heap : n.heap >> 1)] : 0;
        }
        return;
    }
    const NodeStep *sel[6] = {&hs.regular[0], &hs.regular[1], &hs.regular[2], &hs.regular[3], &hs.regular[4], &hs.regular[5]};
    {  // level 7 (depth 2): wavelet_transform.rs:115-177
        const bool alt_down = step_tile(tile, hs.probe[0]) < 0 && step_tile(tile, hs.probe[1]) >= 0;
        const bool alt_up = step_tile(tile, hs.probe[2]) < 0 && step_tile(tile, hs.probe[3]) >= 0;
        if (alt_up) { sel[1] = &hs.alt[0]; sel[2] = &hs.alt[1]; }
        if (alt_down) { sel[4] = &hs.alt[2]; sel[5] = &hs.alt[3]; }
    }
    for (int j = 0; j < 6; ++j) {
        const int t = step_tile(tile, *sel[j]);
        v[j] = t >= 0 ? coefs[(((size_t)t * channels + ch) << kBaseDepth) + (j < 3 ? sel[j]->heap : sel[j]->heap >> 1)] : 0;
    }
}

void Predictor::lf(const int32_t *coefs, int tile, int heap, int ch, int &bucket, int32_t &prediction) const
{
    int32_t v[3];
    for (int j = 0; j < 3; ++j) {
        const int t = adjacent[(size_t)tile * 9 + lf_cell[j]];
        v[j] = t >= 0 ? coefs[(((size_t)t * channels + ch) << kBaseDepth) + heap] : 0;
    }
    const uint32_t width = (uint32_t)std::abs((int64_t)v[0] - v[2]);
    bucket = assign_bucket((float)width);
    const int32_t hi = std::max(v[0], v[2]), lo = std::min(v[0], v[2]);
    prediction = v[1] >= hi ? hi : (v[1] <= lo ? lo : (int32_t)((uint32_t)v[0] + (uint32_t)v[2] - (uint32_t)v[1]));
}

void Predictor::hf(const int32_t *coefs, int tile, int heap, int ch, const float vp_all[3][6], const float wp_all[3][6], int &bucket,
                   int32_t &prediction) const
{
    const int level = 31 - __builtin_clz((unsigned)heap);
    const int layer = level < kBaseDepth - 2 ? 2 : (level == kBaseDepth - 2 ? 1 : 0);
    const float *vp = vp_all[layer], *wp = wp_all[layer];
    int32_t v[6];
    neighbour_values(coefs, tile, heap, ch, v);
    auto absdiff = [](int32_t a, int32_t b) { return (float)std::abs((int32_t)((uint32_t)a - (uint32_t)b)); };
    float width = wp[0];
    width = width + wp[1] * absdiff(v[0], v[3]);
    width = width + wp[2] * absdiff(v[1], v[2]);
    width = width + wp[3] * absdiff(v[4], v[5]);
    width = width + wp[4] * absdiff(v[1], v[5]);
    width = width + wp[5] * absdiff(v[2], v[4]);
    bucket = assign_bucket(width);
    float p = (float)v[0] * vp[0];
    for (int j = 1; j < 6; ++j) p = p + (float)v[j] * vp[j];
    prediction = f32_as_i32(p);
}

// ---------------------------------------------------------------------------------------------
// parameter fit (context_modeling.rs:79-214) — integer normal equations, see the header
// ---------------------------------------------------------------------------------------------
void FitSums::add(const int64_t w[6], int64_t y)
{
    int k = 0;
    for (int i = 0; i < 6; ++i) {
        b[i] += (uint64_t)w[i] * (uint64_t)y;
        for (int j = i; j < 6; ++j) a[k++] += (uint64_t)w[i] * (uint64_t)w[j];
    }
}

void FitSums::merge(const FitSums &o)
{
    for (int i = 0; i < 21; ++i) a[i] += o.a[i];
    for (int i = 0; i < 6; ++i) b[i] += o.b[i];
}

int64_t width_target(float coef, float prediction)
{
    const float r = std::fabs(coef - prediction) * (float)kFitScale;
    return r < 1.0e12f ? (int64_t)r : (int64_t)1 << 40;  // NaN / overflow: the same cap as the kernel's
}

// Minimum-norm least-squares solution of the symmetric system through a Jacobi eigen-decomposition
// (what an SVD-based lstsq returns, up to rounding).
void solve_fit(const FitSums &n, double b_scale, float out[6])
{
    double a[6][6], v[6][6], rhs[6];
    for (int i = 0, k = 0; i < 6; ++i) {
        rhs[i] = (double)(int64_t)n.b[i] / b_scale;
        for (int j = i; j < 6; ++j, ++k) a[i][j] = a[j][i] = (double)(int64_t)n.a[k];
    }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) v[i][j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i + 1; j < 6; ++j) off += a[i][j] * a[i][j];
        if (off < 1e-300) break;
        for (int p = 0; p < 6; ++p)
            for (int q = p + 1; q < 6; ++q) {
                if (std::fabs(a[p][q]) < 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < 6; ++k) {
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 6; ++k) {
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 6; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    double lmax = 0;
    for (int i = 0; i < 6; ++i) lmax = std::max(lmax, std::fabs(a[i][i]));
    double x[6] = {};
    for (int k = 0; k < 6; ++k) {
        const double lam = a[k][k];
        if (lmax == 0 || lam <= 1e-12 * lmax) continue;  // rank-deficient direction: contributes nothing (minimum norm)
        double proj = 0;
        for (int i = 0; i < 6; ++i) proj += v[i][k] * rhs[i];
        for (int i = 0; i < 6; ++i) x[i] += v[i][k] * proj / lam;
    }
    for (int i = 0; i < 6; ++i) out[i] = (float)x[i];
}

void fit_zero_rows(const Plan &plan, const std::vector<uint8_t> &some, uint64_t rows[3])
{
    rows[0] = rows[1] = rows[2] = 0;
    const int n_tiles = plan.geo.n_fractals;
    for (int tile = 0; tile < n_tiles; ++tile) {
        for (int heap = 2; heap < kTileLeaves; ++heap)
            if (!some[(size_t)tile * kTileLeaves + heap]) ++rows[fit_layer_set(31 - __builtin_clz((unsigned)heap))];
        rows[2] += 2;
    }
}

void fit_parameters(const Plan &plan, const LatticeIndex &lat, const std::vector<uint8_t> &some, const int32_t *coefs,
                    float *value_params, float *width_params, int n_threads)
{
    const Geometry &g = plan.geo;
    const int n_tiles = g.n_fractals, C = g.channels;
    const Predictor pred(lat, plan.centers.data(), C);
    n_threads = std::max(1, std::min(n_threads, std::max(1, n_tiles / 64)));
    auto parallel = [&](auto &&body) {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t) th.emplace_back([&, t] { body(t, (int)((int64_t)n_tiles * t / n_threads), (int)((int64_t)n_tiles * (t + 1) / n_threads)); });
        for (auto &x : th) x.join();
    };
    uint64_t zero_rows[3];
    fit_zero_rows(plan, some, zero_rows);
    for (int ch = 0; ch < C; ++ch) {
        // ---- value predictors: y ~ v . p per layer set
        std::vector<FitSums> acc((size_t)n_threads * 3);
        parallel([&](int t, int lo, int hi) {
            for (int tile = lo; tile < hi; ++tile)
                for (int heap = 2; heap < kTileLeaves; ++heap) {
                    if (!some[(size_t)tile * kTileLeaves + heap]) continue;
                    int32_t v[6];
                    pred.neighbour_values(coefs, tile, heap, ch, v);
                    const int64_t w[6] = {v[0], v[1], v[2], v[3], v[4], v[5]};
                    acc[(size_t)t * 3 + fit_layer_set(31 - __builtin_clz((unsigned)heap))].add(
                        w, coefs[(((size_t)tile * C + ch) << kBaseDepth) + heap]);
                }
        });
        float vp[3][6];
        for (int s = 0; s < 3; ++s) {
            FitSums n;
            for (int t = 0; t < n_threads; ++t) n.merge(acc[(size_t)t * 3 + s]);
            solve_fit(n, 1.0, vp[s]);
            std::memcpy(value_params + ((size_t)ch * 3 + s) * 6, vp[s], sizeof(float) * 6);
        }
        // ---- width predictors: |y - v . p| ~ w . [1, |v0-v3|, |v1-v2|, |v4-v5|, |v1-v5|, |v2-v4|]; rows the
        // reference's matrices keep at zero (None coefficients, two spare rows per tile in the last set) count as
        // [1, 0, 0, 0, 0, 0] -> 0 (fit_zero_rows)
        std::vector<FitSums> wacc((size_t)n_threads * 3);
        parallel([&](int t, int lo, int hi) {
            for (int tile = lo; tile < hi; ++tile)
                for (int heap = 2; heap < kTileLeaves; ++heap) {
                    if (!some[(size_t)tile * kTileLeaves + heap]) continue;
                    const int s = fit_layer_set(31 - __builtin_clz((unsigned)heap));
                    int32_t v[6];
                    pred.neighbour_values(coefs, tile, heap, ch, v);
                    float p = (float)v[0] * vp[s][0];
                    for (int j = 1; j < 6; ++j) p = p + (float)v[j] * vp[s][j];
                    auto ad = [](int32_t x, int32_t y) { return std::abs((int64_t)x - (int64_t)y); };
                    const int64_t w[6] = {1, ad(v[0], v[3]), ad(v[1], v[2]), ad(v[4], v[5]), ad(v[1], v[5]), ad(v[2], v[4])};
                    wacc[(size_t)t * 3 + s].add(w, width_target((float)coefs[(((size_t)tile * C + ch) << kBaseDepth) + heap], p));
                }
        });
        for (int s = 0; s < 3; ++s) {
            FitSums n;
            for (int t = 0; t < n_threads; ++t) n.merge(wacc[(size_t)t * 3 + s]);
            n.a[0] += zero_rows[s];
            float wp[6];
            solve_fit(n, (double)kFitScale, wp);
            std::memcpy(width_params + ((size_t)ch * 3 + s) * 6, wp, sizeof(float) * 6);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// entropy coding of one channel
// ---------------------------------------------------------------------------------------------
std::string entropy_encode_channel(const uint16_t *sym, const uint8_t *bucket, size_t count, const uint32_t *hist,
                                   ChannelPayload &out)
{
    out.contexts.clear();
    for (int b = 0; b < kContexts; ++b) out.contexts.push_back(context_from_counts(hist + (size_t)b * kAlphabet, b));
    // per (context, symbol): the division-free form of the coding step
    std::vector<RansEncSymbol> table((size_t)kContexts * kAlphabet);
    for (int b = 0; b < kContexts; ++b) {
        const AnsContext &c = out.contexts[b];
        for (int s = 0; s < kAlphabet; ++s)
            if (c.freqs[s] != 0 && c.max_freq_bits <= 31) table[(size_t)b * kAlphabet + s] = RansEncSymbol(c.cdf[s], c.freqs[s], c.max_freq_bits);
    }
    RansEncoderMulti enc(kContexts);
    enc.reserve_words(count / 2 + 64);
    for (size_t k = count; k-- > 0;) {  // entropy_coding.rs:332-334: pushed in reverse
        const uint32_t s = sym[k];
        const int b = bucket[k];
        if (s >= (uint32_t)kAlphabet)
            return "a residual falls outside the 1024-symbol alphabet (the reference panics at entropy_coding.rs:99)";
        if (b >= kContexts) return "context bucket out of range";
        const RansEncSymbol &e = table[(size_t)b * kAlphabet + s];
        if (!e.codable) return "symbol with zero model frequency (cannot be coded)";
        enc.put_at(b, e);
    }
    enc.flush_all();
    out.data = enc.data();
    return {};
}

std::string entropy_decode_channel(const ChannelPayload &in, const Predictor &pred, const std::vector<uint32_t> &emit_src, int ch,
                                   int32_t *coefs)
{
    if ((int)in.contexts.size() != kContexts) return "channel has " + std::to_string(in.contexts.size()) + " entropy contexts, expected 10";
    for (const AnsContext &c : in.contexts)
        if (c.max_freq_bits > 31) return "context with an impossible max_freq_bits";
    RansDecoderMulti dec(kContexts, in.data.data(), in.data.size());
    const int C = pred.channels;
    // slot -> symbol: a coarse table over the top kCoarseBits bits of the slot bounds the search in the cdf
    constexpr int kCoarseBits = 12;
    struct Coarse {
        int shift;
        std::vector<uint16_t> first;  // symbol owning the first slot of every cell, plus a sentinel
    };
    std::vector<Coarse> coarse(kContexts);
    for (int b = 0; b < kContexts; ++b) {
        const AnsContext &c = in.contexts[b];
        const int bits = (int)c.max_freq_bits, cells_bits = std::min(bits, kCoarseBits);
        coarse[b].shift = bits - cells_bits;
        coarse[b].first.resize(((size_t)1 << cells_bits) + 1);
        for (size_t i = 0; i < ((size_t)1 << cells_bits); ++i) {
            const uint32_t slot = (uint32_t)(i << coarse[b].shift);
            const int sym = (int)(std::upper_bound(c.cdf.begin(), c.cdf.end(), slot) - c.cdf.begin()) - 1;
            coarse[b].first[i] = (uint16_t)std::max(sym, 0);
        }
        coarse[b].first.back() = kAlphabet - 1;
    }
    for (uint32_t src : emit_src) {
        const int tile = (int)(src >> kBaseDepth), heap = (int)(src & (kTileLeaves - 1));
        int bucket;
        int32_t prediction;
        if (heap < 2) pred.lf(coefs, tile, heap, ch, bucket, prediction);
        else pred.hf(coefs, tile, heap, ch, in.value_params, in.width_params, bucket, prediction);
        const AnsContext &c = in.contexts[bucket];
        const int pos = kContexts - bucket - 1;  // entropy_coding.rs:239
        const uint32_t got = dec.get_at(pos, c.max_freq_bits);
        // find_nearest_or_equal + the "last index with that cdf" walk (:244-255): the symbol owning `got`,
        // i.e. the last index whose cdf is <= got; it lies between the owners of the cell's first slot and of
        // the next cell's
        const Coarse &cs = coarse[bucket];
        const uint32_t cell = got >> cs.shift;
        int symbol = cs.first[cell];
        const int sym_hi = cs.first[cell + 1];
        if (symbol != sym_hi) {  // several symbols share the cell: the last one whose cdf is <= got (branch-free count)
            const uint32_t *cdf = c.cdf.data();
            if (sym_hi - symbol <= 8) {
                int n = 0;
                for (int i = symbol + 1; i <= sym_hi; ++i) n += cdf[i] <= got;
                symbol += n;  // cdf is non-decreasing: the qualifying entries are a prefix
            } else {
                symbol = (int)(std::upper_bound(cdf + symbol, cdf + sym_hi + 1, got) - cdf) - 1;
            }
        }
        if (c.freqs[symbol] == 0) return "corrupt stream: decoded a symbol with zero frequency";
        dec.advance_at(pos, c.cdf[symbol], c.freqs[symbol], c.max_freq_bits);
        if (dec.overrun()) return "corrupt stream: entropy-coded data ends early";
        coefs[(((size_t)tile * C + ch) << kBaseDepth) + heap] = (int32_t)((uint32_t)unpack_signed((uint32_t)symbol) + (uint32_t)prediction);
    }
    return {};
}

// ---------------------------------------------------------------------------------------------
// container (serialize.rs)
// ---------------------------------------------------------------------------------------------
namespace {
const uint8_t kEHD[2] = {0xFF, 0xB2}, kDAT[2] = {0xFF, 0xB4}, kEOC[2] = {0xFF, 0xB8}, kPRD[2] = {0xFF, 0xBB}, kEOI[2] = {0xFF, 0xDF};

template <typename T>
void put_le(std::vector<uint8_t> &v, T x)
{
    for (size_t i = 0; i < sizeof(T); ++i) v.push_back((uint8_t)((uint64_t)x >> (8 * i)));
}
void put_f32(std::vector<uint8_t> &v, float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    put_le<uint32_t>(v, u);
}
}  // namespace

std::vector<uint8_t> serialize(uint32_t height, uint32_t width, int colorspace, const std::vector<ChannelPayload> &channels)
{
    std::vector<uint8_t> s;
    s.insert(s.end(), {'f', 'r', 'i', 'f'});
    put_le<uint32_t>(s, height);
    put_le<uint32_t>(s, width);
    uint32_t mdat = 0;
    mdat |= (uint32_t)colorspace << 30;
    mdat |= 1u << 28;  // FractalVariant::TameTwindragon (images.rs:49-55, encoder.rs:96)
    put_le<uint32_t>(s, mdat);
    for (const ChannelPayload &c : channels) {
        s.insert(s.end(), kPRD, kPRD + 2);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j) put_f32(s, c.value_params[i][j]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 6; ++j) put_f32(s, c.width_params[i][j]);
        for (const AnsContext &ctx : c.contexts) {
            s.insert(s.end(), kEHD, kEHD + 2);
            put_le<uint32_t>(s, ctx.max_freq_bits);
            put_le<uint64_t>(s, (uint64_t)ctx.off_distribution_values.size());  // usize
            for (uint16_t v : ctx.off_distribution_values) put_le<uint16_t>(s, v);
        }
        s.insert(s.end(), kDAT, kDAT + 2);
        put_le<uint64_t>(s, (uint64_t)c.data.size());
        s.insert(s.end(), c.data.begin(), c.data.end());
        s.insert(s.end(), kEOC, kEOC + 2);
    }
    s.insert(s.end(), kEOI, kEOI + 2);
    return s;
}

std::string deserialize(const uint8_t *b, size_t len, uint32_t &height, uint32_t &width, int &colorspace,
                        std::vector<ChannelPayload> &channels)
{
    size_t off = 0;
    auto need = [&](size_t n) { return off + n <= len; };
    auto u32 = [&]() { uint32_t v = (uint32_t)b[off] | (uint32_t)b[off + 1] << 8 | (uint32_t)b[off + 2] << 16 | (uint32_t)b[off + 3] << 24; off += 4; return v; };
    auto u64 = [&]() { uint64_t lo = u32(); uint64_t hi = u32(); return lo | hi << 32; };
    if (!need(16) || std::memcmp(b, "frif", 4) != 0) return "Invalid signature for FRIF image.";
    off = 4;
    height = u32();
    width = u32();
    const uint32_t mdat = u32();
    colorspace = (int)(mdat >> 30 & 3);
    if (colorspace == 0) return "Invalid metadata";
    if ((mdat >> 28 & 3) == 0) return "Invalid metadata";
    channels.clear();
    ChannelPayload cur{};
    bool have_params = false;
    for (;;) {
        if (!need(2)) return "Malformed image bytes";
        const uint8_t m0 = b[off], m1 = b[off + 1];
        off += 2;
        if (m0 != 0xFF) return "Malformed image bytes";
        if (m1 == kPRD[1]) {
            if (!need(36 * 4)) return "Malformed image bytes";
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 6; ++j) { uint32_t u = u32(); std::memcpy(&cur.value_params[i][j], &u, 4); }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 6; ++j) { uint32_t u = u32(); std::memcpy(&cur.width_params[i][j], &u, 4); }
            have_params = true;
        } else if (m1 == kEHD[1]) {
            if (!need(12)) return "Malformed image bytes";
            const uint32_t bits = u32();
            const uint64_t n = u64();
            if (n > (uint64_t)kAlphabet * 64 || !need((size_t)n * 2)) return "Malformed image bytes";
            std::vector<uint16_t> offv((size_t)n);
            for (auto &v : offv) { v = (uint16_t)(b[off] | b[off + 1] << 8); off += 2; }
            if (bits > 31) return "Malformed image bytes";
            cur.contexts.push_back(context_from_header(bits, std::move(offv), (int)cur.contexts.size()));
        } else if (m1 == kDAT[1]) {
            if (!need(8)) return "Malformed image bytes";
            const uint64_t n = u64();
            if (n > len || !need((size_t)n)) return "Malformed image bytes";
            cur.data.assign(b + off, b + off + n);
            off += (size_t)n;
        } else if (m1 == kEOC[1]) {
            if (!have_params) std::memset(cur.value_params, 0, sizeof(cur.value_params)), std::memset(cur.width_params, 0, sizeof(cur.width_params));
            channels.push_back(std::move(cur));
            cur = ChannelPayload{};
            have_params = false;
            if (channels.size() > 3) return "Malformed image bytes";
        } else if (m1 == kEOI[1]) {
            return {};
        } else {
            return "Malformed image bytes";
        }
    }
}

}  // namespace codec
}  // namespace fri

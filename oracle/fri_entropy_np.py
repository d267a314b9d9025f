"""Context model, rANS and `frif` container — CPU restatement, encoder side.

TEST INFRASTRUCTURE ONLY, PARITY UNPINNED (see oracle/fri_oracle.h): written independently of
frave_b200/csrc/fri_codec.cpp (Python integers and numpy float32 scalars, no shared code) so that a misreading
of the reference in one of the two shows up as a byte difference.  Paths are relative to /root/reference/.

  crates/libfri/src/stages/entropy_coding.rs:63-74     get_cdf
  crates/libfri/src/stages/entropy_coding.rs:82-96     fill_with_laplace
  crates/libfri/src/stages/entropy_coding.rs:102-117   finalize_context
  crates/libfri/src/stages/entropy_coding.rs:119-159   normalize_freqs
  crates/libfri/src/stages/entropy_coding.rs:266-352   encode (symbols pushed in reverse, flush_all, data)
  crates/libfri/src/stages/prediction.rs:70-84, 220-222, 302-304   widths, laplace_distribution, max_freq_bits
  crates/libfri/src/stages/serialize.rs:48-117         container
  crates/libfri/src/utils.rs:5-14, 42-48               get_prev_power_two, unpack_signed

rANS: the reference calls the un-vendored crate `rans 0.2.1` (B64RansEncoderMulti<10>), a wrapper of ryg_rans'
rans64.h — restated here from the published algorithm: 64-bit state starting at 2^31, one 32-bit word emitted
when the state would overflow, words written back to front; `flush_all` taken to flush the states in index
order (the reference's decoder reads state `CONTEXT_AMOUNT - bucket - 1`, entropy_coding.rs:239).
"""
from __future__ import annotations

import struct

import numpy as np

ALPHABET_SIZE = 1024
CONTEXT_AMOUNT = 10
F = np.float32
M32 = 0xFFFFFFFF
WIDTHS = [2.5, 4.5, 6.3, 8.5, 12.7, 16.0, 20.0, 24.0, 28.0, 36.0]  # prediction.rs:70-84


def prev_power_two(x: int) -> int:  # utils.rs:5-14
    num = x
    for s in (1, 2, 4, 8, 16):
        num |= num >> s
    return num ^ (num >> 1)


def trailing_zeros(x: int) -> int:
    return 64 if x == 0 else (x & -x).bit_length() - 1


def unpack_signed(k: int) -> int:  # utils.rs:42-48
    return k // 2 if k % 2 == 0 else -((k + 1) // 2)


def laplace(x, center, width):  # prediction.rs:220-222, f32 throughout
    x, center, width = F(x), F(center), F(width)
    with np.errstate(all="ignore"):
        return F(np.exp(F(-np.abs(F(x - center)) / width)) / F(F(2.0) * width))


def as_u32(x) -> int:
    x = float(x)
    if not x > 0.0:
        return 0
    return min(int(x), M32)


class AnsContext:
    def __init__(self):
        self.freqs = [0] * ALPHABET_SIZE
        self.cdf = [0] * ALPHABET_SIZE
        self.off_distribution_values: list[int] = []
        self.max_freq_bits = 0

    def get_cdf(self):
        out, acc = [], 0
        for f in self.freqs:
            out.append(acc)
            acc = (acc + f) & M32
        return out

    def fill_with_laplace(self, bucket: int):
        width = WIDTHS[bucket] if bucket < 10 else 50.0
        one = 1 << self.max_freq_bits
        if one >= 1 << 31:  # `1 << bits` is an i32 in the reference
            one -= 1 << 32
        scale = F(one)
        for j in range(ALPHABET_SIZE):
            lv = as_u32(F(laplace(unpack_signed(j), 0.0, width) * scale))
            if lv == 0 and self.freqs[j] == 0 and j in self.off_distribution_values:
                self.freqs[j] = 1
            elif self.freqs[j] != 0 and lv == 0:
                self.freqs[j] = 1
                self.off_distribution_values.append(j)
            else:
                self.freqs[j] = lv

    def normalize_freqs(self, target_total: int):
        cum = self.get_cdf()
        cur_total = (cum[-1] + self.freqs[-1]) & M32
        for i in range(1, ALPHABET_SIZE):
            cum[i] = (target_total * cum[i]) // cur_total
        for i in range(ALPHABET_SIZE - 1):
            if self.freqs[i] != 0 and cum[i + 1] == cum[i]:
                best_freq, best_steal = M32, None
                for j in range(ALPHABET_SIZE - 1):
                    freq = cum[j + 1] - cum[j]
                    if 1 < freq < best_freq:
                        best_freq, best_steal = freq, j
                if best_steal is None:
                    continue
                if best_steal < i:
                    for j in range(best_steal + 1, i + 1):
                        cum[j] -= 1
                else:
                    for j in range(i + 1, best_steal + 1):
                        cum[j] += 1
        for i in range(ALPHABET_SIZE - 1):
            self.freqs[i] = cum[i + 1] - cum[i]
        self.freqs[-1] = (cum[-1] - target_total) & M32
        return cum

    def finalize_context(self, normalize: bool, bucket: int):
        if self.max_freq_bits < 8:
            self.max_freq_bits = 8
        self.fill_with_laplace(bucket)
        self.cdf = self.normalize_freqs(1 << self.max_freq_bits) if normalize else self.get_cdf()
        self.max_freq_bits = trailing_zeros(prev_power_two(sum(self.freqs) & M32))


def context_from_counts(counts, bucket: int) -> AnsContext:  # prediction.rs:302-304
    c = AnsContext()
    c.freqs = [int(x) for x in counts]
    total = sum(c.freqs) & M32
    c.max_freq_bits = trailing_zeros(prev_power_two(total)) if total else 0
    c.finalize_context(True, bucket)
    return c


RANS64_L = 1 << 31


def rans_encode(symbols, buckets, contexts) -> bytes:
    """entropy_coding.rs:332-336 with the rans64 algorithm: put in reverse order, flush states 0..9, data()."""
    state = [RANS64_L] * CONTEXT_AMOUNT
    words = []  # emission order; the stream is the reverse
    for s, b in zip(reversed([int(x) for x in symbols]), reversed([int(x) for x in buckets])):
        ctx = contexts[b]
        start, freq, bits = ctx.cdf[s], ctx.freqs[s], ctx.max_freq_bits
        x = state[b]
        x_max = ((RANS64_L >> bits) << 32) * freq
        if x >= x_max:
            words.append(x & M32)
            x >>= 32
        state[b] = ((x // freq) << bits) + (x % freq) + start
    for x in state:
        words.append(x >> 32)
        words.append(x & M32)
    return b"".join(struct.pack("<I", w) for w in reversed(words))


def serialize(height: int, width: int, colorspace: int, channels) -> bytes:
    """serialize.rs:48-117.  channels: list of (value_params [3][6], width_params [3][6], contexts, data)."""
    out = bytearray(b"frif")
    out += struct.pack("<II", height, width)
    out += struct.pack("<I", (colorspace << 30) | (1 << 28))
    for vp, wp, contexts, data in channels:
        out += b"\xFF\xBB"
        for row in vp:
            for x in row:
                out += struct.pack("<f", float(x))
        for row in wp:
            for x in row:
                out += struct.pack("<f", float(x))
        for ctx in contexts:
            out += b"\xFF\xB2" + struct.pack("<I", ctx.max_freq_bits) + struct.pack("<Q", len(ctx.off_distribution_values))
            for v in ctx.off_distribution_values:
                out += struct.pack("<H", v)
        out += b"\xFF\xB4" + struct.pack("<Q", len(data)) + data + b"\xFF\xB8"
    out += b"\xFF\xDF"
    return bytes(out)


def encode(height, width, colorspace, value_params, width_params, bucket, sym, hist) -> bytes:
    """Whole host tail of the encoder for one frame: bucket / sym [C][n], hist [C][10][1024] -> container bytes."""
    channels = []
    for ch in range(len(bucket)):
        contexts = [context_from_counts(hist[ch][b], b) for b in range(CONTEXT_AMOUNT)]
        data = rans_encode(sym[ch], bucket[ch], contexts)
        channels.append((value_params[ch], width_params[ch], contexts, data))
    return serialize(height, width, colorspace, channels)

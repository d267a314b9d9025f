// fri_kat.rs — known-answer harness for the REAL reference (pagmerek/frave, libfri).
//
// TEST INFRASTRUCTURE.  NEVER RUN IN THIS REPOSITORY'S ENVIRONMENT: the image has no cargo / rustc and no
// network.  oracle/ref_harness/run.sh copies the reference tree to a scratch directory, drops this file into
// crates/libfri/src/stages/ (the stages are private modules, lib.rs:6-10, so the harness has to live inside
// the crate), registers it in stages/mod.rs and runs `cargo test -p libfri fri_kat -- --nocapture`.
//
// For every case it prints one JSON line with the digests SURVEY.md §8(c) defines, computed from the
// reference's own functions:
//   * WaveletImage::from_raster (wavelet_transform.rs:405-432) + quantization::encode (quantization.rs:7-25):
//       "sha256"       tiles sorted by (centre.im, centre.re): <i32 re><i32 im>, then per channel, per
//                      coefficient index 0..511: Some(v) -> <i32 v LE> 01, None -> FF FF FF 7F 00;
//       "retained", "some", "sum"
//   * quantization::decode + RasterImage::from_wavelet (:308-322, :358-381):
//       "recon_sha256" SHA-256 of the reconstructed HWC bytes
//   * sort_lattice (:657-705):
//       "order_sha256" for level 0..8, for every position of sorted_lattice[level]: <i32 re><i32 im>
// tests/test_oracle.py::test_reference_digests compares these lines (oracle/_ref/kat_digests.jsonl, if present)
// with the same digests computed by oracle/fri_oracle.c and by the plan's emission order.
use crate::images::{ImageMetadata, RasterImage};
use crate::stages::quantization;
use crate::stages::wavelet_transform::WaveletImage;

const K: [u32; 64] = [
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2,
];

fn sha256(data: &[u8]) -> String {
    let mut h: [u32; 8] = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19];
    let mut msg = data.to_vec();
    let bits = (data.len() as u64) * 8;
    msg.push(0x80);
    while msg.len() % 64 != 56 {
        msg.push(0);
    }
    msg.extend_from_slice(&bits.to_be_bytes());
    for block in msg.chunks(64) {
        let mut w = [0u32; 64];
        for i in 0..16 {
            w[i] = u32::from_be_bytes([block[4 * i], block[4 * i + 1], block[4 * i + 2], block[4 * i + 3]]);
        }
        for i in 16..64 {
            let s0 = w[i - 15].rotate_right(7) ^ w[i - 15].rotate_right(18) ^ (w[i - 15] >> 3);
            let s1 = w[i - 2].rotate_right(17) ^ w[i - 2].rotate_right(19) ^ (w[i - 2] >> 10);
            w[i] = w[i - 16].wrapping_add(s0).wrapping_add(w[i - 7]).wrapping_add(s1);
        }
        let mut v = h;
        for i in 0..64 {
            let s1 = v[4].rotate_right(6) ^ v[4].rotate_right(11) ^ v[4].rotate_right(25);
            let ch = (v[4] & v[5]) ^ (!v[4] & v[6]);
            let t1 = v[7].wrapping_add(s1).wrapping_add(ch).wrapping_add(K[i]).wrapping_add(w[i]);
            let s0 = v[0].rotate_right(2) ^ v[0].rotate_right(13) ^ v[0].rotate_right(22);
            let maj = (v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]);
            let t2 = s0.wrapping_add(maj);
            v = [t1.wrapping_add(t2), v[0], v[1], v[2], v[3].wrapping_add(t1), v[4], v[5], v[6]];
        }
        for i in 0..8 {
            h[i] = h[i].wrapping_add(v[i]);
        }
    }
    h.iter().map(|x| format!("{:08x}", x)).collect()
}

/// pix(x, y, ch) = (7x + 13y + (x*y mod 11) + 29ch) mod 256, HWC u8.
fn survey_image(w: u32, h: u32, c: u32) -> Vec<u8> {
    let mut v = Vec::with_capacity((w * h * c) as usize);
    for y in 0..h {
        for x in 0..w {
            for ch in 0..c {
                v.push(((7 * x + 13 * y + (x * y) % 11 + 29 * ch) % 256) as u8);
            }
        }
    }
    v
}

fn digest_case(w: u32, h: u32) {
    let c = 3u32; // ImageMetadata::new is RGB; a Luma image panics in the reference (retain over 3 slots, :415-416)
    let image = RasterImage { metadata: ImageMetadata::new(h, w), data: survey_image(w, h, c) };
    let wavelet = quantization::encode(WaveletImage::from_raster(image)).unwrap();
    let mut centers: Vec<_> = wavelet.fractal_lattice.keys().cloned().collect();
    centers.sort_by(|a, b| (a.im, a.re).cmp(&(b.im, b.re)));
    let (mut bytes, mut some, mut sum) = (Vec::<u8>::new(), 0u64, 0i64);
    for ctr in &centers {
        bytes.extend_from_slice(&ctr.re.to_le_bytes());
        bytes.extend_from_slice(&ctr.im.to_le_bytes());
        let f = &wavelet.fractal_lattice[ctr];
        for ch in 0..c as usize {
            for coef in f.coefficients[ch].iter() {
                match coef {
                    Some(v) => {
                        bytes.extend_from_slice(&v.to_le_bytes());
                        bytes.push(1);
                        some += 1;
                        sum += *v as i64;
                    }
                    None => bytes.extend_from_slice(&[0xff, 0xff, 0xff, 0x7f, 0x00]),
                }
            }
        }
    }
    let coef_sha = sha256(&bytes);
    let mut order = Vec::<u8>::new();
    for level in wavelet.sorted_lattice.iter() {
        for p in level.iter() {
            order.extend_from_slice(&p.re.to_le_bytes());
            order.extend_from_slice(&p.im.to_le_bytes());
        }
    }
    let order_sha = sha256(&order);
    let retained = centers.len();
    let recon = RasterImage::from_wavelet(quantization::decode(wavelet).unwrap());
    println!(
        "FRI_KAT {{\"w\": {}, \"h\": {}, \"c\": {}, \"retained\": {}, \"some\": {}, \"sum\": {}, \"sha256\": \"{}\", \"recon_sha256\": \"{}\", \"order_sha256\": \"{}\"}}",
        w, h, c, retained, some, sum, coef_sha, sha256(&recon.data), order_sha
    );
}

#[test]
fn fri_kat() {
    // keep in sync with oracle/ref_harness/cases.json
    for (w, h) in [(64u32, 48u32), (100, 37), (10, 10), (131, 77), (480, 270), (512, 512), (1920, 1080)] {
        digest_case(w, h);
    }
}

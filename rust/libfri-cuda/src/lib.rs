//! Wrapper over libfri_cuda's C ABI (include/fri_cuda.h).
//!
//! NEVER COMPILED IN THIS REPOSITORY'S ENVIRONMENT (no cargo/rustc in the image).  The `extern "C"` block
//! below is checked mechanically against the header by tests/test_host.py::test_rust_extern_block_matches_header
//! (every prototype bound, same argument count); the wrappers are meant to be reviewed by inspection.
//!
//! Soundness rules the wrappers enforce:
//!   * every slice handed to the C side is length-checked against the plan's frame / coefficient sizes
//!     BEFORE the call (a short slice is an `Err`, never an out-of-bounds access);
//!   * host buffers of the fast path are `PinnedBuf`s (fri_host_alloc) — the path the measured end-to-end
//!     numbers use; plain slices are accepted by the synchronous calls only;
//!   * asynchronous mode is `unsafe`: the caller promises that the pinned buffers outlive `sync()`.
use std::ffi::{c_char, c_int, c_void, CStr};
use std::marker::PhantomData;

#[repr(C)]
pub struct FriPlan {
    _private: [u8; 0],
}

pub const FRI_DEQUANT_DIVIDE: c_int = 0; // quantization.rs:37 — the reference divides again
pub const FRI_DEQUANT_MULTIPLY: c_int = 1;
pub const FRI_BASE_DEPTH: u32 = 9; // BASE_FRAC_DEPTH, wavelet_transform.rs:39

extern "C" {
    fn fri_version() -> *const c_char;
    fn fri_last_error() -> *const c_char;
    fn fri_device_count() -> c_int;
    fn fri_plan_create(out: *mut *mut FriPlan, device: c_int, width: u32, height: u32, channels: u32, depth: u32, sample_bytes: u32) -> c_int;
    fn fri_plan_destroy(plan: *mut FriPlan);
    fn fri_plan_num_tiles(plan: *const FriPlan) -> u32;
    fn fri_plan_num_built(plan: *const FriPlan) -> u32;
    fn fri_plan_num_full_tiles(plan: *const FriPlan) -> u32;
    fn fri_plan_coefs_per_frame(plan: *const FriPlan) -> u64;
    fn fri_plan_pixels_covered(plan: *const FriPlan) -> u64;
    fn fri_plan_launch_info(plan: *const FriPlan, info: *mut i32) -> c_int;
    fn fri_plan_centers(plan: *const FriPlan, centers: *mut i32) -> c_int;
    fn fri_plan_masks(plan: *const FriPlan, masks: *mut u32) -> c_int;
    fn fri_encode_tq_device(plan: *const FriPlan, d_pixels: *const c_void, n_frames: u32, q: *const i32, d_coefs: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_decode_tq_device(plan: *const FriPlan, d_coefs: *const i32, n_frames: u32, q: *const i32, dequant_mode: c_int, d_pixels: *mut c_void, stream: *mut c_void) -> c_int;
    fn fri_encode_tq_device16(plan: *const FriPlan, d_pixels: *const c_void, n_frames: u32, q: *const i32, d_coefs: *mut i16, stream: *mut c_void) -> c_int;
    fn fri_decode_tq_device16(plan: *const FriPlan, d_coefs: *const i16, n_frames: u32, q: *const i32, dequant_mode: c_int, d_pixels: *mut c_void, stream: *mut c_void) -> c_int;
    fn fri_encode_tq(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, coefs: *mut i32) -> c_int;
    fn fri_decode_tq(plan: *mut FriPlan, coefs: *const i32, n_frames: u32, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_encode_tq16(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, coefs: *mut i16) -> c_int;
    fn fri_decode_tq16(plan: *mut FriPlan, coefs: *const i16, n_frames: u32, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_plan_emission_count(plan: *mut FriPlan) -> u64;
    fn fri_plan_emission_packed_bytes(plan: *mut FriPlan) -> u64;
    fn fri_plan_emission_packed_size(plan: *mut FriPlan, bits: c_int) -> u64;
    fn fri_emit_device_packed(plan: *mut FriPlan, d_coefs: *const i32, n_frames: u32, bits: c_int, d_out: *mut u8, stream: *mut c_void) -> c_int;
    fn fri_encode_tq_emit_packed(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, bits: c_int, out: *mut u8) -> c_int;
    fn fri_unemit_device_packed(plan: *mut FriPlan, d_streams: *const u8, n_frames: u32, bits: c_int, d_coefs: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_decode_tq_emit_packed(plan: *mut FriPlan, streams: *const u8, n_frames: u32, bits: c_int, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_plan_emission_order(plan: *mut FriPlan, order: *mut u32) -> c_int;
    fn fri_emit_device(plan: *mut FriPlan, d_coefs: *const i32, n_frames: u32, d_out: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_emit_device16(plan: *mut FriPlan, d_coefs: *const i32, n_frames: u32, d_out: *mut i16, stream: *mut c_void) -> c_int;
    fn fri_emit_device10(plan: *mut FriPlan, d_coefs: *const i32, n_frames: u32, d_out: *mut u8, stream: *mut c_void) -> c_int;
    fn fri_encode_tq_emit(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, out: *mut i32) -> c_int;
    fn fri_encode_tq_emit16(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, out: *mut i16) -> c_int;
    fn fri_encode_tq_emit10(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, out: *mut u8) -> c_int;
    fn fri_unemit_device(plan: *mut FriPlan, d_streams: *const i32, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_unemit_device16(plan: *mut FriPlan, d_streams: *const i16, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_unemit_device10(plan: *mut FriPlan, d_streams: *const u8, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> c_int;
    fn fri_decode_tq_emit(plan: *mut FriPlan, streams: *const i32, n_frames: u32, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_decode_tq_emit16(plan: *mut FriPlan, streams: *const i16, n_frames: u32, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_decode_tq_emit10(plan: *mut FriPlan, streams: *const u8, n_frames: u32, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_predict_device(plan: *mut FriPlan, d_coefs: *const i32, n_frames: u32, value_params: *const f32, width_params: *const f32, d_bucket: *mut u8, d_pred: *mut i32, d_sym: *mut u16, d_hist: *mut u32, d_overflow: *mut u32, stream: *mut c_void) -> c_int;
    fn fri_fit_parameters(plan: *mut FriPlan, coefs: *const i32, value_params: *mut f32, width_params: *mut f32) -> c_int;
    fn fri_fit_device(plan: *mut FriPlan, d_coefs: *const i32, value_params: *mut f32, width_params: *mut f32, stream: *mut c_void) -> c_int;
    fn fri_plan_part(plan: *const FriPlan, part: u32, n_parts: u32, group_begin: *mut u32, group_end: *mut u32, tile_begin: *mut u32, tile_end: *mut u32, row_begin: *mut u32, row_end: *mut u32) -> c_int;
    fn fri_encode_tq_device_part(plan: *const FriPlan, d_pixel_rows: *const c_void, q: *const i32, d_coef_tiles: *mut i32, part: u32, n_parts: u32, stream: *mut c_void) -> c_int;
    fn fri_decode_tq_device_part(plan: *const FriPlan, d_coef_tiles: *const i32, q: *const i32, dequant_mode: c_int, d_pixel_rows: *mut c_void, part: u32, n_parts: u32, stream: *mut c_void) -> c_int;
    fn fri_plan_groups_in_rows(plan: *const FriPlan, group_begin: u32, group_end: u32, row_begin: u32, row_end: u32, first: *mut u32, last: *mut u32, span_begin: *mut u32, span_end: *mut u32) -> c_int;
    fn fri_decode_tq_device_groups(plan: *const FriPlan, d_coef_tiles: *const i32, tile_first: u32, q: *const i32, dequant_mode: c_int, d_pixel_rows: *mut c_void, row_first: i32, group_begin: u32, group_end: u32, stream: *mut c_void) -> c_int;
    fn fri_predict_host(plan: *mut FriPlan, coefs: *const i32, value_params: *const f32, width_params: *const f32, bucket: *mut u8, pred: *mut i32, sym: *mut u16, hist: *mut u32, overflow: *mut u32) -> c_int;
    fn fri_frv_pack(plan: *mut FriPlan, colorspace: c_int, value_params: *const f32, width_params: *const f32, bucket: *const u8, sym: *const u16, hist: *const u32, out: *mut *mut u8, out_len: *mut usize) -> c_int;
    fn fri_frv_unpack(plan: *mut FriPlan, bytes: *const u8, len: usize, coefs: *mut i32) -> c_int;
    fn fri_frv_info(bytes: *const u8, len: usize, width: *mut u32, height: *mut u32, channels: *mut u32) -> c_int;
    fn fri_frv_encode(plan: *mut FriPlan, pixels: *const c_void, q: *const i32, colorspace: c_int, out: *mut *mut u8, out_len: *mut usize) -> c_int;
    fn fri_frv_decode(plan: *mut FriPlan, bytes: *const u8, len: usize, q: *const i32, dequant_mode: c_int, pixels: *mut c_void) -> c_int;
    fn fri_frv_free(bytes: *mut u8);
    fn fri_plan_set_bands(plan: *mut FriPlan, bands: c_int) -> c_int;
    fn fri_plan_set_async(plan: *mut FriPlan, on: c_int) -> c_int;
    fn fri_plan_set_independent_calls(plan: *mut FriPlan, on: c_int) -> c_int;
    fn fri_plan_sync(plan: *mut FriPlan) -> c_int;
    fn fri_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    fn fri_host_free(p: *mut c_void);
    fn fri_plan_last_launches(plan: *const FriPlan) -> u32;
    fn fri_quant_divide(value: i32, q: i32) -> i32;
    fn fri_quant_divide_magic(value: i32, q: i32) -> i32;
    fn fri_quant_divide_small(value: i32, q: i32) -> i32;
}

fn check(rc: c_int) -> Result<(), String> {
    if rc == 0 {
        Ok(())
    } else {
        // Err(String) is what every libfri stage returns (encoder.rs:19-46)
        Err(unsafe { CStr::from_ptr(fri_last_error()) }.to_string_lossy().into_owned())
    }
}

fn want_len(what: &str, got: usize, want: usize) -> Result<(), String> {
    if got == want {
        Ok(())
    } else {
        Err(format!("{what}: slice of {got} elements, the plan needs exactly {want}"))
    }
}

pub fn version() -> String { unsafe { CStr::from_ptr(fri_version()) }.to_string_lossy().into_owned() }
pub fn device_count() -> i32 { unsafe { fri_device_count() } }
/// (width, height, channels) of a `frif` container.
pub fn frv_info(bytes: &[u8]) -> Result<(u32, u32, u32), String> {
    let (mut w, mut h, mut c) = (0u32, 0u32, 0u32);
    check(unsafe { fri_frv_info(bytes.as_ptr(), bytes.len(), &mut w, &mut h, &mut c) })?;
    Ok((w, h, c))
}
/// value / q exactly as the kernels compute it (truncation toward zero, quantization.rs:19).
pub fn quant_divide(value: i32, q: i32) -> i32 { unsafe { fri_quant_divide(value, q) } }
pub fn quant_divide_magic(value: i32, q: i32) -> i32 { unsafe { fri_quant_divide_magic(value, q) } }
pub fn quant_divide_small(value: i32, q: i32) -> i32 { unsafe { fri_quant_divide_small(value, q) } }

/// Page-locked host memory (fri_host_alloc): the buffers of the fast path — asynchronous copies at the
/// link's full rate.  Owns its allocation; `T` is a plain integer type.
pub struct PinnedBuf<T: Copy> {
    ptr: *mut T,
    len: usize,
    _own: PhantomData<T>,
}

impl<T: Copy> PinnedBuf<T> {
    pub fn new(len: usize) -> Result<Self, String> {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { fri_host_alloc(&mut p, len * std::mem::size_of::<T>()) })?;
        unsafe { std::ptr::write_bytes(p as *mut u8, 0, len * std::mem::size_of::<T>()) };
        Ok(PinnedBuf { ptr: p as *mut T, len, _own: PhantomData })
    }
    pub fn len(&self) -> usize { self.len }
    pub fn is_empty(&self) -> bool { self.len == 0 }
    pub fn as_slice(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.ptr, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.ptr, self.len) } }
}

impl<T: Copy> Drop for PinnedBuf<T> {
    fn drop(&mut self) { unsafe { fri_host_free(self.ptr as *mut c_void) } }
}
unsafe impl<T: Copy + Send> Send for PinnedBuf<T> {}

/// Lattice + launch geometry for one image size (replaces fractal_divide + Fractal::new + retain,
/// wavelet_transform.rs:450-484, :42-69, :415-416, and from_metadata :392-403 on decode).
pub struct Plan {
    raw: *mut FriPlan,
    frame_bytes: usize,     // width * height * channels (8-bit samples)
    coefs_per_frame: usize, // n_tiles * channels * 512
    channels: usize,
}

impl Plan {
    pub fn new(device: i32, width: u32, height: u32, channels: u32) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        check(unsafe { fri_plan_create(&mut p, device, width, height, channels, FRI_BASE_DEPTH, 1) })?;
        let coefs = unsafe { fri_plan_coefs_per_frame(p) } as usize;
        Ok(Plan { raw: p, frame_bytes: width as usize * height as usize * channels as usize, coefs_per_frame: coefs, channels: channels as usize })
    }
    pub fn num_tiles(&self) -> usize { unsafe { fri_plan_num_tiles(self.raw) as usize } }
    pub fn num_built(&self) -> usize { unsafe { fri_plan_num_built(self.raw) as usize } }
    pub fn num_full_tiles(&self) -> usize { unsafe { fri_plan_num_full_tiles(self.raw) as usize } }
    pub fn coefs_per_frame(&self) -> usize { self.coefs_per_frame }
    pub fn frame_bytes(&self) -> usize { self.frame_bytes }
    pub fn pixels_covered(&self) -> u64 { unsafe { fri_plan_pixels_covered(self.raw) } }
    pub fn last_launches(&self) -> u32 { unsafe { fri_plan_last_launches(self.raw) } }
    pub fn launch_info(&self) -> Result<[i32; 16], String> {
        let mut v = [0i32; 16];
        check(unsafe { fri_plan_launch_info(self.raw, v.as_mut_ptr()) })?;
        Ok(v)
    }
    /// (re, im) of every retained tile, in the order of the coefficient blocks.
    pub fn centers(&self) -> Result<Vec<[i32; 2]>, String> {
        let mut v = vec![[0i32; 2]; self.num_tiles()];
        check(unsafe { fri_plan_centers(self.raw, v.as_mut_ptr() as *mut i32) })?;
        Ok(v)
    }
    /// 512-bit Some/None mask per tile (bit i of word i/32 set <=> coefficient i is Some).
    pub fn masks(&self) -> Result<Vec<[u32; 16]>, String> {
        let mut v = vec![[0u32; 16]; self.num_tiles()];
        check(unsafe { fri_plan_masks(self.raw, v.as_mut_ptr() as *mut u32) })?;
        Ok(v)
    }

    // ---- the two stage calls, synchronous, any host memory (pageable slices take the driver's staging path)
    /// wavelet_transform::encode + quantization::encode (encoder.rs:26-33), fused.
    pub fn encode_tq(&mut self, pixels: &[u8], q: &[i32; 32], coefs: &mut [i32]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        check(unsafe { fri_encode_tq(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), coefs.as_mut_ptr()) })
    }
    /// quantization::decode + wavelet_transform::decode (decoder.rs:27-34), fused.
    pub fn decode_tq(&mut self, coefs: &[i32], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq(self.raw, coefs.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    /// The same with int16 coefficients on the host side (every coefficient of an 8-bit image fits:
    /// |residue| <= 255, wavelet_transform.rs:211-218); the glue widens while applying the mask.
    pub fn encode_tq16(&mut self, pixels: &[u8], q: &[i32; 32], coefs: &mut [i16]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        check(unsafe { fri_encode_tq16(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), coefs.as_mut_ptr()) })
    }
    pub fn decode_tq16(&mut self, coefs: &[i16], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq16(self.raw, coefs.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }

    // ---- the fast path the end-to-end numbers measure: pinned buffers allocated once per image size
    pub fn pinned_frame(&self) -> Result<PinnedBuf<u8>, String> { PinnedBuf::new(self.frame_bytes) }
    pub fn pinned_coefs16(&self) -> Result<PinnedBuf<i16>, String> { PinnedBuf::new(self.coefs_per_frame) }
    pub fn encode_tq16_pinned(&mut self, pixels: &PinnedBuf<u8>, q: &[i32; 32], coefs: &mut PinnedBuf<i16>) -> Result<(), String> {
        let (p, c) = (pixels.as_slice().as_ptr(), coefs.as_mut_slice().as_mut_ptr());
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        check(unsafe { fri_encode_tq16(self.raw, p as *const c_void, 1, q.as_ptr(), c) })
    }
    pub fn decode_tq16_pinned(&mut self, coefs: &PinnedBuf<i16>, q: &[i32; 32], pixels: &mut PinnedBuf<u8>) -> Result<(), String> {
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq16(self.raw, coefs.as_slice().as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_slice().as_mut_ptr() as *mut c_void) })
    }

    // ---- emission order (entropy_coding.rs:283-329 over sort_lattice): flat streams for the host coder
    pub fn emission_count(&mut self) -> usize { unsafe { fri_plan_emission_count(self.raw) as usize } }
    pub fn emission_packed_bytes(&mut self) -> usize { unsafe { fri_plan_emission_packed_bytes(self.raw) as usize } }
    pub fn emission_order(&mut self) -> Result<Vec<u32>, String> {
        let mut v = vec![0u32; self.num_tiles() * 512];
        check(unsafe { fri_plan_emission_order(self.raw, v.as_mut_ptr()) })?;
        Ok(v)
    }
    /// pixels -> [C][emission_count] quantized Some coefficients in the order the entropy coder consumes them.
    pub fn encode_tq_emit(&mut self, pixels: &[u8], q: &[i32; 32], out: &mut [i32]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        let n = self.channels * self.emission_count();
        want_len("streams", out.len(), n)?;
        check(unsafe { fri_encode_tq_emit(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), out.as_mut_ptr()) })
    }
    pub fn encode_tq_emit16(&mut self, pixels: &[u8], q: &[i32; 32], out: &mut [i16]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        let n = self.channels * self.emission_count();
        want_len("streams", out.len(), n)?;
        check(unsafe { fri_encode_tq_emit16(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), out.as_mut_ptr()) })
    }
    /// The same streams as 10-bit zig-zag symbols (utils.rs:34-40), 64 symbols per 80 bytes: [C][emission_packed_bytes].
    pub fn encode_tq_emit10(&mut self, pixels: &[u8], q: &[i32; 32], out: &mut [u8]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        let n = self.channels * self.emission_packed_bytes();
        want_len("packed streams", out.len(), n)?;
        check(unsafe { fri_encode_tq_emit10(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), out.as_mut_ptr()) })
    }
    /// The packed streams at `bits` = 9 or 10 bits per symbol: [C][emission_packed_size(bits)] bytes.
    pub fn emission_packed_size(&mut self, bits: i32) -> usize { unsafe { fri_plan_emission_packed_size(self.raw, bits) as usize } }
    pub fn encode_tq_emit_packed(&mut self, pixels: &[u8], q: &[i32; 32], bits: i32, out: &mut [u8]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        let n = self.channels * self.emission_packed_size(bits);
        want_len("packed streams", out.len(), n)?;
        check(unsafe { fri_encode_tq_emit_packed(self.raw, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), bits, out.as_mut_ptr()) })
    }
    pub fn decode_tq_emit_packed(&mut self, packed: &[u8], bits: i32, q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        let n = self.channels * self.emission_packed_size(bits);
        want_len("packed streams", packed.len(), n)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq_emit_packed(self.raw, packed.as_ptr(), 1, bits, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    /// # Safety
    /// Device pointers; packed output [n_frames][C][emission_packed_size(bits)], 16-byte aligned.
    pub unsafe fn emit_device_packed(&mut self, d_coefs: *const i32, n_frames: u32, bits: i32, d_out: *mut u8, stream: *mut c_void) -> Result<(), String> {
        check(fri_emit_device_packed(self.raw, d_coefs, n_frames, bits, d_out, stream))
    }
    /// # Safety
    /// Device pointers; see `emit_device_packed`.
    pub unsafe fn unemit_device_packed(&mut self, d_streams: *const u8, n_frames: u32, bits: i32, d_coefs: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_unemit_device_packed(self.raw, d_streams, n_frames, bits, d_coefs, stream))
    }
    pub fn decode_tq_emit(&mut self, streams: &[i32], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        let n = self.channels * self.emission_count();
        want_len("streams", streams.len(), n)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq_emit(self.raw, streams.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    pub fn decode_tq_emit16(&mut self, streams: &[i16], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        let n = self.channels * self.emission_count();
        want_len("streams", streams.len(), n)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq_emit16(self.raw, streams.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    pub fn decode_tq_emit10(&mut self, packed: &[u8], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        let n = self.channels * self.emission_packed_bytes();
        want_len("packed streams", packed.len(), n)?;
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_decode_tq_emit10(self.raw, packed.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }

    // ---- the whole codec (this library's own host entropy coder + container; parity with the reference's bytes unpinned)
    /// FRIEncoder::encode (encoder.rs:87-109): pixels -> `frif` container bytes.
    pub fn frv_encode(&mut self, pixels: &[u8], q: &[i32; 32]) -> Result<Vec<u8>, String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        let (mut out, mut len) = (std::ptr::null_mut::<u8>(), 0usize);
        check(unsafe { fri_frv_encode(self.raw, pixels.as_ptr() as *const c_void, q.as_ptr(), 0, &mut out, &mut len) })?;
        let v = unsafe { std::slice::from_raw_parts(out, len) }.to_vec();
        unsafe { fri_frv_free(out) };
        Ok(v)
    }
    /// FRIDecoder::decode (decoder.rs:48-59): container bytes -> pixels.
    pub fn frv_decode(&mut self, bytes: &[u8], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        want_len("pixels", pixels.len(), self.frame_bytes)?;
        check(unsafe { fri_frv_decode(self.raw, bytes.as_ptr(), bytes.len(), q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    /// Entropy decoding only: container bytes -> quantized dense blocks (for a caller that keeps its own lattice).
    pub fn frv_unpack(&mut self, bytes: &[u8], coefs: &mut [i32]) -> Result<(), String> {
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        check(unsafe { fri_frv_unpack(self.raw, bytes.as_ptr(), bytes.len(), coefs.as_mut_ptr()) })
    }
    /// Predictor parameter fit on the host (context_modeling.rs:204-214): ([C][3][6] value, [C][3][6] width).
    pub fn fit_parameters(&mut self, coefs: &[i32]) -> Result<(Vec<[[f32; 6]; 3]>, Vec<[[f32; 6]; 3]>), String> {
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        let (mut v, mut w) = (vec![[[0f32; 6]; 3]; self.channels], vec![[[0f32; 6]; 3]; self.channels]);
        check(unsafe { fri_fit_parameters(self.raw, coefs.as_ptr(), v.as_mut_ptr() as *mut f32, w.as_mut_ptr() as *mut f32) })?;
        Ok((v, w))
    }
    /// The same fit for one frame resident on the device (sums by a kernel, solve on the host; synchronizes `stream`).
    /// # Safety
    /// `d_coefs` must be a device pointer to one frame of dense blocks on this plan's device.
    pub unsafe fn fit_device(&mut self, d_coefs: *const i32, stream: *mut c_void) -> Result<(Vec<[[f32; 6]; 3]>, Vec<[[f32; 6]; 3]>), String> {
        let (mut v, mut w) = (vec![[[0f32; 6]; 3]; self.channels], vec![[[0f32; 6]; 3]; self.channels]);
        check(fri_fit_device(self.raw, d_coefs, v.as_mut_ptr() as *mut f32, w.as_mut_ptr() as *mut f32, stream))?;
        Ok((v, w))
    }
    /// Part `part` of `n_parts` of one image split over several GPUs: (groups, tiles, pixel rows) as half-open ranges.
    pub fn part(&self, part: u32, n_parts: u32) -> Result<[(u32, u32); 3], String> {
        let mut v = [0u32; 6];
        let p = v.as_mut_ptr();
        check(unsafe { fri_plan_part(self.raw, part, n_parts, p, p.add(1), p.add(2), p.add(3), p.add(4), p.add(5)) })?;
        Ok([(v[0], v[1]), (v[2], v[3]), (v[4], v[5])])
    }
    /// # Safety
    /// `d_pixel_rows` / `d_coef_tiles` must be device buffers on this plan's device covering the part's rows / tiles.
    pub unsafe fn encode_device_part(&self, d_pixel_rows: *const c_void, q: &[i32; 32], d_coef_tiles: *mut i32, part: u32, n_parts: u32,
                                     stream: *mut c_void) -> Result<(), String> {
        check(fri_encode_tq_device_part(self.raw, d_pixel_rows, q.as_ptr(), d_coef_tiles, part, n_parts, stream))
    }
    /// # Safety
    /// As `encode_device_part`; writes only the pixels the part's tiles own.
    pub unsafe fn decode_device_part(&self, d_coef_tiles: *const i32, q: &[i32; 32], d_pixel_rows: *mut c_void, part: u32, n_parts: u32,
                                     stream: *mut c_void) -> Result<(), String> {
        check(fri_decode_tq_device_part(self.raw, d_coef_tiles, q.as_ptr(), FRI_DEQUANT_DIVIDE, d_pixel_rows, part, n_parts, stream))
    }
    /// Groups of `[group_begin, group_end)` touching pixel rows `[row_begin, row_end)`: ((first, last), (span_begin, span_end)).
    pub fn groups_in_rows(&self, group_begin: u32, group_end: u32, row_begin: u32, row_end: u32) -> Result<[(u32, u32); 2], String> {
        let mut v = [0u32; 4];
        let p = v.as_mut_ptr();
        check(unsafe { fri_plan_groups_in_rows(self.raw, group_begin, group_end, row_begin, row_end, p, p.add(1), p.add(2), p.add(3)) })?;
        Ok([(v[0], v[1]), (v[2], v[3])])
    }
    /// # Safety
    /// `d_coef_tiles` points at tile `tile_first`'s block, `d_pixel_rows` at frame row `row_first`; both must cover what the
    /// groups touch (the target may be a peer-mapped band of another GPU).
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn decode_device_groups(&self, d_coef_tiles: *const i32, tile_first: u32, q: &[i32; 32], d_pixel_rows: *mut c_void, row_first: i32,
                                       group_begin: u32, group_end: u32, stream: *mut c_void) -> Result<(), String> {
        check(fri_decode_tq_device_groups(self.raw, d_coef_tiles, tile_first, q.as_ptr(), FRI_DEQUANT_DIVIDE, d_pixel_rows, row_first, group_begin, group_end, stream))
    }
    /// Host predictor (what the serial entropy decoder evaluates): see include/fri_cuda.h for the array shapes.
    #[allow(clippy::too_many_arguments)]
    pub fn predict_host(&mut self, coefs: &[i32], value_params: &[[[f32; 6]; 3]], width_params: &[[[f32; 6]; 3]], bucket: &mut [u8], pred: &mut [i32],
                        sym: &mut [u16], hist: &mut [u32]) -> Result<u32, String> {
        let n = self.channels * self.emission_count();
        want_len("coefs", coefs.len(), self.coefs_per_frame)?;
        want_len("value_params", value_params.len(), self.channels)?;
        want_len("width_params", width_params.len(), self.channels)?;
        want_len("bucket", bucket.len(), n)?;
        want_len("pred", pred.len(), n)?;
        want_len("sym", sym.len(), n)?;
        want_len("hist", hist.len(), self.channels * 10 * 1024)?;
        let mut over = 0u32;
        check(unsafe { fri_predict_host(self.raw, coefs.as_ptr(), value_params.as_ptr() as *const f32, width_params.as_ptr() as *const f32,
                                        bucket.as_mut_ptr(), pred.as_mut_ptr(), sym.as_mut_ptr(), hist.as_mut_ptr(), &mut over) })?;
        Ok(over)
    }
    /// Symbols + buckets + histograms -> container bytes (rANS + serialize on the host).
    pub fn frv_pack(&mut self, value_params: &[[[f32; 6]; 3]], width_params: &[[[f32; 6]; 3]], bucket: &[u8], sym: &[u16], hist: &[u32]) -> Result<Vec<u8>, String> {
        let n = self.channels * self.emission_count();
        want_len("value_params", value_params.len(), self.channels)?;
        want_len("width_params", width_params.len(), self.channels)?;
        want_len("bucket", bucket.len(), n)?;
        want_len("sym", sym.len(), n)?;
        want_len("hist", hist.len(), self.channels * 10 * 1024)?;
        let (mut out, mut len) = (std::ptr::null_mut::<u8>(), 0usize);
        check(unsafe { fri_frv_pack(self.raw, 0, value_params.as_ptr() as *const f32, width_params.as_ptr() as *const f32, bucket.as_ptr(), sym.as_ptr(),
                                    hist.as_ptr(), &mut out, &mut len) })?;
        let v = unsafe { std::slice::from_raw_parts(out, len) }.to_vec();
        unsafe { fri_frv_free(out) };
        Ok(v)
    }

    /// Bands per frame of the host-buffer calls: 0 = automatic (single caller), 1 when an encoder and a
    /// decoder thread drive one handle each.
    pub fn set_bands(&mut self, bands: i32) -> Result<(), String> { check(unsafe { fri_plan_set_bands(self.raw, bands) }) }

    /// Asynchronous mode: the host-buffer calls return once enqueued.
    /// # Safety
    /// Every buffer passed to a call made in asynchronous mode must be a `PinnedBuf` (the library rejects
    /// pageable memory with an error) that is neither dropped, read nor written until `sync()` has returned.
    pub unsafe fn set_async(&mut self, on: bool) -> Result<(), String> { check(fri_plan_set_async(self.raw, on as c_int)) }
    pub fn sync(&mut self) -> Result<(), String> { check(unsafe { fri_plan_sync(self.raw) }) }
    /// Hint: consecutive *_device calls on one stream are independent (their kernels may overlap).
    /// # Safety
    /// Dependent calls (e.g. encode then decode of the same coefficient buffer) with the hint set are a data race.
    pub unsafe fn set_independent_calls(&mut self, on: bool) -> Result<(), String> { check(fri_plan_set_independent_calls(self.raw, on as c_int)) }

    // ---- device-resident entry points (raw device pointers from the caller's CUDA context)
    /// # Safety
    /// `d_pixels` / `d_coefs` must be device allocations of n_frames frames / coefficient blocks on the plan's
    /// device, `stream` a cudaStream_t of that device (or null).
    pub unsafe fn encode_tq_device(&self, d_pixels: *const c_void, n_frames: u32, q: &[i32; 32], d_coefs: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_encode_tq_device(self.raw, d_pixels, n_frames, q.as_ptr(), d_coefs, stream))
    }
    /// # Safety
    /// See `encode_tq_device`.
    pub unsafe fn decode_tq_device(&self, d_coefs: *const i32, n_frames: u32, q: &[i32; 32], d_pixels: *mut c_void, stream: *mut c_void) -> Result<(), String> {
        check(fri_decode_tq_device(self.raw, d_coefs, n_frames, q.as_ptr(), FRI_DEQUANT_DIVIDE, d_pixels, stream))
    }
    /// # Safety
    /// See `encode_tq_device`; int16 coefficient arrays.
    pub unsafe fn encode_tq_device16(&self, d_pixels: *const c_void, n_frames: u32, q: &[i32; 32], d_coefs: *mut i16, stream: *mut c_void) -> Result<(), String> {
        check(fri_encode_tq_device16(self.raw, d_pixels, n_frames, q.as_ptr(), d_coefs, stream))
    }
    /// # Safety
    /// See `encode_tq_device`; int16 coefficient arrays.
    pub unsafe fn decode_tq_device16(&self, d_coefs: *const i16, n_frames: u32, q: &[i32; 32], d_pixels: *mut c_void, stream: *mut c_void) -> Result<(), String> {
        check(fri_decode_tq_device16(self.raw, d_coefs, n_frames, q.as_ptr(), FRI_DEQUANT_DIVIDE, d_pixels, stream))
    }
    /// # Safety
    /// Device pointers: dense blocks in, [n_frames][C][emission_count] streams out.
    pub unsafe fn emit_device(&mut self, d_coefs: *const i32, n_frames: u32, d_out: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_emit_device(self.raw, d_coefs, n_frames, d_out, stream))
    }
    /// # Safety
    /// As `emit_device`, int16 streams.
    pub unsafe fn emit_device16(&mut self, d_coefs: *const i32, n_frames: u32, d_out: *mut i16, stream: *mut c_void) -> Result<(), String> {
        check(fri_emit_device16(self.raw, d_coefs, n_frames, d_out, stream))
    }
    /// # Safety
    /// As `emit_device`, 10-bit packed streams ([n_frames][C][emission_packed_bytes], 16-byte aligned).
    pub unsafe fn emit_device10(&mut self, d_coefs: *const i32, n_frames: u32, d_out: *mut u8, stream: *mut c_void) -> Result<(), String> {
        check(fri_emit_device10(self.raw, d_coefs, n_frames, d_out, stream))
    }
    /// Prediction + context bucketing on the device (prediction.rs:224-323): per emitted coefficient the context
    /// bucket, the prediction and the zig-zag symbol, plus the per-context histograms the rANS tables are built from.
    /// # Safety
    /// Device pointers sized as include/fri_cuda.h states; `params` are the [C][3][6] value / width predictors.
    #[allow(clippy::too_many_arguments)]
    pub unsafe fn predict_device(&mut self, d_coefs: *const i32, n_frames: u32, value_params: &[[[f32; 6]; 3]], width_params: &[[[f32; 6]; 3]],
                                 d_bucket: *mut u8, d_pred: *mut i32, d_sym: *mut u16, d_hist: *mut u32, d_overflow: *mut u32, stream: *mut c_void) -> Result<(), String> {
        want_len("value_params", value_params.len(), self.channels)?;
        want_len("width_params", width_params.len(), self.channels)?;
        check(fri_predict_device(self.raw, d_coefs, n_frames, value_params.as_ptr() as *const f32, width_params.as_ptr() as *const f32,
                                 d_bucket, d_pred, d_sym, d_hist, d_overflow, stream))
    }
    /// # Safety
    /// Device pointers: streams in, dense blocks out (None slots zeroed).
    pub unsafe fn unemit_device(&mut self, d_streams: *const i32, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_unemit_device(self.raw, d_streams, n_frames, d_coefs, stream))
    }
    /// # Safety
    /// As `unemit_device`, int16 streams.
    pub unsafe fn unemit_device16(&mut self, d_streams: *const i16, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_unemit_device16(self.raw, d_streams, n_frames, d_coefs, stream))
    }
    /// # Safety
    /// As `unemit_device`, 10-bit packed streams.
    pub unsafe fn unemit_device10(&mut self, d_streams: *const u8, n_frames: u32, d_coefs: *mut i32, stream: *mut c_void) -> Result<(), String> {
        check(fri_unemit_device10(self.raw, d_streams, n_frames, d_coefs, stream))
    }
}

// One handle per host thread: the library keeps no global state besides the thread-local error
// string, so an encoder thread and a decoder thread can each own a Plan and run concurrently.
unsafe impl Send for Plan {}

impl Drop for Plan {
    fn drop(&mut self) { unsafe { fri_plan_destroy(self.raw) } }
}

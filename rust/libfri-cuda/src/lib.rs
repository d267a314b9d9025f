//! Safe wrapper over libfri_cuda's C ABI (include/fri_cuda.h).  Untested here: no cargo/rustc in
//! the build image; every item maps 1:1 onto a prototype of the header.
use std::ffi::{c_char, c_int, c_void, CStr};

#[repr(C)]
pub struct FriPlan {
    _private: [u8; 0],
}

pub const FRI_DEQUANT_DIVIDE: c_int = 0; // quantization.rs:37 — the reference divides again
pub const FRI_DEQUANT_MULTIPLY: c_int = 1;

extern "C" {
    fn fri_last_error() -> *const c_char;
    fn fri_plan_create(out: *mut *mut FriPlan, device: c_int, width: u32, height: u32, channels: u32, depth: u32,
                       sample_bytes: u32) -> c_int;
    fn fri_plan_destroy(plan: *mut FriPlan);
    fn fri_plan_num_tiles(plan: *const FriPlan) -> u32;
    fn fri_plan_coefs_per_frame(plan: *const FriPlan) -> u64;
    fn fri_plan_centers(plan: *const FriPlan, centers: *mut i32) -> c_int;
    fn fri_plan_masks(plan: *const FriPlan, masks: *mut u32) -> c_int;
    fn fri_encode_tq(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, coefs: *mut i32) -> c_int;
    fn fri_decode_tq(plan: *mut FriPlan, coefs: *const i32, n_frames: u32, q: *const i32, dequant_mode: c_int,
                     pixels: *mut c_void) -> c_int;
    fn fri_plan_set_bands(plan: *mut FriPlan, bands: c_int) -> c_int;
    fn fri_plan_set_async(plan: *mut FriPlan, on: c_int) -> c_int;
    fn fri_plan_sync(plan: *mut FriPlan) -> c_int;
    // 16-bit transport of the same two calls (8-bit samples): half the bytes over PCIe
    fn fri_encode_tq16(plan: *mut FriPlan, pixels: *const c_void, n_frames: u32, q: *const i32, coefs: *mut i16) -> c_int;
    fn fri_decode_tq16(plan: *mut FriPlan, coefs: *const i16, n_frames: u32, q: *const i32, dequant_mode: c_int,
                       pixels: *mut c_void) -> c_int;
}

fn check(rc: c_int) -> Result<(), String> {
    if rc == 0 {
        Ok(())
    } else {
        // Err(String) is what every libfri stage returns (encoder.rs:19-46)
        Err(unsafe { CStr::from_ptr(fri_last_error()) }.to_string_lossy().into_owned())
    }
}

/// Lattice + launch geometry for one image size (replaces fractal_divide + Fractal::new + retain,
/// wavelet_transform.rs:450-484, :42-69, :415-416, and from_metadata :392-403 on decode).
pub struct Plan(*mut FriPlan);

impl Plan {
    pub fn new(device: i32, width: u32, height: u32, channels: u32) -> Result<Self, String> {
        let mut p = std::ptr::null_mut();
        check(unsafe { fri_plan_create(&mut p, device, width, height, channels, 9, 1) })?;
        Ok(Plan(p))
    }
    pub fn num_tiles(&self) -> usize { unsafe { fri_plan_num_tiles(self.0) as usize } }
    pub fn coefs_per_frame(&self) -> usize { unsafe { fri_plan_coefs_per_frame(self.0) as usize } }
    /// (re, im) of every retained tile, in the order of the coefficient blocks.
    pub fn centers(&self) -> Result<Vec<[i32; 2]>, String> {
        let mut v = vec![[0i32; 2]; self.num_tiles()];
        check(unsafe { fri_plan_centers(self.0, v.as_mut_ptr() as *mut i32) })?;
        Ok(v)
    }
    /// 512-bit Some/None mask per tile (bit i of word i/32 set <=> coefficient i is Some).
    pub fn masks(&self) -> Result<Vec<[u32; 16]>, String> {
        let mut v = vec![[0u32; 16]; self.num_tiles()];
        check(unsafe { fri_plan_masks(self.0, v.as_mut_ptr() as *mut u32) })?;
        Ok(v)
    }
    /// wavelet_transform::encode + quantization::encode (encoder.rs:26-33), fused.
    pub fn encode_tq(&mut self, pixels: &[u8], q: &[i32; 32]) -> Result<Vec<i32>, String> {
        let mut coefs = vec![0i32; self.coefs_per_frame()];
        check(unsafe { fri_encode_tq(self.0, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), coefs.as_mut_ptr()) })?;
        Ok(coefs)
    }
    /// quantization::decode + wavelet_transform::decode (decoder.rs:27-34), fused.
    pub fn decode_tq(&mut self, coefs: &[i32], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        check(unsafe { fri_decode_tq(self.0, coefs.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
    /// Bands per frame of the host-buffer calls: 0 = automatic (single caller), 1 when an encoder and a
    /// decoder thread drive one handle each.
    pub fn set_bands(&mut self, bands: i32) -> Result<(), String> { check(unsafe { fri_plan_set_bands(self.0, bands) }) }
    /// Asynchronous mode: encode_tq* / decode_tq* return once enqueued (the slices passed must then be
    /// pinned memory from fri_host_alloc and must outlive the next `sync`).
    pub fn set_async(&mut self, on: bool) -> Result<(), String> { check(unsafe { fri_plan_set_async(self.0, on as c_int) }) }
    pub fn sync(&mut self) -> Result<(), String> { check(unsafe { fri_plan_sync(self.0) }) }
    /// encode_tq with int16 coefficients on the host side (every coefficient of an 8-bit image fits:
    /// |residue| <= 255, wavelet_transform.rs:211-218); widen while applying the mask.
    pub fn encode_tq16(&mut self, pixels: &[u8], q: &[i32; 32]) -> Result<Vec<i16>, String> {
        let mut coefs = vec![0i16; self.coefs_per_frame()];
        check(unsafe { fri_encode_tq16(self.0, pixels.as_ptr() as *const c_void, 1, q.as_ptr(), coefs.as_mut_ptr()) })?;
        Ok(coefs)
    }
    /// decode_tq from int16 coefficients (a decodable container never holds more: 1024-symbol
    /// alphabet, entropy_coding.rs:25).
    pub fn decode_tq16(&mut self, coefs: &[i16], q: &[i32; 32], pixels: &mut [u8]) -> Result<(), String> {
        check(unsafe { fri_decode_tq16(self.0, coefs.as_ptr(), 1, q.as_ptr(), FRI_DEQUANT_DIVIDE, pixels.as_mut_ptr() as *mut c_void) })
    }
}

// One handle per host thread: the library keeps no global state besides the thread-local error
// string, so an encoder thread and a decoder thread can each own a Plan and run concurrently.
unsafe impl Send for Plan {}

impl Drop for Plan {
    fn drop(&mut self) { unsafe { fri_plan_destroy(self.0) } }
}

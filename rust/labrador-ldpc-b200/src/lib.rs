//! `LDPCCode` with the reference crate's API, executed on a B200 through the C ABI of
//! `liblabrador_ldpc.so` (include/labrador_ldpc.h).  Same enum discriminants, method names,
//! argument meaning and length assertions as labrador-ldpc 1.2.1 (src/codes/mod.rs:37-66,
//! src/encoder.rs:292-315, src/decoder.rs:93-116, 243, 347, 484, 498), plus `_batch` variants.
//!
//! There is no CPU fallback: every compute method calls CUDA kernels.
//! UNCOMPILED in this repository (no Rust toolchain in the image); kept in sync with the header
//! by tests/test_capi_host.py::test_rust_binding_matches_header.
#![allow(clippy::too_many_arguments)]

use std::os::raw::{c_int, c_void};

#[repr(C)]
#[derive(Copy, Clone, Debug, Eq, PartialEq, Hash)]
pub enum LDPCCode {
    TC128 = 0, TC256 = 1, TC512 = 2,
    TM1280 = 3, TM1536 = 4, TM2048 = 5,
    TM5120 = 6, TM6144 = 7, TM8192 = 8,
    /// Extension: the k = 16384 codes whose parity-check constants the reference carries without supporting them
    /// (src/lib.rs:81-83).  Decoded like every TM code, encoded through the sparse parity-check matrix.
    TM20480 = 9, TM24576 = 10, TM32768 = 11,
}

#[derive(Debug)]
pub struct Error { pub code: i32, pub message: String }

pub mod ffi {
    use super::*;
    extern "C" {
        pub fn labrador_ldpc_code_n(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_code_k(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_bf_working_len(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_ms_working_len(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_ms_working_u8_len(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_output_len(code: LDPCCode) -> usize;
        pub fn labrador_ldpc_cuda_init(devices: *const c_int, n_devices: c_int) -> c_int;
        pub fn labrador_ldpc_cuda_shutdown();
        pub fn labrador_ldpc_last_error() -> *const std::os::raw::c_char;
        pub fn labrador_ldpc_encode_batch(code: LDPCCode, codewords: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_copy_encode_batch(code: LDPCCode, data: *const u8, codewords: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_decode_bf_batch(code: LDPCCode, input: *const u8, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_i8_batch(code: LDPCCode, llrs: *const i8, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_i16_batch(code: LDPCCode, llrs: *const i16, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_i32_batch(code: LDPCCode, llrs: *const i32, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_f32_batch(code: LDPCCode, llrs: *const f32, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_f64_batch(code: LDPCCode, llrs: *const f64, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_batch_async(code: LDPCCode, llr_type: c_int, input: *const u8,
            llrs: *mut c_void, batch: usize, stream: *mut c_void) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_batch_async(code: LDPCCode, llr_type: c_int, llrs: *const c_void,
            output: *mut u8, batch: usize, stream: *mut c_void) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_i8_batch(code: LDPCCode, input: *const u8, llrs: *mut i8, batch: usize) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_i16_batch(code: LDPCCode, input: *const u8, llrs: *mut i16, batch: usize) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_i32_batch(code: LDPCCode, input: *const u8, llrs: *mut i32, batch: usize) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_f32_batch(code: LDPCCode, input: *const u8, llrs: *mut f32, batch: usize) -> c_int;
        pub fn labrador_ldpc_hard_to_llrs_f64_batch(code: LDPCCode, input: *const u8, llrs: *mut f64, batch: usize) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_i8_batch(code: LDPCCode, llrs: *const i8, output: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_i16_batch(code: LDPCCode, llrs: *const i16, output: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_i32_batch(code: LDPCCode, llrs: *const i32, output: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_f32_batch(code: LDPCCode, llrs: *const f32, output: *mut u8, batch: usize) -> c_int;
        pub fn labrador_ldpc_llrs_to_hard_f64_batch(code: LDPCCode, llrs: *const f64, output: *mut u8, batch: usize) -> c_int;
        // fused front ends (include/labrador_ldpc.h, "Fused front ends of decode_ms")
        pub fn labrador_ldpc_decode_ms_i8_soft_batch(code: LDPCCode, soft: *const f32, scale: f32, limit: c_int,
            output: *mut u8, batch: usize, max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_i16_soft_batch(code: LDPCCode, soft: *const f32, scale: f32, limit: c_int,
            output: *mut u8, batch: usize, max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_decode_ms_i8_hard_batch(code: LDPCCode, input: *const u8, output: *mut u8, batch: usize,
            max_iters: usize, success: *mut u8, iters_run: *mut u32) -> c_int;
        pub fn labrador_ldpc_quantise_i8_batch(code: LDPCCode, soft: *const f32, scale: f32, limit: c_int,
            llrs: *mut i8, batch: usize) -> c_int;
        pub fn labrador_ldpc_quantise_i16_batch(code: LDPCCode, soft: *const f32, scale: f32, limit: c_int,
            llrs: *mut i16, batch: usize) -> c_int;
    }
}

fn check(rc: c_int) -> Result<(), Error> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(ffi::labrador_ldpc_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message: msg })
}

/// LLR element types the min-sum decoder accepts (mirror of `DecodeFrom`, src/decoder.rs:22-86).
pub trait DecodeFrom: Copy {
    unsafe fn decode_ms_batch(code: LDPCCode, llrs: *const Self, output: *mut u8, batch: usize, max_iters: usize,
                              success: *mut u8, iters: *mut u32) -> c_int;
    unsafe fn hard_to_llrs_batch(code: LDPCCode, input: *const u8, llrs: *mut Self, batch: usize) -> c_int;
    unsafe fn llrs_to_hard_batch(code: LDPCCode, llrs: *const Self, output: *mut u8, batch: usize) -> c_int;
}
macro_rules! impl_decode_from {
    ($t:ty, $ms:ident, $h2l:ident, $l2h:ident) => {
        impl DecodeFrom for $t {
            unsafe fn decode_ms_batch(code: LDPCCode, llrs: *const Self, output: *mut u8, batch: usize,
                                      max_iters: usize, success: *mut u8, iters: *mut u32) -> c_int {
                ffi::$ms(code, llrs, output, batch, max_iters, success, iters)
            }
            unsafe fn hard_to_llrs_batch(code: LDPCCode, input: *const u8, llrs: *mut Self, batch: usize) -> c_int {
                ffi::$h2l(code, input, llrs, batch)
            }
            unsafe fn llrs_to_hard_batch(code: LDPCCode, llrs: *const Self, output: *mut u8, batch: usize) -> c_int {
                ffi::$l2h(code, llrs, output, batch)
            }
        }
    };
}
impl_decode_from!(i8, labrador_ldpc_decode_ms_i8_batch, labrador_ldpc_hard_to_llrs_i8_batch, labrador_ldpc_llrs_to_hard_i8_batch);
impl_decode_from!(i16, labrador_ldpc_decode_ms_i16_batch, labrador_ldpc_hard_to_llrs_i16_batch, labrador_ldpc_llrs_to_hard_i16_batch);
impl_decode_from!(i32, labrador_ldpc_decode_ms_i32_batch, labrador_ldpc_hard_to_llrs_i32_batch, labrador_ldpc_llrs_to_hard_i32_batch);
impl_decode_from!(f32, labrador_ldpc_decode_ms_f32_batch, labrador_ldpc_hard_to_llrs_f32_batch, labrador_ldpc_llrs_to_hard_f32_batch);
impl_decode_from!(f64, labrador_ldpc_decode_ms_f64_batch, labrador_ldpc_hard_to_llrs_f64_batch, labrador_ldpc_llrs_to_hard_f64_batch);

impl LDPCCode {
    pub fn n(self) -> usize { unsafe { ffi::labrador_ldpc_code_n(self) } }
    pub fn k(self) -> usize { unsafe { ffi::labrador_ldpc_code_k(self) } }
    pub fn output_len(self) -> usize { unsafe { ffi::labrador_ldpc_output_len(self) } }
    pub fn punctured_bits(self) -> usize { self.output_len() * 8 - self.n() }
    pub fn decode_bf_working_len(self) -> usize { unsafe { ffi::labrador_ldpc_bf_working_len(self) } }
    pub fn decode_ms_working_len(self) -> usize { unsafe { ffi::labrador_ldpc_ms_working_len(self) } }
    pub fn decode_ms_working_u8_len(self) -> usize { unsafe { ffi::labrador_ldpc_ms_working_u8_len(self) } }

    // ---- the reference's single-codeword methods (same signatures; scratch arguments are unused) ----

    /// src/encoder.rs:293 (u8 view; the u32/u64 views are byte-identical, src/encoder.rs:157-159).
    pub fn encode<'a>(&self, codeword: &'a mut [u8]) -> &'a mut [u8] {
        assert_eq!(codeword.len() * 8, self.n(), "codeword must be n bits long");
        check(unsafe { ffi::labrador_ldpc_encode_batch(*self, codeword.as_mut_ptr(), 1) }).expect("encode");
        codeword
    }
    /// src/encoder.rs:309
    pub fn copy_encode<'a>(&self, data: &[u8], codeword: &'a mut [u8]) -> &'a mut [u8] {
        assert_eq!(data.len() * 8, self.k(), "data must be k bits long");
        assert_eq!(codeword.len() * 8, self.n(), "codeword must be n bits long");
        check(unsafe { ffi::labrador_ldpc_copy_encode_batch(*self, data.as_ptr(), codeword.as_mut_ptr(), 1) })
            .expect("copy_encode");
        codeword
    }
    /// src/decoder.rs:243
    pub fn decode_bf(self, input: &[u8], output: &mut [u8], working: &mut [u8], maxiters: usize) -> (bool, usize) {
        assert_eq!(input.len(), self.n() / 8, "input.len() != n/8");
        assert_eq!(output.len(), self.output_len(), "output.len != (n+p)/8");
        assert_eq!(working.len(), self.decode_bf_working_len(), "working.len() incorrect");
        let (mut ok, mut it) = (0u8, 0u32);
        check(unsafe { ffi::labrador_ldpc_decode_bf_batch(self, input.as_ptr(), output.as_mut_ptr(), 1, maxiters,
                                                         &mut ok, &mut it) }).expect("decode_bf");
        (ok != 0, it as usize)
    }
    /// src/decoder.rs:347
    pub fn decode_ms<T: DecodeFrom>(self, llrs: &[T], output: &mut [u8], working: &mut [T], working_u8: &mut [u8],
                                    maxiters: usize) -> (bool, usize) {
        assert_eq!(llrs.len(), self.n(), "llrs.len() != n");
        assert_eq!(output.len(), self.output_len(), "output.len() != (n+p)/8");
        assert_eq!(working.len(), self.decode_ms_working_len(), "working.len() incorrect");
        assert_eq!(working_u8.len(), self.decode_ms_working_u8_len(), "working_u8 != (n+p-k)/8");
        let (mut ok, mut it) = (0u8, 0u32);
        check(unsafe { T::decode_ms_batch(self, llrs.as_ptr(), output.as_mut_ptr(), 1, maxiters, &mut ok, &mut it) })
            .expect("decode_ms");
        (ok != 0, it as usize)
    }
    /// src/decoder.rs:484
    pub fn hard_to_llrs<T: DecodeFrom>(self, input: &[u8], llrs: &mut [T]) {
        assert_eq!(input.len(), self.n() / 8, "input.len() != n/8");
        assert_eq!(llrs.len(), self.n(), "llrs.len() != n");
        check(unsafe { T::hard_to_llrs_batch(self, input.as_ptr(), llrs.as_mut_ptr(), 1) }).expect("hard_to_llrs");
    }
    /// src/decoder.rs:498
    pub fn llrs_to_hard<T: DecodeFrom>(self, llrs: &[T], output: &mut [u8]) {
        assert_eq!(llrs.len(), self.n(), "llrs.len() != n");
        assert_eq!(output.len(), self.n() / 8, "output.len() != n/8");
        check(unsafe { T::llrs_to_hard_batch(self, llrs.as_ptr(), output.as_mut_ptr(), 1) }).expect("llrs_to_hard");
    }

    // ---- batched additions: frame-major contiguous slices, device owns all scratch ----

    pub fn copy_encode_batch(&self, data: &[u8], codewords: &mut [u8]) -> Result<(), Error> {
        let batch = data.len() / (self.k() / 8);
        assert_eq!(data.len(), batch * self.k() / 8);
        assert_eq!(codewords.len(), batch * self.n() / 8);
        check(unsafe { ffi::labrador_ldpc_copy_encode_batch(*self, data.as_ptr(), codewords.as_mut_ptr(), batch) })
    }
    pub fn decode_ms_batch<T: DecodeFrom>(self, llrs: &[T], output: &mut [u8], maxiters: usize,
                                          success: &mut [u8], iters: &mut [u32]) -> Result<(), Error> {
        let batch = llrs.len() / self.n();
        assert_eq!(llrs.len(), batch * self.n());
        assert_eq!(output.len(), batch * self.output_len());
        assert_eq!(success.len(), batch);
        assert_eq!(iters.len(), batch);
        check(unsafe { T::decode_ms_batch(self, llrs.as_ptr(), output.as_mut_ptr(), batch, maxiters,
                                          success.as_mut_ptr(), iters.as_mut_ptr()) })
    }
    pub fn decode_bf_batch(self, input: &[u8], output: &mut [u8], maxiters: usize,
                           success: &mut [u8], iters: &mut [u32]) -> Result<(), Error> {
        let batch = input.len() / (self.n() / 8);
        assert_eq!(input.len(), batch * self.n() / 8);
        assert_eq!(output.len(), batch * self.output_len());
        assert_eq!(success.len(), batch);     // the C side writes `batch` entries into both
        assert_eq!(iters.len(), batch);
        check(unsafe { ffi::labrador_ldpc_decode_bf_batch(self, input.as_ptr(), output.as_mut_ptr(), batch, maxiters,
                                                         success.as_mut_ptr(), iters.as_mut_ptr()) })
    }

    /// decode_ms::<i8> of `clamp(round(soft * scale), -limit, limit)`, quantised inside the decoder.
    pub fn decode_ms_soft_i8_batch(self, soft: &[f32], scale: f32, limit: i8, output: &mut [u8], maxiters: usize,
                                   success: &mut [u8], iters: &mut [u32]) -> Result<(), Error> {
        let batch = soft.len() / self.n();
        assert_eq!(soft.len(), batch * self.n());
        assert_eq!(output.len(), batch * self.output_len());
        assert_eq!(success.len(), batch);
        assert_eq!(iters.len(), batch);
        check(unsafe { ffi::labrador_ldpc_decode_ms_i8_soft_batch(self, soft.as_ptr(), scale, limit as c_int,
                                                                 output.as_mut_ptr(), batch, maxiters,
                                                                 success.as_mut_ptr(), iters.as_mut_ptr()) })
    }
    /// decode_ms::<i8> of `hard_to_llrs(input)`, converted inside the decoder.
    pub fn decode_ms_hard_batch(self, input: &[u8], output: &mut [u8], maxiters: usize,
                                success: &mut [u8], iters: &mut [u32]) -> Result<(), Error> {
        let batch = input.len() / (self.n() / 8);
        assert_eq!(input.len(), batch * self.n() / 8);
        assert_eq!(output.len(), batch * self.output_len());
        assert_eq!(success.len(), batch);
        assert_eq!(iters.len(), batch);
        check(unsafe { ffi::labrador_ldpc_decode_ms_i8_hard_batch(self, input.as_ptr(), output.as_mut_ptr(), batch,
                                                                 maxiters, success.as_mut_ptr(), iters.as_mut_ptr()) })
    }
}

//! Drop-in body for src/processing/nms.rs: same signature, same result (indices into `dets`, pick order), computed by
//! the B200 kernels.  `CTX` is the per-thread context (one fd_ctx per GPU per host thread).
use ndarray::Array2;
use crate::ffi;

thread_local! { static CTX: *mut ffi::fd_ctx = unsafe {
    let mut c = std::ptr::null_mut();
    ffi::check(ffi::fd_ctx_create(0, std::ptr::null(), &mut c)).expect("fd_ctx_create");
    c
}; }

pub fn nms(dets: &Array2<f32>, thresh: f32) -> Vec<usize> {
    let dets = dets.as_standard_layout();                      // row-major (K,5) as the FFI expects
    let k = dets.nrows();
    let mut keep = vec![0i32; k.max(1)];
    let mut n = 0;
    CTX.with(|&c| unsafe {
        ffi::check(ffi::fd_nms(c, dets.as_ptr(), k as i32, thresh, keep.as_mut_ptr(), &mut n)).expect("fd_nms");
    });
    keep[..n as usize].iter().map(|&i| i as usize).collect()
}

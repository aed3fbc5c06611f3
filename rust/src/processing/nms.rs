//! Drop-in body for src/processing/nms.rs:3-65: same signature, same result (indices into `dets`, pick order), computed by
//! the B200 kernels (stable descending-score sort + exact greedy IoU suppression on the device).
use ndarray::Array2;
use crate::{ctx::with_ctx, ffi};

pub fn nms(dets: &Array2<f32>, thresh: f32) -> Vec<usize> {
    let dets = dets.as_standard_layout();                      // row-major (K,5) as the FFI expects
    let k = dets.nrows();
    let mut keep = vec![0i32; k.max(1)];
    let mut n = 0;
    let rc = with_ctx(|c| unsafe { ffi::fd_nms(c, dets.as_ptr(), k as i32, thresh, keep.as_mut_ptr(), &mut n) });
    if rc == ffi::FD_ERR_NAN_SCORE {
        // The reference sorts with `partial_cmp(..).unwrap_or(Equal)` (nms.rs:6): it does not panic on a NaN score, but the
        // comparator is then not a total order and the resulting order is unspecified.  This wrapper never panics either and
        // makes the case deterministic: NaN-scored boxes are ranked after every other box, in index order.
        let mut clean = dets.to_owned();
        for mut row in clean.rows_mut() {
            if row[4].is_nan() { row[4] = f32::NEG_INFINITY; }
        }
        let rc2 = with_ctx(|c| unsafe { ffi::fd_nms(c, clean.as_ptr(), k as i32, thresh, keep.as_mut_ptr(), &mut n) });
        ffi::check(rc2).expect("fd_nms");
    } else {
        ffi::check(rc).expect("fd_nms");
    }
    keep[..n as usize].iter().map(|&i| i as usize).collect()
}

//! Drop-in body for src/rcnn/cpu_nms.rs:10-55 — the `>=` variant (a box is suppressed when ovr >= thresh).
use ndarray::ArrayView2;
use crate::{ctx::with_ctx, ffi};

#[allow(dead_code)]
fn cpu_nms(dets: ArrayView2<f32>, thresh: f32) -> Vec<usize> {
    let dets = dets.as_standard_layout();
    let k = dets.nrows();
    let mut keep = vec![0i32; k.max(1)];
    let mut n = 0;
    with_ctx(|c| ffi::check(unsafe { ffi::fd_cpu_nms(c, dets.as_ptr(), k as i32, thresh, keep.as_mut_ptr(), &mut n) })).expect("fd_cpu_nms");
    keep[..n as usize].iter().map(|&i| i as usize).collect()
}

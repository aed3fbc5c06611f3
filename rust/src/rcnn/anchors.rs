//! Drop-in body for src/rcnn/anchors.rs:3-21: the (H,W,A,4) anchor plane of one stride.
use ndarray::{Array2, Array4};
use crate::{ctx::with_ctx, ffi};

pub fn anchors(height: usize, width: usize, stride: usize, base_anchors: &Array2<f32>) -> Array4<f32> {
    let base = base_anchors.as_standard_layout();
    let a = base.nrows();
    let mut out = vec![0f32; (height * width * a * 4).max(1)];
    with_ctx(|c| ffi::check(unsafe {
        ffi::fd_anchors_plane(c, height as i32, width as i32, stride as i32, base.as_ptr(), a as i32, out.as_mut_ptr())
    }))
    .expect("fd_anchors_plane");
    out.truncate(height * width * a * 4);
    Array4::from_shape_vec((height, width, a, 4), out).unwrap()
}

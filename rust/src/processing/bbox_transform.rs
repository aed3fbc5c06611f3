//! Drop-in bodies for src/processing/bbox_transform.rs (same names, argument meaning and results; every op rounded
//! separately on the device, no FMA contraction).
use ndarray::Array2;
use crate::{ctx::with_ctx, ffi};

fn rows_cols(a: &Array2<f32>) -> (i32, i32) { (a.nrows() as i32, a.ncols() as i32) }

/// bbox_transform.rs:2-24 — N x K IoU matrix with the `+1` convention.
pub fn bbox_overlaps_py(boxes: &Array2<f32>, query_boxes: &Array2<f32>) -> Array2<f32> {
    let (b, q) = (boxes.as_standard_layout(), query_boxes.as_standard_layout());
    let (n, k) = (b.nrows(), q.nrows());
    let mut out = vec![0f32; (n * k).max(1)];
    with_ctx(|c| ffi::check(unsafe { ffi::fd_bbox_overlaps(c, b.as_ptr(), n as i32, q.as_ptr(), k as i32, out.as_mut_ptr()) }))
        .expect("fd_bbox_overlaps");
    out.truncate(n * k);
    Array2::from_shape_vec((n, k), out).unwrap()
}

/// bbox_transform.rs:27-45 — in place; im_shape = (height, width).
pub fn clip_boxes(boxes: &mut Array2<f32>, im_shape: (usize, usize)) {
    let mut tmp = boxes.as_standard_layout().to_owned();
    let (r, cl) = rows_cols(&tmp);
    with_ctx(|c| ffi::check(unsafe { ffi::fd_clip_boxes(c, tmp.as_mut_ptr(), r, cl, im_shape.0 as i32, im_shape.1 as i32) }))
        .expect("fd_clip_boxes");
    boxes.assign(&tmp);
}

/// bbox_transform.rs:47-65 — in place.
pub fn clip_points(points: &mut Array2<f32>, im_shape: (usize, usize)) {
    let mut tmp = points.as_standard_layout().to_owned();
    let (r, cl) = rows_cols(&tmp);
    with_ctx(|c| ffi::check(unsafe { ffi::fd_clip_points(c, tmp.as_mut_ptr(), r, cl, im_shape.0 as i32, im_shape.1 as i32) }))
        .expect("fd_clip_points");
    points.assign(&tmp);
}

/// bbox_transform.rs:67-88 — regression targets (dx, dy, dw, dh) of gt_rois w.r.t. ex_rois.
pub fn nonlinear_transform(ex_rois: &Array2<f32>, gt_rois: &Array2<f32>) -> Array2<f32> {
    let (e, g) = (ex_rois.as_standard_layout(), gt_rois.as_standard_layout());
    let n = e.nrows();
    let mut out = vec![0f32; (n * 4).max(1)];
    with_ctx(|c| ffi::check(unsafe { ffi::fd_nonlinear_transform(c, e.as_ptr(), g.as_ptr(), n as i32, out.as_mut_ptr()) }))
        .expect("fd_nonlinear_transform");
    out.truncate(n * 4);
    Array2::from_shape_vec((n, 4), out).unwrap()
}

/// bbox_transform.rs:90-120 — every group of 4 delta columns regressed from `boxes`.
pub fn nonlinear_pred(boxes: &Array2<f32>, box_deltas: &Array2<f32>) -> Array2<f32> {
    let (b, d) = (boxes.as_standard_layout(), box_deltas.as_standard_layout());
    let (n, ncols) = (d.nrows(), d.ncols());
    let mut out = vec![0f32; (n * ncols).max(1)];
    with_ctx(|c| ffi::check(unsafe { ffi::fd_nonlinear_pred(c, b.as_ptr(), d.as_ptr(), n as i32, ncols as i32, out.as_mut_ptr()) }))
        .expect("fd_nonlinear_pred");
    out.truncate(n * ncols);
    Array2::from_shape_vec((n, ncols), out).unwrap()
}

/// bbox_transform.rs:123-160 — flat (N,10) landmark deltas regressed from the ANCHOR boxes.
pub fn landmark_pred(boxes: &Array2<f32>, point_deltas: &Array2<f32>) -> Array2<f32> {
    let (b, d) = (boxes.as_standard_layout(), point_deltas.as_standard_layout());
    let n = d.nrows();
    let mut out = vec![0f32; (n * 10).max(1)];
    with_ctx(|c| ffi::check(unsafe { ffi::fd_landmark_pred(c, b.as_ptr(), d.as_ptr(), n as i32, out.as_mut_ptr()) }))
        .expect("fd_landmark_pred");
    out.truncate(n * 10);
    Array2::from_shape_vec((n, 10), out).unwrap()
}

/// bbox_transform.rs:162-186.
pub fn iou_pred(boxes: &Array2<f32>, box_deltas: &Array2<f32>, num_classes: usize) -> Array2<f32> {
    let (b, d) = (boxes.as_standard_layout(), box_deltas.as_standard_layout());
    let (n, ncols) = (d.nrows(), d.ncols());
    let mut out = vec![0f32; (n * ncols).max(1)];
    with_ctx(|c| ffi::check(unsafe {
        ffi::fd_iou_pred(c, b.as_ptr(), d.as_ptr(), n as i32, ncols as i32, num_classes as i32, out.as_mut_ptr())
    }))
    .expect("fd_iou_pred");
    out.truncate(n * ncols);
    Array2::from_shape_vec((n, ncols), out).unwrap()
}

//! src/rcnn/gpu_nms.rs:21-49, un-commented: the binding the reference sketched for its C symbol `_nms` (gpu_nms.hpp:6-8).
//! libfd_b200 exports that literal symbol with the same contract (boxes sorted by score descending, host pointers,
//! keep = indices into the sorted array), so the body is the reference's own: sort, gather, call, map back.
use ndarray::Array2;
use crate::ffi;

pub fn gpu_nms(dets: Array2<f32>, thresh: f32, device_id: i32) -> Vec<usize> {
    let boxes_num = dets.nrows();
    let boxes_dim = dets.ncols();
    if boxes_num == 0 { return Vec::new(); }
    let mut order: Vec<usize> = (0..boxes_num).collect();
    order.sort_by(|&a, &b| dets[[b, 4]].partial_cmp(&dets[[a, 4]]).unwrap_or(std::cmp::Ordering::Equal));
    let mut sorted = Vec::with_capacity(boxes_num * boxes_dim);
    for &i in order.iter() {
        sorted.extend(dets.row(i).iter().cloned());
    }
    let mut keep = vec![0i32; boxes_num];
    let mut num_out = 0;
    unsafe { ffi::_nms(keep.as_mut_ptr(), &mut num_out, sorted.as_ptr(), boxes_num as i32, boxes_dim as i32, thresh, device_id) };
    keep[..num_out as usize].iter().map(|&k| order[k as usize]).collect()
}

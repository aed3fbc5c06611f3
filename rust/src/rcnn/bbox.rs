//! Drop-in body for src/rcnn/bbox.rs:4-30 (same arithmetic as processing::bbox_transform::bbox_overlaps_py).
use ndarray::Array2;

pub(crate) fn bbox_overlaps(boxes: &Array2<f32>, query_boxes: &Array2<f32>) -> Array2<f32> {
    crate::processing::bbox_transform::bbox_overlaps_py(boxes, query_boxes)
}

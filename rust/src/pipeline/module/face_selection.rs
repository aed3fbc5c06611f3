//! Replacement body for FaceSelection::call (src/pipeline/module/face_selection.rs:72-189) — same signature and results
//! (including the reference's quirks: the size filter squares the box WIDTH, :115; the key points come from the FIRST row
//! within 2 px of the selected box, :160-176).  `new` (:14-70) is unchanged.
use anyhow::Error;
use ndarray::{s, Array1, Array2, Array3};
use opencv::core::Mat;
use opencv::prelude::MatTraitConst;
use crate::{ctx::with_ctx, ffi};

pub struct FaceSelection {
    pub margin_center_left_ratio: f32,
    pub margin_center_right_ratio: f32,
    pub margin_edge_ratio: f32,
    pub minimum_face_ratio: f32,
}

impl FaceSelection {
    pub fn call(&self, img: &Mat, face_boxes: Array2<f32>, key_points: Option<Array3<f32>>, is_enroll: Option<bool>,
                _is_debug: Option<bool>) -> Result<(Option<Array1<f32>>, Option<Array2<f32>>), Error> {
        let boxes = face_boxes.as_standard_layout();
        let kps = key_points.as_ref().map(|k| k.as_standard_layout());
        let params = ffi::fd_select_params {
            margin_center_left_ratio: self.margin_center_left_ratio, margin_center_right_ratio: self.margin_center_right_ratio,
            margin_edge_ratio: self.margin_edge_ratio, minimum_face_ratio: self.minimum_face_ratio,
        };
        let (mut bi, mut ki) = (-1i32, -1i32);
        with_ctx(|c| ffi::check(unsafe {
            ffi::fd_face_selection(c, img.rows(), img.cols(), boxes.as_ptr(), kps.as_ref().map_or(std::ptr::null(), |k| k.as_ptr()),
                                   boxes.nrows() as i32, is_enroll.unwrap_or(false) as i32, &params, &mut bi, &mut ki)
        }))?;
        let sel_box = if bi >= 0 { Some(face_boxes.row(bi as usize).to_owned()) } else { None };
        let sel_kps = match (&key_points, ki >= 0) {
            (Some(k), true) => Some(k.slice(s![ki as usize, .., ..]).to_owned()),
            _ => None,
        };
        Ok((sel_box, sel_kps))
    }
}

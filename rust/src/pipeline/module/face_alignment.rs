//! Replacement body for src/pipeline/module/face_alignment.rs:14-141 — same type, same `new` / `call` signatures.
//! estimateAffinePartial2D(LMEDS) + warpAffine, and the bbox-crop fallback (:64-116) when the estimate is empty, all on the
//! device; errors where the reference returns Err (landmarks == None, fallback ROI outside the image).
use anyhow::Error;
use ndarray::{Array1, Array2};
use opencv::core::{Mat, Scalar, CV_8UC3};
use opencv::prelude::{MatTrait, MatTraitConst};
use crate::{ctx::Ctx, ffi};

pub(crate) struct FaceAlignment {
    image_size: (i32, i32),
    standard_landmarks: Array2<f32>,
    fd: Ctx,
}

impl FaceAlignment {
    pub fn new(image_size: (i32, i32), standard_landmarks: Array2<f32>) -> Self {
        let mut cfg: ffi::fd_config = unsafe { std::mem::zeroed() };
        ffi::check(unsafe { ffi::fd_config_default(&mut cfg) }).expect("fd_config_default");
        cfg.crop_w = image_size.0;
        cfg.crop_h = image_size.1;
        for (i, row) in standard_landmarks.rows().into_iter().enumerate().take(5) {
            cfg.template_landmarks[i] = [row[0], row[1]];
        }
        let fd = Ctx::new(0, Some(&cfg)).expect("fd_ctx_create");
        FaceAlignment { image_size, standard_landmarks, fd }
    }

    pub fn call(&self, img: &Mat, bbox: Option<Array1<f32>>, landmarks: Option<Array2<f32>>, _is_debug: Option<bool>) -> Result<Mat, Error> {
        let lmk: Option<Vec<f32>> = landmarks.map(|l| l.as_standard_layout().iter().cloned().collect());
        let bb: Option<Vec<f32>> = bbox.map(|b| b.iter().cloned().take(4).collect());
        let mut out = Mat::new_rows_cols_with_default(self.image_size.1, self.image_size.0, CV_8UC3, Scalar::all(0.0))?;
        let pitch = img.step1(0)? as i32;
        ffi::check(unsafe {
            ffi::fd_align(self.fd.raw(), img.data(), img.rows(), img.cols(), pitch,
                          bb.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),
                          lmk.as_ref().map_or(std::ptr::null(), |v| v.as_ptr()),     // None -> Err, like the reference
                          out.data_mut(), std::ptr::null_mut(), std::ptr::null_mut())
        })?;
        Ok(out)
    }
}

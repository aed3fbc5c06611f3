//! Replacement bodies for src/pipeline/module/face_detection.rs: `_preprocess` (:131-198) + the tensor loop (:220-230) and
//! the post-CNN half of `_forward` + `_postprocess` (:319-493) run on the B200; the Triton gRPC round trip in the middle
//! (:232-312) is the reference's own code and stays where it is.  `call` keeps its signature (:496).
//!
//! Only the parts that change are shown: the struct gains one field (`fd: Ctx`, created in `new` from the same constants
//! `new` already holds, :41-129), `_triton_infer` is the reference's :232-312 cut at the point where it has
//! `raw_output_contents` and the output shapes in hand.
use anyhow::Error;
use ndarray::{Array2, Array3};
use opencv::core::Mat;
use opencv::prelude::MatTraitConst;
use crate::{ctx::Ctx, ffi};

pub struct RetinaFaceDetectionB200 {
    pub fd: Ctx,                       // fd_ctx built from image_size / thresholds / anchors / pixel constants (fd_config)
    pub confidence_threshold: f32,
    pub iou_threshold: f32,
    pub image_size: (i32, i32),
}

impl RetinaFaceDetectionB200 {
    /// RetinaFaceDetection::new (:41-129): the same constants, handed to the library as one fd_config.
    pub fn new(image_size: (i32, i32), confidence_threshold: f32, iou_threshold: f32, device_id: i32) -> Result<Self, Error> {
        let mut cfg: ffi::fd_config = unsafe { std::mem::zeroed() };
        ffi::check(unsafe { ffi::fd_config_default(&mut cfg) })?;   // strides 32/16/8, anchors of generate_anchors_fpn2, means 0, stds 1
        cfg.image_w = image_size.0;
        cfg.image_h = image_size.1;
        cfg.conf_thr = confidence_threshold;
        cfg.iou_thr = iou_threshold;
        Ok(Self { fd: Ctx::new(device_id, Some(&cfg))?, confidence_threshold, iou_threshold, image_size })
    }

    /// `_preprocess` + tensor loop (:131-230): BGR Mat -> (1,3,H,W) f32 tensor + det_scale.
    pub fn _preprocess(&self, image: &Mat) -> Result<(Vec<f32>, f32), Error> {
        let n = 3 * self.image_size.0 as usize * self.image_size.1 as usize;
        let mut tensor = vec![0f32; n];
        let mut det_scale = 0f32;
        let pitch = image.step1(0)? as i32;                  // bytes per row of a CV_8UC3 Mat
        ffi::check(unsafe { ffi::fd_preprocess(self.fd.raw(), image.data(), image.rows(), image.cols(), pitch, tensor.as_mut_ptr(), &mut det_scale) })?;
        Ok((tensor, det_scale))
    }

    /// Post-CNN half of `_forward` + `_postprocess` (:286-493) on Triton's `raw_output_contents` (little-endian f32 bytes,
    /// utils.rs:126-132) and the outputs' shapes, in net_out order (scores, bbox, landmarks per stride 32/16/8).
    pub fn _decode(&self, raw: &[Vec<u8>], shapes: &[[i64; 4]], det_scale: f32) -> Result<(Array2<f32>, Option<Array3<f32>>), Error> {
        let ptrs: Vec<*const u8> = raw.iter().map(|v| v.as_ptr()).collect();
        let lens: Vec<usize> = raw.iter().map(|v| v.len()).collect();
        let ds = [det_scale];
        ffi::check(unsafe {
            ffi::fd_detect_batch_raw(self.fd.raw(), ptrs.as_ptr(), lens.as_ptr(), shapes.as_ptr(), raw.len() as i32, ds.as_ptr(),
                                     self.confidence_threshold, self.iou_threshold)
        })?;
        let mut cap = 0i32;
        ffi::check(unsafe { ffi::fd_ctx_total_anchors(self.fd.raw(), &mut cap) })?;
        let mut det = vec![0f32; cap as usize * 5];
        let mut lmk = vec![0f32; cap as usize * 10];
        let (mut count, mut total) = (0i32, 0i32);
        ffi::check(unsafe { ffi::fd_detect_fetch(self.fd.raw(), &mut count, det.as_mut_ptr(), lmk.as_mut_ptr(), cap, &mut total) })?;
        let m = total as usize;
        det.truncate(m * 5);
        lmk.truncate(m * 10);
        // empty -> (0,5) and (0,5,2) like face_detection.rs:413-419
        Ok((Array2::from_shape_vec((m, 5), det)?, Some(Array3::from_shape_vec((m, 5, 2), lmk)?)))
    }

    /// `call` (:496-513), with `infer` standing for the reference's gRPC round trip (:232-312):
    /// tensor -> (raw_output_contents, shapes).
    pub async fn call<F, Fut>(&self, image: &Mat, _is_debug: Option<bool>, infer: F) -> Result<(Array2<f32>, Option<Array3<f32>>), Error>
    where
        F: FnOnce(Vec<f32>) -> Fut,
        Fut: std::future::Future<Output = Result<(Vec<Vec<u8>>, Vec<[i64; 4]>), Error>>,
    {
        let (tensor, det_scale) = self._preprocess(image)?;
        let (raw, shapes) = infer(tensor).await?;
        self._decode(&raw, &shapes, det_scale)
    }
}

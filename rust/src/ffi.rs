//! src/ffi.rs — `extern "C"` declarations of include/fd_b200.h (the only unsafe surface of the crate).
//! Replaces the commented-out binding in src/rcnn/gpu_nms.rs:9-19.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const FD_MAX_STRIDES: usize = 8;
pub const FD_MAX_ANCHORS: usize = 4;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct fd_config {
    pub image_w: i32, pub image_h: i32,
    pub conf_thr: f32, pub iou_thr: f32,
    pub n_strides: i32, pub strides: [i32; FD_MAX_STRIDES],
    pub num_anchors: i32,
    pub base_anchors: [[[f32; 4]; FD_MAX_ANCHORS]; FD_MAX_STRIDES],
    pub pixel_means: [f32; 3], pub pixel_stds: [f32; 3], pub pixel_scale: f32,
    pub bbox_stds: [f32; 4], pub landmark_std: f32,
    pub crop_w: i32, pub crop_h: i32,
    pub template_landmarks: [[f32; 2]; 5],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fd_frame { pub data: *const u8, pub height: i32, pub width: i32, pub pitch: i32 }
#[repr(C)]
pub struct fd_host_batch_out {
    pub counts: *mut i32, pub det: *mut f32, pub landmarks: *mut f32, pub crops: *mut u8, pub det_scale: *mut f32,
    pub tensor: *mut f32, pub cap_rows: i32, pub total: i32, pub h2d_bytes: i64, pub d2h_bytes: i64,
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct fd_select_params {
    pub margin_center_left_ratio: f32, pub margin_center_right_ratio: f32, pub margin_edge_ratio: f32, pub minimum_face_ratio: f32,
}
pub enum fd_ctx {}

extern "C" {
    pub fn fd_last_error() -> *const c_char;
    pub fn fd_config_default(cfg: *mut fd_config) -> c_int;
    pub fn fd_ctx_create(device_id: c_int, cfg: *const fd_config, out: *mut *mut fd_ctx) -> c_int;
    pub fn fd_ctx_destroy(ctx: *mut fd_ctx);
    pub fn fd_ctx_synchronize(ctx: *mut fd_ctx) -> c_int;
    pub fn fd_dev_alloc(ctx: *mut fd_ctx, bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn fd_dev_free(ctx: *mut fd_ctx, ptr: *mut c_void) -> c_int;
    pub fn fd_memcpy_h2d(ctx: *mut fd_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    pub fn fd_memcpy_d2h(ctx: *mut fd_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> c_int;
    // processing::nms / rcnn
    pub fn fd_nms(ctx: *mut fd_ctx, dets: *const f32, k: c_int, thresh: f32, keep: *mut i32, num_keep: *mut c_int) -> c_int;
    pub fn fd_cpu_nms(ctx: *mut fd_ctx, dets: *const f32, k: c_int, thresh: f32, keep: *mut i32, num_keep: *mut c_int) -> c_int;
    pub fn _nms(keep: *mut i32, num_out: *mut c_int, boxes: *const f32, boxes_num: c_int, boxes_dim: c_int, thresh: f32, device_id: c_int);
    pub fn fd_argsort_descending(ctx: *mut fd_ctx, scores: *const f32, n: c_int, order: *mut i32) -> c_int;
    pub fn fd_anchors_plane(ctx: *mut fd_ctx, h: c_int, w: c_int, stride: c_int, base: *const f32, a: c_int, out: *mut f32) -> c_int;
    pub fn fd_bbox_pred(ctx: *mut fd_ctx, boxes: *const f32, deltas: *const f32, n: c_int, ncols: c_int, out: *mut f32) -> c_int;
    pub fn fd_nonlinear_pred(ctx: *mut fd_ctx, boxes: *const f32, deltas: *const f32, n: c_int, ncols: c_int, out: *mut f32) -> c_int;
    pub fn fd_landmark_pred(ctx: *mut fd_ctx, boxes: *const f32, deltas: *const f32, n: c_int, out: *mut f32) -> c_int;
    pub fn fd_clip_boxes(ctx: *mut fd_ctx, boxes: *mut f32, rows: c_int, cols: c_int, im_h: c_int, im_w: c_int) -> c_int;
    pub fn fd_clip_points(ctx: *mut fd_ctx, pts: *mut f32, rows: c_int, cols: c_int, im_h: c_int, im_w: c_int) -> c_int;
    pub fn fd_bbox_overlaps(ctx: *mut fd_ctx, boxes: *const f32, n: c_int, query: *const f32, k: c_int, out: *mut f32) -> c_int;
    // detector / aligner, single image, host buffers
    pub fn fd_preprocess(ctx: *mut fd_ctx, img: *const u8, h: c_int, w: c_int, pitch: c_int, out_nchw: *mut f32, det_scale: *mut f32) -> c_int;
    pub fn fd_detect(ctx: *mut fd_ctx, heads: *const *const f32, n_heads: c_int, det_scale: f32, conf_thr: f32, iou_thr: f32,
                     det: *mut f32, landmarks: *mut f32, cap: c_int, num_det: *mut c_int) -> c_int;
    pub fn fd_align(ctx: *mut fd_ctx, img: *const u8, h: c_int, w: c_int, pitch: c_int, landmarks: *const f32, crop: *mut u8, m_out: *mut f64) -> c_int;
    // batched pipeline
    pub fn fd_preprocess_batch(ctx: *mut fd_ctx, frames: *const fd_frame, b: c_int, out_nchw_dev: *mut f32, det_scale_host: *mut f32) -> c_int;
    pub fn fd_detect_batch(ctx: *mut fd_ctx, heads_dev: *const *const f32, n_heads: c_int, b: c_int, det_scale_host: *const f32, conf_thr: f32, iou_thr: f32) -> c_int;
    pub fn fd_detect_fetch(ctx: *mut fd_ctx, counts: *mut i32, det: *mut f32, landmarks: *mut f32, cap_rows: c_int, total: *mut c_int) -> c_int;
    pub fn fd_align_detections(ctx: *mut fd_ctx, frames: *const fd_frame, b: c_int, crops_dev: *mut u8, cap_faces: c_int, m_dev: *mut f64, ok_dev: *mut u8) -> c_int;
    // SURVEY 8(f) "next" rows: model preprocessors on the crops, Triton raw_output_contents, FaceSelection
    pub fn fd_crops_to_tensor(ctx: *mut fd_ctx, crops_dev: *const u8, f: c_int, in_h: c_int, in_w: c_int, out_h: c_int, out_w: c_int,
                              mean_rgb: *const f32, mul_rgb: *const f32, out_nchw_dev: *mut f32, use_detect_count: c_int) -> c_int;
    pub fn fd_model_preprocess(ctx: *mut fd_ctx, img: *const u8, h: c_int, w: c_int, pitch: c_int, out_h: c_int, out_w: c_int,
                               mean_rgb: *const f32, mul_rgb: *const f32, out_nchw: *mut f32) -> c_int;
    pub fn fd_detect_batch_raw(ctx: *mut fd_ctx, raw: *const *const u8, nbytes: *const usize, shape: *const [i64; 4], n_heads: c_int,
                               det_scale_host: *const f32, conf_thr: f32, iou_thr: f32) -> c_int;
    pub fn fd_select_params_default(p: *mut fd_select_params) -> c_int;
    pub fn fd_face_selection(ctx: *mut fd_ctx, img_h: c_int, img_w: c_int, face_boxes: *const f32, key_points: *const f32, m: c_int,
                             is_enroll: c_int, params: *const fd_select_params, box_index: *mut c_int, kp_index: *mut c_int) -> c_int;
    pub fn fd_select_detections(ctx: *mut fd_ctx, frames: *const fd_frame, b: c_int, is_enroll: c_int, params: *const fd_select_params,
                                sel_host: *mut i32) -> c_int;
    pub fn fd_align_selected(ctx: *mut fd_ctx, frames: *const fd_frame, b: c_int, crops_dev: *mut u8, m_dev: *mut f64, ok_dev: *mut u8) -> c_int;
    pub fn fd_pipeline_host(ctx: *mut fd_ctx, frames: *const fd_frame, b: c_int, heads_host: *const *const f32, n_heads: c_int,
                            conf_thr: f32, iou_thr: f32, out: *mut fd_host_batch_out) -> c_int;
}

/// Maps an fd_status to the crate's error type (the reference bubbles `anyhow::Error`, e.g. face_detection.rs:135-138).
pub fn check(rc: c_int) -> anyhow::Result<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(fd_last_error()) }.to_string_lossy().into_owned();
    Err(anyhow::anyhow!("fd_b200 error {}: {}", rc, msg))
}

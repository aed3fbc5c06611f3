//! Replacement body for byte_data_to_opencv (src/utils/utils.rs:8-52): same signature, the decode runs through libfd_b200
//! (device Huffman stage for streams with restart markers, host otherwise; CUDA IDCT / upsampling / colour conversion),
//! bit-identical to cv::imdecode for baseline 3-component JPEG.  Streams the library does not cover (progressive, grayscale,
//! 4-channel PNG, ...) still go to OpenCV exactly as in the reference (:17-49), so behaviour is unchanged for them.
use anyhow::{Error, Result};
use opencv::core::{Mat, Scalar, CV_8UC3};
use opencv::prelude::MatTrait;
use crate::{ctx::with_ctx, ffi};

pub fn byte_data_to_opencv(im_bytes: &[u8]) -> Result<Mat, Error> {
    let (mut h, mut w, mut ss) = (0i32, 0i32, 0i32);
    let covered = unsafe { ffi::fd_jpeg_info(im_bytes.as_ptr(), im_bytes.len(), &mut h, &mut w, &mut ss) } == ffi::FD_OK;
    if !covered {
        return crate::utils::utils_opencv::byte_data_to_opencv(im_bytes); // the reference's own body, moved aside unchanged
    }
    let mut img = Mat::new_rows_cols_with_default(h, w, CV_8UC3, Scalar::all(0.0))?;
    let pitch = w * 3;
    with_ctx(|c| ffi::check(unsafe { ffi::fd_imdecode(c, im_bytes.as_ptr(), im_bytes.len(), img.data_mut(), pitch) }))?;
    Ok(img)
}

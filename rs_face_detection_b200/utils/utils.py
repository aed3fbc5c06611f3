"""utils::utils (src/utils/utils.rs) — the functions of the reference's utility module that sit on the hot path.

byte_data_to_opencv (utils.rs:8-52): encoded image bytes -> BGR image.  The reference calls cv::imdecode(IMREAD_UNCHANGED)
and normalises 4- and 2-channel results; here a baseline 3-component JPEG is decoded by the library (Huffman stage on the
device for streams with restart markers, on the host otherwise; IDCT / upsampling / colour conversion in CUDA kernels),
bit-identical to cv2.imdecode.  Other formats raise FdError: they stay with OpenCV in the reference.
u8_to_f32_vec (utils.rs:126-132) has no counterpart: fd_detect_batch_raw reads the little-endian bytes in place.
"""
from .. import default_context


def byte_data_to_opencv(im_bytes, ctx=None):
    """-> (h, w, 3) BGR u8 numpy array (the reference returns an OpenCV Mat of the same layout)"""
    return (ctx or default_context()).imdecode(im_bytes)

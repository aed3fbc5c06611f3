"""pipeline::module::face_alignment::FaceAlignment (src/pipeline/module/face_alignment.rs:14-141)."""
import numpy as np

from .. import Context
from ..ffi import default_config


class FaceAlignment:
    def __init__(self, image_size=(112, 112), standard_landmarks=None, ctx=None, device=0):
        cfg = default_config()
        cfg.crop_w, cfg.crop_h = image_size
        if standard_landmarks is not None:
            for i, v in enumerate(np.asarray(standard_landmarks, np.float32).ravel()):
                cfg.template_landmarks[i] = v
        self.ctx = ctx or Context(device, cfg)

    def call(self, img, bbox=None, landmarks=None):
        """face_alignment.rs:27-141: the similarity warp (:50-59, :119-126) or, when estimateAffinePartial2D returns an
        empty matrix, the margin-44 bbox crop + resize (:64-116).  Raises FdError where the reference returns Err
        (landmarks=None: OpenCV asserts on the empty Mat; fallback ROI outside the image: Mat::roi)."""
        crop, _ = self.ctx.align(img, landmarks, bbox)
        return crop

"""pipeline::module::face_alignment::FaceAlignment (src/pipeline/module/face_alignment.rs:14-141)."""
import numpy as np

from .. import Context, FdError
from ..ffi import FD_ERR_ESTIMATE, default_config


class FaceAlignment:
    def __init__(self, image_size=(112, 112), standard_landmarks=None, ctx=None, device=0):
        cfg = default_config()
        cfg.crop_w, cfg.crop_h = image_size
        if standard_landmarks is not None:
            for i, v in enumerate(np.asarray(standard_landmarks, np.float32).ravel()):
                cfg.template_landmarks[i] = v
        self.ctx = ctx or Context(device, cfg)

    def call(self, img, bbox=None, landmarks=None):
        """Main branch (:50-59, :119-126).  Where the reference would take its bbox-crop fallback (:64-116, reachable
        only when the estimate is empty) this raises FdError(FD_ERR_ESTIMATE): that branch is out of scope."""
        if landmarks is None:
            raise FdError(FD_ERR_ESTIMATE, "landmarks=None: the reference's bbox-crop fallback is out of scope")
        crop, _ = self.ctx.align(img, landmarks)
        return crop

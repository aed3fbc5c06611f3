"""pipeline::module::face_selection::FaceSelection (src/pipeline/module/face_selection.rs:5-189)."""
import numpy as np

from .. import Context


class FaceSelection:
    def __init__(self, margin_center_left_ratio=0.3, margin_center_right_ratio=0.3, margin_edge_ratio=0.1, minimum_face_ratio=0.0075,
                 ctx=None, device=0):
        # defaults: FaceSelectionConfig::new (face_pipeline/config.rs:107-116)
        self.params = (margin_center_left_ratio, margin_center_right_ratio, margin_edge_ratio, minimum_face_ratio)
        self.ctx = ctx or Context(device)

    def call(self, img, face_boxes, key_points=None, is_enroll=False):
        """-> (Option<Array1<f32>> box (5,), Option<Array2<f32>> key points (5,2))   (face_selection.rs:72)"""
        fb = np.ascontiguousarray(face_boxes, np.float32).reshape(-1, 5)
        bi, ki = self.ctx.face_selection(img.shape[:2], fb, key_points, is_enroll, self.params)
        box = fb[bi].copy() if bi >= 0 else None
        kps = np.asarray(key_points, np.float32).reshape(-1, 5, 2)[ki].copy() if (ki >= 0 and key_points is not None) else None
        return box, kps

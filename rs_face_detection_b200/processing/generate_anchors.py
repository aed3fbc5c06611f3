"""processing::generate_anchors (src/processing/generate_anchors.rs): init-time host tables, computed inside the library."""
from ..ffi import generate_anchors, generate_anchors2, generate_anchors_fpn, generate_anchors_fpn2  # noqa: F401

RETINAFACE_ANCHOR_CFG = {  # face_detection.rs:55-80
    "32": {"base_size": 16, "ratios": [1.0], "scales": [32.0, 16.0], "allowed_border": 9999},
    "16": {"base_size": 16, "ratios": [1.0], "scales": [8.0, 4.0], "allowed_border": 9999},
    "8": {"base_size": 16, "ratios": [1.0], "scales": [2.0, 1.0], "allowed_border": 9999},
}

"""Mirror of the reference's `processing` module (src/processing/mod.rs:2-4)."""
from . import bbox_transform, generate_anchors, nms  # noqa: F401

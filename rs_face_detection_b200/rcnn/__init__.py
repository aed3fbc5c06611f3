"""Mirror of the reference's `rcnn` module (src/rcnn/mod.rs)."""
from . import anchors, bbox, cpu_nms, gpu_nms  # noqa: F401

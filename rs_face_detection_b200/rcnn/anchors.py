"""rcnn::anchors (src/rcnn/anchors.rs:3-21)."""
from .. import default_context


def anchors(height, width, stride, base_anchors, ctx=None):
    return (ctx or default_context()).anchors_plane(height, width, stride, base_anchors)

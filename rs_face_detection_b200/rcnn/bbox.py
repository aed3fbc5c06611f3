"""rcnn::bbox (src/rcnn/bbox.rs:4-30)."""
from .. import default_context


def bbox_overlaps(boxes, query_boxes, ctx=None):
    return (ctx or default_context()).bbox_overlaps(boxes, query_boxes)

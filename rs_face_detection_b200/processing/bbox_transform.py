"""processing::bbox_transform (src/processing/bbox_transform.rs).  Every function runs on the GPU."""
from .. import default_context


def bbox_overlaps_py(boxes, query_boxes, ctx=None):          # bbox_transform.rs:2-24
    return (ctx or default_context()).bbox_overlaps(boxes, query_boxes)


def clip_boxes(boxes, im_shape, ctx=None):                   # :27-45 (returns the clipped copy)
    return (ctx or default_context()).clip_boxes(boxes, im_shape)


def clip_points(points, im_shape, ctx=None):                 # :47-65
    return (ctx or default_context()).clip_points(points, im_shape)


def nonlinear_transform(ex_rois, gt_rois, ctx=None):         # :67-88
    assert len(ex_rois) == len(gt_rois), "inconsistent rois number"
    return (ctx or default_context()).nonlinear_transform(ex_rois, gt_rois)


def nonlinear_pred(boxes, box_deltas, ctx=None):             # :90-120
    return (ctx or default_context()).nonlinear_pred(boxes, box_deltas)


def landmark_pred(boxes, point_deltas, ctx=None):            # :123-160
    return (ctx or default_context()).landmark_pred(boxes, point_deltas)


def iou_pred(boxes, box_deltas, num_classes, ctx=None):      # :162-186
    return (ctx or default_context()).iou_pred(boxes, box_deltas, num_classes)

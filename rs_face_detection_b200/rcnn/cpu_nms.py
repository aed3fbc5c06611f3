"""rcnn::cpu_nms (src/rcnn/cpu_nms.rs:10-55): the `>=`-threshold variant — computed on the GPU here."""
from .. import default_context


def cpu_nms(dets, thresh, ctx=None):
    return (ctx or default_context()).cpu_nms(dets, thresh)

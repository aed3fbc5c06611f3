"""processing::nms (src/processing/nms.rs)."""
from .. import default_context


def nms(dets, thresh, ctx=None):
    """nms(&Array2<f32>, f32) -> Vec<usize> (nms.rs:3-65): indices into dets, pick order.  Runs on the GPU."""
    return (ctx or default_context()).nms(dets, thresh)

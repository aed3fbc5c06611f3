"""rcnn::gpu_nms (src/rcnn/gpu_nms.rs:21-49, commented out in the reference): sort, call `_nms`, map back."""
import ctypes as C

import numpy as np

from .. import ffi


def gpu_nms(dets, thresh, device_id=0):
    """Same contract as the reference's intended wrapper, through the literal C symbol `_nms` (gpu_nms.hpp:6-8)."""
    dets = np.ascontiguousarray(dets, np.float32).reshape(-1, 5)
    order = np.argsort(-dets[:, 4], kind="stable")
    sorted_dets = np.ascontiguousarray(dets[order])
    keep = np.zeros(max(len(dets), 1), np.int32)
    num_out = C.c_int(0)
    ffi.load()._nms(keep.ctypes.data_as(ffi.c_i32p), C.byref(num_out), sorted_dets.ctypes.data_as(ffi.c_f32p),
                    len(dets), 5, C.c_float(thresh), int(device_id))
    return order[keep[:num_out.value]]

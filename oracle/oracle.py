"""ctypes loader for the CPU oracle (oracle/fd_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (rs_face_detection_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfd_oracle.so")

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present) with oracle/Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("fd_oracle.c", "fd_jpeg_oracle.c", "Makefile")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "libfd_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class DetCfg(C.Structure):
    """Mirror of fdo_det_cfg."""
    _fields_ = [
        ("image_w", C.c_int), ("image_h", C.c_int),
        ("n_strides", C.c_int),
        ("strides", C.c_int * 8),
        ("num_anchors", C.c_int),
        ("base_anchors", C.c_float * (8 * 4 * 4)),
        ("bbox_stds", C.c_float * 4),
        ("landmark_std", C.c_float),
        ("conf_thr", C.c_float), ("iou_thr", C.c_float),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.fdo_nms_pairs.restype = C.c_longlong
        _lib.fdo_preprocess_letterbox.restype = C.c_float
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------- anchors
def ratio_enum(anchor, ratios):
    anchor, ratios = _f32(anchor), _f32(ratios)
    out = np.empty((len(ratios), 4), np.float32)
    lib().fdo_ratio_enum(_p(anchor, c_f32p), _p(ratios, c_f32p), len(ratios), _p(out, c_f32p))
    return out


def scale_enum(anchor, scales):
    anchor, scales = _f32(anchor), _f32(scales)
    out = np.empty((len(scales), 4), np.float32)
    lib().fdo_scale_enum(_p(anchor, c_f32p), _p(scales, c_f32p), len(scales), _p(out, c_f32p))
    return out


def generate_anchors2(base_size, ratios, scales, stride=0, dense_anchor=False):
    ratios, scales = _f32(ratios), _f32(scales)
    out = np.empty((len(ratios) * len(scales) * 2, 4), np.float32)
    n = lib().fdo_generate_anchors2(int(base_size), _p(ratios, c_f32p), len(ratios), _p(scales, c_f32p),
                                    len(scales), int(stride), int(dense_anchor), _p(out, c_f32p))
    return out[:n].copy()


def generate_anchors(base_size, ratios, scales):
    return generate_anchors2(base_size, ratios, scales, 0, False)


def generate_anchors_fpn(base_size, ratios, scales):
    bs = np.ascontiguousarray(base_size, np.int32)
    ratios, scales = _f32(ratios), _f32(scales)
    out = np.empty((len(bs), 4), np.float32)
    lib().fdo_generate_anchors_fpn(_p(bs, c_i32p), _p(ratios, c_f32p), _p(scales, c_f32p), len(bs), _p(out, c_f32p))
    return [out[i:i + 1].copy() for i in range(len(bs))]


def generate_anchors_fpn2_retinaface(dense_anchor=False):
    out = np.empty((3, 2, 4), np.float32)
    lib().fdo_generate_anchors_fpn2_retinaface(int(dense_anchor), _p(out, c_f32p))
    return out


def anchors_plane(height, width, stride, base):
    base = _f32(base)
    A = base.shape[0]
    out = np.empty((height, width, A, 4), np.float32)
    lib().fdo_anchors_plane(height, width, stride, _p(base, c_f32p), A, _p(out, c_f32p))
    return out


# ---------------------------------------------------------------- bbox transforms
def nonlinear_pred(boxes, deltas):
    boxes, deltas = _f32(boxes), _f32(deltas)
    out = np.empty_like(deltas)
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), np.float32)
    lib().fdo_nonlinear_pred(_p(boxes, c_f32p), _p(deltas, c_f32p), boxes.shape[0], deltas.shape[1], _p(out, c_f32p))
    return out


def bbox_pred(boxes, deltas):
    boxes, deltas = _f32(boxes), _f32(deltas)
    out = np.empty_like(deltas)
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), np.float32)
    lib().fdo_bbox_pred(_p(boxes, c_f32p), _p(deltas, c_f32p), boxes.shape[0], deltas.shape[1], _p(out, c_f32p))
    return out


def landmark_pred(boxes, deltas):
    boxes, deltas = _f32(boxes), _f32(deltas)
    shape = deltas.shape
    out = np.empty_like(deltas)
    if boxes.shape[0] == 0:
        return np.zeros((0,) + tuple(shape[1:]), np.float32)
    lib().fdo_landmark_pred(_p(boxes, c_f32p), _p(deltas, c_f32p), boxes.shape[0], _p(out, c_f32p))
    return out.reshape(shape)


def clip_boxes(boxes, im_shape):
    boxes = _f32(boxes).copy()
    lib().fdo_clip_boxes(_p(boxes, c_f32p), boxes.shape[0], boxes.shape[1], int(im_shape[0]), int(im_shape[1]))
    return boxes


def clip_points(points, im_shape):
    points = _f32(points).copy()
    lib().fdo_clip_points(_p(points, c_f32p), points.shape[0], points.shape[1], int(im_shape[0]), int(im_shape[1]))
    return points


def iou_pred(boxes, deltas, num_classes):
    boxes, deltas = _f32(boxes), _f32(deltas)
    out = np.empty_like(deltas)
    lib().fdo_iou_pred(_p(boxes, c_f32p), _p(deltas, c_f32p), boxes.shape[0], deltas.shape[1], num_classes, _p(out, c_f32p))
    return out


def nonlinear_transform(ex, gt):
    ex, gt = _f32(ex), _f32(gt)
    out = np.empty((ex.shape[0], 4), np.float32)
    lib().fdo_nonlinear_transform(_p(ex, c_f32p), _p(gt, c_f32p), ex.shape[0], _p(out, c_f32p))
    return out


def bbox_overlaps(boxes, query):
    boxes, query = _f32(boxes), _f32(query)
    out = np.empty((boxes.shape[0], query.shape[0]), np.float32)
    lib().fdo_bbox_overlaps(_p(boxes, c_f32p), boxes.shape[0], _p(query, c_f32p), query.shape[0], _p(out, c_f32p))
    return out


# ---------------------------------------------------------------- sort / nms
def argsort_descending(scores):
    scores = _f32(scores)
    order = np.empty(len(scores), np.int32)
    rc = lib().fdo_argsort_descending(_p(scores, c_f32p), len(scores), _p(order, c_i32p))
    if rc != 0:
        raise ValueError("NaN score (the reference panics in argsort_descending, utils.rs:92)")
    return order


def _nms_like(fn, dets, thresh):
    dets = _f32(dets).reshape(-1, 5)
    keep = np.empty(max(dets.shape[0], 1), np.int32)
    n = fn(_p(dets, c_f32p), dets.shape[0], C.c_float(thresh), _p(keep, c_i32p))
    if n < 0:
        raise ValueError("NaN score")
    return keep[:n].copy()


def nms(dets, thresh):
    """processing::nms::nms (nms.rs:3-65)."""
    return _nms_like(lib().fdo_nms, dets, thresh)


def cpu_nms(dets, thresh):
    """rcnn::cpu_nms::cpu_nms (cpu_nms.rs:10-55)."""
    return _nms_like(lib().fdo_cpu_nms, dets, thresh)


def nms_pairs(dets, thresh):
    dets = _f32(dets).reshape(-1, 5)
    return int(lib().fdo_nms_pairs(_p(dets, c_f32p), dets.shape[0], C.c_float(thresh)))


def nms_sorted(boxes, thresh):
    """Contract of the reference C symbol _nms (gpu_nms.hpp:6-8): boxes pre-sorted."""
    boxes = _f32(boxes)
    keep = np.empty(max(boxes.shape[0], 1), np.int32)
    n = lib().fdo_nms_sorted(_p(boxes, c_f32p), boxes.shape[0], boxes.shape[1], C.c_float(thresh), _p(keep, c_i32p))
    return keep[:n].copy()


# ---------------------------------------------------------------- preprocess
def resize_linear(img, dsize):
    """cv::resize(img, dsize=(w,h), INTER_LINEAR) for u8 HxWx3."""
    img = np.ascontiguousarray(img, np.uint8)
    dw, dh = dsize
    out = np.empty((dh, dw, 3), np.uint8)
    lib().fdo_resize_linear_u8c3(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0],
                                 _p(out, c_u8p), dh, dw, dw * 3)
    return out


def letterbox_geometry(h, w, size=(640, 640)):
    nw, nh, sc = C.c_int(), C.c_int(), C.c_float()
    lib().fdo_letterbox_geometry(h, w, size[0], size[1], C.byref(nw), C.byref(nh), C.byref(sc))
    return nw.value, nh.value, np.float32(sc.value)


def preprocess_letterbox(img, size=(640, 640)):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty((size[1], size[0], 3), np.uint8)
    sc = lib().fdo_preprocess_letterbox(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0],
                                        size[0], size[1], _p(out, c_u8p))
    return out, np.float32(sc)


def to_tensor(det_img, pixel_scale=1.0, means=(0, 0, 0), stds=(1, 1, 1)):
    det_img = np.ascontiguousarray(det_img, np.uint8)
    means, stds = _f32(means), _f32(stds)
    out = np.empty((1, 3, det_img.shape[0], det_img.shape[1]), np.float32)
    lib().fdo_to_tensor(_p(det_img, c_u8p), det_img.shape[0], det_img.shape[1], C.c_float(pixel_scale),
                        _p(means, c_f32p), _p(stds, c_f32p), _p(out, c_f32p))
    return out


# ---------------------------------------------------------------- detect (post-CNN half)
def make_det_cfg(conf_thr=0.7, iou_thr=0.45, image_size=(640, 640), strides=(32, 16, 8), base_anchors=None,
                 bbox_stds=(1, 1, 1, 1), landmark_std=1.0):
    cfg = DetCfg()
    cfg.image_w, cfg.image_h = image_size
    cfg.n_strides = len(strides)
    for i, s in enumerate(strides):
        cfg.strides[i] = s
    if base_anchors is None:
        base_anchors = generate_anchors_fpn2_retinaface(False)
    base_anchors = np.asarray(base_anchors, np.float32)
    cfg.num_anchors = base_anchors.shape[1]
    buf = np.zeros((8, 4, 4), np.float32)
    buf[:base_anchors.shape[0], :base_anchors.shape[1]] = base_anchors
    for i, v in enumerate(buf.ravel()):
        cfg.base_anchors[i] = v
    for i in range(4):
        cfg.bbox_stds[i] = bbox_stds[i]
    cfg.landmark_std = landmark_std
    cfg.conf_thr, cfg.iou_thr = conf_thr, iou_thr
    return cfg


def _head_ptrs(heads):
    heads = [_f32(h) for h in heads]
    arr = (c_f32p * len(heads))(*[_p(h, c_f32p) for h in heads])
    fh = (C.c_int * (len(heads) // 3))(*[heads[3 * s + 1].shape[-2] for s in range(len(heads) // 3)])
    fw = (C.c_int * (len(heads) // 3))(*[heads[3 * s + 1].shape[-1] for s in range(len(heads) // 3)])
    return heads, arr, fh, fw


def detect_post(cfg, heads, det_scale):
    """heads: 9 arrays for ONE image, (C,H,W) or (1,C,H,W).  Returns det (M,5), landmarks (M,5,2), K."""
    heads, arr, fh, fw = _head_ptrs(heads)
    cap = sum(int(fh[s]) * int(fw[s]) for s in range(len(heads) // 3)) * cfg.num_anchors
    det = np.empty((cap, 5), np.float32)
    lmk = np.empty((cap, 10), np.float32)
    K = C.c_int()
    M = lib().fdo_detect_post(C.byref(cfg), arr, fh, fw, C.c_float(det_scale), _p(det, c_f32p), _p(lmk, c_f32p),
                              cap, C.byref(K))
    if M < 0:
        raise ValueError("NaN score")
    return det[:M].copy(), lmk[:M].reshape(M, 5, 2).copy(), K.value


def decode_candidates(cfg, heads):
    heads, arr, fh, fw = _head_ptrs(heads)
    cap = sum(int(fh[s]) * int(fw[s]) for s in range(len(heads) // 3)) * cfg.num_anchors
    box = np.empty((cap, 4), np.float32)
    score = np.empty(cap, np.float32)
    lmk = np.empty((cap, 10), np.float32)
    idx = np.empty(cap, np.int32)
    K = lib().fdo_decode_candidates(C.byref(cfg), arr, fh, fw, _p(box, c_f32p), _p(score, c_f32p), _p(lmk, c_f32p),
                                    _p(idx, c_i32p), cap)
    assert K >= 0
    return box[:K].copy(), score[:K].copy(), lmk[:K].copy(), idx[:K].copy()


# ---------------------------------------------------------------- align
ARCFACE_TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                             [41.5493, 92.3655], [70.7299, 92.2041]], np.float32)  # config.rs:46-52


def estimate_affine_partial_2d(src, dst):
    """cv::estimateAffinePartial2D(src, dst, LMEDS, 3.0, 2000, 0.99, 10) -> (M 2x3 f64 | None, inliers u8)."""
    src, dst = _f32(src).reshape(-1, 2), _f32(dst).reshape(-1, 2)
    M = np.empty(6, np.float64)
    inl = np.zeros(src.shape[0], np.uint8)
    ok = lib().fdo_estimate_affine_partial_2d_lmeds(_p(src, c_f32p), _p(dst, c_f32p), src.shape[0], _p(M, c_f64p),
                                                    _p(inl, c_u8p))
    return (M.reshape(2, 3) if ok else None), inl


def warp_affine(img, M, dsize=(112, 112)):
    img = np.ascontiguousarray(img, np.uint8)
    M = np.ascontiguousarray(M, np.float64).reshape(6)
    dw, dh = dsize
    out = np.empty((dh, dw, 3), np.uint8)
    lib().fdo_warp_affine_u8c3(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0], _p(M, c_f64p),
                               _p(out, c_u8p), dh, dw, dw * 3)
    return out


def align_face(img, lmk, template=ARCFACE_TEMPLATE, dsize=(112, 112), bbox=None, with_mode=False):
    """FaceAlignment::call (face_alignment.rs:27-141).  -> (crop, M) — M is None when the bbox-crop fallback (:64-116) was
    taken, both are None where the reference returns Err; with_mode appends the mode (1 warp, 2 fallback, 0 Err)."""
    img = np.ascontiguousarray(img, np.uint8)
    lmk, template = _f32(lmk).reshape(10), _f32(template).reshape(10)
    bb = _f32(bbox).reshape(-1)[:4].copy() if bbox is not None else None
    dw, dh = dsize
    out = np.empty((dh, dw, 3), np.uint8)
    M = np.empty(6, np.float64)
    mode = lib().fdo_align_face(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0], _p(bb, c_f32p) if bb is not None else None,
                                _p(lmk, c_f32p), _p(template, c_f32p), dw, dh, _p(out, c_u8p), _p(M, c_f64p))
    res = (out, M.reshape(2, 3)) if mode == 1 else ((out, None) if mode == 2 else (None, None))
    return res + (mode,) if with_mode else res


def align_fallback(img, bbox=None, dsize=(112, 112)):
    """The empty-transform branch of FaceAlignment::call on its own (face_alignment.rs:64-116); None where it errs."""
    img = np.ascontiguousarray(img, np.uint8)
    bb = _f32(bbox).reshape(-1)[:4].copy() if bbox is not None else None
    dw, dh = dsize
    out = np.empty((dh, dw, 3), np.uint8)
    ok = lib().fdo_align_fallback(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0], _p(bb, c_f32p) if bb is not None else None,
                                  dw, dh, _p(out, c_u8p))
    return out if ok else None


MODEL_NORMS = {  # (mean_rgb, mul_rgb) of the three post-align models
    "face_extraction": ((127.5, 127.5, 127.5), (0.0078125,) * 3),                                   # face_extraction.rs:69
    "face_quality": ((123.675, 116.28, 103.53), (0.01712475, 0.017507, 0.01742919)),               # face_quality.rs:43-44,93
    "face_quality_assessment": ((127.5, 127.5, 127.5), (0.00784313725,) * 3),                       # face_quality_assessment.rs:78
}


def model_preprocess(img, out_size, mean_rgb, mul_rgb):
    """resize -> BGR2RGB -> (p - mean) * mul -> (3, out_h, out_w)"""
    img = np.ascontiguousarray(img, np.uint8)
    mean_rgb, mul_rgb = _f32(mean_rgb), _f32(mul_rgb)
    ow, oh = out_size
    out = np.empty((3, oh, ow), np.float32)
    lib().fdo_model_preprocess(_p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0], oh, ow, _p(mean_rgb, c_f32p),
                               _p(mul_rgb, c_f32p), _p(out, c_f32p))
    return out


SELECT_DEFAULTS = (0.3, 0.3, 0.1, 0.0075)   # FaceSelectionConfig::new (face_pipeline/config.rs:107-116)


def face_selection(img_hw, face_boxes, key_points=None, is_enroll=False, params=SELECT_DEFAULTS):
    """FaceSelection::call (face_selection.rs:72-189) -> (box row index or -1, key-point row index or -1)."""
    fb = _f32(face_boxes).reshape(-1, 5)
    prm = _f32(params)
    bi, ki = C.c_int(-1), C.c_int(-1)
    lib().fdo_face_selection(int(img_hw[0]), int(img_hw[1]), _p(fb, c_f32p), len(fb), int(key_points is not None), int(bool(is_enroll)),
                             _p(prm, c_f32p), C.byref(bi), C.byref(ki))
    return bi.value, ki.value


def pipeline_frame(cfg, img, heads, template=ARCFACE_TEMPLATE, crop=(112, 112), pixel_scale=1.0,
                   means=(0, 0, 0), stds=(1, 1, 1), bufs=None):
    """Whole reference CPU path for one frame.  Returns tensor, det, landmarks, crops."""
    img = np.ascontiguousarray(img, np.uint8)
    heads, arr, fh, fw = _head_ptrs(heads)
    cap = sum(int(fh[s]) * int(fw[s]) for s in range(len(heads) // 3)) * cfg.num_anchors
    means, stds, template = _f32(means), _f32(stds), _f32(template).reshape(10)
    if bufs is None:
        bufs = (np.empty((1, 3, cfg.image_h, cfg.image_w), np.float32), np.empty((cap, 5), np.float32),
                np.empty((cap, 10), np.float32), np.empty((cap, crop[1], crop[0], 3), np.uint8))
    tensor, det, lmk, crops = bufs
    M = lib().fdo_pipeline_frame(C.byref(cfg), _p(img, c_u8p), img.shape[0], img.shape[1], img.strides[0], arr, fh, fw,
                                 _p(means, c_f32p), _p(stds, c_f32p), C.c_float(pixel_scale), _p(template, c_f32p),
                                 crop[0], crop[1], _p(tensor, c_f32p), _p(det, c_f32p), _p(lmk, c_f32p), cap,
                                 _p(crops, c_u8p))
    if M < 0:
        raise ValueError("NaN score")
    return tensor, det[:M], lmk[:M].reshape(M, 5, 2), crops[:M]


# ---------------------------------------------------------------- N4: JPEG decode (utils.rs:8-52 -> cv::imdecode)
def jpeg_info(data):
    """-> (h, w, subsampling) with subsampling 11 (4:4:4), 21 (4:2:2) or 22 (4:2:0); raises ValueError on unsupported streams"""
    buf = np.frombuffer(bytes(data), np.uint8)
    h, w, ss = C.c_int(0), C.c_int(0), C.c_int(0)
    rc = lib().fdo_jpeg_info(_p(buf, c_u8p), C.c_size_t(len(buf)), C.byref(h), C.byref(w), C.byref(ss))
    if rc != 0:
        raise ValueError("jpeg oracle: error %d" % rc)
    return h.value, w.value, ss.value


def jpeg_decode(data):
    """cv::imdecode(bytes, IMREAD_UNCHANGED) for a baseline 3-component JPEG -> (h, w, 3) BGR u8"""
    buf = np.frombuffer(bytes(data), np.uint8)
    h, w, _ = jpeg_info(data)
    out = np.empty((h, w, 3), np.uint8)
    rc = lib().fdo_jpeg_decode_bgr(_p(buf, c_u8p), C.c_size_t(len(buf)), _p(out, c_u8p), w * 3)
    if rc != 0:
        raise ValueError("jpeg oracle: error %d" % rc)
    return out

"""A second, independent pin for the oracle's Rust-only arithmetic (the reference holds no expected values for these
rows and there is no Rust toolchain here): plain numpy float32 restatements written from the Rust sources — array expressions
in the order the reference evaluates them, every operation rounded to f32 — against the C oracle on random inputs with ties,
degenerate boxes and thresholds on both sides of the decision.  CPU only."""
import numpy as np
import pytest

F = np.float32


def np_nms(dets, thr):
    """processing/nms.rs:3-65, statement by statement (stable descending sort, `ovr <= thresh` survives)."""
    dets = np.asarray(dets, F)
    scores = dets[:, 4]
    # sort_by(|a, b| scores[b].partial_cmp(&scores[a])) is a stable sort by descending score
    order = sorted(range(len(dets)), key=lambda i: -float(scores[i])) if not np.isnan(scores).any() else None
    assert order is not None
    order = np.array(order, np.int64)
    keep = []
    with np.errstate(divide="ignore", invalid="ignore"):
        while len(order):
            i = order[0]
            keep.append(int(i))
            rest = order[1:]
            xx1 = np.maximum(dets[i, 0], dets[rest, 0])
            yy1 = np.maximum(dets[i, 1], dets[rest, 1])
            xx2 = np.minimum(dets[i, 2], dets[rest, 2])
            yy2 = np.minimum(dets[i, 3], dets[rest, 3])
            w = np.maximum(F(0), (xx2 - xx1 + F(1)).astype(F))
            h = np.maximum(F(0), (yy2 - yy1 + F(1)).astype(F))
            inter = (w * h).astype(F)
            area_i = F((dets[i, 2] - dets[i, 0] + F(1)) * (dets[i, 3] - dets[i, 1] + F(1)))
            area_o = ((dets[rest, 2] - dets[rest, 0] + F(1)) * (dets[rest, 3] - dets[rest, 1] + F(1))).astype(F)
            ovr = (inter / ((area_i + area_o).astype(F) - inter).astype(F)).astype(F)
            order = rest[ovr <= F(thr)]          # NaN <= thr is false: the box is removed (nms.rs:58)
    return np.array(keep, np.int64)


def _dets(n, seed, canvas=300.0, levels=None, degenerate=False):
    rng = np.random.default_rng(seed)
    s = rng.uniform(6, 90, n)
    x, y = rng.uniform(0, canvas, n), rng.uniform(0, canvas, n)
    sc = rng.uniform(0.02, 1, n)
    if levels:
        sc = np.round(sc * levels) / levels
    d = np.stack([x, y, x + s, y + s * rng.uniform(0.7, 1.3, n), sc], 1).astype(F)
    d[:, :4] = np.round(d[:, :4] / 2) * 2          # quantised: overlaps land exactly on simple thresholds
    if degenerate:
        d[::7, 2] = d[::7, 0] - 1.0                # zero width: area 0, 0/0 = NaN overlap
        d[::11, 3] = d[::11, 1] - 5.0              # negative height
    return d


@pytest.mark.parametrize("seed,n,levels,deg", [(1, 300, None, False), (2, 800, 20, False), (3, 500, 5, True), (4, 1, None, False), (5, 64, 3, True)])
def test_nms_against_numpy_restatement(oracle, seed, n, levels, deg):
    d = _dets(n, seed, levels=levels, degenerate=deg)
    for thr in (0.4, 0.25, 0.5, 0.0, 1.0):
        np.testing.assert_array_equal(oracle.nms(d, thr), np_nms(d, thr))


def test_argsort_is_the_stable_descending_order(oracle):
    rng = np.random.default_rng(9)
    s = (np.round(rng.uniform(-1, 1, 5000) * 50) / 50).astype(F)     # many ties, both signs, zeros of both signs
    s[::13] = F(-0.0)
    s[1::13] = F(0.0)
    np.testing.assert_array_equal(oracle.argsort_descending(s), np.argsort(-s.astype(np.float64), kind="stable"))


def test_clip_and_landmark_pred_against_numpy(oracle):
    """bbox_transform.rs:27-68 (clip_boxes / clip_points: min then max), :123-160 (landmark_pred: d * size + centre)."""
    rng = np.random.default_rng(11)
    n = 400
    x1, y1 = rng.uniform(-50, 600, n), rng.uniform(-50, 600, n)
    boxes = np.stack([x1, y1, x1 + rng.uniform(1, 300, n), y1 + rng.uniform(1, 300, n)], 1).astype(F)
    got = oracle.clip_boxes(boxes.copy(), (480, 640))
    exp = boxes.copy()
    exp[:, 0::2] = np.maximum(np.minimum(exp[:, 0::2], F(639)), F(0))
    exp[:, 1::2] = np.maximum(np.minimum(exp[:, 1::2], F(479)), F(0))
    np.testing.assert_array_equal(got, exp)
    pts = rng.uniform(-100, 800, (n, 10)).astype(F)
    got = oracle.clip_points(pts.copy(), (480, 640))
    exp = pts.copy()
    exp[:, 0::2] = np.maximum(np.minimum(exp[:, 0::2], F(639)), F(0))
    exp[:, 1::2] = np.maximum(np.minimum(exp[:, 1::2], F(479)), F(0))
    np.testing.assert_array_equal(got, exp)
    deltas = rng.normal(0, 0.4, (n, 10)).astype(F)
    w = (boxes[:, 2] - boxes[:, 0] + F(1)).astype(F)
    h = (boxes[:, 3] - boxes[:, 1] + F(1)).astype(F)
    cx = (boxes[:, 0] + (F(0.5) * (w - F(1)).astype(F)).astype(F)).astype(F)
    cy = (boxes[:, 1] + (F(0.5) * (h - F(1)).astype(F)).astype(F)).astype(F)
    exp = np.empty((n, 10), F)
    for k in range(5):
        exp[:, 2 * k] = ((deltas[:, 2 * k] * w).astype(F) + cx).astype(F)
        exp[:, 2 * k + 1] = ((deltas[:, 2 * k + 1] * h).astype(F) + cy).astype(F)
    np.testing.assert_array_equal(oracle.landmark_pred(boxes, deltas), exp)


def test_nonlinear_pred_against_numpy(oracle):
    """bbox_transform.rs:90-121: centre/size form, exp of the size deltas (libm expf vs numpy: 1e-6 relative)."""
    rng = np.random.default_rng(12)
    n = 400
    x1, y1 = rng.uniform(0, 600, n), rng.uniform(0, 600, n)
    boxes = np.stack([x1, y1, x1 + rng.uniform(1, 300, n), y1 + rng.uniform(1, 300, n)], 1).astype(F)
    d = np.concatenate([rng.normal(0, 0.3, (n, 2)), rng.normal(0, 0.2, (n, 2))], 1).astype(F)
    w = (boxes[:, 2] - boxes[:, 0] + F(1)).astype(F)
    h = (boxes[:, 3] - boxes[:, 1] + F(1)).astype(F)
    cx = (boxes[:, 0] + (F(0.5) * (w - F(1)))).astype(F)
    cy = (boxes[:, 1] + (F(0.5) * (h - F(1)))).astype(F)
    pcx = ((d[:, 0] * w).astype(F) + cx).astype(F)
    pcy = ((d[:, 1] * h).astype(F) + cy).astype(F)
    pw = (np.exp(d[:, 2]).astype(F) * w).astype(F)
    ph = (np.exp(d[:, 3]).astype(F) * h).astype(F)
    exp = np.stack([pcx - F(0.5) * (pw - F(1)), pcy - F(0.5) * (ph - F(1)), pcx + F(0.5) * (pw - F(1)), pcy + F(0.5) * (ph - F(1))], 1).astype(F)
    np.testing.assert_allclose(oracle.nonlinear_pred(boxes, d), exp, rtol=2e-6, atol=1e-4)


def np_forward_postprocess(heads, base_anchors, strides, conf_thr, iou_thr, det_scale, image_hw=(640, 640)):
    """RetinaFaceDetection::_forward after the CNN + _postprocess (face_detection.rs:319-493, bbox_pred :516-549, landmark_pred
    :551-570, rcnn/anchors.rs:3-21) in numpy float32: EVERY anchor is decoded and clipped, then `>= confidence_threshold`
    selects, the strides are stacked 32|16|8, sorted by descending score (stable), NMS, rows gathered, `/= det_scale`."""
    props, scs, lmks = [], [], []
    for s, stride in enumerate(strides):
        scores, deltas, ldeltas = (np.asarray(heads[3 * s + k], F) for k in range(3))   # (C,H,W)
        base = np.asarray(base_anchors[s], F)
        A = base.shape[0]
        H, W = deltas.shape[1:]
        # anchors(): all_anchors[ih, iw, k] = base[k] + (iw*stride, ih*stride, iw*stride, ih*stride)
        sw = (np.arange(W) * stride).astype(F)[None, :, None]
        sh = (np.arange(H) * stride).astype(F)[:, None, None]
        an = np.empty((H, W, A, 4), F)
        an[..., 0] = base[None, None, :, 0] + sw
        an[..., 1] = base[None, None, :, 1] + sh
        an[..., 2] = base[None, None, :, 2] + sw
        an[..., 3] = base[None, None, :, 3] + sh
        an = an.reshape(-1, 4)
        sc = scores[A:].transpose(1, 2, 0).reshape(-1)                     # fg channels, (H,W,A) order
        d = deltas.transpose(1, 2, 0).reshape(-1, 4)                       # bbox_stds = 1
        w = (an[:, 2] - an[:, 0] + F(1)).astype(F)
        h = (an[:, 3] - an[:, 1] + F(1)).astype(F)
        cx = (an[:, 0] + F(0.5) * (w - F(1))).astype(F)
        cy = (an[:, 1] + F(0.5) * (h - F(1))).astype(F)
        pcx = ((d[:, 0] * w).astype(F) + cx).astype(F)
        pcy = ((d[:, 1] * h).astype(F) + cy).astype(F)
        pw = (np.exp(d[:, 2]).astype(F) * w).astype(F)
        ph = (np.exp(d[:, 3]).astype(F) * h).astype(F)
        box = np.stack([pcx - F(0.5) * (pw - F(1)), pcy - F(0.5) * (ph - F(1)), pcx + F(0.5) * (pw - F(1)), pcy + F(0.5) * (ph - F(1))], 1).astype(F)
        box[:, 0::2] = np.maximum(np.minimum(box[:, 0::2], F(image_hw[1] - 1)), F(0))
        box[:, 1::2] = np.maximum(np.minimum(box[:, 1::2], F(image_hw[0] - 1)), F(0))
        ld = ldeltas.transpose(1, 2, 0).reshape(-1, 5, 2)                  # landmark_std = 1
        lm = np.empty_like(ld)
        lm[:, :, 0] = ((ld[:, :, 0] * w[:, None]).astype(F) + cx[:, None]).astype(F)
        lm[:, :, 1] = ((ld[:, :, 1] * h[:, None]).astype(F) + cy[:, None]).astype(F)
        sel = np.nonzero(sc >= F(conf_thr))[0]
        props.append(box[sel]); scs.append(sc[sel]); lmks.append(lm[sel])
    box, sc, lm = np.concatenate(props), np.concatenate(scs), np.concatenate(lmks)
    if len(box) == 0:
        return np.zeros((0, 5), F), np.zeros((0, 5, 2), F), 0
    order = np.argsort(-sc.astype(np.float64), kind="stable")
    pre = np.concatenate([box[order], sc[order, None]], 1).astype(F)
    keep = np_nms(pre, iou_thr)
    det = pre[keep].copy()
    det[:, :4] = (det[:, :4] / F(det_scale)).astype(F)
    return det, (lm[order][keep] / F(det_scale)).astype(F), len(box)


@pytest.mark.parametrize("seed,faces,thr,scale", [(21, 12, 0.7, 1.0), (22, 30, 0.5, 0.3333333), (23, 0, 0.7, 0.5), (24, 6, 0.02, 0.59259259)])
def test_forward_postprocess_against_numpy(oracle, seed, faces, thr, scale):
    from rs_face_detection_b200.utils import synth
    heads, _ = synth.make_heads(1, seed=seed, n_faces=faces)
    heads1 = [h[0] for h in heads]
    cfg = oracle.make_det_cfg(conf_thr=thr, iou_thr=0.4)
    det, lmk, K = oracle.detect_post(cfg, heads1, scale)
    edet, elmk, eK = np_forward_postprocess(heads1, synth.BASE_ANCHORS, synth.STRIDES, thr, 0.4, scale)
    assert K == eK and det.shape == edet.shape
    np.testing.assert_array_equal(det[:, 4], edet[:, 4])                 # same boxes picked, same order
    np.testing.assert_allclose(det[:, :4], edet[:, :4], rtol=2e-6, atol=1e-4)   # expf vs numpy exp
    np.testing.assert_array_equal(lmk, elmk)                             # no transcendental on this path: bit-exact


def test_letterbox_geometry_against_numpy(oracle):
    """_preprocess geometry (face_detection.rs:140-153) in f32 with Rust's truncating `as i32`, over 3 000 random frame sizes
    (incl. aspect ratios on both sides of the model's, where the f32 quotient decides the branch)."""
    rng = np.random.default_rng(31)
    sizes = [(1080, 1920), (2160, 3840), (640, 640), (641, 640), (640, 641), (1, 1), (7, 5000), (5000, 7)]
    sizes += [(int(h), int(w)) for h, w in zip(rng.integers(1, 5000, 3000), rng.integers(1, 5000, 3000))]
    for h, w in sizes:
        im_ratio = F(h) / F(w)
        model_ratio = F(640) / F(640)
        if im_ratio > model_ratio:
            nh = 640
            nw = int(F(nh) / im_ratio)          # (new_height as f32 / im_ratio) as i32: truncation toward zero
        else:
            nw = 640
            nh = int(F(nw) * im_ratio)
        ds = F(nh) / F(h)
        assert oracle.letterbox_geometry(h, w) == (nw, nh, ds), (h, w)


def py_face_selection(img_hw, fb, has_kps, enroll, prm):
    """FaceSelection::call (face_selection.rs:72-189) and get_biggest_area_face (:28-53) in f32, returning row indices."""
    H, W = F(img_hw[0]), F(img_hw[1])
    ml, mr, me, mn = (F(x) for x in prm)
    if enroll:
        best, bi = F(0), -1
        if has_kps:
            for i, b in enumerate(fb):
                a = F(F(b[2] - b[0]) * F(b[3] - b[1]))
                if a > best:
                    best, bi = a, i
        return bi, bi
    mcl, mcr = F(ml * W), F(mr * W)
    edge = min(F(50.0), F(me * W))
    x_cen = F(W / F(2))
    valid = []
    for i, b in enumerate(fb):
        area = F(F(b[2] - b[0]) * F(b[2] - b[0]))                         # (:115) the width, squared
        cw, ch = F(F(b[0] + b[2]) / F(2)), F(F(b[1] + b[3]) / F(2))
        if cw >= edge and cw <= F(W - edge) and ch >= edge and ch <= F(H - edge) and F(area / F(H * W)) >= mn:
            valid.append(i)
    center = [i for i in valid if -mcl <= F(F(F(fb[i][0] + fb[i][2]) / F(2)) - x_cen) <= mcr]
    if not center:
        center = valid if valid else list(range(len(fb)))
    out, mx = -1, F(0)
    for i in center:
        t = F(F(fb[i][2] - fb[i][0]) + F(fb[i][3] - fb[i][1]))
        if t > mx:
            mx, out = t, i
    if out < 0:
        return -1, -1
    ki = -1
    if has_kps:
        o = fb[out]
        for i, b in enumerate(fb):
            if abs(F(o[0] - b[0])) <= 2 and abs(F(o[1] - b[1])) <= 2 and abs(F(o[2] - b[2])) <= 2 and abs(F(o[3] - b[3])) <= 2:
                ki = i
                break
    return out, ki


def test_face_selection_against_python_restatement(oracle):
    rng = np.random.default_rng(41)
    prm = (0.3, 0.3, 0.1, 0.0075)
    for trial in range(400):
        H, W = int(rng.integers(200, 2200)), int(rng.integers(200, 3900))
        n = int(rng.integers(0, 9))
        x1, y1 = rng.uniform(-20, W, n), rng.uniform(-20, H, n)
        fb = np.stack([x1, y1, x1 + rng.uniform(2, W / 2, n), y1 + rng.uniform(2, H / 2, n), rng.uniform(0, 1, n)], 1).astype(F).reshape(-1, 5)
        if n >= 2 and trial % 3 == 0:
            fb[1, :4] = fb[0, :4] + rng.uniform(-2.5, 2.5, 4).astype(F)     # a near-duplicate: the 2 px key-point match (:160-176)
        for has_kps in (True, False):
            for enroll in (False, True):
                got = oracle.face_selection((H, W), fb, np.zeros((n, 5, 2), F) if has_kps else None, is_enroll=enroll, params=prm)
                assert got == py_face_selection((H, W), fb, has_kps, enroll, prm), (trial, H, W, has_kps, enroll, fb)


def test_model_preprocessors_against_cv2_numpy(oracle):
    """The three post-align preprocessors (face_extraction.rs:38-75, face_quality.rs:52-100, face_quality_assessment.rs): cv::resize
    INTER_LINEAR -> BGR2RGB -> (pixel as f32 - mean) * mul -> CHW, restated with this container's cv2 + numpy f32."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(51)
    for name, (mean, mul) in oracle.MODEL_NORMS.items():
        for (h, w), out in (((112, 112), (112, 112)), ((112, 112), (224, 224)), ((97, 131), (112, 112)), ((300, 280), (128, 96))):
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            rgb = cv2.cvtColor(cv2.resize(img, out, interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB).astype(F)
            exp = ((rgb - np.asarray(mean, F)).astype(F) * np.asarray(mul, F)).astype(F).transpose(2, 0, 1)
            np.testing.assert_array_equal(oracle.model_preprocess(img, out, mean, mul), exp, err_msg="%s %s -> %s" % (name, (h, w), out))


def test_to_tensor_against_numpy(oracle):
    """face_detection.rs:222-229: im_tensor[0, i, y, x] = (pixel[2 - i] as f32 / pixel_scale - pixel_means[2 - i]) / pixel_stds[2 - i]
    — BGR -> RGB planes, two f32 divisions, with the reference's identity constants and with non-trivial ones."""
    rng = np.random.default_rng(61)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    for scale, means, stds in ((1.0, (0, 0, 0), (1, 1, 1)), (255.0, (0.406, 0.456, 0.485), (0.225, 0.224, 0.229)), (1.7, (12.5, 0.0, 99.0), (3.0, 0.1, 7.0))):
        exp = np.empty((1, 3, 37, 53), F)
        for i in range(3):
            exp[0, i] = (((img[:, :, 2 - i].astype(F) / F(scale)).astype(F) - F(means[2 - i])).astype(F) / F(stds[2 - i])).astype(F)
        np.testing.assert_array_equal(oracle.to_tensor(img, scale, means, stds), exp)


def np_cpu_nms(dets, thr):
    """rcnn/cpu_nms.rs:10-55: every kept box marks EVERY box (earlier ones and itself included) with `ovr >= thresh`; a NaN overlap
    (zero-area boxes) marks nothing.  (The reference sorts with sort_unstable_by: distinct scores here, so the order is defined.)"""
    dets = np.asarray(dets, F)
    x1, y1, x2, y2, sc = (dets[:, k] for k in range(5))
    areas = ((x2 - x1 + F(1)).astype(F) * (y2 - y1 + F(1)).astype(F)).astype(F)
    order = np.argsort(-sc.astype(np.float64), kind="stable")
    sup = np.zeros(len(dets), bool)
    keep = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in order:
            if sup[i]:
                continue
            keep.append(int(i))
            w = np.maximum((np.minimum(x2[i], x2) - np.maximum(x1[i], x1) + F(1)).astype(F), F(0))
            h = np.maximum((np.minimum(y2[i], y2) - np.maximum(y1[i], y1) + F(1)).astype(F), F(0))
            inter = (w * h).astype(F)
            ovr = (inter / ((areas[i] + areas).astype(F) - inter).astype(F)).astype(F)
            sup |= ovr >= F(thr)
    return np.array(keep, np.int64)


@pytest.mark.parametrize("seed,n,deg", [(71, 400, False), (72, 700, True), (73, 1, False)])
def test_cpu_nms_against_numpy_restatement(oracle, seed, n, deg):
    d = _dets(n, seed, degenerate=deg)
    assert len(np.unique(d[:, 4])) == n
    for thr in (0.3, 0.4, 0.5, 1.0):
        np.testing.assert_array_equal(oracle.cpu_nms(d, thr), np_cpu_nms(d, thr))

"""Oracle vs the reference's own test INPUTS (the reference tests only println!, so expected values are derived by
hand from the reference source and stated here).  SURVEY.md §4."""
import numpy as np


def test_nms_reference_vector(oracle):
    # src/processing/nms.rs:76-83 ; IoU(0,3)=1.0 -> 3 suppressed by 0 ; IoU(1,2)=0.1657 < 0.4
    dets = np.array([[100, 100, 210, 210, 0.72], [250, 250, 420, 420, 0.8], [220, 220, 320, 330, 0.92],
                     [100, 100, 210, 210, 0.6]], np.float32)
    assert oracle.nms(dets, 0.4).tolist() == [2, 1, 0]
    ov = oracle.bbox_overlaps(dets[:, :4], dets[:, :4])
    assert ov[0, 3] == np.float32(1.0)
    assert abs(ov[1, 2] - 0.16573009) < 1e-7


def test_cpu_nms_reference_vector(oracle):
    # src/rcnn/cpu_nms.rs:65-72, thr 0.3
    dets = np.array([[100, 100, 210, 210, 0.72], [250, 250, 420, 420, 0.8], [220, 220, 320, 330, 0.92],
                     [100, 100, 210, 210, 0.6]], np.float32)
    assert oracle.cpu_nms(dets, 0.3).tolist() == [2, 1, 0]


def test_nms_threshold_semantics(oracle):
    # two boxes with IoU exactly 0.5 (areas 100 / 50 overlap... built from integers so the quotient is exact)
    a = [0, 0, 9, 9, 0.9]      # area 100
    b = [0, 0, 9, 4, 0.8]      # area 50, inter 50 -> IoU 50/100 = 0.5
    dets = np.array([a, b], np.float32)
    assert oracle.nms(dets, 0.5).tolist() == [0, 1]        # survivor iff ovr <= thr (nms.rs:58)
    assert oracle.cpu_nms(dets, 0.5).tolist() == [0]       # suppress iff ovr >= thr (cpu_nms.rs:48)
    assert oracle.nms(dets, 0.49).tolist() == [0]


def test_nms_stable_ties(oracle):
    # equal scores keep original index order (stable sort_by, nms.rs:6)
    dets = np.array([[0, 0, 10, 10, 0.5], [100, 100, 110, 110, 0.5], [0, 0, 10, 10, 0.5], [200, 200, 210, 210, 0.7]], np.float32)
    assert oracle.nms(dets, 0.4).tolist() == [3, 0, 1]
    assert oracle.argsort_descending(dets[:, 4]).tolist() == [3, 0, 1, 2]


def test_anchors_reference_vector(oracle):
    # src/rcnn/anchors.rs:30-37
    base = np.array([[0, 0, 15, 15], [0, 0, 31, 31]], np.float32)
    out = oracle.anchors_plane(2, 2, 16, base)
    assert out.shape == (2, 2, 2, 4)
    for ih in range(2):
        for iw in range(2):
            for k in range(2):
                np.testing.assert_array_equal(out[ih, iw, k], base[k] + np.array([iw * 16, ih * 16, iw * 16, ih * 16], np.float32))


def test_retinaface_base_anchors(oracle):
    # generate_anchors.rs:218-251 with the cfg of face_detection.rs:55-80 ; SURVEY §3.2
    a = oracle.generate_anchors_fpn2_retinaface(False)
    exp = np.array([[[-248, -248, 263, 263], [-120, -120, 135, 135]], [[-56, -56, 71, 71], [-24, -24, 39, 39]],
                    [[-8, -8, 23, 23], [0, 0, 15, 15]]], np.float32)
    np.testing.assert_array_equal(a, exp)


def test_generate_anchors_classic_unrounded_hs(oracle):
    # generate_anchors.rs:195-201 ; hs is NOT rounded (:146): ratio 0.5 -> ws=round(sqrt(512))=23, hs=11.5
    a = oracle.generate_anchors(16, [0.5, 1.0, 2.0], [8.0, 16.0, 32.0])
    assert a.shape == (9, 4)
    np.testing.assert_array_equal(a[0], np.array([-84, -38, 99, 53], np.float32))   # py-faster-rcnn would give [-84,-40,99,55]
    np.testing.assert_array_equal(a[3], np.array([-56, -56, 71, 71], np.float32))
    r = oracle.ratio_enum([0, 0, 15, 15], [0.5, 1.0, 2.0])                           # :166-172
    np.testing.assert_array_equal(r[0], np.array([-3.5, 2.25, 18.5, 12.75], np.float32))
    s = oracle.scale_enum([0, 0, 15, 15], [0.5, 1.0, 2.0])                           # :174-180
    np.testing.assert_array_equal(s, np.array([[4, 4, 11, 11], [0, 0, 15, 15], [-8, -8, 23, 23]], np.float32))


def test_generate_anchors_fpn(oracle):
    # generate_anchors.rs:203-215
    a = oracle.generate_anchors_fpn([64, 32, 16, 8, 4], [0.5, 1.0, 2.0, 1.0, 1.0], [8.0] * 5)
    assert len(a) == 5 and all(x.shape == (1, 4) for x in a)
    np.testing.assert_array_equal(a[1][0], np.array([-112, -112, 143, 143], np.float32))


def test_bbox_transform_vectors(oracle):
    # bbox_transform.rs:198-278
    boxes = np.array([[50, 50, 100, 100], [30, 30, 70, 70]], np.float32)
    d = np.array([[1, 1, 1, 1, 2, 2, 2, 2]] * 2, np.float32)
    out = oracle.iou_pred(boxes, d, 2)
    np.testing.assert_array_equal(out[0], [51, 51, 101, 101, 52, 52, 102, 102])
    cb = oracle.clip_boxes(np.array([[50, 50, 150, 150, 60, 60, 160, 160], [30, 30, 200, 200, 40, 40, 220, 220]], np.float32), (100, 100))
    np.testing.assert_array_equal(cb[0], [50, 50, 99, 99, 60, 60, 99, 99])
    cp = oracle.clip_points(np.array([[50, 50, 150, 150, 60, 60, 160, 160, 70, 70]], np.float32), (100, 100))
    np.testing.assert_array_equal(cp[0], [50, 50, 99, 99, 60, 60, 99, 99, 70, 70])
    ex = np.array([[50, 50, 150, 150], [30, 30, 200, 200]], np.float32)
    gt = np.array([[60, 60, 170, 170], [35, 35, 210, 210]], np.float32)
    t = oracle.nonlinear_transform(ex, gt)
    np.testing.assert_allclose(t[0], [15 / 101, 15 / 101, np.log(111 / 101), np.log(111 / 101)], rtol=1e-6)
    # nonlinear_pred is the inverse of nonlinear_transform
    p = oracle.nonlinear_pred(ex, t)
    np.testing.assert_allclose(p, gt, rtol=1e-5, atol=1e-3)
    lp = oracle.landmark_pred(ex, np.array([[0.1, 0.2, 0.1, 0.2, 0.2, 0.1, 0.2, 0.1, 0.3, 0.3]] * 2, np.float32))
    np.testing.assert_allclose(lp[0, :2], [0.1 * 101 + 100, 0.2 * 101 + 100], rtol=1e-6)


def test_letterbox_geometry(oracle):
    # face_detection.rs:140-153 ; SURVEY §8 a1
    assert oracle.letterbox_geometry(1080, 1920) == (640, 360, np.float32(360 / 1080))
    assert oracle.letterbox_geometry(2160, 3840) == (640, 360, np.float32(360 / 2160))
    nw, nh, sc = oracle.letterbox_geometry(1000, 700)
    assert (nw, nh) == (448, 640) and sc == np.float32(0.64)


def test_resize_exact_integer_scales(oracle):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
    out = oracle.resize_linear(img, (640, 360))
    np.testing.assert_array_equal(out, img[1::3, 1::3])    # scale 3 -> pure point sampling src[3y+1][3x+1]
    small = img[:64, :64]
    avg = ((small[0::2, 0::2].astype(int) + small[0::2, 1::2] + small[1::2, 0::2] + small[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    np.testing.assert_array_equal(oracle.resize_linear(small, (32, 32)), avg)


def test_detect_post_empty_and_order(oracle):
    from rs_face_detection_b200.utils import synth
    heads, _ = synth.make_heads(1, seed=5, n_faces=0)
    cfg = oracle.make_det_cfg(conf_thr=0.7, iou_thr=0.4)
    det, lmk, K = oracle.detect_post(cfg, [h[0] for h in heads], 1.0)
    assert det.shape == (0, 5) and lmk.shape == (0, 5, 2) and K == 0      # face_detection.rs:413-419
    heads, faces = synth.make_heads(1, seed=1234, n_faces=20)
    det, lmk, K = oracle.detect_post(cfg, [h[0] for h in heads], 0.5)
    assert K > 20 and 10 <= len(det) <= 25
    assert np.all(np.diff(det[:, 4]) <= 0)                                 # pick order is score-descending
    # every planted face is recovered by a detection with IoU > 0.5 (boxes were divided by det_scale 0.5)
    ov = oracle.bbox_overlaps(faces[0] / 0.5, det[:, :4])
    assert (ov.max(1) > 0.5).mean() > 0.8


def test_model_preprocess_known_values(oracle):
    # face_extraction.rs:69: (p - 127.5) * 0.0078125, RGB order, NCHW
    img = np.zeros((2, 2, 3), np.uint8)
    img[0, 0] = (10, 20, 30)          # B, G, R
    out = oracle.model_preprocess(img, (2, 2), *oracle.MODEL_NORMS["face_extraction"])
    assert out.shape == (3, 2, 2)
    np.testing.assert_array_equal(out[:, 0, 0], np.float32([(30 - 127.5) * 0.0078125, (20 - 127.5) * 0.0078125, (10 - 127.5) * 0.0078125]))
    np.testing.assert_array_equal(out[:, 1, 1], np.float32([-127.5 * 0.0078125] * 3))


def test_face_selection_hand_derived(oracle):
    """FaceSelection::call (face_selection.rs:72-189) on hand-derived cases (the reference has no test for it)."""
    fb = np.array([[100, 100, 200, 220, .9],      # left of the centre band (|cx - 960| > 0.3 * 1920)
                   [900, 400, 1100, 640, .8],     # central, size 200 + 240 = 440  -> selected
                   [10, 10, 60, 70, .99],         # centre inside the 50 px edge margin -> invalid
                   [902, 401, 1099, 641, .7]], np.float32)   # central, size 437
    kps = np.zeros((4, 5, 2), np.float32)
    assert oracle.face_selection((1080, 1920), fb, kps) == (1, 1)
    assert oracle.face_selection((1080, 1920), fb, None) == (1, -1)                 # key_points = None
    assert oracle.face_selection((1080, 1920), fb[:0], kps[:0]) == (-1, -1)         # no detections -> (None, None)
    # key points come from the FIRST row within 2 px of the selected box (:160-176), not necessarily the box itself
    fb2 = np.array([[899, 399, 1099, 639, .5], [900, 400, 1101, 640.5, .8]], np.float32)
    assert oracle.face_selection((1080, 1920), fb2, kps[:2]) == (1, 0)
    # no central box -> the valid boxes compete; none valid -> every box competes (:137-143)
    assert oracle.face_selection((1080, 1920), fb[[0, 2]], kps[:2]) == (0, 0)
    assert oracle.face_selection((1080, 1920), fb[[2]], kps[:1]) == (0, 0)
    # the minimum-size test squares the WIDTH (:115): a 300 x 20 sliver is "big enough", a 20 x 300 one is not
    sl = np.array([[800, 500, 1100, 520, .9], [950, 300, 970, 600, .9]], np.float32)
    assert oracle.face_selection((1080, 1920), sl, kps[:2]) == (0, 0)
    # enroll: biggest (x2-x1)*(y2-y1), first maximum wins, nothing without key points (:28-53)
    assert oracle.face_selection((1080, 1920), fb, kps, is_enroll=True) == (1, 1)
    assert oracle.face_selection((1080, 1920), fb, None, is_enroll=True) == (-1, -1)
    eq = np.array([[0, 0, 10, 10, .1], [5, 5, 15, 15, .2]], np.float32)
    assert oracle.face_selection((100, 100), eq, kps[:2], is_enroll=True) == (0, 0)

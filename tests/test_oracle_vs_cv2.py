"""Pins the oracle's restatement of the three OpenCV calls: against the committed cv2 4.13.0 golden fixtures (always)
and against a live cv2 when it is importable."""
import numpy as np
import pytest

try:
    import cv2
    cv2.setNumThreads(1)
except Exception:  # pragma: no cover
    cv2 = None


def test_golden_resize(oracle, golden):
    for i in range(int(golden["resize_n"])):
        dw, dh = golden["resize_dsize_%d" % i]
        np.testing.assert_array_equal(oracle.resize_linear(golden["resize_in_%d" % i], (int(dw), int(dh))), golden["resize_out_%d" % i])


def test_golden_warp(oracle, golden):
    for i in range(int(golden["warp_n"])):
        np.testing.assert_array_equal(oracle.warp_affine(golden["warp_in_%d" % i], golden["warp_M_%d" % i], (112, 112)),
                                      golden["warp_out_%d" % i])


def test_golden_estimate(oracle, golden):
    pts, Mg, inl, ok = golden["est_pts"], golden["est_M"], golden["est_inliers"], golden["est_ok"]
    rejected = 0
    for t in range(len(pts)):
        M, mask = oracle.estimate_affine_partial_2d(pts[t], oracle.ARCFACE_TEMPLATE)
        assert (M is not None) == bool(ok[t])
        if M is None:
            continue
        np.testing.assert_array_equal(mask != 0, inl[t] != 0)
        np.testing.assert_allclose(M, Mg[t], rtol=0, atol=1e-7)   # LM refinement vs closed-form least squares
        rejected += int(mask.sum() < 5)
    assert rejected > 10   # the fixture does exercise LMedS outlier rejection


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
@pytest.mark.parametrize("src,dst", [((1080, 1920), (640, 360)), ((2160, 3840), (640, 360)), ((480, 640), (640, 480)),
                                     ((1000, 700), (448, 640)), ((333, 517), (640, 412)), ((97, 131), (640, 473)),
                                     ((50, 2000), (640, 16)), ((3000, 40), (8, 640))])
def test_live_resize(oracle, src, dst):
    rng = np.random.default_rng(src[0] * 7 + dst[0])
    img = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    np.testing.assert_array_equal(oracle.resize_linear(img, dst), cv2.resize(img, dst, interpolation=cv2.INTER_LINEAR))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_warp(oracle):
    rng = np.random.default_rng(3)
    for t in range(8):
        sh, sw = [(1080, 1920), (2160, 3840), (480, 640), (300, 200)][t % 4]
        img = rng.integers(0, 256, (sh, sw, 3), dtype=np.uint8)
        s, th = rng.uniform(0.3, 3.0), rng.uniform(-0.7, 0.7)
        a, b = s * np.cos(th), s * np.sin(th)
        cx, cy = rng.uniform(-30, sw + 30), rng.uniform(-30, sh + 30)
        M = np.array([[a, -b, 56 - (a * cx - b * cy)], [b, a, 56 - (b * cx + a * cy)]], np.float64)
        ref = cv2.warpAffine(img, M, (112, 112), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        np.testing.assert_array_equal(oracle.warp_affine(img, M, (112, 112)), ref)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_live_estimate(oracle):
    from rs_face_detection_b200.utils import synth
    pts = synth.make_landmarks(400, seed=11)
    for t in range(len(pts)):
        Mc, inl = cv2.estimateAffinePartial2D(pts[t], oracle.ARCFACE_TEMPLATE, method=cv2.LMEDS, ransacReprojThreshold=3.0,
                                              maxIters=2000, confidence=0.99, refineIters=10)
        Mo, mask = oracle.estimate_affine_partial_2d(pts[t], oracle.ARCFACE_TEMPLATE)
        assert (Mc is None) == (Mo is None)
        if Mc is not None:
            np.testing.assert_array_equal(inl.ravel() != 0, mask != 0)
            np.testing.assert_allclose(Mo, Mc, rtol=0, atol=1e-7)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_align_fallback_vs_cv2(oracle):
    """FaceAlignment::call's empty-transform branch (face_alignment.rs:64-116) restated with numpy f32 + cv2.resize on the
    ROI, against the oracle: same crops, same Err cases (Mat::roi range check, empty ROI)."""
    f32 = np.float32
    rng = np.random.default_rng(8)

    def reference(img, bbox):
        H, W = img.shape[:2]
        if bbox is None:
            d0, d1 = f32(W) * f32(0.0625), f32(H) * f32(0.0625)
            det = [d0, d1, f32(W) - d0, f32(H) - d1]
        else:
            det = [f32(v) for v in bbox[:4]]
        fmax = lambda a, b: b if np.isnan(a) else (a if np.isnan(b) else max(a, b))            # Rust f32::max
        as_i32 = lambda v: 0 if np.isnan(v) else int(np.clip(np.trunc(v), -2**31, 2**31 - 1))   # Rust `as i32`
        bb = [fmax(det[0] - f32(22), f32(0)), fmax(det[1] - f32(22), f32(0)), fmax(det[2] + f32(22), f32(W)), fmax(det[1] + f32(22), f32(H))]
        x0, y0, x1, y1 = (as_i32(v) for v in bb)
        wd, ht = x1 - x0, y1 - y0
        if not (0 <= x0 and 0 <= wd and x0 + wd <= W and 0 <= y0 and 0 <= ht and y0 + ht <= H) or wd == 0 or ht == 0:
            return None                                                                         # Mat::roi / cv::resize -> Err
        return cv2.resize(np.ascontiguousarray(img[y0:y0 + ht, x0:x0 + wd]), (112, 112), interpolation=cv2.INTER_LINEAR)

    n_ok = n_err = 0
    for (h, w) in [(480, 640), (1080, 1920), (113, 131), (112, 134), (60, 45)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cases = [None, [100.5, 60.25, 180.0, 140.0], [10, 5, 60, 70, 0.9], [w - 150.0, 30, w - 22.0, 90], [w - 150.0, 30, w - 21.5, 90],
                 [50, h - 21.0, 120, h - 5.0], [np.nan, 40, np.nan, 90], [w + 40.0, 40, w - 100.0, 90], [22.0 + w - 112, 22.0 + h - 112, 5, 5]]
        for bbox in cases:
            want = reference(img, bbox)
            got = oracle.align_fallback(img, None if bbox is None else np.array(bbox, np.float32))
            assert (want is None) == (got is None), (h, w, bbox)
            if want is not None:
                np.testing.assert_array_equal(got, want)
                n_ok += 1
            else:
                n_err += 1
    assert n_ok >= 15 and n_err >= 8
    # and the estimate really is empty for the landmark sets the fallback tests use (cv2 returns None)
    same = np.tile(np.array([[300.0, 200.0]], np.float32), (5, 1))
    M, _ = cv2.estimateAffinePartial2D(same, oracle.ARCFACE_TEMPLATE, method=cv2.LMEDS, ransacReprojThreshold=3.0, maxIters=2000,
                                       confidence=0.99, refineIters=10)
    assert M is None and oracle.estimate_affine_partial_2d(same, oracle.ARCFACE_TEMPLATE)[0] is None


# ---- N4: byte_data_to_opencv (utils.rs:8-52) = cv::imdecode -> the JPEG restatement of oracle/fd_jpeg_oracle.c -------------
def _jpeg_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_golden.npz"))


def test_jpeg_golden(oracle):
    g = _jpeg_golden()
    for i in range(int(g["n"])):
        np.testing.assert_array_equal(oracle.jpeg_decode(g["jpeg_%d" % i].tobytes()), g["bgr_%d" % i])
    for k in ("unsupported_progressive", "unsupported_gray"):
        with pytest.raises(ValueError):
            oracle.jpeg_decode(g[k].tobytes())


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_jpeg_vs_live_cv2(oracle):
    """bit-exact against cv2.imdecode over samplings, qualities, odd sizes, restart intervals and optimised Huffman tables"""
    SS = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420}
    rng = np.random.default_rng(5)
    n = 0
    for (h, w) in [(480, 640), (123, 77), (17, 33), (8, 8), (3, 2), (5, 4), (9, 6), (1, 7), (31, 250)]:
        for ss in SS.values():
            for q in (20, 75, 92, 100):
                img = np.clip(rng.normal(128, 50, (h, w, 3)) + np.linspace(0, 60, w)[None, :, None], 0, 255).astype(np.uint8)
                params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss]
                if (h + q) % 3 == 0:
                    params += [cv2.IMWRITE_JPEG_RST_INTERVAL, 1 + (w % 4)]
                if q == 92:
                    params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
                ok, buf = cv2.imencode(".jpg", img, params)
                np.testing.assert_array_equal(oracle.jpeg_decode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_UNCHANGED))
                n += 1
    assert n == 108

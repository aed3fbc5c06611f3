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

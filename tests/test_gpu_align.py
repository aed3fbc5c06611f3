"""Align parity: similarity estimate (M bit-identical to the oracle, 1e-7 to cv2) and the fixed-point warp (crops
bit-identical to the oracle and to the cv2 golden fixtures)."""
import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu


def test_estimate_bit_exact_vs_oracle(ctx, oracle):
    pts = synth.make_landmarks(3000, seed=21)
    pts[::9, 2] += np.random.default_rng(0).normal(0, 40, (len(pts[::9]), 2)).astype(np.float32)   # gross outliers -> LMedS rejects
    M, ok = ctx.estimate_affine_partial_2d(pts)
    rejected = 0
    for t in range(len(pts)):
        Mo, mask = oracle.estimate_affine_partial_2d(pts[t], oracle.ARCFACE_TEMPLATE)
        assert bool(ok[t]) == (Mo is not None)
        if Mo is not None:
            np.testing.assert_array_equal(M[t], Mo)      # same fp64 operation order on both sides
            rejected += int(mask.sum() < 5)
    assert rejected > 100


def test_estimate_golden_cv2(ctx, golden):
    M, ok = ctx.estimate_affine_partial_2d(golden["est_pts"])
    np.testing.assert_array_equal(ok, golden["est_ok"])
    np.testing.assert_allclose(M[ok != 0], golden["est_M"][ok != 0], rtol=0, atol=1e-7)


def test_estimate_degenerate(ctx, oracle):
    same = np.tile(np.array([[10.0, 20.0]], np.float32), (5, 1))[None]
    M, ok = ctx.estimate_affine_partial_2d(same)
    Mo, _ = oracle.estimate_affine_partial_2d(same[0], oracle.ARCFACE_TEMPLATE)
    assert bool(ok[0]) == (Mo is not None)
    explicit = synth.make_landmarks(4, seed=1)
    M1, ok1 = ctx.estimate_affine_partial_2d(explicit, oracle.ARCFACE_TEMPLATE)     # explicit destination == template
    M2, ok2 = ctx.estimate_affine_partial_2d(explicit)
    np.testing.assert_array_equal(M1, M2)


def test_warp_golden_cv2(ctx, golden):
    for i in range(int(golden["warp_n"])):
        np.testing.assert_array_equal(ctx.warp_affine(golden["warp_in_%d" % i], golden["warp_M_%d" % i], (112, 112)), golden["warp_out_%d" % i])


@pytest.mark.parametrize("h,w", [(1080, 1920), (2160, 3840), (480, 641), (300, 200)])
def test_warp_bit_exact_vs_oracle(ctx, oracle, h, w):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    for t in range(4):
        s, th = rng.uniform(0.3, 3.0), rng.uniform(-0.7, 0.7)
        a, b = s * np.cos(th), s * np.sin(th)
        cx, cy = rng.uniform(-30, w + 30), rng.uniform(-30, h + 30)
        M = np.array([[a, -b, 56 - (a * cx - b * cy)], [b, a, 56 - (b * cx + a * cy)]], np.float64)
        np.testing.assert_array_equal(ctx.warp_affine(img, M, (112, 112)), oracle.warp_affine(img, M, (112, 112)))


@pytest.mark.parametrize("dsize", [(96, 80), (1, 5), (7, 1), (200, 131), (113, 112), (256, 256), (1024, 3)])
def test_warp_other_crop_sizes(ctx, oracle, dsize):
    """crop sizes other than the reference's 112x112 take the generic kernel (flattened bands, byte or packed stores)"""
    rng = np.random.default_rng(dsize[0] * 1000 + dsize[1])
    img = rng.integers(0, 256, (360, 500, 3), dtype=np.uint8)
    for t in range(3):
        s, th = rng.uniform(0.3, 2.5), rng.uniform(-0.7, 0.7)
        a, b = s * np.cos(th), s * np.sin(th)
        cx, cy = rng.uniform(-20, 520), rng.uniform(-20, 380)
        M = np.array([[a, -b, dsize[0] / 2 - (a * cx - b * cy)], [b, a, dsize[1] / 2 - (b * cx + a * cy)]], np.float64)
        np.testing.assert_array_equal(ctx.warp_affine(img, M, dsize), oracle.warp_affine(img, M, dsize))


def test_align_single_call(ctx, oracle):
    from rs_face_detection_b200.pipeline import FaceAlignment
    img = synth.make_frame(720, 1280, 5)
    pts = synth.make_landmarks(6, seed=3, frame_hw=(720, 1280))
    fa = FaceAlignment(ctx=ctx)
    for t in range(len(pts)):
        crop, M = oracle.align_face(img, pts[t])
        np.testing.assert_array_equal(fa.call(img, None, pts[t]), crop)


def test_align_batch_c4_shape(ctx, oracle):
    """BASELINE config 4 at reduced batch: frames + ~50 faces/frame, crops bit-exact vs the oracle."""
    B, per = 4, 50
    frames = [synth.make_frame(1080, 1920, 2000 + i) for i in range(B)]
    pts = synth.make_landmarks(B * per, seed=77).reshape(B * per, 10)
    fidx = np.repeat(np.arange(B, dtype=np.int32), per)
    devs = [ctx.to_device(f) for f in frames]
    F = B * per
    crops = ctx.alloc(F * 112 * 112 * 3)
    Md = ctx.alloc(F * 6 * 8)
    okd = ctx.alloc(F)
    ctx.align_batch([(d.ptr, 1080, 1920, 5760) for d in devs], ctx.to_device(pts), ctx.to_device(fidx), F, crops, Md, okd)
    ctx.synchronize()
    got = crops.download((F, 112, 112, 3), np.uint8)
    M = Md.download((F, 2, 3), np.float64)
    ok = okd.download((F,), np.uint8)
    outside = 0
    for f in range(F):
        crop, Mo = oracle.align_face(frames[fidx[f]], pts[f])
        assert bool(ok[f]) == (crop is not None)
        if crop is not None:
            np.testing.assert_array_equal(M[f], Mo)
            np.testing.assert_array_equal(got[f], crop)
            outside += int((crop == 0).all(-1).mean() > 0.2)
    assert outside > 0        # some crops hang off the frame border (BORDER_CONSTANT taps)


def test_round_trip_property(ctx):
    """Size-independent property: warping with the identity-scale transform that maps a 112x112 window to itself
    reproduces the source window exactly."""
    img = synth.make_frame(400, 500, 9)
    M = np.array([[1, 0, -100], [0, 1, -50]], np.float64)
    np.testing.assert_array_equal(ctx.warp_affine(img, M, (112, 112)), img[50:162, 100:212])

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
        crop, Mo, mode = oracle.align_face(frames[fidx[f]], pts[f], with_mode=True)
        assert int(ok[f]) == mode == 1
        np.testing.assert_array_equal(M[f], Mo)
        np.testing.assert_array_equal(got[f], crop)
        outside += int((crop == 0).all(-1).mean() > 0.2)
    assert outside > 0        # some crops hang off the frame border (BORDER_CONSTANT taps)


# ---- FaceAlignment::call's bbox-crop fallback (face_alignment.rs:64-116): taken when the estimate is empty ----------------
DEGENERATE = np.tile(np.array([[300.0, 200.0]], np.float32), (5, 1))          # five equal points: no 2-point sample is valid


def _fallback_cases(h, w):
    """(bbox, expected mode): 2 = the reference crops (x0,y0)..(W,H) and resizes, 0 = Mat::roi rejects the rectangle -> Err"""
    return [
        (None, 2),                                                           # bbox None -> the image inset by 1/16 (:67-72)
        (np.array([100.5, 60.25, 180.0, 140.0], np.float32), 2),
        (np.array([10.0, 5.0, 60.0, 70.0, 0.93], np.float32), 2),            # a (5,) detection row: x1-22 and y1-22 clamp to 0
        (np.array([w - 150.0, 30.0, w - 22.0, 90.0], np.float32), 2),        # x2 + 22 == W exactly: still inside
        (np.array([w - 150.0, 30.0, w - 21.5, 90.0], np.float32), 2),        # x2 + 22 = W + 0.5: `as i32` truncates back to W
        (np.array([w - 150.0, 30.0, w - 20.0, 90.0], np.float32), 0),        # x2 + 22 > W: `max` keeps it, roi is out of range
        (np.array([50.0, h - 21.0, 120.0, h - 5.0], np.float32), 0),         # det[1] + 22 > H (the :82 quirk uses det[1])
        (np.array([np.nan, 40.0, np.nan, 90.0], np.float32), 2),             # f32::max drops NaN; `as i32` maps NaN to 0
        (np.array([w + 40.0, 40.0, w - 100.0, 90.0], np.float32), 0),        # x0 >= W: negative width
    ]


@pytest.mark.parametrize("h,w", [(480, 640), (1080, 1920), (113, 131)])
def test_align_fallback_single(ctx, oracle, h, w):
    from rs_face_detection_b200 import FdError
    from rs_face_detection_b200.ffi import FD_ERR_ESTIMATE
    img = synth.make_frame(h, w, 3)
    seen = set()
    for bbox, mode in _fallback_cases(h, w):
        ecrop, eM, emode = oracle.align_face(img, DEGENERATE, bbox=bbox, with_mode=True)
        assert emode == mode or (h, w) == (113, 131)      # the expectations are written for frames wider than the boxes
        mode = emode
        seen.add(mode)
        if mode == 0:
            with pytest.raises(FdError) as e:
                ctx.align(img, DEGENERATE, bbox)
            assert e.value.code == FD_ERR_ESTIMATE
        else:
            crop, M, got_mode = ctx.align(img, DEGENERATE, bbox, with_mode=True)
            assert got_mode == 2 and M is None
            np.testing.assert_array_equal(crop, ecrop)
    assert seen == {0, 2}


def test_align_none_landmarks_is_an_error_like_the_reference(ctx):
    """landmarks == None: the reference hands an empty Mat to cv::estimateAffinePartial2D, which asserts -> Err"""
    from rs_face_detection_b200 import FdError
    with pytest.raises(FdError):
        ctx.align(synth.make_frame(100, 100, 1), None, np.array([1, 2, 30, 40], np.float32))


@pytest.mark.parametrize("crop", [(112, 112), (96, 128)])
def test_align_batch_mixed_fallback(oracle, crop):
    """one batch mixing warps, fallback crops and reference errors, on both warp kernels (112x112 and generic)"""
    from rs_face_detection_b200 import Context
    from rs_face_detection_b200.ffi import default_config
    cfg = default_config()
    cfg.crop_w, cfg.crop_h = crop
    c = Context(0, cfg)
    try:
        sizes = [(480, 640), (1080, 1920)]
        frames = [synth.make_frame(h, w, 40 + i) for i, (h, w) in enumerate(sizes)]
        devs = [c.to_device(f) for f in frames]
        good = synth.make_landmarks(6, seed=5, frame_hw=(480, 640))
        lmk, fidx, bbs = [], [], []
        for b, (h, w) in enumerate(sizes):
            for bbox, _ in _fallback_cases(h, w)[1:]:
                lmk.append(DEGENERATE); fidx.append(b); bbs.append(bbox[:4])
            for t in range(3):
                lmk.append(good[3 * b + t]); fidx.append(b); bbs.append(np.array([10, 10, 50, 50], np.float32))
        lmk = np.stack(lmk).reshape(-1, 10).astype(np.float32)
        fidx = np.array(fidx, np.int32)
        bbs = np.stack(bbs).astype(np.float32)
        F = len(lmk)
        nb = crop[0] * crop[1] * 3
        crops, okd = c.alloc(F * nb), c.alloc(F)
        c.align_batch([(d.ptr, f.shape[0], f.shape[1], f.strides[0]) for d, f in zip(devs, frames)], c.to_device(lmk), c.to_device(fidx),
                      F, crops, None, okd, bbox_dev=c.to_device(bbs))
        c.synchronize()
        got = crops.download((F, crop[1], crop[0], 3), np.uint8)
        mode = okd.download((F,), np.uint8)
        seen = set()
        for f in range(F):
            ecrop, _, emode = oracle.align_face(frames[fidx[f]], lmk[f], bbox=bbs[f], dsize=crop, with_mode=True)
            assert int(mode[f]) == emode
            seen.add(emode)
            np.testing.assert_array_equal(got[f], ecrop if ecrop is not None else np.zeros_like(got[f]))
        assert seen == {0, 1, 2}
    finally:
        c.close()


def test_round_trip_property(ctx):
    """Size-independent property: warping with the identity-scale transform that maps a 112x112 window to itself
    reproduces the source window exactly."""
    img = synth.make_frame(400, 500, 9)
    M = np.array([[1, 0, -100], [0, 1, -50]], np.float64)
    np.testing.assert_array_equal(ctx.warp_affine(img, M, (112, 112)), img[50:162, 100:212])

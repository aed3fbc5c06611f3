"""FaceSelection::call (face_selection.rs:72-189) through the C ABI vs the oracle, single image and batched
on the device between NMS and the warp (detect -> select -> align, face_pipeline/pipeline.rs:196-232)."""
import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu


def _random_boxes(rng, M, h, w):
    cx, cy = rng.uniform(-20, w + 20, M), rng.uniform(-20, h + 20, M)
    bw, bh = rng.uniform(1, 0.5 * w, M), rng.uniform(1, 0.5 * h, M)
    fb = np.stack([cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, rng.uniform(0, 1, M)], 1).astype(np.float32)
    if M > 3:      # near-duplicates (key-point row != box row), exact ties, degenerate boxes
        fb[M // 2] = fb[0] + np.float32([1.5, -1.0, 0.5, 2.0, 0])
        fb[M // 3, :4] = fb[1, :4]
        fb[M - 1, 2] = fb[M - 1, 0] - 5
    return fb


@pytest.mark.parametrize("hw", [(1080, 1920), (2160, 3840), (480, 300)])
def test_face_selection_vs_oracle(ctx, oracle, hw):
    rng = np.random.default_rng(hw[0])
    for M in (0, 1, 2, 5, 31, 32, 33, 200, 1500):
        for trial in range(4):
            fb = _random_boxes(rng, M, *hw)
            kps = rng.uniform(0, hw[1], (M, 5, 2)).astype(np.float32)
            for enroll in (False, True):
                for kp in (kps, None):
                    assert ctx.face_selection(hw, fb, kp, enroll) == oracle.face_selection(hw, fb, kp, enroll), (M, trial, enroll, kp is None)
    # other ratios than the reference defaults
    fb = _random_boxes(rng, 64, *hw)
    prm = (0.1, 0.45, 0.02, 0.05)
    assert ctx.face_selection(hw, fb, np.zeros((64, 5, 2)), False, prm) == oracle.face_selection(hw, fb, np.zeros((64, 5, 2)), False, prm)


def test_python_mirror(ctx, oracle):
    from rs_face_detection_b200.pipeline import FaceSelection
    rng = np.random.default_rng(5)
    fb = _random_boxes(rng, 40, 1080, 1920)
    kps = rng.uniform(0, 1920, (40, 5, 2)).astype(np.float32)
    img = np.zeros((1080, 1920, 3), np.uint8)
    box, kp = FaceSelection(ctx=ctx).call(img, fb, kps)
    bi, ki = oracle.face_selection((1080, 1920), fb, kps)
    np.testing.assert_array_equal(box, fb[bi])
    np.testing.assert_array_equal(kp, kps[ki])
    assert FaceSelection(ctx=ctx).call(img, fb[:0], kps[:0]) == (None, None)


@pytest.mark.parametrize("enroll", [False, True])
def test_detect_select_align_on_device(ctx, oracle, enroll):
    """The reference's per-image flow at batch granularity: detect -> FaceSelection -> FaceAlignment of the selected face,
    nothing but the crops and the (B,2) selection leaving the device."""
    B = 6
    frames = [synth.make_frame(1080, 1920, 100 + i) for i in range(B)]
    heads, _ = synth.make_heads(B, seed=77, n_faces=8, content_hw=(360, 640))
    for h in heads[0::3]:
        h[4, 2:] = 0.0        # image 4: no detections -> no selection -> zero crop, ok = 0
        h[4, :2] = 1.0
    fdev = [ctx.to_device(f) for f in frames]
    fl = [(d.ptr, 1080, 1920, 5760) for d in fdev]
    devs = [ctx.to_device(h) for h in heads]
    ds = np.full(B, 1 / 3, np.float32)
    ctx.detect_batch(devs, B, ds, 0.7, 0.4)
    sel = ctx.select_detections(fl, is_enroll=enroll)
    crops, okd = ctx.alloc(B * 112 * 112 * 3), ctx.alloc(B)
    ctx.align_selected(fl, crops, None, okd)
    counts, det, lmk = ctx.detect_fetch(B)
    got = crops.download((B, 112, 112, 3), np.uint8)
    ok = okd.download((B,), np.uint8)
    off = 0
    for b in range(B):
        d, l = det[off:off + counts[b]], lmk[off:off + counts[b]].reshape(-1, 5, 2)
        bi, ki = oracle.face_selection((1080, 1920), d, l, enroll)
        assert (sel[b, 0], sel[b, 1]) == ((off + bi) if bi >= 0 else -1, (off + ki) if ki >= 0 else -1)
        if ki >= 0:
            crop, _ = oracle.align_face(frames[b], l[ki])
            assert bool(ok[b]) == (crop is not None)
            if crop is not None:
                np.testing.assert_array_equal(got[b], crop)
        else:
            assert ok[b] == 0 and not got[b].any()
        off += counts[b]
    assert counts[4] == 0 and sel[4, 0] == -1

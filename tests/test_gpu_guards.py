"""Memory-safety checks without compute-sanitizer (closed on this pool): every device OUTPUT buffer is embedded between
canary regions that must come back untouched, and size-independent properties are checked at BASELINE sizes."""
import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu
GUARD = 4096


class Guarded:
    """device buffer [canary | payload | canary]"""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, nbytes
        self.pad = (-nbytes) % 256
        self.buf = ctx.alloc(GUARD + nbytes + self.pad + GUARD)
        self.buf.upload(np.full(GUARD + nbytes + self.pad + GUARD, 0xCD, np.uint8))
        self.ptr = self.buf.ptr + GUARD

    def check(self):
        raw = self.buf.download((GUARD + self.nbytes + self.pad + GUARD,), np.uint8)
        assert (raw[:GUARD] == 0xCD).all(), "write before the buffer"
        assert (raw[GUARD + self.nbytes:] == 0xCD).all(), "write past the buffer"
        return raw[GUARD:GUARD + self.nbytes]


def test_preprocess_output_bounds(ctx):
    shapes = [(1080, 1920), (97, 131), (2160, 3840), (641, 643)]
    imgs = [synth.make_frame(h, w, 50 + i) for i, (h, w) in enumerate(shapes)]
    devs = [ctx.to_device(im) for im in imgs]
    out = Guarded(ctx, len(imgs) * 3 * 640 * 640 * 4)
    ctx.preprocess_batch([(d.ptr, im.shape[0], im.shape[1], im.strides[0]) for d, im in zip(devs, imgs)], out.ptr)
    ctx.synchronize()
    t = out.check().view(np.float32).reshape(len(imgs), 3, 640, 640)
    assert np.isfinite(t).all() and t.min() >= 0 and t.max() <= 255


def test_align_output_bounds(ctx):
    B, F = 2, 37
    frames = [synth.make_frame(480, 640, 5 + i) for i in range(B)]
    devs = [ctx.to_device(f) for f in frames]
    pts = synth.make_landmarks(F, seed=9, frame_hw=(480, 640)).reshape(F, 10)
    fidx = (np.arange(F) % B).astype(np.int32)
    crops, M, ok = Guarded(ctx, F * 112 * 112 * 3), Guarded(ctx, F * 48), Guarded(ctx, F)
    ctx.align_batch([(d.ptr, 480, 640, 1920) for d in devs], ctx.to_device(pts), ctx.to_device(fidx), F, crops.ptr, M.ptr, ok.ptr)
    ctx.synchronize()
    crops.check(); M.check()
    assert ok.check().sum() > 0


@pytest.mark.parametrize("n", [300, 5000, 100000])
def test_nms_output_bounds_and_idempotence(ctx, oracle, n):
    dets = synth.make_crowd_boxes(n, seed=n, n_faces=max(1, n // 20))
    d = ctx.to_device(dets)
    keep, num = Guarded(ctx, 4 * n), Guarded(ctx, 8)
    ctx.nms_device(d, n, 0.4, keep.ptr, num.ptr)
    ctx.synchronize()
    cnt = int(num.check().view(np.int32)[0])
    k = keep.check().view(np.int32)[:cnt]
    assert len(set(k.tolist())) == cnt and k.min() >= 0 and k.max() < n
    # idempotence: the kept boxes are mutually compatible, so NMS over them alone keeps every one, in the same order
    again = ctx.nms(dets[k], 0.4)
    np.testing.assert_array_equal(again, np.arange(cnt))
    # kept scores are non-increasing (pick order)
    assert np.all(np.diff(dets[k, 4]) <= 0)


def test_full_size_property_4k(ctx):
    """4K -> 640x360 is exactly scale 6: fractions 0.5 on both axes = the rounded 2x2 average of pixels (6d+2, 6d+3)
    through OpenCV's fixed-point formula; padding rows are zero."""
    img = synth.make_frame(2160, 3840, 123)
    got, sc = ctx.preprocess(img)
    assert sc == np.float32(360 / 2160)
    a = img[2::6, 2::6].astype(np.int32); b = img[2::6, 3::6].astype(np.int32)
    c = img[3::6, 2::6].astype(np.int32); d = img[3::6, 3::6].astype(np.int32)
    # T = (p0+p1)*1024 ; ((1024*(T>>4))>>16) = (p0+p1) ; ((top + bottom) + 2) >> 2
    exp = ((a + b) + (c + d) + 2) >> 2
    np.testing.assert_array_equal(got[0, :, :360], exp[:, :, ::-1].transpose(2, 0, 1).astype(np.float32))
    assert not got[0, :, 360:].any()


def test_detect_results_bounds_via_pipeline(ctx):
    """cap_rows smaller than the number of faces must be reported, not overrun."""
    from rs_face_detection_b200 import FdError, ffi
    frames = [synth.make_frame(270, 480, 1)]
    heads, _ = synth.make_heads(1, seed=11, n_faces=12, content_hw=(360, 640))
    with pytest.raises(FdError) as e:
        ctx.pipeline_host(frames, heads, cap_rows=2, conf_thr=0.7, iou_thr=0.4)
    assert e.value.code == ffi.FD_ERR_CAPACITY

"""Preprocess parity: letterbox resize + pad + BGR->RGB + normalise -> NCHW fp32, bit-exact vs the oracle (which is
pinned bit-exact against cv2)."""
import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu

SHAPES = [(1080, 1920), (2160, 3840), (480, 640), (720, 1280), (1000, 700), (333, 517), (97, 131), (640, 640), (1280, 1280),
          (50, 2000), (3000, 40), (7, 9), (641, 643)]


def _oracle_tensor(oracle, img, cfg=None):
    det, sc = oracle.preprocess_letterbox(img, (640, 640))
    if cfg is None:
        return oracle.to_tensor(det), sc
    return oracle.to_tensor(det, cfg.pixel_scale, list(cfg.pixel_means), list(cfg.pixel_stds)), sc


@pytest.mark.parametrize("h,w", SHAPES)
def test_single_image(ctx, oracle, h, w):
    rng = np.random.default_rng(h * 3 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got, sc = ctx.preprocess(img)
    exp, sce = _oracle_tensor(oracle, img)
    assert sc == sce
    np.testing.assert_array_equal(got, exp)
    assert ctx.letterbox_geometry(h, w) == oracle.letterbox_geometry(h, w)


def test_unaligned_pitch_and_pointer(ctx, oracle):
    rng = np.random.default_rng(9)
    big = rng.integers(0, 256, (300, 411 * 3 + 7), dtype=np.uint8)
    img = big[:, 5:5 + 401 * 3].reshape(300, 401, 3)       # pitch 1240 (not a multiple of 16), base offset 5
    dev = ctx.to_device(big)
    out = ctx.alloc(3 * 640 * 640 * 4)
    ds = ctx.preprocess_batch([(dev.ptr + 5, 300, 401, big.shape[1])], out)
    got = out.download((1, 3, 640, 640), np.float32)
    exp, sce = _oracle_tensor(oracle, np.ascontiguousarray(img))
    assert ds[0] == sce
    np.testing.assert_array_equal(got, exp)


def test_batch_mixed_sizes(ctx, oracle):
    shapes = [(1080, 1920), (2160, 3840), (480, 640), (1000, 700), (97, 131), (1080, 1920)]
    imgs = [synth.make_frame(h, w, 2000 + i) for i, (h, w) in enumerate(shapes)]
    devs = [ctx.to_device(im) for im in imgs]
    out = ctx.alloc(len(imgs) * 3 * 640 * 640 * 4)
    ds = ctx.preprocess_batch([(d.ptr, im.shape[0], im.shape[1], im.strides[0]) for d, im in zip(devs, imgs)], out)
    got = out.download((len(imgs), 3, 640, 640), np.float32)
    for i, im in enumerate(imgs):
        exp, sce = _oracle_tensor(oracle, im)
        assert ds[i] == sce
        np.testing.assert_array_equal(got[i], exp[0])


def test_normalisation_constants(oracle):
    """Non-identity mean/std/scale: the general (p/scale - mean)/std expression (face_detection.rs:227)."""
    from rs_face_detection_b200 import Context, default_config
    cfg = default_config()
    cfg.pixel_scale = 255.0
    for i, (m, s) in enumerate(zip((0.406, 0.456, 0.485), (0.225, 0.224, 0.229))):
        cfg.pixel_means[i], cfg.pixel_stds[i] = m, s
    c = Context(0, cfg)
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (360, 500, 3), dtype=np.uint8)
    got, _ = c.preprocess(img)
    exp, _ = _oracle_tensor(oracle, img, cfg)
    np.testing.assert_array_equal(got, exp)
    c.close()


@pytest.mark.parametrize("src,dst", [((1080, 1920), (640, 360)), ((333, 517), (640, 412)), ((97, 131), (640, 473)), ((64, 64), (64, 64)),
                                     ((90, 120), (40, 30))])
def test_resize_linear(ctx, oracle, src, dst):
    rng = np.random.default_rng(src[0] + dst[0])
    img = rng.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    np.testing.assert_array_equal(ctx.resize_linear(img, dst), oracle.resize_linear(img, dst))


def test_golden_resize_through_gpu(ctx, golden):
    for i in range(int(golden["resize_n"])):
        dw, dh = golden["resize_dsize_%d" % i]
        np.testing.assert_array_equal(ctx.resize_linear(golden["resize_in_%d" % i], (int(dw), int(dh))), golden["resize_out_%d" % i])


def test_full_size_property_1080p(ctx):
    """At BASELINE size without the oracle: 1080p -> 640x360 is exactly scale 3 == point sampling src[3y+1][3x+1];
    the padding rows are the normalised zero."""
    img = synth.make_frame(1080, 1920, 77)
    got, sc = ctx.preprocess(img)
    assert sc == np.float32(360 / 1080)
    np.testing.assert_array_equal(got[0, :, :360], img[1::3, 1::3, ::-1].transpose(2, 0, 1).astype(np.float32))
    assert not got[0, :, 360:].any()

"""post-align model preprocessors (face_extraction.rs:38-77, face_quality.rs:43-101,
face_quality_assessment.rs:48-88) through the C ABI vs the oracle: bit-exact (no transcendental involved)."""
import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["face_extraction", "face_quality", "face_quality_assessment"])
def test_single_image_112(ctx, oracle, model):
    mean, mul = oracle.MODEL_NORMS[model]
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (112, 112, 3), dtype=np.uint8)
    np.testing.assert_array_equal(ctx.model_preprocess(img, (112, 112), mean, mul), oracle.model_preprocess(img, (112, 112), mean, mul))


@pytest.mark.parametrize("hw,out", [((200, 150), (112, 112)), ((64, 64), (112, 112)), ((113, 97), (100, 90)), ((112, 112), (224, 224))])
def test_with_resize(ctx, oracle, hw, out):
    mean, mul = oracle.MODEL_NORMS["face_quality"]
    rng = np.random.default_rng(hw[0])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    np.testing.assert_array_equal(ctx.model_preprocess(img, out, mean, mul), oracle.model_preprocess(img, out, mean, mul))


def test_fused_on_aligned_crops(ctx, oracle):
    """frames -> align -> model tensor, all on the device: the crops never visit the host."""
    B, F = 2, 24
    frames = [synth.make_frame(480, 640, 5 + i) for i in range(B)]
    devs = [ctx.to_device(f) for f in frames]
    pts = synth.make_landmarks(F, seed=9, frame_hw=(480, 640)).reshape(F, 10)
    fidx = (np.arange(F) % B).astype(np.int32)
    crops = ctx.alloc(F * 112 * 112 * 3)
    ctx.align_batch([(d.ptr, 480, 640, 1920) for d in devs], ctx.to_device(pts), ctx.to_device(fidx), F, crops)
    mean, mul = oracle.MODEL_NORMS["face_extraction"]
    out = ctx.alloc(F * 3 * 112 * 112 * 4)
    ctx.crops_to_tensor(crops, F, (112, 112), (112, 112), mean, mul, out)
    ctx.synchronize()
    got = out.download((F, 3, 112, 112), np.float32)
    c = crops.download((F, 112, 112, 3), np.uint8)
    for f in range(F):
        np.testing.assert_array_equal(got[f], oracle.model_preprocess(c[f], (112, 112), mean, mul))
    assert got.min() >= -1.0 and got.max() <= 1.0

"""SURVEY 8(f) N4 — byte_data_to_opencv (utils.rs:8-52 = cv::imdecode): host Huffman pass + CUDA IDCT / upsampling / colour
kernels, bit-exact against the cv2-generated golden vectors and the CPU oracle; decoded frames feed the detection path directly."""
import os

import numpy as np
import pytest

from rs_face_detection_b200.utils import synth

pytestmark = pytest.mark.gpu
try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def _golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_golden.npz"))


def test_imdecode_golden_cv2(ctx):
    g = _golden()
    for i in range(int(g["n"])):
        np.testing.assert_array_equal(ctx.imdecode(g["jpeg_%d" % i].tobytes()), g["bgr_%d" % i])


def test_byte_data_to_opencv_mirror(ctx):
    from rs_face_detection_b200.utils.utils import byte_data_to_opencv
    g = _golden()
    np.testing.assert_array_equal(byte_data_to_opencv(g["jpeg_7"].tobytes(), ctx), g["bgr_7"])


def test_imdecode_unsupported_streams_are_errors(ctx):
    from rs_face_detection_b200 import FdError
    g = _golden()
    for k in ("unsupported_progressive", "unsupported_gray"):
        with pytest.raises(FdError):
            ctx.imdecode(g[k].tobytes())
    with pytest.raises(FdError):
        ctx.imdecode(b"\x00\x01\x02\x03 not a jpeg")
    with pytest.raises(FdError):
        ctx.imdecode(g["jpeg_0"].tobytes()[:40])        # truncated before SOS


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_imdecode_vs_oracle_and_cv2_at_frame_size(ctx, oracle):
    """the bench's 1080p frame and a 4K frame, 4:2:0 and 4:4:4: GPU == oracle == cv2.imdecode"""
    for (h, w, ss, q) in [(1080, 1920, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 90), (1080, 1920, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 75),
                          (2160, 3840, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 85), (1081, 1923, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 95)]:
        ok, buf = cv2.imencode(".jpg", synth.make_frame(h, w, h + q), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss])
        want = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        np.testing.assert_array_equal(oracle.jpeg_decode(buf.tobytes()), want)
        np.testing.assert_array_equal(ctx.imdecode(buf.tobytes()), want)


def test_decode_batch_feeds_the_detection_path(ctx, oracle):
    """JPEG bytes -> device frames (mixed sizes and samplings in one batch) -> fd_preprocess_batch: same CNN input tensors as the
    oracle's decode + letterbox + tensor path; and the frames themselves equal the oracle's decode."""
    g = _golden()
    idx = [0, 7, 8, 1, 9]
    jpegs = [g["jpeg_%d" % i] for i in idx]
    for threads in (1, 3):
        frames = ctx.decode_jpeg_batch(jpegs, n_threads=threads)
        B = len(idx)
        tensor = ctx.alloc(B * 3 * 640 * 640 * 4)
        ds = ctx.preprocess_batch(frames, tensor)
        ctx.synchronize()
        t = tensor.download((B, 3, 640, 640), np.float32)
        for b, i in enumerate(idx):
            want = g["bgr_%d" % i]
            assert (frames[b].height, frames[b].width) == want.shape[:2]
            row = np.empty((want.shape[0], frames[b].pitch), np.uint8)
            import ctypes as C
            ctx.lib.fd_memcpy_d2h(ctx.handle, row.ctypes.data_as(C.c_void_p), C.c_void_p(frames[b].data), C.c_size_t(row.nbytes))
            np.testing.assert_array_equal(row[:, :want.shape[1] * 3].reshape(want.shape), want)
            det_img, sc = oracle.preprocess_letterbox(want)
            np.testing.assert_array_equal(t[b], oracle.to_tensor(det_img)[0])
            assert ds[b] == sc
        tensor.free()


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable (only to ENCODE the test streams)")
def test_pipeline_host_jpeg_equals_pipeline_on_decoded_frames(ctx, oracle):
    """fd_pipeline_host_jpeg(JPEG bytes) == fd_pipeline_host(cv2.imdecode of the same bytes): detections, landmarks, crops"""
    B = 3
    frames = [synth.make_frame(1080, 1920, 2600 + i) for i in range(B - 1)] + [synth.make_frame(720, 1280, 2650)]
    streams = [np.asarray(cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 88])[1], np.uint8).ravel() for f in frames]
    decoded = [cv2.imdecode(s, cv2.IMREAD_UNCHANGED) for s in streams]
    heads, _ = synth.make_heads(B, seed=3400, n_faces=8, content_hw=(360, 640))
    ref, total, _, _ = ctx.pipeline_host(decoded, heads, cap_rows=B * 64, conf_thr=0.7, iou_thr=0.4, want_tensor=True)
    got, total2, h2d, _ = ctx.pipeline_host(streams, heads, cap_rows=B * 64, conf_thr=0.7, iou_thr=0.4, want_tensor=True, jpeg=True, jpeg_threads=2)
    assert total2 == total > 0 and h2d > 0
    for k in ("counts", "tensor", "det_scale"):
        np.testing.assert_array_equal(got[k], ref[k])
    for k in ("det", "lmk", "crops", "align_mode"):
        np.testing.assert_array_equal(got[k][:total], ref[k][:total])
    crop, _ = oracle.align_face(oracle.jpeg_decode(streams[0].tobytes()), got["lmk"][0])
    np.testing.assert_array_equal(got["crops"][0], crop)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable (only to ENCODE the test streams)")
def test_restart_marker_streams_are_huffman_decoded_on_the_device(ctx):
    """Streams with short restart intervals take jpeg_huffman_kernel (one interval per thread); streams without markers, or with
    intervals of 32 MCUs and more, the self-synchronising sub-sequence decoder (restart markers as known-state boundaries); only
    compressed bytes cross PCIe.  All in one batch, every frame bit-identical to cv2.imdecode."""
    import ctypes as C
    SS = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    cases = [(1080, 1920, 0, 16, 90), (1080, 1920, 0, 0, 90), (720, 1280, 1, 1, 75), (333, 517, 2, 7, 95), (2160, 3840, 0, 240, 85),
             (64, 48, 0, 1000, 60), (17, 33, 0, 2, 100), (480, 640, 0, 40, 30)]
    streams, want = [], []
    for i, (h, w, ss, rst, q) in enumerate(cases):
        params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss]]
        if rst:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        if i == 3:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        buf = np.asarray(cv2.imencode(".jpg", synth.make_frame(h, w, 900 + i), params)[1], np.uint8).ravel()
        streams.append(buf)
        want.append(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED))
    frames = ctx.decode_jpeg_batch(streams, n_threads=2)
    ctx.synchronize()
    st = ctx.jpeg_last_stats()
    assert st["device_entropy_images"] == len(cases) and st["host_entropy_images"] == 0 and st["selfsync_images"] == 4
    assert st["h2d_bytes"] < sum(s.size for s in streams) + 1e6               # only compressed bytes (+ tables) cross PCIe
    for b, w_ in enumerate(want):
        row = np.empty((w_.shape[0], frames[b].pitch), np.uint8)
        ctx.lib.fd_memcpy_d2h(ctx.handle, row.ctypes.data_as(C.c_void_p), C.c_void_p(frames[b].data), C.c_size_t(row.nbytes))
        np.testing.assert_array_equal(row[:, :w_.shape[1] * 3].reshape(w_.shape), w_, err_msg="case %d %r" % (b, cases[b]))
    for s_, w_ in zip(streams[:3], want[:3]):                                  # and through the single-image call
        np.testing.assert_array_equal(ctx.imdecode(s_.tobytes()), w_)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable (only to ENCODE the test streams)")
def test_long_restart_intervals_take_the_selfsync_decoder(ctx):
    """DRI of one MCU row and more: too few intervals for one thread each, so the markers become boundaries of the
    self-synchronising chain (dropped by the device unstuffing, DC prediction segmented at them).  Bit-identical to cv2."""
    import ctypes as C
    SS = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    cases = [(1080, 1920, 0, 120, 90), (1080, 1920, 0, 480, 90), (1080, 1920, 0, 32, 50), (720, 1280, 1, 80, 75), (333, 517, 2, 33, 95),
             (2160, 3840, 0, 240, 85), (601, 799, 0, 100000, 100), (480, 640, 2, 64, 10), (1080, 1920, 0, 121, 97)]
    streams, want = [], []
    for i, (h, w, ss, rst, q) in enumerate(cases):
        params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss], cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        img = synth.make_frame(h, w, 1300 + i) if i != 6 else np.random.default_rng(3).integers(0, 256, (h, w, 3), dtype=np.uint8)
        buf = np.asarray(cv2.imencode(".jpg", img, params)[1], np.uint8).ravel()
        streams.append(buf)
        want.append(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED))
    for group in (list(range(len(cases))), [0], [6, 1]):
        frames = ctx.decode_jpeg_batch([streams[i] for i in group], n_threads=2)
        ctx.synchronize()
        st = ctx.jpeg_last_stats()
        assert st["selfsync_images"] == len(group) and st["host_entropy_images"] == 0, st
        for b, i in enumerate(group):
            w_ = want[i]
            row = np.empty((w_.shape[0], frames[b].pitch), np.uint8)
            ctx.lib.fd_memcpy_d2h(ctx.handle, row.ctypes.data_as(C.c_void_p), C.c_void_p(frames[b].data), C.c_size_t(row.nbytes))
            np.testing.assert_array_equal(row[:, :w_.shape[1] * 3].reshape(w_.shape), w_, err_msg="case %d %r" % (i, cases[i]))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable (only to ENCODE the test streams)")
def test_every_restart_interval_through_the_selfsync_decoder():
    """FD_JPEG_RST_SYNC_MIN=1: also the short intervals (1, 2, 7, 16 MCUs: several boundaries inside one sub-sequence)."""
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, cv2\n"
            "from rs_face_detection_b200 import Context\n"
            "from rs_face_detection_b200.utils import synth\n"
            "c = Context(0)\n"
            "SS = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]\n"
            "for i, (h, w, ss, rst, q) in enumerate([(1080, 1920, 0, 1, 90), (720, 1280, 1, 2, 75), (333, 517, 2, 7, 95), (1080, 1920, 0, 16, 90), (64, 48, 0, 3, 60), (17, 33, 0, 2, 100)]):\n"
            "    buf = np.asarray(cv2.imencode('.jpg', synth.make_frame(h, w, 1500 + i), [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[ss], cv2.IMWRITE_JPEG_RST_INTERVAL, rst])[1], np.uint8).ravel()\n"
            "    np.testing.assert_array_equal(c.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_UNCHANGED), err_msg=str((h, w, ss, rst, q)))\n"
            "    assert c.jpeg_last_stats()['selfsync_images'] == 1\n"
            "    if i in (0, 2, 3):\n"                      # corrupt copies (bit flips, damaged / missing markers, truncation): return, never wedge
            "        rng = np.random.default_rng(40 + i)\n"
            "        for trial in range(8):\n"
            "            bad = buf.copy(); start = len(bad) // 3\n"
            "            if trial < 3:\n"
            "                for pos in rng.integers(start, len(bad) - 2, 12): bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))\n"
            "            elif trial < 6:\n"
            "                ff = np.flatnonzero((bad[start:-2] == 0xFF) & ((bad[start + 1:-1] & 0xF8) == 0xD0)) + start\n"
            "                for pos in rng.choice(ff, min(3, len(ff)), replace=False): bad[pos + 1] = [0xD9, 0x00, 0xD0 | int(rng.integers(0, 8))][trial - 3]\n"
            "            else:\n"
            "                bad = bad[:int(rng.integers(start, len(bad) - 2))]\n"
            "            try:\n"
            "                assert c.imdecode(bad.tobytes()).shape == (h, w, 3)\n"
            "            except Exception as e:\n"
            "                assert type(e).__name__ == 'FdError', e\n"
            "        np.testing.assert_array_equal(c.imdecode(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_UNCHANGED))\n"
            "print('ok')\n" % (root,))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FD_JPEG_RST_SYNC_MIN="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-1500:]


def test_corrupted_streams_fail_or_decode_but_never_wedge_the_context(ctx):
    """Bit flips / truncations in the entropy-coded segment (with and without restart markers): the call returns — an error
    or some image of the right shape — and the context keeps decoding good streams bit-exactly afterwards."""
    from rs_face_detection_b200 import FdError
    g = _golden()
    rng = np.random.default_rng(11)
    for i in (0, 2, 7, 8):                       # 2 and 8 carry restart markers (device Huffman), 0 and 7 do not (host Huffman)
        good = g["jpeg_%d" % i]
        want = g["bgr_%d" % i]
        for trial in range(6):
            bad = good.copy()
            start = len(bad) // 2                # well inside the entropy-coded segment
            if trial < 4:
                for pos in rng.integers(start, len(bad) - 2, 12):
                    bad[pos] ^= np.uint8(1 << int(rng.integers(0, 8)))
            else:
                bad = bad[:int(rng.integers(start, len(bad) - 2))]
            try:
                out = ctx.imdecode(bad.tobytes())
                assert out.shape == want.shape
            except FdError:
                pass
        np.testing.assert_array_equal(ctx.imdecode(good.tobytes()), want)


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable (only to ENCODE the test streams)")
def test_selfsync_decoder_streams_without_restart_markers(ctx):
    """Ordinary JPEG files (no DRI): the sub-sequence states must reach their fixed point in a few rounds and the frames must be
    bit-identical to cv2.imdecode — smooth and noisy content, all samplings, tiny and 4K images, optimised Huffman tables."""
    import ctypes as C
    SS = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    rng = np.random.default_rng(3)
    imgs = [synth.make_frame(1080, 1920, 31), synth.make_frame(2160, 3840, 32), synth.make_frame(333, 517, 33), synth.make_frame(8, 8, 34),
            np.full((480, 640, 3), 128, np.uint8), (np.indices((720, 1280)).sum(0)[..., None] // 8 % 256).astype(np.uint8).repeat(3, 2),
            rng.integers(0, 256, (97, 1201, 3), dtype=np.uint8), synth.make_frame(1, 1, 35)]
    streams, want = [], []
    for i, im in enumerate(imgs):
        params = [cv2.IMWRITE_JPEG_QUALITY, [90, 75, 95, 50, 90, 85, 100, 90][i], cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SS[i % 3]]
        if i == 2:
            params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
        buf = np.asarray(cv2.imencode(".jpg", np.ascontiguousarray(im), params)[1], np.uint8).ravel()
        streams.append(buf)
        want.append(cv2.imdecode(buf, cv2.IMREAD_UNCHANGED))
    # textured content synchronises within a sub-sequence or two: the fixed point is confirmed after a handful of rounds.
    # Two kinds of stream synchronise slowly and are repaired from the true start, up to one sub-sequence per round (still exact,
    # rounds <= number of sub-sequences): PERIODIC ones (a constant image, a pure ramp: every block codes to the same few bits, a
    # decoder in the wrong phase stays there) and streams without end-of-block codes (uniform noise at quality 100: all 63 AC
    # coefficients of every block are coded, so nothing resets the coefficient position of a misaligned decoder).
    for group, max_rounds in (([0, 1, 2, 3, 7], 8), ([4, 5, 6], 800), (list(range(len(imgs))), 800)):
        frames = ctx.decode_jpeg_batch([streams[i] for i in group], n_threads=4)
        ctx.synchronize()
        st = ctx.jpeg_last_stats()
        assert st["selfsync_images"] == len(group) and st["host_entropy_images"] == 0
        assert 1 <= st["selfsync_rounds"] <= max_rounds, st
        for b, i in enumerate(group):
            w_ = want[i]
            row = np.empty((w_.shape[0], frames[b].pitch), np.uint8)
            ctx.lib.fd_memcpy_d2h(ctx.handle, row.ctypes.data_as(C.c_void_p), C.c_void_p(frames[b].data), C.c_size_t(row.nbytes))
            np.testing.assert_array_equal(row[:, :w_.shape[1] * 3].reshape(w_.shape), w_, err_msg="image %d" % i)


def test_host_huffman_path_still_bit_exact():
    """FD_JPEG_HOST_HUFFMAN=1 forces the serial host Huffman pass (what streams with > 2 tables per class take): same frames."""
    import subprocess, sys
    code = ("import os, sys, numpy as np; sys.path.insert(0, %r); from rs_face_detection_b200 import Context; c = Context(0); "
            "g = np.load(%r); "
            "[np.testing.assert_array_equal(c.imdecode(g['jpeg_%%d' %% i].tobytes()), g['bgr_%%d' %% i]) for i in range(int(g['n']))]; "
            "assert c.jpeg_last_stats()['host_entropy_images'] == 1; print('ok')"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jpeg_golden.npz")))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, FD_JPEG_HOST_HUFFMAN="1"), capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-800:]

"""Host-side logic that needs no GPU: the library's init-time anchor tables vs the oracle, the synthetic workload
generator, and the bench sharding helpers."""
import numpy as np

from rs_face_detection_b200 import ffi
from rs_face_detection_b200.processing import generate_anchors as GA
from rs_face_detection_b200.utils import synth


def test_anchor_tables_match_oracle(oracle):
    np.testing.assert_array_equal(ffi.generate_anchors(16, [0.5, 1.0, 2.0], [8.0, 16.0, 32.0]),
                                  oracle.generate_anchors(16, [0.5, 1.0, 2.0], [8.0, 16.0, 32.0]))
    np.testing.assert_array_equal(ffi.generate_anchors2(16, [0.5, 1.0, 2.0], [8.0, 16.0, 32.0], 16, True),
                                  oracle.generate_anchors2(16, [0.5, 1.0, 2.0], [8.0, 16.0, 32.0], 16, True))
    a = ffi.generate_anchors_fpn([64, 32, 16, 8, 4], [0.5, 1.0, 2.0, 1.0, 1.0], [8.0] * 5)
    b = oracle.generate_anchors_fpn([64, 32, 16, 8, 4], [0.5, 1.0, 2.0, 1.0, 1.0], [8.0] * 5)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    fpn2 = ffi.generate_anchors_fpn2(False, GA.RETINAFACE_ANCHOR_CFG)
    np.testing.assert_array_equal(np.stack(fpn2), oracle.generate_anchors_fpn2_retinaface(False))
    np.testing.assert_array_equal(np.stack(fpn2), synth.BASE_ANCHORS)
    cfg = ffi.default_config()
    base = np.array(list(cfg.base_anchors), np.float32).reshape(8, 4, 4)[:3, :2]
    np.testing.assert_array_equal(base, synth.BASE_ANCHORS)
    # dict order must not matter: strides are sorted descending (generate_anchors.rs:123-124)
    rev = dict(reversed(list(GA.RETINAFACE_ANCHOR_CFG.items())))
    np.testing.assert_array_equal(np.stack(ffi.generate_anchors_fpn2(False, rev)), synth.BASE_ANCHORS)


def test_synth_shapes():
    heads, faces = synth.make_heads(2, seed=3, n_faces=4)
    shapes = [h.shape for h in heads]
    assert shapes == [(2, 4, 20, 20), (2, 8, 20, 20), (2, 20, 20, 20), (2, 4, 40, 40), (2, 8, 40, 40), (2, 20, 40, 40),
                      (2, 4, 80, 80), (2, 8, 80, 80), (2, 20, 80, 80)]
    assert sum(h[0].size for h in heads) == 268800           # 1.075 MB of fp32 per image
    np.testing.assert_allclose(heads[0][:, :2] + heads[0][:, 2:], 1.0, atol=1e-6)   # softmax pair
    assert faces.shape == (2, 4, 4)
    d = synth.make_crowd_boxes(2000, seed=1, n_faces=100)
    assert d.shape == (2000, 5) and d.dtype == np.float32
    assert len(np.unique(d[:, 4])) < 2000                     # forced score duplicates
    assert synth.make_frame(36, 64, 0).shape == (36, 64, 3)
    assert synth.make_landmarks(7, 0).shape == (7, 5, 2)


def test_bench_sharding_helpers():
    import bench
    # contiguous image shards [g*n/N, (g+1)*n/N)
    for n, world in [(512, 8), (64, 4), (10, 3)]:
        spans = [bench.shard_range(n, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    line = bench.result_line(frames=128, seconds=0.5, n_gpus=2, steps=4, warmup=3, extra={})
    assert line["value"] == 512.0 and line["unit"] == "frames/s" and line["n_gpus"] == 2 and line["scaling"] == "weak"


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path on the host cores) prints ONE JSON line with the contract's
    keys; under torchrun only rank 0 prints, the other ranks exit 0 without work."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]

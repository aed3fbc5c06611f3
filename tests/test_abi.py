"""The C-ABI library loads without a GPU, exports every symbol include/fd_b200.h declares, and the product path fails
loudly (no CPU fallback) when no device is present."""
import os
import re
import subprocess

import pytest

from conftest import HAS_GPU, ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "fd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(fd_[a-z0-9_]+|_nms|_set_device)\s*\(", src))
    return names


def test_library_builds_and_exports_header():
    from rs_face_detection_b200 import build as B, ffi
    lib_path = B.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    header = _header_symbols()
    assert header, "no symbols parsed from the header"
    assert header <= exported, "declared but not exported: %s" % sorted(header - exported)
    assert set(ffi.SYMBOLS) == header, "ffi.SYMBOLS out of sync with the header: %s" % sorted(set(ffi.SYMBOLS) ^ header)
    lib = ffi.load()
    assert lib.fd_abi_version() == 1


def test_struct_layout_matches_c():
    from rs_face_detection_b200 import ffi
    import ctypes as C
    assert C.sizeof(ffi.FdConfig) == 4 * (2 + 2 + 1 + 8 + 1 + 8 * 4 * 4 + 3 + 3 + 1 + 4 + 1 + 2 + 10)
    assert C.sizeof(ffi.FdFrame) == 24
    cfg = ffi.default_config()
    assert (cfg.image_w, cfg.image_h, cfg.n_strides, cfg.num_anchors) == (640, 640, 3, 2)
    assert abs(cfg.conf_thr - 0.7) < 1e-7 and abs(cfg.iou_thr - 0.45) < 1e-7
    assert list(cfg.strides)[:3] == [32, 16, 8]
    assert (cfg.crop_w, cfg.crop_h) == (112, 112)


@pytest.mark.skipif(HAS_GPU, reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_gpu():
    from rs_face_detection_b200 import Context, FdError, ffi
    with pytest.raises(FdError) as e:
        Context(0)
    assert e.value.code == ffi.FD_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may use oracle/."""
    pkg = os.path.join(ROOT, "rs_face_detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "fd_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt and "libfd_oracle" not in txt, f

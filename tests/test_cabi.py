"""The C-ABI library loads on a GPU-less host and exports every symbol include/b200seg.h declares."""
import ctypes
import os
import re

from b200seg import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200seg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.EXPORTS) == names       # the Python binding covers the whole header


def test_version_and_error_channel():
    assert _cabi.lib.b200seg_version() == 100
    # argument validation happens before any CUDA call: usable without a GPU
    rc = _cabi.lib.b200seg_dwconv3x3(0, 0, 0, 0, _cabi.BF16, 1, 8, 8, 12, 1, 0, 0)
    assert rc < 0
    assert b"multiple of 8" in _cabi.lib.b200seg_last_error()
    rc = _cabi.lib.b200seg_conv_tc(0, 0, 0, 0, 0, 1, 8, 8, 16, 16, 4, 0, 0, 0)
    assert rc < 0 and b"taps" in _cabi.lib.b200seg_last_error()


def test_library_is_sm100a_with_tcgen05_and_tma():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnem in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM"):
        assert mnem in sass, mnem

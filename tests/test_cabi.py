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


def test_streaming_kernels_keep_a_batch_of_loads_in_flight():
    """Round-2 finding (DESIGN 4b ix): ptxas sinks independent global loads into the arithmetic that consumes them when that
    saves registers -- the large-layer BatchNorm passes ran with ONE 16-byte load pair in flight per thread (2.3-2.9 TB/s)
    until the loads were issued as an explicit batch.  This pins the property in the built library: the largest run of
    back-to-back 16-byte global loads (at most 6 other instructions between two of them) per streaming kernel."""
    import importlib.util
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump) or not shutil.which("c++filt"):
        import pytest
        pytest.skip("cuobjdump / c++filt not available")
    tool = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sass_ldg_clusters.py")
    spec = importlib.util.spec_from_file_location("sass_ldg_clusters", tool)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    got = mod.ldg_clusters(_cabi.LIB_PATH, "bn_bwd_reduce|bn_bwd_apply|bn_apply_kernel|upcat_bwd|upsample2x_concat_kernel|dw_wgrad_bf16",
                           wide_only=True)
    want = {
        "bn_bwd_reduce_kernel<__nv_bfloat16, 4, 2>": 8,       # 4 pixel rows x (da, z)
        "bn_bwd_apply_kernel<__nv_bfloat16, true>": 8,
        "bn_apply_kernel<__nv_bfloat16, false, true>": 8,     # 8 rows of z; pinned (the finalize prologue made ptxas split it 4 + 4)
        "upcat_bwd_kernel<__nv_bfloat16>": 12,                # the 4 x 4 gather (16 loads, pinned by a data dependence)
        "upsample2x_concat_kernel<__nv_bfloat16>": 8,         # the 3 x 3 source neighbourhood (9 loads)
        "dw_wgrad_bf16_kernel<1, 4>": 12,                     # 4 dz + 3 x 6 x vectors of a 4-pixel group
        "dw_wgrad_bf16_kernel<2, 4>": 12,
    }
    for key, n in want.items():
        hits = [cl for name, (_, cl) in got.items() if key in name]
        assert hits, f"kernel {key} not found in the library"
        assert hits[0][0] >= n, f"{key}: largest batch of 16-byte loads is {hits[0][0]}, expected >= {n} (clusters {hits[0][:6]})"

"""ctypes loader of tools/_build/libb200seg_tools.so (python tools/build_tools.py): diagnostics and experiments only."""
import ctypes as C
import os

_vp, _i = C.c_void_p, C.c_int
_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libb200seg_tools.so")
if not os.path.exists(_path):
    raise ImportError(f"{_path} missing: run `python tools/build_tools.py`")
lib = C.CDLL(_path)
lib.b200seg_probe_mma.argtypes = [_i, _i, _i, _i, _vp, _vp]
lib.b200seg_mbconv_tc.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]
lib.b200seg_last_error.restype = C.c_char_p


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError(f"tools lib {what} failed (rc={rc}): {lib.b200seg_last_error().decode()}")

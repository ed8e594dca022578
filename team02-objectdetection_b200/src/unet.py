"""Call-site shim: the reference scripts do ``from src.unet import UNet, MobileNetV2UNet``
(main.py:7, inference.py:8, convert.py:8).  Put this directory
(team02-objectdetection_b200/) first on sys.path and those imports resolve to the B200 path."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from b200seg.unet import (LightUNet, MobileNetV2UNet, UNet, double_conv, down, inconv,  # noqa: E402,F401
                          outconv, up)

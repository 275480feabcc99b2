"""``models.conv`` — same import path as /root/reference/models/conv.py, B200-native implementation."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)

import sirgcn_b200  # noqa: E402  (registers the package under an importable name)
from sirgcn_b200.conv import SIRConv, SIREConv, SIRConvBase, SIREConvBase  # noqa: E402,F401

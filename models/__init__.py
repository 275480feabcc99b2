"""Drop-in ``models`` package: the reference's scripts do ``sys.path.append('../..')`` and then
``from models.conv import SIRConv, SIREConv`` / ``from models.utils import MLP, DropEdge`` / ``from models.norm import
GetNorm`` (e.g. /root/reference/benchmark-datasets/zinc/model.py:1-9).  Putting this repository root on ``sys.path``
AHEAD of the reference root (``PYTHONPATH=<this repo>``) makes ``models.conv`` resolve to the B200-native layers,
while every other sub-module (``models.utils``, ``models.norm`` — stock PyTorch/DGL glue outside the hot path) still
resolves to the reference's own file: the package path below is extended with any other ``models`` directory found on
``sys.path``, this repository's first."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
__path__ = [_here]
for _p in list(sys.path):
    _cand = os.path.abspath(os.path.join(_p or os.curdir, "models"))
    if _cand != _here and _cand not in __path__ and os.path.isfile(os.path.join(_cand, "conv.py")):
        __path__.append(_cand)

"""Drop-in ``models`` package: the reference's scripts do ``sys.path.append('../..')`` and then
``from models.conv import SIRConv, SIREConv`` (e.g. /root/reference/benchmark-datasets/zinc/model.py:1-9).
Putting this repository root on ``sys.path`` instead makes them pick up the B200-native layers."""

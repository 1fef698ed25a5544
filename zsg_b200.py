"""Import alias: the package directory is named `zsgnet-pytorch_b200` (not a valid identifier), so
`import zsg_b200` resolves to it."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("zsgnet-pytorch_b200")
sys.modules[__name__] = _pkg

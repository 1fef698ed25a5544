"""Import alias: the package directory is named `zsgnet-pytorch_b200` (not a valid identifier), so
`import zsg_b200` and `import zsg_b200.<sub>` resolve to the very same module objects."""
import importlib
import importlib.abc
import importlib.machinery
import os
import sys

_REAL = "zsgnet-pytorch_b200"
_ALIAS = "zsg_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == _ALIAS or fullname.startswith(_ALIAS + "."):
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(_REAL + spec.name[len(_ALIAS):])

    def exec_module(self, module):
        pass


if not any(isinstance(f, _AliasFinder) for f in sys.meta_path):
    sys.meta_path.insert(0, _AliasFinder())
sys.modules[__name__] = importlib.import_module(_REAL)

"""Importable alias: ``import splade_b200`` -> the package in ``opensearch-neural-pre-train_b200/``
(the directory name required by the project layout is not a valid Python identifier).  Sub-modules are
aliased too, so ``splade_b200.ops`` and the real module are ONE object (no duplicate ctypes types)."""
import importlib
import os
import sys

_REAL = "opensearch-neural-pre-train_b200"
_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg

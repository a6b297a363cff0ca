"""Importable alias of the hyphenated package directory `mi-seg_b200/`: `import mi_seg_b200`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
sys.modules[__name__] = importlib.import_module("mi-seg_b200")

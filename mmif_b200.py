"""Importable alias of the ``multi-modal-image-fusion_b200`` package (its directory name is not
a Python identifier).  ``import mmif_b200; mmif_b200.core.loss.SSIMLoss(...)``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module('multi-modal-image-fusion_b200')
sys.modules[__name__] = _pkg
sys.modules.setdefault('mmif_b200', _pkg)

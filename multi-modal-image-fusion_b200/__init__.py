"""multi-modal-image-fusion_b200 — B200-native fusion loss and metric suite.

``core.loss`` / ``core.metric`` mirror the reference's call surface; ``_lib`` binds the C ABI
(include/mmif_b200.h) of the in-tree ``libmmif_b200.so``.  Import as
``importlib.import_module('multi-modal-image-fusion_b200')`` or through the ``mmif_b200`` alias
module at the repository root."""
from . import _lib  # noqa: F401

__version__ = '0.1.0'

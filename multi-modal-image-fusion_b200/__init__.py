"""multi-modal-image-fusion_b200 — B200-native fusion loss and metric suite.

``core.loss`` / ``core.metric`` mirror the reference's call surface; ``_lib`` binds the C ABI
(include/mmif_b200.h) of the in-tree ``libmmif_b200.so``.  Import as
``importlib.import_module('multi-modal-image-fusion_b200')`` or through the ``mmif_b200`` alias
module at the repository root."""
from . import _lib  # noqa: F401

__version__ = '0.1.0'

import os as _os

if _os.environ.get('MMIF_COUNTS_FILE'):         # integration tests: which kernels did an unmodified reference script launch?
    import atexit as _atexit
    import json as _json

    def _dump_counts(path=_os.environ['MMIF_COUNTS_FILE']):
        try:
            with open(path, 'w') as fh:
                _json.dump(_lib.launch_counts() if _lib._lib is not None else {}, fh)
        except Exception:       # pragma: no cover
            pass
    _atexit.register(_dump_counts)

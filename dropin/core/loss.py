"""Reference-compatible `core.loss` (loss.py:16-19) backed by libmmif_b200.so."""
import mmif_b200  # noqa: F401
from mmif_b200.core.loss import *  # noqa: F401,F403
from mmif_b200.core.loss import __all__  # noqa: F401

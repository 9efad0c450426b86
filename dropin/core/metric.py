"""Reference-compatible `core.metric` (metric.py:16-21) backed by libmmif_b200.so."""
import mmif_b200  # noqa: F401
from mmif_b200.core.metric import *  # noqa: F401,F403
from mmif_b200.core.metric import __all__  # noqa: F401
